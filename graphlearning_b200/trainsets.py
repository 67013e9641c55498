"""Training-set generation (mirror of reference graphlearning/trainsets.py:47-156, the in-memory part)."""
from __future__ import annotations

import numpy as np


def generate(labels, rate=1, num_trials=1, mask=None, seed=None):
    """Random stratified label sets.  Same draws as the reference for the same numpy seed
    (np.random.choice per class, in np.unique order; trainsets.py:86-131)."""
    if seed is not None:
        np.random.seed(seed)
    labels = np.asarray(labels)
    unique_labels = np.unique(labels)
    num_per_class = np.bincount(labels)
    num_classes = len(unique_labels)
    num_points = len(labels)
    if type(rate) == int:
        rate = (np.ones(num_classes)[None, :] * rate).astype(int)
    elif type(rate) == float:
        rate = (rate * num_per_class[None, :]).astype(int)
    elif type(rate) == np.ndarray:
        ratetype = rate.dtype
        if rate.ndim != 2:
            raise ValueError("Must provide a 2-dimensional array for rate")
        if rate.shape[1] == 1:
            rate = rate @ np.ones((1, num_classes))
        if np.issubdtype(ratetype, np.integer):
            rate = rate.astype(int)
        elif np.issubdtype(ratetype, np.floating):
            rate = (rate * num_per_class).astype(int)
        else:
            raise ValueError("Invalid numpy array type " + str(rate.dtype))
    else:
        raise ValueError("Invalid rate type " + str(type(rate)))
    if mask is None:
        mask = np.ones(num_points, dtype=bool)
    trainset = []
    for _ in range(num_trials):
        for i in range(rate.shape[0]):
            L = []
            for j, l in enumerate(unique_labels):
                p = ((labels == l) & mask).astype(float)
                p = p / np.sum(p)
                L = L + np.random.choice(num_points, size=rate[i, j], p=p, replace=False).tolist()
            trainset.append(np.array(L))
    if len(trainset) == 1:
        trainset = trainset[0]
    return trainset
