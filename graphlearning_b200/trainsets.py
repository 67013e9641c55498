"""Training sets (mirror of reference graphlearning/trainsets.py:17-156).

`generate` draws exactly the reference's random numbers: for every requested rate row and every class, in np.unique
order, one np.random.choice over all points with the class indicator as probability vector - so a numpy seed gives
the same label sets as the reference.  `load` reads the reference's saved permutation files; nothing is downloaded."""
from __future__ import annotations

import os

import numpy as np

from . import utils

trainset_dir = os.path.abspath(os.path.join(os.getcwd(), "trainsets"))     # reference trainsets.py:15


def _file_for(dataset, trainset_name):
    return dataset.lower() + trainset_name.lower() + "_permutations.npz"


def load(dataset, trainset_name=""):
    """The 'perm' array of trainsets/<dataset><trainset_name>_permutations.npz (reference trainsets.py:17-45 without
    the download branch :39-41; LabelPermutations/ of the reference's repository holds the files)."""
    name = _file_for(dataset, trainset_name)
    path = os.path.join(trainset_dir, name)
    if not os.path.exists(path) and os.path.isdir(trainset_dir):
        hits = [f for f in os.listdir(trainset_dir) if f.lower() == name.lower()]
        if hits:
            path = os.path.join(trainset_dir, hits[0])
    if not os.path.exists(path):
        raise FileNotFoundError("%s not found in %s (this backend does not download; copy LabelPermutations/%s of the "
                                "GraphLearning repository there)" % (name, trainset_dir, name))
    return utils.numpy_load(path, "perm")


def _rate_table(rate, classes_present, per_class):
    """rows = label rates to generate, columns = labelled points per class (reference trainsets.py:88-113)."""
    width = len(classes_present)
    if type(rate) == int:
        return np.full((1, width), rate, dtype=int)
    if type(rate) == float:
        return (rate * per_class[None, :]).astype(int)
    if type(rate) == np.ndarray:
        if rate.ndim != 2:
            raise ValueError("Must provide a 2-dimensional array for rate")
        kind = rate.dtype
        table = rate @ np.ones((1, width)) if rate.shape[1] == 1 else rate
        if np.issubdtype(kind, np.integer):
            return table.astype(int)
        if np.issubdtype(kind, np.floating):
            return (table * per_class).astype(int)
        raise ValueError("Invalid numpy array type " + str(kind))
    raise ValueError("Invalid rate type " + str(type(rate)))


def generate(labels, rate=1, num_trials=1, mask=None, dataset=None, trainset_name="", overwrite=False, seed=None):
    """Random stratified label sets; same arguments, same random stream and same return value (a single index array, or
    a list of them) as the reference (trainsets.py:47-156).  With `dataset` the sets are also saved the way the
    reference saves them."""
    if seed is not None:
        np.random.seed(seed)
    labels = np.asarray(labels)
    classes_present = np.unique(labels)
    table = _rate_table(rate, classes_present, np.bincount(labels))
    n = len(labels)
    eligible = np.ones(n, dtype=bool) if mask is None else mask
    drawn = []
    for _ in range(num_trials):
        for counts in table:
            picks = []
            for cls, how_many in zip(classes_present, counts):
                weight = ((labels == cls) & eligible).astype(float)
                picks.extend(np.random.choice(n, size=how_many, p=weight / np.sum(weight), replace=False).tolist())
            drawn.append(np.array(picks))
    result = drawn[0] if len(drawn) == 1 else drawn
    if dataset is not None:
        result = np.array(result, dtype=object)
        os.makedirs(trainset_dir, exist_ok=True)
        path = os.path.join(trainset_dir, _file_for(dataset, trainset_name))
        if os.path.isfile(path) and not overwrite:
            print("Training set file " + path + " already exists. Not saving.")
        else:
            np.savez_compressed(path, perm=result)
    return result
