"""Host-side helpers on the hot path (mirror of reference graphlearning/utils.py)."""
from __future__ import annotations

import numpy as np


def labels_to_onehot(labels, k=None):
    """One-hot encoding, width max(k, max(label)+1).  Reference graphlearning/utils.py:536-572."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    kk = int(np.max(labels)) + 1
    if k is not None:
        kk = max(kk, int(k))
    labels = labels.astype(int)
    onehot = np.zeros((n, kk))
    onehot[range(n), labels] = 1
    return onehot
