"""Labels and feature files of the datasets the examples use (mirror of reference graphlearning/datasets.py:73-152, the
file-reading part).  The reference downloads missing files from github.com / umn.edu; this backend never opens a
connection: a missing file is an error that says where to put it.  Files are the reference's own .npz files
(`Data/*_labels.npz` of its repository -> ./data/<dataset>_labels.npz)."""
from __future__ import annotations

import os

import numpy as np

from . import utils

data_dir = os.path.abspath(os.path.join(os.getcwd(), "data"))      # reference datasets.py:25


def _find(directory, name):
    """The reference's repository spells its files 'MNIST_labels.npz', its loader 'mnist_labels.npz': accept either."""
    path = os.path.join(directory, name)
    if os.path.exists(path):
        return path
    if os.path.isdir(directory):
        for f in os.listdir(directory):
            if f.lower() == name.lower():
                return os.path.join(directory, f)
    return None


def load(dataset, metric="raw", labels_only=False):
    """labels, or (data, labels): reference datasets.py:73-152 without the download branch (:139-141, :149-151)."""
    labels_file = dataset.lower() + "_labels.npz"
    data_file = dataset.lower() + "_" + metric.lower() + ".npz"
    path = _find(data_dir, labels_file)
    if path is None:
        raise FileNotFoundError("%s not found in %s (this backend does not download; copy Data/%s of the GraphLearning "
                                "repository there)" % (labels_file, data_dir, labels_file))
    labels = utils.numpy_load(path, "labels")
    if labels_only:
        return labels
    path = _find(data_dir, data_file)
    if path is None:
        raise FileNotFoundError("%s not found in %s (this backend does not download)" % (data_file, data_dir))
    return utils.numpy_load(path, "data"), labels
