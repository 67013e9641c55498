"""Drop-in name for the reference package: `import graphlearning as gl` resolves to the B200 backend
(graphlearning_b200) for the hot-path API - gl.weightmatrix, gl.graph, gl.ssl, gl.utils, gl.trainsets, gl.clustering
(reference graphlearning/__init__.py:1-8 exports the same submodule names).  Everything outside the hot path
(datasets download, active learning, plotting) is not provided here."""
import sys as _sys

import graphlearning_b200 as _b

from graphlearning_b200 import clustering, datasets, ssl, trainsets, utils, weightmatrix  # noqa: F401
from graphlearning_b200 import graph as _graph_module
from graphlearning_b200.graph import graph  # noqa: F401

for _name in ("clustering", "datasets", "ssl", "trainsets", "utils", "weightmatrix"):
    _sys.modules[__name__ + "." + _name] = getattr(_b, _name)
_sys.modules[__name__ + ".graph"] = _graph_module          # `import graphlearning.graph` resolves to the module, gl.graph to the class
__version__ = _b.__version__
