"""graphlearning_b200 - B200-native backend for GraphLearning's data-parallel hot path.

Mirrors the part of the reference API that sits on the path kNN graph build -> Laplacian normalisation ->
Poisson / Laplace iterate (reference graphlearning/__init__.py:1-8 exports the same submodule names):

    import graphlearning_b200 as gl
    W = gl.weightmatrix.knn(X, 10)
    model = gl.ssl.poisson(W, solver='gradient_descent')
    pred = model.fit_predict(train_ind, train_labels)

Host code is Python; all arithmetic on the path runs in hand-written sm_100a CUDA behind the C-ABI of
libglb200.so (include/glb200.h).  There is no CPU fallback: without the library / a GPU the calls raise.
"""
from . import utils        # noqa: F401
from . import trainsets    # noqa: F401
from . import datasets     # noqa: F401
from . import weightmatrix  # noqa: F401
from . import graph as _graph_module
from . import ssl          # noqa: F401
from . import clustering   # noqa: F401
from .graph import graph   # noqa: F401  (reference: `from .graph import graph`, graphlearning/__init__.py:8)

__version__ = "0.1.0"
