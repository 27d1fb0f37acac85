"""newsrecommend_b200 -- B200-native candidate retrieval (the Retrieval.py stage of
YuxuanZhao/NewsRecommend): exact top-k, IVF-Flat and IVF k-means over article / user
embeddings, behind a faiss-compatible surface (newsrecommend_b200.faiss).

Importing the package loads libnrb200.so (hand-written sm_100a CUDA behind the C-ABI of
include/nrb200.h) and fails loudly if it has not been built; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when libnrb200.so is missing)
from . import faiss  # noqa: F401
from .faiss import (METRIC_INNER_PRODUCT, METRIC_L2, Clustering, ClusteringParameters, IndexFlat,  # noqa: F401
                    IndexFlatIP, IndexFlatL2, IndexHNSWFlat, IndexIVFFlat, normalize_L2,
                    vector_float_to_array)

__version__ = "0.1.0"
