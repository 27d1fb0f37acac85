"""`import faiss` -> the CPU oracle (test infrastructure only): lets the unmodified reference
script run a second time on the oracle so that its output can be compared with the B200 run."""
from oracle.faiss_oracle import (METRIC_INNER_PRODUCT, METRIC_L2, Clustering, ClusteringParameters,  # noqa: F401
                                 IndexFlat, IndexFlatIP, IndexFlatL2, IndexHNSWFlat, IndexIVFFlat,
                                 normalize_L2, vector_float_to_array)
