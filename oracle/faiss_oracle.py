"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

A faiss-equivalent restatement of the candidate-retrieval hot path of the reference
(/root/reference/Retrieval.py:11-34) and of the index surface BASELINE.json's north_star names.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; newsrecommend_b200/ never does.

PARITY UNPINNED: faiss is a third-party dependency that is absent from /root/reference (no
requirements file, version unpinned) and not installable in this image. The algorithms below
restate the published faiss 1.7.x-1.9.x behaviour (SURVEY.md sections 3.1-3.3, 8b) and are pinned
only by the KATs in tests/test_oracle.py and by fp64 brute force.

The C side (oracle/faiss_oracle.c) holds heaps, MT19937, the centroid update, cluster
splitting and the IVF list scanner; the BLAS sgemm blocks (4096 queries x 1024 items, faiss's
distance_compute_blas_{query,database}_bs) are numpy/OpenBLAS matmuls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfaiss_oracle.so")

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
FLT_MAX = np.float32(3.4028234663852886e38)

BLAS_THRESHOLD = 20   # distance_compute_blas_threshold
BLAS_QUERY_BS = 4096  # distance_compute_blas_query_bs
BLAS_DB_BS = 1024     # distance_compute_blas_database_bs


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "faiss_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfaiss_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int64)
        i64 = C.c_int64
        L.fo_mt19937_nth.restype = C.c_uint32
        L.fo_mt19937_nth.argtypes = [C.c_uint32, C.c_int]
        L.fo_rand_perm.argtypes = [C.POINTER(C.c_int32), i64, i64]
        L.fo_heap_init.argtypes = [C.c_int, i64, i64, fp, ip]
        L.fo_heap_add_block.argtypes = [C.c_int, i64, i64, fp, ip, fp, i64, i64, i64]
        L.fo_heap_add_block_l2.argtypes = [i64, i64, fp, ip, fp, i64, i64, i64, fp, fp]
        L.fo_heap_end.argtypes = [C.c_int, i64, i64, fp, ip]
        L.fo_norms_l2sqr.argtypes = [fp, i64, i64, fp]
        L.fo_search_seq.argtypes = [C.c_int, fp, fp, i64, i64, i64, i64, fp, ip]
        L.fo_truth_topk.argtypes = [C.c_int, fp, fp, i64, i64, i64, i64, C.POINTER(C.c_double), ip]
        L.fo_compute_centroids.argtypes = [i64, i64, i64, fp, ip, fp, fp]
        L.fo_split_clusters.restype = C.c_int
        L.fo_split_clusters.argtypes = [i64, i64, i64, fp, fp]
        L.fo_imbalance_factor.restype = C.c_double
        L.fo_imbalance_factor.argtypes = [i64, i64, ip]
        L.fo_ivf_search.argtypes = [C.c_int, fp, i64, i64, i64, i64, ip, ip, fp, ip, fp, ip]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _as_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    return x


# ---------------------------------------------------------------------------------- random
def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss::rand_perm (utils/random.cpp)."""
    perm = np.empty(n, dtype=np.int32)
    lib().fo_rand_perm(perm.ctypes.data_as(C.POINTER(C.c_int32)), n, seed)
    return perm


def mt19937_nth(seed: int, n: int) -> int:
    return int(lib().fo_mt19937_nth(seed, n))


# ---------------------------------------------------------------------------------- flat kNN
def norms_l2sqr(x: np.ndarray) -> np.ndarray:
    x = _as_f32(x)
    out = np.empty(x.shape[0], dtype=np.float32)
    lib().fo_norms_l2sqr(_f(x), x.shape[0], x.shape[1], _f(out))
    return out


def knn(xq, xb, k: int, metric: int):
    """knn_inner_product / knn_L2sqr (utils/distances.cpp): seq path for nq < 20, else BLAS
    blocks + HeapBlockResultHandler. Returns (D f32[nq,k], I i64[nq,k]) best-first."""
    xq = _as_f32(xq)
    xb = _as_f32(xb)
    nq, d = xq.shape
    nb = xb.shape[0]
    assert xb.shape[1] == d and k > 0
    is_l2 = int(metric == METRIC_L2)
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if nq == 0:
        return D, I
    L = lib()
    if nq < BLAS_THRESHOLD or nb == 0:
        L.fo_search_seq(is_l2, _f(xq), _f(xb), d, nq, nb, k, _f(D), _i(I))
        return D, I
    L.fo_heap_init(is_l2, nq, k, _f(D), _i(I))
    if is_l2:
        xn = norms_l2sqr(xq)
        yn = norms_l2sqr(xb)
    for i0 in range(0, nq, BLAS_QUERY_BS):
        i1 = min(nq, i0 + BLAS_QUERY_BS)
        Db = D[i0:i1]
        Ib = I[i0:i1]
        for j0 in range(0, nb, BLAS_DB_BS):
            j1 = min(nb, j0 + BLAS_DB_BS)
            ip = np.ascontiguousarray(xq[i0:i1] @ xb[j0:j1].T)  # sgemm_
            if is_l2:
                L.fo_heap_add_block_l2(i1 - i0, k, _f(Db), _i(Ib), _f(ip), ip.shape[1], j0, j1,
                                       _f(xn[i0:i1]), _f(yn))
            else:
                L.fo_heap_add_block(0, i1 - i0, k, _f(Db), _i(Ib), _f(ip), ip.shape[1], j0, j1)
    L.fo_heap_end(is_l2, nq, k, _f(D), _i(I))
    return D, I


def knn_fast(xq, xb, k: int, metric: int, db_bs: int = 16384):
    """Same algorithm with a larger database block (fewer Python round trips); used as the CPU
    throughput baseline. Result is identical to knn() up to BLAS blocking effects."""
    global BLAS_DB_BS
    old = BLAS_DB_BS
    BLAS_DB_BS = db_bs
    try:
        return knn(xq, xb, k, metric)
    finally:
        BLAS_DB_BS = old


def truth_topk(xq, xb, k: int, metric: int):
    """fp64 brute force, ordered by (score, id)."""
    xq = _as_f32(xq)
    xb = _as_f32(xb)
    nq, d = xq.shape
    D = np.empty((nq, k), dtype=np.float64)
    I = np.empty((nq, k), dtype=np.int64)
    lib().fo_truth_topk(int(metric == METRIC_L2), _f(xq), _f(xb), d, nq, xb.shape[0], k,
                        D.ctypes.data_as(C.POINTER(C.c_double)), _i(I))
    return D, I


# ---------------------------------------------------------------------------------- indexes
class IndexFlat:
    def __init__(self, d: int, metric: int = METRIC_L2):
        self.d = d
        self.metric_type = metric
        self.is_trained = True
        self.xb = np.empty((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.xb.shape[0]

    def add(self, x):
        x = _as_f32(x)
        assert x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x], axis=0)

    def reset(self):
        self.xb = np.empty((0, self.d), dtype=np.float32)

    def train(self, x):
        pass

    def search(self, x, k):
        x = _as_f32(x)
        assert x.shape[1] == self.d
        assert k > 0
        return knn(x, self.xb, k, self.metric_type)

    def assign(self, x, k=1):
        return self.search(x, k)[1]


class IndexFlatL2(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_L2)


class IndexFlatIP(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexHNSWFlat(IndexFlatL2):
    """Retrieval.py:16 uses HNSW(M=32) as the k-means assigner. Its parallel graph build is
    nondeterministic, so neither faiss nor this oracle can reproduce it bit for bit; the oracle
    (like the product) treats it as an exact L2 assigner. Documented divergence, SURVEY 8a-a3."""

    def __init__(self, d, M=32):
        super().__init__(d)
        self.M = M


class ClusteringParameters:
    def __init__(self):
        self.niter = 25
        self.nredo = 1
        self.verbose = False
        self.spherical = False
        self.int_centroids = False
        self.update_index = False
        self.frozen_centroids = False
        self.min_points_per_centroid = 39
        self.max_points_per_centroid = 256
        self.seed = 1234
        self.decode_block_size = 32768


class ClusteringIterationStats:
    def __init__(self, obj, imbalance_factor, nsplit):
        self.obj = obj
        self.imbalance_factor = imbalance_factor
        self.nsplit = nsplit


def compute_centroids(x, assign, k):
    x = _as_f32(x)
    n, d = x.shape
    assign = np.ascontiguousarray(assign, dtype=np.int64)
    hassign = np.empty(k, dtype=np.float32)
    cent = np.empty((k, d), dtype=np.float32)
    lib().fo_compute_centroids(d, k, n, _f(x), _i(assign), _f(hassign), _f(cent))
    return cent, hassign


def split_clusters(cent, hassign, n):
    k, d = cent.shape
    return int(lib().fo_split_clusters(d, k, n, _f(hassign), _f(cent)))


def imbalance_factor(assign, k):
    assign = np.ascontiguousarray(assign, dtype=np.int64)
    return float(lib().fo_imbalance_factor(assign.shape[0], k, _i(assign)))


class Clustering(ClusteringParameters):
    """Clustering::train (Clustering.cpp), used at Retrieval.py:12-19."""

    def __init__(self, d: int, k: int, cp: ClusteringParameters | None = None):
        super().__init__()
        if cp is not None:
            self.__dict__.update(cp.__dict__)
        self.d = d
        self.k = k
        self.centroids = np.empty(0, dtype=np.float32)
        self.iteration_stats = []
        # hook for teacher-forced parity tests: called as hook(it, centroids_in, assign, centroids_out)
        self.trace = None

    def subsample(self, x):
        n = x.shape[0]
        if n > self.k * self.max_points_per_centroid:
            perm = rand_perm(n, self.seed)
            n2 = self.k * self.max_points_per_centroid
            return np.ascontiguousarray(x[perm[:n2]])
        return x

    def train(self, x, index):
        x = _as_f32(x)
        n, d = x.shape
        k = self.k
        assert d == self.d
        if n < k:
            raise RuntimeError(
                "Number of training points (%d) should be at least as large as number of clusters (%d)" % (n, k))
        if not np.isfinite(x).all():
            raise RuntimeError("input contains NaN's or Inf's")
        x = self.subsample(x)
        nx = x.shape[0]
        if nx < k * self.min_points_per_centroid:
            print("WARNING clustering %d points to %d centroids: please provide at least %d training points"
                  % (nx, k, k * self.min_points_per_centroid), file=sys.stderr)
        if nx == k:
            self.centroids = x.reshape(-1).copy()
            index.reset()
            index.add(x)
            return
        best_obj = None
        best = None
        for redo in range(self.nredo):
            perm = rand_perm(nx, self.seed + 1 + redo * 15486557)
            cent = np.ascontiguousarray(x[perm[:k]])
            if self.spherical:
                normalize_L2(cent)
            if index.ntotal != 0:
                index.reset()
            index.add(cent)
            stats = []
            for it in range(self.niter):
                Dd, assign = index.search(x, 1)
                assign = assign.reshape(-1)
                obj = float(np.sum(Dd.reshape(-1).astype(np.float32), dtype=np.float32))
                cent_in = cent
                cent, hassign = compute_centroids(x, assign, k)
                nsplit = split_clusters(cent, hassign, nx)
                if self.spherical:
                    normalize_L2(cent)
                stats.append(ClusteringIterationStats(obj, imbalance_factor(assign, k), nsplit))
                if self.verbose:
                    print("  Iteration %d objective=%g imbalance=%.3f nsplit=%d" %
                          (it, obj, stats[-1].imbalance_factor, nsplit))
                if self.trace is not None:
                    self.trace(it, cent_in, assign, cent)
                index.reset()
                index.add(cent)
            if self.nredo > 1:
                better = best_obj is None or (obj > best_obj if index.metric_type == METRIC_INNER_PRODUCT else obj < best_obj)
                if better:
                    best_obj, best = obj, (cent.copy(), stats)
            else:
                best = (cent, stats)
        cent, stats = best
        if self.nredo > 1:
            index.reset()
            index.add(cent)
        self.centroids = cent.reshape(-1).copy()
        self.iteration_stats = stats


def vector_float_to_array(v):
    return np.array(v, dtype=np.float32, copy=True)


def normalize_L2(x):
    """fvec_renorm_L2: in place; zero rows untouched."""
    for i in range(x.shape[0]):
        nr = np.float32(np.sqrt(np.dot(x[i], x[i])))
        if nr > 0:
            x[i] *= np.float32(1.0) / nr


class IndexIVFFlat:
    """IndexIVFFlat (Level1Quantizer::train_q1 with cp.niter = 10, add_core into insertion-
    ordered ArrayInvertedLists, search_preassigned + IVFFlatScanner)."""

    def __init__(self, quantizer, d, nlist, metric=METRIC_L2):
        self.quantizer = quantizer
        self.d = d
        self.nlist = nlist
        self.metric_type = metric
        self.nprobe = 1
        self.is_trained = False
        self.cp = ClusteringParameters()
        self.cp.niter = 10
        self._lists = [[] for _ in range(nlist)]  # list of (ids array, rows array) chunks
        self.ntotal = 0
        self._csr = None

    def train(self, x):
        x = _as_f32(x)
        if self.quantizer.is_trained and self.quantizer.ntotal == self.nlist:
            self.is_trained = True
            return
        clus = Clustering(self.d, self.nlist, self.cp)
        self.quantizer.reset()
        clus.train(x, self.quantizer)
        self.is_trained = True
        self.clustering = clus

    def add(self, x):
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed")
        x = _as_f32(x)
        assign = self.quantizer.assign(x, 1).reshape(-1)
        ids = np.arange(self.ntotal, self.ntotal + x.shape[0], dtype=np.int64)
        order = np.argsort(assign, kind="stable")
        sa = assign[order]
        bounds = np.searchsorted(sa, np.arange(self.nlist + 1))
        for l in range(self.nlist):
            sel = order[bounds[l]:bounds[l + 1]]
            if sel.size:
                self._lists[l].append((ids[sel], x[sel]))
        self.ntotal += x.shape[0]
        self._csr = None

    def reset(self):
        self._lists = [[] for _ in range(self.nlist)]
        self.ntotal = 0
        self._csr = None

    def list_sizes(self):
        return np.array([sum(c[0].shape[0] for c in l) for l in self._lists], dtype=np.int64)

    def _build_csr(self):
        if self._csr is None:
            sizes = self.list_sizes()
            off = np.zeros(self.nlist + 1, dtype=np.int64)
            np.cumsum(sizes, out=off[1:])
            ids = np.empty(self.ntotal, dtype=np.int64)
            rows = np.empty((self.ntotal, self.d), dtype=np.float32)
            for l in range(self.nlist):
                p = off[l]
                for cid, cx in self._lists[l]:
                    ids[p:p + cid.shape[0]] = cid
                    rows[p:p + cid.shape[0]] = cx
                    p += cid.shape[0]
            self._csr = (off, ids, rows)
        return self._csr

    def search(self, x, k):
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed")
        x = _as_f32(x)
        assert x.shape[1] == self.d and k > 0
        nprobe = min(self.nprobe, self.nlist)
        _, coarse = self.quantizer.search(x, nprobe)
        coarse = np.ascontiguousarray(coarse, dtype=np.int64)
        off, ids, rows = self._build_csr()
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        lib().fo_ivf_search(int(self.metric_type == METRIC_L2), _f(x), nq, self.d, k, nprobe,
                            _i(coarse), _i(off), _f(rows), _i(ids), _f(D), _i(I))
        return D, I
