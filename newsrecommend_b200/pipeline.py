"""The retrieval stage and its two neighbours, batched.

`retrieve_candidates` is /root/reference/Retrieval.py:11-34 as ONE call: k-means over the
article embeddings, nearest centroid of every article, inverted lists, nearest centroid per
user, and each user's whole list as its candidates -- same outputs as the script (ids in
ascending row order inside a list), without the 50,000-iteration Python loop and the 300
boolean masks. `finalize_candidates` is finialize_retrieval.py:6-15, `hit_rate` is
utils.py:12-22. File formats are those the scripts exchange under `news/`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import faiss as nf
from ._lib import check, lib


def _dev(a, dtype):
    t = torch.as_tensor(a)
    return t.to(device=nf._device(), dtype=dtype).contiguous()


# ---------------------------------------------------------------------------- file formats
def load_article_table(path):
    """news/article_table.npy (embedding_generate.py:124-131): rows = embedding values followed
    by the article id; written with dtype=object upstream, so allow_pickle is needed (the
    reference's own np.load at Retrieval.py:6 only works on a float re-save). Returns
    (article_ids i64[n], embeddings f32[n, d] C-contiguous) as Retrieval.py:7-8 builds them."""
    arr = np.load(path, allow_pickle=True)
    if arr.dtype == object:
        arr = np.asarray(arr.tolist(), dtype=np.float64)
    ids = arr[:, -1].astype(np.int64)
    emb = np.ascontiguousarray(arr[:, :-1], dtype=np.float32)
    return ids, emb


def load_article_table_device(path, chunk_rows: int = 32768):
    """news/article_table.npy straight into HBM: (article_ids i64[n], embeddings f32[n, d]) as CUDA
    tensors, equal to load_article_table() bit for bit.

    A float table (what Retrieval.py:6 can actually read: its np.load has no allow_pickle) never
    becomes a numpy array: the .npy header is parsed, the float64 payload is read in chunks into two
    page-locked staging buffers (file.readinto), each chunk goes host-to-device on a copy stream while
    the next one is being read, and nrb_split_table_f64 does the casts of Retrieval.py:7-8 on the
    device. An object-dtype table (what embedding_generate.py:124-131 writes) is pickled Python
    floats; it has to go through the unpickler once, then takes the same device path."""
    dev = nf._device()
    with open(path, "rb") as f:
        major, _ = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0(f) if major == 1
                                 else np.lib.format.read_array_header_2_0(f))
        plain = not dtype.hasobject and dtype == np.float64 and not fortran and len(shape) == 2
        if plain:
            n, w = shape
            table = torch.empty((n, w), dtype=torch.float64, device=dev)
            stage = [torch.empty((chunk_rows, w), dtype=torch.float64, pin_memory=True) for _ in range(2)]
            done = [torch.cuda.Event(), torch.cuda.Event()]
            h2d, _ = nf._copy_streams()
            h2d.wait_stream(torch.cuda.current_stream())
            for c, r0 in enumerate(range(0, n, chunk_rows)):
                r1 = min(n, r0 + chunk_rows)
                buf = stage[c & 1]
                if c >= 2:
                    done[c & 1].synchronize()  # the copy that last used this staging buffer has finished
                got = f.readinto(memoryview(buf.numpy()).cast("B")[: (r1 - r0) * w * 8])
                if got != (r1 - r0) * w * 8:
                    raise IOError("article table is truncated")
                with torch.cuda.stream(h2d):
                    table[r0:r1].copy_(buf[: r1 - r0], non_blocking=True)
                    done[c & 1].record(h2d)
            torch.cuda.current_stream().wait_stream(h2d)
    if not plain:
        arr = np.load(path, allow_pickle=True)
        if arr.dtype == object:
            arr = np.asarray(arr.tolist(), dtype=np.float64)
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.ndim == 2
        n, w = arr.shape
        table = torch.from_numpy(arr).to(dev)
    ids = torch.empty(n, dtype=torch.int64, device=dev)
    emb = torch.empty((n, w - 1), dtype=torch.float32, device=dev)
    check(lib.nrb_split_table_f64(table.data_ptr(), n, w, emb.data_ptr(), ids.data_ptr(), nf._stream()), "split_table_f64")
    return ids, emb


def load_user_profiles(path):
    """news/test_user_profile.npy: pickled dict uid -> vector (Retrieval.py:28). Returns
    (uids i64[nu], profiles f32[nu, d]) in the dict's iteration order."""
    d = np.load(path, allow_pickle=True).item()
    uids = np.fromiter(d.keys(), dtype=np.int64, count=len(d))
    prof = np.stack([np.asarray(v, dtype=np.float32).reshape(-1) for v in d.values()])
    return uids, np.ascontiguousarray(prof)


def save_recommendations(path, uids, offsets, candidates):
    """news/test_user_recommendations.npy as Retrieval.py:36 writes it: dict uid -> i64 array."""
    offsets = np.asarray(offsets)
    candidates = np.asarray(candidates)
    rec = {int(u): candidates[offsets[i]:offsets[i + 1]].copy() for i, u in enumerate(np.asarray(uids))}
    np.save(path, rec, allow_pickle=True)


def load_recommendations(path):
    """Inverse of save_recommendations: (uids, offsets i64[nu+1], candidates i64[total])."""
    d = np.load(path, allow_pickle=True).item()
    uids = np.fromiter(d.keys(), dtype=np.int64, count=len(d))
    lens = np.fromiter((len(v) for v in d.values()), dtype=np.int64, count=len(d))
    off = np.zeros(len(d) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cand = np.concatenate([np.asarray(v, dtype=np.int64) for v in d.values()]) if len(d) else np.empty(0, np.int64)
    return uids, off, cand


# ---------------------------------------------------------------------------- the stage
def retrieve_candidates(article_ids, embeddings, user_profiles, num_clusters=300, niter=80, verbose=False,
                        centroids=None):
    """Retrieval.py:11-34 batched. Returns a dict with
      offsets i64[nu+1], candidates i64[total]  -- CSR: user u's candidates (article ids)
      user_list i64[nu], centroids f32[k, d], assignments i64[n], list_sizes i64[k].
    centroids (optional, f32[k, d]): skip the k-means of lines 12-19 and use these (teacher-forced
    comparison with another implementation of the stage)."""
    emb, _ = nf._to_device_f32(embeddings)
    users, _ = nf._to_device_f32(user_profiles)
    n, d = emb.shape
    ids = _dev(article_ids, torch.int64)
    index = nf.IndexHNSWFlat(d, 32)                           # :16 (exact L2)
    if centroids is None:
        clustering = nf.Clustering(d, num_clusters)          # :12
        clustering.niter = niter                              # :13
        clustering.verbose = verbose                          # :14
        clustering.train(emb, index)                          # :18
        centroids = nf.vector_float_to_array(clustering.centroids).reshape(num_clusters, d)  # :19
    else:
        centroids = np.ascontiguousarray(centroids, dtype=np.float32).reshape(num_clusters, d)
        index.add(centroids)
    _, assign = index.search(emb, 1)                          # :21
    assign = assign.reshape(-1)                               # :22
    # :23 -- the 300 boolean masks become one stable counting sort
    dev = emb.device
    off = torch.empty(num_clusters + 1, dtype=torch.int32, device=dev)
    order = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    wsb = lib.nrb_ivf_build_lists_workspace(n, num_clusters)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = nf._stream()
    check(lib.nrb_ivf_build_lists(assign.data_ptr(), n, num_clusters, off.data_ptr(), order.data_ptr(),
                                  ws.data_ptr(), wsb, st), "ivf_build_lists")
    list_ids = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    check(lib.nrb_gather_i64(ids.data_ptr(), order.data_ptr(), n, list_ids.data_ptr(), st), "gather_i64")
    centroid_index = nf.IndexFlatL2(d)                        # :25
    centroid_index.add(centroids)                             # :26
    _, I = centroid_index.search(users, 1)                    # :30-32, one batched call
    user_list = I.reshape(-1)
    sizes = (off[1:] - off[:-1]).to(torch.int64)
    lens = torch.where(user_list >= 0, sizes[user_list.clamp(min=0)], torch.zeros_like(user_list))
    out_off = torch.zeros(users.shape[0] + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=out_off[1:])
    total = int(out_off[-1])
    cand = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    check(lib.nrb_expand_lists(user_list.data_ptr(), off.data_ptr(), list_ids.data_ptr(), out_off.data_ptr(),
                               users.shape[0], cand.data_ptr(), st), "expand_lists")   # :33-34
    return dict(offsets=out_off, candidates=cand[:total], user_list=user_list, centroids=centroids,
                assignments=assign, list_sizes=sizes)


def csr_contains(offsets, candidates, targets):
    """u8[nu]: is targets[u] among user u's candidates (utils.py:14-16)?"""
    off = _dev(offsets, torch.int64)
    cand = _dev(candidates, torch.int64)
    tgt = _dev(targets, torch.int64)
    out = torch.empty(tgt.shape[0], dtype=torch.uint8, device=off.device)
    check(lib.nrb_csr_contains(off.data_ptr(), cand.data_ptr(), tgt.data_ptr(), tgt.shape[0], out.data_ptr(),
                               nf._stream()), "csr_contains")
    return out


def hit_rate(offsets, candidates, ground_truth):
    """utils.py:12-22: (users whose ground-truth article is among their candidates,
    {list length: number of users})."""
    hits = int(csr_contains(offsets, candidates, ground_truth).sum())
    off = torch.as_tensor(offsets)
    lens = (off[1:] - off[:-1]).cpu().numpy()
    vals, counts = np.unique(lens, return_counts=True)
    return hits, dict(zip(vals.tolist(), counts.tolist()))


def finalize_candidates(offsets, candidates, ground_truth, cap=None, seed=0):
    """finialize_retrieval.py:6-15: append the ground-truth article to a user's list when it is
    missing. The script's `np.random.choice(...)` result at :8 is discarded, i.e. its cap of 400
    is a no-op; cap=None reproduces that. cap=N implements the evidently intended truncation
    (a behaviour change, off by default): lists longer than N keep N entries sampled without
    replacement. Returns (offsets, candidates) CSR on the device."""
    off = _dev(offsets, torch.int64)
    cand = _dev(candidates, torch.int64)
    gt = _dev(ground_truth, torch.int64)
    nu = gt.shape[0]
    if cap is not None:
        g = torch.Generator(device=off.device).manual_seed(seed)
        lens = off[1:] - off[:-1]
        row = torch.repeat_interleave(torch.arange(nu, device=off.device), lens)
        # random order inside each row: shuffle everything, then a STABLE sort by row (the row must
        # not be folded into a float32 key -- at row ~50,000 the spacing of float32 is 2^-8 and
        # neighbouring rows' keys collide)
        perm1 = torch.argsort(torch.rand(cand.shape[0], generator=g, device=off.device))
        perm = perm1[torch.sort(row[perm1], stable=True)[1]]
        rank = torch.arange(cand.shape[0], device=off.device) - off[:-1][row]
        keep = rank < cap
        cand = cand[perm][keep]
        lens = torch.minimum(lens, torch.full_like(lens, cap))
        off = torch.zeros(nu + 1, dtype=torch.int64, device=off.device)
        torch.cumsum(lens, 0, out=off[1:])
    missing = csr_contains(off, cand, gt) == 0
    lens = off[1:] - off[:-1]
    new_lens = lens + missing.to(torch.int64)
    new_off = torch.zeros(nu + 1, dtype=torch.int64, device=off.device)
    torch.cumsum(new_lens, 0, out=new_off[1:])
    out = torch.empty(int(new_off[-1]), dtype=torch.int64, device=off.device)
    row = torch.repeat_interleave(torch.arange(nu, device=off.device), lens)
    pos = torch.arange(cand.shape[0], device=off.device) - off[:-1][row] + new_off[:-1][row]
    out[pos] = cand
    out[new_off[1:][missing] - 1] = gt[missing]
    return new_off, out
