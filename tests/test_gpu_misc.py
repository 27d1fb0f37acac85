"""GPU tests: k-way merge (K4). (The unmodified reference scripts run in
tests/test_gpu_reference_script.py.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("metric", [0, 1])
def test_merge_topk(nf, metric):
    import torch
    from newsrecommend_b200._lib import check, lib
    rng = np.random.default_rng(0)
    G, nq, k = 8, 300, 50
    Dp = rng.standard_normal((G, nq, k)).astype(np.float32)
    Dp = np.sort(Dp, axis=2)
    if metric == 0:
        Dp = Dp[:, :, ::-1].copy()
    else:
        Dp = np.abs(Dp)
        Dp = np.sort(Dp, axis=2)
    Ip = rng.permutation(G * nq * k).reshape(G, nq, k).astype(np.int64)
    Ip[3, :, 40:] = -1  # a shard with fewer than k results
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dp, ipt = torch.from_numpy(Dp).cuda(), torch.from_numpy(Ip).cuda()
    check(lib.nrb_merge_topk(dp.data_ptr(), ipt.data_ptr(), G, nq, k, metric, D.data_ptr(), I.data_ptr(), None))
    D, I = D.cpu().numpy(), I.cpu().numpy()
    for q in range(nq):
        d = Dp[:, q, :].reshape(-1)
        i = Ip[:, q, :].reshape(-1)
        keep = i >= 0
        d, i = d[keep], i[keep]
        order = np.argsort(-d if metric == 0 else d, kind="stable")[:k]
        assert np.array_equal(D[q], d[order]) and np.array_equal(I[q], i[order])
