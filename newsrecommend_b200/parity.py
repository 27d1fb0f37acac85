"""Gap-aware comparison of top-k results (the parity rule of BASELINE.json's north_star):

* ids must match position-wise, except inside runs of adjacent reference scores whose gaps are
  below `gap_rtol` (1e-5) relative -- inside such a run the ids are compared as a set; the run
  that touches position k-1 may also exchange members with items ranked beyond k, which is
  accepted when the candidate's own score is within the gap of the reference's k-th score;
* scores must agree within `score_rtol` (1e-4) relative. For L2 the tolerance is relative to the
  largest distance in the row as well (or to `scale`, e.g. |q|^2 + max|x|^2, when given), because
  faiss's and our squared distance is a difference of O(|x|^2) terms (a self-match has true
  distance 0, and in low dimensions every near neighbour is a cancellation).

Pure numpy; no oracle import -- callers pass the reference (D, I).
"""
from __future__ import annotations

import numpy as np

GAP_RTOL = 1e-5
SCORE_RTOL = 1e-4


def compare_topk(D, I, D_ref, I_ref, metric: int = 0, gap_rtol: float = GAP_RTOL,
                 score_rtol: float = SCORE_RTOL, scale=None) -> dict:
    D = np.asarray(D, dtype=np.float64)
    D_ref = np.asarray(D_ref, dtype=np.float64)
    I = np.asarray(I)
    I_ref = np.asarray(I_ref)
    assert D.shape == D_ref.shape == I.shape == I_ref.shape, (D.shape, D_ref.shape, I.shape, I_ref.shape)
    nq, k = D.shape
    valid = I_ref >= 0
    # per-row scale for tolerances
    absd = np.where(valid, np.abs(D_ref), 0.0)
    row_scale = absd.max(axis=1, keepdims=True) if k else np.zeros((nq, 1))
    if scale is not None:
        row_scale = np.maximum(row_scale, scale)
    tol = score_rtol * np.maximum(np.abs(D_ref), row_scale if metric == 1 else 0.0) + 1e-30
    gap_tol = gap_rtol * np.maximum(np.abs(D_ref), row_scale if metric == 1 else 1e-30)

    pad_ok = np.all((I < 0) == (~valid))
    score_err = np.where(valid & (I >= 0), np.abs(D - D_ref), 0.0)
    rel = np.where(valid, score_err / np.maximum(np.abs(D_ref), np.maximum(row_scale if metric == 1 else 0, 1e-30)), 0.0)

    exact = (I == I_ref)
    hard_bad = 0
    tie_exempt = 0
    bad_queries = []
    rows = np.nonzero(~exact.all(axis=1))[0]
    for q in rows:
        ok = True
        # tie runs on the reference scores
        j = 0
        while j < k:
            e = j
            while e + 1 < k and valid[q, e + 1] and abs(D_ref[q, e + 1] - D_ref[q, e]) <= max(gap_tol[q, e], gap_tol[q, e + 1]):
                e += 1
            if not exact[q, j:e + 1].all():
                ref_set = set(I_ref[q, j:e + 1].tolist())
                got = I[q, j:e + 1].tolist()
                extra = [g for g in got if g not in ref_set]
                if len(set(got)) != len(got):
                    ok = False
                elif extra:
                    if e == k - 1:
                        # boundary run: outsiders must score within the gap of the reference k-th
                        kth = D_ref[q, k - 1]
                        for pos in range(j, e + 1):
                            if I[q, pos] in ref_set:
                                continue
                            if I[q, pos] < 0 or abs(D[q, pos] - kth) > gap_tol[q, k - 1] * (e - j + 2) + tol[q, k - 1]:
                                ok = False
                    else:
                        ok = False
            j = e + 1
        if ok:
            tie_exempt += 1
        else:
            hard_bad += 1
            if len(bad_queries) < 8:
                bad_queries.append(int(q))
    # scores compared after sorting ids inside tie runs is overkill: position-wise scores must
    # already agree because runs have near-equal scores.
    score_bad = int((score_err > tol * (1 + 0)).sum() - 0)
    # allow run-internal permutations: their score differences are below gap_tol << tol
    inter = 0
    for q in range(nq):
        inter += len(set(I[q][I[q] >= 0].tolist()) & set(I_ref[q][I_ref[q] >= 0].tolist()))
    denom = int(valid.sum())
    rep = dict(
        ok=bool(hard_bad == 0 and score_bad == 0 and pad_ok),
        n_queries=int(nq), k=int(k),
        exact_ordered=float(exact.all(axis=1).mean()) if nq else 1.0,
        id_mismatch_queries=int(hard_bad), tie_exempt_queries=int(tie_exempt),
        score_violations=score_bad, max_rel_score_err=float(rel.max()) if rel.size else 0.0,
        recall=float(inter / denom) if denom else 1.0, pad_ok=bool(pad_ok), bad_queries=bad_queries,
    )
    return rep


def recall_at_k(I, I_ref) -> float:
    I = np.asarray(I)
    I_ref = np.asarray(I_ref)
    inter = 0
    tot = 0
    for a, b in zip(I, I_ref):
        bs = set(b[b >= 0].tolist())
        inter += len(set(a[a >= 0].tolist()) & bs)
        tot += len(bs)
    return inter / tot if tot else 1.0
