"""Oracle for scoring, top-k and retrieval metrics.  TEST INFRASTRUCTURE.

Follows, in the reference (/root/reference):
  * one-vs-all hyperbolic scoring   src/train.py:3259       (``pmath.dist(q[1,D], G[P,D])``)
  * top-k semantics                 src/auxiliary.py:374    (``torch.topk(largest=False)``)
                                    notebooks/retrieval.ipynb:202 (``argsort(sim)[-k:][::-1]``)
  * cosine scoring                  notebooks/retrieval.ipynb:368 (sklearn ``cosine_similarity``)
  * full ranking                    notebooks/retrieval.ipynb:383 (``np.argsort(sim)[::-1]``)
  * MRR@k / Precision@k             notebooks/retrieval.ipynb:310-324
  * AP (ranking order)              notebooks/retrieval.ipynb:411-420
  * nDCG                            notebooks/retrieval.ipynb:430-437
  * Recall@k                        notebooks/retrieval.ipynb:439-443
  * sklearn AP (ties grouped)       src/train.py:3285, src/auxiliary.py:200-224
  * evaluate_retrieval              src/train.py:3108-3296

The hyperbolic arithmetic is PARITY UNPINNED (oracle/pmath.py).  The cosine and
metric functions are pinned by golden vectors produced by the reference's own
code/dependencies (tests/golden/make_golden.py).

Tie policy (the reference leaves it unspecified -- numpy introsort reversed,
``torch.topk``): ascending distance / descending similarity, equal scores ->
lower gallery index first.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import pmath


# --------------------------------------------------------------------------- scoring
def hyperbolic_dist_rows(q: torch.Tensor, g: torch.Tensor, c: float, form: str = "geoopt",
                         block: int = 64) -> torch.Tensor:
    """[Q,D] x [N,D] -> [Q,N] distances.  ``form='geoopt'`` is the reference path
    (per-query ``pmath.dist(q[1,D], G)`` exactly as src/train.py:3259, looped);
    ``form='arcosh'`` is the closed form, evaluated in blocks."""
    k = torch.tensor(-float(c), dtype=q.dtype)
    out = torch.empty(q.shape[0], g.shape[0], dtype=q.dtype)
    if form == "geoopt":
        for i in range(q.shape[0]):
            out[i] = pmath.dist(q[i].unsqueeze(0), g, k=k)
    elif form == "arcosh":
        gs = g.pow(2).sum(-1)
        for i0 in range(0, q.shape[0], block):
            qq = q[i0:i0 + block]
            qs = qq.pow(2).sum(-1)
            # explicit differences: no ||x||^2+||y||^2-2<x,y> cancellation
            s = torch.cdist(qq, g, p=2, compute_mode="donot_use_mm_for_euclid_dist").pow(2)
            t = 2 * c * s / ((1 - c * qs)[:, None] * (1 - c * gs)[None, :])
            out[i0:i0 + block] = torch.log1p(t + torch.sqrt(t * (t + 2))) / math.sqrt(c)
    else:
        raise ValueError(form)
    return out


def topk_smallest(d: torch.Tensor, k: int):
    """Ascending, ties -> lower index (stable).  Returns (values[Q,k], idx[Q,k] int64)."""
    idx = torch.from_numpy(np.argsort(d.numpy(), axis=-1, kind="stable")[..., :k].copy())
    return torch.gather(d, -1, idx), idx


def hyperbolic_topk(q, g, c, k, form="geoopt"):
    return topk_smallest(hyperbolic_dist_rows(q, g, c, form=form), k)


def cosine_similarity(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise.cosine_similarity restated: rows L2-normalised
    (zero rows left as zero), then a dense product."""
    def _normalize(a):
        n = np.sqrt((a * a).sum(axis=1))
        n = np.where(n == 0.0, 1.0, n).astype(a.dtype)
        return a / n[:, None]
    return _normalize(x) @ _normalize(y).T


def cosine_topk(q: np.ndarray, g: np.ndarray, k: int):
    sim = cosine_similarity(q, g)
    idx = np.argsort(-sim, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(sim, idx, axis=1), idx.astype(np.int64)


# --------------------------------------------------------------------------- metrics (notebook)
def mrr_at_k(ranked, positives: set, k: int) -> float:
    for rank, item in enumerate(ranked[:k], 1):
        if item in positives:
            return 1.0 / rank
    return 0.0


def precision_at_k(ranked, positives: set, k: int) -> float:
    retrieved = ranked[:k]
    rel = len(set(retrieved).intersection(positives))
    return rel / k if k <= len(ranked) else 0.0


def average_precision_ranked(ranked, positives: set) -> float:
    relevant = 0
    ap = 0.0
    for j, item in enumerate(ranked, 1):
        if item in positives:
            relevant += 1
            ap += relevant / j
    return ap / len(positives) if len(positives) > 0 else 0


def ndcg_ranked(ranked, positives: set) -> float:
    idcg = sum(1 / np.log2(j + 2) for j in range(len(positives)))
    dcg = 0.0
    for j, item in enumerate(ranked):
        if item in positives:
            dcg += 1 / np.log2(j + 2)
    return dcg / idcg if idcg > 0 else 0


def recall_at_k(ranked, positives: set, k: int) -> float:
    return len(set(ranked[:k]).intersection(positives)) / len(positives) if len(positives) > 0 else 0


def notebook_metrics(ranked_lists, positives_list, ks=(5, 10, 20)):
    """Means over queries, like notebooks/retrieval.ipynb:446-456.  ``ranked_lists``
    may be truncated to the top-K (then AP / nDCG / MRR are their @K variants)."""
    out = {"mrr": [], "ap": [], "ndcg": []}
    for k in ks:
        out[f"mrr@{k}"] = []
        out[f"precision@{k}"] = []
        out[f"recall@{k}"] = []
    for ranked, pos in zip(ranked_lists, positives_list):
        ranked = [int(r) for r in ranked]
        pos = set(int(p) for p in pos)
        out["mrr"].append(mrr_at_k(ranked, pos, len(ranked)))
        out["ap"].append(average_precision_ranked(ranked, pos))
        out["ndcg"].append(ndcg_ranked(ranked, pos))
        for k in ks:
            out[f"mrr@{k}"].append(mrr_at_k(ranked, pos, k))
            out[f"precision@{k}"].append(precision_at_k(ranked, pos, k))
            out[f"recall@{k}"].append(recall_at_k(ranked, pos, k))
    return {name: float(np.mean(v)) if len(v) else 0.0 for name, v in out.items()}, out


# --------------------------------------------------------------------------- metrics (sklearn)
def average_precision_sklearn(target: np.ndarray, scores: np.ndarray) -> float:
    """sklearn.metrics.average_precision_score (binary) restated: thresholds are the
    DISTINCT scores (ties grouped), AP = sum_n (R_n - R_{n-1}) P_n."""
    target = np.asarray(target, dtype=np.float64)
    scores = np.asarray(scores)
    order = np.argsort(scores, kind="mergesort")[::-1]
    s = scores[order]
    t = target[order]
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, t.size - 1]
    tps = np.cumsum(t)[thr]
    fps = 1 + thr - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    # sklearn reverses and appends (P=1, R=0); AP = -sum(diff(recall) * precision[:-1])
    precision = np.r_[precision[::-1], 1.0]
    recall = np.r_[recall[::-1], 0.0]
    return float(-np.sum(np.diff(recall) * precision[:-1]))


def evaluate_retrieval(fig_emb: torch.Tensor, patent_emb: torch.Tensor, positives_list, c: float,
                       form: str = "geoopt") -> float:
    """Core of src/train.py:3221-3293 given already-encoded queries: per query
    ``dist(q, patents)`` -> ``scores=-dist`` -> sklearn AP -> mean.  Queries with no
    in-range positive, or with a NaN/inf distance (3262), are skipped."""
    k = torch.tensor(-float(c), dtype=fig_emb.dtype)
    num_patents = patent_emb.shape[0]
    aps = []
    for i, pos in enumerate(positives_list):
        valid = [p for p in pos if 0 <= p < num_patents]
        if not valid:
            continue
        if form == "geoopt":
            d = pmath.dist(fig_emb[i].unsqueeze(0), patent_emb, k=k)
        else:
            d = pmath.dist_arcosh(fig_emb[i].unsqueeze(0), patent_emb, k=k)
        if torch.isnan(d).any() or torch.isinf(d).any():
            continue
        target = np.zeros(num_patents, dtype=np.float32)
        target[valid] = 1
        ap = average_precision_sklearn(target, (-d).numpy())
        if not np.isnan(ap):
            aps.append(ap)
    return float(np.mean(aps)) if aps else 0.0
