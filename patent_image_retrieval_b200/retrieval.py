"""Gallery index + search: the user-facing call of the retrieval hot path.

Replaces, in the reference:
  * the per-query loop ``pmath.dist(q[1,D], patents[P,D])`` + ranking of
    ``evaluate_retrieval``                         (src/train.py:3221-3293)
  * ``cosine_similarity(Q, G)`` + ``np.argsort``   (notebooks/retrieval.ipynb:368-383)
  * ``ImageRetrieval.retrieve_similar_images``     (notebooks/retrieval.ipynb:190-206)

``GalleryIndex`` keeps, resident in HBM, the fp32 gallery rows (exact-rerank operand) and
the bf16 tensor-core operand built by the fused projection kernel.  ``search`` runs
projection(queries) -> tcgen05 scoring + streaming top-k' -> exact rerank.  All compute is
in libhypret.so; torch only owns the buffers.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def default_kprime(k: int) -> int:
    if k > ops.MAX_KPRIME:
        raise ValueError(f"k={k} exceeds the supported maximum {ops.MAX_KPRIME}")
    return min(ops.MAX_KPRIME, max(16, k + 6))


class GalleryIndex:
    """Row-shard of a gallery, resident on one GPU.

    metric='hyperbolic': ``features`` are Euclidean backbone features (``space='euclidean'``,
    mapped with project(expmap0(.)), reference src/models.py:310,317) or points already on the
    Poincare ball (``space='ball'``, e.g. outputs of ``encode_figures``).
    metric='cosine': ``features`` are raw embeddings (normalised like sklearn does)."""

    def __init__(self, features: torch.Tensor, c: float = 1.0, metric: str = "hyperbolic",
                 space: str = "euclidean", idx_offset: int = 0, device: Optional[torch.device] = None):
        if metric not in ("hyperbolic", "cosine"):
            raise ValueError(metric)
        if space not in ("euclidean", "ball"):
            raise ValueError(space)
        device = torch.device(device) if device is not None else (
            features.device if features.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        feats = features.to(device=device, dtype=torch.float32, non_blocking=True)
        self.c = float(c)
        self.metric = metric
        self.space = space
        self.idx_offset = int(idx_offset)
        self.device = device
        self.n, self.d = feats.shape
        if metric == "hyperbolic":
            mode = "expmap0" if space == "euclidean" else "onball"
            self.rows32, self.operand, _ = ops.project_rows(feats, c, mode=mode, side="gallery")
        else:
            self.rows32 = feats.contiguous()
            _, self.operand, _ = ops.project_rows(feats, 1.0, mode="cosine", side="gallery", want_point=False)
        self._cand = {}

    def _query_mode(self):
        if self.metric == "cosine":
            return "cosine"
        return "expmap0" if self.space == "euclidean" else "onball"

    def search(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None, return_margin: bool = False,
               max_ctas: int = 0, kernel_events: Optional[list] = None):
        """queries [Q,D] fp32 (host or device) -> (score [Q,k] f32, idx [Q,k] i64) on the device.
        score = Poincare distance ascending, or cosine similarity descending; ties -> lower index.
        ``kernel_events``: if a list, a (start, end) CUDA-event pair bracketing the scoring kernel
        on the launching stream is appended per call (bench.py's roofline measurement)."""
        kprime = default_kprime(k) if kprime is None else int(kprime)
        kprime = min(kprime, ops.MAX_KPRIME)
        if k > kprime:
            raise ValueError("k must be <= kprime")
        q = queries.to(device=self.device, dtype=torch.float32, non_blocking=True)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be [Q, {self.d}]")
        if self.metric == "hyperbolic":
            q32, q_op, _ = ops.project_rows(q, self.c, mode=self._query_mode(), side="query")
        else:
            q32 = q.contiguous()
            _, q_op, _ = ops.project_rows(q, 1.0, mode="cosine", side="query", want_point=False)
        key = (q.shape[0], kprime, max_ctas)
        plan = ops.score_plan(q.shape[0], self.n, self.d, kprime, max_ctas)
        buf = self._cand.get(key)
        if buf is None:
            self._cand.clear()
            buf = (torch.empty(q.shape[0], plan["n_lists"], kprime, dtype=torch.float32, device=self.device),
                   torch.empty(q.shape[0], plan["n_lists"], kprime, dtype=torch.int32, device=self.device),
                   torch.empty(q.shape[0], dtype=torch.int32, device=self.device))
            self._cand[key] = buf
        if kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        cs, ci = ops.score_topk(q_op, self.operand, self.d, kprime, max_ctas, out=buf[:2], thr_workspace=buf[2])
        if kernel_events is not None:
            e1.record()
            kernel_events.append((e0, e1))
        return ops.rerank(q32, self.rows32, cs, ci, self.c, self.metric, k, idx_offset=self.idx_offset,
                          want_margin=return_margin)
