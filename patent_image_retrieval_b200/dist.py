"""Multi-GPU: gallery row-sharding and the candidate exchange.

One process per GPU (torch.distributed, NCCL over NVLink).  Gallery rows are sharded
contiguously.  Two query layouts:

* ``queries="replicated"`` -- every rank holds the same ``[Q,D]`` batch, searches its shard
  (projection, tcgen05 scoring, exact rerank -- all shard-local) and only the ``[Q,k]`` (score,
  global index) lists cross NVLink: one ``all_gather`` each, then a merge kernel on every rank.
  Payload Q*k*12 bytes per rank (1.2 MB at Q=10k, k=10): latency-bound.
* ``queries="sharded"`` -- the serving layout: every rank is fed its OWN ``[Ql,D]`` batch (its own
  host link), the batches are all-gathered over NVLink (W*Ql*D*4 bytes, 164 MB at W=8, Ql=10k,
  D=512), every rank searches all W*Ql queries against its shard, and an ``all_to_all`` returns
  each query's W per-shard lists to the rank that owns the query, which merges them.  Per-GPU
  work is Ql x N_total pairs whatever W is, so query throughput grows with the number of GPUs.

The reference has no distributed path at all (single process, single GPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from .retrieval import GalleryIndex


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of ``rank`` (first ``n_rows % world`` ranks get one more)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_candidates(score: torch.Tensor, idx: torch.Tensor, group=None):
    """all_gather the per-shard ``[Q,k]`` lists -> ``([W,Q,k] scores, [W,Q,k] global idx)``.
    Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    Q = score.shape[0]
    gs = torch.empty((world * Q,) + tuple(score.shape[1:]), dtype=score.dtype, device=score.device)
    gi = torch.empty((world * Q,) + tuple(idx.shape[1:]), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gs, score.contiguous(), group=group)     # concatenated along dim 0
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    return gs.view((world,) + tuple(score.shape)), gi.view((world,) + tuple(idx.shape))


def gather_queries(q_local: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather the per-rank query batches ``[Ql,D]`` -> ``[W*Ql,D]`` (rank-major).  Every rank
    must pass the same ``Ql`` (pad the last batch)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * q_local.shape[0],) + tuple(q_local.shape[1:]), dtype=q_local.dtype,
                      device=q_local.device)
    dist.all_gather_into_tensor(out, q_local.contiguous(), group=group)
    return out


def return_lists_to_owners(score: torch.Tensor, idx: torch.Tensor, group=None):
    """``score/idx [W*Ql,k]`` = this shard's lists for ALL queries (rank-major).  all_to_all: block r
    goes to rank r.  Returns ``([W,Ql,k], [W,Ql,k])``: the W per-shard lists of this rank's own queries."""
    world = dist.get_world_size(group)
    rs, ri = torch.empty_like(score), torch.empty_like(idx)
    dist.all_to_all_single(rs, score.contiguous(), group=group)
    dist.all_to_all_single(ri, idx.contiguous(), group=group)
    ql = score.shape[0] // world
    return rs.view((world, ql) + tuple(score.shape[1:])), ri.view((world, ql) + tuple(idx.shape[1:]))


class ShardedGalleryIndex:
    """This rank's row-shard of a global gallery of ``n_total`` rows + the exchange step."""

    def __init__(self, shard_features: torch.Tensor, row_offset: int, n_total: int, c: float = 1.0,
                 metric: str = "hyperbolic", space: str = "euclidean", group=None,
                 device: Optional[torch.device] = None, queries: str = "replicated"):
        if queries not in ("replicated", "sharded"):
            raise ValueError(queries)
        self.queries = queries
        self.group = group
        self.n_total = int(n_total)
        self.local = GalleryIndex(shard_features, c=c, metric=metric, space=space, idx_offset=row_offset,
                                  device=device)
        self.metric = metric

    def _single(self) -> bool:
        return not dist.is_initialized() or dist.get_world_size(self.group) == 1

    def search(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
               kernel_events: Optional[list] = None):
        """``queries`` = the replicated batch, or this rank's own batch when the index was built with
        ``queries="sharded"``.  Returns the global top-k of the queries passed in."""
        if self.queries == "sharded" and not self._single():
            return self.search_sharded(queries, k=k, kprime=kprime, kernel_events=kernel_events)
        return self.search_replicated(queries, k=k, kprime=kprime, kernel_events=kernel_events)

    def search_replicated(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
                          kernel_events: Optional[list] = None):
        """Every rank passes the SAME batch; every rank returns the same global ``[Q,k]`` lists."""
        score, idx = self.local.search(queries, k=k, kprime=kprime, kernel_events=kernel_events)
        if self._single():
            return score, idx
        gs, gi = gather_candidates(score, idx, self.group)
        return ops.merge_topk(gs, gi, descending=(self.metric == "cosine"))

    def search_sharded(self, q_local: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
                       kernel_events: Optional[list] = None, prune: Optional[bool] = None):
        """Serving layout: all_gather the per-rank batches, score them all against this shard,
        all_to_all the lists back to the query owners, merge.  Returns ``[Ql,k]`` for ``q_local``.

        ``prune`` (default: on for k <= 32): a query's exact rescoring needs only its GLOBAL approximate
        top-k', of which this shard holds k'/W on average.  The shards' k' best surrogate scores go to the
        query owner (all_to_all), which takes the k'-th smallest of the W*k' values (``hypret_kth_smallest``)
        and all_gathers that threshold; each shard then rescores exactly only the candidates at or below it
        (``hypret_rerank_pruned``), so the gather traffic of the rerank is shared between the shards instead
        of being repeated on each of them.  The merged result is the list the single-GPU path returns."""
        from .retrieval import default_kprime
        world = dist.get_world_size(self.group)
        q_local = q_local.to(device=self.local.device, dtype=torch.float32, non_blocking=True)
        q_all = gather_queries(q_local, self.group)
        kp = min(default_kprime(k) if kprime is None else int(kprime), ops.MAX_KPRIME)
        if prune is None:
            prune = k <= 32 and kp <= 32 and k <= kp
        q32, cs, ci, cnt = self.local.score_candidates(q_all, k=k, kprime=kprime, kernel_events=kernel_events)
        thr_all = None
        if prune:
            sel_s, sel_i = ops.cand_select(cs, ci, cnt)                              # [W*Ql, k']
            recv = torch.empty_like(sel_s)
            dist.all_to_all_single(recv, sel_s, group=self.group)                    # [W, Ql, k'] at the owner
            thr = ops.kth_smallest(recv.view(world, q_local.shape[0], kp), kp)       # global k'-th best surrogate
            thr_all = torch.empty(world * q_local.shape[0], dtype=torch.float32, device=thr.device)
            dist.all_gather_into_tensor(thr_all, thr, group=self.group)
            cs, ci, cnt = sel_s.unsqueeze(1), sel_i.unsqueeze(1), None               # one merged list per query
        score, idx = self.local.rerank_candidates(q32, cs, ci, k, prune_thr=thr_all, kernel_events=kernel_events,
                                                  list_count=cnt)
        rs, ri = return_lists_to_owners(score, idx, self.group)
        return ops.merge_topk(rs, ri, descending=(self.metric == "cosine"))


def full_ranking_ap(q32: torch.Tensor, shard_rows32: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor,
                    c: float = 1.0, metric: str = "hyperbolic", row_offset: int = 0, n_total: Optional[int] = None,
                    grouped_ties: bool = True, group=None):
    """Exact AP over the FULL ranking of every query (reference src/train.py:3259-3293, grouped ties; or
    notebooks/retrieval.ipynb:411-420, index tie-break) without ever forming the [Q,N] score matrix, with the
    gallery row-sharded across ranks ("collective 2").

    ``q32`` [Q,D] replicated exact query rows (points on the ball / raw features), ``shard_rows32`` this rank's
    gallery rows (global ids ``row_offset ..``), positives as a CSR of GLOBAL gallery ids.  Two all-reduces cross
    NVLink: the [nnz] keys of the (query, positive) pairs (each computed by the shard that owns the positive) and
    the [nnz,3] rank counts.  Works unsharded too (no process group).  Returns ``(mean_ap, ap [Q], valid [Q])``."""
    sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    n_total = int(n_total) if n_total is not None else int(shard_rows32.shape[0])
    keys = ops.pair_keys(q32, shard_rows32, pos_offsets, pos_items, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=group)          # exactly one shard contributes per pair
    counts, bad = ops.rank_count(q32, shard_rows32, pos_offsets, pos_items, keys, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(bad, op=dist.ReduceOp.SUM, group=group)
    return ops.ap_from_counts(pos_offsets, pos_items, keys, counts, bad, n_total, grouped_ties=grouped_ties)
