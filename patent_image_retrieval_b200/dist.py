"""Multi-GPU: gallery row-sharding and the candidate exchange.

One process per GPU (torch.distributed, NCCL over NVLink).  Gallery rows are sharded
contiguously; queries are replicated.  Each rank searches its shard (projection, tcgen05
scoring, exact rerank -- all shard-local) and only the ``[Q,k]`` (score, global index) lists
cross NVLink: one ``all_gather`` each, then a merge kernel.  The payload is Q*k*12 bytes per
rank (1.2 MB at Q=10k, k=10), i.e. latency-bound; nothing else is exchanged.

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


class ShardedGalleryIndex:
    """This rank's row-shard of a global gallery of ``n_total`` rows + the exchange step."""

    def __init__(self, shard_features: torch.Tensor, row_offset: int, n_total: int, c: float = 1.0,
                 metric: str = "hyperbolic", space: str = "euclidean", group=None,
                 device: Optional[torch.device] = None):
        self.group = group
        self.n_total = int(n_total)
        self.local = GalleryIndex(shard_features, c=c, metric=metric, space=space, idx_offset=row_offset,
                                  device=device)
        self.metric = metric

    def search(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
               kernel_events: Optional[list] = None):
        score, idx = self.local.search(queries, k=k, kprime=kprime, kernel_events=kernel_events)
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return score, idx
        gs, gi = gather_candidates(score, idx, self.group)
        return ops.merge_topk(gs, gi, descending=(self.metric == "cosine"))
