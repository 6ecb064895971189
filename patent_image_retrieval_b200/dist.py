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

On one NVLink / NVSwitch box (the default deployment) no exchange of ``queries="sharded"`` is an NCCL call: the
projection kernel stores the operand rows into every rank's exchange buffer (``PeerQueryExchange``, CUDA IPC +
plain NVLink stores), the fp32 rows follow by copy engine under the scoring kernel, and ``cand_select`` /
``kth_smallest`` / the pruned rerank store their outputs into the receivers' regions, with stream-ordered step
counters instead of collectives (csrc/peer.cu).  ``HYPRET_PEER_EXCHANGE=0`` / ``HYPRET_PEER_ROUTE=0`` select the NCCL
forms; all variants return bit-identical lists.

The reference has no distributed path at all (single process, single GPU).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from .retrieval import GalleryIndex, _span


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


class _RawCuda:
    """``__cuda_array_interface__`` view of raw device memory (a buffer libhypret allocated)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank cannot allocate or map the exchange buffers (CUDA IPC refused, devices
    hidden from each other): the caller falls back to the NCCL all_gather."""


class PeerQueryExchange:
    """The all-gather of the per-rank query batches, done through peer memory instead of NCCL (csrc/peer.cu).

    Every rank owns an exchange buffer that all ranks of the box map over NVLink (CUDA IPC).  ``publish``:
      1. the projection kernel reads this rank's ``[Ql,D]`` raw rows once and stores each fp16 operand row into the
         buffers of ALL ranks (posted NVLink stores) -- no rank projects another rank's queries, and the operands
         are in place when the kernel ends;
      2. the exact fp32 rows (needed only by the rerank) follow through the copy engines on a side stream while
         the scoring kernel, which owns every SM, runs;
      3. per-source step counters in the destination buffers announce arrival (``hypret_peer_signal`` behind the
         producer, ``hypret_peer_wait`` in front of the consumer, both stream-ordered; no host synchronisation).
    Two buffer slots: a rank cannot start step i+2 before every peer has finished step i (see csrc/peer.cu)."""

    FLAGS_BYTES = 1024           # [0:64) operand counters, [64:128) point counters, [128:132) error word,
                                 # [192:256) scratch counters (kernel preload), then the counters of the routed
                                 # exchanges: [256:320) surrogate lists, [320:384) thresholds, [384:448) result lists
    F_OP, F_PT, F_ERR, F_SCRATCH, F_SEL, F_THR, F_LST = 0, 64, 128, 192, 256, 320, 384
    SLOTS = 2
    MAX_K = 32                   # the pruned protocol serves k <= k' <= 32

    def __init__(self, n_queries: int, d: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise ValueError("PeerQueryExchange addresses at most 16 ranks of one box")
        self.device = torch.device(device)
        self.ql, self.d = int(n_queries), int(d)
        self.kpad = ops.operand_kpad(d)
        rnd = lambda b: (b + 255) // 256 * 256
        self.op_blk = self.ql * self.kpad * 2
        self.pt_blk = self.ql * self.d * 4
        self.op_bytes, self.pt_bytes = rnd(self.world * self.op_blk), rnd(self.world * self.pt_blk)
        self.slot_bytes = self.op_bytes + self.pt_bytes
        # receive regions of the routed exchanges (single-buffered: everything behind the operand wait of step i+1
        # on a peer is behind the end of step i here)
        rows = self.world * self.ql
        self.off_sel = self.FLAGS_BYTES + self.SLOTS * self.slot_bytes      # [W,Ql,k'] fp32 surrogates per shard
        self.off_thr = self.off_sel + rnd(rows * self.MAX_K * 4)           # [W*Ql] fp32 global k'-th best surrogate
        self.off_ls = self.off_thr + rnd(rows * 4)                         # [W,Ql,k] fp32 per-shard result lists
        self.off_li = self.off_ls + rnd(rows * self.MAX_K * 4)             # [W,Ql,k] int64
        self.nbytes = self.off_li + rnd(rows * self.MAX_K * 8)
        lib = _lib.load()
        handle = (ctypes.c_ubyte * 64)()
        base = ctypes.c_void_p()
        self.base, self.peer_base, failure = None, [], None
        # every rank goes through every collective below whatever fails locally, and all ranks agree on the outcome
        with torch.cuda.device(self.device):
            rc = lib.hypret_peer_alloc(self.nbytes, ctypes.byref(base), handle)
        if rc != 0:
            failure = f"hypret_peer_alloc rc={rc}"
        else:
            self.base = int(base.value)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        ok = torch.tensor([0 if failure else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if bool(ok.item()):
            every = every.cpu().view(self.world, 64)
            with torch.cuda.device(self.device):
                for r in range(self.world):
                    if r == self.rank:
                        self.peer_base.append(self.base)
                        continue
                    h = (ctypes.c_ubyte * 64)(*every[r].tolist())
                    ptr = ctypes.c_void_p()
                    rc = lib.hypret_peer_open(h, ctypes.byref(ptr))
                    if rc != 0:
                        failure = failure or f"hypret_peer_open(rank {r}) rc={rc}"
                        self.peer_base.append(None)
                    else:
                        self.peer_base.append(int(ptr.value))
            ok = torch.tensor([0 if failure else 1], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not bool(ok.item()):
            with torch.cuda.device(self.device):
                for r, ptr in enumerate(self.peer_base):
                    if r != self.rank and ptr is not None:
                        lib.hypret_peer_close(ctypes.c_void_p(ptr))
                if self.base is not None:
                    lib.hypret_peer_free(ctypes.c_void_p(self.base))
            self.base = None
            raise PeerExchangeUnavailable(failure or "a peer could not map the exchange buffers")
        raw = torch.as_tensor(_RawCuda(self.base, self.nbytes), device=self.device)
        self._raw = raw
        self.op_all, self.pt_all = [], []
        for s in range(self.SLOTS):
            o = self.FLAGS_BYTES + s * self.slot_bytes
            self.op_all.append(raw[o:o + self.world * self.op_blk].view(torch.float16).view(self.world * self.ql,
                                                                                              self.kpad))
            o += self.op_bytes
            self.pt_all.append(raw[o:o + self.world * self.pt_blk].view(torch.float32).view(self.world * self.ql,
                                                                                            self.d))
        self.err = raw[128:132].view(torch.int32)
        self.route = _lib.PeerRoute()
        self.route.n_ranks, self.route.me, self.route.ql = self.world, self.rank, self.ql
        for r in range(self.world):
            self.route.base[r] = self.peer_base[r]
        self.side = torch.cuda.Stream(device=self.device)
        self.projected = torch.cuda.Event()
        self.copied = [torch.cuda.Event() for _ in range(self.SLOTS)]
        self.step = 0
        self._arr = ctypes.c_void_p * self.world
        # run both flag kernels once while nothing spins: a kernel's first launch loads its module (lazy loading),
        # and that waits for running kernels -- it must never happen behind a wait that is already spinning
        with torch.cuda.device(self.device):
            cur = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            scratch = self._arr(*[self.base + 192 + 4 * r for r in range(self.world)])
            _lib.check(lib.hypret_peer_signal(scratch, self.world, 1, cur))
            _lib.check(lib.hypret_peer_wait(ctypes.c_void_p(self.base + 192), self.world, 1,
                                            ctypes.c_void_p(self.base + 128), cur))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)              # every rank has mapped every buffer before the first store

    def _off(self, slot: int, points: bool) -> int:
        return self.FLAGS_BYTES + slot * self.slot_bytes + (self.op_bytes if points else 0)

    def publish(self, q_local: torch.Tensor, c: float, mode: str, op_err: Optional[torch.Tensor] = None):
        """Queue projection + exchange of this rank's batch on the current stream.  Returns
        ``(q_op_all [W*Ql,kpad] fp16, q32_all [W*Ql,D] fp32)`` (rank-major); ``q_op_all`` is complete for whatever is
        queued behind this call, ``q32_all`` only behind ``wait_points()``.  ``op_err`` [Ql] fp32 (optional) receives the
        rounding-residual norms of this rank's operand rows (the certificate's ``q_err``)."""
        if tuple(q_local.shape) != (self.ql, self.d) or q_local.dtype != torch.float32 or not q_local.is_cuda:
            raise ValueError(f"expected a CUDA float32 batch of shape [{self.ql}, {self.d}]")
        q_local = q_local.contiguous()
        lib = _lib.load()
        self.step += 1
        s = self.step % self.SLOTS
        cur = torch.cuda.current_stream(self.device)
        cur_p, side_p = ctypes.c_void_p(cur.cuda_stream), ctypes.c_void_p(self.side.cuda_stream)
        W, me = self.world, self.rank
        my_pt = self.base + self._off(s, True) + me * self.pt_blk
        with torch.cuda.device(self.device):
            cur.wait_event(self.copied[s])     # the copies that last read this slot's local block are done
            dsts = self._arr(*[self.peer_base[r] + self._off(s, False) + me * self.op_blk for r in range(W)])
            cosine = mode == "cosine"
            _lib.check(lib.hypret_project_rows_peers(ctypes.c_void_p(q_local.data_ptr()), self.ql, self.d, float(c),
                                                     ops.MODE[mode], None if cosine else ctypes.c_void_p(my_pt), dsts,
                                                     W, ops._ptr(op_err), cur_p))
            if cosine:                         # cosine reranks from the raw rows
                _lib.check(lib.hypret_peer_copy(ctypes.c_void_p(my_pt), ctypes.c_void_p(q_local.data_ptr()),
                                                self.pt_blk, cur_p))
            flags = self._arr(*[self.peer_base[r] + 4 * me for r in range(W)])
            _lib.check(lib.hypret_peer_signal(flags, W, self.step, cur_p))
            self.projected.record(cur)
            self.side.wait_event(self.projected)
            for r in range(W):
                if r != me:
                    _lib.check(lib.hypret_peer_copy(
                        ctypes.c_void_p(self.peer_base[r] + self._off(s, True) + me * self.pt_blk),
                        ctypes.c_void_p(my_pt), self.pt_blk, side_p))
            flags = self._arr(*[self.peer_base[r] + 64 + 4 * me for r in range(W)])
            _lib.check(lib.hypret_peer_signal(flags, W, self.step, side_p))
            self.copied[s].record(self.side)
            _lib.check(lib.hypret_peer_wait(ctypes.c_void_p(self.base), W, self.step, ctypes.c_void_p(self.base + 128),
                                            cur_p))
        q_local.record_stream(cur)
        return self.op_all[s], self.pt_all[s]

    def wait_points(self):
        """Hold the current stream until every rank's fp32 rows of the last published step have landed."""
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hypret_peer_wait(ctypes.c_void_p(self.base + 64), self.world, self.step,
                                                    ctypes.c_void_p(self.base + 128),
                                                    ctypes.c_void_p(cur.cuda_stream)))

    # ---- the exchanges behind the scoring kernel, fused into the kernels that produce the data --------------
    def _region(self, off: int, width: int, dtype) -> torch.Tensor:
        n = self.world * self.ql * width * torch.empty((), dtype=dtype).element_size()
        return self._raw[off:off + n].view(dtype).view(self.world, self.ql, width)

    def _signal(self, flag_off: int):
        cur = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        flags = self._arr(*[self.peer_base[r] + flag_off + 4 * self.rank for r in range(self.world)])
        _lib.check(_lib.load().hypret_peer_signal(flags, self.world, self.step, cur))

    def _wait(self, flag_off: int):
        cur = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.load().hypret_peer_wait(ctypes.c_void_p(self.base + flag_off), self.world, self.step,
                                                ctypes.c_void_p(self.base + self.F_ERR), cur))

    def select_and_send(self, cand_score, cand_idx, list_count):
        """``ops.cand_select`` whose kernel also stores every query's ``[k']`` surrogates into the receive region of
        the query's owner (the all_to_all).  Returns the local ``(sel_score [W*Ql,k'], sel_idx)``; behind this call
        the region ``[W,Ql,k']`` of this rank holds every shard's list of its own queries."""
        Q, S, kp = cand_score.shape
        if kp > self.MAX_K or Q != self.world * self.ql:
            raise ValueError("select_and_send: shape does not match the exchange")
        ss = torch.empty(Q, kp, dtype=torch.float32, device=self.device)
        si = torch.empty(Q, kp, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hypret_cand_select_route(
                ops._ptr(cand_score.contiguous()), ops._ptr(cand_idx.contiguous()), ops._ptr(list_count), Q, S, kp,
                ops._ptr(ss), ops._ptr(si), ctypes.byref(self.route), self.off_sel, ops._stream()))
            self._signal(self.F_SEL)
            self._wait(self.F_SEL)
        return ss, si, self._region(self.off_sel, kp, torch.float32)

    def threshold_to_all(self, recv: torch.Tensor, kth: int) -> torch.Tensor:
        """``ops.kth_smallest`` over the received lists of this rank's queries, stored into EVERY rank's threshold
        region (the all_gather).  Returns this rank's ``[W*Ql]`` region, complete behind this call."""
        W, Ql, m = recv.shape
        own = torch.empty(Ql, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hypret_kth_smallest_route(ops._ptr(recv), W, Ql, m, int(kth), ops._ptr(own),
                                                             ctypes.byref(self.route), self.off_thr, ops._stream()))
            self._signal(self.F_THR)
            self._wait(self.F_THR)
        return self._region(self.off_thr, 1, torch.float32).view(self.world * self.ql)

    def rerank_to_owners(self, local: GalleryIndex, q32, sel_s, sel_i, k: int, thr_all):
        """Pruned exact rerank whose ``[k]`` result lists go straight into the query owners' receive regions (the
        two all_to_alls).  Returns ``(scores [W,Ql,k], idx [W,Ql,k])`` of this rank's own queries."""
        Q, kp = sel_s.shape
        if k > self.MAX_K or kp > self.MAX_K:
            raise ValueError("rerank_to_owners serves k <= k' <= 32")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hypret_rerank_pruned_route(
                ops._ptr(q32), ops._ptr(local.rows32), Q, local.n, local.d, float(local.c), ops.METRIC[local.metric],
                ops._ptr(sel_s), ops._ptr(sel_i), 1, kp, int(k), int(local.idx_offset), ops._ptr(thr_all),
                ctypes.byref(self.route), self.off_ls, self.off_li, ops._stream()))
            self._signal(self.F_LST)
            self._wait(self.F_LST)
        return self._region(self.off_ls, k, torch.float32), self._region(self.off_li, k, torch.int64)

    def check(self):
        """Synchronise and raise if a wait ran into its 20 s bound (a peer died or fell out of step)."""
        torch.cuda.synchronize(self.device)
        e = int(self.err.item())
        if e:
            raise RuntimeError(f"peer exchange: rank {self.rank} timed out waiting for rank {e - 1}")

    def release(self):
        """``close`` without the barrier, for an owner that is being garbage-collected: the caller vouches that no
        peer is still inside a step that writes here."""
        self.close(barrier=False)

    def close(self, barrier: bool = True):
        if self.base is None:
            return
        torch.cuda.synchronize(self.device)
        if barrier:
            dist.barrier(group=self.group)     # nobody still writes into a buffer that is about to go
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for r, ptr in enumerate(self.peer_base):
                if r != self.rank:
                    lib.hypret_peer_close(ctypes.c_void_p(ptr))
            self._raw = self.op_all = self.pt_all = self.err = None
            lib.hypret_peer_free(ctypes.c_void_p(self.base))
        self.base = None


def peer_exchange_available(device: torch.device, group=None) -> bool:
    """Peer-memory exchange needs NCCL-style deployment: one process per GPU of ONE box, every device visible to
    every process and peer access between them.  ``HYPRET_PEER_EXCHANGE=0`` forces the NCCL all_gather."""
    if os.environ.get("HYPRET_PEER_EXCHANGE", "1") == "0" or not dist.is_initialized():
        return False
    world = dist.get_world_size(group)
    if dist.get_backend(group) != "nccl" or world < 2 or world > 16 or torch.cuda.device_count() < world:
        return False
    if int(os.environ.get("LOCAL_WORLD_SIZE", world)) != world:
        return False
    dev = torch.device(device).index
    ok = all(r == dev or torch.cuda.can_device_access_peer(dev, r) for r in range(world))
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class ShardedGalleryIndex:
    """This rank's row-shard of a global gallery of ``n_total`` rows + the exchange step."""

    def __init__(self, shard_features: torch.Tensor, row_offset: int, n_total: int, c: float = 1.0,
                 metric: str = "hyperbolic", space: str = "euclidean", group=None,
                 device: Optional[torch.device] = None, queries: str = "replicated", exact: bool = True):
        if queries not in ("replicated", "sharded"):
            raise ValueError(queries)
        self.queries = queries
        self.exact = bool(exact)        # sharded queries: certify the merged lists and rescan what cannot be proven
        self.group = group
        self.n_total = int(n_total)
        self.local = GalleryIndex(shard_features, c=c, metric=metric, space=space, idx_offset=row_offset,
                                  device=device)
        self.metric = metric
        self._exchange = None
        self._exchange_ok = None
        self._stats_all = None          # maxima over ALL shards of the gallery statistics (certificate of merged lists)
        self.uncertified = None         # [Ql] int32 of the last search_sharded: 1 = recomputed by the exact scan

    def stats_all(self) -> torch.Tensor:
        """``local.stats`` maximised over the shards (one all_reduce, at first use; collective)."""
        if self._stats_all is None:
            st = self.local.stats.clone()
            if not self._single():
                dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
            self._stats_all = st
        return self._stats_all

    def _peer_exchange(self, n_queries: int) -> Optional[PeerQueryExchange]:
        """The peer-memory exchange for ``n_queries``-row batches (built at first use; collective), or None."""
        if self._exchange_ok is None:
            self._exchange_ok = peer_exchange_available(self.local.device, self.group)
        if not self._exchange_ok:
            return None
        if self._exchange is None or self._exchange.ql != n_queries:
            if self._exchange is not None:
                self._exchange.close()
                self._exchange = None
            try:
                self._exchange = PeerQueryExchange(n_queries, self.local.d, self.local.device, self.group)
            except PeerExchangeUnavailable as exc:      # raised on all ranks together
                import warnings
                warnings.warn(f"peer-memory query exchange unavailable ({exc}); using the NCCL all_gather")
                self._exchange_ok = False
        return self._exchange

    def _single(self) -> bool:
        return not dist.is_initialized() or dist.get_world_size(self.group) == 1

    def check(self) -> None:
        """Synchronise and raise if a peer-exchange wait ran into its time bound since the last check."""
        if self._exchange is not None:
            self._exchange.check()

    def exchange_failed(self) -> bool:
        """The exchange's error word, read WITHOUT a device synchronise of its own (call it where the results of a
        step have already been waited for, e.g. SearchPipeline.result)."""
        return self._exchange is not None and self._exchange.err is not None and bool(int(self._exchange.err.item()))

    def close(self) -> None:
        """Release the peer-memory exchange buffers (collective: every rank of the group calls it)."""
        if self._exchange is not None:
            self._exchange.close()
            self._exchange = None

    def __del__(self):
        # an index dropped without close(): unmap what this rank mapped, without the collective barrier -- the peers'
        # own buffers stay valid until they drop theirs
        ex = getattr(self, "_exchange", None)
        if ex is not None:
            try:
                ex.release()
            except Exception:
                pass

    def search(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
               kernel_events: Optional[list] = None):
        """``queries`` = the replicated batch, or this rank's own batch when the index was built with
        ``queries="sharded"``.  Returns the global top-k of the queries passed in."""
        if self.queries == "sharded" and not self._single():
            return self.search_sharded(queries, k=k, kprime=kprime, kernel_events=kernel_events, exact=self.exact)
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
                       kernel_events: Optional[list] = None, prune: Optional[bool] = None, exact: bool = True):
        """Serving layout: all_gather the per-rank batches, score them all against this shard,
        all_to_all the lists back to the query owners, merge.  Returns ``[Ql,k]`` for ``q_local``.
        On one NVLink box the all_gather is the projection kernel itself storing into every rank's exchange
        buffer (``PeerQueryExchange``); the result is bit-identical to the NCCL path (``HYPRET_PEER_EXCHANGE=0``).

        ``prune`` (default: on for k <= 32): a query's exact rescoring needs only its GLOBAL approximate
        top-k', of which this shard holds k'/W on average.  The shards' k' best surrogate scores go to the
        query owner (all_to_all), which takes the k'-th smallest of the W*k' values (``hypret_kth_smallest``)
        and all_gathers that threshold; each shard then rescores exactly only the candidates at or below it
        (``hypret_rerank_pruned``), so the gather traffic of the rerank is shared between the shards instead
        of being repeated on each of them.  The merged result is the list the single-GPU path returns.

        ``exact`` (with ``prune``): the owner proves its merged lists exact (``hypret_cert_merged``: global k'-th best
        filter score minus the exact surrogate of the k-th result against the rounding bound); the flags of the queries
        it cannot prove are all_gathered, every shard rescans those queries exactly (``hypret_exact_topk``, list
        built on the device) and the owners merge the scans -- fixed-size collectives, no host round trip;
        ``self.uncertified`` keeps the flags."""
        from .retrieval import default_kprime
        world = dist.get_world_size(self.group)
        q_local = q_local.to(device=self.local.device, dtype=torch.float32, non_blocking=True)
        kp = min(default_kprime(k) if kprime is None else int(kprime), ops.MAX_KPRIME)
        if prune is None:
            prune = k <= 32 and kp <= 32 and k <= kp
        exact = bool(exact and prune)
        ql, me = q_local.shape[0], dist.get_rank(self.group)
        stats_all = self.stats_all() if exact else None
        ex = self._peer_exchange(ql)
        q_err = None
        if ex is not None:
            q_err = torch.empty(ql, dtype=torch.float32, device=q_local.device) if exact else None
            with _span(kernel_events, "project"):
                q_op, q32 = ex.publish(q_local, self.local.c, self.local._query_mode(), op_err=q_err)
            q32, cs, ci, cnt = self.local.score_projected(q32, q_op, k=k, kprime=kprime, kernel_events=kernel_events)
        else:
            q_all = gather_queries(q_local, self.group)
            q32, cs, ci, cnt, q_err = self.local.score_candidates(q_all, k=k, kprime=kprime,
                                                                  kernel_events=kernel_events, want_err=exact)
            q_err = q_err[me * ql:(me + 1) * ql] if exact else None

        def poisoned(score, idx):
            # a wait of the peer exchange that ran into its time bound (dead / out-of-step peer) leaves stale data in
            # the exchange buffers: such a step hands out idx = -1 everywhere instead of plausible wrong lists, and
            # SearchPipeline.result / PeerQueryExchange.check raise (no host sync here: the error word is read on the device)
            if ex is None:
                return score, idx
            return score, torch.where(ex.err.view(1, 1) != 0, -1, idx)

        def finish(rs, ri, thr_all):
            score, idx = ops.merge_topk(rs, ri, descending=(self.metric == "cosine"))
            if not exact:
                self.uncertified = None
                return poisoned(score, idx)
            with _span(kernel_events, "certify"):
                own = slice(me * ql, (me + 1) * ql)
                flags = ops.cert_merged(q32[own], score, idx, thr_all[own], q_err, stats_all, self.local.c, self.metric)
                # one all_gather carries the flag and the owner's k-th merged score (the rescans' warm-start bound)
                fb = torch.stack([flags.float(), score[:, k - 1]], dim=1)
                fb_all = torch.empty(world * ql, 2, dtype=torch.float32, device=flags.device)
                dist.all_gather_into_tensor(fb_all, fb, group=self.group)
                flags_all = (fb_all[:, 0] != 0).to(torch.int32)
                xs, xi = ops.exact_topk_flagged(q32, self.local.rows32, self.local.rows_sq64, flags_all, self.local.c,
                                                self.metric, k, idx_offset=self.local.idx_offset,
                                                init_bound=fb_all[:, 1])
                xs, xi = return_lists_to_owners(xs, xi, self.group)
                xs, xi = ops.merge_topk(xs, xi, descending=(self.metric == "cosine"))
                redo = flags.bool()[:, None]
                self.uncertified = flags
                return poisoned(torch.where(redo, xs, score), torch.where(redo, xi, idx))

        thr_all = None
        if prune and ex is not None and os.environ.get("HYPRET_PEER_ROUTE", "1") != "0":
            # every exchange is done by the kernel that produces the data (NVLink stores into the receivers'
            # regions + stream-ordered counters): no NCCL call on the data path of a step
            sel_s, sel_i, recv = ex.select_and_send(cs, ci, cnt)
            thr_all = ex.threshold_to_all(recv, kp)
            ex.wait_points()
            with _span(kernel_events, "rerank"):
                rs, ri = ex.rerank_to_owners(self.local, q32, sel_s, sel_i, k, thr_all)
            return finish(rs, ri, thr_all)
        if prune:
            sel_s, sel_i = ops.cand_select(cs, ci, cnt)                              # [W*Ql, k']
            recv = torch.empty_like(sel_s)
            dist.all_to_all_single(recv, sel_s, group=self.group)                    # [W, Ql, k'] at the owner
            thr = ops.kth_smallest(recv.view(world, q_local.shape[0], kp), kp)       # global k'-th best surrogate
            thr_all = torch.empty(world * q_local.shape[0], dtype=torch.float32, device=thr.device)
            dist.all_gather_into_tensor(thr_all, thr, group=self.group)
            cs, ci, cnt = sel_s.unsqueeze(1), sel_i.unsqueeze(1), None               # one merged list per query
        if ex is not None:
            ex.wait_points()
        score, idx = self.local.rerank_candidates(q32, cs, ci, k, prune_thr=thr_all, kernel_events=kernel_events,
                                                  list_count=cnt)
        rs, ri = return_lists_to_owners(score, idx, self.group)
        if thr_all is None:
            self.uncertified = None
            return poisoned(*ops.merge_topk(rs, ri, descending=(self.metric == "cosine")))
        return finish(rs, ri, thr_all)


def _explicitly_sharded(sharded: Optional[bool], group) -> bool:
    """The row-shard collectives of the full-ranking metrics run only when the caller asks for them."""
    want = bool(sharded) if sharded is not None else group is not None
    if want and not dist.is_initialized():
        raise RuntimeError("sharded=True needs an initialised torch.distributed process group")
    return want and dist.get_world_size(group) > 1


def full_ranking_ap(q32: torch.Tensor, shard_rows32: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor,
                    c: float = 1.0, metric: str = "hyperbolic", row_offset: int = 0, n_total: Optional[int] = None,
                    grouped_ties: bool = True, group=None, sharded: Optional[bool] = None):
    """Exact AP over the FULL ranking of every query (reference src/train.py:3259-3293, grouped ties; or
    notebooks/retrieval.ipynb:411-420, index tie-break) without ever forming the [Q,N] score matrix, with the
    gallery row-sharded across ranks ("collective 2").

    ``q32`` [Q,D] replicated exact query rows (points on the ball / raw features), ``shard_rows32`` this rank's
    gallery rows (global ids ``row_offset ..``), positives as a CSR of GLOBAL gallery ids.  Two all-reduces cross
    NVLink: the [nnz] keys of the (query, positive) pairs (each computed by the shard that owns the positive) and
    the [nnz,3] rank counts.  Returns ``(mean_ap, ap [Q], valid [Q])``.

    Sharding is EXPLICIT: the collectives run only with ``sharded=True`` (or a ``group``); a call without either is
    the unsharded path even inside an initialised multi-rank job -- ``evaluate_retrieval`` passes the whole patent
    table on every rank, and reducing that would multiply keys and counts by the world size."""
    sharded = _explicitly_sharded(sharded, group)
    n_total = int(n_total) if n_total is not None else int(shard_rows32.shape[0])
    keys = ops.pair_keys(q32, shard_rows32, pos_offsets, pos_items, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=group)          # exactly one shard contributes per pair
    counts, bad = ops.rank_count(q32, shard_rows32, pos_offsets, pos_items, keys, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(bad, op=dist.ReduceOp.SUM, group=group)
    return ops.ap_from_counts(pos_offsets, pos_items, keys, counts, bad, n_total, grouped_ties=grouped_ties)
