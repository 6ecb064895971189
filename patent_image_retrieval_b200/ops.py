"""Torch-facing wrappers over the C ABI (include/hypret.h).

PyTorch is plumbing here: it owns device memory and streams; every operator below
passes raw device pointers into libhypret.so.  CUDA tensors only -- there is no CPU
or eager fallback, a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

MODE = {"expmap0": 0, "onball": 1, "cosine": 2}
SIDE = {"query": 0, "gallery": 1}
METRIC = {"cosine": 0, "hyperbolic": 1}
MAX_KPRIME = 64      # list slots per strip
MAX_K = 128          # results per query (k > 32 uses the wide rerank kernel)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hypret operators run on CUDA tensors only (no CPU fallback)")


def operand_kpad(d: int) -> int:
    return int(_lib.load().hypret_operand_kpad(int(d)))


def project_rows(u: torch.Tensor, c: float = 1.0, mode: str = "expmap0", side: str = "query",
                 want_point: bool = True, want_operand: bool = True, want_sqnorm: bool = False,
                 want_err: bool = False, stats: Optional[torch.Tensor] = None):
    """Fused expmap0/project (or project only, or L2-normalise) + fp16 GEMM operand + ||y||^2.

    Returns ``(y32 | None, operand | None, sqnorm | None)`` -- with ``want_err`` a fourth element, the ``[n]`` fp32
    rounding-residual norms of the operand rows; ``stats`` (``[4]`` fp32, zeroed once per gallery) accumulates the
    gallery maxima.  Both feed the exact-top-k certificate of ``rerank_cert`` (``hypret_project_rows_cert``).
    Mirrors pmath.expmap0 -> pmath.project (reference src/models.py:310,317)."""
    _need_cuda(u, stats)
    if u.dtype != torch.float32 or u.dim() != 2:
        raise ValueError("u must be a [n, d] float32 tensor")
    u = u.contiguous()
    n, d = u.shape
    y = torch.empty_like(u) if want_point else None
    op = torch.empty(n, operand_kpad(d), dtype=torch.float16, device=u.device) if want_operand else None
    sq = torch.empty(n, dtype=torch.float32, device=u.device) if want_sqnorm else None
    if want_err or stats is not None:
        if stats is not None and (stats.dtype != torch.float32 or stats.numel() != 4 or not stats.is_contiguous()):
            raise ValueError("stats must be a contiguous [4] float32 CUDA tensor")
        err = torch.empty(n, dtype=torch.float32, device=u.device) if want_err else None
        with torch.cuda.device(u.device):
            _lib.check(_lib.load().hypret_project_rows_cert(_ptr(u), n, d, float(c), MODE[mode], SIDE[side], _ptr(y),
                                                            _ptr(op), _ptr(sq), _ptr(err), _ptr(stats), _stream()))
        return (y, op, sq, err) if want_err else (y, op, sq)
    with torch.cuda.device(u.device):
        _lib.check(_lib.load().hypret_project_rows(_ptr(u), n, d, float(c), MODE[mode], SIDE[side], _ptr(y), _ptr(op),
                                                   _ptr(sq), _stream()))
    return y, op, sq


def score_plan(Q: int, N: int, d: int, kprime: int, max_ctas: int = 0, min_lists: int = 0) -> dict:
    """Strip schedule + shared-memory plan of ``score_topk`` (host-only; works without a GPU)."""
    return _score_plan_struct(Q, N, d, kprime, max_ctas, min_lists).asdict()


def _score_plan_struct(Q, N, d, kprime, max_ctas=0, min_lists=0):
    plan = _lib.ScorePlan()
    _lib.check(_lib.load().hypret_score_plan(int(Q), int(N), int(d), int(kprime), int(max_ctas), int(min_lists),
                                             ctypes.byref(plan)))
    return plan


def score_strips(Q: int, N: int, d: int, kprime: int, max_ctas: int = 0, min_lists: int = 0):
    """All strips of the schedule as ``(unit, step, first_query_tile, g0, g1, slot)`` tuples (host-only).
    A strip covers ``plan["pair"]`` consecutive query tiles."""
    plan = _score_plan_struct(Q, N, d, kprime, max_ctas, min_lists)
    out = (ctypes.c_int32 * 4)()
    strips = []
    lib = _lib.load()
    for cta in range(plan.grid // plan.pair):
        for step in range(plan.n_steps):
            rc = lib.hypret_score_strip(ctypes.byref(plan), cta, step, out)
            if rc < 0:
                _lib.check(rc)
            if rc == 1:
                strips.append((cta, step, int(out[0]), int(out[1]), int(out[2]), int(out[3])))
    return strips


def score_topk(q_op: torch.Tensor, g_op: torch.Tensor, d: int, kprime: int, max_ctas: int = 0,
               debug: bool = False, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               share_thresholds: bool = True, thr_workspace: Optional[torch.Tensor] = None, min_lists: int = 0,
               list_count: Optional[torch.Tensor] = None, kbound: Optional[int] = None):
    """tcgen05 scoring GEMM + streaming top-k'.  Returns ``(cand_score [Q,L,k'], cand_idx [Q,L,k'] int32)``
    with L = plan["n_lists"] (+ the full ``[Q,N]`` surrogate matrix when ``debug`` -- tests only).
    ``share_thresholds``: strips of a query exchange their running k'-th best score (fast path);
    off, every list is exactly the top-k' of its own strip.
    ``list_count`` ([Q] int32, out): compact list slots -- the lists of query q are its first ``list_count[q]`` slots
    (pass the same tensor to ``rerank`` / ``cand_select``); without it slots are the schedule's and unwritten ones
    read as empty.
    ``kbound`` (k' <= 16 < kbound <= 32, shared thresholds): the lists share the bound of the kbound-th best score -- the
    union holds the query's top-kbound (``hypret_score_topk_bound``; ``rerank_cert(ksel=kbound)`` certifies it)."""
    _need_cuda(q_op, g_op, list_count)
    kbound = int(kprime) if kbound is None else int(kbound)
    if kbound != kprime and not (share_thresholds and kprime <= 16 < kbound <= 32):
        raise ValueError("kbound needs shared thresholds and k' <= 16 < kbound <= 32")
    if list_count is not None and (list_count.dtype != torch.int32 or list_count.numel() < q_op.shape[0]):
        raise ValueError("list_count must be an int32 tensor with one entry per query")
    if q_op.dtype != torch.float16 or g_op.dtype != torch.float16:
        raise ValueError("operands must be fp16 rows from project_rows")
    kpad = operand_kpad(d)
    if q_op.shape[1] != kpad or g_op.shape[1] != kpad or not q_op.is_contiguous() or not g_op.is_contiguous():
        raise ValueError(f"operands must be contiguous [rows, {kpad}]")
    Q, N = q_op.shape[0], g_op.shape[0]
    with torch.cuda.device(q_op.device):
        plan = score_plan(Q, N, d, kprime, max_ctas, min_lists)
        S = plan["n_lists"]
        if out is None:
            cs = torch.empty(Q, S, kprime, dtype=torch.float32, device=q_op.device)
            ci = torch.empty(Q, S, kprime, dtype=torch.int32, device=q_op.device)
        else:
            cs, ci = out
            if tuple(cs.shape) != (Q, S, kprime) or tuple(ci.shape) != (Q, S, kprime):
                raise ValueError("preallocated candidate buffers do not match the score plan")
        dbg = torch.empty(Q, N, dtype=torch.float32, device=q_op.device) if debug else None
        ws = None
        if share_thresholds:
            ws = thr_workspace if thr_workspace is not None else torch.empty(Q, dtype=torch.int32, device=q_op.device)
            if ws.numel() < Q or ws.element_size() != 4 or not ws.is_cuda:
                raise ValueError("thr_workspace must be a CUDA tensor of >= Q 32-bit elements")
        _lib.check(_lib.load().hypret_score_topk_bound(_ptr(q_op), Q, _ptr(g_op), N, int(d), int(kprime), kbound, S,
                                                       int(max_ctas), int(min_lists), _ptr(cs), _ptr(ci), _ptr(ws),
                                                       _ptr(list_count), _ptr(dbg), _stream()))
    return (cs, ci, dbg) if debug else (cs, ci)


def rerank(q32: torch.Tensor, g32: torch.Tensor, cand_score: torch.Tensor, cand_idx: torch.Tensor, c: float,
           metric: str, k: int, idx_offset: int = 0, want_margin: bool = False,
           prune_thr: Optional[torch.Tensor] = None, list_count: Optional[torch.Tensor] = None,
           g_sqnorm64: Optional[torch.Tensor] = None):
    """Merge candidate lists, exact fp64-accumulated rescoring, sorted top-k.
    Returns ``(score [Q,k] f32, idx [Q,k] i64[, margin [Q] f32])``.
    ``g_sqnorm64 [N]`` fp64 (``row_sqnorm64(g32)``, once per index): lets the wide path (k > 32) skip the
    per-survivor row norm.
    ``prune_thr [Q]`` (multi-GPU): candidates whose surrogate exceeds it are skipped (``hypret_rerank_pruned``)."""
    _need_cuda(q32, g32, cand_score, cand_idx, prune_thr, list_count)
    if prune_thr is not None:
        if list_count is not None:
            raise ValueError("a pruned rerank takes the selected lists of cand_select (no list_count)")
        if want_margin:
            raise ValueError("the margin certificate is not defined for a pruned rerank")
        q32, g32 = q32.contiguous(), g32.contiguous()
        Q, d = q32.shape
        _, S, kprime = cand_score.shape
        out_s = torch.empty(Q, k, dtype=torch.float32, device=q32.device)
        out_i = torch.empty(Q, k, dtype=torch.int64, device=q32.device)
        thr = prune_thr.contiguous().float()
        if thr.numel() != Q:
            raise ValueError("prune_thr must hold one threshold per query")
        with torch.cuda.device(q32.device):
            _lib.check(_lib.load().hypret_rerank_pruned(_ptr(q32), _ptr(g32), Q, g32.shape[0], d, float(c),
                                                        METRIC[metric], _ptr(cand_score.contiguous()),
                                                        _ptr(cand_idx.contiguous()), S, kprime, int(k),
                                                        int(idx_offset), _ptr(thr), _ptr(out_s), _ptr(out_i),
                                                        _stream()))
        return out_s, out_i
    q32 = q32.contiguous()
    g32 = g32.contiguous()
    Q, d = q32.shape
    N = g32.shape[0]
    _, S, kprime = cand_score.shape
    out_s = torch.empty(Q, k, dtype=torch.float32, device=q32.device)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=q32.device)
    margin = torch.empty(Q, dtype=torch.float32, device=q32.device) if want_margin else None
    if g_sqnorm64 is not None and (g_sqnorm64.dtype != torch.float64 or g_sqnorm64.numel() != N or
                                   not g_sqnorm64.is_cuda or not g_sqnorm64.is_contiguous()):
        raise ValueError("g_sqnorm64 must be a contiguous CUDA float64 tensor with one entry per gallery row")
    with torch.cuda.device(q32.device):
        _lib.check(_lib.load().hypret_rerank(_ptr(q32), _ptr(g32), Q, N, d, float(c), METRIC[metric],
                                             _ptr(cand_score), _ptr(cand_idx), _ptr(list_count), S, kprime, int(k),
                                             int(idx_offset),
                                             _ptr(out_s), _ptr(out_i), _ptr(margin), _ptr(g_sqnorm64), _stream()))
    return (out_s, out_i, margin) if want_margin else (out_s, out_i)


class CertBuffers:
    """Device state of the exact-top-k guarantee for batches of up to ``Q`` queries: the uncertified-query list and its
    count, the lock words of the fallback merge, and the per-query ``certified`` flags (1 = the filter pass was proven
    exact, 0 = recomputed by the full scan).  Nothing here is read by the host on the search path."""

    def __init__(self, Q: int, device):
        self.Q = int(Q)
        self.state = torch.empty(2 * Q, dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        self.list = torch.empty(Q, dtype=torch.int32, device=device)
        self.certified = torch.empty(Q, dtype=torch.uint8, device=device)
        self.bound = torch.empty(Q, dtype=torch.float32, device=device)      # k-th filtered score of uncertified queries


def rerank_cert(q32: torch.Tensor, g32: torch.Tensor, cand_score: torch.Tensor, cand_idx: torch.Tensor, c: float,
                metric: str, k: int, q_err: torch.Tensor, g_stats: torch.Tensor, g_sqnorm64: torch.Tensor,
                bufs: CertBuffers, idx_offset: int = 0, want_margin: bool = False,
                list_count: Optional[torch.Tensor] = None, fallback: bool = True, ksel: int = 0):
    """``rerank`` (k <= k' <= 32) with the exact-top-k guarantee: ``hypret_rerank_cert`` proves per query that no row
    outside the fp16-filtered candidate set can precede the k-th result, and ``hypret_exact_topk`` -- queued right
    behind it, reading the list length on the device -- recomputes the queries it could not prove from all gallery
    rows.  Returns ``(score [Q,k], idx [Q,k][, margin [Q]])``; ``bufs.certified`` / ``bufs.count`` tell what happened.
    ``fallback=False`` (tests): certificate only.  ``ksel`` (> k'): the lists were built with ``score_topk(kbound=ksel)``."""
    _need_cuda(q32, g32, cand_score, cand_idx, q_err, g_stats, g_sqnorm64, list_count)
    q32, g32 = q32.contiguous(), g32.contiguous()
    Q, d = q32.shape
    N = g32.shape[0]
    _, S, kprime = cand_score.shape
    if bufs.Q < Q:
        raise ValueError("CertBuffers too small for this batch")
    out_s = torch.empty(Q, k, dtype=torch.float32, device=q32.device)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=q32.device)
    margin = torch.empty(Q, dtype=torch.float32, device=q32.device) if want_margin else None
    lib = _lib.load()
    with torch.cuda.device(q32.device):
        _lib.check(lib.hypret_rerank_cert(_ptr(q32), _ptr(g32), Q, N, d, float(c), METRIC[metric], _ptr(cand_score),
                                          _ptr(cand_idx), _ptr(list_count), S, kprime, int(ksel), int(k),
                                          int(idx_offset), _ptr(out_s), _ptr(out_i), _ptr(margin), _ptr(q_err),
                                          _ptr(g_stats),
                                          _ptr(bufs.state), _ptr(bufs.count), _ptr(bufs.list), _ptr(bufs.bound),
                                          _ptr(bufs.certified), _stream()))
        if fallback:
            _lib.check(lib.hypret_exact_topk(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, N, d, float(c), METRIC[metric],
                                             int(k), int(idx_offset), _ptr(bufs.list), _ptr(bufs.count),
                                             _ptr(bufs.state), _ptr(bufs.bound), _ptr(out_s), _ptr(out_i), _stream()))
    return (out_s, out_i, margin) if want_margin else (out_s, out_i)


def exact_topk(q32: torch.Tensor, g32: torch.Tensor, g_sqnorm64: torch.Tensor, c: float, metric: str, k: int,
               idx_offset: int = 0):
    """Exact top-k of EVERY query by a full scan (``hypret_exact_topk`` over the identity list): the reference's
    per-query loop + ``torch.topk`` (src/train.py:3259, src/auxiliary.py:374) as one kernel.  Returns
    ``(score [Q,k], idx [Q,k])``."""
    _need_cuda(q32, g32, g_sqnorm64)
    q32, g32 = q32.contiguous().float(), g32.contiguous().float()
    Q, d = q32.shape
    dev = q32.device
    out_s = torch.empty(Q, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=dev)
    lst = torch.arange(Q, dtype=torch.int32, device=dev)
    cnt = torch.full((1,), Q, dtype=torch.int32, device=dev)
    state = torch.zeros(2 * Q, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hypret_exact_topk(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d, float(c),
                                                 METRIC[metric], int(k), int(idx_offset), _ptr(lst), _ptr(cnt),
                                                 _ptr(state), None, _ptr(out_s), _ptr(out_i), _stream()))
    return out_s, out_i


def cert_merged(q32: torch.Tensor, score: torch.Tensor, idx: torch.Tensor, thr: torch.Tensor, q_err: torch.Tensor,
                g_stats: torch.Tensor, c: float, metric: str, want_margin: bool = False):
    """Certificate of merged lists (``hypret_cert_merged``): ``flags [Q] int32``, 1 = NOT proven exact."""
    _need_cuda(q32, score, idx, thr, q_err, g_stats)
    q32, score, idx = q32.contiguous(), score.contiguous(), idx.contiguous()
    Q, d = q32.shape
    flags = torch.empty(Q, dtype=torch.int32, device=q32.device)
    margin = torch.empty(Q, dtype=torch.float32, device=q32.device) if want_margin else None
    with torch.cuda.device(q32.device):
        _lib.check(_lib.load().hypret_cert_merged(_ptr(q32), Q, d, float(c), METRIC[metric], _ptr(score), _ptr(idx),
                                                  score.shape[1], _ptr(thr.contiguous()), _ptr(q_err.contiguous()),
                                                  _ptr(g_stats), _ptr(flags), _ptr(margin), _stream()))
    return (flags, margin) if want_margin else flags


def exact_topk_flagged(q32: torch.Tensor, g32: torch.Tensor, g_sqnorm64: torch.Tensor, flags: torch.Tensor, c: float,
                       metric: str, k: int, idx_offset: int = 0, init_bound: Optional[torch.Tensor] = None):
    """Exact top-k over this shard of the FLAGGED queries only (``hypret_flag_compact`` + ``hypret_exact_topk``, the
    list built and read on the device).  Returns ``(score [Q,k], idx [Q,k])``; rows of unflagged queries are +-inf / -1.
    ``init_bound [Q]``: a score no better than each query's true k-th best (warm start, ``hypret_exact_topk``)."""
    _need_cuda(q32, g32, g_sqnorm64, flags, init_bound)
    q32, g32 = q32.contiguous(), g32.contiguous()
    Q, d = q32.shape
    dev = q32.device
    out_s = torch.full((Q, k), float("inf") if metric == "hyperbolic" else float("-inf"), dtype=torch.float32, device=dev)
    out_i = torch.full((Q, k), -1, dtype=torch.int64, device=dev)
    lst = torch.empty(Q, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    state = torch.empty(2 * Q, dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.hypret_flag_compact(_ptr(flags.contiguous()), Q, _ptr(lst), _ptr(cnt), _ptr(state), _stream()))
        _lib.check(lib.hypret_exact_topk(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d, float(c),
                                         METRIC[metric], int(k), int(idx_offset), _ptr(lst), _ptr(cnt), _ptr(state),
                                         _ptr(init_bound.contiguous().float() if init_bound is not None else None),
                                         _ptr(out_s), _ptr(out_i), _stream()))
    return out_s, out_i


def _packed_keys(score: torch.Tensor, idx: torch.Tensor, metric: str, idx_offset: int) -> torch.Tensor:
    """The 64-bit ordering keys of result entries: ordered fp32 key << 32 | local row id (csrc/exact.cu)."""
    key = score if metric == "hyperbolic" else -score
    bits = key.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    hi = torch.where(bits >= 0x80000000, (~bits) & 0xFFFFFFFF, bits | 0x80000000)
    packed = (hi << 32) | ((idx - idx_offset) & 0xFFFFFFFF)
    return torch.where(idx >= 0, packed, torch.full_like(packed, -1))       # -1 = all ones: nothing lies above it


def exact_topk_any(q32: torch.Tensor, g32: torch.Tensor, g_sqnorm64: torch.Tensor, c: float, metric: str, k: int,
                   idx_offset: int = 0):
    """Exact top-k for ANY k by paging through the exact ranking 32 rows at a time (``hypret_exact_topk_after``): one
    full scan per page.  For the k beyond the filtered path's 128 (``retrieve_similar_images(k=500)``)."""
    _need_cuda(q32, g32, g_sqnorm64)
    q32, g32 = q32.contiguous().float(), g32.contiguous().float()
    Q, d = q32.shape
    dev = q32.device
    lst = torch.arange(Q, dtype=torch.int32, device=dev)
    cnt = torch.full((1,), Q, dtype=torch.int32, device=dev)
    after = torch.zeros(Q, dtype=torch.int64, device=dev)
    pages_s, pages_i = [], []
    lib = _lib.load()
    first = True
    for k0 in range(0, k, 32):
        kk = min(32, k - k0)
        out_s = torch.empty(Q, kk, dtype=torch.float32, device=dev)
        out_i = torch.empty(Q, kk, dtype=torch.int64, device=dev)
        state = torch.zeros(2 * Q, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            if first:
                _lib.check(lib.hypret_exact_topk(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d, float(c),
                                                 METRIC[metric], kk, int(idx_offset), _ptr(lst), _ptr(cnt), _ptr(state),
                                                 None, _ptr(out_s), _ptr(out_i), _stream()))
            else:
                _lib.check(lib.hypret_exact_topk_after(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d,
                                                       float(c), METRIC[metric], kk, int(idx_offset), _ptr(lst),
                                                       _ptr(cnt), _ptr(state), _ptr(after), _ptr(out_s), _ptr(out_i),
                                                       _stream()))
        first = False
        pages_s.append(out_s)
        pages_i.append(out_i)
        after = _packed_keys(out_s[:, -1], out_i[:, -1], metric, idx_offset)
    return torch.cat(pages_s, dim=1), torch.cat(pages_i, dim=1)


def exact_topk_any_flagged(q32: torch.Tensor, g32: torch.Tensor, g_sqnorm64: torch.Tensor, flags: torch.Tensor, c: float,
                           metric: str, k: int, idx_offset: int = 0):
    """``exact_topk_any`` for the FLAGGED queries only (list built on the device, nothing read by the host): pages of 32
    through the exact ranking.  Returns ``(score [Q,k], idx [Q,k])``; rows of unflagged queries are +-inf / -1."""
    _need_cuda(q32, g32, g_sqnorm64, flags)
    q32, g32 = q32.contiguous().float(), g32.contiguous().float()
    Q, d = q32.shape
    dev = q32.device
    fill = float("inf") if metric == "hyperbolic" else float("-inf")
    lst = torch.empty(Q, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    state = torch.empty(2 * Q, dtype=torch.int32, device=dev)
    flags = flags.contiguous().to(torch.int32)
    after = torch.zeros(Q, dtype=torch.int64, device=dev)
    pages_s, pages_i = [], []
    lib = _lib.load()
    for k0 in range(0, k, 32):
        kk = min(32, k - k0)
        out_s = torch.full((Q, kk), fill, dtype=torch.float32, device=dev)
        out_i = torch.full((Q, kk), -1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hypret_flag_compact(_ptr(flags), Q, _ptr(lst), _ptr(cnt), _ptr(state), _stream()))
            if k0 == 0:
                _lib.check(lib.hypret_exact_topk(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d, float(c),
                                                 METRIC[metric], kk, int(idx_offset), _ptr(lst), _ptr(cnt), _ptr(state),
                                                 None, _ptr(out_s), _ptr(out_i), _stream()))
            else:
                _lib.check(lib.hypret_exact_topk_after(_ptr(q32), _ptr(g32), _ptr(g_sqnorm64), Q, g32.shape[0], d,
                                                       float(c), METRIC[metric], kk, int(idx_offset), _ptr(lst),
                                                       _ptr(cnt), _ptr(state), _ptr(after), _ptr(out_s), _ptr(out_i),
                                                       _stream()))
        pages_s.append(out_s)
        pages_i.append(out_i)
        after = _packed_keys(out_s[:, -1], out_i[:, -1], metric, idx_offset)
    return torch.cat(pages_s, dim=1), torch.cat(pages_i, dim=1)


def certificate_bound(q32: torch.Tensor, q_err: torch.Tensor, g_stats: torch.Tensor, c: float, metric: str,
                      d: int) -> torch.Tensor:
    """The rounding bound E of the fp16 filter per query (the formula hypret_rerank_cert evaluates in the
    kernel), for paths that compare a margin with it on the device: ``[Q]`` fp32."""
    st = g_stats.double()
    qn = (float(c) * row_sqnorm(q32).double()).sqrt() if metric == "hyperbolic" else torch.ones_like(q_err, dtype=torch.float64)
    slack = (operand_kpad(d) / 16 + 8) * 2.0 ** -22
    return (q_err.double() * st[0] + qn * st[1] + slack * (qn * st[0] + qn * qn * st[2] + st[3])).float()


def row_sqnorm64(x: torch.Tensor) -> torch.Tensor:
    """``||x_i||^2`` accumulated in fp64 (``hypret_row_sqnorm64``): the per-row constant of the exact rerank."""
    _need_cuda(x)
    x = x.contiguous().float()
    out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hypret_row_sqnorm64(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream()))
    return out


def cand_select(cand_score: torch.Tensor, cand_idx: torch.Tensor, list_count: Optional[torch.Tensor] = None):
    """The k' best candidates of each query by surrogate score: ``[Q,L,k'] -> ([Q,k'] f32 ascending, [Q,k'] i32)``,
    padded with (+inf, -1)."""
    _need_cuda(cand_score, cand_idx)
    Q, S, kprime = cand_score.shape
    ss = torch.empty(Q, kprime, dtype=torch.float32, device=cand_score.device)
    si = torch.empty(Q, kprime, dtype=torch.int32, device=cand_score.device)
    with torch.cuda.device(cand_score.device):
        _lib.check(_lib.load().hypret_cand_select(_ptr(cand_score.contiguous()), _ptr(cand_idx.contiguous()),
                                                  _ptr(list_count), Q, S, kprime, _ptr(ss), _ptr(si), _stream()))
    return ss, si


def kth_smallest(vals: torch.Tensor, kth: int) -> torch.Tensor:
    """``vals [W,Q,m]`` fp32 -> ``[Q]``: the kth smallest (1-based) of each query's W*m values (+inf if fewer)."""
    _need_cuda(vals)
    vals = vals.contiguous().float()
    W, Q, m = vals.shape
    out = torch.empty(Q, dtype=torch.float32, device=vals.device)
    with torch.cuda.device(vals.device):
        _lib.check(_lib.load().hypret_kth_smallest(_ptr(vals), W, Q, m, int(kth), _ptr(out), _stream()))
    return out


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, descending: bool = False):
    """Merge per-shard lists ``scores/idx [W,Q,k]`` (global indices) into the global top-k ``[Q,k]``."""
    _need_cuda(scores, idx)
    scores = scores.contiguous()
    idx = idx.contiguous()
    W, Q, k = scores.shape
    out_s = torch.empty(Q, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.load().hypret_merge_topk(_ptr(scores), _ptr(idx), W, Q, k, int(bool(descending)), _ptr(out_s),
                                                 _ptr(out_i), _stream()))
    return out_s, out_i


def pairdist(a: torch.Tensor, p: torch.Tensor, c: float) -> torch.Tensor:
    """Exact Poincare distance matrix ``[n,m]`` (fp32) between on-ball points ``a [n,d]`` and ``p [m,d]``."""
    _need_cuda(a, p)
    a = a.contiguous().float()
    p = p.contiguous().float()
    n, d = a.shape
    m = p.shape[0]
    out = torch.empty(n, m, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hypret_pairdist(_ptr(a), _ptr(p), n, m, d, float(c), _ptr(out), _stream()))
    return out


METRIC_COLUMNS = ("mrr", "ap", "ndcg")


def metric_names(ks):
    names = list(METRIC_COLUMNS)
    for k in ks:
        names += [f"mrr@{k}", f"precision@{k}", f"recall@{k}"]
    return names


def retrieval_metrics(ranked: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor, ks=(5, 10, 20),
                      n_pos_total: Optional[torch.Tensor] = None):
    """Metrics of ranked lists ``[Q,K]`` (reference notebooks/retrieval.ipynb:310-324,411-456).
    Returns ``(means: dict name -> float, per_query [Q, 3+3*len(ks)] fp64 on the device)``."""
    _need_cuda(ranked, pos_offsets, pos_items, n_pos_total)
    ranked = ranked.contiguous().to(torch.int64)
    pos_offsets = pos_offsets.contiguous().to(torch.int64)
    pos_items = pos_items.contiguous().to(torch.int64)
    if n_pos_total is not None:
        n_pos_total = n_pos_total.contiguous().to(torch.int32)
    Q, K = ranked.shape
    ks = [int(k) for k in ks]
    ncol = 3 + 3 * len(ks)
    per_query = torch.empty(Q, ncol, dtype=torch.float64, device=ranked.device)
    means = torch.zeros(ncol, dtype=torch.float64, device=ranked.device)
    ks_arr = (ctypes.c_int32 * max(1, len(ks)))(*ks)
    with torch.cuda.device(ranked.device):
        _lib.check(_lib.load().hypret_retrieval_metrics(_ptr(ranked), Q, K, _ptr(pos_offsets), _ptr(pos_items),
                                                        _ptr(n_pos_total), ks_arr, len(ks), _ptr(per_query),
                                                        _ptr(means), _stream()))
    return dict(zip(metric_names(ks), means.tolist())), per_query


def ap_full(scores: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor, grouped_ties: bool = True):
    """AP over full score rows ``[Q,N]`` (higher = better).  ``grouped_ties`` = sklearn semantics
    (reference src/train.py:3285).  Returns ``(mean_ap float, ap [Q] fp64, valid [Q] int32)``."""
    _need_cuda(scores, pos_offsets, pos_items)
    scores = scores.contiguous().float()
    pos_offsets = pos_offsets.contiguous().to(torch.int64)
    pos_items = pos_items.contiguous().to(torch.int64)
    Q, N = scores.shape
    ap = torch.empty(Q, dtype=torch.float64, device=scores.device)
    valid = torch.empty(Q, dtype=torch.int32, device=scores.device)
    mean = torch.zeros(1, dtype=torch.float64, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.load().hypret_ap_full(_ptr(scores), Q, N, _ptr(pos_offsets), _ptr(pos_items),
                                              int(bool(grouped_ties)), _ptr(ap), _ptr(valid), _ptr(mean), _stream()))
    return float(mean.item()), ap, valid


def pair_keys(q32: torch.Tensor, g32: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor, c: float,
              metric: str, idx_offset: int = 0) -> torch.Tensor:
    """Key (distance / minus cosine) of every (query, positive) pair whose gallery row is on this shard
    (global ids ``idx_offset .. idx_offset + len(g32) - 1``), 0 elsewhere: sum over shards = all keys."""
    _need_cuda(q32, g32, pos_offsets, pos_items)
    q32, g32 = q32.contiguous().float(), g32.contiguous().float()
    pos_offsets, pos_items = pos_offsets.contiguous().to(torch.int64), pos_items.contiguous().to(torch.int64)
    keys = torch.zeros(max(1, pos_items.numel()), dtype=torch.float32, device=q32.device)
    with torch.cuda.device(q32.device):
        _lib.check(_lib.load().hypret_pair_keys(_ptr(q32), _ptr(g32), q32.shape[0], g32.shape[0], q32.shape[1],
                                                float(c), METRIC[metric], _ptr(pos_offsets), _ptr(pos_items),
                                                int(idx_offset), _ptr(keys), _stream()))
    return keys[:pos_items.numel()]


def rank_count(q32: torch.Tensor, g32: torch.Tensor, pos_offsets: torch.Tensor, pos_items: torch.Tensor,
               pos_keys: torch.Tensor, c: float, metric: str, idx_offset: int = 0):
    """This shard's rank counts of every (query, positive) pair: ``counts [nnz,3]`` int64
    {#better, #tied with a lower global id, #tied} and ``bad [Q]`` int32 (non-finite scores); sum over shards."""
    _need_cuda(q32, g32, pos_offsets, pos_items, pos_keys)
    q32, g32 = q32.contiguous().float(), g32.contiguous().float()
    pos_offsets, pos_items = pos_offsets.contiguous().to(torch.int64), pos_items.contiguous().to(torch.int64)
    nnz = pos_items.numel()
    counts = torch.zeros(max(1, nnz), 3, dtype=torch.int64, device=q32.device)
    bad = torch.zeros(q32.shape[0], dtype=torch.int32, device=q32.device)
    keys = pos_keys.contiguous().float()
    if keys.numel() == 0:
        keys = torch.zeros(1, dtype=torch.float32, device=q32.device)
    with torch.cuda.device(q32.device):
        _lib.check(_lib.load().hypret_rank_count(_ptr(q32), _ptr(g32), q32.shape[0], g32.shape[0], q32.shape[1],
                                                 float(c), METRIC[metric], _ptr(pos_offsets), _ptr(pos_items),
                                                 _ptr(keys), int(idx_offset), _ptr(counts), _ptr(bad), _stream()))
    return counts[:nnz], bad


def ap_from_counts(pos_offsets: torch.Tensor, pos_items: torch.Tensor, pos_keys: torch.Tensor, counts: torch.Tensor,
                   bad: torch.Tensor, n_total: int, grouped_ties: bool = True):
    """AP per query from global rank counts.  Returns ``(mean_ap float, ap [Q] fp64, valid [Q] int32)``."""
    _need_cuda(pos_offsets, pos_items, pos_keys, counts, bad)
    pos_offsets, pos_items = pos_offsets.contiguous().to(torch.int64), pos_items.contiguous().to(torch.int64)
    Q = pos_offsets.numel() - 1
    dev = pos_offsets.device
    ap = torch.empty(Q, dtype=torch.float64, device=dev)
    valid = torch.empty(Q, dtype=torch.int32, device=dev)
    mean = torch.zeros(1, dtype=torch.float64, device=dev)
    keys = pos_keys.contiguous().float()
    cnt = counts.contiguous().to(torch.int64)
    if keys.numel() == 0:
        keys = torch.zeros(1, dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, 3, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hypret_ap_from_counts(_ptr(pos_offsets), _ptr(pos_items), _ptr(keys), _ptr(cnt),
                                                     _ptr(bad.contiguous().to(torch.int32)), Q, int(n_total),
                                                     int(bool(grouped_ties)), _ptr(ap), _ptr(valid), _ptr(mean),
                                                     _stream()))
    return float(mean.item()), ap, valid


GRAM_MAX_D = 1024      # tensor-core distance matrix: feature dimensions it supports
GRAM_MIN_PAIRS = 1 << 16   # below this the CUDA-core tile kernel is as fast (launch-bound either way)


def gram_dist(a: torch.Tensor, p: torch.Tensor, c: float):
    """Distance matrix ``[n,m]`` on the tensor cores (3-way bf16 split operands, exact recompute of near pairs);
    training accuracy (~1e-6 relative).  Returns ``(dmat, |a|^2 [n], |p|^2 [m])``."""
    _need_cuda(a, p)
    a, p = a.contiguous().float(), p.contiguous().float()
    n, d = a.shape
    m = p.shape[0]
    lib = _lib.load()
    kp = int(lib.hypret_gram_kpad(d))
    a_op = torch.empty(n, kp, dtype=torch.bfloat16, device=a.device)
    p_op = torch.empty(m, kp, dtype=torch.bfloat16, device=a.device)
    asq = torch.empty(n, dtype=torch.float32, device=a.device)
    psq = torch.empty(m, dtype=torch.float32, device=a.device)
    out = torch.empty(n, m, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.hypret_gram_split(_ptr(a), n, d, 0, _ptr(a_op), _ptr(asq), _stream()))
        _lib.check(lib.hypret_gram_split(_ptr(p), m, d, 1, _ptr(p_op), _ptr(psq), _stream()))
        _lib.check(lib.hypret_gram_dist(_ptr(a_op), _ptr(p_op), _ptr(a), _ptr(p), _ptr(asq), _ptr(psq), n, m, d,
                                        float(c), _ptr(out), _stream()))
    return out, asq, psq


def pairdist_ce_fwd(a: torch.Tensor, p: torch.Tensor, c: float, inv_tau: float, want_cols: bool,
                    tensor_cores: Optional[bool] = None):
    """Distance matrix (fp32 tail) + row / column log-sum-exps of ``-D * inv_tau``.
    ``tensor_cores`` (default: whenever the shape allows): the Gram matrix comes from the tcgen05 kernel of
    ``gram_dist`` instead of the FP32-FMA tile kernel.  Returns ``(dmat [n,m], row_lse [n], col_lse [m] | None)``."""
    _need_cuda(a, p)
    a, p = a.contiguous().float(), p.contiguous().float()
    n, d = a.shape
    m = p.shape[0]
    if tensor_cores is None:
        tensor_cores = d <= GRAM_MAX_D and d % 4 == 0 and n * m >= GRAM_MIN_PAIRS
    row_lse = torch.empty(n, dtype=torch.float32, device=a.device)
    col_lse = torch.empty(m, dtype=torch.float32, device=a.device) if want_cols else None
    n_part = max(1, min(64, (n + 127) // 128))
    scratch = torch.empty(2 * n_part * m, dtype=torch.float32, device=a.device) if want_cols else None
    if tensor_cores:
        dmat, _, _ = gram_dist(a, p, c)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().hypret_neg_lse(_ptr(dmat), n, m, float(inv_tau), int(bool(want_cols)), _ptr(row_lse),
                                                  _ptr(col_lse), _ptr(scratch), n_part, _stream()))
        return dmat, row_lse, col_lse
    dmat = torch.empty(n, m, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hypret_pairdist_ce_fwd(_ptr(a), _ptr(p), n, m, d, float(c), float(inv_tau),
                                                      int(bool(want_cols)), _ptr(dmat), _ptr(row_lse), _ptr(col_lse),
                                                      _ptr(scratch), n_part, _stream()))
    return dmat, row_lse, col_lse


# ---- flash-style train_hyp step (csrc/flash.cu): no [n,m] array in memory, every dense product on tcgen05 -----------
FLASH_MAX_D = 128
FLASH_MIN_PAIRS = 1 << 16


def flash_ok(n: int, m: int, d: int) -> bool:
    """Shapes the flash kernels serve (others use the matrix kernels above)."""
    return 16 <= d <= FLASH_MAX_D and d % 16 == 0 and n * m >= FLASH_MIN_PAIRS


class FlashOperands:
    """Per-tensor operands of the flash kernels (``hypret_flash_prep``): fp16 2-way split Gram operands in the row and
    the column layout, transposed bf16 hi/mid planes for the gradient product, squared norms."""

    def __init__(self, x: torch.Tensor, want_row=True, want_col=True, want_t=True):
        _need_cuda(x)
        self.x = x.contiguous().float()
        n, d = self.x.shape
        lib = _lib.load()
        kp = int(lib.hypret_flash_kpad(d))
        dev = x.device
        self.n, self.d = n, d
        self.row = torch.empty(n, kp, dtype=torch.float16, device=dev) if want_row else None
        self.col = torch.empty(n, kp, dtype=torch.float16, device=dev) if want_col else None
        self.t_cols = (n + 63) // 64 * 64
        self.t = torch.empty(2, d, self.t_cols, dtype=torch.bfloat16, device=dev) if want_t else None
        self.sq = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hypret_flash_prep(_ptr(self.x), n, d, _ptr(self.row), _ptr(self.col), _ptr(self.t),
                                             self.t_cols, _ptr(self.sq), _stream()))


def _flash_workspace(n: int, m: int, d: int, device) -> torch.Tensor:
    return torch.empty(int(_lib.load().hypret_flash_workspace(n, m, d)), dtype=torch.float32, device=device)


def flash_lse(xo: FlashOperands, yo: FlashOperands, c: float, inv_tau: float) -> torch.Tensor:
    """``lse[i] = logsumexp_j(-dist(x_i, y_j) * inv_tau)`` without the distance matrix (``hypret_flash_lse``)."""
    out = torch.empty(xo.n, dtype=torch.float32, device=xo.x.device)
    ws = _flash_workspace(xo.n, yo.n, xo.d, xo.x.device)
    with torch.cuda.device(xo.x.device):
        _lib.check(_lib.load().hypret_flash_lse(_ptr(xo.row), _ptr(yo.col), _ptr(xo.x), _ptr(yo.x), _ptr(xo.sq),
                                                _ptr(yo.sq), xo.n, yo.n, xo.d, float(c), float(inv_tau), _ptr(ws),
                                                _ptr(out), _stream()))
    return out


def flash_grad(xo: FlashOperands, yo: FlashOperands, c: float, inv_tau: float, x_lse: Optional[torch.Tensor],
               y_lse: Optional[torch.Tensor], w_rows: float, w_cols: float, grad_scale: Optional[torch.Tensor] = None,
               diag_offset: int = 0, n_total: Optional[int] = None) -> torch.Tensor:
    """Gradient of the in-batch InfoNCE with respect to the ROW operand ``x`` (``hypret_flash_grad``); the gradient
    with respect to ``y`` is the same call with the roles swapped."""
    out = torch.empty(xo.n, xo.d, dtype=torch.float32, device=xo.x.device)
    ws = _flash_workspace(xo.n, yo.n, xo.d, xo.x.device)
    gs = grad_scale.reshape(1).contiguous().float() if grad_scale is not None else None
    with torch.cuda.device(xo.x.device):
        _lib.check(_lib.load().hypret_flash_grad(_ptr(xo.row), _ptr(yo.col), _ptr(yo.t), yo.t_cols, _ptr(xo.x),
                                                 _ptr(yo.x), _ptr(xo.sq), _ptr(yo.sq), _ptr(x_lse), _ptr(y_lse), xo.n,
                                                 yo.n, xo.d, float(c), float(inv_tau), float(w_rows), float(w_cols),
                                                 _ptr(gs), int(diag_offset), int(xo.n if n_total is None else n_total),
                                                 _ptr(ws), _ptr(out), _stream()))
    return out


def split_operand(x: torch.Tensor, side: str, want_sqnorm: bool = False):
    """fp16 2-way split GEMM operand of ``x [n,d]``: ``side='row'`` -> ``[hi|lo|hi]``, ``'col'`` -> ``[hi|hi|lo]``
    (``hypret_flash_prep``); a row operand times a column operand is the fp32 product to 2^-22.
    ``want_sqnorm``: also return ``||x_i||^2`` from the same pass."""
    _need_cuda(x)
    x = x.detach().contiguous().float()
    n, d = x.shape
    lib = _lib.load()
    op = torch.empty(n, int(lib.hypret_flash_kpad(d)), dtype=torch.float16, device=x.device)
    sq = torch.empty(n, dtype=torch.float32, device=x.device) if want_sqnorm else None
    with torch.cuda.device(x.device):
        _lib.check(lib.hypret_flash_prep(_ptr(x), n, d, _ptr(op) if side == "row" else None,
                                         _ptr(op) if side == "col" else None, None, 0, _ptr(sq), _stream()))
    return (op, sq) if want_sqnorm else op


def sgemm(a: torch.Tensor, b: torch.Tensor, row_scale: Optional[torch.Tensor] = None,
          addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a @ b (+ row_scale[:, None] * addend)`` for fp32 2-D tensors of ANY strides (transposed views cost nothing):
    the batch-sized products of the head's backward pass (``hypret_sgemm_strided``)."""
    _need_cuda(a, b, row_scale, addend)
    if a.dtype != torch.float32 or b.dtype != torch.float32 or a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[0]:
        raise ValueError("sgemm: fp32 [m,k] @ [k,n]")
    m, k = a.shape
    n = b.shape[1]
    out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    if row_scale is not None:
        row_scale, addend = row_scale.contiguous().float(), addend.contiguous().float()
        if row_scale.numel() != m or tuple(addend.shape) != (m, n):
            raise ValueError("sgemm: row_scale [m], addend [m,n]")
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hypret_sgemm_strided(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1),
                                                    m, n, k, _ptr(row_scale), _ptr(addend), _ptr(out), _stream()))
    return out


def mobius_epilogue_bwd(mx: torch.Tensor, gy: torch.Tensor, c: float, bias: Optional[torch.Tensor] = None,
                        xsq: Optional[torch.Tensor] = None, post_tanh: bool = False, n_project: int = 1):
    """Backward of one MobiusLinear epilogue (``hypret_mobius_epilogue_bwd``): ``(gmx, gbias | None, gxn | None)`` with
    ``dL/dx_in = gmx @ W + gxn[:, None] * x_in`` for a hyperbolic input."""
    _need_cuda(mx, gy, bias, xsq)
    mx, gy = mx.contiguous().float(), gy.contiguous().float()
    n, d = mx.shape
    gmx = torch.empty_like(mx)
    b = bias.detach().contiguous().float() if bias is not None else None
    gb = torch.zeros(d, dtype=torch.float32, device=mx.device) if bias is not None else None
    xs = xsq.contiguous().float() if xsq is not None else None
    gxn = torch.empty(n, dtype=torch.float32, device=mx.device) if xsq is not None else None
    with torch.cuda.device(mx.device):
        _lib.check(_lib.load().hypret_mobius_epilogue_bwd(_ptr(mx), n, d, _ptr(xs), _ptr(b), float(c),
                                                          int(bool(post_tanh)), int(n_project), _ptr(gy), _ptr(gmx),
                                                          _ptr(gb), _ptr(gxn), _stream()))
    return gmx, gb, gxn


class MobiusLinearFn(torch.autograd.Function):
    """One MobiusLinear layer, forward AND backward, as kernels of this library: forward = two operand splits + one
    ``hypret_mobius_gemm``; backward = one ``hypret_mobius_epilogue_bwd`` + the dense products dW = gmx^T x and
    dx = gmx W (+ gxn x).  Replaces the ~35 autograd nodes per layer of src/models.py:291-318."""

    @staticmethod
    def forward(ctx, x, weight, bias, c: float, hyperbolic_input: bool, post_tanh: bool, n_project: int):
        x32 = x.detach().contiguous().float()
        w32 = weight.detach().contiguous().float()
        n_out, d_in = w32.shape
        xo, xsq = split_operand(x32, "row", want_sqnorm=True)
        out = mobius_gemm(xo, split_operand(w32, "col"), d_in, n_out, c, bias=bias,
                          xsq=xsq if hyperbolic_input else None, post_tanh=post_tanh, n_project=n_project, want_y=True,
                          want_mx=True)
        ctx.save_for_backward(x32, w32, bias.detach() if bias is not None else None, out["mx"], xsq)
        ctx.cfg = (float(c), bool(hyperbolic_input), bool(post_tanh), int(n_project), x.dtype, weight.dtype,
                   bias.dtype if bias is not None else None)
        return out["y"].to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        x32, w32, bias, mx, xsq = ctx.saved_tensors
        c, hyp, post_tanh, n_project, xdt, wdt, bdt = ctx.cfg
        gmx, gb, gxn = mobius_epilogue_bwd(mx, gy, c, bias=bias, xsq=xsq if hyp else None, post_tanh=post_tanh,
                                           n_project=n_project)
        gw = sgemm(gmx.t(), x32).to(wdt) if ctx.needs_input_grad[1] else None
        gx = None
        if ctx.needs_input_grad[0]:
            gx = sgemm(gmx, w32, gxn, x32 if gxn is not None else None).to(xdt)
        return gx, gw, (gb.to(bdt) if gb is not None and ctx.needs_input_grad[2] else None), None, None, None, None


def mobius_gemm_ok(d_in: int, n_out: int) -> bool:
    return d_in % 4 == 0 and d_in <= 4096 and n_out % 16 == 0 and 16 <= n_out <= 256


def mobius_gemm(x_op: torch.Tensor, w_op: torch.Tensor, d_in: int, n_out: int, c: float,
                bias: Optional[torch.Tensor] = None, xsq: Optional[torch.Tensor] = None, post_tanh: bool = False,
                n_project: int = 1, want_y: bool = True, want_op: bool = False, want_sqnorm: bool = False,
                want_mx: bool = False):
    """One MobiusLinear layer as one kernel (``hypret_mobius_gemm``): tcgen05 GEMM of split operands + the hyperbolic
    epilogue on the accumulator.  Returns a dict with the requested outputs: ``y`` [n,n_out] fp32, ``op`` the next
    layer's row operand, ``sq`` ||y||^2, ``mx`` the raw product (kept by the training path for its backward)."""
    _need_cuda(x_op, w_op, bias, xsq)
    n = x_op.shape[0]
    dev = x_op.device
    lib = _lib.load()
    out = {
        "y": torch.empty(n, n_out, dtype=torch.float32, device=dev) if want_y else None,
        "op": torch.empty(n, int(lib.hypret_flash_kpad(n_out)), dtype=torch.float16, device=dev) if want_op else None,
        "sq": torch.empty(n, dtype=torch.float32, device=dev) if want_sqnorm else None,
        "mx": torch.empty(n, n_out, dtype=torch.float32, device=dev) if want_mx else None,
    }
    b = bias.detach().contiguous().float() if bias is not None else None
    xs = xsq.contiguous().float() if xsq is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.hypret_mobius_gemm(_ptr(x_op), _ptr(w_op), n, int(d_in), int(n_out), _ptr(xs), _ptr(b), float(c),
                                          int(bool(post_tanh)), int(n_project), _ptr(out["mx"]), _ptr(out["y"]),
                                          _ptr(out["sq"]), _ptr(out["op"]), _stream()))
    return out


def sum_parts(parts: torch.Tensor) -> torch.Tensor:
    """``parts.sum(dim=0)`` in a fixed order for per-CTA partials ``[P, n]`` (``hypret_sum_parts``)."""
    _need_cuda(parts)
    parts = parts.contiguous().float()
    out = torch.empty(parts.shape[1], dtype=torch.float32, device=parts.device)
    with torch.cuda.device(parts.device):
        _lib.check(_lib.load().hypret_sum_parts(_ptr(parts), parts.shape[0], parts.shape[1], _ptr(out), _stream()))
    return out


def lse_combine(parts: torch.Tensor) -> torch.Tensor:
    """``logsumexp(parts, dim=0)`` for per-rank partial log-sum-exps ``[W, n]`` (``hypret_lse_combine``)."""
    _need_cuda(parts)
    parts = parts.contiguous().float()
    out = torch.empty(parts.shape[1], dtype=torch.float32, device=parts.device)
    with torch.cuda.device(parts.device):
        _lib.check(_lib.load().hypret_lse_combine(_ptr(parts), parts.shape[0], parts.shape[1], _ptr(out), _stream()))
    return out


BWD_ROWS = 16       # HYPRET_BWD_ROWS: matrix rows per CTA of the backward pass (one col_partial row each)


def _bwd_chunks(n: int, m: int) -> int:
    """Column chunks of the backward pass: ~7 waves of 2 CTAs per SM out of the ceil(n/16) row blocks."""
    row_blocks = max(1, (n + BWD_ROWS - 1) // BWD_ROWS)
    return max(1, min((m + 255) // 256, round(7 * 2 * 148 / row_blocks)))


def _w_buffer(n: int, m: int, split: bool, device):
    if split:
        return torch.empty(3, n, m, dtype=torch.bfloat16, device=device)
    return torch.empty(n, m, dtype=torch.float32, device=device)


def split3(x: torch.Tensor) -> torch.Tensor:
    """fp32 ``[...]`` -> bf16 planes ``[3, ...]`` with hi + mid + lo = x to fp32 accuracy (``hypret_split3``)."""
    _need_cuda(x)
    x = x.contiguous().float()
    if x.numel() % 4:
        raise ValueError("split3 needs a multiple of 4 elements")
    out = torch.empty((3,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hypret_split3(_ptr(x), x.numel(), _ptr(out), _stream()))
    return out


class SplitW:
    """W of the distance-matrix backward as three bf16 planes (hi + mid + lo = W): ``data`` is ``[3,n,m]``, cut from
    the fp32 W by ``hypret_split3``, or written by the backward kernel itself (``w_format`` 1) when n*m is not a
    multiple of 4."""

    def __init__(self, data: torch.Tensor):
        self.data = data

    def float(self) -> torch.Tensor:
        return self.data.float().sum(dim=0)


def _two_pass_split(n: int, m: int) -> bool:
    """Planes cut from an fp32 W by ``split3`` (faster than writing them from inside the backward pass) unless the
    element count is not a multiple of 4."""
    return (n * m) % 4 == 0


def pairdist_ce_bwd(dmat: torch.Tensor, asq: torch.Tensor, psq: torch.Tensor, c: float, row_lse: torch.Tensor,
                    col_lse: Optional[torch.Tensor], inv_tau: float, w_rows: float, w_cols: float,
                    grad_scale: Optional[torch.Tensor] = None, split: bool = False, diag_offset: int = 0,
                    n_total: Optional[int] = None):
    """Backward weights of the in-batch InfoNCE: ``(W, row_sum [n], col_sum [m])``; the upstream gradient is
    formed inside the kernel from the log-sum-exps (``grad_scale``: device scalar dL/dloss).  ``W`` is ``[n,m]``
    fp32, or with ``split`` three bf16 planes ``[3,n,m]`` (hi + mid + lo = W) for ``split_products``.
    ``diag_offset`` / ``n_total``: this is the row block ``[diag_offset, diag_offset + n)`` of an ``n_total``-row
    batch whose negatives are sharded across ranks (``train.ShardedInBatchInfoNCE``)."""
    _need_cuda(dmat, asq, psq, row_lse, col_lse, grad_scale)
    n, m = dmat.shape
    two_pass = split and _two_pass_split(n, m)
    if two_pass:
        split = False
    w = _w_buffer(n, m, split, dmat.device)
    n_rp = _bwd_chunks(n, m)
    rp = torch.empty(n_rp, n, dtype=torch.float32, device=dmat.device)
    cp = torch.empty((n + BWD_ROWS - 1) // BWD_ROWS, m, dtype=torch.float32, device=dmat.device)
    gs = grad_scale.reshape(1).contiguous().float() if grad_scale is not None else None
    with torch.cuda.device(dmat.device):
        _lib.check(_lib.load().hypret_pairdist_ce_bwd(_ptr(dmat.contiguous()), _ptr(asq.contiguous().float()),
                                                      _ptr(psq.contiguous().float()), n, m, float(c), _ptr(row_lse),
                                                      _ptr(col_lse), float(inv_tau), float(w_rows), float(w_cols),
                                                      _ptr(gs), _ptr(w), int(split), _ptr(rp), n_rp, _ptr(cp),
                                                      int(diag_offset), int(n if n_total is None else n_total),
                                                      _stream()))
    if two_pass:
        w = SplitW(split3(w))
    elif split:
        w = SplitW(w)
    return w, sum_parts(rp), sum_parts(cp)


def pairdist_bwd(grad_out: torch.Tensor, dmat: torch.Tensor, asq: torch.Tensor, psq: torch.Tensor, c: float,
                 n_partial: Optional[int] = None, split: bool = False):
    """Weights of the distance-matrix backward: returns ``(W, row_sum [n], col_sum [m])`` (``split`` as in
    ``pairdist_ce_bwd``)."""
    _need_cuda(grad_out, dmat, asq, psq)
    grad_out = grad_out.contiguous().float()
    dmat = dmat.contiguous()
    n, m = dmat.shape
    n_partial = max((n + BWD_ROWS - 1) // BWD_ROWS, n_partial or 0, 1)
    two_pass = split and _two_pass_split(n, m)
    if two_pass:
        split = False
    w = _w_buffer(n, m, split, dmat.device)
    n_rp = _bwd_chunks(n, m)
    rp = torch.empty(n_rp, n, dtype=torch.float32, device=dmat.device)
    cp = torch.empty(n_partial, m, dtype=torch.float32, device=dmat.device)
    with torch.cuda.device(dmat.device):
        _lib.check(_lib.load().hypret_pairdist_bwd(_ptr(grad_out), _ptr(dmat), _ptr(asq.contiguous()),
                                                   _ptr(psq.contiguous()), n, m, float(c), _ptr(w), int(split), _ptr(rp),
                                                   n_rp, _ptr(cp), n_partial, _stream()))
    if two_pass:
        w = SplitW(split3(w))
    elif split:
        w = SplitW(w)
    return w, sum_parts(rp), sum_parts(cp)


SPLIT_MIN_PAIRS = 1 << 20     # below this the two products are launch-bound either way: plain fp32 GEMMs


def _split3(x: torch.Tensor) -> torch.Tensor:
    """[r,d] fp32 -> [r,3d] bf16 = [hi | mid | lo] with hi + mid + lo = x to fp32 accuracy."""
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return torch.cat([hi, mid, lo], dim=1)


def _split_parts(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def split_products(w: SplitW, a: torch.Tensor, p: torch.Tensor):
    """``(W @ p, W.T @ a)`` for ``W = hi + mid + lo`` (bf16 planes) at fp32-GEMM accuracy on the tensor cores: the six
    bf16 cross products of order <= 2^-16, hi x (hi|mid|lo) + mid x (hi|mid) + lo x hi, as library GEMMs with fp32
    accumulation (the plain dense products of the backward; cuBLAS SGEMM before): three GEMMs per product.
    The non-flash path only (D not served by ``flash_ok``); the flash backward does these products on-chip."""
    d = a.shape[1]
    a, p = a.float(), p.float()
    w3 = w.data
    outs = []
    for planes, x in (((w3[0], w3[1], w3[2]), p), ((w3[0].t(), w3[1].t(), w3[2].t()), a)):
        x3 = torch.cat(_split_parts(x), dim=1)
        o = torch.mm(planes[0], x3, out_dtype=torch.float32)
        acc = o[:, 2 * d:] + o[:, d:2 * d]                       # smallest terms first
        o1 = torch.mm(planes[1], x3[:, :2 * d], out_dtype=torch.float32)
        acc = acc + o1[:, d:] + torch.mm(planes[2], x3[:, :d], out_dtype=torch.float32)
        outs.append(acc + o1[:, :d] + o[:, :d])
    return outs[0], outs[1]


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    """||x_i||^2 per row, fp32 (the sq output of ``hypret_flash_prep``)."""
    _need_cuda(x)
    x = x.detach().contiguous().float()
    n, d = x.shape
    if d % 4 or d > 4096:
        raise ValueError("row_sqnorm: D % 4 == 0, D <= 4096")
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hypret_flash_prep(_ptr(x), n, d, None, None, None, 0, _ptr(out), _stream()))
    return out


def mobius_epilogue(mx: torch.Tensor, c: float, bias: Optional[torch.Tensor] = None,
                    xsq: Optional[torch.Tensor] = None, post_tanh: bool = False, n_project: int = 1,
                    want_sqnorm: bool = True):
    """Everything a MobiusLinear layer does after its GEMM, in one kernel (inference path).
    ``xsq`` given -> hyperbolic input (mobius_matvec rescale), else expmap0.  Returns ``(y, ||y||^2 | None)``."""
    _need_cuda(mx, bias, xsq)
    mx = mx.contiguous().float()
    n, d = mx.shape
    y = torch.empty_like(mx)
    sq = torch.empty(n, dtype=torch.float32, device=mx.device) if want_sqnorm else None
    b = bias.detach().contiguous().float() if bias is not None else None
    xs = xsq.contiguous().float() if xsq is not None else None
    with torch.cuda.device(mx.device):
        _lib.check(_lib.load().hypret_mobius_epilogue(_ptr(mx), n, d, _ptr(xs), _ptr(b), float(c), int(xs is not None),
                                                      int(bool(post_tanh)), int(n_project), _ptr(y), _ptr(sq),
                                                      _stream()))
    return y, sq
