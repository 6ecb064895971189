"""Gallery index + search: the user-facing call of the retrieval hot path.

Replaces, in the reference:
  * the per-query loop ``pmath.dist(q[1,D], patents[P,D])`` + ranking of
    ``evaluate_retrieval``                         (src/train.py:3221-3293)
  * ``cosine_similarity(Q, G)`` + ``np.argsort``   (notebooks/retrieval.ipynb:368-383)
  * ``ImageRetrieval.retrieve_similar_images``     (notebooks/retrieval.ipynb:190-206)

``GalleryIndex`` keeps, resident in HBM, the fp32 gallery rows (exact-rerank operand) and
the fp16 tensor-core operand built by the fused projection kernel.  ``search`` runs
projection(queries) -> tcgen05 scoring + streaming top-k' -> exact rerank.  All compute is
in libhypret.so; torch only owns the buffers.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import ops


class StageEvents:
    """CUDA-event pairs around the kernels of a search, recorded on the launching stream (bench.py's roofline
    measurements; "certify" = the merged-list certificate + flagged rescans of the sharded path).  ``ms(name)`` = mean
    duration after a synchronize."""

    STAGES = ("project", "score", "rerank", "certify")

    def __init__(self):
        self.pairs = {s: [] for s in self.STAGES}

    @contextlib.contextmanager
    def span(self, name):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self.pairs[name].append((e0, e1))

    def ms(self, name):
        p = self.pairs[name]
        return sum(a.elapsed_time(b) for a, b in p) / len(p) if p else float("nan")


@contextlib.contextmanager
def _span(events, name):
    """``events``: None, a list (receives the scoring kernel's pair only) or a ``StageEvents``."""
    if isinstance(events, StageEvents):
        with events.span(name):
            yield
    elif events is not None and name == "score":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        events.append((e0, e1))
    else:
        yield


WIDE_MIN_LISTS = 4     # wide top-k: candidate lists per query (4 x 64 = 256 candidates for k <= 128)
WIDE_SWITCH_RATE = 0.02        # adaptive list width: share of uncertified queries above which the wide lists are cheaper
WIDE_SWITCH_MIN_ROWS = 1024    # ... only for galleries where a full scan per query costs more than the wide pass
WIDE_PROBE_EVERY = 32          # ... and how often the narrow path is tried again


def default_kprime(k: int) -> int:
    """List slots per strip.  k <= 10: 16-slot lists (the register-list epilogue of the scoring kernel, query tile
    resident in shared memory); k <= 26: one list of k+14 slots (<= 32; threshold sharing keeps the union equal to the
    global approximate top-k').  Larger k ("wide", up to 128): 64-slot lists, at least ``WIDE_MIN_LISTS`` independent
    lists per query, no threshold sharing -- the top-k is in the union unless one strip holds more than 64 of the k
    best rows, which ``margin`` certifies per query."""
    if k > ops.MAX_K:
        raise ValueError(f"k={k} exceeds the supported maximum {ops.MAX_K}")
    if k <= 10:
        return 16
    return min(32, k + 14) if k <= 26 else 64


def default_kbound(k: int, kprime: int) -> int:
    """Candidates the exact-top-k certificate works with: k+14 (<= 32).  The certificate needs the k'-th..k-th best
    candidates to be further apart than the rounding bound E; with k+6 candidates about 1.5e-4 of the (query,
    1M..10M-row shard) pairs fail that and fall to the exact scan (4.5 ms for a few queries over 5M rows), with k+14 the
    1 % margin quantile is 8 E (tools/diag_cert.py).  16-slot lists reach k+14 by sharing the bound of the (k+14)-th
    best score (``ops.score_topk(kbound=...)``): measured 10 % faster at C4 and 20 % at C2 than 24-slot lists in shared
    memory, which lose the resident query tile."""
    return max(kprime, min(32, k + 14)) if kprime <= 16 else kprime


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
        # gallery maxima for the exact-top-k certificate (csrc/project.cu): filled by the projection pass itself
        self.stats = torch.zeros(4, dtype=torch.float32, device=device)
        if metric == "hyperbolic":
            mode = "expmap0" if space == "euclidean" else "onball"
            self.rows32, self.operand, _ = ops.project_rows(feats, c, mode=mode, side="gallery", stats=self.stats)
        else:
            self.rows32 = feats.contiguous()
            _, self.operand, _ = ops.project_rows(feats, 1.0, mode="cosine", side="gallery", want_point=False,
                                                  stats=self.stats)
        self.rows_sq64 = ops.row_sqnorm64(self.rows32)       # fp64 row norms: exact scan, wide exact rerank (8 B per row)
        self._cand = {}
        self._cert = None
        self.certificate = None     # CertBuffers of the last exact search (device-side; see ops.rerank_cert)
        self.uncertified_wide = None   # [Q] int32 flags of the last exact wide search (26 < k <= 128, or wide mode)
        # adaptive list width (search(kprime=None, k <= 26)): the share of queries the narrow certificate could not
        # prove, mirrored to pinned host memory by an asynchronous copy and read WITHOUT synchronising
        self.adaptive = True
        self.fallback_rate = 0.0       # of the last narrow search whose count has arrived
        self.last_mode = "narrow"
        self._h_count = None
        self._pending = None           # (event, n_queries) of the copy in flight
        self._since_probe = 0

    def _query_mode(self):
        if self.metric == "cosine":
            return "cosine"
        return "expmap0" if self.space == "euclidean" else "onball"

    def search(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None, return_margin: bool = False,
               max_ctas: int = 0, kernel_events: Optional[list] = None, exact: bool = True,
               kbound: Optional[int] = None):
        """queries [Q,D] fp32 (host or device) -> (score [Q,k] f32, idx [Q,k] i64) on the device.
        score = Poincare distance ascending, or cosine similarity descending; ties -> lower index.

        ``exact`` (default): the result is GUARANTEED to be the exact top-k (by exact distance; equal fp32 distances
        in index order).  The fp16 tensor-core pass is only a filter: per query the rerank kernel proves, from the
        rounding residuals of the operands, that no row outside the candidate set can precede the k-th result, and the
        queries it cannot prove (near-duplicate galleries) are recomputed by a full exact scan queued on the same
        stream -- no host synchronisation either way.  ``self.certificate`` (``ops.CertBuffers``) holds the per-query
        flags and the count of rescanned queries (k <= 26; for 26 < k <= 128 ``self.uncertified_wide`` holds the flags
        of the queries that were paged through the exact ranking instead).  ``kbound``: see ``default_kbound`` (None = default; = k' for plain k'-slot semantics).
        ``kernel_events``: a list (receives a (start, end) CUDA-event pair bracketing the scoring kernel
        on the launching stream per call) or a ``StageEvents`` (all three kernels): bench.py's rooflines."""
        if queries.shape[0] == 0:                 # an empty batch is an empty result (sklearn / torch.topk semantics)
            empty = (torch.empty(0, k, dtype=torch.float32, device=self.device),
                     torch.empty(0, k, dtype=torch.int64, device=self.device))
            return empty + (torch.empty(0, dtype=torch.float32, device=self.device),) if return_margin else empty
        if k > ops.MAX_K:
            # beyond the filtered path: page through the exact ranking by full scans (the reference ranks the whole
            # gallery, notebooks/retrieval.ipynb:383; its metrics are served by rank counting, this serves its lists)
            q = queries.to(device=self.device, dtype=torch.float32, non_blocking=True)
            if self.metric == "hyperbolic":
                q = ops.project_rows(q, self.c, mode=self._query_mode(), side="query", want_operand=False)[0]
            res = ops.exact_topk_any(q, self.rows32, self.rows_sq64, self.c, self.metric, min(k, self.n),
                                     idx_offset=self.idx_offset)
            return res + (torch.full((q.shape[0],), float("inf"), device=self.device),) if return_margin else res
        auto = (kprime is None and exact and self.adaptive and k <= 26 and self.n > 4 * WIDE_SWITCH_MIN_ROWS
                and not torch.cuda.is_current_stream_capturing())
        if auto and self._wide_mode():
            if self._since_probe == 0:     # now and then: the narrow filter + certificate alone (no scans), for the rate
                pq32, pcs, pci, pcnt, perr = self.score_candidates(queries, k=k, want_err=True,
                                                                   kbound=default_kbound(k, default_kprime(k)))
                self.rerank_candidates(pq32, pcs, pci, k, list_count=pcnt, q_err=perr,
                                       ksel=default_kbound(k, default_kprime(k)), fallback=False)
                self._mirror_fallback_count(pq32.shape[0])
            kprime = 64                # tight classes: 64-slot lists + 256 exactly rescored survivors certify them
        self.last_mode = "wide" if (kprime is not None and int(kprime) > 32) else "narrow"
        kp = min(default_kprime(k) if kprime is None else int(kprime), ops.MAX_KPRIME)
        if kbound is None:
            kbound = default_kbound(k, kp) if (exact and k <= kp) else kp
        q32, cs, ci, cnt, q_err = self.score_candidates(queries, k=k, kprime=kp, max_ctas=max_ctas,
                                                        kernel_events=kernel_events, want_err=exact, kbound=kbound)
        if exact and (k > kp or k > 32 or kp > 32):
            # wide top-k (26 < k <= 128): the wide rerank's margin (smallest filter score a non-candidate can have - exact
            # surrogate of the k-th result) against the same rounding bound E; queries it does not clear are paged
            # through the exact ranking by full scans (device list, no host round trip) and overwrite their rows
            score, idx, margin = self.rerank_candidates(q32, cs, ci, k, return_margin=True, kernel_events=kernel_events,
                                                        list_count=cnt)
            bound = ops.certificate_bound(q32, q_err, self.stats, self.c, self.metric, self.d)
            flags = (~(margin > bound)).to(torch.int32)
            xs, xi = ops.exact_topk_any_flagged(q32, self.rows32, self.rows_sq64, flags, self.c, self.metric,
                                                min(k, self.n), idx_offset=self.idx_offset)
            if xs.shape[1] < k:                   # fewer gallery rows than k: the tail reads as empty, like the rerank's
                pad = k - xs.shape[1]
                fill = float("inf") if self.metric == "hyperbolic" else float("-inf")
                xs = torch.nn.functional.pad(xs, (0, pad), value=fill)
                xi = torch.nn.functional.pad(xi, (0, pad), value=-1)
            redo = flags.bool()[:, None]
            self.uncertified_wide = flags
            score, idx = torch.where(redo, xs, score), torch.where(redo, xi, idx)
            return (score, idx, margin) if return_margin else (score, idx)
        out = self.rerank_candidates(q32, cs, ci, k, return_margin=return_margin, kernel_events=kernel_events,
                                     list_count=cnt, q_err=q_err if exact else None, ksel=kbound)
        if auto and self.certificate is not None:
            self._mirror_fallback_count(q32.shape[0])
        return out

    # ---- adaptive list width -------------------------------------------------------------------------------------
    # The certificate needs the k-th .. k_b-th candidates to be further apart than the rounding bound E.  On galleries of
    # tight classes (the reference's figures of one patent family) a class of more than k_b - k near-equidistant rows
    # defeats it, and every such query costs a full exact scan: 10k queries x 300k rows, 30 rows per class at 10 % noise:
    # 21 % uncertified, 112 ms per search instead of 2.3.  The wide path (64-slot lists, 256 survivors rescored exactly,
    # the same bound) proves all of them in 9.6 ms (tools/bench_clustered.py).  So the index watches the share of
    # uncertified queries -- an asynchronous 4-byte copy per search, read one or more searches later, never waited
    # for -- and serves a gallery that keeps defeating the narrow certificate with the wide lists, probing the narrow
    # certificate (filter + proof only, no scans: +1 narrow pass) every WIDE_PROBE_EVERY searches.  The results are exact
    # either way; only the cost moves.
    def _mirror_fallback_count(self, n_queries: int) -> None:
        if self._h_count is None:
            self._h_count = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        if self._pending is not None and not self._pending[0].query():
            return                                 # the previous copy is still in flight: keep its buffer untouched
        self._consume_pending()
        self._h_count.copy_(self.certificate.count, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending = (ev, int(n_queries))

    def _consume_pending(self) -> None:
        if self._pending is not None and self._pending[0].query():
            self.fallback_rate = float(self._h_count[0]) / max(self._pending[1], 1)
            self._pending = None

    def _wide_mode(self) -> bool:
        self._consume_pending()
        if self.fallback_rate <= WIDE_SWITCH_RATE:
            self._since_probe = 0
            return False
        self._since_probe = (self._since_probe + 1) % WIDE_PROBE_EVERY     # 0: this call also probes the narrow path
        return True

    def score_candidates(self, queries: torch.Tensor, k: int = 10, kprime: Optional[int] = None, max_ctas: int = 0,
                         kernel_events: Optional[list] = None, want_err: bool = False, kbound: Optional[int] = None):
        """First half of ``search``: projection + tcgen05 scoring / streaming top-k'.
        Returns ``(q32 [Q,D] exact-rerank operand, cand_score [Q,L,k'], cand_idx [Q,L,k'] int32, list_count [Q]
        int32, q_err [Q] | None)``: the lists of query q are its first ``list_count[q]`` slots (compact, arrival
        order); the other slots hold stale data; ``q_err`` (``want_err``) = rounding-residual norms of the query
        operand rows for the certificate.  The candidate buffers are reused by the next call."""
        q = queries.to(device=self.device, dtype=torch.float32, non_blocking=True)
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be [Q, {self.d}]")
        with _span(kernel_events, "project"):
            if self.metric == "hyperbolic":
                res = ops.project_rows(q, self.c, mode=self._query_mode(), side="query", want_err=want_err)
                q32, q_op = res[0], res[1]
            else:
                q32 = q.contiguous()
                res = ops.project_rows(q, 1.0, mode="cosine", side="query", want_point=False, want_err=want_err)
                q_op = res[1]
        q_err = res[3] if want_err else None
        return self.score_projected(q32, q_op, k=k, kprime=kprime, max_ctas=max_ctas,
                                    kernel_events=kernel_events, kbound=kbound) + (q_err,)

    def score_projected(self, q32: torch.Tensor, q_op: torch.Tensor, k: int = 10, kprime: Optional[int] = None,
                        max_ctas: int = 0, kernel_events: Optional[list] = None, kbound: Optional[int] = None):
        """``score_candidates`` for queries that are already projected: ``q32`` the exact-rerank rows, ``q_op`` their
        fp16 operand rows (``ops.project_rows`` / the peer exchange of ``dist.PeerQueryExchange``)."""
        kprime = default_kprime(k) if kprime is None else int(kprime)
        kprime = min(kprime, ops.MAX_KPRIME)
        wide = k > kprime or k > 32 or kprime > 32
        min_lists = WIDE_MIN_LISTS if wide else 0
        n_q = q_op.shape[0]
        key = (n_q, kprime, max_ctas, min_lists)
        plan = ops.score_plan(n_q, self.n, self.d, kprime, max_ctas, min_lists)
        if k > plan["n_lists"] * kprime:
            raise ValueError("k exceeds the number of candidates the plan can hold")
        buf = self._cand.get(key)
        if buf is None:
            self._cand.clear()
            buf = (torch.empty(n_q, plan["n_lists"], kprime, dtype=torch.float32, device=self.device),
                   torch.empty(n_q, plan["n_lists"], kprime, dtype=torch.int32, device=self.device),
                   torch.empty(n_q, dtype=torch.int32, device=self.device),
                   torch.empty(n_q, dtype=torch.int32, device=self.device))
            self._cand[key] = buf
        with _span(kernel_events, "score"):
            cs, ci = ops.score_topk(q_op, self.operand, self.d, kprime, max_ctas, out=buf[:2], thr_workspace=buf[2],
                                    share_thresholds=not wide, min_lists=min_lists, list_count=buf[3],
                                    kbound=None if wide else kbound)
        return q32, cs, ci, buf[3]

    def rerank_candidates(self, q32, cand_score, cand_idx, k: int, return_margin: bool = False,
                          prune_thr: Optional[torch.Tensor] = None, kernel_events: Optional[list] = None,
                          list_count: Optional[torch.Tensor] = None, q_err: Optional[torch.Tensor] = None,
                          ksel: int = 0, fallback: bool = True):
        """Second half of ``search``: candidate merge + exact rerank against this shard's fp32 rows.  With ``q_err``
        (and k <= k' <= 32, no pruning): the certified rerank + exact-scan fallback of ``ops.rerank_cert``."""
        kprime = cand_score.shape[2]
        if q_err is not None and prune_thr is None and k <= kprime <= 32:
            n_q = q32.shape[0]
            if self._cert is None or self._cert.Q < n_q:
                self._cert = ops.CertBuffers(n_q, self.device)
            self.certificate = self._cert
            with _span(kernel_events, "rerank"):
                return ops.rerank_cert(q32, self.rows32, cand_score, cand_idx, self.c, self.metric, k, q_err,
                                       self.stats, self.rows_sq64, self._cert, idx_offset=self.idx_offset,
                                       want_margin=return_margin, list_count=list_count,
                                       ksel=ksel if ksel > kprime else 0, fallback=fallback)
        with _span(kernel_events, "rerank"):
            out = ops.rerank(q32, self.rows32, cand_score, cand_idx, self.c, self.metric, k,
                             idx_offset=self.idx_offset, want_margin=return_margin, prune_thr=prune_thr,
                             list_count=list_count, g_sqnorm64=self.rows_sq64 if prune_thr is None else None)
        return out


class SearchPipeline:
    """Host-buffer serving loop: pinned-host query batches in, pinned-host ``[Q,k]`` results out,
    with the H2D copy of batch i+1 and the D2H copy of batch i-1 overlapping the search of batch i
    (three streams, double-buffered device and host buffers).  ``index`` is a ``GalleryIndex`` or a
    ``dist.ShardedGalleryIndex``.

        pipe = SearchPipeline(index, Q, k)
        slot = pipe.submit(q_host_pinned)          # asynchronous
        score, idx = pipe.result(slot)             # blocks until that batch's results are on the host
    """

    def __init__(self, index, n_queries: int, k: int = 10, kprime: Optional[int] = None, depth: int = 2):
        local = getattr(index, "local", index)
        self.index = index
        self.k, self.kprime, self.depth = k, kprime, depth
        self.device = local.device
        self.copy_in = torch.cuda.Stream(device=self.device)
        self.compute = torch.cuda.Stream(device=self.device)
        self.copy_out = torch.cuda.Stream(device=self.device)
        self.q_dev = [torch.empty(n_queries, local.d, dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.out_s = [torch.empty(n_queries, k, dtype=torch.float32, pin_memory=True) for _ in range(depth)]
        self.out_i = [torch.empty(n_queries, k, dtype=torch.int64, pin_memory=True) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0
        self.bytes_in = n_queries * local.d * 4
        self.bytes_out = n_queries * k * 12

    def submit(self, q_host: torch.Tensor, kernel_events: Optional[list] = None) -> int:
        slot = self.n % self.depth
        if self.n >= self.depth:
            # the device query buffer of this slot was last read by the search submitted `depth` steps ago,
            # and its host result buffers by the caller: both must be done
            self.copy_in.wait_event(self.ev_done[slot])
            self.ev_out[slot].synchronize()
        with torch.cuda.stream(self.copy_in):
            self.q_dev[slot].copy_(q_host, non_blocking=True)
            self.ev_in[slot].record(self.copy_in)
        self.compute.wait_event(self.ev_in[slot])
        with torch.cuda.stream(self.compute):
            kw = {"kernel_events": kernel_events} if kernel_events is not None else {}
            score, idx = self.index.search(self.q_dev[slot], k=self.k, kprime=self.kprime, **kw)
            self.ev_done[slot].record(self.compute)
        score.record_stream(self.copy_out)
        idx.record_stream(self.copy_out)
        self.copy_out.wait_event(self.ev_done[slot])
        with torch.cuda.stream(self.copy_out):
            self.out_s[slot].copy_(score, non_blocking=True)
            self.out_i[slot].copy_(idx, non_blocking=True)
            self.ev_out[slot].record(self.copy_out)
        self.n += 1
        return slot

    def result(self, slot: int):
        self.ev_out[slot].synchronize()
        failed = getattr(self.index, "exchange_failed", None)
        if failed is not None and failed():      # sharded index: a peer-exchange wait timed out, the lists are poisoned
            raise RuntimeError("peer exchange timed out during this batch (a rank died or fell out of step)")
        return self.out_s[slot], self.out_i[slot]

    def drain(self):
        for ev in self.ev_out:
            ev.synchronize()
