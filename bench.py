#!/usr/bin/env python
"""bench.py -- headline benchmark of the retrieval hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl native|reference]

A *step* = one pass of the hot path over one batch of synthetic queries:
    project(queries) -> tcgen05 scoring + streaming top-k' -> exact rerank [-> all_gather + merge]
against a gallery index that is resident in HBM (built once, untimed, like the reference's
``load_embeddings()`` cache, notebooks/retrieval.ipynb:155-163).

* ``value``  queries/s, raw query features already resident in HBM (CUDA events, max over ranks)
* ``e2e``    the same through the public API with HOST buffers (``SearchPipeline``): every step
             copies that step's pinned-host queries H2D, searches, and copies the [Q,k] result
             (distances + indices) D2H; copies of neighbouring steps overlap the search (three
             streams, three buffer slots).  ``e2e.serial`` is the same without any overlap
* ``roofline``  scoring kernel only: 2*Q*N_local*D algorithmic flops / its CUDA-event duration,
             against the measured bf16 tensor peak in MEASURED_PEAKS.json
* ``cpu_baseline``  the oracle (reference torch-fp32 path restated, oracle/) timed on this
             box's host cores on a bounded sample of the same workload (rank 0, N=1 only)

Multi-GPU (torchrun, one rank per GPU), gallery row-sharded across the ranks in both modes:
* ``--scaling weak`` (default): the serving layout.  Every rank is fed its OWN batch of Q queries
  per step (W*Q queries per step in total, per-GPU work = Q x N pairs whatever W is); batches are
  all-gathered over NVLink, each rank scores all W*Q queries against its N/W-row shard, and an
  all_to_all returns the per-shard [Q,k] lists to the query owners, which merge them.
  ``value`` = W*Q / step time.
* ``--scaling strong``: the SAME global workload (Q queries, replicated), [Q,k] candidate lists
  all-gathered and merged on every rank (BASELINE config 4 = ``--workload c4 --scaling strong``).
``--impl reference`` times the CPU oracle alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (Q, N, D, k, c, description)
    "c1": (1000, 10_000, 2048, 10, 1.0, "C1: 1k queries x 10k gallery, D=2048, top-10, c=1"),
    "c2": (10_000, 300_000, 512, 10, 1.0, "C2: 10k queries x 300k gallery, D=512, top-10, c=1"),
    "c4": (10_000, 10_000_000, 512, 10, 1.0, "C4: 10k queries x 10M gallery, D=512, top-10, c=1"),
    "c5": (8192, 8192, 128, 0, 0.5,
           "C5: train_hyp in-batch InfoNCE over the 8192 x 8192 Poincare distance matrix, forward + backward, D=128 "
           "(src/train.py:4009), tau=0.07, c=0.5"),
    "c3": (100_000, 1_000_000, 768, 100, 1.0,
           "C3: 100k queries x 1M gallery, D=768, top-100 under BOTH metrics (cosine + hyperbolic), c=1"),
}
METRIC = "queries/sec at top-10 over N-gallery"
UNIT = "queries/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)),
                "tflops_burst": float(d.get("bf16_tflops", 1590.0)), "hbm_gbs": float(d.get("hbm_gbs", 6650.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_info():
    model = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"os_cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cpu_model": model}


def default_scaling(workload: str) -> str:
    """C4 is the 10M-row gallery north_star's scaling target is written for: the SAME global workload at every GPU
    count (strong).  The small-gallery workloads scale as a serving loop (each rank fed its own batch: weak)."""
    return "strong" if workload == "c4" else "weak"


def search_config(workload: str, world: int, scaling: str) -> dict:
    """The workload description -- identical in the native and the reference arm (same keys, same seeds)."""
    Q, N, D, k, c, desc = WORKLOADS[workload]
    weak = world > 1 and scaling == "weak"
    return {"workload": desc, "Q": Q, "N": N, "D": D, "k": k, "c": c, "scaling": scaling,
            "queries_per_step_total": Q * world if weak else Q,
            "gallery": "synth.gallery_rows: N(0,(0.45/sqrt(D))^2), 2^18-row blocks seeded 0+7919*(block+1) -- the same "
                       "gallery at every GPU count, row-sharded",
            "queries": "synth.gaussian_features seed 1" + (" + 100*rank (each rank its own batch)" if weak else ""),
            "cache": "inputs larger than L2 (fp16 gallery operand %.0f MB per GPU vs 126 MB L2); no flush" %
                     (-(-N // world) * (D + 16 + (-D) % 64) * 2 / 1e6)}


def cpu_oracle_topk(q_u, g_pts, c, k):
    """The reference's path restated (oracle/): embed the raw query features, per-query pmath.dist over the gallery
    points (src/train.py:3259), top-k (src/auxiliary.py:374)."""
    from oracle import head, retrieval
    return retrieval.hyperbolic_topk(head.embed_rows(q_u, c), g_pts, c, k, form="geoopt")


def time_cpu_baseline(q_u_cpu, g_pts_cpu, c, k, budget_s=15.0, fraction=1.0):
    """Queries/s of the oracle on a bounded sample (gallery points pre-embedded: the index is resident for the CPU
    arm too).  ``fraction`` < 1: ``g_pts_cpu`` is that fraction of the gallery's rows; the path is a linear scan per
    query, so queries/s over the whole gallery = fraction * (queries/s over the slice).
    Returns (qps, n_queries, seconds, (dist, idx))."""
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    cpu_oracle_topk(q_u_cpu[:1], g_pts_cpu, c, k)
    one = time.perf_counter() - t0
    n = int(max(2, min(64, budget_s / max(one, 1e-3))))
    n = min(n, q_u_cpu.shape[0])
    t0 = time.perf_counter()
    res = cpu_oracle_topk(q_u_cpu[:n], g_pts_cpu, c, k)
    dt = time.perf_counter() - t0
    return fraction * n / dt, n, dt, res


def oracle_parity(rows32_dev, row_lo, q_u_dev, c, k, gpu_d, gpu_i, world, rank, dist, chunk_rows=1 << 20, n_keep=256):
    """Driver-visible parity at EVERY GPU count and at full gallery size.  The first rows of this
    step's batch are answered by the CPU oracle -- per-query ``pmath.dist(q[1,D], G)`` in the reference's fp32
    geoopt form + top-k (src/train.py:3259, src/auxiliary.py:374) -- and compared with the GPU lists.

    Running the oracle over all N rows costs ~13 s per query at N = 10M, so it is evaluated on the ``n_keep`` rows per
    query that a coarse CPU prefilter keeps (fp32 GEMM squared distances, >= 25x more rows than k).  ``pmath.dist``
    treats gallery rows independently, so the kept rows get bit-identical oracle distances, and the oracle top-k
    over the whole gallery equals its top-k over the kept rows as long as the kept set covers it -- asserted through
    the margin between the oracle's k-th distance and the worst kept row's.  Every rank does this for its own shard
    on its share of the host cores; rank 0 merges the per-shard oracle lists (top-k of a union)."""
    from oracle import head, pmath, retrieval
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    n_par = q_u_dev.shape[0]
    t0 = time.perf_counter()
    q_pts = head.embed_rows(q_u_dev.cpu(), c)                        # the reference's own embedding of the raw queries
    n_local = rows32_dev.shape[0]
    stage = torch.empty(min(chunk_rows, n_local), rows32_dev.shape[1], dtype=torch.float32, pin_memory=True)
    keep_v, keep_i, keep_rows = [], [], []
    for r0 in range(0, n_local, chunk_rows):
        r1 = min(n_local, r0 + chunk_rows)
        g = stage[:r1 - r0]
        g.copy_(rows32_dev[r0:r1])
        d2 = g.pow(2).sum(1)[None, :] - 2.0 * (q_pts @ g.t())        # + ||q||^2: constant per query
        kk = min(64, r1 - r0)
        v, i = torch.topk(d2, kk, dim=1, largest=False)
        keep_v.append(v)
        keep_i.append(i + r0)
        keep_rows.append(g[i])                                       # [n_par, kk, D]
    v, i, rows = torch.cat(keep_v, 1), torch.cat(keep_i, 1), torch.cat(keep_rows, 1)
    kk = min(n_keep, v.shape[1])
    v, sel = torch.topk(v, kk, dim=1, largest=False)
    i = torch.gather(i, 1, sel)
    rows = torch.gather(rows, 1, sel[:, :, None].expand(-1, -1, rows.shape[2]))
    order = torch.argsort(i, dim=1)                                  # ascending row id: ties -> lower index
    i = torch.gather(i, 1, order)
    rows = torch.gather(rows, 1, order[:, :, None].expand(-1, -1, rows.shape[2]))
    kt = torch.tensor(-float(c))
    d32 = torch.stack([pmath.dist(q_pts[q:q + 1], rows[q], k=kt) for q in range(n_par)])          # [n_par, kk] fp32
    d64 = torch.stack([retrieval.hyperbolic_dist_rows(q_pts[q:q + 1].double(), rows[q].double(), c, form="arcosh")[0]
                       for q in range(n_par)])
    kq = min(k, kk)
    o32_v, o32_j = retrieval.topk_smallest(d32, kq)
    o64_v, o64_j = retrieval.topk_smallest(d64, kq)
    worst_kept = d64.max(dim=1).values
    lists = {"d32": o32_v, "i32": torch.gather(i, 1, o32_j) + row_lo, "d64": o64_v, "i64": torch.gather(i, 1, o64_j) + row_lo,
             "worst": worst_kept}
    if world > 1:
        dev = rows32_dev.device
        out = {}
        for name, t in lists.items():
            t = t.to(dev).contiguous()
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            out[name] = torch.stack(g).cpu()                          # [W, n_par, ...]
        if rank != 0:
            return None
        def merge(dv, iv):
            dv = dv.permute(1, 0, 2).reshape(n_par, -1)
            iv = iv.permute(1, 0, 2).reshape(n_par, -1)
            o = torch.from_numpy(__import__("numpy").lexsort((iv.numpy(), dv.numpy()), axis=1)[:, :k].copy())
            return torch.gather(dv, 1, o), torch.gather(iv, 1, o)
        o32_v, o32_i = merge(out["d32"], out["i32"])
        o64_v, o64_i = merge(out["d64"], out["i64"])
        worst_kept = out["worst"].min(dim=0).values
    else:
        o32_i, o64_i = lists["i32"], lists["i64"]
    gd, gi = gpu_d[:n_par].cpu(), gpu_i[:n_par].cpu()
    margin_ok = bool((worst_kept > o64_v[:, -1] * (1 + 1e-3)).all())   # kept sets reach well beyond the oracle's k-th
    same32 = (gi == o32_i).all(dim=1).float().mean()
    same64 = (gi == o64_i).all(dim=1).float().mean()
    sets32 = torch.tensor([set(a.tolist()) == set(b.tolist()) for a, b in zip(gi, o32_i)]).float().mean()
    rel32 = ((gd - o32_v).abs() / o32_v).max()
    rel64 = ((gd.double() - o64_v).abs() / o64_v).max()
    return {"n_queries": n_par, "gallery_rows": "all (every shard)",
            "oracle": "CPU, per-query pmath.dist in the reference's fp32 geoopt form + top-k over the %d rows per query "
                      "and shard kept by a coarse fp32 prefilter (row-independent arithmetic: identical distances; "
                      "coverage asserted by kept_margin_ok); fp64 closed form beside it" % kk,
            "topk_lists_identical_frac_fp32_oracle": float(same32), "topk_sets_identical_frac_fp32_oracle": float(sets32),
            "topk_lists_identical_frac_fp64_oracle": float(same64),
            "max_rel_dist_diff_fp32_oracle": float(rel32), "max_rel_dist_diff_fp64_oracle": float(rel64),
            "kept_margin_ok": margin_ok, "seconds": round(time.perf_counter() - t0, 1),
            "threads_per_rank": torch.get_num_threads()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="c5: launch the step's kernels one by one (no CUDA graph)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else max(args.warmup, 0)
    if args.scaling is None:
        args.scaling = default_scaling(args.workload)

    Q, N, D, k, c, desc = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, Q, N, D, k, c, desc, world, rank, emit)
    if args.workload == "c3":
        return run_c3(args, Q, N, D, k, c, desc, world, rank, local_rank, emit)
    if args.workload == "c5":
        return run_c5(args, Q, D, c, desc, world, local_rank, emit)

    from patent_image_retrieval_b200 import SearchPipeline, StageEvents, ops, synth
    from patent_image_retrieval_b200.dist import ShardedGalleryIndex, shard_range
    from patent_image_retrieval_b200.retrieval import default_kbound, default_kprime
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- build the resident gallery shard (untimed) -------------------------------------------
    lo, hi = shard_range(N, rank, world)
    t_build = time.perf_counter()
    g_u = synth.gallery_rows(lo, hi, D, device=dev)
    weak = world > 1 and args.scaling == "weak"
    index = ShardedGalleryIndex(g_u, row_offset=lo, n_total=N, c=c, metric="hyperbolic", space="euclidean",
                                queries="sharded" if weak else "replicated",
                                exact=os.environ.get("HYPRET_BENCH_EXACT", "1") != "0")   # 0: diagnostics only
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    # gallery-side projection roofline: the same kernel call the index build made, timed alone
    n_proj = min(hi - lo, 2_000_000)
    ev_p = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(4):
        if it == 1:
            ev_p[0].record()
        ops.project_rows(g_u[:n_proj], c, mode="expmap0", side="gallery")
    ev_p[1].record()
    torch.cuda.synchronize()
    gallery_project_ms = ev_p[0].elapsed_time(ev_p[1]) / 3
    del g_u
    torch.cuda.empty_cache()
    q_dev = synth.gaussian_features(Q, D, seed=synth.SEED_QUERY + (100 * rank if weak else 0), device=dev)
    q_total = Q * world if weak else Q          # queries the whole job answers per step
    q_host = torch.empty(Q, D, dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    out_d_host = torch.empty(Q, k, dtype=torch.float32, pin_memory=True)
    out_i_host = torch.empty(Q, k, dtype=torch.int64, pin_memory=True)
    kprime = default_kprime(k)                  # what search() uses when the caller does not choose (16-slot lists at
                                                # k=10, sharing the bound of the 24th best score: default_kbound)
    plan = ops.score_plan(q_total if weak else Q, hi - lo, D, kprime)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(events=None):
        return index.search(q_dev, k=k, kprime=kprime, kernel_events=events)

    def step_e2e():
        qd = q_host.to(dev, non_blocking=True)
        dd, ii = index.search(qd, k=k, kprime=kprime)
        out_d_host.copy_(dd, non_blocking=True)
        out_i_host.copy_(ii, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller holds the result on the host

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- timed: device-resident ------------------------------------------------------------------
    kernel_events = StageEvents()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident(kernel_events)
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1) / args.steps
    score_ms, project_ms, rerank_ms = (kernel_events.ms(n_) for n_ in ("score", "project", "rerank"))
    certify_ms = kernel_events.ms("certify")                    # sharded serving: owner's certificate + flagged rescans

    # ---- timed: end to end with host buffers, no overlap ----------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e_serial = e0.elapsed_time(e1) / args.steps

    # ---- timed: end to end with host buffers, pipelined (the serving loop) ----------------------------
    # three buffer slots: the host consumes a step's result while the two following steps are already queued, so a
    # late wake-up of the host thread (8 ranks share the box's cores) does not leave the GPU idle
    pipe = SearchPipeline(index, Q, k=k, kprime=kprime, depth=3)
    q_hosts = [q_host, q_host.clone().pin_memory()]
    for s_ in range(4):
        pipe.submit(q_hosts[s_ % 2])
    pipe.drain()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(pipe.copy_in)
    pending = []
    for s_ in range(args.steps):
        pending.append(pipe.submit(q_hosts[s_ % 2]))
        if len(pending) > 2:
            pipe.result(pending.pop(0))           # the host consumes results two steps behind the submissions
    last = pending[-1]
    for slot in pending:
        pipe.result(slot)
    t_end.record(pipe.copy_out)
    barrier()
    ms_e2e = t_start.elapsed_time(t_end) / args.steps
    e2e_ok = bool(torch.equal(pipe.out_i[last], out_i_host))
    clocks = sampler.stop() if sampler is not None else None

    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e, score_ms, ms_e2e_serial, project_ms, rerank_ms, gallery_project_ms],
                         device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, score_ms, ms_e2e_serial, project_ms, rerank_ms, gallery_project_ms = (
            float(x) for x in t.tolist())

    # ---- size-independent result properties at full size ----------------------------------------------
    dd, ii = step_resident()
    torch.cuda.synchronize()
    props_ok = bool((dd[:, 1:] >= dd[:, :-1]).all()) and bool((ii >= 0).all()) and bool((ii < N).all())
    props_ok &= bool((ii.sort(dim=1).values[:, 1:] != ii.sort(dim=1).values[:, :-1]).all())   # no duplicates
    props_ok &= bool(torch.equal(out_i_host.to(dev), ii)) and e2e_ok                          # e2e == resident
    # exact-top-k guarantee: queries proven by the filter pass / recomputed by the full scan (this rank's shard)
    cert = index.local.certificate
    cert_stats = None
    if weak and getattr(index, "uncertified", None) is not None:
        # sharded serving: the OWNER certifies its merged lists; flagged queries are rescanned on every shard
        n_unc = index.uncertified.float().sum()
        t = torch.stack([Q - n_unc, n_unc, torch.tensor(float(Q), device=dev)]).double()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        cert_stats = {"certified_frac": float(t[0] / t[2]), "fallback_queries": int(t[1]),
                      "note": "per query, at its owner: merged list proven exact by hypret_cert_merged (global k'-th best "
                              "filter score - exact surrogate of the k-th result > rounding bound), else rescanned "
                              "exactly on every shard; summed over ranks"}
    elif cert is not None and not weak:
        t = torch.stack([cert.certified[:Q].float().sum(), cert.count[0].float(),
                         torch.tensor(float(Q), device=dev)]).double()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        cert_stats = {"certified_frac": float(t[0] / t[2]), "fallback_queries": int(t[1]),
                      "note": "per (query, gallery shard): proven exact by the margin > rounding-bound test of the rerank "
                              "kernel, else recomputed by hypret_exact_topk; summed over ranks"}
    q0 = q_dev
    if weak:
        # the two exchange patterns must agree: rank 0's batch through the replicated path
        # (shard-local search -> all_gather -> merge on every rank) == its result through the sharded path
        q0 = q_dev.clone()
        dist.broadcast(q0, src=0)
        dd2, ii2 = index.search_replicated(q0, k=k, kprime=kprime)
        torch.cuda.synchronize()
        if rank == 0:
            props_ok &= bool(torch.equal(ii2, ii)) and bool(torch.equal(dd2, dd))
    parity = None
    if not args.no_parity:
        gd, gi = dd, ii
        if world > 1:                                   # rank 0's result, visible to the merge on rank 0 only
            gd, gi = dd.clone(), ii.clone()
            dist.broadcast(gd, src=0)
            dist.broadcast(gi, src=0)
        parity = oracle_parity(index.local.rows32, lo, q0[:32], c, k, gd, gi, world, rank, dist)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    n_local = hi - lo
    q_scored = q_total if weak else Q           # query rows one rank's scoring kernel sees per step
    flops = 2.0 * q_scored * n_local * D
    kpad = ops.operand_kpad(D)
    project_bytes = q_scored * (4 * D + 4 * D + 2 * kpad)          # read f32 row, write f32 point + fp16 operand row
    if world > 1 and weak and getattr(index, "_exchange", None) is not None:
        project_bytes = Q * (4 * D + 4 * D + world * 2 * kpad)     # own rows only; operand row stored to every rank
    # exact rescoring: k' gathered fp32 rows per query; with the cross-shard surrogate threshold (weak mode) the
    # W shards share one query's k' rows between them
    rerank_bytes = q_scored * kprime * D * 4 // (world if weak else 1)
    peer_x = weak and getattr(index, "_exchange", None) is not None    # query exchange through peer memory
    peer_r = peer_x and os.environ.get("HYPRET_PEER_ROUTE", "1") != "0"  # every exchange fused into its producer
    # own kernels per step: project_rows, score_topk, rerank(+certificate), exact_topk (empty list: exits at once)
    # weak: + cert_merged, flag_compact, exact_topk, merge_topk and 3 fixed-size NCCL calls (flags, the rescans' lists)
    n_own = 4 if world == 1 else ((20 if peer_r else 14 if peer_x else 10) if weak else 5)
    n_nccl = 0 if world == 1 else ((3 if peer_r else 7 if peer_x else 8) if weak else 2)
    achieved = flops / (score_ms * 1e-3) / 1e12
    timed_region_s = args.steps * ms_resident * 1e-3
    # a region well under a second from an idle board runs at burst clocks; seconds of dense MMA settle at the power cap
    regime = "burst" if timed_region_s < 1.0 else "sustained"
    peak = peaks["tflops_burst"] if regime == "burst" else peaks["tflops_sustained"]
    traffic = None
    tp = ROOT / "profiles" / "score_topk_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload, {}).get(str(world))
        except Exception:
            traffic = None
    gproj_bytes = n_proj * (4 * D + 4 * D + 2 * kpad)
    line = {
        "metric": METRIC, "value": q_total / (ms_resident * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None, "dtype": "fp16 tensor-core filter (fp32 accumulate) + fp32/fp64 exact rerank, certified",
        "data": "synthetic",
        "config": search_config(args.workload, world, args.scaling),
        "run": {"kprime": kprime, "kbound": default_kbound(k, kprime), "gallery_rows_per_gpu": n_local,
                "stage_ms": {"project": project_ms, "score": score_ms, "rerank": rerank_ms,
                             "certify": None if certify_ms != certify_ms else certify_ms},
                "parallelism": (f"gallery row-shard x{world}; " +
                                ("each rank fed its own Q-query batch per step: " +
                                 ("projection kernel stores the operand rows into every rank's buffer over NVLink "
                                  "(peer memory), fp32 rows follow by copy engine under the scoring kernel"
                                  if peer_x else "all_gather(queries)") + " -> shard-local "
                                 "search of all W*Q -> " +
                                 ("cand_select / kth_smallest / pruned rerank store their outputs into the query "
                                  "owners' buffers (NVLink), counters instead of collectives"
                                  if peer_r else "all_to_all([Q,k] lists)") + " -> merge at the owner" if weak else
                                 "queries replicated: shard-local certified search -> all_gather([Q,k] lists) -> merge"))
                if world > 1 else "single GPU",
                "plan": {kk: plan[kk] for kk in ("grid", "n_lists", "stages", "resident", "l1", "l2")},
                "index_build_s": round(build_s, 3), "timed_region_s": round(timed_region_s, 3)},
        "e2e": {"value": q_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * k * 12, "bytes_are": "per rank",
                "mode": "SearchPipeline: per-step H2D + search + D2H, copies of neighbouring steps overlapped",
                "serial": {"value": q_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial}},
        "gpu_launches": args.steps * n_own,
        "gpu_launches_note": "own kernels per step per rank: project_rows, score_topk, " +
                             (("peer_signal x5, peer_wait x5, " if peer_r else "peer_signal x2, peer_wait x2, "
                               if peer_x else "") +
                              "cand_select, kth_smallest, rerank (pruned), merge_topk" if weak else
                              "rerank (certified), exact_topk, merge_topk" if world > 1 else
                              "rerank (certified), exact_topk") +
                             (" (+ %d NCCL collectives)" % n_nccl if world > 1 else ""),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "score_topk_kernel", "kernel_ms": score_ms, "algorithmic_flops": flops,
                     "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                     "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                     "peak_source": peaks["source"] + f", {regime} figure: the timed region is {timed_region_s:.2f} s "
                                    "(< 1 s from an idle board = burst clocks; longer = power-capped)"},
        "roofline_hbm": [
            {"kernel": "project_rows_kernel (gallery side, index build)", "bound": "hbm", "kernel_ms": gallery_project_ms,
             "algorithmic_bytes": gproj_bytes, "achieved": gproj_bytes / (gallery_project_ms * 1e-3) / 1e9,
             "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": gproj_bytes / (gallery_project_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
             "note": "%d gallery rows: read f32 row, write f32 point + fp16 operand row (+ certificate maxima)" % n_proj},
            {"kernel": "project_rows_kernel (query side, per step)", "bound": "hbm", "kernel_ms": project_ms,
             "algorithmic_bytes": project_bytes,
             "achieved": project_bytes / (project_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": project_bytes / (project_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
             "note": ("own %d rows, operand stored into all %d ranks' buffers over NVLink + arrival wait" % (Q, world))
                     if peer_x else "%d rows: launch-latency sized at this Q" % q_scored},
            {"kernel": "rerank_kernel (+ certificate) and exact_topk", "bound": "hbm", "kernel_ms": rerank_ms,
             "algorithmic_bytes": rerank_bytes,
             "achieved": rerank_bytes / (rerank_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": rerank_bytes / (rerank_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
             "note": "gather of k' random fp32 gallery rows per query"}],
        "clocks": clocks,
        "result_properties_ok": props_ok,
        "exactness": cert_stats,
        "parity": parity,
    }

    if world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path on a bounded sample: all rows of the gallery for small N, else a contiguous slice
        # (linear scan per query: queries/s scale with the fraction of rows)
        n_slice = min(N, 1_250_000)
        g_pts_cpu = index.local.rows32[:n_slice].cpu()
        frac = n_slice / N
        qps, n_q, secs, (d_cpu, i_cpu) = time_cpu_baseline(q_host, g_pts_cpu, c, k, fraction=frac)
        line["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": (f"first {n_q} queries of the same workload against " +
                                           (f"the full {N}-row gallery" if frac == 1.0 else
                                            f"rows [0, {n_slice}) of the {N}-row gallery (1/{N // n_slice} of the linear "
                                            f"scan; value = queries/s over the slice x {frac:.4f})") +
                                           f" (points pre-embedded), {secs:.1f} s"),
                                **cpu_info()}
        if frac == 1.0:
            line["cpu_baseline"]["topk_lists_identical_frac"] = float((i_cpu == ii[:n_q].cpu()).all(dim=1).float().mean())
            line["cpu_baseline"]["max_rel_dist_diff"] = float(((d_cpu - dd[:n_q].cpu()).abs() / d_cpu).max())
        # the two other CPU paths beside the faithful per-query loop -- (ii) a best-effort blocked
        # closed-form arccosh + top-k, (iii) the notebook's cosine_similarity + argsort -- on 64 queries each
        from oracle import head, retrieval
        n_b = min(64, Q)
        qb = head.embed_rows(q_host[:n_b], c)
        t0 = time.perf_counter()
        retrieval.hyperbolic_topk(qb, g_pts_cpu, c, k, form="arcosh")
        t_b = time.perf_counter() - t0
        t0 = time.perf_counter()
        retrieval.cosine_topk(q_host[:n_b].numpy(), g_pts_cpu.numpy(), k)
        t_c = time.perf_counter() - t0
        line["cpu_baseline"]["other_paths"] = {
            "best_effort_blocked_arccosh": {"value": frac * n_b / t_b, "unit": UNIT, "queries": n_b},
            "cosine_similarity_argsort": {"value": frac * n_b / t_c, "unit": UNIT, "queries": n_b}}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_c5(args, n, D, c, desc, world, local_rank, emit):
    """BASELINE config 5 on ONE GPU (replicas only beyond that): a step = forward + backward of the in-batch
    InfoNCE over n anchors x n positives through ``train.in_batch_contrastive_loss`` (gradients for both inputs).
    Data: anchors / positives = expmap0(mu_i + 0.5 eps) -- close pairs on the diagonal, an O(1) loss (the softmax is
    not saturated: every gradient is signal)."""
    if world != 1:
        raise SystemExit("--workload c5 is a single-GPU line (train_hyp scales as replicas)")
    from patent_image_retrieval_b200 import synth, train
    from patent_image_retrieval_b200.geoopt_shim import pmath
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    tau, kk, noise = 0.07, torch.tensor([-c]), 1.0
    mu = synth.gaussian_features(n, D, seed=2, scale=1.0, device=dev)
    mk = lambda seed: pmath.project(pmath.expmap0(mu + noise * synth.gaussian_features(n, D, seed=seed, scale=1.0,
                                                                                       device=dev), k=kk), k=kk)
    a, p = mk(3).requires_grad_(True), mk(4).requires_grad_(True)
    a_host, p_host = a.detach().cpu().pin_memory(), p.detach().cpu().pin_memory()
    loss_host = torch.empty(1, pin_memory=True)

    def step(host=False):
        if host:
            a.data.copy_(a_host, non_blocking=True)
            p.data.copy_(p_host, non_blocking=True)
        a.grad = p.grad = None
        loss = train.in_batch_contrastive_loss(a, p, kk, tau)
        loss.backward()
        if host:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / args.steps
    # The step is ~10 kernels of 30-160 us: launched one by one from Python it is host-bound on a slow host core, so
    # the timed step replays ONE CUDA graph of the same forward + backward (same kernels, same stream order; every
    # operator of the step is capture-safe: no host synchronisation, lengths and flags stay on the device).
    graph, g_loss, graph_note = None, None, "eager launches"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            a.grad = p.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_loss = train.in_batch_contrastive_loss(a, p, kk, tau)
                g_loss.backward()
            graph_note = "one CUDA graph of forward + backward per step"
        except Exception as exc:                                  # report, and time the eager step
            graph, graph_note = None, f"eager launches (graph capture failed: {type(exc).__name__}: {exc})"
            torch.cuda.synchronize()

    def step_timed(host=False):
        if graph is None:
            return step(host)
        if host:
            a.data.copy_(a_host, non_blocking=True)
            p.data.copy_(p_host, non_blocking=True)
        graph.replay()
        if host:
            loss_host.copy_(g_loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return g_loss

    for _ in range(3):
        step_timed()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    e0.record()
    for _ in range(args.steps):
        loss = step_timed()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    step_timed(host=True)
    e0.record()
    for _ in range(args.steps):
        step_timed(host=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    peaks = load_peaks()
    flops = 6.0 * n * n * D                                     # A P^T, (G o W) P, (G o W)^T A
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks["tflops_burst"] if args.steps * ms < 1000 else peaks["tflops_sustained"]
    flash = D % 16 == 0 and 16 <= D <= 128
    # the epilogues are transcendental-bound: rsqrt + lg2 + ex2 per pair and pass (forward, dA pass, dP pass)
    mufu_ops = 9.0 * n * n
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    mufu_peak = 16.0 * 148 * sm_mhz * 1e6                        # 16 MUFU results / clk / SM
    line = {
        "metric": "in-batch pairs/sec, train_hyp distance matrix forward+backward", "value": n * n / (ms * 1e-3),
        "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 2-way split Gram + bf16 2-plane gradient products on tcgen05 (fp32 accumulate), fp32 epilogues",
        "data": "synthetic",
        "config": {"workload": desc, "n": n, "D": D, "c": c, "tau": tau, "noise": noise, "parallelism": "single GPU",
                   "cache": "no [n,n] array exists; operands (~20 MB) live in L2", "launch": graph_note,
                   "eager_ms_per_step": ms_eager},
        "e2e": {"value": n * n / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n * D * 4, "d2h_bytes_per_step": 4,
                "mode": "per step: H2D of anchors and positives, forward + backward, D2H of the loss; no overlap"},
        "gpu_launches": args.steps * (10 if flash else 8),
        "gpu_launches_note": ("own kernels per step: flash_prep x2 (+ transposes x2), flash_lse + finish, rowpair_dist, "
                              "flash_grad + finish x2; no library GEMM, no [n,n] array" if flash else
                              "gram_split x2, gram_dist, lse_rows, pairdist_bwd_fused (+ library GEMMs of the split products)"),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": None, "kernel": "whole step",
                     "kernel_ms": ms, "algorithmic_flops": flops,
                     "note": "algorithmic 6 n^2 D; the kernels issue 3x that (2-way split operands: 3 products each)",
                     "peak_source": peaks["source"]},
        "roofline_mufu": {"bound": "mufu", "ops": mufu_ops, "peak_ops_per_s": mufu_peak, "floor_ms": mufu_ops / mufu_peak * 1e3,
                          "frac": mufu_ops / mufu_peak * 1e3 / ms,
                          "note": "9 transcendentals per pair (rsqrt, lg2, ex2 in each of the three passes) at 16 / clk / SM"},
        "clocks": clocks, "loss": float(loss.detach()),
        "result_properties_ok": bool(torch.isfinite(a.grad).all() and torch.isfinite(p.grad).all()),
    }
    if not args.no_parity:
        # parity at FULL size: the loss over all n rows and the gradient of a sample of anchor rows against fp64 on the
        # CPU (closed-form distances in blocks, softmax in fp64, autograd for the sampled rows)
        from oracle import retrieval
        t0 = time.perf_counter()
        a64, p64 = a.detach().cpu().double(), p.detach().cpu().double()
        d64 = retrieval.hyperbolic_dist_rows(a64, p64, c, form="arcosh", block=512)
        logits = -d64 / tau
        loss64 = float((torch.logsumexp(logits, dim=1) - logits.diagonal()).mean())
        rows = torch.arange(0, n, max(1, n // 64))[:64]
        ar = a64[rows].clone().requires_grad_(True)
        from oracle import pmath as opm
        dr = opm.dist(ar[:, None, :], p64[None, :, :], k=torch.tensor(-float(c), dtype=torch.float64))
        lr = -dr / tau
        ((torch.logsumexp(lr, dim=1) - lr[torch.arange(len(rows)), rows]).sum() / n).backward()
        ga = a.grad.detach().cpu().double()[rows]
        line["parity"] = {"loss_fp64": loss64, "loss_rel_diff": abs(line["loss"] - loss64) / abs(loss64),
                          "grad_rows_checked": int(len(rows)),
                          "grad_max_abs_diff_over_max_abs": float((ga - ar.grad).abs().max() / ar.grad.abs().max()),
                          "oracle": "fp64 closed-form distances + logsumexp on the CPU over all n x n pairs (loss); fp64 "
                                    "autograd of the geoopt form for the sampled anchor rows (gradient)",
                          "seconds": round(time.perf_counter() - t0, 1)}
    if not args.no_cpu_baseline:
        # the reference's literal double loop of 1x1 pmath.dist + autograd (src/train.py:1832-1846) at its own
        # batch size; the loop is O(n^2), so pairs/s is the size-independent figure
        from oracle import contrastive, head
        torch.set_num_threads(os.cpu_count() or 1)
        ns = 96
        ac = head.embed_rows(synth.gaussian_features(ns, D, seed=3, scale=1.0), c).requires_grad_(True)
        pc = head.embed_rows(synth.gaussian_features(ns, D, seed=4, scale=1.0), c).requires_grad_(True)
        t0 = time.perf_counter()
        contrastive.contrastive_loss(ac, pc, kk, temperature=tau, loop=True).backward()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns * ns / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"double loop + autograd at n={ns} ({dt:.1f} s)", **cpu_info()}
    emit(line)
    return 0


def run_c3(args, Q, N, D, k, c, desc, world, rank, local_rank, emit):
    """BASELINE config 3 on ONE GPU: a step answers all Q queries under both metrics (fused projection, tcgen05
    scoring with 64-slot lists, wide exact rerank with the per-query certificate), in chunks of 20k queries."""
    if world != 1:
        raise SystemExit("--workload c3 is a single-GPU line (replicas only beyond that)")
    from patent_image_retrieval_b200 import GalleryIndex, StageEvents, synth
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    chunk = 20_000
    t_build = time.perf_counter()
    g_u = synth.gaussian_features(N, D, seed=synth.SEED_GALLERY, device=dev)
    indexes = {m: GalleryIndex(g_u, c=c, metric=m) for m in ("hyperbolic", "cosine")}
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    q_dev = synth.gaussian_features(Q, D, seed=synth.SEED_QUERY, device=dev)
    q_host = torch.empty(Q, D, dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    out_s = {m: torch.empty(Q, k, dtype=torch.float32, pin_memory=True) for m in indexes}
    out_i = {m: torch.empty(Q, k, dtype=torch.int64, pin_memory=True) for m in indexes}

    def step(events=None, host=False, margins=None):
        for q0 in range(0, Q, chunk):
            qc = q_host[q0:q0 + chunk].to(dev, non_blocking=True) if host else q_dev[q0:q0 + chunk]
            for m, index in indexes.items():
                res = index.search(qc, k=k, kernel_events=events, return_margin=margins is not None)
                if margins is not None:
                    margins.append(index.uncertified_wide.clone())
                if host:
                    out_s[m][q0:q0 + chunk].copy_(res[0], non_blocking=True)
                    out_i[m][q0:q0 + chunk].copy_(res[1], non_blocking=True)
        if host:
            torch.cuda.current_stream().synchronize()
        return res

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    ev = StageEvents()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    step(host=True)
    e0.record()
    for _ in range(args.steps):
        step(host=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    margins = []
    step(margins=margins)
    torch.cuda.synchronize()
    unc = torch.cat(margins)                                   # 1 = margin <= rounding bound: paged through the exact scan
    certified, n_fallback = float(1.0 - unc.float().mean()), int(unc.sum())
    ok = all(bool((out_s["hyperbolic"][:, 1:] >= out_s["hyperbolic"][:, :-1]).all()) for _ in (0,))
    ok &= bool((out_s["cosine"][:, 1:] <= out_s["cosine"][:, :-1]).all())
    ok &= all(bool((out_i[m] >= 0).all()) and bool((out_i[m] < N).all()) for m in indexes)
    peaks = load_peaks()
    calls = (Q // chunk) * 2                                   # scoring kernels per step
    score_ms = ev.ms("score") * calls                          # per step
    flops = 2.0 * 2.0 * Q * N * D                              # two GEMM passes (one per metric)
    achieved = flops / (score_ms * 1e-3) / 1e12
    line = {
        "metric": "queries/sec at top-100 over N-gallery, both metrics", "value": Q / (ms * 1e-3), "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16 tensor-core filter (fp32 accumulate) + fp32/fp64 exact rerank",
        "data": "synthetic",
        "config": {"workload": desc, "Q": Q, "N": N, "D": D, "k": k, "c": c, "kprime": 64, "query_chunk": chunk,
                   "parallelism": "single GPU", "index_build_s": round(build_s, 3),
                   "cache": "inputs larger than L2 (fp16 gallery operand %.0f MB vs 126 MB L2); no flush" %
                            (N * (D + 16) * 2 / 1e6)},
        "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": Q * D * 4,
                "d2h_bytes_per_step": 2 * Q * k * 12, "mode": "per chunk: H2D, both searches, D2H; no overlap"},
        "gpu_launches": args.steps * calls * 3,
        "gpu_launches_note": "per chunk and metric: project_rows, score_topk, rerank_wide",
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": None, "kernel": "score_topk_kernel",
                     "kernel_ms": score_ms, "algorithmic_flops": flops,
                     "peak_source": peaks["source"] + ", sustained figure"},
        "stage_ms_per_step": {n_: ev.ms(n_) * calls for n_ in ev.STAGES},
        "clocks": clocks, "certified_frac": certified, "result_properties_ok": bool(ok),
        "exactness": {"certified_frac": certified, "fallback_queries": n_fallback,
                      "note": "per (query, metric): wide-rerank margin > rounding bound E, else paged through the exact "
                              "ranking by full scans (hypret_exact_topk / _after) on the same stream"},
    }
    if not args.no_cpu_baseline:
        # the reference's two CPU paths on a bounded sample: per-query pmath.dist + topk (src/train.py:3259) and
        # sklearn-style cosine + argsort (notebooks/retrieval.ipynb:368,383), restated in oracle/
        from oracle import head, retrieval
        torch.set_num_threads(os.cpu_count() or 1)
        n_s = 8
        g_cpu = g_u.cpu()
        g_pts = indexes["hyperbolic"].rows32.cpu()
        t0 = time.perf_counter()
        d_h, i_h = retrieval.hyperbolic_topk(head.embed_rows(q_host[:n_s], c), g_pts, c, k, form="geoopt")
        _, i_c = retrieval.cosine_topk(q_host[:n_s].numpy(), g_cpu.numpy(), k)
        i_c = torch.from_numpy(i_c)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"first {n_s} queries under both metrics against the full gallery, {dt:.1f} s",
                                "hyperbolic_lists_identical_frac": float((i_h == out_i["hyperbolic"][:n_s]).all(1).float().mean()),
                                "cosine_sets_identical_frac": float(torch.tensor(
                                    [set(i_c[r].tolist()) == set(out_i["cosine"][r].tolist()) for r in range(n_s)]).float().mean())}
    emit(line)
    return 0


def run_reference(args, Q, N, D, k, c, desc, world, rank, emit):
    """CPU arm: the oracle port of the reference path on this box's host cores (geoopt is not
    installable, so the unmodified reference cannot run)."""
    if rank != 0:
        return 0
    from oracle import head
    torch.set_num_threads(os.cpu_count() or 1)
    if args.workload == "c5":
        # the reference's double loop of 1x1 pmath.dist + autograd (src/train.py:1832-1846); O(n^2): pairs/s
        from oracle import contrastive
        from patent_image_retrieval_b200 import synth
        ns, tau, kk = 64, 0.07, torch.tensor([-c])
        times = []
        for s_ in range(args.warmup + args.steps):
            ac = head.embed_rows(synth.gaussian_features(ns, D, seed=3 + 2 * s_, scale=1.0), c).requires_grad_(True)
            pc = head.embed_rows(synth.gaussian_features(ns, D, seed=4 + 2 * s_, scale=1.0), c).requires_grad_(True)
            t0 = time.perf_counter()
            contrastive.contrastive_loss(ac, pc, kk, temperature=tau, loop=True).backward()
            if s_ >= args.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        val = ns * ns / (ms * 1e-3)
        sample = f"double loop + autograd over a {ns} x {ns} batch per step"
        emit({"impl": "reference", "metric": "in-batch pairs/sec, train_hyp distance matrix forward+backward",
              "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic", "config": {"workload": desc, "n": Q, "D": D, "c": c, "tau": tau, "sample": sample},
              "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": sample},
              "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return 0
    k = max(k, 1)
    from patent_image_retrieval_b200 import synth       # the data generator only: no kernel, no engine on this path
    scaling = args.scaling or default_scaling(args.workload)
    # bounded sample: a contiguous slice of the SAME gallery (same block seeds as the native arm; CPU generator) and
    # the first queries of the SAME batch; the path is a linear scan per query, so queries/s over the whole gallery
    # = (rows of the slice / N) * queries/s over the slice
    n_slice = min(N, 1_250_000)
    frac = n_slice / N
    g_pts = torch.empty(n_slice, D)
    for r0 in range(0, n_slice, 1 << 18):          # chunked: bounded temporaries
        r1 = min(n_slice, r0 + (1 << 18))
        g_pts[r0:r1] = head.embed_rows(synth.gallery_rows(r0, r1, D), c)
    q_u = synth.gaussian_features(64, D, seed=synth.SEED_QUERY)
    per_step = 2 if n_slice < 100_000 else 1        # ~0.1-2 s of CPU work per step
    times = []
    for s in range(args.warmup + args.steps):
        qs = q_u[(s * per_step) % 64:(s * per_step) % 64 + per_step]
        t0 = time.perf_counter()
        cpu_oracle_topk(qs, g_pts, c, k)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = frac * per_step / (ms * 1e-3)
    sample = (f"{per_step} queries per step against " +
              (f"the full {N}-row gallery" if frac == 1.0 else
               f"rows [0, {n_slice}) of the {N}-row gallery (1/{N // n_slice} of the linear scan; value = queries/s over "
               f"the slice x {frac:.4f})") + " (points pre-embedded); oracle port of the reference's per-query "
              "pmath.dist + top-k (geoopt is not installable: the unmodified reference cannot run)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": search_config(args.workload, world, scaling)["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": search_config(args.workload, world, scaling),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample, **cpu_info()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
