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


def cpu_oracle_topk(q_u, g_u, c, k):
    """The reference's path restated (oracle/): embed, per-query pmath.dist over the gallery
    (src/train.py:3259), top-k (src/auxiliary.py:374)."""
    from oracle import head, retrieval
    q = head.embed_rows(q_u, c)
    g = head.embed_rows(g_u, c) if g_u.shape[1] == q_u.shape[1] else g_u
    return retrieval.hyperbolic_topk(q, g, c, k, form="geoopt")


def time_cpu_baseline(q_u_cpu, g_pts_cpu, c, k, budget_s=15.0, per_step=None):
    """Queries/s of the oracle on a bounded sample (gallery points pre-embedded: the index is
    resident for the CPU arm too).  Returns (qps, n_queries, seconds, (dist, idx))."""
    from oracle import head, retrieval
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    q = head.embed_rows(q_u_cpu[:1], c)
    retrieval.hyperbolic_topk(q, g_pts_cpu, c, k, form="geoopt")
    one = time.perf_counter() - t0
    n = per_step if per_step is not None else int(max(2, min(64, budget_s / max(one, 1e-3))))
    n = min(n, q_u_cpu.shape[0])
    t0 = time.perf_counter()
    q = head.embed_rows(q_u_cpu[:n], c)
    res = retrieval.hyperbolic_topk(q, g_pts_cpu, c, k, form="geoopt")
    dt = time.perf_counter() - t0
    return n / dt, n, dt, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else max(args.warmup, 0)

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
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- build the resident gallery shard (untimed) -------------------------------------------
    lo, hi = shard_range(N, rank, world)
    t_build = time.perf_counter()
    g_u = synth.gaussian_features(hi - lo, D, seed=synth.SEED_GALLERY + 1000 * rank, device=dev)
    weak = world > 1 and args.scaling == "weak"
    index = ShardedGalleryIndex(g_u, row_offset=lo, n_total=N, c=c, metric="hyperbolic", space="euclidean",
                                queries="sharded" if weak else "replicated")
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    q_dev = synth.gaussian_features(Q, D, seed=synth.SEED_QUERY + (100 * rank if weak else 0), device=dev)
    q_total = Q * world if weak else Q          # queries the whole job answers per step
    q_host = torch.empty(Q, D, dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    out_d_host = torch.empty(Q, k, dtype=torch.float32, pin_memory=True)
    out_i_host = torch.empty(Q, k, dtype=torch.int64, pin_memory=True)
    kprime = 16
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
        t = torch.tensor([ms_resident, ms_e2e, score_ms, ms_e2e_serial, project_ms, rerank_ms], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, score_ms, ms_e2e_serial, project_ms, rerank_ms = (float(x) for x in t.tolist())

    # ---- size-independent result properties at full size ----------------------------------------------
    dd, ii = step_resident()
    torch.cuda.synchronize()
    props_ok = bool((dd[:, 1:] >= dd[:, :-1]).all()) and bool((ii >= 0).all()) and bool((ii < N).all())
    props_ok &= bool((ii.sort(dim=1).values[:, 1:] != ii.sort(dim=1).values[:, :-1]).all())   # no duplicates
    props_ok &= bool(torch.equal(out_i_host.to(dev), ii)) and e2e_ok                          # e2e == resident
    if weak:
        # the two exchange patterns must agree: rank 0's batch through the replicated path
        # (shard-local search -> all_gather -> merge on every rank) == its result through the sharded path
        q0 = q_dev.clone()
        dist.broadcast(q0, src=0)
        dd2, ii2 = index.search_replicated(q0, k=k, kprime=kprime)
        torch.cuda.synchronize()
        if rank == 0:
            props_ok &= bool(torch.equal(ii2, ii)) and bool(torch.equal(dd2, dd))

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
    project_bytes = q_scored * (4 * D + 4 * D + 2 * kpad)          # read f32 row, write f32 point + bf16 operand row
    if world > 1 and weak and getattr(index, "_exchange", None) is not None:
        project_bytes = Q * (4 * D + 4 * D + world * 2 * kpad)     # own rows only; operand row stored to every rank
    # exact rescoring: k' gathered fp32 rows per query; with the cross-shard surrogate threshold (weak mode) the
    # W shards share one query's k' rows between them
    rerank_bytes = q_scored * kprime * D * 4 // (world if weak else 1)
    peer_x = weak and getattr(index, "_exchange", None) is not None    # query exchange through peer memory
    peer_r = peer_x and os.environ.get("HYPRET_PEER_ROUTE", "1") != "0"  # every exchange fused into its producer
    n_own = 3 if world == 1 else ((16 if peer_r else 10 if peer_x else 6) if weak else 4)   # own kernels per step
    n_nccl = 0 if world == 1 else ((0 if peer_r else 4 if peer_x else 5) if weak else 2)
    achieved = flops / (score_ms * 1e-3) / 1e12
    traffic = None
    tp = ROOT / "profiles" / "score_topk_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload, {}).get(str(world))
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": q_total / (ms_resident * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True,
        "scaling": "weak" if (weak or world == 1) else "strong",
        "vs_baseline": None, "dtype": "bf16 tensor-core filter + fp32/fp64 exact rerank", "data": "synthetic",
        "config": {"workload": desc, "Q": Q, "N": N, "D": D, "k": k, "c": c, "kprime": kprime,
                   "queries_per_step_total": q_total, "gallery_rows_per_gpu": n_local,
                   "parallelism": (f"gallery row-shard x{world}; " +
                                   ("each rank fed its own Q-query batch per step: " +
                                    ("projection kernel stores the operand rows into every rank's buffer over NVLink "
                                     "(peer memory), fp32 rows follow by copy engine under the scoring kernel"
                                     if peer_x else "all_gather(queries)") + " -> shard-local "
                                    "search of all W*Q -> " +
                                    ("cand_select / kth_smallest / pruned rerank store their outputs into the query "
                                     "owners' buffers (NVLink), counters instead of collectives"
                                     if peer_r else "all_to_all([Q,k] lists)") + " -> merge at the owner" if weak else
                                    "queries replicated: shard-local search -> all_gather([Q,k] lists) -> merge"))
                   if world > 1 else "single GPU",
                   "cache": "inputs larger than L2 (bf16 gallery operand %.0f MB vs 126 MB L2); no flush" %
                            (n_local * ops.operand_kpad(D) * 2 / 1e6),
                   "plan": {kk: plan[kk] for kk in ("grid", "n_lists", "stages", "resident", "l1", "l2")},
                   "index_build_s": round(build_s, 3)},
        "e2e": {"value": q_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * k * 12, "bytes_are": "per rank",
                "mode": "SearchPipeline: per-step H2D + search + D2H, copies of neighbouring steps overlapped",
                "serial": {"value": q_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial}},
        "gpu_launches": args.steps * n_own,
        "gpu_launches_note": "own kernels per step per rank: project_rows, score_topk, " +
                             (("peer_signal x5, peer_wait x5, " if peer_r else "peer_signal x2, peer_wait x2, "
                               if peer_x else "") +
                              "cand_select, kth_smallest, rerank (pruned), merge_topk" if weak else
                              "rerank, merge_topk" if world > 1 else "rerank") +
                             (" (+ %d NCCL collectives)" % n_nccl if world > 1 else ""),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                     "kernel": "score_topk_kernel", "kernel_ms": score_ms, "algorithmic_flops": flops,
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)"},
        "roofline_hbm": [
            {"kernel": "project_rows_kernel", "bound": "hbm", "kernel_ms": project_ms, "algorithmic_bytes": project_bytes,
             "achieved": project_bytes / (project_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": project_bytes / (project_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
             "note": ("own %d rows, operand stored into all %d ranks' buffers over NVLink + arrival wait" % (Q, world))
                     if peer_x else "query side only (%d rows): launch-latency sized at this Q" % q_scored},
            {"kernel": "rerank_kernel", "bound": "hbm", "kernel_ms": rerank_ms, "algorithmic_bytes": rerank_bytes,
             "achieved": rerank_bytes / (rerank_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": rerank_bytes / (rerank_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
             "note": "gather of k' random fp32 gallery rows per query"}],
        "clocks": clocks,
        "result_properties_ok": props_ok,
    }

    if world == 1 and not args.no_cpu_baseline:
        g_pts_cpu = index.local.rows32.cpu()
        qps, n_q, secs, (d_cpu, i_cpu) = time_cpu_baseline(q_host, g_pts_cpu, c, k)
        same = float((i_cpu == ii[:n_q].cpu()).all(dim=1).float().mean())
        rel = float(((d_cpu - dd[:n_q].cpu()).abs() / d_cpu).max())
        line["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"first {n_q} queries of the same workload against the full {N}-row "
                                          f"gallery (points pre-embedded), {secs:.1f} s",
                                "topk_lists_identical_frac": same, "max_rel_dist_diff": rel}
        # the two other CPU paths beside the faithful per-query loop -- (ii) a best-effort blocked
        # closed-form arccosh + top-k, (iii) the notebook's cosine_similarity + argsort -- on 64 queries each
        from oracle import head, retrieval
        n_b = min(64, Q)
        qb = head.embed_rows(q_host[:n_b], c)
        t0 = time.perf_counter()
        _, i_b = retrieval.hyperbolic_topk(qb, g_pts_cpu, c, k, form="arcosh")
        t_b = time.perf_counter() - t0
        g_raw = g_u.cpu().numpy()
        t0 = time.perf_counter()
        retrieval.cosine_topk(q_host[:n_b].numpy(), g_raw, k)
        t_c = time.perf_counter() - t0
        line["cpu_baseline"]["other_paths"] = {
            "best_effort_blocked_arccosh": {"value": n_b / t_b, "unit": UNIT, "queries": n_b,
                                            "topk_lists_identical_frac": float((i_b == ii[:n_b].cpu()).all(dim=1).float().mean())},
            "cosine_similarity_argsort": {"value": n_b / t_c, "unit": UNIT, "queries": n_b}}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_c5(args, n, D, c, desc, world, local_rank, emit):
    """BASELINE config 5 on ONE GPU (replicas only beyond that): a step = forward + backward of the in-batch
    InfoNCE over n anchors x n positives through ``train.in_batch_contrastive_loss`` (gradients for both inputs)."""
    if world != 1:
        raise SystemExit("--workload c5 is a single-GPU line (train_hyp scales as replicas)")
    from patent_image_retrieval_b200 import synth, train
    from patent_image_retrieval_b200.geoopt_shim import pmath
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    tau, kk = 0.07, torch.tensor([-c])
    mu = synth.gaussian_features(n, D, seed=2, scale=1.0, device=dev)
    mk = lambda seed: pmath.project(pmath.expmap0(mu + 0.1 * synth.gaussian_features(n, D, seed=seed, scale=1.0,
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
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
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
    peaks = load_peaks()
    flops = 6.0 * n * n * D                                     # A P^T, (G o W) P, (G o W)^T A
    achieved = flops / (ms * 1e-3) / 1e12
    line = {
        "metric": "in-batch pairs/sec, train_hyp distance matrix forward+backward", "value": n * n / (ms * 1e-3),
        "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "3 x bf16 split operands on tcgen05 (fp32 accumulate), fp32 epilogues", "data": "synthetic",
        "config": {"workload": desc, "n": n, "D": D, "c": c, "tau": tau, "parallelism": "single GPU",
                   "cache": "the [n,n] matrices (268 MB each) exceed the 126 MB L2; no flush"},
        "e2e": {"value": n * n / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * n * D * 4, "d2h_bytes_per_step": 4,
                "mode": "per step: H2D of anchors and positives, forward + backward, D2H of the loss; no overlap"},
        "gpu_launches": args.steps * 8,
        "gpu_launches_note": "own kernels per step: gram_split x2, gram_dist, lse_rows, pairdist_bwd_fused (+ torch "
                             "reductions, 6 library bf16 GEMMs of the split products)",
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": None, "kernel": "whole step",
                     "kernel_ms": ms, "algorithmic_flops": flops,
                     "note": "the step materialises the [n,n] distance and weight matrices (DESIGN 4.6): it is bound "
                             "by their HBM passes and the fp32 epilogues, not by the tensor pipe",
                     "peak_source": peaks["source"] + ", sustained figure"},
        "clocks": clocks, "loss": float(loss.detach()),
        "result_properties_ok": bool(torch.isfinite(a.grad).all() and torch.isfinite(p.grad).all()),
    }
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
                                "sample": f"double loop + autograd at n={ns} ({dt:.1f} s)"}
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
                    margins.append(res[2])
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
    certified = float(torch.cat(margins).gt(0).float().mean())
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
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 tensor-core filter + fp32/fp64 exact rerank",
        "data": "synthetic",
        "config": {"workload": desc, "Q": Q, "N": N, "D": D, "k": k, "c": c, "kprime": 64, "query_chunk": chunk,
                   "parallelism": "single GPU", "index_build_s": round(build_s, 3),
                   "cache": "inputs larger than L2 (bf16 gallery operand %.0f MB vs 126 MB L2); no flush" %
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
    gen = torch.Generator().manual_seed(0)
    sigma = 0.45 / D ** 0.5
    g_pts = torch.empty(N, D)
    for r0 in range(0, N, 1 << 18):          # chunked: bounded temporaries
        r1 = min(N, r0 + (1 << 18))
        g_pts[r0:r1] = head.embed_rows(torch.randn(r1 - r0, D, generator=gen) * sigma, c)
    q_u = torch.randn(64, D, generator=torch.Generator().manual_seed(1)) * sigma
    per_step = 2 if N < 100_000 else 1        # bounded sample: ~1 s of CPU work per step at C2
    times = []
    for s in range(args.warmup + args.steps):
        qs = q_u[(s * per_step) % 64:(s * per_step) % 64 + per_step]
        t0 = time.perf_counter()
        cpu_oracle_topk(qs, g_pts, c, k)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = per_step / (ms * 1e-3)
    sample = f"{per_step} queries per step against the full {N}-row gallery (points pre-embedded)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "Q": Q, "N": N, "D": D, "k": k, "c": c, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
