"""Scoring kernel alone (hypret_score_topk_bound, shared thresholds, k' = 16, kbound = 24) on the shapes where the
cold phase of a strip start matters: C1 (4-tile strips), the 37.5k-row shards of 8-GPU sharded serving, C2.
Checks the union of each query's lists against torch.topk of the kernel's own score matrix on a small case first.
Usage (GPU box): python tools/bench_score_shapes.py [--reps 20]"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from patent_image_retrieval_b200 import ops  # noqa: E402

SHAPES = [("C1", 1000, 10000, 2048), ("shard37k", 80000, 37500, 512), ("C2", 10000, 300000, 512),
          ("C4/4", 10000, 2500000, 512)]


def operands(Q, N, d, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    u = torch.randn(Q, d, device="cuda", generator=g) * (0.45 / d ** 0.5)
    v = torch.randn(N, d, device="cuda", generator=g) * (0.45 / d ** 0.5)
    _, q_op, _ = ops.project_rows(u, 1.0, "expmap0", "query")
    _, g_op, _ = ops.project_rows(v, 1.0, "expmap0", "gallery")
    return q_op, g_op


def union_check(Q=700, N=9000, d=512, kprime=16, kbound=24):
    """Every query's global top-kbound (by the kernel's own scores) must sit in the union of its lists unless one
    half-strip list overflowed (then at least its top-kprime must)."""
    q_op, g_op = operands(Q, N, d, seed=3)
    cs, ci, dbg = ops.score_topk(q_op, g_op, d, kprime, debug=True, kbound=kbound)
    torch.cuda.synchronize()
    want = torch.topk(dbg, kprime, dim=1, largest=False).indices
    have = ci.reshape(Q, -1)
    miss = 0
    for q in range(Q):
        s = set(have[q].tolist())
        miss += sum(1 for j in want[q].tolist() if j not in s)
    return miss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    print("union check (top-16 of 700 x 9000 inside the lists): missing =", union_check(), flush=True)
    for name, Q, N, d in SHAPES:
        q_op, g_op = operands(Q, N, d)
        plan = ops.score_plan(Q, N, d, 16)
        S = plan["n_lists"]
        cs = torch.empty(Q, S, 16, dtype=torch.float32, device="cuda")
        ci = torch.empty(Q, S, 16, dtype=torch.int32, device="cuda")
        ws = torch.empty(Q, dtype=torch.int32, device="cuda")
        lc = torch.empty(Q, dtype=torch.int32, device="cuda")
        for _ in range(5):
            ops.score_topk(q_op, g_op, d, 16, out=(cs, ci), thr_workspace=ws, list_count=lc, kbound=24)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        ev[0].record()
        for i in range(a.reps):
            ops.score_topk(q_op, g_op, d, 16, out=(cs, ci), thr_workspace=ws, list_count=lc, kbound=24)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
        med = ts[len(ts) // 2]
        tf = 2.0 * Q * N * d / (med * 1e-3) / 1e12
        print(f"{name:9s} Q={Q} N={N} D={d} grid={plan['grid']} lists={S}: median {med:.4f} ms (min {ts[0]:.4f}) "
              f"= {tf:.0f} TFLOP/s  [memsets of the launch included]", flush=True)
        del q_op, g_op, cs, ci


if __name__ == "__main__":
    main()
