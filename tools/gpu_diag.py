"""First-contact diagnostics on a B200: localises descriptor / swizzle / pipeline bugs of the
tcgen05 scoring kernel in ONE run (block-by-block operand masking), then times it.
Usage: python tools/gpu_diag.py [--quick]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from patent_image_retrieval_b200 import ops  # noqa: E402


def ref_scores(q_op, g_op):
    return q_op.float() @ g_op.float().t()


def check(Q, N, d, kprime=16, hint=0, label=""):
    torch.manual_seed(0)
    dev = "cuda"
    u = torch.randn(Q, d, device=dev) * (0.45 / d ** 0.5)
    v = torch.randn(N, d, device=dev) * (0.45 / d ** 0.5)
    _, q_op, _ = ops.project_rows(u, 1.0, "expmap0", "query")
    _, g_op, _ = ops.project_rows(v, 1.0, "expmap0", "gallery")
    plan = ops.score_plan(Q, N, d, kprime, hint)
    print(f"[{label}] Q={Q} N={N} d={d} plan={plan}", flush=True)
    kpad = q_op.shape[1]
    dpad = kpad - 16
    ok_all = True
    blocks = [("all", 0, kpad)] + [(f"kb{b}", b * 64, b * 64 + 64) for b in range(dpad // 64)] + [("ext", dpad, kpad)]
    for name, c0, c1 in blocks:
        qm = torch.zeros_like(q_op)
        gm = torch.zeros_like(g_op)
        qm[:, c0:c1] = q_op[:, c0:c1]
        gm[:, c0:c1] = g_op[:, c0:c1]
        cs, ci, dbg = ops.score_topk(qm, gm, d, kprime, hint, debug=True, share_thresholds=False)
        torch.cuda.synchronize()
        ref = ref_scores(qm, gm)
        err = (dbg - ref).abs()
        mx = float(err.max())
        scale = float(ref.abs().max()) + 1e-30
        bad = int((err > 1e-4 * scale + 1e-6).sum())
        ok = bad == 0
        ok_all &= ok
        print(f"   block {name:>4}: max|err|={mx:.3e} ref_max={scale:.3e} bad={bad}/{err.numel()} {'OK' if ok else 'FAIL'}",
              flush=True)
        if not ok and name != "all":
            ij = torch.nonzero(err > 1e-4 * scale + 1e-6)[:8]
            for i, j in ij.tolist():
                print(f"      ({i},{j}) got {float(dbg[i, j]):.6e} want {float(ref[i, j]):.6e}")
        if name == "all":
            # candidate lists vs torch.topk on the kernel's own score matrix
            miss = 0
            for cta, step, qt, g0, g1, slot in ops.score_strips(Q, N, d, kprime, hint):
                r0, r1 = qt * 128, min(Q, (qt + plan["pair"]) * 128)
                lo, hi = g0 * 256, min(N, g1 * 256)
                kk = min(kprime, hi - lo)
                want = torch.topk(dbg[r0:r1, lo:hi], kk, dim=1, largest=False).values.sort(dim=1).values
                eg = plan.get("epi_groups", 1)      # a strip owns eg consecutive list slots (one per epilogue warpgroup)
                got = cs[r0:r1, slot:slot + eg, :].reshape(r1 - r0, -1).sort(dim=1).values[:, :kk]
                miss += int((want != got).sum())
            print(f"   top-k' lists vs torch.topk(debug matrix): mismatching entries = {miss}", flush=True)
            ok_all &= miss == 0
    return ok_all


def bench(Q, N, d, kprime=16, iters=5):
    dev = "cuda"
    u = torch.randn(Q, d, device=dev) * (0.45 / d ** 0.5)
    v = torch.randn(N, d, device=dev) * (0.45 / d ** 0.5)
    _, q_op, _ = ops.project_rows(u, 1.0, "expmap0", "query")
    _, g_op, _ = ops.project_rows(v, 1.0, "expmap0", "gallery")
    plan = ops.score_plan(Q, N, d, kprime)
    cs = torch.empty(Q, plan["n_lists"], kprime, device=dev)
    ci = torch.empty(Q, plan["n_lists"], kprime, device=dev, dtype=torch.int32)
    ws = torch.empty(Q, device=dev, dtype=torch.int32)
    for share in (False, True):
        for _ in range(2):
            ops.score_topk(q_op, g_op, d, kprime, out=(cs, ci), share_thresholds=share, thr_workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            ops.score_topk(q_op, g_op, d, kprime, out=(cs, ci), share_thresholds=share, thr_workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * Q * N * d / ms / 1e9
        print(f"[bench share={int(share)}] Q={Q} N={N} d={d} grid={plan['grid']} lists={plan['n_lists']} "
              f"stages={plan['stages']} resident={plan['resident']} pair={plan['pair']}: {ms:.3f} ms  {tf:.1f} TFLOP/s (algorithmic)  "
              f"{Q / ms * 1e3:.0f} q/s", flush=True)


if __name__ == "__main__":
    if "--bench" in sys.argv:
        i = sys.argv.index("--bench")
        Q, N, d = (int(x) for x in sys.argv[i + 1:i + 4])
        bench(Q, N, d, iters=int(sys.argv[i + 4]) if len(sys.argv) > i + 4 else 5)
        sys.exit(0)
    t0 = time.time()
    print(torch.cuda.get_device_name(0), flush=True)
    ok = check(200, 1000, 512, label="resident-small")
    ok &= check(900, 5000, 128, label="resident-d128-3phase", kprime=8, hint=5)
    ok &= check(130, 700, 768, label="stream-d768")
    ok &= check(64, 520, 2048, label="stream-d2048")
    print("DIAG", "PASS" if ok else "FAIL", f"{time.time() - t0:.1f}s", flush=True)
    if ok and "--quick" not in sys.argv:
        bench(10000, 300000, 512)
        bench(10000, 300000, 256)
        bench(10000, 300000, 128)
        bench(10000, 300000, 768)
        bench(1000, 10000, 2048, iters=20)
