"""Time of the exact full-scan fallback (hypret_exact_topk) for a few listed queries: python tools/bench_exact.py [N] [D]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g, _, _ = ops.project_rows(synth.gaussian_features(N, D, seed=0, device="cuda"), 1.0, want_operand=False)
sq = ops.row_sqnorm64(g)
for Q in (1, 3, 4, 8, 32):
    q, _, _ = ops.project_rows(synth.gaussian_features(Q, D, seed=1, device="cuda"), 1.0, want_operand=False)
    for _ in range(2):
        ops.exact_topk(q, g, sq, 1.0, "hyperbolic", 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.exact_topk(q, g, sq, 1.0, "hyperbolic", 10)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"N={N} D={D} Q={Q}: {ms:.3f} ms  ({N * D * 4 / ms / 1e6:.0f} GB/s of gallery rows per pass-equivalent)")
    # warm start: the bound a filtered result provides (here the true 16th best distance)
    bound = ops.exact_topk(q, g, sq, 1.0, "hyperbolic", 16)[0][:, 15].contiguous()
    flags = torch.ones(Q, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ops.exact_topk_flagged(q, g, sq, flags, 1.0, "hyperbolic", 10, init_bound=bound)
    e0.record()
    for _ in range(5):
        ops.exact_topk_flagged(q, g, sq, flags, 1.0, "hyperbolic", 10, init_bound=bound)
    e1.record()
    torch.cuda.synchronize()
    print(f"      warm start (init_bound = 16th best): {e0.elapsed_time(e1) / 5:.3f} ms")
