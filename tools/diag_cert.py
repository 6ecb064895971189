"""Certificate diagnostics: distribution of the margin (k'-th best filter score - exact surrogate of the k-th result)
against the rounding bound E, per gallery size / k'.  python tools/diag_cert.py [N ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import GalleryIndex, ops, synth  # noqa: E402

import os
d, c, k, Q = 512, 1.0, 10, int(os.environ.get('DIAG_Q', 2048))
for N in [int(a) for a in sys.argv[1:]] or [30_000, 300_000]:
    index = GalleryIndex(synth.gallery_rows(0, N, d, "cuda"), c=c)
    qry = synth.gaussian_features(Q, d, seed=1, device="cuda")
    for kp, kb in ((16, 16), (16, 24), (24, 24), (32, 32)):
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            q32, cs, ci, cnt, q_err = index.score_candidates(qry, k=k, kprime=kp, want_err=True, kbound=kb)
            e1.record()
            torch.cuda.synchronize()
        print(f"   score_candidates k'={kp} kbound={kb}: {e0.elapsed_time(e1):.2f} ms")
        bufs = ops.CertBuffers(Q, index.device)
        _, _, margin = ops.rerank_cert(q32, index.rows32, cs, ci, c, "hyperbolic", k, q_err, index.stats,
                                       index.rows_sq64, bufs, list_count=cnt, fallback=False, want_margin=True,
                                       ksel=kb if kb > kp else 0)
        st = index.stats.double()
        qn = q32.double().norm(dim=1)
        slack = (ops.operand_kpad(d) / 16 + 8) * 2.0 ** -22
        E = q_err.double() * st[0] + qn * st[1] + slack * (qn * st[0] + qn * qn * st[2] + st[3])
        qs = torch.tensor([0.01, 0.1, 0.5, 0.9], device="cuda", dtype=torch.float64)
        print(f"N={N} k'={kp} kbound={kb}: certified {float(bufs.certified[:Q].float().mean()):.4f}  stats {index.stats.tolist()}\n"
              f"   margin quantiles {torch.quantile(margin.double(), qs).tolist()}\n"
              f"   E quantiles      {torch.quantile(E, qs).tolist()}  q_err median {float(q_err.median()):.3e}")
