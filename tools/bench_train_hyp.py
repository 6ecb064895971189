"""C5: train_hyp in-batch n x n Poincare distance matrix, forward + backward (BASELINE.json configs[4]).
GPU: hypret_pairdist_ce_fwd + hypret_pairdist_ce_bwd (+ two cuBLAS fp32 GEMMs) through train.in_batch_contrastive_loss.
CPU: the reference's literal double loop (src/train.py:1833-1840) at its own batch sizes, restated by the oracle."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from patent_image_retrieval_b200 import ops, synth, train  # noqa: E402
from patent_image_retrieval_b200.geoopt_shim import pmath  # noqa: E402


def gpu_case(n, d, c=0.5, tau=0.07, iters=10):
    k = torch.tensor([-c])
    mu = synth.gaussian_features(n, d, seed=2, scale=1.0, device="cuda")
    a = pmath.project(pmath.expmap0(mu + 0.1 * synth.gaussian_features(n, d, seed=3, scale=1.0, device="cuda"), k=k), k=k)
    p = pmath.project(pmath.expmap0(mu + 0.1 * synth.gaussian_features(n, d, seed=4, scale=1.0, device="cuda"), k=k), k=k)
    a.requires_grad_(True)
    p.requires_grad_(True)

    def step():
        a.grad = p.grad = None
        loss = train.in_batch_contrastive_loss(a, p, k, tau)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # forward kernel alone
    e0.record()
    for _ in range(iters):
        ops.pairdist(a.detach(), p.detach(), c)
    e1.record()
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / iters
    return {"n": n, "d": d, "ms_fwd_bwd": ms, "ms_pairdist_kernel": fwd_ms, "algorithmic_tflops": 6.0 * n * n * d / ms / 1e9,
            "loss": float(loss), "grad_finite": bool(torch.isfinite(a.grad).all() and torch.isfinite(p.grad).all())}


def cpu_case(n, d, c=0.5, tau=0.07):
    from oracle import contrastive, head
    kk = torch.tensor([-c])
    a = head.embed_rows(synth.gaussian_features(n, d, seed=3, scale=1.0), c).requires_grad_(True)
    p = head.embed_rows(synth.gaussian_features(n, d, seed=4, scale=1.0), c).requires_grad_(True)
    t0 = time.perf_counter()
    loss = contrastive.contrastive_loss(a, p, kk, temperature=tau, loop=True)
    loss.backward()
    return {"n": n, "d": d, "s_fwd_bwd_double_loop": time.perf_counter() - t0}


if __name__ == "__main__":
    out = {"gpu": [gpu_case(8192, 128), gpu_case(8192, 256), gpu_case(128, 256, iters=50)],
           "cpu_reference_double_loop": [cpu_case(64, 256), cpu_case(128, 256)]}
    c = out["cpu_reference_double_loop"][-1]
    out["cpu_extrapolated_to_n8192_s"] = c["s_fwd_bwd_double_loop"] * (8192 / c["n"]) ** 2
    for g in out["gpu"]:
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in g.items()}))
    print(json.dumps({"cpu_reference_double_loop": out["cpu_reference_double_loop"],
                      "cpu_extrapolated_to_n8192_s": out["cpu_extrapolated_to_n8192_s"]}))
