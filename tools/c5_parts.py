"""Per-kernel times of the flash-style train_hyp step (bench.py --workload c5 shape): python tools/c5_parts.py [n] [D]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import ops, synth, train  # noqa: E402
from patent_image_retrieval_b200.geoopt_shim import pmath  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
c, tau, dev = 0.5, 0.07, "cuda"
kk = torch.tensor([-c])
mu = synth.gaussian_features(n, D, seed=2, scale=1.0, device=dev)
mk = lambda s: pmath.project(pmath.expmap0(mu + 1.0 * synth.gaussian_features(n, D, seed=s, scale=1.0, device=dev), k=kk), k=kk)
a, p = mk(3), mk(4)
inv_tau = 1.0 / tau


def timed(name, fn, reps=20):
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:42s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us")
    return out


ao = timed("flash_prep (row+col+t+sq), one operand", lambda: ops.FlashOperands(a))
po = ops.FlashOperands(p)
row_lse = timed("flash_lse (rows)", lambda: ops.flash_lse(ao, po, c, inv_tau))
timed("diag distances (rowpair_dist)", lambda: train._diag_distance(a, p, c))
gs = torch.tensor(1.0, device=dev)
timed("flash_grad dA (+finish)", lambda: ops.flash_grad(ao, po, c, inv_tau, row_lse, None, 1.0, 0.0, grad_scale=gs))
timed("flash_grad dP (+finish)", lambda: ops.flash_grad(po, ao, c, inv_tau, None, row_lse, 0.0, 1.0, grad_scale=gs))
ar, pr = a.clone().requires_grad_(True), p.clone().requires_grad_(True)


def step():
    ar.grad = pr.grad = None
    train.in_batch_contrastive_loss(ar, pr, kk, tau).backward()


timed("whole step (autograd)", step)
