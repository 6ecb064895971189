"""Row-local manifold kernels at the reference's sizes -- ~14k label embeddings of D = 128, 50k hierarchy
pairs, 100k figure pairs -- against the same arithmetic op by op through the geoopt shim on the GPU (what the reference's
code would execute): forward + backward of each loss term, and one RiemannianAdam step.  python tools/bench_manifold.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import manifold  # noqa: E402
from patent_image_retrieval_b200 import geoopt_shim as gt  # noqa: E402
from patent_image_retrieval_b200.geoopt_shim import pmath  # noqa: E402
from patent_image_retrieval_b200.geoopt_shim.optim import RiemannianAdam  # noqa: E402
from patent_image_retrieval_b200.models import _hmi_terms  # noqa: E402


def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


torch.manual_seed(0)
L, D, c = 14000, 128, 1.0
k = torch.tensor([-c], device="cuda")
ball = gt.PoincareBall(c=c)
emb = pmath.project(pmath.expmap0(torch.randn(L, D, device="cuda") * 0.3, k=k), k=k)
pairs = torch.randint(0, L, (50000, 2), device="cuda")
fig = pmath.project(pmath.expmap0(torch.randn(20000, D, device="cuda") * 0.3, k=k), k=k)
fpairs = torch.randint(0, 20000, (100000, 2), device="cuda")


def fb(loss_fn, x):
    x = x.detach().requires_grad_(True)
    loss_fn(x).backward()


def eager_hmi(e):
    ra, rb, dcen = _hmi_terms(ball, k, e[pairs[:, 0]], e[pairs[:, 1]])
    return torch.relu(0.1 - (ra - rb - dcen)).mean()


def eager_dist0(e):
    d0 = pmath.dist0(e, k=k)
    return (torch.relu(0.5 - d0) + torch.relu(d0 - 3.0)).mean()


rows = [
    ("row-pair distances, 100k pairs (fwd+bwd)",
     lambda: fb(lambda x: manifold.rowpair_dist(x, x, fpairs[:, 0], fpairs[:, 1], k).sum(), fig),
     lambda: fb(lambda x: pmath.dist(x[fpairs[:, 0]], x[fpairs[:, 1]], k=k).sum(), fig)),
    ("HMI insideness hinge, 50k pairs (fwd+bwd)",
     lambda: fb(lambda x: manifold.hmi_pair_loss(x, pairs, k, "insideness", 0.1), emb),
     lambda: fb(eager_hmi, emb)),
    ("dist0 regulariser, 14k labels (fwd+bwd)",
     lambda: fb(lambda x: manifold.dist0_reg_loss(x, k, 0.5, 3.0), emb),
     lambda: fb(eager_dist0, emb)),
]
for name, kern, eager in rows:
    print(f"{name:48s} kernels {timed(kern):8.1f} us   op-by-op {timed(eager):8.1f} us")

for fused in (True, False):
    # the fused kernel serves contiguous fp32 [n,d] ball parameters; a column-major copy of the same values takes the
    # optimiser's op-by-op path
    data = emb.clone() if fused else emb.t().contiguous().t()
    p = gt.ManifoldParameter(data, manifold=ball)
    opt = RiemannianAdam([p], lr=1e-3)
    g = torch.randn(L, D, device="cuda") * 1e-2

    def step():
        p.grad = g
        opt.step()
    print(f"RiemannianAdam step, 14k x 128 ({'fused kernel' if fused else 'op-by-op'}): {timed(step):8.1f} us")
