"""GPU parity: flash-style train_hyp kernels (csrc/flash.cu) -- row log-sum-exps and gradients of the in-batch InfoNCE
without the [n,m] distance matrix -- against fp64 autograd through the oracle's restatement of the reference loss
(src/train.py:1832-1846 rows-only, 2304-2334 symmetric)."""
import pytest
import torch

from oracle import contrastive, head
from oracle import pmath as opm
from patent_image_retrieval_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _batch(n, m, d, c, noise, seed=0):
    mu = synth.gaussian_features(max(n, m), d, seed=seed, scale=1.0)
    a = head.embed_rows((mu[:n] + noise * synth.gaussian_features(n, d, seed=seed + 1, scale=1.0)).double(), c)
    p = head.embed_rows((mu[:m] + noise * synth.gaussian_features(m, d, seed=seed + 2, scale=1.0)).double(), c)
    return a, p


@pytest.mark.parametrize("n,m,d,c,tau", [(300, 520, 128, 1.0, 0.2), (1000, 700, 64, 0.5, 0.5), (129, 257, 32, 2.0, 0.1),
                                         (512, 512, 128, 1.0, 0.07)])
def test_flash_lse_matches_fp64(n, m, d, c, tau):
    a, p = _batch(n, m, d, c, 0.5)
    k = torch.tensor(-c, dtype=torch.float64)
    dm = opm.dist(a[:, None, :], p[None, :, :], k=k)
    want_r = torch.logsumexp(-dm / tau, dim=1)
    want_c = torch.logsumexp(-dm / tau, dim=0)
    ao, po = ops.FlashOperands(a.float().cuda()), ops.FlashOperands(p.float().cuda())
    got_r = ops.flash_lse(ao, po, c, 1.0 / tau).cpu().double()
    got_c = ops.flash_lse(po, ao, c, 1.0 / tau).cpu().double()
    # fp32 inputs + lg2.approx: absolute accuracy of the logits (the loss only sees differences of these)
    assert float((got_r - want_r).abs().max()) < 2e-4
    assert float((got_c - want_c).abs().max()) < 2e-4


@pytest.mark.parametrize("n,d,c,tau,noise,symmetric", [(384, 128, 1.0, 0.2, 0.5, False), (700, 64, 0.5, 0.5, 0.3, True),
                                                       (1100, 128, 1.0, 0.07, 1.0, True), (200, 16, 0.9, 0.1, 0.6, False)])
def test_flash_grad_matches_fp64_autograd(n, d, c, tau, noise, symmetric):
    a, p = _batch(n, n, d, c, noise)
    k = torch.tensor([-c], dtype=torch.float64)
    ar, pr = a.clone().requires_grad_(True), p.clone().requires_grad_(True)
    loss = contrastive.contrastive_loss(ar, pr, k, temperature=tau, symmetric=symmetric)
    loss.backward()
    ao, po = ops.FlashOperands(a.float().cuda()), ops.FlashOperands(p.float().cuda())
    inv_tau = 1.0 / tau
    row_lse = ops.flash_lse(ao, po, c, inv_tau)
    col_lse = ops.flash_lse(po, ao, c, inv_tau) if symmetric else None
    wr, wc = (0.5, 0.5) if symmetric else (1.0, 0.0)
    gs = torch.tensor(1.0, device="cuda")
    da = ops.flash_grad(ao, po, c, inv_tau, row_lse, col_lse, wr, wc, grad_scale=gs).cpu().double()
    dp = ops.flash_grad(po, ao, c, inv_tau, col_lse, row_lse, wc, wr, grad_scale=gs).cpu().double()
    assert float(loss) > 1e-3                                  # not a saturated softmax: the gradients mean something
    for got, want, name in ((da, ar.grad, "dA"), (dp, pr.grad, "dP")):
        err = float((got - want).abs().max()) / float(want.abs().max())
        assert err < 1e-4, f"{name}: {err:.2e} of max |grad|"
