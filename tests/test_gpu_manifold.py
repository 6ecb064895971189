"""GPU parity: row-local manifold kernels of the train_hyp step (csrc/manifold.cu) -- hierarchy / regulariser losses
against the reference's own models.py (golden_r2.npz: values + autograd gradients), the row-pair distance against fp64
autograd of the oracle, and the fused RiemannianAdam step against the op-by-op restatement of geoopt's optimiser."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pmath as opm
from patent_image_retrieval_b200 import geoopt_shim as gs, manifold, models, train

pytestmark = pytest.mark.gpu
G2 = Path(__file__).resolve().parent / "golden" / "golden_r2.npz"


def test_hierarchy_and_reg_losses_match_reference_on_gpu():
    g = np.load(G2)
    c = float(g["refshim2_c"])
    m = models.HyperbolicEmbeddingModel(32, 16, label_num=60, hidden_dims=[24], c=c).cuda()
    with torch.no_grad():       # fp64 parameter, as in the reference (default dtype flipped at import): projx eps 1e-5
        m.label_emb.data = torch.from_numpy(g["refshim2_label_emb"]).cuda()
    imp, exc = torch.from_numpy(g["refshim2_imp"]).cuda(), torch.from_numpy(g["refshim2_exc"]).cuda()
    figs = torch.from_numpy(g["refshim2_figs"]).cuda().requires_grad_(True)
    inside, disjoint = m.calculate_hierarchical_loss(imp, exc)
    label_reg, instance_reg = m.calculate_reg_loss(figs)
    # instance_reg: three figure rows sit 1e-5 inside the boundary (fp64 projx); the kernel reads fp32 rows, and
    # artanh near 1 amplifies their 6e-8 rounding
    for got, key, tol in ((inside, "inside", 5e-6), (disjoint, "disjoint", 5e-6), (label_reg, "label_reg", 5e-6),
                          (instance_reg, "instance_reg", 1e-4)):
        np.testing.assert_allclose(got.item(), float(g["refshim2_" + key]), rtol=tol, err_msg=key)
    # three label rows lie outside the ball (x40): fp32 rows of the golden differ from the fp64 ones by 1e-7 relative,
    # and the clip Jacobian divides by small margins -- compare against the gradient scale
    for loss, wrt, key in ((inside, m.label_emb, "inside_grad"), (disjoint, m.label_emb, "disjoint_grad"),
                           (label_reg, m.label_emb, "label_reg_grad"), (instance_reg, figs, "instance_reg_grad")):
        got = torch.autograd.grad(loss, wrt, retain_graph=True)[0].cpu().double().numpy()
        want = g["refshim2_" + key]
        assert np.abs(got - want).max() <= (2e-3 if key == "instance_reg_grad" else 2e-5) * np.abs(want).max() + 1e-9, key
    v_in = manifold.hmi_values(m.label_emb, imp, m.k, "insideness").cpu().double().numpy()
    v_dj = manifold.hmi_values(m.label_emb, exc, m.k, "disjointedness").cpu().double().numpy()
    np.testing.assert_allclose(v_in, g["refshim2_insideness"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(v_dj, g["refshim2_disjointedness"], rtol=2e-5, atol=2e-6)
    with pytest.raises(IndexError):
        m.calculate_hierarchical_loss(torch.tensor([[0, 60]], device="cuda"), None)


@pytest.mark.parametrize("d,c", [(128, 1.0), (20, 0.5), (256, 2.0)])
def test_rowpair_distance_forward_backward(d, c):
    torch.manual_seed(2)
    k = torch.tensor(-c, dtype=torch.float64)
    x = opm.project(opm.expmap0(torch.randn(50, d, dtype=torch.float64) * 0.6 / d ** 0.5, k=k), k=k)
    y = opm.project(opm.expmap0(torch.randn(70, d, dtype=torch.float64) * 0.6 / d ** 0.5, k=k), k=k)
    y[3] = x[5] * (1 + 1e-2)                                         # a near pair
    ia = torch.randint(0, 50, (400,))
    ib = torch.randint(0, 70, (400,))
    ia[0], ib[0] = 5, 3
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    want = opm.dist(xr[ia], yr[ib], k=k)
    w = torch.randn(400, dtype=torch.float64)
    (want * w).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    yg = y.float().cuda().requires_grad_(True)
    got = manifold.rowpair_dist(xg, yg, ia.cuda(), ib.cuda(), torch.tensor([-c]))
    (got * w.float().cuda()).sum().backward()
    # fp32 inputs: the near pair's distance is conditioned by the fp32 rounding of the points themselves
    far = torch.ones(400, dtype=torch.bool)
    far[0] = False
    far = (ia != 5) | (ib != 3)
    assert float(((got.detach().cpu().double() - want.detach()).abs() / want.detach())[far].max()) < 5e-6
    assert float((xg.grad.cpu().double() - xr.grad).abs().max()) < 1e-4 * float(xr.grad.abs().max())
    assert float((yg.grad.cpu().double() - yr.grad).abs().max()) < 1e-4 * float(yr.grad.abs().max())


def test_pair_losses_use_the_rowpair_kernel_and_match_cpu_path():
    torch.manual_seed(4)
    m = models.FigureOnlyHyperbolicModel(64, 32, hidden_dims=[48], c=1.0).eval()
    x = torch.randn(40, 64) * 0.5
    pos = torch.randint(0, 40, (30, 2))
    neg = torch.randint(0, 40, (60, 2))
    pos[:, 1] = (pos[:, 0] + 1 + pos[:, 1] % 39) % 40                 # no (i, i) pairs: dist(x, x) has no gradient
    neg[:, 1] = (neg[:, 0] + 1 + neg[:, 1] % 39) % 40
    emb = m.encode_figures(x).detach()
    # truth: the op-by-op CPU path in fp64 (its fp32 autograd is itself only good to ~1e-3 on close pairs)
    emb_c = emb.double().requires_grad_(True)
    m.k = m.k.double()
    want = m.calculate_pair_loss(emb_c, pos, neg)
    want.backward()
    mg = m.cuda()
    mg.k = mg.k.float()
    emb_g = emb.cuda().requires_grad_(True)
    got = mg.calculate_pair_loss(emb_g, pos.cuda(), neg.cuda())
    assert float(got) == pytest.approx(float(want), rel=2e-5)
    got.backward()
    assert float((emb_g.grad.cpu().double() - emb_c.grad).abs().max()) < 1e-4 * float(emb_c.grad.abs().max())


def test_fused_riemannian_adam_step_matches_opwise_optimiser():
    torch.manual_seed(6)
    c = 0.8
    ball = gs.PoincareBall(c=c)
    pts = gs.pmath.expmap0(torch.randn(300, 64) * 0.08, k=ball.k)
    pts[:5] *= 7.0                                                     # near the boundary: the retraction clips
    pts = gs.pmath.project(pts, k=ball.k)
    p_cpu = gs.ManifoldParameter(pts.clone(), manifold=ball)
    p_gpu = gs.ManifoldParameter(pts.clone().cuda(), manifold=gs.PoincareBall(c=c).cuda())
    o_cpu = gs.optim.RiemannianAdam([p_cpu], lr=5e-2, weight_decay=1e-3)
    o_gpu = gs.optim.RiemannianAdam([p_gpu], lr=5e-2, weight_decay=1e-3)
    for step in range(4):
        gsd = torch.randn(300, 64, generator=torch.Generator().manual_seed(step))
        p_cpu.grad, p_gpu.grad = gsd.clone(), gsd.clone().cuda()
        o_cpu.step()
        o_gpu.step()
        assert float((p_gpu.detach().cpu() - p_cpu.detach()).abs().max()) < 2e-6
        st_c, st_g = o_cpu.state[p_cpu], o_gpu.state[p_gpu]
        scale = float(st_c["exp_avg"].abs().max())
        assert float((st_g["exp_avg"].cpu() - st_c["exp_avg"]).abs().max()) < 1e-5 * scale
        torch.testing.assert_close(st_g["exp_avg_sq"].cpu(), st_c["exp_avg_sq"].expand(300, 64), rtol=1e-4, atol=1e-12)
    assert ball.check_point_on_manifold(p_gpu.detach().cpu())


def test_train_hyperbolic_contrastive_returns_the_model_like_the_reference(tmp_path):
    """src/train.py:1792-1910 returns the model (its caller: src/train.py:3900); the history rides on the model."""
    torch.manual_seed(8)
    m = models.FigureOnlyHyperbolicModel(32, 16, hidden_dims=[24], c=1.0)
    X = torch.randn(64, 32) * 0.5
    pos = {i: [(i + 1) % 64, (i + 7) % 64] for i in range(64)}
    out = train.train_hyperbolic_contrastive(m, X, pos, list(range(48)), list(range(48, 64)), epochs=2, batch_size=16,
                                             lr=1e-3, temperature=0.2, device="cuda", save_path=str(tmp_path / "m.pt"))
    assert isinstance(out, models.FigureOnlyHyperbolicModel) and len(out._train_history) == 2
    out.eval()
    assert out.encode_figures(X[:4].cuda()).shape == (4, 16)
