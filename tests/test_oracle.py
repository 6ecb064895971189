"""CPU: the oracle against (a) vectors produced by the reference's own code (ref_*),
(b) its own frozen fp64 known answers (kat_*, parity unpinned), (c) mathematical identities."""
import math

import numpy as np
import pytest
import torch

from oracle import contrastive, head, pmath, retrieval

CS = (1.0, 0.5, 2.0)


def _tag(c):
    return str(c).replace(".", "p")


# ------------------------------------------------------------------ reference-pinned pieces
def test_cosine_matches_sklearn_golden(golden):
    sim = retrieval.cosine_similarity(golden["ref_cos_q"], golden["ref_cos_g"])
    np.testing.assert_allclose(sim, golden["ref_cos_sim"], rtol=0, atol=2e-6)
    assert np.all(sim[:, 7] == 0.0)      # zero gallery row


def test_notebook_metrics_match_reference(golden):
    order = golden["ref_cos_order"]
    off, items = golden["ref_nb_pos_offsets"], golden["ref_nb_pos_items"]
    pos = [items[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    _, per = retrieval.notebook_metrics(order, pos, ks=(5, 10, 20))
    np.testing.assert_allclose(per["ap"], golden["ref_nb_ap_scores"], rtol=1e-12)
    np.testing.assert_allclose(per["ndcg"], golden["ref_nb_ndcg_scores"], rtol=1e-12)
    np.testing.assert_allclose(per["mrr"], golden["ref_nb_reciprocal_ranks"], rtol=1e-12)
    for k in (5, 20):
        np.testing.assert_allclose(per[f"mrr@{k}"], golden[f"ref_nb_reciprocal_ranks_{k}"], rtol=1e-12)
    for k in (5, 10, 20):
        np.testing.assert_allclose(per[f"recall@{k}"], golden[f"ref_nb_recall_{k}"], rtol=1e-12)
        np.testing.assert_allclose(per[f"precision@{k}"], golden[f"ref_nb_precision_{k}"], rtol=1e-12)


def test_sklearn_ap_restatement_with_ties(golden):
    s, t = golden["ref_ap_scores"], golden["ref_ap_target"]
    got = [retrieval.average_precision_sklearn(t[i], s[i]) for i in range(s.shape[0])]
    np.testing.assert_allclose(got, golden["ref_ap_values"], rtol=1e-12)
    # auxiliary.mean_average_precision: mean over label columns with >= 1 positive
    np.testing.assert_allclose(np.mean(got), float(golden["ref_aux_map"]), rtol=1e-12)


def test_ap_conventions_agree_without_ties():
    rng = np.random.default_rng(0)
    scores = rng.standard_normal(500)
    target = (rng.random(500) < 0.03).astype(np.float64)
    target[3] = 1
    ranked = list(np.argsort(-scores, kind="stable"))
    pos = set(np.nonzero(target)[0].tolist())
    assert math.isclose(retrieval.average_precision_ranked(ranked, pos),
                        retrieval.average_precision_sklearn(target, scores), rel_tol=1e-12)


# ------------------------------------------------------------------ frozen KATs (fp64)
@pytest.mark.parametrize("c", CS)
def test_pmath_kats(golden, c):
    t = _tag(c)
    k = torch.tensor(-c, dtype=torch.float64)
    u = torch.from_numpy(golden[f"kat_u_c{t}"])
    x = torch.from_numpy(golden[f"kat_x_c{t}"])
    y = torch.from_numpy(golden[f"kat_y_c{t}"])
    w = torch.from_numpy(golden[f"kat_w_c{t}"])
    b = torch.from_numpy(golden[f"kat_b_c{t}"])
    tt = lambda name: torch.from_numpy(golden[f"kat_{name}_c{t}"])
    torch.testing.assert_close(pmath.project(pmath.expmap0(u, k=k), k=k), x, rtol=1e-13, atol=0)
    torch.testing.assert_close(pmath.dist(x[:, None], y[None], k=k), tt("dist"), rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(pmath.dist0(x, k=k), tt("dist0"), rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(pmath.mobius_add(x, y, k=k), tt("madd"), rtol=1e-12, atol=1e-15)
    torch.testing.assert_close(pmath.mobius_matvec(w, x, k=k), tt("matvec"), rtol=1e-12, atol=1e-15)
    torch.testing.assert_close(pmath.mobius_fn_apply(torch.tanh, x, k=k), tt("tanh"), rtol=1e-12, atol=1e-15)
    torch.testing.assert_close(head.mobius_linear(x, w, b, hyperbolic_input=True, k=k), tt("mlin"), rtol=1e-12,
                               atol=1e-15)
    torch.testing.assert_close(head.mobius_linear(u, w, b, hyperbolic_input=False, k=k), tt("mlin_e"), rtol=1e-12,
                               atol=1e-15)


# ------------------------------------------------------------------ identities
@pytest.mark.parametrize("c", CS)
def test_distance_identities(c):
    torch.manual_seed(0)
    k = torch.tensor(-c, dtype=torch.float64)
    x = pmath.project(pmath.expmap0(torch.randn(40, 24, dtype=torch.float64) * 0.3, k=k), k=k)
    y = pmath.project(pmath.expmap0(torch.randn(40, 24, dtype=torch.float64) * 0.3, k=k), k=k)
    z = pmath.project(pmath.expmap0(torch.randn(40, 24, dtype=torch.float64) * 0.3, k=k), k=k)
    dxy, dyx = pmath.dist(x, y, k=k), pmath.dist(y, x, k=k)
    torch.testing.assert_close(dxy, dyx, rtol=1e-12, atol=1e-14)                         # symmetry
    assert float(pmath.dist(x, x, k=k).abs().max()) < 1e-7                               # d(x,x)=0 (artanh clamp)
    torch.testing.assert_close(pmath.dist0(x, k=k), pmath.dist(torch.zeros_like(x), x, k=k), rtol=1e-12, atol=1e-14)
    assert bool((pmath.dist(x, z, k=k) <= dxy + pmath.dist(y, z, k=k) + 1e-12).all())     # triangle
    torch.testing.assert_close(pmath.dist_arcosh(x, y, k=k), dxy, rtol=1e-10, atol=1e-12)  # closed form
    # closed form of dist0: 2/sqrt(c) artanh(sqrt(c)|x|)
    ref0 = 2 / math.sqrt(c) * torch.atanh(math.sqrt(c) * x.norm(dim=-1))
    torch.testing.assert_close(pmath.dist0(x, k=k), ref0, rtol=1e-12, atol=1e-14)
    # expmap0 / logmap0 inverse
    u = torch.randn(40, 24, dtype=torch.float64) * 0.2
    torch.testing.assert_close(pmath.logmap0(pmath.expmap0(u, k=k), k=k), u, rtol=1e-10, atol=1e-12)


def test_monotone_surrogate_same_ranking():
    """s_ij / (1 - c|y_j|^2) orders the gallery like the distance (what the GPU filter relies on)."""
    torch.manual_seed(1)
    c = 1.0
    q = head.embed_rows(torch.randn(5, 32, dtype=torch.float64) * 0.08, c)
    g = head.embed_rows(torch.randn(400, 32, dtype=torch.float64) * 0.08, c)
    d = retrieval.hyperbolic_dist_rows(q, g, c, form="arcosh")
    sur = torch.cdist(q, g).pow(2) / (1 - c * g.pow(2).sum(-1))[None]
    assert torch.equal(torch.argsort(d, dim=1, stable=True), torch.argsort(sur, dim=1, stable=True))


def test_project_clip_and_fp32_eps():
    u = torch.randn(8, 16) * 5.0
    x = head.embed_rows(u, 1.0)
    assert float(x.norm(dim=-1).max()) <= (1 - 4e-3) + 1e-6
    x64 = head.embed_rows(u.double(), 1.0)
    assert float(x64.norm(dim=-1).max()) <= (1 - 1e-5) + 1e-12
    assert pmath.check_point_on_manifold(x, torch.tensor(-1.0))


def test_fp32_geoopt_form_vs_fp64_truth_well_conditioned():
    torch.manual_seed(2)
    d = 512
    u = torch.randn(64, d) * (0.45 / d ** 0.5)
    v = torch.randn(2000, d) * (0.45 / d ** 0.5)
    q32, g32 = head.embed_rows(u, 1.0), head.embed_rows(v, 1.0)
    d32 = retrieval.hyperbolic_dist_rows(q32, g32, 1.0, form="geoopt")
    d64 = retrieval.hyperbolic_dist_rows(q32.double(), g32.double(), 1.0, form="arcosh")
    rel = ((d32.double() - d64).abs() / d64).max()
    assert float(rel) < 5e-6


def test_contrastive_loop_equals_broadcast_and_grad():
    torch.manual_seed(3)
    k = torch.tensor([-0.5], dtype=torch.float64)
    a = head.embed_rows(torch.randn(6, 8, dtype=torch.float64) * 0.3, 0.5).requires_grad_(True)
    p = head.embed_rows(torch.randn(6, 8, dtype=torch.float64) * 0.3, 0.5).requires_grad_(True)
    l1 = contrastive.contrastive_loss(a, p, k, 0.1, loop=True)
    l2 = contrastive.contrastive_loss(a, p, k, 0.1, loop=False)
    torch.testing.assert_close(l1, l2, rtol=1e-12, atol=0)
    ls = contrastive.contrastive_loss(a, p, k, 0.07, symmetric=True)
    ls.backward()
    assert torch.isfinite(a.grad).all() and torch.isfinite(p.grad).all()


def test_topk_tie_policy():
    d = torch.tensor([[0.5, 0.1, 0.1, 0.7, 0.1]])
    v, i = retrieval.topk_smallest(d, 3)
    assert i.tolist() == [[1, 2, 4]]
