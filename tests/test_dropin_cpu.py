"""CPU: the oracle and the drop-in host layer against vectors produced by the reference's OWN
``models.py`` / ``train.py`` code (imported from /root/reference/src by tests/golden/make_golden.py,
with geoopt replaced by this repo's shim -- ``refshim_*``), plus the world-size-2 gloo test of the
multi-GPU host logic."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import contrastive, head, retrieval


def _t(golden, key, dtype=None):
    t = torch.from_numpy(np.array(golden[key]))
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("c", [1.0, 0.5])
def test_oracle_head_matches_reference_models_py(golden, c):
    tag = str(c).replace(".", "p")
    x = _t(golden, f"refshim_head_x_c{tag}")
    w1 = _t(golden, f"refshim_head_encoder.first_layer.weight_c{tag}")
    b1 = _t(golden, f"refshim_head_encoder.first_layer.bias_c{tag}")
    w2 = _t(golden, f"refshim_head_encoder.final_layer.weight_c{tag}")
    b2 = _t(golden, f"refshim_head_encoder.final_layer.bias_c{tag}")
    want = _t(golden, f"refshim_head_y_c{tag}")
    k = torch.tensor([-c], dtype=torch.float32)
    got = head.encoder_forward(x, w1, b1, w2, b2, k)       # fp32 input, fp64 weights cast to fp32 (models.py:301-303)
    assert want.dtype == torch.float32
    torch.testing.assert_close(got, want, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("c", [1.0, 0.5])
def test_dropin_model_loads_reference_state_dict(golden, c):
    from patent_image_retrieval_b200 import models
    tag = str(c).replace(".", "p")
    m = models.FigureOnlyHyperbolicModel(32, 16, hidden_dims=[24], c=c, dropout_rate=0.3).eval()
    # geoopt-shaped key set (PoincareBall is an nn.Module there: curvature parameter ``isp_c`` per ball)
    ball_keys = ["ball.isp_c", "encoder.ball.isp_c", "encoder.final_layer.ball.isp_c", "encoder.first_layer.ball.isp_c"]
    weight_keys = ["encoder.final_layer.bias", "encoder.final_layer.weight", "encoder.first_layer.bias",
                   "encoder.first_layer.weight"]
    assert sorted(m.state_dict().keys()) == sorted(ball_keys + weight_keys)
    sd = {k: _t(golden, f"refshim_head_{k}_c{tag}") for k in weight_keys}
    m.load_state_dict(sd)                      # a checkpoint without the curvature keys (round-1 layout) still loads
    # ... and so does one with them, as the real reference writes it: isp_c = log(exp(c) - 1), 0-dim
    ref_style = dict(sd, **{k: torch.tensor(c, dtype=torch.float64).exp().sub(1).log() for k in ball_keys})
    m.load_state_dict(ref_style, strict=True)
    assert float(m.ball.c) == pytest.approx(c, rel=1e-6) and float(m.encoder.first_layer.ball.k) == pytest.approx(-c, rel=1e-6)
    with torch.no_grad():
        y = m(_t(golden, f"refshim_head_x_c{tag}"))
    torch.testing.assert_close(y, _t(golden, f"refshim_head_y_c{tag}"), rtol=2e-6, atol=1e-7)
    assert m.k.dtype == torch.float32 and "k" not in m.state_dict()
    full = models.HyperbolicEmbeddingModel(32, 16, label_num=7, hidden_dims=[24], c=c)
    assert "label_emb" in full.state_dict() and full.label_emb.shape == (7, 16)
    assert "ball.isp_c" in full.state_dict() and not full.ball.isp_c.requires_grad


def test_oracle_contrastive_matches_reference_train_py(golden):
    k = torch.tensor([-0.5], dtype=torch.float64)
    a = _t(golden, "refshim_hcl_a").requires_grad_(True)
    p = _t(golden, "refshim_hcl_p").requires_grad_(True)
    loss = contrastive.contrastive_loss(a, p, k, temperature=0.07, symmetric=True)
    loss.backward()
    # the reference writes fp64 distances into a float32-default... here: default fp64 at its import
    np.testing.assert_allclose(loss.item(), float(golden["refshim_hcl_loss"]), rtol=1e-10)
    np.testing.assert_allclose(a.grad.numpy(), golden["refshim_hcl_da"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(p.grad.numpy(), golden["refshim_hcl_dp"], rtol=1e-8, atol=1e-12)
    s2p = contrastive.sample_to_prototype_loss(_t(golden, "refshim_s2p_s"), _t(golden, "refshim_s2p_pos"),
                                               _t(golden, "refshim_s2p_neg"), 3, k, margin=0.1)
    np.testing.assert_allclose(s2p.item(), float(golden["refshim_s2p_loss"]), rtol=1e-12)


def _eval_case(golden):
    f2p = {int(k): v for k, v in json.loads(bytes(golden["refshim_eval_f2p_json"]).decode()).items()}
    X = _t(golden, "refshim_eval_X")
    sd = {k[len("refshim_eval_"):]: _t(golden, k) for k in golden.files
          if k.startswith("refshim_eval_encoder") or k == "refshim_eval_label_emb"}
    return X, sd, f2p, {"patents": 0, "medium_cpcs": 45, "big_cpcs": 55, "main_cpcs": 58}


def test_oracle_evaluate_retrieval_matches_reference(golden):
    X, sd, f2p, _ = _eval_case(golden)
    k = torch.tensor([-2.0], dtype=torch.float32)
    emb = head.encoder_forward(X, sd["encoder.first_layer.weight"], sd["encoder.first_layer.bias"],
                               sd["encoder.final_layer.weight"], sd["encoder.final_layer.bias"], k)
    patents = sd["label_emb"][:45]
    pos = []
    for i in range(40):
        e = f2p.get(i, -1)
        pos.append(e if isinstance(e, list) else ([e] if e != -1 else []))
    # reference: fp32 query against fp64 label embeddings -> promoted to fp64 (src/train.py:3259)
    got = retrieval.evaluate_retrieval(emb.double(), patents, pos, 2.0, form="geoopt")
    np.testing.assert_allclose(got, float(golden["refshim_eval_map"]), rtol=1e-9)
    assert float(golden["refshim_eval_empty"]) == 0.0 and float(golden["refshim_eval_no_offset"]) == -1.0


def test_shim_matches_oracle_and_installs_as_geoopt():
    from patent_image_retrieval_b200 import geoopt_shim as gs
    from oracle import pmath as op
    torch.manual_seed(0)
    k = torch.tensor(-0.7, dtype=torch.float64)
    x = op.project(op.expmap0(torch.randn(9, 10, dtype=torch.float64) * 0.3, k=k), k=k)
    y = op.project(op.expmap0(torch.randn(9, 10, dtype=torch.float64) * 0.3, k=k), k=k)
    w = torch.randn(6, 10, dtype=torch.float64)
    torch.testing.assert_close(gs.pmath.dist(x, y, k=k), op.dist(x, y, k), rtol=0, atol=0)
    torch.testing.assert_close(gs.pmath.mobius_matvec(w, x, k=k), op.mobius_matvec(w, x, k), rtol=0, atol=0)
    torch.testing.assert_close(gs.pmath.mobius_fn_apply(torch.tanh, x, k=k), op.mobius_fn_apply(torch.tanh, x, k=k),
                               rtol=0, atol=0)
    ball = gs.PoincareBall(c=0.7)
    assert ball.check_point_on_manifold(x.float()) and float(ball.k) == pytest.approx(-0.7)
    with pytest.raises(ValueError):
        ball.assert_check_point_on_manifold(torch.ones(2, 4))
    p = gs.ManifoldParameter(x.clone(), manifold=ball)
    opt = gs.optim.RiemannianAdam([p], lr=1e-2)
    before = gs.pmath.dist(p[:4], p[4:8], k=ball.k).sum()
    before.backward()
    opt.step()
    after = gs.pmath.dist(p[:4], p[4:8], k=ball.k).sum()
    assert float(after.detach()) < float(before.detach()) and ball.check_point_on_manifold(p.detach())
    pts = gs.pmath.expmap0(torch.randn(3, 4) * 0.1, k=ball.k)
    v = torch.randn(3, 4)
    # parallel transport preserves the Riemannian norm
    tv = gs.pmath.parallel_transport(pts, pts.flip(0), v, k=ball.k)
    n0 = gs.pmath.lambda_x(pts, k=ball.k) * v.norm(dim=-1)
    n1 = gs.pmath.lambda_x(pts.flip(0), k=ball.k) * tv.norm(dim=-1)
    torch.testing.assert_close(n0, n1, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ multi-GPU host logic on gloo
def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from patent_image_retrieval_b200.dist import _explicitly_sharded, gather_candidates, shard_range
    # the full-ranking metrics reduce across ranks only when the caller says the rows are sharded --
    # evaluate_retrieval (whole patent table on every rank) must stay collective-free under an initialised group
    assert not _explicitly_sharded(None, None) and not _explicitly_sharded(False, None)
    assert _explicitly_sharded(True, None) and _explicitly_sharded(None, dist.group.WORLD)
    torch.manual_seed(3)
    Q, N, D, k, c = 12, 301, 16, 5, 1.0
    qu = torch.randn(Q, D) * 0.2
    gu = torch.randn(N, D) * 0.2
    lo, hi = shard_range(N, rank, world)
    qp, gp = head.embed_rows(qu, c), head.embed_rows(gu[lo:hi], c)
    d, i = retrieval.hyperbolic_topk(qp, gp, c, k, form="arcosh")       # this rank's shard (oracle as the shard searcher)
    gs, gi = gather_candidates(d, i + lo)
    q.put((rank, lo, hi, gs.numpy(), gi.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_merge_equals_single_shard():
    world, port = 2, 29531 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=60) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 151, 151, 301)
    np.testing.assert_array_equal(res[0][3], res[1][3])      # every rank holds all lists after the gather
    np.testing.assert_array_equal(res[0][4], res[1][4])
    gs, gi = torch.from_numpy(res[0][3]), torch.from_numpy(res[0][4])      # [W,Q,k]
    W, Q, k = gs.shape
    flat_s = gs.permute(1, 0, 2).reshape(Q, W * k)
    flat_i = gi.permute(1, 0, 2).reshape(Q, W * k)
    order = np.lexsort((flat_i.numpy(), flat_s.numpy()), axis=1)[:, :k]
    merged_i = np.take_along_axis(flat_i.numpy(), order, 1)
    torch.manual_seed(3)
    qu = torch.randn(Q, 16) * 0.2
    gu = torch.randn(301, 16) * 0.2
    _, want = retrieval.hyperbolic_topk(head.embed_rows(qu, 1.0), head.embed_rows(gu, 1.0), 1.0, k, form="arcosh")
    np.testing.assert_array_equal(merged_i, want.numpy())


def _gloo_sharded_queries_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from patent_image_retrieval_b200.dist import gather_queries, return_lists_to_owners, shard_range
    Ql, N, D, k, c = 7, 203, 16, 5, 1.0
    gu = torch.randn(N, D, generator=torch.Generator().manual_seed(5)) * 0.2
    qu = torch.randn(Ql, D, generator=torch.Generator().manual_seed(10 + rank)) * 0.2     # this rank's own batch
    lo, hi = shard_range(N, rank, world)
    q_all = gather_queries(qu)                                                            # [W*Ql, D], rank-major
    d, i = retrieval.hyperbolic_topk(head.embed_rows(q_all, c), head.embed_rows(gu[lo:hi], c), c, k, form="arcosh")
    rs, ri = return_lists_to_owners(d, i + lo)                                            # [W, Ql, k] for own queries
    q.put((rank, q_all.numpy(), rs.numpy(), ri.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_queries_alltoall_equals_single_shard():
    """Serving layout (dist.ShardedGalleryIndex.search_sharded): all_gather(query batches) -> shard-local
    search -> all_to_all(lists) -> merge at the owner == searching the whole gallery for the owner's batch."""
    world, port = 2, 29741 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_sharded_queries_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=60) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gu = torch.randn(203, 16, generator=torch.Generator().manual_seed(5)) * 0.2
    g = head.embed_rows(gu, 1.0)
    np.testing.assert_array_equal(res[0][1], res[1][1])                   # same gathered batch everywhere
    for rank, q_all, rs, ri in res:
        W, Ql, k = rs.shape
        own = torch.randn(Ql, 16, generator=torch.Generator().manual_seed(10 + rank)) * 0.2
        np.testing.assert_array_equal(q_all[rank * Ql:(rank + 1) * Ql], own.numpy())
        flat_s = np.transpose(rs, (1, 0, 2)).reshape(Ql, W * k)
        flat_i = np.transpose(ri, (1, 0, 2)).reshape(Ql, W * k)
        order = np.lexsort((flat_i, flat_s), axis=1)[:, :k]
        _, want = retrieval.hyperbolic_topk(head.embed_rows(own, 1.0), g, 1.0, k, form="arcosh")
        np.testing.assert_array_equal(np.take_along_axis(flat_i, order, 1), want.numpy())


def test_shard_range_partitions():
    from patent_image_retrieval_b200.dist import shard_range
    for n, w in [(10, 3), (300000, 8), (7, 8), (10_000_000, 4)]:
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_hierarchy_and_regulariser_losses_match_reference_models_py():
    """calculate_hierarchical_loss / calculate_reg_loss / _hmi_* (src/models.py:550-674) on the CPU path against
    values and autograd gradients produced by the reference's own models.py (tests/golden/make_golden_r2.py)."""
    from pathlib import Path
    from patent_image_retrieval_b200 import models
    g = np.load(Path(__file__).resolve().parent / "golden" / "golden_r2.npz")
    c = float(g["refshim2_c"])
    torch.set_default_dtype(torch.float64)                      # the reference flips the default dtype at import
    try:
        m = models.HyperbolicEmbeddingModel(32, 16, label_num=60, hidden_dims=[24], c=c)
    finally:
        torch.set_default_dtype(torch.float32)
    # geoopt-shaped state dict: the key list of the REFERENCE model under the module-shim
    assert sorted(m.state_dict().keys()) == bytes(g["refshim2_state_dict_keys"]).decode().split("\n")
    with torch.no_grad():
        m.label_emb.data = torch.from_numpy(g["refshim2_label_emb"])
    m.k = m.k.double()
    imp, exc = torch.from_numpy(g["refshim2_imp"]), torch.from_numpy(g["refshim2_exc"])
    figs = torch.from_numpy(g["refshim2_figs"]).requires_grad_(True)
    inside, disjoint = m.calculate_hierarchical_loss(imp, exc)
    label_reg, instance_reg = m.calculate_reg_loss(figs)
    for got, key in ((inside, "inside"), (disjoint, "disjoint"), (label_reg, "label_reg"), (instance_reg, "instance_reg")):
        np.testing.assert_allclose(got.item(), float(g["refshim2_" + key]), rtol=1e-7)
    np.testing.assert_allclose(torch.autograd.grad(inside, m.label_emb, retain_graph=True)[0].numpy(),
                               g["refshim2_inside_grad"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(torch.autograd.grad(label_reg, m.label_emb)[0].numpy(), g["refshim2_label_reg_grad"],
                               rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m._hmi_insideness(m.label_emb[imp[:, 0]], m.label_emb[imp[:, 1]]).detach().numpy(),
                               g["refshim2_insideness"], rtol=1e-7, atol=1e-9)
    with pytest.raises(IndexError):
        m.calculate_hierarchical_loss(torch.tensor([[0, 60]]), None)
    zero = m.calculate_hierarchical_loss(None, torch.zeros(0, 2, dtype=torch.int64))
    assert float(zero[0]) == 0.0 and float(zero[1]) == 0.0
