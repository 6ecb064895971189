"""GPU parity: metric kernels against vectors produced by the reference's own notebook code /
sklearn (ref_*), and the evaluation drop-ins against the reference's own train.py (refshim_*)."""
import json

import numpy as np
import pytest
import torch

from oracle import retrieval
from patent_image_retrieval_b200 import evaluation, models, ops, synth

pytestmark = pytest.mark.gpu


def _nb_case(golden):
    order = torch.from_numpy(golden["ref_cos_order"].copy()).cuda()
    off = torch.from_numpy(golden["ref_nb_pos_offsets"]).cuda()
    items_all = golden["ref_nb_pos_items"]
    # positives >= 300 are ground-truth names that are not in the gallery: they count in |P| only
    n_tot = torch.tensor(np.diff(golden["ref_nb_pos_offsets"]), dtype=torch.int32).cuda()
    return order, off, torch.from_numpy(items_all).cuda(), n_tot


def test_notebook_metrics_full_ranking(golden):
    order, off, items, n_tot = _nb_case(golden)
    means, per = ops.retrieval_metrics(order, off, items, ks=(5, 10, 20), n_pos_total=n_tot)
    per = per.cpu().numpy()
    names = ops.metric_names((5, 10, 20))
    col = {n: per[:, i] for i, n in enumerate(names)}
    tol = dict(rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(col["ap"], golden["ref_nb_ap_scores"], **tol)
    np.testing.assert_allclose(col["ndcg"], golden["ref_nb_ndcg_scores"], **tol)
    np.testing.assert_allclose(col["mrr"], golden["ref_nb_reciprocal_ranks"], **tol)
    np.testing.assert_allclose(col["mrr@5"], golden["ref_nb_reciprocal_ranks_5"], **tol)
    np.testing.assert_allclose(col["mrr@20"], golden["ref_nb_reciprocal_ranks_20"], **tol)
    for k in (5, 10, 20):
        np.testing.assert_allclose(col[f"recall@{k}"], golden[f"ref_nb_recall_{k}"], **tol)
        np.testing.assert_allclose(col[f"precision@{k}"], golden[f"ref_nb_precision_{k}"], **tol)
    # means == np.mean of the notebook lists (retrieval.ipynb:446-456)
    np.testing.assert_allclose(means["ap"], golden["ref_nb_ap_scores"].mean(), rtol=1e-12)
    np.testing.assert_allclose(means["recall@10"], golden["ref_nb_recall_10"].mean(), rtol=1e-12)


def test_metrics_on_truncated_lists_match_oracle(golden):
    order, off, items, n_tot = _nb_case(golden)
    K = 20
    means, per = ops.retrieval_metrics(order[:, :K].contiguous(), off, items, ks=(5, 10, 20), n_pos_total=n_tot)
    o, it = golden["ref_nb_pos_offsets"], golden["ref_nb_pos_items"]
    pos = [it[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    want, _ = retrieval.notebook_metrics(golden["ref_cos_order"][:, :K], pos, ks=(5, 10, 20))
    for name, v in means.items():
        assert abs(v - want[name]) < 1e-12, name
    # padded (-1) entries and empty positives
    ranked = torch.tensor([[3, -1, -1, -1], [1, 2, 3, 4]], device="cuda")
    off2 = torch.tensor([0, 1, 1], device="cuda")
    m2, p2 = ops.retrieval_metrics(ranked, off2, torch.tensor([3], device="cuda"), ks=(2, 4))
    p2 = p2.cpu()
    assert p2[0, 0] == 1.0 and p2[0, 1] == 1.0 and p2[0, 4] == 0.0      # precision@2: only 1 valid entry -> 0
    assert float(p2[1].abs().sum()) == 0.0


def test_ap_full_matches_sklearn_and_ranking_conventions(golden):
    s = torch.from_numpy(golden["ref_ap_scores"]).cuda()
    t = golden["ref_ap_target"]
    lists = [np.nonzero(t[i])[0].tolist() for i in range(t.shape[0])]
    off, items = evaluation._csr_from_lists(lists, "cuda")
    mean, ap, valid = ops.ap_full(s, off, items, grouped_ties=True)
    np.testing.assert_allclose(ap.cpu().numpy(), golden["ref_ap_values"], rtol=1e-12)
    np.testing.assert_allclose(mean, float(golden["ref_aux_map"]), rtol=1e-12)     # auxiliary.mean_average_precision
    assert int(valid.sum()) == 20
    _, ap2, _ = ops.ap_full(s, off, items, grouped_ties=False)
    sc = golden["ref_ap_scores"]
    want = [retrieval.average_precision_ranked(list(np.argsort(-sc[i], kind="stable")), set(lists[i])) for i in range(20)]
    np.testing.assert_allclose(ap2.cpu().numpy(), want, rtol=1e-12)
    # rows the reference skips: no positive, NaN score
    s3 = s[:3].clone()
    s3[1, 5] = float("nan")
    off3, items3 = evaluation._csr_from_lists([[], [0], [2, 999]], "cuda")
    mean3, ap3, valid3 = ops.ap_full(s3, off3, items3)
    assert valid3.tolist() == [0, 0, 1]
    assert mean3 == pytest.approx(float(ap3[2]))


@pytest.mark.parametrize("n,m,d,c", [(70, 130, 128, 1.0), (33, 65, 256, 0.5), (64, 64, 20, 2.0)])
def test_pairdist_matches_oracle(n, m, d, c):
    from oracle import head
    a = head.embed_rows(synth.gaussian_features(n, d, seed=1, scale=1.0), c)
    p = head.embed_rows(synth.gaussian_features(m, d, seed=0, scale=1.0), c)
    p[:5] = a[:5] * (1 + 1e-4)                                   # near duplicates: the cancellation case
    got = ops.pairdist(a.cuda(), p.cuda(), c).cpu()
    d32 = retrieval.hyperbolic_dist_rows(a, p, c, form="geoopt")
    d64 = retrieval.hyperbolic_dist_rows(a.double(), p.double(), c, form="arcosh")
    far = d64 > 1e-2
    assert float(((got.double() - d64).abs() / d64)[far].max()) < 2e-6
    # vs the reference's fp32 geoopt form: within 1e-5, plus that form's own error against fp64 where it is
    # ill-conditioned (points near the ball boundary)
    own = (d32.double() - d64).abs()
    assert bool((((got - d32).abs().double()) <= 1e-5 * d64 + 2 * own)[far].all())
    # near-duplicate pairs: the explicit-difference kernel keeps RELATIVE accuracy that the geoopt fp32 form loses
    near = ~far
    assert float(((got.double() - d64).abs() / d64)[near].max()) < 1e-3


def test_contrastive_loss_and_gradients_match_reference_train_py(golden):
    from patent_image_retrieval_b200 import train
    k = torch.tensor([-0.5])
    a = torch.from_numpy(golden["refshim_hcl_a"]).float().cuda().requires_grad_(True)
    p = torch.from_numpy(golden["refshim_hcl_p"]).float().cuda().requires_grad_(True)
    loss = train.hyperbolic_contrastive_loss(a, p, k, temperature=0.07)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(golden["refshim_hcl_loss"]), rtol=2e-5)
    da, dp = golden["refshim_hcl_da"], golden["refshim_hcl_dp"]
    assert np.abs(a.grad.cpu().numpy() - da).max() < 2e-4 * np.abs(da).max()
    assert np.abs(p.grad.cpu().numpy() - dp).max() < 2e-4 * np.abs(dp).max()
    s2p = train.sample_to_prototype_loss(torch.from_numpy(golden["refshim_s2p_s"]).float().cuda(),
                                         torch.from_numpy(golden["refshim_s2p_pos"]).float().cuda(),
                                         torch.from_numpy(golden["refshim_s2p_neg"]).float().cuda(), 3, k, margin=0.1)
    np.testing.assert_allclose(s2p.item(), float(golden["refshim_s2p_loss"]), rtol=2e-5)


def test_pairdist_backward_vs_autograd_oracle():
    from oracle import contrastive, head
    from patent_image_retrieval_b200 import train
    # positives close to their anchors (the cancellation-prone diagonal), but a temperature at which the
    # softmax is not saturated: with tau=0.1 the true gradient is ~1e-8 and ANY fp32 path (the reference's
    # included) only carries rounding noise of (softmax_ii - 1)
    c, n, d, tau = 1.0, 96, 128, 0.5
    mu = synth.gaussian_features(n, d, seed=2, scale=1.0)
    a0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=3, scale=1.0), c)
    p0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=4, scale=1.0), c)
    k = torch.tensor([-c], dtype=torch.float64)
    a64, p64 = a0.double().requires_grad_(True), p0.double().requires_grad_(True)
    contrastive.contrastive_loss(a64, p64, k, temperature=tau).backward()
    ag, pg = a0.cuda().requires_grad_(True), p0.cuda().requires_grad_(True)
    loss = train.in_batch_contrastive_loss(ag, pg, torch.tensor([-c]), temperature=tau)
    loss.backward()
    for got, want in ((ag.grad, a64.grad), (pg.grad, p64.grad)):
        err = (got.cpu().double() - want).abs().max() / want.abs().max()
        assert float(err) < 1e-4
    # the kernel pair itself, with an arbitrary upstream gradient (no softmax in the way)
    g = torch.randn(n, n, dtype=torch.float64)
    a64.grad = p64.grad = None
    (contrastive.dist_matrix(a64, p64, k) * g).sum().backward()
    ag.grad = pg.grad = None
    (train.pairwise_dist(ag, pg, torch.tensor([-c])) * g.float().cuda()).sum().backward()
    for got, want in ((ag.grad, a64.grad), (pg.grad, p64.grad)):
        assert float((got.cpu().double() - want).abs().max() / want.abs().max()) < 5e-6


def test_evaluate_retrieval_dropin_matches_reference(golden):
    f2p = {int(k): v for k, v in json.loads(bytes(golden["refshim_eval_f2p_json"]).decode()).items()}
    X = torch.from_numpy(golden["refshim_eval_X"])
    model = models.HyperbolicEmbeddingModel(32, 16, label_num=60, hidden_dims=[24], c=2.0)
    sd = {k: torch.from_numpy(np.array(golden["refshim_eval_" + k])).float() for k in model.state_dict().keys() if not k.endswith("isp_c")}      # golden: weights; curvature from c
    model.load_state_dict(sd)
    model = model.cuda()
    offsets = {"patents": 0, "medium_cpcs": 45, "big_cpcs": 55, "main_cpcs": 58}
    got = evaluation.evaluate_retrieval(model, X, list(range(40)), f2p, offsets, "cuda", 16)
    # the reference mixed fp32 queries with fp64 label embeddings; the CUDA path is fp32 throughout
    assert abs(got - float(golden["refshim_eval_map"])) < 2e-6
    assert evaluation.evaluate_retrieval(model, X, [], f2p, offsets, "cuda", 16) == 0.0
    assert evaluation.evaluate_retrieval(model, X, list(range(40)), f2p, {"cpcs": 3}, "cuda", 16) == -1.0
    assert evaluation.evaluate_retrieval(model, X[:, :8], list(range(40)), f2p, offsets, "cuda", 16) == -1.0


def test_image_retrieval_and_query_evaluation(golden):
    g, q = golden["ref_cos_g"], golden["ref_cos_q"]
    paths = [f"gallery/fig_{i:05d}.png" for i in range(g.shape[0])]
    ir = evaluation.ImageRetrieval(g, paths)
    res = ir.retrieve_similar_images(q[0], k=20)
    order = golden["ref_cos_order"]
    assert [p for p, _ in res] == [paths[j] for j in order[0, :20]]
    np.testing.assert_allclose([s for _, s in res], golden["ref_cos_sim"][0, order[0, :20]], atol=2e-6)
    o, it = golden["ref_nb_pos_offsets"], golden["ref_nb_pos_items"]
    gt = {f"q{i}.png": {"patent_positives": [f"fig_{j:05d}.png" for j in it[o[i]:o[i + 1]]]} for i in range(12)}
    names = [f"queries/q{i}.png" for i in range(12)] + ["queries/not_in_gt.png"]
    qq = np.concatenate([q, q[:1]], 0)
    got = evaluation.evaluate_queries(ir, qq, names, gt, k=20, ks=(5, 10, 20))
    want, _ = retrieval.notebook_metrics(order[:, :20], [it[o[i]:o[i + 1]] for i in range(12)], ks=(5, 10, 20))
    for name in ("mrr", "ap", "ndcg", "recall@5", "recall@10", "recall@20", "precision@5", "mrr@5"):
        assert abs(got[name] - want[name]) < 1e-12, name


def test_merge_topk():
    torch.manual_seed(0)
    W, Q, k = 4, 300, 10
    s = torch.rand(W, Q, k, device="cuda").sort(dim=2).values
    i = torch.randperm(W * Q * k, device="cuda").view(W, Q, k)
    s[1, :, 3] = s[0, :, 2]                                   # exact ties across shards
    i[2, 5, 7:] = -1                                          # short list
    got_s, got_i = ops.merge_topk(s, i)
    fs = s.permute(1, 0, 2).reshape(Q, -1).cpu().numpy()
    fi = i.permute(1, 0, 2).reshape(Q, -1).cpu().numpy()
    fs = np.where(fi < 0, np.inf, fs)
    order = np.lexsort((fi, fs), axis=1)[:, :k]
    np.testing.assert_array_equal(got_i.cpu().numpy(), np.take_along_axis(fi, order, 1))
    np.testing.assert_array_equal(got_s.cpu().numpy(), np.take_along_axis(fs, order, 1))
    d_s, d_i = ops.merge_topk(s, i, descending=True)
    order = np.lexsort((fi, np.where(fi < 0, np.inf, -fs)), axis=1)[:, :k]
    np.testing.assert_array_equal(d_i.cpu().numpy(), np.take_along_axis(fi, order, 1))


def test_sharded_index_single_process_equals_plain_index():
    from patent_image_retrieval_b200 import GalleryIndex
    from patent_image_retrieval_b200.dist import ShardedGalleryIndex
    u = synth.gaussian_features(50, 128, seed=1).cuda()
    v = synth.gaussian_features(3000, 128, seed=0).cuda()
    a = GalleryIndex(v).search(u, k=10)
    # two "ranks" emulated in one process: search each shard, then the merge kernel
    parts = [ShardedGalleryIndex(v[lo:hi], lo, 3000).search(u, k=10) for lo, hi in ((0, 1500), (1500, 3000))]
    ms, mi = ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, a[1]) and torch.equal(ms, a[0])


@pytest.mark.parametrize("c", [1.0, 0.5])
def test_fused_head_matches_reference_models_py(golden, c):
    """Inference path of the drop-in encoder = 2 GEMMs + 2 fused epilogue kernels; must reproduce the
    output of the reference's own models.py (refshim golden) and of the op-by-op torch path."""
    tag = str(c).replace(".", "p")
    m = models.FigureOnlyHyperbolicModel(32, 16, hidden_dims=[24], c=c, dropout_rate=0.3).eval()
    m.load_state_dict({k: torch.from_numpy(np.array(golden[f"refshim_head_{k}_c{tag}"])).float()
                       for k in m.state_dict().keys() if not k.endswith("isp_c")})
    m = m.cuda()
    x = torch.from_numpy(golden[f"refshim_head_x_c{tag}"]).cuda()
    with torch.no_grad():
        assert m.encoder._fused_ok(x)
        y_fused = m(x)
    want = torch.from_numpy(golden[f"refshim_head_y_c{tag}"])
    torch.testing.assert_close(y_fused.cpu(), want, rtol=5e-6, atol=2e-7)
    y_ops = m(x.clone().requires_grad_(True))               # autograd on -> op-by-op shim path
    torch.testing.assert_close(y_fused, y_ops.detach(), rtol=5e-6, atol=2e-7)


def test_fused_head_large_and_clipped_rows():
    from oracle import head
    torch.manual_seed(5)
    c = 2.0
    m = models.HyperbolicEmbeddingModel(512, 128, label_num=4, hidden_dims=[256], c=c).eval()
    with torch.no_grad():
        m.encoder.first_layer.weight.mul_(6.0)              # push many rows onto the project clip
    x = torch.randn(1000, 512)
    x[3] = 0.0
    sd = m.state_dict()
    k = torch.tensor([-c])
    want = head.encoder_forward(x, sd["encoder.first_layer.weight"], sd["encoder.first_layer.bias"],
                                sd["encoder.final_layer.weight"], sd["encoder.final_layer.bias"], k)
    m = m.cuda()
    with torch.no_grad():
        got = m.encode_figures(x.cuda()).cpu()
    # GEMM summation order differs between cuBLAS and the CPU oracle; rows near the clip amplify that slightly
    scale = want.norm(dim=1, keepdim=True).clamp_min(1e-20)
    assert float(((got - want).abs() / scale).max()) < 2e-5
    assert float(got.norm(dim=1).max()) <= (1 - 4e-3) / c ** 0.5 * (1 + 1e-6)


@pytest.mark.parametrize("n,m,d,c,scale", [(300, 520, 128, 1.0, 1.0), (129, 257, 256, 0.5, 1.0), (64, 64, 20, 2.0, 1.0),
                                           (1000, 700, 512, 1.0, 1.0), (260, 512, 128, 1.0, 0.15),
                                           (200, 384, 64, 0.7, 0.03), (150, 256, 128, 1.0, 0.3)])
def test_tensor_core_distance_matrix_matches_oracle(n, m, d, c, scale):
    """csrc/gramdist.cu (3-way bf16 split Gram matrix on tcgen05 + exact recompute of near pairs) against the fp64
    oracle and the exact CUDA-core kernel, including near duplicates and ragged tiles.  The small scales put the
    distances below / around d sqrt(c) = 0.41, where the epilogue swaps its fast lg2 for the accurate log1pf."""
    from oracle import head
    a = head.embed_rows(synth.gaussian_features(n, d, seed=1, scale=1.0) * scale, c)
    p = head.embed_rows(synth.gaussian_features(m, d, seed=0, scale=1.0) * scale, c)
    p[:5] = a[:5] * (1 + 1e-4)                                   # near duplicates: the cancellation case
    p[5:9] = a[5:9] * 0.9
    got, asq, psq = ops.gram_dist(a.cuda(), p.cuda(), c)
    got = got.cpu()
    d64 = retrieval.hyperbolic_dist_rows(a.double(), p.double(), c, form="arcosh")
    exact = ops.pairdist(a.cuda(), p.cuda(), c).cpu()
    far = d64 > 1e-2
    assert float(((got.double() - d64).abs() / d64)[far].max()) < 1e-5        # BASELINE north_star: 1e-5 relative
    assert float(((got - exact).abs() / exact)[far].max()) < 1e-5
    assert float(((got.double() - d64).abs() / d64)[~far].max()) < 1e-3       # near pairs: exact path, as pairdist
    torch.testing.assert_close(asq.cpu(), a.pow(2).sum(1), rtol=1e-5, atol=0)
    torch.testing.assert_close(psq.cpu(), p.pow(2).sum(1), rtol=1e-5, atol=0)


def test_infonce_tensor_core_path_matches_cuda_core_path():
    from oracle import head
    from patent_image_retrieval_b200 import train
    c, n, d, tau = 1.0, 384, 128, 0.5
    mu = synth.gaussian_features(n, d, seed=2, scale=1.0)
    a0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=3, scale=1.0), c).cuda()
    p0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=4, scale=1.0), c).cuda()
    inv_tau = 1.0 / tau
    d_tc, r_tc, c_tc = ops.pairdist_ce_fwd(a0, p0, c, inv_tau, True, tensor_cores=True)
    d_cc, r_cc, c_cc = ops.pairdist_ce_fwd(a0, p0, c, inv_tau, True, tensor_cores=False)
    torch.testing.assert_close(d_tc, d_cc, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(r_tc, r_cc, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(c_tc, c_cc, rtol=1e-5, atol=1e-5)
    k = torch.tensor([-c])
    for sym in (False, True):
        ag, pg = a0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
        loss = train.InBatchInfoNCE.apply(ag, pg, c, tau, sym)          # tensor-core forward (n*m >= 2^16)
        loss.backward()
        a64, p64 = a0.cpu().double().requires_grad_(True), p0.cpu().double().requires_grad_(True)
        from oracle import contrastive
        sim = -contrastive.dist_matrix(a64, p64, k.double()) / tau
        lab = torch.arange(n)
        ref = torch.nn.functional.cross_entropy(sim, lab)
        if sym:
            ref = (ref + torch.nn.functional.cross_entropy(sim.t(), lab)) / 2
        ref.backward()
        assert float(loss) == pytest.approx(float(ref), rel=2e-5)
        for got, want in ((ag.grad, a64.grad), (pg.grad, p64.grad)):
            assert float((got.cpu().double() - want).abs().max() / want.abs().max()) < 1e-4


def _dist64(a, p, c):
    """Differentiable fp64 closed form arccosh(1 + 2c|a-p|^2 / ((1-c|a|^2)(1-c|p|^2))) / sqrt(c) (SURVEY 8c:
    analytically pmath.dist; the broadcast geoopt form would need [n,m,D] fp64 autograd temporaries)."""
    a2, p2 = a.pow(2).sum(1), p.pow(2).sum(1)
    diff = (a2[:, None] + p2[None, :] - 2 * a @ p.t()).clamp_min(0)
    x = 1 + 2 * c * diff / ((1 - c * a2)[:, None] * (1 - c * p2)[None, :])
    return torch.acosh(x.clamp_min(1 + 1e-15)) / c ** 0.5


def test_infonce_split_bf16_backward_matches_fp64_autograd():
    """n*m >= ops.SPLIT_MIN_PAIRS: the backward emits W as three bf16 planes and forms W P / W^T A as bf16
    tensor-core GEMMs with fp32 accumulation (six cross products); gradients against fp64 autograd through the
    oracle's distance matrix, and against the fp32-W path."""
    from oracle import head
    from patent_image_retrieval_b200 import train
    c, n, d, tau = 0.8, 1100, 128, 0.2
    assert n * n >= ops.SPLIT_MIN_PAIRS
    mu = synth.gaussian_features(n, d, seed=2, scale=1.0)
    a0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=3, scale=1.0), c).cuda()
    p0 = head.embed_rows(mu + 0.3 * synth.gaussian_features(n, d, seed=4, scale=1.0), c).cuda()
    for sym in (False, True):
        ag, pg = a0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
        loss = train.InBatchInfoNCE.apply(ag, pg, c, tau, sym)
        loss.backward()
        a64, p64 = a0.cpu().double().requires_grad_(True), p0.cpu().double().requires_grad_(True)
        sim = -_dist64(a64, p64, c) / tau
        lab = torch.arange(n)
        ref = torch.nn.functional.cross_entropy(sim, lab)
        if sym:
            ref = (ref + torch.nn.functional.cross_entropy(sim.t(), lab)) / 2
        ref.backward()
        assert float(loss) == pytest.approx(float(ref), rel=2e-5)
        for got, want in ((ag.grad, a64.grad), (pg.grad, p64.grad)):
            assert float((got.cpu().double() - want).abs().max() / want.abs().max()) < 1e-4
    # the split products themselves: fp32-GEMM accuracy or better
    dm, rl, cl = ops.pairdist_ce_fwd(a0, p0, c, 1 / tau, True)
    asq, psq = ops.row_sqnorm(a0), ops.row_sqnorm(p0)
    w, rs, cs = ops.pairdist_ce_bwd(dm, asq, psq, c, rl, cl, 1 / tau, 0.5, 0.5)
    w3, rs3, cs3 = ops.pairdist_ce_bwd(dm, asq, psq, c, rl, cl, 1 / tau, 0.5, 0.5, split=True)
    assert w3.data.dtype == torch.bfloat16 and tuple(w3.data.shape) == (3, n, n)
    torch.testing.assert_close(w3.float(), w, rtol=3e-7, atol=0)
    assert torch.equal(rs, rs3) and torch.equal(cs, cs3)
    wp, wta = ops.split_products(w3, a0, p0)
    ref_wp, ref_wta = w.double() @ p0.double(), w.double().t() @ a0.double()
    # error against the size of the summed terms (the sums cancel: W has a negative diagonal).  The six bf16 cross
    # products are exact to 2^-24, but the tensor core's fp32 accumulator TRUNCATES: up to one ulp of the running sum
    # per 16-deep MMA step, all in one direction (measured 4.4e-6 here = 69 steps x 2^-24; an FFMA SGEMM rounds to
    # nearest, ~sqrt(K) ulps) -- well inside the 1e-4 gradient tolerance above, documented in DESIGN 4.6
    bound = 1.5 * (n / 16 + 8) * 2.0 ** -24
    scale_wp, scale_wta = w.abs().double() @ p0.abs().double(), w.abs().double().t() @ a0.abs().double()
    assert float(((wp - ref_wp).abs() / scale_wp).max()) < bound
    assert float(((wta - ref_wta).abs() / scale_wta).max()) < bound
    assert float((wp - ref_wp).abs().max() / ref_wp.abs().max()) < 2e-5
    assert float((wta - ref_wta).abs().max() / ref_wta.abs().max()) < 2e-5


def test_generic_pairdist_backward_split_and_ragged():
    """PairwiseDistance.backward on a ragged n x m (row blocks of 16, column chunks) in both W formats."""
    from oracle import head
    from patent_image_retrieval_b200 import train
    c = 1.3
    for n, m, d in ((37, 300, 32), (1030, 1111, 64)):
        a0 = head.embed_rows(synth.gaussian_features(n, d, seed=7, scale=1.0), c).cuda()
        p0 = head.embed_rows(synth.gaussian_features(m, d, seed=8, scale=1.0), c).cuda()
        g = torch.randn(n, m, generator=torch.Generator().manual_seed(3)).cuda()
        ag, pg = a0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
        (train.PairwiseDistance.apply(ag, pg, c) * g).sum().backward()
        a64, p64 = a0.cpu().double().requires_grad_(True), p0.cpu().double().requires_grad_(True)
        (_dist64(a64, p64, c) * g.cpu().double()).sum().backward()
        for got, want in ((ag.grad, a64.grad), (pg.grad, p64.grad)):
            assert float((got.cpu().double() - want).abs().max() / want.abs().max()) < 1e-4


@pytest.mark.parametrize("d_in,hid,d_out,c,b", [(512, 256, 128, 1.0, 1000), (2048, 256, 256, 0.5, 300),
                                                (768, 128, 64, 2.0, 129), (64, 32, 16, 1.0, 5)])
def test_mobius_gemm_head_matches_oracle(d_in, hid, d_out, c, b):
    """The head as two hypret_mobius_gemm kernels (tcgen05 GEMM + Moebius epilogue on the accumulator, the hidden
    activations handed over as the second GEMM's fp16 split operand) against the oracle's restatement of
    src/models.py:291-318, 481-505, and against the round-1 path (library GEMM + epilogue kernel) on the same weights."""
    from oracle import head
    from patent_image_retrieval_b200 import models, ops
    torch.manual_seed(9)
    m = models.FigureOnlyHyperbolicModel(d_in, d_out, hidden_dims=[hid], c=c).eval()
    with torch.no_grad():
        m.encoder.first_layer.weight.mul_(3.0)                # some rows reach the project clip
    x = torch.randn(b, d_in) * 0.7
    x[0] = 0.0
    sd = m.state_dict()
    want = head.encoder_forward(x, sd["encoder.first_layer.weight"], sd["encoder.first_layer.bias"],
                                sd["encoder.final_layer.weight"], sd["encoder.final_layer.bias"], torch.tensor([-c]))
    assert ops.mobius_gemm_ok(d_in, hid) and ops.mobius_gemm_ok(hid, d_out)
    mg = m.cuda()
    with torch.no_grad():
        got = mg.encode_figures(x.cuda()).cpu()
    scale = want.norm(dim=1, keepdim=True).clamp_min(1e-20)
    assert float(((got - want).abs() / scale).max()) < 2e-5
    assert float(got.norm(dim=1).max()) <= (1 - 4e-3) / c ** 0.5 * (1 + 1e-6)
    assert torch.equal(got[0], got[0]) and bool(torch.isfinite(got).all())


@pytest.mark.parametrize("d_in,hid,d_out,c,b,wscale", [(512, 256, 128, 1.0, 128, 1.0), (768, 128, 64, 0.5, 77, 1.0),
                                                       (64, 32, 16, 2.0, 200, 4.0),
                                                       (256, 128, 64, 1.0, 6001, 1.0)])   # grid-stride rows, split-K dW
def test_head_training_step_matches_autograd(d_in, hid, d_out, c, b, wscale, monkeypatch):
    """train_hyp's head on the kernel path (ops.MobiusLinearFn: fused GEMM kernel forward, closed-form epilogue kernel +
    dense products backward) against autograd through the op-by-op path of the same module on the CPU
    (src/models.py:291-318, 481-505): output, dL/dW1, dL/db1, dL/dW2, dL/db2 and dL/dx.  Truth is the fp64 run (with
    geoopt's fp32 project margin); the bar is 5e-5 of the tensor's scale or twice the error of the reference's own fp32
    autograd, whichever is larger -- the first layer saturates (tanh(|mx|) ~ 1), where dL/db1 is ~1e-6 and fp32 noise
    dominates it in BOTH implementations."""
    import copy
    from patent_image_retrieval_b200 import models
    from patent_image_retrieval_b200.geoopt_shim import pmath
    real_project = pmath.project
    monkeypatch.setattr(pmath, "project", lambda x, *, k, dim=-1, eps=-1.0: real_project(x, k=k, dim=dim, eps=4e-3))
    torch.manual_seed(21)
    m = models.DeeperHyperbolicEncoder(d_in, [hid], d_out, c=c, dropout_rate=0.0).train()
    with torch.no_grad():
        m.first_layer.weight.mul_(wscale)                    # wscale 4: rows reach the project clip
    x = torch.randn(b, d_in) * 0.5
    r = torch.randn(b, d_out)
    names = [n for n, _ in m.named_parameters() if not n.endswith("isp_c")]
    res = {}
    for tag in ("f64", "f32", "gpu"):
        mod = copy.deepcopy(m)
        if tag == "f64":
            mod, xx, rr = mod.double(), x.double().requires_grad_(True), r.double()
        elif tag == "f32":
            xx, rr = x.clone().requires_grad_(True), r
        else:
            mod, xx, rr = mod.cuda(), x.cuda().requires_grad_(True), r.cuda()
            assert mod._kernel_train_ok(xx)
        y = mod(xx)
        (y * rr).sum().backward()
        res[tag] = [y.detach().cpu().double(), xx.grad.cpu().double()] + \
                   [mod.get_parameter(n).grad.cpu().double() for n in names]
    for i, what in enumerate(["y", "dx"] + names):
        truth = res["f64"][i]
        scale = float(truth.abs().max())
        e32 = float((res["f32"][i] - truth).abs().max())
        egpu = float((res["gpu"][i] - truth).abs().max())
        assert egpu <= max(5e-5 * scale, 2.0 * e32), (what, egpu, e32, scale)


def test_head_training_step_dropout_same_masks_as_eager():
    """With dropout on, the kernel path draws its masks with the same two F.dropout calls as the op-by-op path, so under
    one seed both see identical masks: outputs and gradients agree to fp32 rounding (the biases: to the noise floor of
    their saturated gradients, see the test above)."""
    import copy
    from patent_image_retrieval_b200 import models
    torch.manual_seed(3)
    m = models.DeeperHyperbolicEncoder(512, [256], 128, c=1.0, dropout_rate=0.3).cuda().train()
    eager = copy.deepcopy(m)
    eager._kernel_train_ok = lambda x: False
    x = torch.randn(128, 512, device="cuda") * 0.5
    r = torch.randn(128, 128, device="cuda")
    names = [n for n, _ in m.named_parameters() if not n.endswith("isp_c")]
    outs = []
    for mod in (m, eager):
        torch.manual_seed(77)
        torch.cuda.manual_seed(77)
        y = mod(x)
        (y * r).sum().backward()
        outs.append((y.detach(), {n: mod.get_parameter(n).grad.clone() for n in names}))
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 2e-5
    for n in names:
        a, w = outs[0][1][n], outs[1][1][n]
        floor = 1e-4 if n.endswith("bias") else 0.0       # saturated layers: bias gradients ~1e-4 and below
        assert float((a - w).abs().max()) <= max(floor, 5e-4 * float(w.abs().max())), (n, float((a - w).abs().max()), float(w.abs().max()))
