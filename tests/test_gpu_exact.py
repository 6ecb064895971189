"""GPU: the exact-top-k guarantee of ``GalleryIndex.search``.

The fp16 tensor-core pass is a filter; ``hypret_rerank_cert`` must PROVE per query that nothing outside the candidate
set can precede the k-th result, and ``hypret_exact_topk`` must recompute the queries it cannot prove.  Checked here:
the error bound really bounds the filter's error (all pairs, random / clustered / adversarially aligned rows), the
full-scan kernel equals the oracle's loop + top-k (reference src/train.py:3259, src/auxiliary.py:374), certified
queries never differ from the oracle even with the fallback switched off, and on near-duplicate galleries -- the
reference's patent figures -- the returned LISTS equal the fp64 oracle's for every query, at BASELINE config 1's full
size too."""
import numpy as np
import pytest
import torch

from oracle import head, retrieval
from patent_image_retrieval_b200 import GalleryIndex, ops, synth

pytestmark = pytest.mark.gpu


def _expected_lists(d64: torch.Tensor, k: int):
    """The documented order of the GPU path: ascending fp32-rounded exact distance, equal fp32 values in index order
    (the reference's torch.topk / argsort leave exact ties unspecified)."""
    d32 = d64.float().numpy()
    idx = np.lexsort((np.broadcast_to(np.arange(d32.shape[1]), d32.shape), d32), axis=1)[:, :k]
    return torch.from_numpy(idx.copy())


TIE_RTOL = 3e-7     # two fp64 distances closer than this may round to one fp32 value: "documented exact-distance ties"


def _assert_lists(idx: torch.Tensor, d64: torch.Tensor, k: int, what: str = "") -> int:
    """Index LISTS must equal the fp64 oracle's.  Rows that differ are accepted only if the difference is explained
    by a tie: the fp32 distances the GPU orders by (explicit fp32 differences, fp64 accumulation) can merge two fp64
    values that are within TIE_RTOL of each other, which then rank in index order.  Returns the number of such rows."""
    want = _expected_lists(d64, k)
    diff = torch.nonzero((idx != want).any(dim=1)).flatten().tolist()
    kth = torch.gather(d64, 1, want[:, k - 1:k]).squeeze(1)
    for r in diff:
        got_d = d64[r, idx[r]]
        tol = TIE_RTOL * float(kth[r].abs()) + 1e-300
        assert len(set(idx[r].tolist())) == k, f"{what} row {r}: duplicate ids"
        assert float(got_d.max()) <= float(kth[r]) + tol, f"{what} row {r}: a non-top-k row was returned"
        assert bool((got_d[1:] - got_d[:-1] >= -tol).all()), f"{what} row {r}: order violates the distances"
        better = torch.nonzero(d64[r] < float(kth[r]) - tol).flatten().tolist()
        assert set(better) <= set(idx[r].tolist()), f"{what} row {r}: a strictly better row is missing"
    return len(diff)


def _oracle_d64(qry, gal, c, metric):
    """fp64 truth from the SAME fp32 rows the GPU index holds (on-ball points / raw features)."""
    if metric == "hyperbolic":
        return retrieval.hyperbolic_dist_rows(qry.double(), gal.double(), c, form="arcosh")
    return -torch.from_numpy(retrieval.cosine_similarity(qry.double().numpy(), gal.double().numpy()))


def _near_duplicates(n_gallery, n_query, d, noise, c, metric, per_class=8):
    """Clustered rows whose within-class noise -> 0: many rows inside the fp16 error band of the k-th best.  For the
    ball they are returned as ON-BALL points (``space='ball'``): between near duplicates the distance is so much
    smaller than the norms that a one-ulp difference between two implementations of expmap0 (CPU tanh vs GPU tanhf)
    already reorders them -- the guarantee is about scoring and ranking given the points, so both sides get the same
    fp32 points (the projection's own parity is tests/test_gpu_project.py)."""
    gal, qry, g_cls, q_cls = synth.clustered_features(n_gallery, n_query, d, noise=noise, per_class=per_class)
    if metric == "hyperbolic":
        gal, qry = head.embed_rows(gal, c), head.embed_rows(qry, c)
    return gal, qry


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
@pytest.mark.parametrize("kind", ["gaussian", "near_dup", "aligned"])
def test_error_bound_covers_every_pair(metric, kind):
    d, c, Q, N = 512, 0.8, 200, 3000
    if kind == "gaussian":
        qry, gal = synth.gaussian_features(Q, d, seed=1), synth.gaussian_features(N, d, seed=0)
    elif kind == "near_dup":
        gal, qry, _, _ = synth.clustered_features(N, Q, d, noise=1e-3)
    else:       # queries parallel to gallery rows, all entries of one sign and rounding-unfriendly: Cauchy-Schwarz is tight
        gal = synth.gaussian_features(N, d, seed=0).abs() * (1.0 + 2.0 ** -12) + 1e-3
        qry = gal[:Q] * 0.8 + 1e-4
    index = GalleryIndex(gal.cuda(), c=c, metric=metric)
    mode = index._query_mode()
    if metric == "hyperbolic":
        q32, q_op, _, q_err = ops.project_rows(qry.cuda(), c, mode=mode, side="query", want_err=True)
    else:
        _, q_op, _, q_err = ops.project_rows(qry.cuda(), 1.0, mode=mode, side="query", want_point=False, want_err=True)
        q32 = qry.cuda()
    _, _, approx = ops.score_topk(q_op, index.operand, d, 16, debug=True)
    approx = approx.double().cpu()
    # exact surrogate in fp64 from the fp32 rows the rerank reads
    qx, gy = q32.double().cpu(), index.rows32.double().cpu()
    if metric == "hyperbolic":
        s = torch.cdist(qx, gy, compute_mode="donot_use_mm_for_euclid_dist").pow(2)
        exact = c * s / (1 - c * gy.pow(2).sum(1))[None, :]          # unit-ball coordinates, like the filter
        qn = qx.norm(dim=1) * c ** 0.5
    else:
        exact = _oracle_d64(qry, gal, c, "cosine")
        qn = torch.ones(Q, dtype=torch.float64)
    st = index.stats.double().cpu()
    slack = (ops.operand_kpad(d) / 16 + 8) * 2.0 ** -22
    E = q_err.double().cpu() * st[0] + qn * st[1] + slack * (qn * st[0] + qn * qn * st[2] + st[3])
    worst = ((approx - exact).abs() / E[:, None]).max()
    assert float(worst) <= 1.0, f"filter error exceeds the certificate's bound by {float(worst):.3f}x"
    assert float(worst) > 1e-3                      # ... and the bound is not vacuous


@pytest.mark.parametrize("metric,d,n,k", [("hyperbolic", 512, 5000, 10), ("hyperbolic", 100, 2049, 32),
                                          ("cosine", 768, 3000, 20), ("hyperbolic", 2048, 700, 1),
                                          ("hyperbolic", 512, 7, 10)])
def test_exact_scan_kernel_equals_oracle(metric, d, n, k):
    c, Q = 0.7, 37
    gal, qry = _near_duplicates(n, Q, d, 0.05, c, metric, per_class=4)
    d64 = _oracle_d64(qry, gal, c, metric)
    index = GalleryIndex(gal.cuda(), c=c, metric=metric, space="ball" if metric == "hyperbolic" else "euclidean")
    assert torch.equal(index.rows32.cpu(), gal)
    q32 = qry.cuda()
    score, idx = ops.exact_topk(q32, index.rows32, index.rows_sq64, c, metric, k, idx_offset=1000)
    idx, score = idx.cpu(), score.cpu()
    kk = min(k, n)
    _assert_lists(idx[:, :kk] - 1000, d64, kk, "exact scan:")
    want = idx[:, :kk] - 1000
    assert bool((idx[:, kk:] == -1).all())
    ref = torch.gather(d64, 1, want)
    got = score[:, :kk].double() if metric == "hyperbolic" else -score[:, :kk].double()
    assert float(((got - ref).abs() / ref.abs().clamp_min(1e-30)).max()) < 2e-6


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
@pytest.mark.parametrize("noise", [0.3, 1e-2, 1e-3, 0.0])
def test_search_lists_equal_oracle_on_near_duplicate_gallery(metric, noise):
    d, c, k, Q, N = 512, 1.0, 10, 256, 12_000
    gal, qry = _near_duplicates(N, Q, d, noise, c, metric)
    index = GalleryIndex(gal.cuda(), c=c, metric=metric, space="ball" if metric == "hyperbolic" else "euclidean")
    score, idx = index.search(qry.cuda(), k=k)
    n_fallback = int(index.certificate.count.item())
    certified = index.certificate.certified[:Q].bool().cpu()
    d64 = _oracle_d64(qry, gal, c, metric)
    want = _expected_lists(d64, k)
    n_tie_rows = _assert_lists(idx.cpu(), d64, k, f"noise {noise}, {n_fallback} rescanned:")
    assert n_tie_rows <= Q // 50
    assert int((~certified).sum()) == n_fallback
    if noise == 0.0:
        assert n_fallback > 0                       # exact duplicates cannot be certified: the scan must have run
    # soundness of the certificate alone: with the fallback switched off, certified queries are already right
    q32, cs, ci, cnt, q_err = index.score_candidates(qry.cuda(), k=k, want_err=True)
    bufs = ops.CertBuffers(Q, index.device)
    _, idx_nf = ops.rerank_cert(q32, index.rows32, cs, ci, c, metric, k, q_err, index.stats, index.rows_sq64, bufs,
                                list_count=cnt, fallback=False)
    ok = bufs.certified[:Q].bool().cpu()
    _assert_lists(idx_nf.cpu()[ok], d64[ok], k, "certified, no fallback:")


def test_gaussian_queries_are_certified_without_rescan():
    """i.i.d. features (the bench workloads): the gap between the k-th and the k'-th best is far above the rounding
    bound, so the guarantee costs one empty launch."""
    d, c, k = 512, 1.0, 10
    index = GalleryIndex(synth.gaussian_features(30_000, d, seed=0).cuda(), c=c)
    index.search(synth.gaussian_features(1024, d, seed=1).cuda(), k=k)
    assert float(index.certificate.certified[:1024].float().mean()) > 0.99


def test_full_size_config1_lists_distances_and_metrics():
    """BASELINE config 1 at FULL size (1k queries x 10k gallery x 2048, top-10, c=1) against the reference's fp32 path
    and the fp64 truth: lists, distances, recall@k / mAP (clustered variant so that recall is non-trivial)."""
    Q, N, d, c, k = 1000, 10_000, 2048, 1.0, 10
    for clustered in (False, True):
        if clustered:
            gal, qry, g_cls, q_cls = synth.clustered_features(N, Q, d)
        else:
            gal, qry = synth.gaussian_features(N, d, seed=0), synth.gaussian_features(Q, d, seed=1)
        index = GalleryIndex(gal.cuda(), c=c)
        dist, idx = index.search(qry.cuda(), k=k)
        dist, idx = dist.cpu(), idx.cpu()
        q32, g32 = head.embed_rows(qry, c), head.embed_rows(gal, c)
        d64 = retrieval.hyperbolic_dist_rows(q32.double(), g32.double(), c, form="arcosh")
        want = _expected_lists(d64, k)
        assert _assert_lists(idx, d64, k, f"C1 clustered={clustered}:") <= 2       # of 1000 rows, tie-explained
        ref64 = torch.gather(d64, 1, idx)
        assert float(((dist.double() - ref64).abs() / ref64).max()) < 2e-6
        # the reference's own fp32 arithmetic (per-query pmath.dist, src/train.py:3259) on a 128-query sample
        d32 = retrieval.hyperbolic_dist_rows(q32[:128], g32, c, form="geoopt")
        ref32 = torch.gather(d32, 1, idx[:128])
        assert float(((dist[:128] - ref32).abs() / ref32).max()) < 1e-5         # north_star tolerance
        _, want32 = retrieval.topk_smallest(d32, k)
        same_sets = sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(idx[:128], want32))
        assert same_sets >= 127                      # fp32 reference noise can swap the 10th / 11th of a near tie
        if clustered:
            off, items = synth.positives_csr(q_cls, g_cls)
            pos = [items[off[i]:off[i + 1]].tolist() for i in range(Q)]
            m_gpu, _ = retrieval.notebook_metrics(idx.tolist(), pos, ks=(5, 10))
            m_ref, _ = retrieval.notebook_metrics(want.tolist(), pos, ks=(5, 10))
            assert m_ref["recall@10"] > 0.5
            for name in ("recall@5", "recall@10", "ap", "mrr"):
                assert abs(m_gpu[name] - m_ref[name]) < 1e-12, name


def test_exact_search_overhead_is_device_side_only():
    """No host synchronisation on the guaranteed path: the search can be captured into a CUDA graph."""
    d, c, k, Q = 256, 1.0, 10, 512
    gal, qry = _near_duplicates(4000, Q, d, 1e-3, c, "hyperbolic")
    index = GalleryIndex(gal.cuda(), c=c, space="ball")
    qd = qry.cuda()
    want = index.search(qd, k=k)[1].clone()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        index.search(qd, k=k)                       # warm the allocator for the capture
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = index.search(qd, k=k)
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1], want)


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_search_beyond_128_pages_through_the_exact_ranking(metric):
    """k > 128 (the reference ranks the whole gallery, notebooks/retrieval.ipynb:383,202): exact lists by paging."""
    d, c, k, Q, N = 128, 1.0, 300, 9, 2500
    gal, qry = _near_duplicates(N, Q, d, 0.05, c, metric)
    index = GalleryIndex(gal.cuda(), c=c, metric=metric, space="ball" if metric == "hyperbolic" else "euclidean")
    score, idx = index.search(qry.cuda(), k=k)
    assert tuple(idx.shape) == (Q, k)
    d64 = _oracle_d64(qry, gal, c, metric)
    _assert_lists(idx.cpu(), d64, k, "paged:")
    got = score.cpu().double() if metric == "hyperbolic" else -score.cpu().double()
    ref = torch.gather(d64, 1, idx.cpu())
    assert float(((got - ref).abs() / ref.abs().clamp_min(1e-30)).max()) < 2e-6
    # more rows requested than the gallery has: padded with -1
    small = GalleryIndex(gal[:150].cuda(), c=c, metric=metric, space="ball" if metric == "hyperbolic" else "euclidean")
    _, i2 = small.search(qry.cuda(), k=200)
    assert tuple(i2.shape) == (Q, 150) and bool((i2 >= 0).all())


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_lists_of_16_sharing_the_bound_of_the_24th_best(metric):
    """score_topk(kbound=24) with 16-slot register lists: the union of a query's lists holds its 24 best filter scores
    unless one list ran out of slots, and rerank_cert(ksel=24) measures its margin against the smaller of the 24th best
    filter score and the worst entry of any full list -- checked against the full [Q,N] filter-score matrix."""
    Q, N, d, c, k, kp, kb = 300, 40000, 256, 1.0, 10, 16, 24
    index = GalleryIndex(synth.gaussian_features(N, d, seed=0).cuda(), c=c, metric=metric)
    qry = synth.gaussian_features(Q, d, seed=1).cuda()
    if metric == "hyperbolic":
        q32, q_op, _, q_err = ops.project_rows(qry, c, mode=index._query_mode(), side="query", want_err=True)
    else:
        _, q_op, _, q_err = ops.project_rows(qry, 1.0, mode="cosine", side="query", want_point=False, want_err=True)
        q32 = qry
    cnt = torch.zeros(Q, dtype=torch.int32, device="cuda")
    cs, ci, dbg = ops.score_topk(q_op, index.operand, d, kp, debug=True, share_thresholds=True, list_count=cnt, kbound=kb)
    torch.cuda.synchronize()
    L = cs.shape[1]
    valid = (torch.arange(L, device="cuda")[None, :, None] < cnt[:, None, None]) & (ci >= 0)
    full = valid.all(dim=2)                                                     # [Q,L]
    worst_full = torch.where(full, torch.where(valid, cs, torch.full_like(cs, float("-inf"))).amax(dim=2),
                             torch.full((Q, L), float("inf"), device="cuda")).amin(dim=1)
    top_s, top_i = torch.sort(dbg, dim=1, stable=True)
    top_s, top_i = top_s[:, :kb], top_i[:, :kb]
    held = torch.zeros(Q, N, dtype=torch.bool, device="cuda")
    rows = torch.arange(Q, device="cuda")[:, None, None].expand_as(ci)
    held[rows[valid], ci[valid].long()] = True
    # every one of the 24 best filter scores below the truncation limit is in the union
    need = top_s < worst_full[:, None]
    assert bool((held.gather(1, top_i) | ~need).all())
    assert float((worst_full >= top_s[:, -1]).float().mean()) > 0.95            # truncation is the exception
    bufs = ops.CertBuffers(Q, "cuda")
    out_s, out_i, margin = ops.rerank_cert(q32, index.rows32, cs, ci, c, metric, k, q_err, index.stats, index.rows_sq64,
                                           bufs, list_count=cnt, fallback=False, want_margin=True, ksel=kb)
    want_s, want_i = ops.exact_topk(q32, index.rows32, index.rows_sq64, c, metric, k)
    ok = bufs.certified[:Q].bool()
    assert float(ok.float().mean()) > 0.99
    assert torch.equal(out_i[ok], want_i[ok]) and torch.equal(out_s[ok], want_s[ok])
    # the margin's limit: min(24th best filter score, worst entry of a full list), minus the exact surrogate of the
    # k-th result
    x, y = q32.double(), index.rows32.double()[want_i[:, k - 1]]
    if metric == "hyperbolic":
        sur = c * (x - y).pow(2).sum(1) / (1 - c * y.pow(2).sum(1))
    else:
        sur = -(x * y).sum(1) / (x.norm(dim=1) * y.norm(dim=1))
    limit = torch.minimum(top_s[:, -1], worst_full).double()
    assert float(((limit - sur) - margin.double()).abs().max()) < 1e-5


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_exact_scan_warm_start_bound(metric):
    """hypret_exact_topk with init_bound (the k-th score of any list of real rows): CTAs whose chunk holds no contender
    skip the merge, and the result is still the exact top-k -- for a loose bound, the TIGHT bound (the true k-th score
    itself: ties at the bound must stay contenders), no bound (+-inf) and a mixture; unflagged rows stay untouched."""
    Q, N, d, c, k = 37, 30000, 256, 1.0, 10
    gal = synth.gaussian_features(N, d, seed=0).cuda()
    qry = synth.gaussian_features(Q, d, seed=1).cuda()
    if metric == "hyperbolic":
        gal = ops.project_rows(gal, c, want_operand=False)[0]
        qry = ops.project_rows(qry, c, want_operand=False)[0]
    sq = ops.row_sqnorm64(gal)
    want_s, want_i = ops.exact_topk(qry, gal, sq, c, metric, k)
    deep_s, _ = ops.exact_topk(qry, gal, sq, c, metric, 32)
    none = float("inf") if metric == "hyperbolic" else float("-inf")
    flags = torch.ones(Q, dtype=torch.int32, device="cuda")
    flags[5] = 0
    for bound in (deep_s[:, 31], want_s[:, k - 1], torch.full((Q,), none, device="cuda"),
                  torch.where(torch.arange(Q, device="cuda") % 2 == 0, want_s[:, k - 1], deep_s[:, 20])):
        got_s, got_i = ops.exact_topk_flagged(qry, gal, sq, flags, c, metric, k, init_bound=bound.contiguous())
        keep = flags.bool()
        assert torch.equal(got_i[keep], want_i[keep]) and torch.equal(got_s[keep], want_s[keep])
        assert bool((got_i[5] == -1).all())


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_wide_topk_is_exact_too(metric):
    """26 < k <= 128: search() compares the wide rerank's margin with the rounding bound and pages the queries that do
    not clear it through the exact ranking; on a near-duplicate gallery most queries take that path, and every list
    equals the full exact scan's (ops.exact_topk_any)."""
    Q, d, c, k = 48, 128, 1.0, 100
    for N, noise, per_class in ((6000, 0.3, 50), (30000, 1e-4, 600)):
        gal, qry = _near_duplicates(N, Q, d, noise, c, metric, per_class=per_class)
        index = GalleryIndex(gal.cuda(), c=c, metric=metric, space="ball" if metric == "hyperbolic" else "euclidean")
        got_s, got_i = index.search(qry.cuda(), k=k)
        want_s, want_i = ops.exact_topk_any(qry.cuda().contiguous(), index.rows32, index.rows_sq64, c, metric, k)
        assert torch.equal(got_i, want_i), (metric, noise, int((got_i != want_i).any(dim=1).sum()))
        assert torch.equal(got_s, want_s)
        if noise < 1e-2:      # 600 rows per class inside the rounding band: more than the 256 rescored survivors
            assert int(index.uncertified_wide.sum()) > 0


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_adaptive_list_width_on_tight_classes(metric):
    """A gallery of tight classes defeats the narrow certificate (classes of ~30 near-equidistant rows against k_b = 24
    candidates); the index notices from the mirrored count (no synchronisation of its own) and serves the next searches
    with the wide lists, which certify them.  Same exact lists either way; easy data stays on the narrow path."""
    Q, N, d, c, k = 256, 20000, 128, 1.0, 10
    gal, qry = _near_duplicates(N, Q, d, 0.03, c, metric, per_class=30)
    space = "ball" if metric == "hyperbolic" else "euclidean"
    index = GalleryIndex(gal.cuda(), c=c, metric=metric, space=space)
    want_s, want_i = ops.exact_topk(qry.cuda().contiguous(), index.rows32, index.rows_sq64, c, metric, k)
    got_s, got_i = index.search(qry.cuda(), k=k)
    assert index.last_mode == "narrow" and torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
    narrow_scans = int(index.certificate.count[0])                  # synchronises: the mirrored count has arrived
    assert narrow_scans > 0.02 * Q
    for _ in range(3):
        got_s, got_i = index.search(qry.cuda(), k=k)
        assert index.last_mode == "wide" and index.fallback_rate > 0.02
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
        assert int(index.uncertified_wide.sum()) < narrow_scans
    index.adaptive = False
    got_s, got_i = index.search(qry.cuda(), k=k)
    assert index.last_mode == "narrow" and torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
    easy = GalleryIndex(synth.gaussian_features(N, d, seed=0).cuda(), c=c, metric=metric)
    for _ in range(3):
        easy.search(synth.gaussian_features(Q, d, seed=1).cuda(), k=k)
        torch.cuda.synchronize()
        assert easy.last_mode == "narrow"
