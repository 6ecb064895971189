"""GPU parity of the multi-GPU pruning primitives (include/hypret.h: hypret_cand_select,
hypret_kth_smallest, hypret_rerank_pruned) against torch restatements, and of the pruned
shard protocol emulated on one GPU against the single-index search."""
import pytest
import torch

from patent_image_retrieval_b200 import GalleryIndex, ops, synth

pytestmark = pytest.mark.gpu


def _random_lists(Q, L, kp, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    s = torch.rand(Q, L, kp, generator=g)
    if ties:
        s = (s * 8).floor() / 8                       # many equal surrogates: the index tie-break decides
    i = torch.stack([torch.randperm(100000, generator=g)[:L * kp] for _ in range(Q)]).view(Q, L, kp).int()
    empty = torch.rand(Q, L, kp, generator=g) < 0.4
    i[empty] = -1
    i[0] = -1                                         # a query without any candidate
    i[1, :, :] = -1
    i[1, 0, :3] = torch.tensor([7, 5, 6], dtype=torch.int32)     # fewer than k' candidates
    return s.cuda(), i.cuda()


@pytest.mark.parametrize("L,kp,ties", [(13, 16, False), (40, 16, True), (1, 16, False), (5, 32, True), (200, 8, False)])
def test_cand_select_matches_sorted_union(L, kp, ties):
    Q = 67
    s, i = _random_lists(Q, L, kp, seed=L + kp, ties=ties)
    ss, si = ops.cand_select(s, i)
    flat_s = s.view(Q, -1).clone()
    flat_i = i.view(Q, -1).long()
    flat_s[flat_i < 0] = float("inf")
    big_i = torch.where(flat_i < 0, torch.full_like(flat_i, 1 << 40), flat_i)
    # lexicographic (score, index) order
    order = torch.argsort(big_i, dim=1, stable=True)
    order = torch.gather(order, 1, torch.argsort(torch.gather(flat_s, 1, order), dim=1, stable=True))[:, :kp]
    want_s = torch.gather(flat_s, 1, order)
    want_i = torch.gather(flat_i, 1, order)
    want_i[want_s == float("inf")] = -1
    assert torch.equal(si.long(), want_i)
    assert torch.equal(ss, want_s)


@pytest.mark.parametrize("W,m,kth", [(8, 16, 16), (2, 16, 16), (4, 22, 22), (3, 5, 20), (8, 32, 1)])
def test_kth_smallest_matches_sort(W, m, kth):
    Q = 301
    g = torch.Generator().manual_seed(W * 100 + m)
    v = torch.rand(W, Q, m, generator=g)
    v = (v * 64).floor() / 64                          # ties
    v[torch.rand(W, Q, m, generator=g) < 0.3] = float("inf")
    out = ops.kth_smallest(v.cuda(), kth).cpu()
    flat = v.permute(1, 0, 2).reshape(Q, W * m)
    want = flat.sort(dim=1).values[:, kth - 1] if kth <= W * m else torch.full((Q,), float("inf"))
    assert torch.equal(out, want)


@pytest.mark.parametrize("metric", ["hyperbolic", "cosine"])
def test_pruned_rerank_equals_rerank_of_surviving_candidates(metric):
    Q, N, D, kp, k = 150, 5000, 256, 16, 10
    index = GalleryIndex(synth.gaussian_features(N, D, seed=0).cuda(), metric=metric)
    q32, cs, ci, cnt, _ = index.score_candidates(synth.gaussian_features(Q, D, seed=1).cuda(), k=k, kprime=kp)
    sel_s, sel_i = ops.cand_select(cs, ci, cnt)
    thr = sel_s[:, 5].clone()                          # keep the 6 best (plus surrogate ties) of every query
    thr[3] = float("-inf")                             # a query that loses every candidate on this shard
    d_p, i_p = index.rerank_candidates(q32, sel_s.unsqueeze(1), sel_i.unsqueeze(1), k, prune_thr=thr)
    keep = sel_s <= thr[:, None]
    masked_i = torch.where(keep, sel_i, torch.full_like(sel_i, -1))
    d_w, i_w = index.rerank_candidates(q32, sel_s.unsqueeze(1).contiguous(), masked_i.unsqueeze(1).contiguous(), k)
    assert torch.equal(i_p, i_w) and torch.equal(d_p, d_w)
    assert bool((i_p[3] == -1).all())
    n_kept = keep.sum(dim=1).clamp(max=k)
    assert torch.equal((i_p >= 0).sum(dim=1), n_kept)


@pytest.mark.parametrize("metric,W", [("hyperbolic", 4), ("cosine", 3)])
def test_pruned_shard_protocol_on_one_gpu_equals_single_index(metric, W):
    """The sharded-serving protocol of dist.ShardedGalleryIndex.search_sharded with the collectives
    replaced by tensor stacking on one GPU: per-shard scoring -> cand_select -> global k'-th surrogate ->
    pruned exact rerank per shard -> merge == search of the whole gallery."""
    from patent_image_retrieval_b200.dist import shard_range
    Q, N, D, kp, k = 200, 12001, 128, 16, 10
    g = synth.gaussian_features(N, D, seed=0).cuda()
    q = synth.gaussian_features(Q, D, seed=1).cuda()
    full = GalleryIndex(g, metric=metric)
    want_d, want_i = full.search(q, k=k, kprime=kp)
    shards, staged = [], []
    for r in range(W):
        lo, hi = shard_range(N, r, W)
        sh = GalleryIndex(g[lo:hi], metric=metric, idx_offset=lo)
        q32, cs, ci, cnt, _ = sh.score_candidates(q, k=k, kprime=kp)
        sel_s, sel_i = ops.cand_select(cs, ci, cnt)
        shards.append(sh)
        staged.append((q32, sel_s, sel_i))
    thr = ops.kth_smallest(torch.stack([s for _, s, _ in staged]), kp)
    lists = [sh.rerank_candidates(q32, s.unsqueeze(1), i.unsqueeze(1), k, prune_thr=thr)
             for sh, (q32, s, i) in zip(shards, staged)]
    n_rescored = sum(int((s <= thr[:, None]).sum()) for _, s, _ in staged)
    got_d, got_i = ops.merge_topk(torch.stack([d for d, _ in lists]), torch.stack([i for _, i in lists]),
                                  descending=(metric == "cosine"))
    assert torch.equal(got_i, want_i) and torch.equal(got_d, want_d)
    assert n_rescored <= Q * kp * 1.05                 # the shards together rescore ~k' rows per query, not W*k'


def test_routed_kernels_store_what_the_collectives_would_deliver():
    """hypret_{cand_select,kth_smallest,rerank_pruned}_route with a 3-"rank" route whose receive buffers all live on
    this GPU: every routed row must land where an all_to_all / all_gather of equal blocks would put it
    (row me*Ql + q % Ql of the owner q // Ql), with the values of the unrouted kernels."""
    import ctypes
    from patent_image_retrieval_b200 import _lib
    lib = _lib.load()
    W, me, Ql, N, D, kp, k = 3, 1, 70, 4000, 128, 16, 10
    Q = W * Ql
    index = GalleryIndex(synth.gaussian_features(N, D, seed=0).cuda())
    q32, cs, ci, cnt, _ = index.score_candidates(synth.gaussian_features(Q, D, seed=1).cuda(), k=k, kprime=kp)
    sel_s, sel_i = ops.cand_select(cs, ci, cnt)
    off_sel, off_thr, off_ls, off_li = 0, 1 << 16, 1 << 17, 1 << 18
    bufs = [torch.full((1 << 19,), 255, dtype=torch.uint8, device="cuda") for _ in range(W)]
    route = _lib.PeerRoute()
    route.n_ranks, route.me, route.ql = W, me, Ql
    for r in range(W):
        route.base[r] = bufs[r].data_ptr()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    region = lambda r, off, width, dt: bufs[r][off:off + Q * width * torch.empty((), dtype=dt).element_size()].view(dt).view(W, Ql, width)

    ss, si = torch.empty_like(sel_s), torch.empty_like(sel_i)
    _lib.check(lib.hypret_cand_select_route(ptr(cs), ptr(ci), ptr(cnt), Q, cs.shape[1], kp, ptr(ss), ptr(si),
                                            ctypes.byref(route), off_sel, stream))
    torch.cuda.synchronize()
    assert torch.equal(ss, sel_s) and torch.equal(si, sel_i)
    for r in range(W):
        assert torch.equal(region(r, off_sel, kp, torch.float32)[me], sel_s[r * Ql:(r + 1) * Ql])

    recv = torch.rand(W, Ql, kp, device="cuda")
    want = ops.kth_smallest(recv, kp)
    own = torch.empty(Ql, device="cuda")
    _lib.check(lib.hypret_kth_smallest_route(ptr(recv), W, Ql, kp, kp, ptr(own), ctypes.byref(route), off_thr, stream))
    torch.cuda.synchronize()
    assert torch.equal(own, want)
    for r in range(W):
        assert torch.equal(region(r, off_thr, 1, torch.float32).view(W, Ql)[me], want)

    thr = sel_s[:, 5].contiguous()
    d_w, i_w = index.rerank_candidates(q32, sel_s.unsqueeze(1), sel_i.unsqueeze(1), k, prune_thr=thr)
    _lib.check(lib.hypret_rerank_pruned_route(ptr(q32), ptr(index.rows32), Q, N, D, 1.0, ops.METRIC["hyperbolic"],
                                              ptr(sel_s), ptr(sel_i), 1, kp, k, 0, ptr(thr), ctypes.byref(route),
                                              off_ls, off_li, stream))
    torch.cuda.synchronize()
    for r in range(W):
        assert torch.equal(region(r, off_ls, k, torch.float32)[me], d_w[r * Ql:(r + 1) * Ql])
        assert torch.equal(region(r, off_li, k, torch.int64)[me], i_w[r * Ql:(r + 1) * Ql])
        other = [b for b in range(W) if b != me]
        assert bool((bufs[r][off_ls:off_ls + Q * k * 4].view(W, -1)[other] == 255).all())     # only block `me` written
    # argument checks: Q must match the route
    assert lib.hypret_cand_select_route(ptr(cs), ptr(ci), ptr(cnt), Q - 1, cs.shape[1], kp, ptr(ss), ptr(si),
                                        ctypes.byref(route), off_sel, stream) == -1
