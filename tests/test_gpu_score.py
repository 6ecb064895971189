"""GPU parity: tcgen05 scoring kernel.  The debug matrix must equal the fp32 product of the very
same bf16 operands (bf16 x bf16 is exact in fp32; only the accumulation order differs), and the
streamed top-k' lists must equal torch.topk over that matrix, for every split."""
import pytest
import torch

from patent_image_retrieval_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _operands(Q, N, d, c=1.0, metric="hyperbolic"):
    u = synth.gaussian_features(Q, d, seed=1).cuda()
    v = synth.gaussian_features(N, d, seed=0).cuda()
    mode = "expmap0" if metric == "hyperbolic" else "cosine"
    _, q_op, _ = ops.project_rows(u, c, mode, "query")
    _, g_op, _ = ops.project_rows(v, c, mode, "gallery")
    return q_op, g_op


CASES = [
    # Q, N, d, kprime, max_ctas
    (200, 1000, 512, 16, 0),     # resident query tile, ragged last tiles
    (128, 256, 512, 16, 0),      # exactly one tile
    (1, 1, 128, 4, 0),           # degenerate
    (300, 3000, 128, 8, 0),      # many short strips, small d (deep ring)
    (130, 700, 768, 32, 0),      # streamed query tile (d > 512)
    (64, 520, 2048, 16, 0),      # C1-style d
    (257, 5000, 256, 20, 0),     # 32-slot lists
    (900, 5000, 128, 8, 5),      # forced small grid: full wave + phase 1 + phase 2 in one CTA
    (700, 9000, 512, 16, 7),     # resident tile reloaded between strips
    (128 * 3, 256 * 40, 64, 16, 6),
    # the dense cold start of a strip (first 64 columns of each warpgroup's half tile through the sorting network):
    # galleries that end inside / at / just after its 16-column groups, k' below the 16 register slots
    (130, 15, 128, 16, 0),
    (130, 17, 128, 10, 0),
    (256, 63, 512, 16, 0),
    (256, 65, 512, 12, 0),
    (140, 130, 256, 16, 0),
    (300, 257, 128, 5, 0),
    (260, 300, 2048, 16, 0),
    (260, 193, 64, 16, 0),
]


@pytest.mark.parametrize("Q,N,d,kprime,cap", CASES)
def test_scores_and_lists(Q, N, d, kprime, cap):
    q_op, g_op = _operands(Q, N, d)
    cs, ci, dbg = ops.score_topk(q_op, g_op, d, kprime, cap, debug=True, share_thresholds=False)
    ref = q_op.float() @ g_op.float().t()
    scale = float(ref.abs().max())
    assert float((dbg - ref).abs().max()) <= 2e-5 * scale + 1e-6
    plan = ops.score_plan(Q, N, d, kprime, cap)
    assert cs.shape == (Q, plan["n_lists"], kprime)
    written = set()
    pair = plan["pair"]
    eg = plan["epi_groups"]      # a strip owns eg consecutive list slots: one per epilogue warpgroup (alternate tiles)
    for cta, step, qt, g0, g1, slot in ops.score_strips(Q, N, d, kprime, cap):
        r0, r1 = qt * 128, min(Q, (qt + pair) * 128)
        lo, hi = g0 * 256, min(N, g1 * 256)
        for t in range(pair):
            for e in range(eg):
                written.add((qt + t, slot + e))
        kk = min(kprime, hi - lo)
        want_v, _ = torch.topk(dbg[r0:r1, lo:hi], kk, dim=1, largest=False)
        un_s = cs[r0:r1, slot:slot + eg, :].reshape(r1 - r0, -1)
        un_i = ci[r0:r1, slot:slot + eg, :].reshape(r1 - r0, -1)
        un_s = torch.where(un_i >= 0, un_s, torch.full_like(un_s, float("inf")))
        got_v, order = un_s.sort(dim=1)
        got_i = torch.gather(un_i, 1, order)
        # the union of the strip's lists holds exactly the strip's top-k' at its head
        assert torch.equal(got_v[:, :kk], want_v.sort(dim=1).values)
        # indices point at the scores they claim, inside the strip; unused slots are (-1, +inf)
        picked = torch.gather(dbg[r0:r1], 1, got_i[:, :kk].long())
        assert torch.equal(picked, got_v[:, :kk])
        assert int(got_i[:, :kk].min()) >= lo and int(got_i[:, :kk].max()) < hi
        n_valid = (un_i >= 0).sum(dim=1)
        assert bool((n_valid <= hi - lo).all())
        if eg == 1 and kk < kprime:
            assert bool((got_i[:, kk:] == -1).all()) and bool(torch.isinf(got_v[:, kk:]).all())
        if hi - lo <= kprime:        # short strip: every column is kept exactly once
            assert bool((n_valid == hi - lo).all())
    # list slots that no strip owns read as empty
    for qt in range(plan["n_qtiles"]):
        for slot in range(plan["n_lists"]):
            if (qt, slot) not in written:
                assert bool((ci[qt * 128:qt * 128 + 128, slot, :] == -1).all())


@pytest.mark.parametrize("Q,N,d,kprime,cap", CASES)
def test_shared_thresholds_keep_the_global_topk(Q, N, d, kprime, cap):
    """Production mode: strips of a query exchange thresholds.  Lists are then only subsets of
    their strip's top-k', but (a) every entry is a genuine (score, index) pair of its strip, with
    no duplicates, and (b) the union of a query's lists contains its global top-k'."""
    q_op, g_op = _operands(Q, N, d)
    cs, ci, dbg = ops.score_topk(q_op, g_op, d, kprime, cap, debug=True, share_thresholds=True)
    plan = ops.score_plan(Q, N, d, kprime, cap)
    L = plan["n_lists"]
    eg = plan["epi_groups"]
    for cta, step, qt, g0, g1, slot in ops.score_strips(Q, N, d, kprime, cap):
        r0, r1 = qt * 128, min(Q, (qt + plan["pair"]) * 128)
        lo, hi = g0 * 256, min(N, g1 * 256)
        for e in range(eg):
            idx = ci[r0:r1, slot + e, :].long()
            valid = idx >= 0
            assert bool(((idx >= lo) & (idx < hi))[valid].all())
            picked = torch.gather(dbg[r0:r1], 1, idx.clamp_min(0))
            assert torch.equal(picked[valid], cs[r0:r1, slot + e, :][valid])
    flat_i = ci.reshape(Q, L * kprime).long()
    flat_s = torch.where(flat_i >= 0, cs.reshape(Q, L * kprime), torch.full_like(cs.reshape(Q, -1), float("inf")))
    srt = flat_i.sort(dim=1).values
    dup = (srt[:, 1:] == srt[:, :-1]) & (srt[:, 1:] >= 0)
    assert not bool(dup.any())
    kk = min(kprime, N)
    want = torch.topk(dbg, kk, dim=1, largest=False).values.sort(dim=1).values
    got = flat_s.sort(dim=1).values[:, :kk]
    assert torch.equal(got, want)


def test_cosine_operands_give_minus_cosine():
    Q, N, d = 150, 900, 512
    q_op, g_op = _operands(Q, N, d, metric="cosine")
    _, _, dbg = ops.score_topk(q_op, g_op, d, 16, debug=True)
    u = synth.gaussian_features(Q, d, seed=1).cuda()
    v = synth.gaussian_features(N, d, seed=0).cuda()
    cos = torch.nn.functional.normalize(u.double(), dim=1) @ torch.nn.functional.normalize(v.double(), dim=1).t()
    assert float((dbg.double() + cos).abs().max()) < 4e-3       # bf16 operand rounding


def test_plan_mismatch_is_an_error():
    q_op, g_op = _operands(64, 2048, 128)
    cs = torch.empty(64, 77, 16, device="cuda")
    ci = torch.empty(64, 77, 16, device="cuda", dtype=torch.int32)
    with pytest.raises(ValueError):
        ops.score_topk(q_op, g_op, 128, 16, 0, out=(cs, ci))


@pytest.mark.parametrize("Q,N,d,kprime,cap", [(200, 1000, 512, 16, 0), (900, 5000, 128, 8, 5), (257, 5000, 256, 20, 0),
                                              (130, 700, 768, 32, 0), (700, 9000, 512, 16, 7)])
def test_compact_list_slots_hold_the_same_candidates(Q, N, d, kprime, cap):
    """list_count mode: slots are handed out per query in arrival order and empty lists take none -- the multiset of
    (score, index) entries in a query's first list_count[q] slots equals the legacy layout's, and nothing is lost."""
    q_op, g_op = _operands(Q, N, d)
    cs0, ci0 = ops.score_topk(q_op, g_op, d, kprime, cap, share_thresholds=False)
    cnt = torch.full((Q,), -7, dtype=torch.int32, device="cuda")
    plan = ops.score_plan(Q, N, d, kprime, cap)
    cs1 = torch.full((Q, plan["n_lists"], kprime), float("nan"), device="cuda")
    ci1 = torch.full((Q, plan["n_lists"], kprime), 123456789, dtype=torch.int32, device="cuda")     # stale garbage
    ops.score_topk(q_op, g_op, d, kprime, cap, share_thresholds=False, out=(cs1, ci1), list_count=cnt)
    assert int(cnt.min()) >= 0 and int(cnt.max()) <= plan["n_lists"]
    n_nonempty = (ci0 >= 0).any(dim=2).sum(dim=1)
    assert torch.equal(cnt.long(), n_nonempty)
    used = torch.arange(plan["n_lists"], device="cuda")[None, :] < cnt[:, None]                     # [Q, L]
    flat0 = torch.where(ci0 >= 0, ci0, torch.full_like(ci0, 1 << 30)).reshape(Q, -1).sort(dim=1).values
    ci1m = torch.where(used[:, :, None] & (ci1 >= 0), ci1, torch.full_like(ci1, 1 << 30))
    flat1 = ci1m.reshape(Q, -1).sort(dim=1).values
    assert torch.equal(flat0, flat1)
    s0 = torch.where(ci0 >= 0, cs0, torch.full_like(cs0, float("inf"))).reshape(Q, -1).sort(dim=1).values
    s1 = torch.where(used[:, :, None] & (ci1 >= 0), cs1, torch.full_like(cs1, float("inf"))).reshape(Q, -1).sort(dim=1).values
    assert torch.equal(s0, s1)
    # consumers read only the used slots: same selection from both layouts
    a = ops.cand_select(cs0, ci0) if kprime <= 32 else None
    b = ops.cand_select(cs1, ci1, cnt) if kprime <= 32 else None
    if a is not None:
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
