"""GPU parity of the matrix-free full-ranking AP (include/hypret.h: hypret_pair_keys, hypret_rank_count,
hypret_ap_from_counts; SURVEY.md 8e collective 2) against the dense path it replaces
(hypret_pairdist -> hypret_ap_full, itself pinned to sklearn / the notebook loops by the ref_* goldens),
unsharded and with the gallery cut into shards whose keys / counts are summed like the all-reduces do."""
import numpy as np
import pytest
import torch

from oracle import head, retrieval
from patent_image_retrieval_b200 import ops, synth
from patent_image_retrieval_b200.dist import full_ranking_ap, shard_range

pytestmark = pytest.mark.gpu


def _case(Q, N, D, seed, max_pos=6, dup_rows=True):
    g = torch.Generator().manual_seed(seed)
    gal = head.embed_rows(synth.gaussian_features(N, D, seed=seed), 1.0)
    qry = head.embed_rows(synth.gaussian_features(Q, D, seed=seed + 1), 1.0)
    if dup_rows and N > 40:                      # exact score ties: duplicated gallery rows
        gal[10] = gal[3]
        gal[N - 1] = gal[3]
        gal[20:24] = gal[7]
    lists = []
    for q in range(Q):
        n = int(torch.randint(0, max_pos + 1, (1,), generator=g))
        ids = torch.randint(0, N, (n,), generator=g).tolist()
        if q % 7 == 0 and N > 40:
            ids += [3, 10, 22]                   # positives inside tie groups
        if q % 11 == 0:
            ids += [-1, N, N + 5]                # out-of-range ids are ignored (src/train.py:3224-3242)
        lists.append(ids)
    lists[1] = []                                # a query without positives
    off = torch.zeros(Q + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor([len(x) for x in lists]), 0)
    items = torch.tensor([v for x in lists for v in x], dtype=torch.int64)
    return qry.cuda(), gal.cuda(), off.cuda(), items.cuda()


@pytest.mark.parametrize("Q,N,D,c", [(150, 3000, 128, 1.0), (64, 257, 20, 0.5), (5, 40, 256, 2.0), (130, 9000, 512, 1.0)])
@pytest.mark.parametrize("grouped", [True, False])
def test_rank_count_ap_equals_dense_ap(Q, N, D, c, grouped):
    qry, gal, off, items = _case(Q, N, D, seed=Q + N)
    if c != 1.0:                                 # keep the points inside the ball of curvature c
        qry, gal = qry * 0.5, gal * 0.5
    d = ops.pairdist(qry, gal, c)
    want_mean, want_ap, want_valid = ops.ap_full(-d, off, items.clamp(min=-1), grouped_ties=grouped)
    got_mean, got_ap, got_valid = full_ranking_ap(qry, gal, off, items, c=c, metric="hyperbolic", grouped_ties=grouped)
    assert torch.equal(got_valid, want_valid)
    # same ranks -> the same rational numbers; only the fp64 summation order over a query's positives differs
    torch.testing.assert_close(got_ap, want_ap, rtol=1e-13, atol=1e-16)
    assert got_mean == pytest.approx(want_mean, rel=1e-13)
    # keys are the dense kernel's distances and the counts are the dense matrix's rank counts, bit for bit
    keys = ops.pair_keys(qry, gal, off, items, c, "hyperbolic")
    rows = torch.repeat_interleave(torch.arange(Q, device="cuda"), off[1:] - off[:-1])
    ok = (items >= 0) & (items < N)
    assert torch.equal(keys[ok], d[rows[ok], items[ok]])
    counts, bad = ops.rank_count(qry, gal, off, items, keys, c, "hyperbolic")
    dr = d[rows[ok]]                                              # [nnz_ok, N]
    kk = keys[ok][:, None]
    col = torch.arange(N, device="cuda")[None, :]
    assert torch.equal(counts[ok, 0], (dr < kk).sum(dim=1))
    assert torch.equal(counts[ok, 2], (dr == kk).sum(dim=1))
    assert torch.equal(counts[ok, 1], ((dr == kk) & (col < items[ok][:, None])).sum(dim=1))
    assert int(bad.sum()) == 0


def test_rank_count_more_positives_than_one_sweep_holds():
    """> 1024 positives in one 64-query tile: the kernel sweeps its columns once per chunk of positives."""
    Q, N, D = 70, 1500, 64
    qry, gal, _, _ = _case(Q, N, D, seed=5)
    g = torch.Generator().manual_seed(9)
    lists = [torch.randint(0, N, (40,), generator=g).tolist() for _ in range(Q)]
    off = torch.arange(0, 40 * Q + 1, 40, dtype=torch.int64).cuda()
    items = torch.tensor([v for x in lists for v in x], dtype=torch.int64).cuda()
    d = ops.pairdist(qry, gal, 1.0)
    for grouped in (True, False):
        _, want_ap, want_valid = ops.ap_full(-d, off, items, grouped_ties=grouped)
        _, got_ap, got_valid = full_ranking_ap(qry, gal, off, items, grouped_ties=grouped)
        assert torch.equal(got_valid, want_valid)
        torch.testing.assert_close(got_ap, want_ap, rtol=1e-13, atol=1e-16)


def test_rank_count_flags_queries_with_nonfinite_scores():
    qry, gal, off, items = _case(40, 300, 32, seed=3)
    qry[4] = float("nan")
    d = ops.pairdist(qry, gal, 1.0)
    _, want_ap, want_valid = ops.ap_full(-d, off, items.clamp(min=-1), grouped_ties=True)
    _, got_ap, got_valid = full_ranking_ap(qry, gal, off, items, grouped_ties=True)
    assert int(got_valid[4]) == 0
    assert torch.equal(got_valid, want_valid)
    torch.testing.assert_close(got_ap, want_ap, rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("W", [2, 5])
@pytest.mark.parametrize("grouped", [True, False])
def test_sharded_keys_and_counts_sum_to_the_unsharded_result(W, grouped):
    Q, N, D = 150, 4001, 128
    qry, gal, off, items = _case(Q, N, D, seed=17)
    want = full_ranking_ap(qry, gal, off, items, grouped_ties=grouped)
    keys = torch.zeros(items.numel(), device="cuda")
    for r in range(W):                                            # all_reduce(SUM) of the keys
        lo, hi = shard_range(N, r, W)
        keys += ops.pair_keys(qry, gal[lo:hi], off, items, 1.0, "hyperbolic", idx_offset=lo)
    counts = torch.zeros(items.numel(), 3, dtype=torch.int64, device="cuda")
    bad = torch.zeros(Q, dtype=torch.int32, device="cuda")
    for r in range(W):                                            # all_reduce(SUM) of the counts
        lo, hi = shard_range(N, r, W)
        c_r, b_r = ops.rank_count(qry, gal[lo:hi], off, items, keys, 1.0, "hyperbolic", idx_offset=lo)
        counts += c_r
        bad += b_r
    got = ops.ap_from_counts(off, items, keys, counts, bad, N, grouped_ties=grouped)
    assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]) and got[0] == want[0]
    # every valid positive is ranked somewhere in 1..N and ties include itself exactly once
    ok = (items >= 0) & (items < N)
    assert bool((counts[ok, 2] >= 1).all()) and bool((counts[ok, 0] + counts[ok, 2] <= N).all())


def test_cosine_full_ranking_ap_matches_oracle_notebook_ap():
    """Ranking convention of notebooks/retrieval.ipynb:383,411-420 on cosine scores."""
    Q, N, D = 60, 2000, 96
    g = torch.Generator().manual_seed(2)
    gal = torch.randn(N, D, generator=g)
    qry = torch.randn(Q, D, generator=g)
    gal[5] = 0.0                                                  # zero row: similarity 0 (sklearn normalize)
    lists = [torch.randint(0, N, (4,), generator=g).tolist() for _ in range(Q)]
    off = torch.arange(0, 4 * Q + 1, 4, dtype=torch.int64)
    items = torch.tensor([v for x in lists for v in x], dtype=torch.int64)
    sim = retrieval.cosine_similarity(qry.double().numpy(), gal.double().numpy())
    want = []
    for q in range(Q):
        order = np.lexsort((np.arange(N), -sim[q]))              # descending similarity, ties -> lower index
        rank = np.empty(N, dtype=np.int64)
        rank[order] = np.arange(1, N + 1)
        rs = np.sort(rank[np.array(lists[q])])
        want.append(float(np.mean([(i + 1) / r for i, r in enumerate(rs)])))
    _, ap, valid = full_ranking_ap(qry.cuda(), gal.cuda(), off.cuda(), items.cuda(), metric="cosine",
                                   grouped_ties=False)
    assert bool(valid.bool().all())
    np.testing.assert_allclose(ap.cpu().numpy(), np.array(want), rtol=0, atol=2e-3)
    assert abs(float(ap.mean()) - float(np.mean(want))) < 2e-4


def test_full_ranking_metric_suite_matches_notebook_loops(tmp_path):
    """Every metric of the notebook's evaluation cell from rank counts == the notebook loops (oracle restatement,
    pinned to the notebook's own code by the ref_nb_* goldens) over the fully sorted lists; plus the results JSON."""
    import json
    from patent_image_retrieval_b200 import evaluation, io as pio
    Q, N, D = 40, 700, 64
    g = torch.Generator().manual_seed(11)
    gal = torch.randn(N, D, generator=g)
    qry = torch.randn(Q, D, generator=g)
    gal[9] = gal[4]                                                # a tie: lower index first
    gallery_paths = [f"/gallery/p{i // 3}/img_{i}.png" for i in range(N)]
    query_names = [f"/queries/q_{i}.png" for i in range(Q)]
    gt = {}
    for i in range(Q):
        if i == 7:
            continue                                              # no ground truth: skipped
        ids = torch.randint(0, N, (int(torch.randint(1, 6, (1,), generator=g)),), generator=g).tolist()
        if i % 5 == 0:
            ids += [4, 9]
        names = [f"img_{j}.png" for j in ids] + (["not_in_gallery.png"] if i % 4 == 0 else [])
        gt[f"q_{i}.png"] = {"patent_positives": names}
    (tmp_path / "gt.json").write_text(json.dumps(gt))
    res = evaluation.evaluate_test_set(gal.numpy(), gallery_paths, qry.numpy(), query_names, tmp_path / "gt.json",
                                       results_path=tmp_path / "results" / "evaluation_results_test.json")
    sim = retrieval.cosine_similarity(qry.double().numpy(), gal.double().numpy())
    keep = [i for i in range(Q) if i != 7]
    ranked = [np.lexsort((np.arange(N), -sim[i])).tolist() for i in keep]
    row_of = {f"img_{j}.png": j for j in range(N)}
    want = {"mrr": [], "ap": [], "ndcg": []}
    for k in (5, 10, 20):
        want.update({f"mrr@{k}": [], f"precision@{k}": [], f"recall@{k}": []})
    for r, i in zip(ranked, keep):
        names = set(gt[f"q_{i}.png"]["patent_positives"])
        pos = {row_of.get(n, N + 1000 + hash(n) % 1000) for n in names}      # missing names never match a row
        assert len(pos) == len(names)
        want["mrr"].append(retrieval.mrr_at_k(r, pos, N))
        want["ap"].append(retrieval.average_precision_ranked(r, pos))
        want["ndcg"].append(retrieval.ndcg_ranked(r, pos))
        for k in (5, 10, 20):
            want[f"mrr@{k}"].append(retrieval.mrr_at_k(r, pos, k))
            want[f"precision@{k}"].append(retrieval.precision_at_k(r, pos, k))
            want[f"recall@{k}"].append(retrieval.recall_at_k(r, pos, k))
    qw = res["query_wise_metrics"]
    pairs = [("reciprocal_ranks", "mrr"), ("reciprocal_ranks@5", "mrr@5"), ("reciprocal_ranks@20", "mrr@20"),
             ("ap_scores", "ap"), ("ndcg_scores", "ndcg"), ("recall_5", "recall@5"), ("recall_10", "recall@10"),
             ("recall_20", "recall@20"), ("precision_5", "precision@5"), ("precision_10", "precision@10"),
             ("precision_20", "precision@20")]
    for js_name, name in pairs:
        np.testing.assert_allclose(qw[js_name], want[name], rtol=1e-12, atol=1e-15, err_msg=name)
    assert res["summary_metrics"]["mAP"] == pytest.approx(np.mean(want["ap"]), rel=1e-12)
    assert json.load(open(tmp_path / "results" / "evaluation_results_test.json")) == res
