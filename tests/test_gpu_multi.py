"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): both exchange patterns of
dist.ShardedGalleryIndex over NCCL must reproduce the single-GPU search of the whole gallery
(the reference has no distributed path)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HYPRET_ROOT"])
from patent_image_retrieval_b200 import GalleryIndex, ops, synth
from patent_image_retrieval_b200.dist import ShardedGalleryIndex, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
N, D, Ql, k = 20011, 256, 301, 10
g = synth.gaussian_features(N, D, seed=0).to(dev)
lo, hi = shard_range(N, rank, world)
full = GalleryIndex(g, c=1.0)
for metric in ("hyperbolic", "cosine"):
    full = GalleryIndex(g, c=1.0, metric=metric)
    rep = ShardedGalleryIndex(g[lo:hi], lo, N, metric=metric, queries="replicated")
    shd = ShardedGalleryIndex(g[lo:hi], lo, N, metric=metric, queries="sharded")
    q_same = synth.gaussian_features(Ql, D, seed=1).to(dev)
    q_own = synth.gaussian_features(Ql, D, seed=10 + rank).to(dev)
    d0, i0 = full.search(q_same, k=k)
    d1, i1 = rep.search(q_same, k=k)
    assert torch.equal(i0, i1) and torch.equal(d0, d1), metric + " replicated"
    d2, i2 = full.search(q_own, k=k)
    d3, i3 = shd.search(q_own, k=k)
    assert torch.equal(i2, i3) and torch.equal(d2, d3), metric + " sharded"
    # the peer-memory exchange (projection stores into every rank's buffer, fp32 rows by copy engine) must be the
    # path that ran, over several steps (two buffer slots), and must equal the NCCL all_gather path bit for bit
    assert shd._exchange is not None, "peer exchange not used"
    nccl = ShardedGalleryIndex(g[lo:hi], lo, N, metric=metric, queries="sharded")
    nccl._exchange_ok = False
    for step in range(5):
        q_step = synth.gaussian_features(Ql, D, seed=100 + 10 * step + rank).to(dev)
        d4, i4 = shd.search(q_step, k=k)
        d5, i5 = nccl.search(q_step, k=k)
        d6, i6 = full.search(q_step, k=k)
        assert torch.equal(i4, i5) and torch.equal(d4, d5), metric + " peer vs nccl"
        assert torch.equal(i4, i6) and torch.equal(d4, d6), metric + " peer vs single"
    assert nccl._exchange is None
    # peer-memory query exchange + NCCL for the rest of the protocol
    os.environ["HYPRET_PEER_ROUTE"] = "0"
    d7, i7 = shd.search(q_step, k=k)
    del os.environ["HYPRET_PEER_ROUTE"]
    assert torch.equal(i7, i6) and torch.equal(d7, d6), metric + " peer queries + nccl lists"
    shd._exchange.check()
    shd._exchange.close()
# the exact-top-k guarantee of the sharded path: on a near-duplicate gallery the owners cannot certify most merged lists;
# the flagged queries are rescanned exactly on every shard and the result must be the exact scan of the whole gallery
from oracle import head as ohead
for metric in ("hyperbolic", "cosine"):
    gal, qry, _, _ = synth.clustered_features(6000, 2 * 128, 256, noise=1e-3, per_class=24)
    space = "euclidean"
    if metric == "hyperbolic":
        gal, qry, space = ohead.embed_rows(gal, 1.0), ohead.embed_rows(qry, 1.0), "ball"
    gal, qry = gal.to(dev), qry.to(dev)
    lo2, hi2 = shard_range(gal.shape[0], rank, world)
    full = GalleryIndex(gal, c=1.0, metric=metric, space=space)
    want_d, want_i = ops.exact_topk(qry[rank * 128:(rank + 1) * 128].contiguous(), full.rows32, full.rows_sq64,
                                    1.0, metric, k)
    for peer in (True, False):
        shd = ShardedGalleryIndex(gal[lo2:hi2], lo2, gal.shape[0], metric=metric, space=space, queries="sharded")
        if not peer:
            shd._exchange_ok = False
        got_d, got_i = shd.search(qry[rank * 128:(rank + 1) * 128], k=k)
        n_unc = int(shd.uncertified.sum())
        assert n_unc > 0, "near-duplicate gallery: expected uncertified merged lists"
        assert torch.equal(got_i, want_i) and torch.equal(got_d, want_d), (metric, peer, "sharded exactness", n_unc)
        loose_d, loose_i = shd.search_sharded(qry[rank * 128:(rank + 1) * 128], k=k, exact=False)
        assert shd.uncertified is None
        if shd._exchange is not None:
            shd._exchange.check()
            shd._exchange.close()
    # gaussian data: everything certified, flags all zero
shd = ShardedGalleryIndex(g[lo:hi], lo, N, queries="sharded")
shd.search(synth.gaussian_features(Ql, D, seed=10 + rank).to(dev), k=k)
assert int(shd.uncertified.sum()) == 0
shd._exchange.close()
# collective 2: exact full-ranking AP from all-reduced keys / rank counts == the unsharded computation
from patent_image_retrieval_b200.dist import full_ranking_ap
from patent_image_retrieval_b200 import ops
pts, _, _ = ops.project_rows(g, 1.0, mode="expmap0", side="gallery", want_operand=False)
qp, _, _ = ops.project_rows(synth.gaussian_features(Ql, D, seed=1).to(dev), 1.0, mode="expmap0", side="query",
                            want_operand=False)
gen = torch.Generator().manual_seed(4)
items = torch.randint(0, N, (Ql * 3,), generator=gen).to(dev)
off = torch.arange(0, Ql * 3 + 1, 3, dtype=torch.int64, device=dev)
for grouped in (True, False):
    single_keys = ops.pair_keys(qp, pts, off, items, 1.0, "hyperbolic")
    c1, b1 = ops.rank_count(qp, pts, off, items, single_keys, 1.0, "hyperbolic")
    want = ops.ap_from_counts(off, items, single_keys, c1, b1, N, grouped_ties=grouped)
    got = full_ranking_ap(qp, pts[lo:hi].contiguous(), off, items, row_offset=lo, n_total=N, grouped_ties=grouped,
                          sharded=True)
    # unsharded call under an initialised multi-rank group: no collective, the single-GPU result
    solo = full_ranking_ap(qp, pts, off, items, n_total=N, grouped_ties=grouped)
    assert torch.equal(solo[1], want[1]) and solo[0] == want[0], "unsharded full_ranking_ap under a process group"
    assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]) and got[0] == want[0], "full_ranking_ap"
# sharded negatives: the in-batch InfoNCE over a batch split across the ranks == the single-GPU loss / gradients
from patent_image_retrieval_b200 import train
from oracle import head
nl, dd, cc, tau = 640, 64, 0.9, 0.2
n_all = nl * world
mu = synth.gaussian_features(n_all, dd, seed=21, scale=1.0)
a_all = head.embed_rows(mu + 0.3 * synth.gaussian_features(n_all, dd, seed=22, scale=1.0), cc).to(dev)
p_all = head.embed_rows(mu + 0.3 * synth.gaussian_features(n_all, dd, seed=23, scale=1.0), cc).to(dev)
kk = torch.tensor([-cc])
for sym in (False, True):
    ag, pg = a_all.clone().requires_grad_(True), p_all.clone().requires_grad_(True)
    ref = train.InBatchInfoNCE.apply(ag, pg, cc, tau, sym)
    ref.backward()
    al = a_all[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    pl = p_all[rank * nl:(rank + 1) * nl].clone().requires_grad_(True)
    got = train.sharded_in_batch_contrastive_loss(al, pl, kk, tau, symmetric=sym)
    got.backward()
    assert abs(float(got) - float(ref)) <= 2e-6 * abs(float(ref)), ("sharded loss", float(got), float(ref))
    for g_, w_ in ((al.grad, ag.grad[rank * nl:(rank + 1) * nl]), (pl.grad, pg.grad[rank * nl:(rank + 1) * nl])):
        assert float((g_ - w_).abs().max() / w_.abs().max()) < 2e-5, "sharded grads"
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_exchange_patterns_equal_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HYPRET_ROOT=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(29800 + os.getpid() % 100), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-8000:]
    assert out.stdout.count("OK") == 2
