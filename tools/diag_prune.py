"""Diagnostic: where does the pruned shard protocol differ from the single-index search?"""
import sys, torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, ops, synth
from patent_image_retrieval_b200.dist import shard_range
metric, W = "hyperbolic", 4
Q, N, D, kp, k = 200, 12001, 128, 16, 10
g = synth.gaussian_features(N, D, seed=0).cuda()
q = synth.gaussian_features(Q, D, seed=1).cuda()
full = GalleryIndex(g, metric=metric)
q32f, csf, cif, cntf, _ = full.score_candidates(q, k=k, kprime=kp)
self_s, self_i = ops.cand_select(csf, cif, cntf)
want_d, want_i = full.rerank_candidates(q32f, csf, cif, k, list_count=cntf)
staged, shards = [], []
for r in range(W):
    lo, hi = shard_range(N, r, W)
    sh = GalleryIndex(g[lo:hi], metric=metric, idx_offset=lo)
    q32, cs, ci, cnt, _ = sh.score_candidates(q, k=k, kprime=kp)
    s, i = ops.cand_select(cs, ci, cnt)
    shards.append(sh); staged.append((q32, s, i, lo))
thr = ops.kth_smallest(torch.stack([s for _, s, _, _ in staged]), kp)
print("thr == single k'-th surrogate:", float((thr == self_s[:, kp - 1]).float().mean()))
lists = [sh.rerank_candidates(q32, s.unsqueeze(1), i.unsqueeze(1), k, prune_thr=thr) for sh, (q32, s, i, _) in zip(shards, staged)]
got_d, got_i = ops.merge_topk(torch.stack([d for d, _ in lists]), torch.stack([i for _, i in lists]))
bad = torch.nonzero(~((got_i == want_i).all(1) & (got_d == want_d).all(1))).flatten().tolist()
print("rows differing:", len(bad), bad[:10])
for r in bad[:3]:
    print("row", r)
    print(" want", want_i[r].tolist(), want_d[r].tolist())
    print(" got ", got_i[r].tolist(), got_d[r].tolist())
    print(" single sel", self_i[r].tolist(), self_s[r].tolist())
    allc = torch.cat([torch.where(i[r] >= 0, i[r].long() + lo, i[r].long()) for _, s, i, lo in staged]); alls = torch.cat([s[r] for _, s, i, lo in staged])
    o = alls.argsort()[:kp + 2]
    print(" union  ", allc[o].tolist(), alls[o].tolist(), "thr", float(thr[r]))
