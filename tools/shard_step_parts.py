"""Where one rank's step of the sharded-serving protocol goes, emulated on ONE GPU at the N=8 shape of bench.py's
weak-scaling run (80k queries against a 37.5k-row shard; the NCCL exchanges are left out, thresholds are faked so
that ~k'/W candidates per query survive the pruning).  CUDA-event time per kernel."""
import sys
import torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, ops, synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Ql, N, D, k, kp = 10_000, 300_000, 512, 10, 16


def tm(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


g = synth.gaussian_features(N // W, D, seed=0, device="cuda")
q = synth.gaussian_features(W * Ql, D, seed=1, device="cuda")
sh = GalleryIndex(g)
q32, q_op, _ = ops.project_rows(q, 1.0, side="query")
_, cs, ci, cnt = sh.score_projected(q32, q_op, k=k, kprime=kp)
sel_s, sel_i = ops.cand_select(cs, ci, cnt)
recv = sel_s.view(W, Ql, kp).contiguous()
thr_own = ops.kth_smallest(recv, kp)
thr = sel_s[:, max(1, kp // W) - 1].contiguous()          # keeps k'/W candidates per query on this shard
lists = sh.rerank_candidates(q32, sel_s.unsqueeze(1), sel_i.unsqueeze(1), k, prune_thr=thr)
ls, li = lists[0].view(W, Ql, k).contiguous(), lists[1].view(W, Ql, k).contiguous()
print(f"W={W}: shard {N // W} rows, {W * Ql} queries, lists per query: mean {float(cnt.float().mean()):.2f} max {int(cnt.max())}")
print("project (own 10k rows)   %.3f ms" % tm(lambda: ops.project_rows(q[:Ql], 1.0, side="query")))
print("project (all W*Ql rows)  %.3f ms" % tm(lambda: ops.project_rows(q, 1.0, side="query")))
print("score_topk               %.3f ms" % tm(lambda: sh.score_projected(q32, q_op, k=k, kprime=kp)))
print("cand_select              %.3f ms" % tm(lambda: ops.cand_select(cs, ci, cnt)))
print("kth_smallest             %.3f ms" % tm(lambda: ops.kth_smallest(recv, kp)))
print("rerank pruned            %.3f ms" % tm(lambda: sh.rerank_candidates(q32, sel_s.unsqueeze(1), sel_i.unsqueeze(1), k, prune_thr=thr)))
print("rerank unpruned          %.3f ms" % tm(lambda: sh.rerank_candidates(q32, cs, ci, k, list_count=cnt)))
print("merge_topk               %.3f ms" % tm(lambda: ops.merge_topk(ls, li)))
