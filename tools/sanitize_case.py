"""Smallest end-to-end case for compute-sanitizer: every kernel of the retrieval path once."""
import sys
import torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, ops, synth, train  # noqa: E402

u = synth.gaussian_features(300, 512, seed=1).cuda()
v = synth.gaussian_features(3000, 512, seed=0).cuda()
idx = GalleryIndex(v)
d, i = idx.search(u, k=10)
idx_c = GalleryIndex(v, metric="cosine")
idx_c.search(u, k=10)
off = torch.arange(0, 301, device="cuda")
ops.retrieval_metrics(i, off, i[:, 0].contiguous(), ks=(5, 10))
dm = ops.pairdist(idx.rows32[:64], idx.rows32[:96], 1.0)
ops.ap_full(-dm, torch.arange(0, 65, device="cuda"), torch.arange(0, 64, device="cuda"))
a = idx.rows32[:64].clone().requires_grad_(True)
train.in_batch_contrastive_loss(a, idx.rows32[64:128], torch.tensor([-1.0]), 0.5).backward()
ops.merge_topk(torch.stack([d, d + 1]), torch.stack([i, i + 3000]))
torch.cuda.synchronize()
print("sanitize case done", float(d.sum()))
