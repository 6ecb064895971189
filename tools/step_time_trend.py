import sys, torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, synth
g = synth.gaussian_features(300000, 512, seed=0, device="cuda")
q = synth.gaussian_features(10000, 512, seed=1, device="cuda")
idx = GalleryIndex(g)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(401)]
evs[0].record()
for i in range(400):
    idx.search(q, k=10, kprime=16)
    evs[i + 1].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(400)]
for lo in (0, 5, 10, 20, 40, 80, 160, 320):
    hi = min(400, lo * 2 if lo else 5)
    print(f"steps {lo:3d}-{hi:3d}: {sum(ms[lo:hi]) / (hi - lo):.3f} ms")
