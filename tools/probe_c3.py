"""C3-shaped probe (BASELINE.json configs[2]: D=768, top-100, cosine + hyperbolic, 1M gallery), query count
scaled down so that it fits a short run: per-metric step time, queries/s, certified fraction."""
import sys
import torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, StageEvents, synth  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
v = synth.gaussian_features(1_000_000, 768, seed=0, device="cuda")
u = synth.gaussian_features(Q, 768, seed=1, device="cuda")
for metric in ("hyperbolic", "cosine"):
    idx = GalleryIndex(v, metric=metric)
    for _ in range(2):
        idx.search(u, k=100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    ev = StageEvents()
    for _ in range(3):
        s, i, m = idx.search(u, k=100, return_margin=True, kernel_events=ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("   stages (ms): " + ", ".join(f"{n} {ev.ms(n):.2f}" for n in ev.STAGES), flush=True)
    print(f"{metric}: {Q} x 1M x 768 top-100: {ms:.1f} ms/step  {Q / ms * 1e3:.0f} q/s  "
          f"{2 * Q * 1e6 * 768 / ms / 1e9:.0f} TFLOP/s over the whole step; margin>0 on {float((m > 0).float().mean()) * 100:.1f}% of queries",
          flush=True)
    del idx
