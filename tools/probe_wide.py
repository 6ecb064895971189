"""Where the wide path (adaptive list width) spends its time on a clustered gallery: per-stage CUDA-event
times of search(kprime=64) and search(kprime=32), C2-sized, 30 rows per class at 3 % within-class noise.
Usage (GPU box): python tools/probe_wide.py"""
import sys, torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import GalleryIndex, synth
from patent_image_retrieval_b200.retrieval import StageEvents
N, Q, D, k = 300_000, 10_000, 512, 10
gal, qry, _, _ = synth.clustered_features(N, Q, D, device="cuda", per_class=30, noise=0.03)
index = GalleryIndex(gal, c=1.0, metric="hyperbolic")
for kp in (64, 32):
    for _ in range(2):
        index.search(qry, k=k, kprime=kp)
    torch.cuda.synchronize()
    ev = StageEvents()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        index.search(qry, k=k, kprime=kp, kernel_events=ev)
    e1.record()
    torch.cuda.synchronize()
    extra = int(index.uncertified_wide.sum()) if kp > 32 else int(index.certificate.count[0])
    print(f"kprime {kp}: {e0.elapsed_time(e1)/3:.2f} ms per search; stages", {s: round(ev.ms(s), 3) for s in ev.STAGES}, "uncertified", extra, flush=True)
