"""The certified search on CLUSTERED galleries (the reference's data are figures of patent families: tight classes),
C2-sized: 10k queries x 300k rows x 512, top-10.  Per within-class noise level: time per search, certified fraction,
number of queries that fell to the exact scan.   python tools/bench_clustered.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import GalleryIndex, synth  # noqa: E402

N, Q, D, k = 300_000, 10_000, 512, 10
for metric in ("hyperbolic", "cosine"):
    for per_class, noise in ((8, 0.3), (30, 0.1), (30, 0.03), (30, 0.01), (100, 0.003)):
        gal, qry, _, _ = synth.clustered_features(N, Q, D, device="cuda", per_class=per_class, noise=noise)
        index = GalleryIndex(gal, c=1.0, metric=metric)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        index.adaptive = False                       # the narrow path alone: k' = 16 lists, bound of the 24th best
        index.search(qry, k=k)
        torch.cuda.synchronize()
        e0.record()
        index.search(qry, k=k)
        e1.record()
        torch.cuda.synchronize()
        cert = index.certificate
        line = (f"{metric:10s} {per_class:4d} rows/class, noise {noise:5.3f}: narrow {e0.elapsed_time(e1):8.2f} ms "
                f"(certified {float(cert.certified[:Q].float().mean()):.4f}, exact scans {int(cert.count[0])})")
        index.adaptive = True                        # the default: the index picks the list width from what it saw
        for _ in range(3):
            index.search(qry, k=k)
            torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            index.search(qry, k=k)
        e1.record()
        torch.cuda.synchronize()
        print(line + f" | adaptive {e0.elapsed_time(e1) / 5:8.2f} ms ({index.last_mode})", end="")
        # the wide path (64-slot lists, 256 survivors rescored exactly, margin against the same bound)
        for _ in range(2):
            index.search(qry, k=k, kprime=64)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            index.search(qry, k=k, kprime=64)
        e1.record()
        torch.cuda.synchronize()
        print(f"   | wide lists: {e0.elapsed_time(e1) / 3:8.2f} ms, exact scans {int(index.uncertified_wide.sum())}")
        del index, gal
        torch.cuda.empty_cache()
