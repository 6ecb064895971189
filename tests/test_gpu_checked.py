"""The CHECKED build (libhypret_checked.so, -DHYPRET_CHECKED: a device-side assert on every guarded index) under ragged
and degenerate shapes.  compute-sanitizer is closed on this GPU pool, so this is the memory-safety net of the
warp-specialised kernels: an out-of-range index traps (the subprocess dies with cudaErrorAssert) instead of passing
silently.  Runs in a subprocess because the library is chosen at first load (HYPRET_CHECKED=1)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ["HYPRET_ROOT"])
from patent_image_retrieval_b200 import GalleryIndex, _lib, ops, synth, train, models
assert _lib.load()._name.endswith("libhypret_checked.so"), _lib.load()._name
torch.manual_seed(0)
# search: ragged Q / N / D, tiny galleries (open candidate sets), both metrics, k up to the wide path, duplicates
for (Q, N, D, k) in [(1, 5, 64, 3), (130, 257, 96, 10), (77, 4099, 512, 10), (300, 20011, 256, 10), (65, 3001, 768, 26),
                     (40, 9000, 128, 100), (129, 513, 2048, 10)]:
    for metric in ("hyperbolic", "cosine"):
        g = synth.gaussian_features(N, D, seed=0).cuda()
        g[N // 2:] = g[: N - N // 2].clone()              # exact duplicates: the certificate must fall back
        index = GalleryIndex(g, c=0.7, metric=metric)
        d, i = index.search(synth.gaussian_features(Q, D, seed=1).cuda(), k=min(k, N))
        torch.cuda.synchronize()
        assert bool((i[:, 0] >= 0).all()) and bool((i < N).all())
want_s, want_i = ops.exact_topk(index.rows32[:7].contiguous(), index.rows32, index.rows_sq64, 0.7, "cosine", 10)
# train_hyp: flash path on ragged n / m, generic path on D = 200
for (n, D) in [(1, 16), (129, 128), (300, 64), (257, 200)]:
    a = torch.nn.functional.normalize(torch.randn(n, D, device="cuda"), dim=1) * 0.6
    p = (a + 0.05 * torch.randn_like(a)) * 0.9
    a.requires_grad_(True); p.requires_grad_(True)
    loss = train.in_batch_contrastive_loss(a, p, torch.tensor([-1.0]), 0.2)
    loss.backward()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(a.grad).all()) and bool(torch.isfinite(p.grad).all())
# projection head: forward + backward kernels, ragged batch
m = models.DeeperHyperbolicEncoder(96, [48], 32, c=1.0, dropout_rate=0.1).cuda().train()
y = m(torch.randn(131, 96, device="cuda"))
y.square().sum().backward()
m.eval()
with torch.no_grad():
    y2 = m(torch.randn(3, 96, device="cuda"))
torch.cuda.synchronize()
assert bool(torch.isfinite(y2).all())
print("CHECKED OK")
"""


def test_checked_build_runs_ragged_shapes_without_tripping_an_assert(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HYPRET_ROOT=str(ROOT), HYPRET_CHECKED="1")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "CHECKED OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
