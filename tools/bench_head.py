"""The projection head (DeeperHyperbolicEncoder 512 -> 256 -> 128) on the kernel path against the op-by-op path of the
same module (what autograd ran before): inference and training step (forward + backward), CUDA events, and the number
of kernel launches per call (torch profiler).   python tools/bench_head.py"""
import copy
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from patent_image_retrieval_b200 import models  # noqa: E402


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def launches(fn):
    from torch.profiler import ProfilerActivity, profile
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    return sum(e.count for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)


torch.manual_seed(0)
for B in (128, 8192, 65536):
    m = models.DeeperHyperbolicEncoder(512, [256], 128, c=1.0, dropout_rate=0.3).cuda()
    eager = copy.deepcopy(m)
    eager._kernel_train_ok = lambda x: False
    eager._fused_ok = lambda x: False
    x = torch.randn(B, 512, device="cuda") * 0.5
    r = torch.randn(B, 128, device="cuda")

    def train_step(mod):
        mod.train()
        for p in mod.parameters():
            p.grad = None
        (mod(x) * r).sum().backward()

    def infer(mod):
        mod.eval()
        with torch.no_grad():
            mod(x)

    row = {}
    for name, mod in (("kernels", m), ("op-by-op", eager)):
        row[name] = (timed(lambda: infer(mod)), launches(lambda: infer(mod)), timed(lambda: train_step(mod)),
                     launches(lambda: train_step(mod)))
    for name, (ti, li, tt, lt) in row.items():
        print(f"B={B:6d} {name:9s}: inference {ti:8.1f} us ({li:3d} launches)   training step {tt:8.1f} us ({lt:3d} launches)")
