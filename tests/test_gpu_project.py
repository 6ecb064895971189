"""GPU parity: fused projection kernel vs the oracle (reference src/models.py:310,317; 504)."""
import numpy as np
import pytest
import torch

from oracle import head, pmath
from patent_image_retrieval_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _f16_split3(v):
    a = v.to(torch.float16).float()
    b = (v - a).to(torch.float16).float()
    c = (v - a - b).to(torch.float16).float()
    return a, b, c


@pytest.mark.parametrize("d", [128, 512, 768, 2048, 100])
@pytest.mark.parametrize("c", [1.0, 0.5, 2.0])
def test_expmap0_project_matches_oracle(d, c):
    u = synth.gaussian_features(777, d, seed=3)
    u[0] = 0.0                                    # zero row: clamp_min(1e-15) path
    u[1] *= 60.0                                  # far outside: project clip, tanh clamp
    u[2] *= 6.0
    want32 = head.embed_rows(u, c)
    want64 = head.embed_rows(u.double(), c)
    y, op, sq = ops.project_rows(u.cuda(), c, "expmap0", "query", want_sqnorm=True)
    y = y.cpu()
    # tolerance: 1e-6 relative to the row norm (fp32 rounding of an HBM-bound elementwise chain)
    scale = want32.norm(dim=1, keepdim=True).clamp_min(1e-30)
    assert float(((y - want32).abs() / scale).max()) < 1e-6
    # clipped rows sit on the fp32 clip sphere (eps = 4e-3), never outside the ball
    assert float(y.norm(dim=1).max()) <= (1 - 4e-3) / c ** 0.5 * (1 + 1e-6)
    assert torch.equal(y[0], torch.zeros(d))
    # fp64 truth (except clipped rows, whose fp64 clip radius differs by design)
    unclipped = want64.norm(dim=1) < (1 - 5e-3) / c ** 0.5
    assert float(((y.double() - want64)[unclipped].abs().max())) < 1e-6
    torch.testing.assert_close(sq.cpu(), y.pow(2).sum(1), rtol=2e-6, atol=1e-12)
    # operand row (unit-ball coordinates): main columns = fp16(sqrt(c) y), padding zero, extension = 3-way split of
    # c ||y||^2 and ones
    kpad = ops.operand_kpad(d)
    assert op.dtype == torch.float16
    op = op.cpu().float()
    assert op.shape == (777, kpad)
    sc = torch.tensor(c, dtype=torch.float32).sqrt()
    torch.testing.assert_close(op[:, :d], (y * sc).to(torch.float16).float(), rtol=0, atol=0)
    assert float(op[:, d:kpad - 16].abs().max() if kpad - 16 > d else 0.0) == 0.0
    x1, x2, x3 = _f16_split3(torch.tensor(c, dtype=torch.float32) * sq.cpu())
    ext = op[:, kpad - 16:]
    want_ext = torch.stack([x1, x1, x2, x1, x2, x3] + [torch.ones_like(x1)] * 3 + [torch.zeros_like(x1)] * 7, 1)
    torch.testing.assert_close(ext, want_ext, rtol=0, atol=0)


def test_gallery_operand_is_surrogate():
    """<query operand, gallery operand> == c * rb_j * ||x_i - y_j||^2 up to fp16 rounding of the main columns."""
    c, d = 0.6, 512
    u = synth.gaussian_features(64, d, seed=1)
    v = synth.gaussian_features(96, d, seed=0)
    x, q_op, _ = ops.project_rows(u.cuda(), c, "expmap0", "query")
    y, g_op, _ = ops.project_rows(v.cuda(), c, "expmap0", "gallery")
    s = q_op.double() @ g_op.double().t()
    xd, yd = x.double(), y.double()
    want = c * torch.cdist(xd, yd).pow(2) / (1 - c * yd.pow(2).sum(1))[None]
    # error budget: fp16 rounding (2^-11) of both operands on the 2*rb*<x,y> term only
    assert float((s - want).abs().max()) < 2.5e-4
    ext_only = q_op[:, -16:].double() @ g_op[:, -16:].double().t()
    want_ext = c * (xd.pow(2).sum(1)[:, None] + yd.pow(2).sum(1)[None]) / (1 - c * yd.pow(2).sum(1))[None]
    assert float(((ext_only - want_ext).abs() / want_ext).max()) < 2e-6


def test_onball_and_cosine_modes():
    d = 256
    u = synth.gaussian_features(300, d, seed=5, scale=3.0)
    x = head.embed_rows(u, 1.0)
    y, _, _ = ops.project_rows(x.cuda(), 1.0, "onball", "query")
    torch.testing.assert_close(y.cpu(), pmath.project(x, torch.tensor(-1.0)), rtol=1e-6, atol=1e-9)
    raw = synth.gaussian_features(300, d, seed=6, scale=5.0)
    raw[4] = 0.0
    _, opq, sq = ops.project_rows(raw.cuda(), 1.0, "cosine", "query", want_point=False, want_sqnorm=True)
    _, opg, _ = ops.project_rows(raw.cuda(), 1.0, "cosine", "gallery", want_point=False)
    nrm = raw.norm(dim=1, keepdim=True)
    unit = raw / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
    torch.testing.assert_close(opq.cpu().float()[:, :d], unit.to(torch.float16).float(), rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(opg.cpu().float()[:, :d], -opq.cpu().float()[:, :d], rtol=0, atol=0)
    assert float(opq.cpu().float()[:, d:].abs().max()) == 0.0
    assert sq.cpu()[4] == 0.0 and float((sq.cpu()[5:] - 1).abs().max()) == 0.0


def test_argument_errors():
    with pytest.raises(RuntimeError, match="invalid argument"):
        ops.project_rows(torch.zeros(4, 6, device="cuda"))          # d % 4 != 0
    with pytest.raises(RuntimeError, match="invalid argument"):
        ops.project_rows(torch.zeros(4, 64, device="cuda"), c=-1.0)
    y, op, _ = ops.project_rows(torch.zeros(0, 64, device="cuda"))   # empty input is fine
    assert y.shape == (0, 64) and op.shape == (0, 80)


def test_project_rows_to_several_destinations_and_stream_flags():
    """hypret_project_rows_peers stores the operand row into every destination buffer (here: three local buffers,
    at the block offset of a middle rank) exactly as hypret_project_rows does into one; hypret_peer_signal /
    hypret_peer_wait order two streams through counters in device memory."""
    import ctypes
    from patent_image_retrieval_b200 import _lib
    lib = _lib.load()
    n, d, c = 333, 192, 0.7
    u = synth.gaussian_features(n, d, seed=5, scale=1.5).cuda()
    y_ref, op_ref, _ = ops.project_rows(u, c, mode="expmap0", side="query")
    kpad = ops.operand_kpad(d)
    bufs = [torch.full((3 * n, kpad), 7.0, dtype=torch.float16, device="cuda") for _ in range(3)]
    y = torch.zeros(n, d, device="cuda")
    arr = (ctypes.c_void_p * 3)(*[b.data_ptr() + n * kpad * 2 for b in bufs])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    err = torch.zeros(n, device="cuda")
    _lib.check(lib.hypret_project_rows_peers(ctypes.c_void_p(u.data_ptr()), n, d, c, ops.MODE["expmap0"],
                                             ctypes.c_void_p(y.data_ptr()), arr, 3, ctypes.c_void_p(err.data_ptr()),
                                             stream))
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    assert torch.equal(err, ops.project_rows(u, c, mode="expmap0", side="query", want_err=True)[3])
    for b in bufs:
        assert torch.equal(b[n:2 * n], op_ref)
        assert bool((b[:n] == 7.0).all()) and bool((b[2 * n:] == 7.0).all())
    # flags: stream B waits for counters that stream A raises after its copy
    flags = torch.zeros(33, dtype=torch.int32, device="cuda")
    src = torch.arange(1 << 20, device="cuda", dtype=torch.float32)
    dst = torch.zeros_like(src)
    out = torch.zeros_like(src)
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    fl = (ctypes.c_void_p * 2)(flags.data_ptr(), flags.data_ptr() + 4)
    # both kernels run once before anything spins: the first launch of a kernel loads its module (lazy loading),
    # which waits for running kernels -- PeerQueryExchange does the same at construction
    _lib.check(lib.hypret_peer_signal(fl, 2, 1, stream))
    _lib.check(lib.hypret_peer_wait(ctypes.c_void_p(flags.data_ptr()), 2, 1, ctypes.c_void_p(flags.data_ptr() + 128),
                                    stream))
    torch.cuda.synchronize()
    with torch.cuda.stream(b):
        _lib.check(lib.hypret_peer_wait(ctypes.c_void_p(flags.data_ptr()), 2, 3, ctypes.c_void_p(flags.data_ptr() + 128),
                                        ctypes.c_void_p(b.cuda_stream)))
        out.copy_(dst)
    with torch.cuda.stream(a):
        _lib.check(lib.hypret_peer_copy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), src.numel() * 4,
                                        ctypes.c_void_p(a.cuda_stream)))
        _lib.check(lib.hypret_peer_signal(fl, 2, 4, ctypes.c_void_p(a.cuda_stream)))   # counters may run ahead: >=
    torch.cuda.synchronize()
    assert torch.equal(out, src)
    assert flags[:2].tolist() == [4, 4] and int(flags[32]) == 0
