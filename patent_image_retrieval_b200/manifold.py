"""Row-local manifold operators of the ``train_hyp`` step on the CUDA path (csrc/manifold.cu; SURVEY 8f-4 and the
per-pair loops of 8a-5), as autograd Functions over the C ABI:

    rowpair_dist(x, y, ia, ib, k)             the per-pair ``pmath.dist`` loops   /root/reference/src/models.py:712-719, 824-829
                                              and the s2p negatives               src/train.py:1036, 1433-1443
    hmi_pair_loss(emb, pairs, k, mode, m)     insideness / disjointedness hinge   src/models.py:550-604, 630-674
    dist0_reg_loss(x, k, lo, hi)              dist0 regulariser                   src/models.py:606-628
    radam_ball_step(...)                      geoopt RiemannianAdam on label_emb  src/train.py:1362

CUDA fp32 tensors only -- like every operator of this package there is no CPU fallback here; the drop-in model methods
(``models.py``) keep an op-by-op torch path for CPU tensors, as the projection head does.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream


def _c_of(k) -> float:
    return float(-torch.as_tensor(k).reshape(-1)[0])


class RowPairDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, ia, ib, c: float):
        _need_cuda(x, y, ia, ib)
        x32, y32 = x.contiguous().float(), y.contiguous().float()
        ia, ib = ia.contiguous().to(torch.int64), ib.contiguous().to(torch.int64)
        if x32.shape[1] != y32.shape[1] or ia.numel() != ib.numel():
            raise ValueError("rowpair_dist: shape mismatch")
        n = ia.numel()
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hypret_rowpair_dist(_ptr(x32), _ptr(y32), _ptr(ia), _ptr(ib), n, x32.shape[1],
                                                       float(c), _ptr(out), _stream()))
        ctx.save_for_backward(x32, y32, ia, ib)
        ctx.c, ctx.dtypes = c, (x.dtype, y.dtype)
        return out.to(torch.promote_types(x.dtype, y.dtype))

    @staticmethod
    def backward(ctx, grad_out):
        x, y, ia, ib = ctx.saved_tensors
        gx = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.zeros_like(y) if ctx.needs_input_grad[1] else None
        if gx is None and gy is None:
            return None, None, None, None, None
        g = grad_out.contiguous().float()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hypret_rowpair_dist_bwd(_ptr(x), _ptr(y), _ptr(ia), _ptr(ib), ia.numel(), x.shape[1],
                                                           float(ctx.c), _ptr(g), _ptr(gx), _ptr(gy), _stream()))
        return (gx.to(ctx.dtypes[0]) if gx is not None else None, gy.to(ctx.dtypes[1]) if gy is not None else None,
                None, None, None)


def rowpair_dist(x: torch.Tensor, y: torch.Tensor, ia: torch.Tensor, ib: torch.Tensor, k) -> torch.Tensor:
    """``d[t] = pmath.dist(x[ia[t]], y[ib[t]], k)`` for index pairs, one kernel forward and one backward."""
    return RowPairDistance.apply(x, y, ia, ib, _c_of(k))


HMI_MODES = {"insideness": 0, "disjointedness": 1}


def _proj_eps(dtype) -> float:
    """geoopt's ``project`` margin depends on the parameter's dtype (4e-3 for fp32, 1e-5 for fp64; the reference's label
    embeddings are fp64 because src/models.py:248-249 flips the default dtype)."""
    return 1e-5 if dtype == torch.float64 else 4e-3


def hmi_values(emb: torch.Tensor, pairs: torch.Tensor, k, mode: str) -> torch.Tensor:
    """``_hmi_insideness`` / ``_hmi_disjointedness`` of the label pairs (no gradient): ``[n_pairs, 1]``."""
    _need_cuda(emb, pairs)
    e32 = emb.detach().contiguous().float()
    pairs = pairs.contiguous().to(torch.int64).view(-1, 2)
    out = torch.empty(pairs.shape[0], dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.load().hypret_hmi_pairs(_ptr(e32), _ptr(pairs), pairs.shape[0], e32.shape[1], _c_of(k),
                                                HMI_MODES[mode], 0.0, _proj_eps(emb.dtype), _ptr(out), None, None, None,
                                                _stream()))
    return out[:, None]


class HmiPairLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, pairs, c: float, mode: int, margin: float):
        _need_cuda(emb, pairs)
        e32 = emb.contiguous().float()
        pairs = pairs.contiguous().to(torch.int64).view(-1, 2)
        if pairs.numel() and (int(pairs.min()) < 0 or int(pairs.max()) >= e32.shape[0]):
            raise IndexError("Invalid index detected in label pairs")           # as src/models.py:563-579
        n = pairs.shape[0]
        total = torch.zeros(1, dtype=torch.float64, device=emb.device)
        with torch.cuda.device(emb.device):
            _lib.check(_lib.load().hypret_hmi_pairs(_ptr(e32), _ptr(pairs), n, e32.shape[1], float(c), int(mode),
                                                    float(margin), _proj_eps(emb.dtype), None, _ptr(total), None, None,
                                                    _stream()))
        ctx.save_for_backward(e32, pairs)
        ctx.c, ctx.mode, ctx.margin, ctx.dtype = c, mode, margin, emb.dtype
        return (total[0] / max(n, 1)).to(emb.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        e32, pairs = ctx.saved_tensors
        grad = torch.zeros_like(e32)
        gs = grad_loss.reshape(1).contiguous().float()
        with torch.cuda.device(e32.device):
            _lib.check(_lib.load().hypret_hmi_pairs(_ptr(e32), _ptr(pairs), pairs.shape[0], e32.shape[1], float(ctx.c),
                                                    int(ctx.mode), float(ctx.margin), _proj_eps(ctx.dtype), None, None,
                                                    _ptr(gs), _ptr(grad), _stream()))
        return grad.to(ctx.dtype), None, None, None, None


def hmi_pair_loss(emb: torch.Tensor, pairs: torch.Tensor, k, mode: str, margin: float) -> torch.Tensor:
    """``relu(margin - hmi(emb[pairs[:,0]], emb[pairs[:,1]])).mean()`` with the analytic gradient."""
    return HmiPairLoss.apply(emb, pairs, _c_of(k), HMI_MODES[mode], float(margin))


class Dist0Reg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c: float, lo: float, hi: float):
        _need_cuda(x)
        x32 = x.contiguous().float()
        n, d = x32.shape
        total = torch.zeros(1, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hypret_dist0_reg(_ptr(x32), n, d, float(c), float(lo), float(hi), _ptr(total), None,
                                                    None, _stream()))
        ctx.save_for_backward(x32)
        ctx.c, ctx.lo, ctx.hi, ctx.dtype = c, lo, hi, x.dtype
        return (total[0] / max(n, 1)).to(x.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        x32, = ctx.saved_tensors
        grad = torch.zeros_like(x32)
        gs = grad_loss.reshape(1).contiguous().float()
        with torch.cuda.device(x32.device):
            _lib.check(_lib.load().hypret_dist0_reg(_ptr(x32), x32.shape[0], x32.shape[1], float(ctx.c), float(ctx.lo),
                                                    float(ctx.hi), None, _ptr(gs), _ptr(grad), _stream()))
        return grad.to(ctx.dtype), None, None, None


def dist0_reg_loss(x: torch.Tensor, k, lo: Optional[float], hi: float) -> torch.Tensor:
    """``(relu(lo - dist0(x)) + relu(dist0(x) - hi)).mean()`` (``lo=None``: upper hinge only)."""
    return Dist0Reg.apply(x, _c_of(k), -1.0 if lo is None else float(lo), float(hi))


def radam_ball_step(point: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, c: float,
                    lr: float, betas, eps: float, weight_decay: float, step: int) -> None:
    """One fused RiemannianAdam step on a ``[n,d]`` fp32 Poincare-ball parameter, in place (``hypret_radam_ball_step``)."""
    _need_cuda(point, grad, exp_avg, exp_avg_sq)
    for t in (point, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != point.shape or t.dim() != 2:
            raise ValueError("radam_ball_step: contiguous [n,d] float32 tensors of one shape")
    with torch.cuda.device(point.device):
        _lib.check(_lib.load().hypret_radam_ball_step(_ptr(point), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq),
                                                      point.shape[0], point.shape[1], float(c), float(lr),
                                                      float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                                      int(step), _stream()))
