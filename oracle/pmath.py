"""Restatement of geoopt's stereographic-model math for negative curvature.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: geoopt is not
available offline; formulas follow its published
``geoopt/manifolds/stereographic/math.py`` (k = -c < 0 branch only), which is
what the reference calls as ``pmath`` (src/models.py:7, src/train.py:18):

    pmath.dist          src/train.py:1837,2315,3259  src/models.py:718,828
    pmath.dist0         src/models.py:586-587
    pmath.expmap0       src/models.py:263,310,313
    pmath.project       src/models.py:317,504
    pmath.mobius_add    src/models.py:314
    pmath.mobius_matvec src/models.py:307
    pmath.mobius_fn_apply src/models.py:316,491

Everything is dtype-generic torch on CPU: call with fp32 tensors for "the
reference's fp32 path", with fp64 tensors for the conditioning truth.
``k`` is a tensor holding ``-c`` exactly like the reference's ``self.k``
(src/models.py:462,519,793).
"""
from __future__ import annotations

import torch

MIN_NORM = 1e-15


def _k(k, like: torch.Tensor) -> torch.Tensor:
    if not torch.is_tensor(k):
        k = torch.tensor(float(k))
    return k.to(dtype=like.dtype, device=like.device)


def sabs(x: torch.Tensor, eps: float = 1e-15) -> torch.Tensor:
    # geoopt.utils.sabs
    return x.abs().add(eps)


def tanh(x: torch.Tensor) -> torch.Tensor:
    # geoopt.utils.tanh: clamp to +-15 before tanh
    return x.clamp(-15, 15).tanh()


def artanh(x: torch.Tensor) -> torch.Tensor:
    # geoopt.utils.artanh
    x = x.clamp(-1 + 1e-7, 1 - 1e-7)
    return (torch.log(1 + x).sub(torch.log(1 - x))).mul(0.5)


def tan_k(x: torch.Tensor, k) -> torch.Tensor:
    k = _k(k, x)
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * tanh(x * k_sqrt)


def artan_k(x: torch.Tensor, k) -> torch.Tensor:
    k = _k(k, x)
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * artanh(x * k_sqrt)


def project(x: torch.Tensor, k, dim: int = -1, eps: float = -1.0) -> torch.Tensor:
    k = _k(k, x)
    if eps < 0:
        eps = 4e-3 if x.dtype == torch.float32 else 1e-5
    maxnorm = (1 - eps) / (sabs(k) ** 0.5)
    norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    cond = norm > maxnorm
    projected = x / norm * maxnorm
    return torch.where(cond, projected, x)


def expmap0(u: torch.Tensor, k, dim: int = -1) -> torch.Tensor:
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return tan_k(u_norm, k) * (u / u_norm)


def logmap0(y: torch.Tensor, k, dim: int = -1) -> torch.Tensor:
    y_norm = y.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return (y / y_norm) * artan_k(y_norm, k)


def mobius_add(x: torch.Tensor, y: torch.Tensor, k, dim: int = -1) -> torch.Tensor:
    k = _k(k, x)
    x2 = x.pow(2).sum(dim=dim, keepdim=True)
    y2 = y.pow(2).sum(dim=dim, keepdim=True)
    xy = (x * y).sum(dim=dim, keepdim=True)
    num = (1 - 2 * k * xy - k * y2) * x + (1 + k * x2) * y
    denom = 1 - 2 * k * xy + k ** 2 * x2 * y2
    return num / denom.clamp_min(MIN_NORM)


def mobius_matvec(m: torch.Tensor, x: torch.Tensor, k, dim: int = -1) -> torch.Tensor:
    if dim != -1:
        raise NotImplementedError("oracle restates the dim=-1 call sites only")
    x_norm = x.norm(dim=-1, keepdim=True, p=2).clamp_min(MIN_NORM)
    mx = x @ m.transpose(-1, -2)
    mx_norm = mx.norm(dim=-1, keepdim=True, p=2).clamp_min(MIN_NORM)
    res_c = tan_k(mx_norm / x_norm * artan_k(x_norm, k), k) * (mx / mx_norm)
    cond = (mx == 0).prod(dim=-1, keepdim=True, dtype=torch.bool)
    res_0 = torch.zeros(1, dtype=res_c.dtype, device=res_c.device)
    return torch.where(cond, res_0, res_c)


def mobius_fn_apply(fn, x: torch.Tensor, *args, k, dim: int = -1, **kwargs) -> torch.Tensor:
    ex = logmap0(x, k, dim=dim)
    ex = fn(ex, *args, **kwargs)
    return expmap0(ex, k, dim=dim)


def dist(x: torch.Tensor, y: torch.Tensor, k, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    """geoopt form: 2 * artan_k(|| (-x) (+) y ||)."""
    return 2.0 * artan_k(mobius_add(-x, y, k, dim=dim).norm(dim=dim, p=2, keepdim=keepdim), k)


def dist0(x: torch.Tensor, k, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    return 2.0 * artan_k(x.norm(dim=dim, p=2, keepdim=keepdim), k)


def dist_arcosh(x: torch.Tensor, y: torch.Tensor, k, dim: int = -1) -> torch.Tensor:
    """Closed form named by BASELINE.json north_star (analytically == ``dist``):

        d = arccosh(1 + 2c||x-y||^2 / ((1-c||x||^2)(1-c||y||^2))) / sqrt(c)

    evaluated as log1p(t + sqrt(t(t+2))) so that small distances keep their
    relative accuracy.  This is the form the CUDA rerank kernel evaluates
    (in fp64); it is kept here so the discrepancy between the two forms is
    measured, not assumed.
    """
    k = _k(k, x)
    c = -k
    s = (x - y).pow(2).sum(dim=dim)
    a = 1 - c * x.pow(2).sum(dim=dim)
    b = 1 - c * y.pow(2).sum(dim=dim)
    t = 2 * c * s / (a * b)
    return torch.log1p(t + torch.sqrt(t * (t + 2))) / c.sqrt()


def check_point_on_manifold(x: torch.Tensor, k, atol: float = 1e-5, dim: int = -1) -> bool:
    """geoopt Stereographic._check_point_on_manifold: ||project(x) - x|| small."""
    px = project(x, k, dim=dim)
    return bool(torch.allclose(x, px, atol=atol, rtol=1e-5))
