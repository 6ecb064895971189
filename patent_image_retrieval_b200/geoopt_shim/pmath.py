"""``geoopt.manifolds.stereographic.math`` look-alike for the negative-curvature (Poincare
ball) branch -- the subset the reference calls as ``pmath`` (src/models.py:7, src/train.py:18).

Plain differentiable torch on whatever device the tensors live on.  It exists so that the
reference's model / training code imports and runs without geoopt (which is not installable
offline); it is NOT the retrieval hot path -- that lives in libhypret.so (``ops``), and the
matrix-shaped, no-grad ``dist`` calls below are routed there when the tensors are on a GPU.
"""
from __future__ import annotations

import torch

MIN_NORM = 1e-15


def _k(k, like):
    if not torch.is_tensor(k):
        k = torch.tensor(float(k))
    return k.to(dtype=like.dtype, device=like.device)


def sabs(x, eps: float = 1e-15):
    return x.abs().add(eps)


def tanh(x):
    return x.clamp(-15, 15).tanh()


def artanh(x):
    x = x.clamp(-1 + 1e-7, 1 - 1e-7)
    return (torch.log(1 + x).sub(torch.log(1 - x))).mul(0.5)


def tan_k(x, k):
    k_sqrt = sabs(_k(k, x)).sqrt()
    return k_sqrt.reciprocal() * tanh(x * k_sqrt)


def artan_k(x, k):
    k_sqrt = sabs(_k(k, x)).sqrt()
    return k_sqrt.reciprocal() * artanh(x * k_sqrt)


def lambda_x(x, *, k, keepdim=False, dim=-1):
    k = _k(k, x)
    return 2 / (1 + k * x.pow(2).sum(dim=dim, keepdim=keepdim)).clamp_min(MIN_NORM)


def project(x, *, k, dim=-1, eps=-1.0):
    k = _k(k, x)
    if eps < 0:
        eps = 4e-3 if x.dtype == torch.float32 else 1e-5
    maxnorm = (1 - eps) / (sabs(k) ** 0.5)
    norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    return torch.where(norm > maxnorm, x / norm * maxnorm, x)


def expmap0(u, *, k, dim=-1):
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return tan_k(u_norm, k) * (u / u_norm)


def logmap0(y, *, k, dim=-1):
    y_norm = y.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return (y / y_norm) * artan_k(y_norm, k)


def mobius_add(x, y, *, k, dim=-1):
    k = _k(k, x)
    x2 = x.pow(2).sum(dim=dim, keepdim=True)
    y2 = y.pow(2).sum(dim=dim, keepdim=True)
    xy = (x * y).sum(dim=dim, keepdim=True)
    num = (1 - 2 * k * xy - k * y2) * x + (1 + k * x2) * y
    denom = 1 - 2 * k * xy + k ** 2 * x2 * y2
    return num / denom.clamp_min(MIN_NORM)


def mobius_matvec(m, x, *, k, dim=-1):
    if dim != -1:
        x = x.transpose(dim, -1)
    x_norm = x.norm(dim=-1, keepdim=True, p=2).clamp_min(MIN_NORM)
    mx = x @ m.transpose(-1, -2)
    mx_norm = mx.norm(dim=-1, keepdim=True, p=2).clamp_min(MIN_NORM)
    res_c = tan_k(mx_norm / x_norm * artan_k(x_norm, k), k) * (mx / mx_norm)
    cond = (mx == 0).prod(dim=-1, keepdim=True, dtype=torch.bool)
    res = torch.where(cond, torch.zeros(1, dtype=res_c.dtype, device=res_c.device), res_c)
    return res.transpose(dim, -1) if dim != -1 else res


def mobius_fn_apply(fn, x, *args, k, dim=-1, **kwargs):
    return expmap0(fn(logmap0(x, k=k, dim=dim), *args, **kwargs), k=k, dim=dim)


def _matrix_shaped(x, y):
    """[n,1,D] x [1,m,D] (or [1,D] x [m,D]) broadcasts -> (a [n,D], p [m,D]) for the CUDA matrix kernel."""
    if x.dim() == 3 and y.dim() == 3 and x.shape[1] == 1 and y.shape[0] == 1 and x.shape[2] == y.shape[2]:
        return x[:, 0, :], y[0], (x.shape[0], y.shape[1])
    if x.dim() == 2 and y.dim() == 2 and x.shape[0] == 1 and y.shape[0] > 1 and x.shape[1] == y.shape[1]:
        return x, y, (y.shape[0],)
    return None


def dist(x, y, *, k, keepdim=False, dim=-1):
    """2 artan_k(|(-x) (+) y|).  Matrix-shaped, gradient-free fp32 CUDA calls (the reference's
    one-vs-all and B x B uses, src/train.py:3259,1033) run in the exact pairdist kernel instead of
    materialising [n,m,D] Moebius-addition temporaries."""
    if (dim == -1 and not keepdim and x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32
            and not (torch.is_grad_enabled() and (x.requires_grad or y.requires_grad)) and x.shape[-1] % 4 == 0):
        ms = _matrix_shaped(x, y)
        if ms is not None:
            from .. import ops
            a, p, shape = ms
            c = float(-_k(k, x).reshape(-1)[0])
            return ops.pairdist(a, p, c).reshape(shape)
    return 2.0 * artan_k(mobius_add(-x, y, k=k, dim=dim).norm(dim=dim, p=2, keepdim=keepdim), k)


def dist0(x, *, k, keepdim=False, dim=-1):
    return 2.0 * artan_k(x.norm(dim=dim, p=2, keepdim=keepdim), k)


def gyration(a, b, u, *, k, dim=-1):
    k = _k(k, a)
    a2 = a.pow(2).sum(dim=dim, keepdim=True)
    b2 = b.pow(2).sum(dim=dim, keepdim=True)
    ab = (a * b).sum(dim=dim, keepdim=True)
    au = (a * u).sum(dim=dim, keepdim=True)
    bu = (b * u).sum(dim=dim, keepdim=True)
    K2 = k ** 2
    aa = -K2 * au * b2 - k * bu + 2 * K2 * ab * bu
    bb = -K2 * bu * a2 + k * au
    d = 1 - 2 * k * ab + K2 * a2 * b2
    return u + 2 * (aa * a + bb * b) / d.clamp_min(MIN_NORM)


def parallel_transport(x, y, v, *, k, dim=-1):
    return gyration(y, -x, v, k=k, dim=dim) * lambda_x(x, k=k, keepdim=True, dim=dim) / \
        lambda_x(y, k=k, keepdim=True, dim=dim)


def egrad2rgrad(x, grad, *, k, dim=-1):
    return grad / lambda_x(x, k=k, keepdim=True, dim=dim) ** 2


def inner(x, u, v, *, k, keepdim=False, dim=-1):
    return lambda_x(x, k=k, keepdim=True, dim=dim) ** 2 * (u * v).sum(dim=dim, keepdim=keepdim) \
        if keepdim else (lambda_x(x, k=k, keepdim=True, dim=dim) ** 2 * (u * v).sum(dim=dim, keepdim=True)).squeeze(dim)
