"""A geoopt-shaped shim: just enough of ``geoopt`` for the reference's hot-path callers
(``PoincareBall``, ``ManifoldParameter``, ``optim.RiemannianAdam``,
``manifolds.stereographic.math``) -- see SURVEY.md 8b.  ``install()`` registers it under the
name ``geoopt`` when the real package is absent, so ``import geoopt as gt`` and
``import geoopt.manifolds.stereographic.math as pmath`` (src/models.py:5-7, src/train.py:15-18)
resolve unmodified.
"""
from __future__ import annotations

import sys
import types

import torch

from . import pmath  # noqa: F401


class PoincareBall(torch.nn.Module):
    """geoopt.PoincareBall(c): curvature holder + the handful of methods the reference uses
    (src/models.py:258,360,381,461,520,547,612-634,794,806).

    As in geoopt it is an ``nn.Module`` whose curvature lives in the parameter ``isp_c = log(exp(c) - 1)``
    (``c = softplus(isp_c)``, ``requires_grad = learnable``), so a model that owns a ball has the reference's
    state-dict keys ``ball.isp_c``, ``encoder.ball.isp_c``, ``encoder.{first,final}_layer.ball.isp_c`` and
    checkpoints written by the real reference load strictly.  Checkpoints written without those keys (round 1 of
    this repo kept the ball outside the module tree) load too: a pre-hook fills a missing ``isp_c`` with the
    constructed value."""

    name = "Poincare ball"
    ndim = 1

    def __init__(self, c=1.0, learnable=False):
        super().__init__()
        c = torch.as_tensor(c)
        if not torch.is_floating_point(c):
            c = c.to(torch.get_default_dtype())
        with torch.no_grad():
            isp = c.detach().clone().exp_().sub_(1).log_()
        self.isp_c = torch.nn.Parameter(isp, requires_grad=bool(learnable))
        self._register_load_state_dict_pre_hook(self._fill_missing_curvature)

    def _fill_missing_curvature(self, state_dict, prefix, *unused):
        state_dict.setdefault(prefix + "isp_c", self.isp_c.detach())

    @property
    def c(self):
        return torch.nn.functional.softplus(self.isp_c)

    @property
    def k(self):
        return -self.c

    def projx(self, x, *, dim=-1):
        return pmath.project(x, k=self.k, dim=dim)

    def expmap0(self, u, *, dim=-1, project=True):
        res = pmath.expmap0(u, k=self.k, dim=dim)
        return pmath.project(res, k=self.k, dim=dim) if project else res

    def logmap0(self, x, *, dim=-1):
        return pmath.logmap0(x, k=self.k, dim=dim)

    def dist(self, x, y, *, keepdim=False, dim=-1):
        return pmath.dist(x, y, k=self.k, keepdim=keepdim, dim=dim)

    def dist0(self, x, *, dim=-1, keepdim=False):
        return pmath.dist0(x, k=self.k, dim=dim, keepdim=keepdim)

    def mobius_add(self, x, y, *, dim=-1, project=True):
        res = pmath.mobius_add(x, y, k=self.k, dim=dim)
        return pmath.project(res, k=self.k, dim=dim) if project else res

    def egrad2rgrad(self, x, u, *, dim=-1):
        return pmath.egrad2rgrad(x, u, k=self.k, dim=dim)

    def retr(self, x, u, *, dim=-1):
        return pmath.project(x + u, k=self.k, dim=dim)

    def transp(self, x, y, v, *, dim=-1):
        return pmath.parallel_transport(x, y, v, k=self.k, dim=dim)

    def retr_transp(self, x, u, v, *, dim=-1):
        y = self.retr(x, u, dim=dim)
        return y, self.transp(x, y, v, dim=dim)

    def component_inner(self, x, u, v=None):
        # geoopt Manifold.component_inner = inner(x, u, v, keepdim=True): one value per point, broadcast by the caller
        v = u if v is None else v
        return pmath.lambda_x(x, k=self.k, keepdim=True) ** 2 * (u * v).sum(dim=-1, keepdim=True)

    def check_point_on_manifold(self, x, *, explain=False, atol=1e-5, rtol=1e-5):
        px = pmath.project(x, k=self.k)
        ok = bool(torch.allclose(x, px, atol=atol, rtol=rtol))
        return (ok, None if ok else "'x' norm lies out of the bounds [-1/sqrt(c)+eps, 1/sqrt(c)-eps]") if explain else ok

    def assert_check_point_on_manifold(self, x, *, atol=1e-5, rtol=1e-5):
        ok, reason = self.check_point_on_manifold(x, explain=True, atol=atol, rtol=rtol)
        if not ok:
            raise ValueError(f"`x` seems to be a tensor not lying on {self.name} manifold.\\nerror: {reason}")


class ManifoldParameter(torch.nn.Parameter):
    """geoopt.ManifoldParameter(data, manifold=...): an nn.Parameter that remembers its manifold."""

    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.empty(0)
        inst = torch.Tensor._make_subclass(cls, data.detach() if isinstance(data, torch.Tensor) else data,
                                           requires_grad)
        inst.manifold = manifold
        return inst

    def __repr__(self):
        return f"Parameter on {getattr(self.manifold, 'name', '?')} manifold containing:\n" + torch.Tensor.__repr__(self)

    def __deepcopy__(self, memo):
        out = type(self)(self.data.clone(memory_format=torch.preserve_format), manifold=self.manifold,
                         requires_grad=self.requires_grad)
        memo[id(self)] = out
        return out

    def __reduce_ex__(self, proto):
        return _rebuild_manifold_parameter, (self.data, self.manifold, self.requires_grad)


def _rebuild_manifold_parameter(data, manifold, requires_grad):
    return ManifoldParameter(data, manifold=manifold, requires_grad=requires_grad)


from . import optim  # noqa: E402,F401


def install(force: bool = False) -> bool:
    """Expose this shim as ``geoopt`` if the real package cannot be imported."""
    if not force:
        try:
            import geoopt  # noqa: F401
            return False
        except Exception:
            pass
    me = sys.modules[__name__]
    manifolds = types.ModuleType("geoopt.manifolds")
    stereo = types.ModuleType("geoopt.manifolds.stereographic")
    stereo.math = pmath
    manifolds.stereographic = stereo
    manifolds.PoincareBall = PoincareBall
    me.manifolds = manifolds
    sys.modules["geoopt"] = me
    sys.modules["geoopt.manifolds"] = manifolds
    sys.modules["geoopt.manifolds.stereographic"] = stereo
    sys.modules["geoopt.manifolds.stereographic.math"] = pmath
    sys.modules["geoopt.optim"] = optim
    return True
