"""``geoopt.optim.RiemannianAdam`` restated (reference use: src/train.py:1362,2177,2643).
Euclidean parameters get plain Adam; ``ManifoldParameter`` s get the Riemannian update
(egrad2rgrad, moments in the tangent space, retraction + parallel transport of the first
moment).  A ``[n,d]`` fp32 CUDA parameter on the Poincare ball (``label_emb``) takes the whole update as ONE fused
kernel (``hypret_radam_ball_step``, csrc/manifold.cu); everything else runs op by op."""
from __future__ import annotations

import torch


class _Euclidean:
    @staticmethod
    def egrad2rgrad(x, u):
        return u

    @staticmethod
    def component_inner(x, u, v=None):
        return u * (u if v is None else v)

    @staticmethod
    def retr_transp(x, u, v):
        return x + u, v


class RiemannianAdam(torch.optim.Adam):
    def __init__(self, *args, stabilize=None, **kwargs):
        self._stabilize = stabilize
        super().__init__(*args, **kwargs)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            wd, eps, lr, amsgrad = group["weight_decay"], group["eps"], group["lr"], group["amsgrad"]
            for point in group["params"]:
                grad = point.grad
                if grad is None:
                    continue
                manifold = getattr(point, "manifold", None) or _Euclidean
                state = self.state[point]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(point)
                    state["exp_avg_sq"] = torch.zeros_like(point)
                    if amsgrad:
                        state["max_exp_avg_sq"] = torch.zeros_like(point)
                state["step"] += 1
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                if (not amsgrad and point.is_cuda and point.dtype == torch.float32 and point.dim() == 2
                        and point.is_contiguous() and hasattr(manifold, "isp_c") and grad.dtype == torch.float32):
                    from ..manifold import radam_ball_step
                    radam_ball_step(point.data, grad.contiguous(), exp_avg, exp_avg_sq, float(manifold.c), lr,
                                    (b1, b2), eps, wd, state["step"])
                    continue
                grad = grad.add(point, alpha=wd)
                grad = manifold.egrad2rgrad(point, grad)
                exp_avg.mul_(b1).add_(grad, alpha=1 - b1)
                exp_avg_sq.mul_(b2).add_(manifold.component_inner(point, grad), alpha=1 - b2)
                bc1 = 1 - b1 ** state["step"]
                bc2 = 1 - b2 ** state["step"]
                if amsgrad:
                    torch.max(state["max_exp_avg_sq"], exp_avg_sq, out=state["max_exp_avg_sq"])
                    denom = state["max_exp_avg_sq"].div(bc2).sqrt_()
                else:
                    denom = exp_avg_sq.div(bc2).sqrt_()
                direction = exp_avg.div(bc1) / denom.add_(eps)
                new_point, exp_avg_new = manifold.retr_transp(point, -lr * direction, exp_avg)
                point.copy_(new_point)
                exp_avg.copy_(exp_avg_new)
        return loss
