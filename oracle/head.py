"""Oracle for the hyperbolic projection head.  TEST INFRASTRUCTURE.  PARITY UNPINNED.

Follows /root/reference/src/models.py:
  * ``mobius_linear``              291-318
  * ``DeeperHyperbolicEncoder.forward`` 481-505
  * ``encode_figures``             537-548, 803-807

Deviations that restate *intended* behaviour instead of a crash:
  * models.py:306 applies ``F.dropout(weight, dropout)`` with ``dropout``
    undefined (NameError as shipped).  Restated as identity.
  * ``F.dropout`` on activations (486, 500, 804) is identity in eval mode, which
    is the only mode the oracle restates.
"""
from __future__ import annotations

import torch

from . import pmath


def mobius_linear(input, weight, bias=None, hyperbolic_input=True, hyperbolic_bias=True,
                  nonlin=None, k=-1.0):
    weight = weight.to(input.dtype)
    if bias is not None:
        bias = bias.to(input.dtype)
    if hyperbolic_input:
        output = pmath.mobius_matvec(weight, input, k=k)
    else:
        output = torch.nn.functional.linear(input, weight)
        output = pmath.expmap0(output, k=k)
    if bias is not None:
        if not hyperbolic_bias:
            bias = pmath.expmap0(bias, k=k)
        output = pmath.mobius_add(output, bias, k=k)
    if nonlin is not None:
        output = pmath.mobius_fn_apply(nonlin, output, k=k)
    output = pmath.project(output, k=k)
    return output


def encoder_forward(x, w1, b1, w2, b2, k):
    """DeeperHyperbolicEncoder.forward in eval mode (models.py:481-505)."""
    x = mobius_linear(x, w1, b1, hyperbolic_input=False, k=k)      # first_layer, 489
    x = pmath.mobius_fn_apply(torch.tanh, x, k=k)                  # 491
    x = mobius_linear(x, w2, b2, hyperbolic_input=True, k=k)       # final_layer, 501
    x = pmath.project(x, k=k)                                      # 504
    return x


def embed_rows(u, c: float):
    """The synthetic-benchmark projection: x = project(expmap0(u)).

    This is ``mobius_linear(hyperbolic_input=False)`` with identity weight and
    no bias, i.e. models.py:309-310,317 without the ``F.linear``.
    """
    k = torch.tensor(-float(c), dtype=u.dtype)
    return pmath.project(pmath.expmap0(u, k=k), k=k)
