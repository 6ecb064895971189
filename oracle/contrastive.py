"""Oracle for the ``train_hyp`` in-batch losses.  TEST INFRASTRUCTURE.  PARITY UNPINNED.

Follows /root/reference/src/train.py:
  * ``train_hyperbolic_contrastive`` hot loop 1832-1844: n x n matrix of
    ``pmath.dist(a_i, p_j)`` -> ``-D/tau`` -> ``F.cross_entropy(., arange(n))``
  * ``hyperbolic_contrastive_loss`` 2291-2336: same matrix, CE over rows and
    columns, averaged
  * ``sample_to_prototype_loss`` 1010-1045: broadcast [B,1,D] x [1,B,D] -> [B,B]
    (line 1033), [B,neg] (1036), hinge

The double Python loop of 1x1 ``pmath.dist`` calls is restated literally in
``dist_matrix_loop`` (small n only) and as one broadcast in ``dist_matrix``
(identical arithmetic per pair, vectorised over pairs).  Autograd gives the
reference backward.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import pmath


def dist_matrix_loop(a: torch.Tensor, p: torch.Tensor, k) -> torch.Tensor:
    n = a.shape[0]
    rows = []
    for i in range(n):
        row = []
        for j in range(p.shape[0]):
            row.append(pmath.dist(a[i:i + 1], p[j:j + 1], k=k).squeeze())
        rows.append(torch.stack(row))
    return torch.stack(rows)


def dist_matrix(a: torch.Tensor, p: torch.Tensor, k, block: int = 256) -> torch.Tensor:
    out = []
    for i0 in range(0, a.shape[0], block):
        out.append(pmath.dist(a[i0:i0 + block, None, :], p[None, :, :], k=k))
    return torch.cat(out, 0)


def contrastive_loss(a, p, k, temperature=0.1, symmetric=False, loop=False):
    d = dist_matrix_loop(a, p, k) if loop else dist_matrix(a, p, k)
    sim = -d / temperature
    labels = torch.arange(a.shape[0])
    if not symmetric:
        return F.cross_entropy(sim, labels)                      # train.py:1842-1844
    return (F.cross_entropy(sim, labels) + F.cross_entropy(sim.t(), labels)) / 2   # 2330-2334


def sample_to_prototype_loss(samples, pos_prototypes, neg_prototypes, num_neg_samples, k,
                             margin=0.1, temperature=0.07):
    b, d = samples.shape
    neg = neg_prototypes.view(b, num_neg_samples, d)
    pos_d = pmath.dist(samples.unsqueeze(1), pos_prototypes.unsqueeze(0), k=k).squeeze(1)   # [B,B] (1033)
    neg_d = pmath.dist(samples.unsqueeze(1), neg, k=k).mean(dim=1)                          # [B] (1036-1038)
    return torch.relu(pos_d.unsqueeze(1) - neg_d + margin).mean()                           # 1040-1043
