"""Drop-in ``train_hyp`` losses of the reference on the CUDA path.

    hyperbolic_contrastive_loss(anchor, positive, k, temperature)     /root/reference/src/train.py:2291-2336
    in-batch loss of train_hyperbolic_contrastive                     src/train.py:1832-1844
    sample_to_prototype_loss(...)                                     src/train.py:1010-1045
    train_hyperbolic_contrastive(...)                                 src/train.py:1792-1910

The reference builds the n x n matrix with an O(n^2) Python double loop of 1x1 ``pmath.dist``
calls (~40 kernel launches each) and differentiates through all of them.  Here the in-batch loss is
one autograd Function over the flash kernels (``hypret_flash_lse`` / ``hypret_flash_grad``: no
``[n,n]`` array, closed-form backward on the tensor cores) for 16 <= D <= 128, D % 16 == 0,
and over the matrix kernels (``hypret_pairdist_ce_fwd/bwd`` + split products) for other widths;
``pairwise_dist`` is the explicit matrix (``hypret_pairdist`` + ``hypret_pairdist_bwd``).  CUDA tensors only.
"""
from __future__ import annotations

import os
import random
from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from .geoopt_shim import pmath


# Tensor-core training path (csrc/gramdist.cu forward, three-plane W + bf16 GEMMs backward).  The bf16 x 3 splits carry
# fp32 values exactly, but the tensor core's fp32 accumulator truncates (<= K/16 ulps of the running sum, one-sided;
# DESIGN 4.6).  HYPRET_TRAIN_FP32=1 keeps every product on the FP32 pipe (FFMA distance tiles, SGEMM backward).
TENSOR_CORES = os.environ.get("HYPRET_TRAIN_FP32", "0") != "1"


def _c_of(k) -> float:
    return float(-torch.as_tensor(k).reshape(-1)[0])


class PairwiseDistance(torch.autograd.Function):
    """D[i,j] = dist(a_i, p_j) on the Poincare ball, exact fp32, with the analytic backward."""

    @staticmethod
    def forward(ctx, a, p, c: float):
        a32, p32 = a.contiguous().float(), p.contiguous().float()
        d = ops.pairdist(a32, p32, c)
        ctx.save_for_backward(a32, p32, d)
        ctx.c = c
        ctx.in_dtypes = (a.dtype, p.dtype)
        return d

    @staticmethod
    def backward(ctx, grad_out):
        a, p, d = ctx.saved_tensors
        split = TENSOR_CORES and d.numel() >= ops.SPLIT_MIN_PAIRS
        w, rs, cs = ops.pairdist_bwd(grad_out, d, ops.row_sqnorm(a), ops.row_sqnorm(p), ctx.c, split=split)
        wp, wta = ops.split_products(w, a, p) if split else (w @ p, w.t() @ a)     # plain library GEMMs
        da = a * rs[:, None] - wp
        dp = p * cs[:, None] - wta
        return da.to(ctx.in_dtypes[0]), dp.to(ctx.in_dtypes[1]), None


def pairwise_dist(a, p, k):
    """[n,D] x [m,D] -> [n,m] Poincare distances (differentiable)."""
    return PairwiseDistance.apply(a, p, _c_of(k))


def _diag_distance(a32, p32, c: float, offset: int = 0):
    """Exact d(a_i, p_{i+offset}) for every local row (``hypret_rowpair_dist``, fp64 inside): the targets' logits."""
    from .manifold import RowPairDistance
    ia = torch.arange(a32.shape[0], device=a32.device)
    return RowPairDistance.apply(a32.detach(), p32.detach(), ia, ia + offset, c)


class InBatchInfoNCE(torch.autograd.Function):
    """CE over the rows (``symmetric=False``) or rows and columns (``True``) of ``-D / tau`` with the diagonal as
    targets, D the n x n Poincare distance matrix -- forward and backward in libhypret.so.

    Flash path (16 <= D <= 128, D % 16 == 0; csrc/flash.cu): D is NEVER written to memory.  Forward = row (and column)
    log-sum-exps from tcgen05 Gram tiles; backward = the tiles recomputed, the weights W formed in registers, written
    to shared memory as bf16 planes and multiplied by the other operand with a second tcgen05.mma (one launch per
    operand, roles swapped).  Other shapes: the matrix kernels (``hypret_pairdist_ce_fwd/bwd`` + split products)."""

    @staticmethod
    def forward(ctx, a, p, c: float, temperature: float, symmetric: bool):
        a32, p32 = a.contiguous().float(), p.contiguous().float()
        if a32.shape != p32.shape:
            raise ValueError("anchors and positives must have the same shape")
        inv_tau = 1.0 / float(temperature)
        n, d = a32.shape
        ctx.flash = TENSOR_CORES and ops.flash_ok(n, n, d)
        if ctx.flash:
            need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
            ao = ops.FlashOperands(a32, want_col=symmetric or need_grad, want_t=need_grad)
            po = ops.FlashOperands(p32, want_row=symmetric or need_grad, want_t=need_grad)
            row_lse = ops.flash_lse(ao, po, c, inv_tau)
            diag_sim = -_diag_distance(a32, p32, c) * inv_tau
            loss = (row_lse - diag_sim).mean()
            col_lse = None
            if symmetric:
                col_lse = ops.flash_lse(po, ao, c, inv_tau)
                loss = (loss + (col_lse - diag_sim).mean()) / 2
            ctx.ops = (ao, po)
            ctx.save_for_backward(row_lse, col_lse if symmetric else row_lse)
            ctx.c, ctx.inv_tau, ctx.symmetric = c, inv_tau, symmetric
            ctx.in_dtypes = (a.dtype, p.dtype)
            return loss
        d, row_lse, col_lse = ops.pairdist_ce_fwd(a32, p32, c, inv_tau, want_cols=symmetric,
                                                  tensor_cores=None if TENSOR_CORES else False)
        diag_sim = -torch.diagonal(d) * inv_tau
        loss = (row_lse - diag_sim).mean()
        if symmetric:
            loss = (loss + (col_lse - diag_sim).mean()) / 2
        ctx.save_for_backward(a32, p32, d, row_lse, col_lse if symmetric else row_lse)
        ctx.c, ctx.inv_tau, ctx.symmetric = c, inv_tau, symmetric
        ctx.in_dtypes = (a.dtype, p.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        wr, wc = (0.5, 0.5) if ctx.symmetric else (1.0, 0.0)
        if ctx.flash:
            row_lse, col_lse = ctx.saved_tensors
            ao, po = ctx.ops
            cl = col_lse if ctx.symmetric else None
            da = ops.flash_grad(ao, po, ctx.c, ctx.inv_tau, row_lse, cl, wr, wc, grad_scale=grad_loss)
            dp = ops.flash_grad(po, ao, ctx.c, ctx.inv_tau, cl, row_lse, wc, wr, grad_scale=grad_loss)
            return da.to(ctx.in_dtypes[0]), dp.to(ctx.in_dtypes[1]), None, None, None
        a, p, d, row_lse, col_lse = ctx.saved_tensors
        split = TENSOR_CORES and d.numel() >= ops.SPLIT_MIN_PAIRS
        w, rs, cs = ops.pairdist_ce_bwd(d, ops.row_sqnorm(a), ops.row_sqnorm(p), ctx.c, row_lse,
                                        col_lse if ctx.symmetric else None, ctx.inv_tau, wr, wc, grad_scale=grad_loss,
                                        split=split)
        wp, wta = ops.split_products(w, a, p) if split else (w @ p, w.t() @ a)     # plain library GEMMs
        da = a * rs[:, None] - wp
        dp = p * cs[:, None] - wta
        return da.to(ctx.in_dtypes[0]), dp.to(ctx.in_dtypes[1]), None, None, None


class ShardedInBatchInfoNCE(torch.autograd.Function):
    """``InBatchInfoNCE`` over a batch whose rows are sharded across the ranks of ``group`` (sharded negatives;
    SURVEY 8e "next"): rank r holds anchors / positives ``[nl,D]`` = rows ``[r*nl, (r+1)*nl)`` of the global batch of
    ``n = W*nl`` pairs.  Forward: all_gather the positives, this rank's ``[nl,n]`` row block of the distance matrix
    (same kernels; the target of local row i is column ``r*nl + i``), row log-sum-exps local, column log-sum-exps
    combined across ranks (all_gather of the per-rank partials), loss all-reduced: every rank returns the GLOBAL
    mean.  Backward: dA is complete locally; the partial dP of all n positives is reduce-scattered to the owners.
    Gradients are those of the global loss with respect to the local rows (no further averaging)."""

    @staticmethod
    def forward(ctx, a, p, c: float, temperature: float, symmetric: bool, group):
        import torch.distributed as dist
        a32, p32 = a.contiguous().float(), p.contiguous().float()
        if a32.shape != p32.shape:
            raise ValueError("anchors and positives must have the same shape")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nl, d = a32.shape
        n = world * nl
        p_all = torch.empty(n, d, dtype=torch.float32, device=a32.device)
        dist.all_gather_into_tensor(p_all, p32, group=group)
        inv_tau = 1.0 / float(temperature)
        off = rank * nl
        ctx.flash = TENSOR_CORES and ops.flash_ok(nl, n, d)
        dm = ao = po = None
        if ctx.flash:                                                           # no [nl, n] array (csrc/flash.cu)
            ao, po = ops.FlashOperands(a32), ops.FlashOperands(p_all)
            row_lse = ops.flash_lse(ao, po, c, inv_tau)
            col_part = ops.flash_lse(po, ao, c, inv_tau) if symmetric else None  # over this rank's rows only
            diag_sim = -_diag_distance(a32, p_all, c, off) * inv_tau            # sim[i, off + i]
        else:
            dm, row_lse, col_part = ops.pairdist_ce_fwd(a32, p_all, c, inv_tau, want_cols=symmetric,
                                                        tensor_cores=None if TENSOR_CORES else False)
            diag_sim = -torch.diagonal(dm, offset=off) * inv_tau
        local = (row_lse - diag_sim).sum()
        col_lse = None
        if symmetric:
            parts = torch.empty(world, n, dtype=torch.float32, device=a32.device)
            dist.all_gather_into_tensor(parts, col_part, group=group)
            col_lse = ops.lse_combine(parts)                                    # over the row blocks of all ranks
            local = (local + (col_lse[off:off + nl] - diag_sim).sum()) / 2
        total = local.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        ctx.ops = (ao, po)
        ctx.save_for_backward(a32, p_all, dm if dm is not None else row_lse, row_lse,
                              col_lse if symmetric else row_lse)
        ctx.c, ctx.inv_tau, ctx.symmetric, ctx.group = c, inv_tau, symmetric, group
        ctx.off, ctx.n, ctx.nl = off, n, nl
        ctx.in_dtypes = (a.dtype, p.dtype)
        return total / n

    @staticmethod
    def backward(ctx, grad_loss):
        import torch.distributed as dist
        a, p_all, dm, row_lse, col_lse = ctx.saved_tensors
        wr, wc = (0.5, 0.5) if ctx.symmetric else (1.0, 0.0)
        if ctx.flash:
            ao, po = ctx.ops
            cl = col_lse if ctx.symmetric else None
            da = ops.flash_grad(ao, po, ctx.c, ctx.inv_tau, row_lse, cl, wr, wc, grad_scale=grad_loss,
                                diag_offset=ctx.off, n_total=ctx.n)
            dp_all = ops.flash_grad(po, ao, ctx.c, ctx.inv_tau, cl, row_lse, wc, wr, grad_scale=grad_loss,
                                    diag_offset=-ctx.off, n_total=ctx.n)
            dp = torch.empty(ctx.nl, dp_all.shape[1], dtype=torch.float32, device=dp_all.device)
            dist.reduce_scatter_tensor(dp, dp_all, op=dist.ReduceOp.SUM, group=ctx.group)
            return da.to(ctx.in_dtypes[0]), dp.to(ctx.in_dtypes[1]), None, None, None, None
        split = TENSOR_CORES and dm.numel() >= ops.SPLIT_MIN_PAIRS
        w, rs, cs = ops.pairdist_ce_bwd(dm, ops.row_sqnorm(a), ops.row_sqnorm(p_all), ctx.c, row_lse,
                                        col_lse if ctx.symmetric else None, ctx.inv_tau, wr, wc, grad_scale=grad_loss,
                                        split=split, diag_offset=ctx.off, n_total=ctx.n)
        wp, wta = ops.split_products(w, a, p_all) if split else (w @ p_all, w.t() @ a)
        da = a * rs[:, None] - wp
        dp_all = (p_all * cs[:, None] - wta).contiguous()                       # this rank's rows' share, all n positives
        dp = torch.empty(ctx.nl, dp_all.shape[1], dtype=torch.float32, device=dp_all.device)
        dist.reduce_scatter_tensor(dp, dp_all, op=dist.ReduceOp.SUM, group=ctx.group)
        return da.to(ctx.in_dtypes[0]), dp.to(ctx.in_dtypes[1]), None, None, None, None


def sharded_in_batch_contrastive_loss(anchors, positives, k, temperature=0.1, symmetric=False, group=None):
    """The in-batch loss of src/train.py:1832-1844 (``symmetric``: 2304-2334) over a batch sharded across ranks."""
    return ShardedInBatchInfoNCE.apply(anchors, positives, _c_of(k), float(temperature), bool(symmetric), group)


def in_batch_contrastive_loss(anchors, positives, k, temperature=0.1):
    """The loss inside train_hyperbolic_contrastive (src/train.py:1832-1844): CE over rows."""
    return InBatchInfoNCE.apply(anchors, positives, _c_of(k), float(temperature), False)


def hyperbolic_contrastive_loss(anchor_embeddings, positive_embeddings, k, temperature=0.07):
    """Symmetric InfoNCE in hyperbolic space (src/train.py:2291-2336)."""
    return InBatchInfoNCE.apply(anchor_embeddings, positive_embeddings, _c_of(k), float(temperature), True)


def sample_to_prototype_loss(samples, pos_prototypes, neg_prototypes, num_neg_samples, k, margin=0.1,
                             temperature=0.07):
    """src/train.py:1010-1045, including its (probably unintended) [B,B] positive matrix
    (line 1033 broadcasts [B,1,D] against [1,B,D])."""
    batch_size, embed_dim = samples.shape
    neg = neg_prototypes.view(batch_size, num_neg_samples, embed_dim)
    pos_distances = pairwise_dist(samples, pos_prototypes.to(samples.dtype), k)                    # [B,B]
    if samples.is_cuda:
        # [B, neg] distances sample i <-> its own negatives: one row-pair kernel (hypret_rowpair_dist) fwd + bwd
        from .manifold import rowpair_dist
        ia = torch.arange(batch_size, device=samples.device).repeat_interleave(num_neg_samples)
        ib = torch.arange(batch_size * num_neg_samples, device=samples.device)
        neg_distances = rowpair_dist(samples, neg_prototypes.reshape(-1, embed_dim).to(samples.dtype), ia, ib,
                                     k).view(batch_size, num_neg_samples).mean(dim=1)               # [B]
    else:
        neg_distances = pmath.dist(samples.unsqueeze(1), neg.to(samples.dtype), k=k).mean(dim=1)   # [B]
    return torch.relu(pos_distances.unsqueeze(1) - neg_distances + margin).mean()


def create_n_pair_batch(indices, batch_size, figure_to_pos_figures, X_figures_tensor, device):
    """src/train.py:1758-1789 (host-side sampling; unchanged semantics)."""
    random.shuffle(indices)
    for i in range(0, len(indices), batch_size):
        anchors = indices[i:i + batch_size]
        positives, valid_anchors = [], []
        for anchor in anchors:
            cand = figure_to_pos_figures.get(anchor, [])
            if cand:
                positives.append(random.choice(cand))
                valid_anchors.append(anchor)
        pairs = [(a, p) for a, p in zip(valid_anchors, positives) if a != p]
        if not pairs:
            continue
        a_idx, p_idx = zip(*pairs)
        batch_x = X_figures_tensor[list(a_idx) + list(p_idx)].to(device)
        yield batch_x, len(a_idx), list(a_idx), list(p_idx)


def train_hyperbolic_contrastive(model, X_figures, figure_to_pos_figures, train_indices, val_indices, epochs=1,
                                 batch_size=128, lr=1e-3, temperature=0.07, device=None, save_path="best_model.pt",
                                 patience=5, return_history=False):
    """src/train.py:1792-1910 with the double loop replaced by the CUDA distance matrix.  Returns the trained model,
    as the reference does (its caller: ``trained_model = train_hyperbolic_contrastive(...)``, src/train.py:3900);
    the per-epoch ``(epoch, train_loss, val_loss)`` list is kept on ``model._train_history`` (and returned as a
    second value only with ``return_history=True``)."""
    device = torch.device(device) if device is not None else torch.device("cuda")
    model = model.to(device)
    model.k = model.k.to(device)
    X = torch.as_tensor(X_figures, dtype=torch.float32)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    best_val, bad_epochs = float("inf"), 0
    history = []
    for epoch in range(1, epochs + 1):
        model.train()
        tot, nb = 0.0, 0
        for batch_x, n, _, _ in create_n_pair_batch(list(train_indices), batch_size, figure_to_pos_figures, X, device):
            enc = model.encode_figures(batch_x)
            loss = in_batch_contrastive_loss(enc[:n], enc[n:], model.k, temperature)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            tot += float(loss.item())
            nb += 1
        model.eval()
        vtot, vb = 0.0, 0
        with torch.no_grad():
            for batch_x, n, _, _ in create_n_pair_batch(list(val_indices), batch_size, figure_to_pos_figures, X, device):
                enc = model.encode_figures(batch_x)
                vtot += float(in_batch_contrastive_loss(enc[:n], enc[n:], model.k, temperature).item())
                vb += 1
        train_loss, val_loss = tot / max(nb, 1), vtot / max(vb, 1)
        history.append((epoch, train_loss, val_loss))
        if val_loss < best_val:
            best_val, bad_epochs = val_loss, 0
            if save_path:
                torch.save(model.state_dict(), save_path)
        else:
            bad_epochs += 1
            if bad_epochs >= patience:
                break
    if save_path:
        try:
            model.load_state_dict(torch.load(save_path))
        except Exception:
            pass
    model._train_history = history
    return (model, history) if return_history else model
