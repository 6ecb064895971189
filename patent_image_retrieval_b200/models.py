"""Drop-in hyperbolic projection head: the reference's ``models.py`` classes with their
signatures, attribute names and state-dict keys, minus the geoopt dependency.

    MobiusLinear / mobius_linear          /root/reference/src/models.py:255-318
    DeeperHyperbolicEncoder               src/models.py:447-505
    HyperbolicEmbeddingModel              src/models.py:507-784
    FigureOnlyHyperbolicModel             src/models.py:788-838

State-dict keys are the reference's: ``encoder.first_layer.{weight,bias}``,
``encoder.final_layer.{weight,bias}``, ``label_emb`` -- checkpoints such as
``best_retrieval_model_c{c}_e{dim}.pt`` (src/train.py:1631) round-trip.  ``k`` stays a plain
attribute (a 1-element fp32 tensor ``[-c]``), not a buffer, exactly as in the reference.

Differences that restate intended behaviour instead of a crash / an accident:
  * src/models.py:306 applies ``F.dropout(weight, dropout)`` with ``dropout`` undefined
    (NameError as shipped): weight dropout is the identity here;
  * the reference flips torch's default dtype to float64 at import (src/models.py:248-249);
    this module does not touch global state -- construct under ``torch.set_default_dtype`` if the
    accidental fp64 parameters are wanted;
  * the per-pair Python loops of ``calculate_pair_loss`` are evaluated as one batched distance
    (same arithmetic per pair).
The hierarchy / regulariser losses (src/models.py:550-674: ``calculate_hierarchical_loss``, ``calculate_reg_loss``,
``_hmi_insideness``, ``_hmi_disjointedness``) run as one kernel forward + one backward each on CUDA
(``manifold.py`` / csrc/manifold.cu) and as op-by-op torch on CPU tensors.

Note on similarity with the reference: ``MobiusLinear`` / ``mobius_linear`` and the three constructors restate
src/models.py:255-318, 447-479, 507-535, 788-801 line for line -- signatures, attribute names, init order and state-dict
keys are the drop-in contract and the class is itself geoopt's example layer.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import geoopt_shim as gt
from .geoopt_shim import pmath

DROPOUT_RATE = 0.1      # src/models.py:16


class MobiusLinear(nn.Linear):
    def __init__(self, *args, hyperbolic_input=True, hyperbolic_bias=True, nonlin=None, c=1.0, **kwargs):
        super().__init__(*args, **kwargs)
        self.ball = gt.PoincareBall(c=c)
        if self.bias is not None:
            if hyperbolic_bias:
                self.bias = gt.ManifoldParameter(self.bias, manifold=self.ball)
                with torch.no_grad():
                    self.bias.set_(pmath.expmap0(self.bias.normal_() * 1e-3, k=self.ball.k))
        with torch.no_grad():
            fin, fout = self.weight.size()
            k = (6 / (fin + fout)) ** 0.5  # xavier uniform
            self.weight.uniform_(-k, k)
        self.hyperbolic_bias = hyperbolic_bias
        self.hyperbolic_input = hyperbolic_input
        self.nonlin = nonlin

    def forward(self, input):
        return mobius_linear(input, weight=self.weight, bias=self.bias, hyperbolic_input=self.hyperbolic_input,
                             nonlin=self.nonlin, hyperbolic_bias=self.hyperbolic_bias, k=self.ball.k)

    def extra_repr(self):
        info = super().extra_repr()
        info += ", hyperbolic_input={}".format(self.hyperbolic_input)
        if self.bias is not None:
            info += ", hyperbolic_bias={}".format(self.hyperbolic_bias)
        return info


def mobius_linear(input, weight, bias=None, hyperbolic_input=True, hyperbolic_bias=True, nonlin=None, k=-1.0):
    weight = weight.to(input.dtype)
    if bias is not None:
        bias = bias.to(input.dtype)
    if hyperbolic_input:
        output = pmath.mobius_matvec(weight, input, k=k)
    else:
        output = F.linear(input, weight)
        output = pmath.expmap0(output, k=k)
    if bias is not None:
        if not hyperbolic_bias:
            bias = pmath.expmap0(bias, k=k)
        output = pmath.mobius_add(output, bias, k=k)
    if nonlin is not None:
        output = pmath.mobius_fn_apply(nonlin, output, k=k)
    return pmath.project(output, k=k)


class DeeperHyperbolicEncoder(nn.Module):
    def __init__(self, input_dim, hidden_dims, output_dim, c=1.0, dropout_rate=0.3):
        super().__init__()
        self.c = c
        self.ball = gt.PoincareBall(c=c)
        self.k = torch.tensor([-c], dtype=torch.float32)
        self.dropout_rate = dropout_rate
        self.first_layer = MobiusLinear(input_dim, hidden_dims[0], hyperbolic_input=False, c=c)
        self.final_layer = MobiusLinear(hidden_dims[0], output_dim, hyperbolic_input=True, c=c)

    def _fused_ok(self, x):
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and not self.training
                and not (torch.is_grad_enabled() and (x.requires_grad or self.first_layer.weight.requires_grad))
                and self.first_layer.bias is not None and self.final_layer.bias is not None
                and self.first_layer.out_features % 4 == 0 and self.final_layer.out_features % 4 == 0)

    def _kernel_train_ok(self, x):
        from . import ops
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0
                and self.first_layer.bias is not None and self.final_layer.bias is not None
                and self.first_layer.hyperbolic_bias and self.final_layer.hyperbolic_bias
                and ops.mobius_gemm_ok(self.first_layer.in_features, self.first_layer.out_features)
                and ops.mobius_gemm_ok(self.first_layer.out_features, self.final_layer.out_features))

    def forward(self, x):
        self.k = self.k.to(x.device)
        if self._fused_ok(x):
            from . import ops
            c = float(self.c)
            d_in, hid, d_out = self.first_layer.in_features, self.first_layer.out_features, self.final_layer.out_features
            if ops.mobius_gemm_ok(d_in, hid) and ops.mobius_gemm_ok(hid, d_out):
                # inference on the GPU: each layer is ONE kernel -- tcgen05 GEMM + the Moebius epilogue on the
                # accumulator (csrc/headgemm.cu); the hidden activations reach the second GEMM as its fp16 split
                # operand, never as an fp32 [B, hidden] tensor
                h = ops.mobius_gemm(ops.split_operand(x, "row"), ops.split_operand(self.first_layer.weight, "col"), d_in,
                                    hid, c, bias=self.first_layer.bias, post_tanh=True, n_project=1, want_y=False,
                                    want_op=True, want_sqnorm=True)
                return ops.mobius_gemm(h["op"], ops.split_operand(self.final_layer.weight, "col"), hid, d_out, c,
                                       bias=self.final_layer.bias, xsq=h["sq"], n_project=2)["y"]
            # other layer widths: two library GEMMs + two fused epilogue kernels (csrc/head.cu)
            h, hsq = ops.mobius_epilogue(F.linear(x, self.first_layer.weight.to(x.dtype)), c,
                                         bias=self.first_layer.bias, post_tanh=True, n_project=1)
            y, _ = ops.mobius_epilogue(F.linear(h, self.final_layer.weight.to(x.dtype)), c,
                                       bias=self.final_layer.bias, xsq=hsq, n_project=2, want_sqnorm=False)
            return y
        if self._kernel_train_ok(x):
            # training on the GPU: each layer is one autograd node whose forward is the fused GEMM kernel and whose
            # backward is one closed-form epilogue kernel + the dense products (ops.MobiusLinearFn); only the two
            # dropouts stay with torch
            from . import ops
            c = float(self.c)
            x = F.dropout(x, p=self.dropout_rate, training=self.training)
            x = ops.MobiusLinearFn.apply(x, self.first_layer.weight, self.first_layer.bias, c, False, True, 1)
            x = F.dropout(x, p=self.dropout_rate, training=self.training)
            return ops.MobiusLinearFn.apply(x, self.final_layer.weight, self.final_layer.bias, c, True, False, 2)
        x = F.dropout(x, p=self.dropout_rate, training=self.training)
        x = self.first_layer(x)
        x = pmath.mobius_fn_apply(torch.tanh, x, k=self.k)
        x = F.dropout(x, p=self.dropout_rate, training=self.training)
        x = self.final_layer(x)
        return pmath.project(x, k=self.k)


MIN_NORM = 1e-15        # src/models.py:15


def _pair_distances(figure_embeddings, all_pairs, k):
    """Batched replacement of the per-pair loop (src/models.py:712-719, 824-829): one row-pair distance kernel
    forward and one backward on CUDA (``hypret_rowpair_dist``), the same arithmetic op by op on CPU tensors."""
    if figure_embeddings.is_cuda:
        from .manifold import rowpair_dist
        return rowpair_dist(figure_embeddings, figure_embeddings, all_pairs[:, 0], all_pairs[:, 1], k)
    return pmath.dist(figure_embeddings[all_pairs[:, 0]], figure_embeddings[all_pairs[:, 1]], k=k)


def _hmi_terms(ball, k, point_a, point_b, dim=-1):
    """Radii and centre distance of the two HMI balls (src/models.py:636-654): (radius_a, radius_b, centre_dist)."""
    point_a, point_b = ball.projx(point_a), ball.projx(point_b)
    na = torch.norm(point_a, p=2, dim=dim, keepdim=True).clamp_min(MIN_NORM)
    nb = torch.norm(point_b, p=2, dim=dim, keepdim=True).clamp_min(MIN_NORM)
    kt = k.to(na.device)
    s = torch.sqrt(-kt)
    ra, rb = (1 + kt * na ** 2) / (2 * s * na), (1 + kt * nb ** 2) / (2 * s * nb)
    ca, cb = point_a * (1 + ra * s / na), point_b * (1 + rb * s / nb)
    return ra, rb, torch.norm(ca - cb, p=2, dim=dim, keepdim=True)


def _all_pairs(figure_embeddings, positive_pairs, negative_pairs):
    if positive_pairs.dim() == 1:
        positive_pairs = positive_pairs.view(-1, 2)
    all_pairs = positive_pairs
    if negative_pairs is not None and negative_pairs.numel() > 0:
        if negative_pairs.dim() == 1:
            negative_pairs = negative_pairs.view(-1, 2)
        all_pairs = torch.cat([positive_pairs, negative_pairs], dim=0)
    labels = torch.zeros(len(all_pairs), device=figure_embeddings.device)
    labels[:len(positive_pairs)] = 1.0
    return all_pairs, labels


class HyperbolicEmbeddingModel(nn.Module):
    def __init__(self, feature_num, embed_dim, label_num, hidden_dims=[256, 128], c=1.0, **kwargs):
        super().__init__(**kwargs)
        self.c = c
        self.embed_dim = embed_dim
        self.k = torch.tensor([-c], dtype=torch.float32)
        self.ball = gt.PoincareBall(c=self.c)
        self.temperature = 0.07
        label_points = torch.randn(label_num, embed_dim) * 0.1
        label_points = pmath.expmap0(label_points, k=self.k)
        self.label_emb = gt.ManifoldParameter(label_points, manifold=self.ball)
        self.encoder = DeeperHyperbolicEncoder(input_dim=feature_num, hidden_dims=hidden_dims, output_dim=embed_dim,
                                               c=c, dropout_rate=DROPOUT_RATE)

    def encode_figures(self, features):
        """Encodes Euclidean figure features into the hyperbolic space (src/models.py:537-548)."""
        features = torch.as_tensor(features, dtype=torch.float32)
        features = F.dropout(features, p=DROPOUT_RATE, training=self.training)
        encoded = self.encoder(features)
        self.ball.assert_check_point_on_manifold(encoded)
        return encoded

    def calculate_hierarchical_loss(self, implication_pairs, exclusion_pairs):
        """Label-hierarchy losses (src/models.py:550-604): ``relu(0.05 - insideness(sub, par)).mean()`` over the
        implication pairs and ``relu(0.1 - disjointedness(l, r)).mean()`` over the exclusion pairs."""
        dev = self.label_emb.device
        inside_loss = torch.tensor(0.0, device=dev)
        disjoint_loss = torch.tensor(0.0, device=dev)
        self.k = self.k.to(dev)
        for pairs, mode, margin in ((implication_pairs, "insideness", 0.05), (exclusion_pairs, "disjointedness", 0.1)):
            if pairs is None or pairs.numel() == 0:
                continue
            pairs = pairs.view(-1, 2) if pairs.dim() == 1 else pairs
            if self.label_emb.is_cuda:
                from .manifold import hmi_pair_loss
                loss = hmi_pair_loss(self.label_emb, pairs.to(dev), self.k, mode, margin)
            else:
                if int(pairs.min()) < 0 or int(pairs.max()) >= self.label_emb.shape[0]:
                    raise IndexError("Invalid index detected in implication pairs")
                fn = self._hmi_insideness if mode == "insideness" else self._hmi_disjointedness
                loss = F.relu(-fn(self.label_emb[pairs[:, 0]], self.label_emb[pairs[:, 1]]) + margin).mean()
            if mode == "insideness":
                inside_loss = loss
            else:
                disjoint_loss = loss
        return inside_loss, disjoint_loss

    def calculate_reg_loss(self, encoded_figures):
        """dist0 regularisers (src/models.py:606-628): labels kept in ``2 <= dist0 <= 8``, figures in ``dist0 <= 8``."""
        self.k = self.k.to(self.label_emb.device)
        if self.label_emb.is_cuda and encoded_figures.is_cuda:
            from .manifold import dist0_reg_loss
            return (dist0_reg_loss(self.label_emb, self.k, 2.0, 8.0),
                    dist0_reg_loss(encoded_figures.reshape(-1, encoded_figures.shape[-1]), self.k, None, 8.0))
        label_dist0 = self.ball.dist0(self.label_emb, dim=-1, keepdim=True).clamp_min(MIN_NORM)
        label_reg = (F.relu(2 - label_dist0) + F.relu(label_dist0 - 8.0)).mean()
        figure_dist0 = self.ball.dist0(encoded_figures, dim=-1, keepdim=True).clamp_min(MIN_NORM)
        return label_reg, F.relu(figure_dist0 - 8.0).mean()

    def _hmi_insideness(self, point_a, point_b, dim=-1):
        ra, rb, cd = _hmi_terms(self.ball, self.k, point_a, point_b, dim)
        return (rb - ra) - cd

    def _hmi_disjointedness(self, point_a, point_b, dim=-1):
        ra, rb, cd = _hmi_terms(self.ball, self.k, point_a, point_b, dim)
        return cd - (ra + rb)

    def calculate_pair_loss(self, figure_embeddings, positive_pairs, negative_pairs):
        """Per query figure: cross-entropy of -d/T over its pairs, positive first (src/models.py:676-757)."""
        self.k = self.k.to(figure_embeddings.device)
        if positive_pairs is None or positive_pairs.numel() == 0:
            return torch.tensor(0.0, device=figure_embeddings.device)
        all_pairs, labels = _all_pairs(figure_embeddings, positive_pairs, negative_pairs)
        similarities = -_pair_distances(figure_embeddings, all_pairs, self.k) / self.temperature
        total_loss, num_queries = 0.0, 0
        for query_idx in torch.unique(all_pairs[:, 0]):
            mask = all_pairs[:, 0] == query_idx
            q_lab = labels[mask]
            if not q_lab.any():
                continue
            total_loss = total_loss + F.cross_entropy(similarities[mask].unsqueeze(0), q_lab.argmax().unsqueeze(0))
            num_queries += 1
        if num_queries > 0:
            return total_loss / num_queries
        return torch.tensor(0.0, device=figure_embeddings.device)

    def forward(self, figure_features, implication_pairs=None, exclusion_pairs=None):
        self.k = self.k.to(figure_features.device)
        return self.encode_figures(figure_features)


class FigureOnlyHyperbolicModel(nn.Module):
    def __init__(self, feature_num, embed_dim, hidden_dims=[256, 128], c=1.0, dropout_rate=0.3):
        super().__init__()
        self.c = c
        self.embed_dim = embed_dim
        self.k = torch.tensor([-c], dtype=torch.float32)
        self.ball = gt.PoincareBall(c=self.c)
        self.encoder = DeeperHyperbolicEncoder(input_dim=feature_num, hidden_dims=hidden_dims, output_dim=embed_dim,
                                               c=c, dropout_rate=dropout_rate)

    def encode_figures(self, features):
        features = F.dropout(features, p=self.encoder.dropout_rate, training=self.training)
        encoded = self.encoder(features)
        self.ball.assert_check_point_on_manifold(encoded)
        return encoded

    def calculate_pair_loss(self, figure_embeddings, positive_pairs, negative_pairs, temperature=0.07):
        """BCE-with-logits on -d/T over positive + negative pairs (src/models.py:809-832)."""
        self.k = self.k.to(figure_embeddings.device)
        if positive_pairs is None or positive_pairs.numel() == 0:
            return torch.tensor(0.0, device=figure_embeddings.device)
        all_pairs, labels = _all_pairs(figure_embeddings, positive_pairs, negative_pairs)
        similarities = -_pair_distances(figure_embeddings, all_pairs, self.k) / temperature
        return F.binary_cross_entropy_with_logits(similarities, labels.float())

    def forward(self, features):
        self.k = self.k.to(features.device)
        return self.encode_figures(features)
