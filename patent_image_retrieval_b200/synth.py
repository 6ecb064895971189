"""Seeded synthetic inputs for the retrieval path.

Euclidean "backbone features" ``u ~ N(0, sigma^2 I_D)`` with ``sigma = 0.45/sqrt(D)``
(so ``||u|| ~ 0.45`` and the projected point sits at radius ~0.42/sqrt(c), where
the geoopt fp32 form and the arccosh form agree to < 1e-6), a *clustered* variant
that gives non-trivial recall (``u = (mu_cls + 0.3 eps) sigma`` with N/8 classes;
positives of a query = gallery rows of its class), and a boundary-stress set that
hits the ``project`` clip.

Seeds follow the survey: 0 = gallery, 1 = queries, 2 = class ids / centres.
CPU generation is bit-reproducible; CUDA generation (used for the full-size
bench shapes that the CPU oracle cannot touch anyway) is reproducible per device
type only.
"""
from __future__ import annotations

import torch

SEED_GALLERY, SEED_QUERY, SEED_LABEL = 0, 1, 2


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def gaussian_features(n: int, d: int, seed: int, device="cpu", scale: float = 0.45,
                      chunk: int = 1 << 20) -> torch.Tensor:
    """[n,d] f32, ``N(0, (scale/sqrt(d))^2)``; generated in row chunks so that 10M x 512
    never needs a second full-size temporary."""
    g = _gen(seed, device)
    out = torch.empty(n, d, dtype=torch.float32, device=device)
    sigma = scale / (d ** 0.5)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        out[r0:r1].normal_(0.0, sigma, generator=g)
    return out


GALLERY_BLOCK = 1 << 18


def gallery_rows(lo: int, hi: int, d: int, device="cpu", scale: float = 0.45, seed: int = SEED_GALLERY) -> torch.Tensor:
    """Rows ``[lo, hi)`` of THE bench gallery: ``N(0, (scale/sqrt(d))^2)`` in blocks of 2^18 rows, block b seeded
    ``seed + 7919 * (b + 1)`` -- so the gallery is the same whatever the number of GPUs it is sharded over (a rank
    generates the blocks its row range touches).  Reproducible per device type, like ``gaussian_features``."""
    out = torch.empty(max(hi - lo, 0), d, dtype=torch.float32, device=device)
    sigma = scale / (d ** 0.5)
    for b in range(lo // GALLERY_BLOCK, (max(hi, lo + 1) - 1) // GALLERY_BLOCK + 1 if hi > lo else 0):
        b0 = b * GALLERY_BLOCK
        r0, r1 = max(lo, b0), min(hi, b0 + GALLERY_BLOCK)
        if r0 == b0 and r1 == b0 + GALLERY_BLOCK:
            out[r0 - lo:r1 - lo].normal_(0.0, sigma, generator=_gen(seed + 7919 * (b + 1), device))
        else:                                     # partial block: generate it whole, keep the slice
            blk = torch.empty(GALLERY_BLOCK, d, dtype=torch.float32, device=device)
            blk.normal_(0.0, sigma, generator=_gen(seed + 7919 * (b + 1), device))
            out[r0 - lo:r1 - lo] = blk[r0 - b0:r1 - b0]
    return out


def clustered_features(n_gallery: int, n_query: int, d: int, device="cpu", per_class: int = 8,
                       noise: float = 0.3, scale: float = 0.45):
    """Returns (gallery_u [N,d], query_u [Q,d], gallery_cls [N] i64, query_cls [Q] i64)."""
    n_cls = max(1, n_gallery // per_class)
    gl = _gen(SEED_LABEL, device)
    mu = torch.randn(n_cls, d, generator=gl, device=device)
    g_cls = torch.randint(0, n_cls, (n_gallery,), generator=gl, device=device)
    q_cls = torch.randint(0, n_cls, (n_query,), generator=gl, device=device)
    sigma = scale / (d ** 0.5)
    gg, gq = _gen(SEED_GALLERY, device), _gen(SEED_QUERY, device)
    gal = (mu[g_cls] + noise * torch.randn(n_gallery, d, generator=gg, device=device)) * sigma
    qry = (mu[q_cls] + noise * torch.randn(n_query, d, generator=gq, device=device)) * sigma
    return gal, qry, g_cls, q_cls


def boundary_features(n: int, d: int, seed: int, device="cpu", max_norm: float = 3.0) -> torch.Tensor:
    """Rows with ||u|| spread over (0, max_norm]; after expmap0 many of them hit the
    ``project`` clip.  Graded against the fp64 oracle only."""
    g = _gen(seed, device)
    x = torch.randn(n, d, generator=g, device=device)
    x = x / x.norm(dim=1, keepdim=True)
    r = torch.rand(n, 1, generator=g, device=device) * max_norm
    return (x * r).float()


def positives_csr(query_cls: torch.Tensor, gallery_cls: torch.Tensor):
    """CSR (offsets[Q+1] i64, items[nnz] i64 ascending per query) of the gallery rows
    sharing each query's class."""
    order = torch.argsort(gallery_cls, stable=True)
    sorted_cls = gallery_cls[order]
    lo = torch.searchsorted(sorted_cls, query_cls, right=False)
    hi = torch.searchsorted(sorted_cls, query_cls, right=True)
    counts = hi - lo
    offsets = torch.zeros(query_cls.numel() + 1, dtype=torch.int64, device=query_cls.device)
    offsets[1:] = torch.cumsum(counts, 0)
    nnz = int(offsets[-1])
    # expand ranges
    rep = torch.repeat_interleave(torch.arange(query_cls.numel(), device=query_cls.device), counts)
    within = torch.arange(nnz, device=query_cls.device) - offsets[:-1][rep]
    items = order[lo[rep] + within]
    return offsets, items
