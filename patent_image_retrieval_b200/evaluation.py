"""Drop-in evaluation entry points of the reference, on the CUDA path.

    evaluate_retrieval(...) -> float          /root/reference/src/train.py:3108-3296
    ImageRetrieval.retrieve_similar_images    /root/reference/notebooks/retrieval.ipynb:190-206
    evaluate_queries (notebook "Test" cell)   /root/reference/notebooks/retrieval.ipynb:368-456

``evaluate_retrieval`` keeps the reference's signature and sentinel returns (0.0 for an empty
evaluation set / no patents, -1.0 for encoding or shape errors).  The per-query Python loop
(``pmath.dist`` one-vs-all -> D2H -> sklearn AP, with a host sync per query) becomes: the exact keys of
the (query, positive) pairs, one sweep of exact distance tiles whose epilogue only counts ranks
(``hypret_rank_count``), and AP from the counts with sklearn's tie-grouping semantics -- no sort, no
[Q,P] matrix, no host round trip per query, and shardable over gallery rows (``dist.full_ranking_ap``).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .retrieval import GalleryIndex


def _csr_from_lists(lists: Sequence[Sequence[int]], device) -> tuple:
    counts = torch.tensor([len(x) for x in lists], dtype=torch.int64)
    offsets = torch.zeros(len(lists) + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    flat = [int(v) for x in lists for v in x]
    items = torch.tensor(flat, dtype=torch.int64) if flat else torch.zeros(0, dtype=torch.int64)
    return offsets.to(device), items.to(device)


def evaluate_retrieval(model, X_figures_tensor, eval_indices, figure_to_pos_patent, label_offsets, device, batch_size):
    """Evaluates retrieval performance using mAP (multiple positive patents per figure supported).
    Signature and return conventions of src/train.py:3108."""
    if not eval_indices:
        return 0.0
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("evaluate_retrieval runs on a CUDA device only (no CPU fallback)")
    model.eval()
    all_label_emb = model.label_emb.detach()
    num_labels = all_label_emb.shape[0]
    patent_start = label_offsets.get("patents", -1)
    if patent_start == -1:
        return -1.0
    patent_end = num_labels + patent_start
    nxt = [v for _, v in label_offsets.items() if v > patent_start]
    if nxt:
        patent_end = min(nxt)
    num_patents = patent_end - patent_start
    if num_patents <= 0:
        return 0.0
    patents = all_label_emb[0:num_patents].to(device=device, dtype=torch.float32).contiguous()
    c = float(-model.k.reshape(-1)[0])

    # ---- encode figures in batches (src/train.py:3152-3176) ---------------------------------------
    chunks = []
    with torch.no_grad():
        for i in range(0, len(eval_indices), batch_size):
            batch_indices = eval_indices[i:i + batch_size]
            if not batch_indices:
                continue
            try:
                if max(batch_indices) >= X_figures_tensor.shape[0]:
                    continue
                batch_x = X_figures_tensor[batch_indices].to(device)
                if batch_x.shape[0] == 0:
                    continue
                encoded = model.encode_figures(batch_x)
                if encoded.shape[0] != len(batch_indices):
                    continue
                chunks.append(encoded.to(torch.float32))
            except Exception:
                return -1.0
    if not chunks:
        return 0.0
    figure_embeddings = torch.cat(chunks, dim=0)
    if figure_embeddings.shape[0] != len(eval_indices):
        return -1.0

    # ---- positives per query, filtered to the patent range (src/train.py:3224-3242) ----------------
    pos_lists: List[List[int]] = []
    for fig_idx in eval_indices:
        entry = figure_to_pos_patent.get(fig_idx, -1)
        if isinstance(entry, list):
            cand = entry
        else:
            cand = [entry] if entry != -1 else []
        # target[idx] = 1 in the reference is idempotent: a patent listed twice counts once
        pos_lists.append(sorted({int(p) for p in cand if 0 <= p < num_patents}))
    offsets, items = _csr_from_lists(pos_lists, device)

    # ---- exact one-vs-all distances + sklearn-style AP: rank counting, no [Q,P] matrix, no per-query sync ----
    from .dist import full_ranking_ap
    if items.numel() == 0:
        return 0.0
    figure_embeddings = figure_embeddings.contiguous()
    if figure_embeddings.shape[1] != patents.shape[1]:
        return -1.0                                   # the reference's pmath.dist raises -> its except returns -1.0
    pad = (-patents.shape[1]) % 4                     # the tile kernels read 128-bit chunks: zero columns change nothing
    if pad:
        figure_embeddings = F.pad(figure_embeddings, (0, pad))
        patents = F.pad(patents, (0, pad))
    try:
        # always the unsharded path: every rank holds the whole patent table (sharded=False even under torchrun)
        mean_ap, _, valid = full_ranking_ap(figure_embeddings, patents, offsets, items, c=c, metric="hyperbolic",
                                            n_total=num_patents, grouped_ties=True, sharded=False)
    except Exception:
        return -1.0
    return mean_ap if int(valid.sum().item()) > 0 else 0.0


class ImageRetrieval:
    """The retrieval half of the notebook's ``ImageRetrieval`` (embeddings in, ranked paths out).
    The CLIP encoder half (retrieval.ipynb:88-153,165-188) is backbone code and out of scope: pass
    embeddings, e.g. the cached ``embeddings/{model}.npy`` + ``.json`` path list (retrieval.ipynb:155-163)."""

    def __init__(self, embeddings=None, image_paths: Optional[Sequence[str]] = None, device=None):
        self.embeddings = None
        self.image_paths = list(image_paths) if image_paths is not None else None
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._index = None
        if embeddings is not None:
            self.set_embeddings(embeddings, image_paths)

    def set_embeddings(self, embeddings, image_paths=None):
        self.embeddings = torch.as_tensor(np.asarray(embeddings) if not torch.is_tensor(embeddings) else embeddings,
                                          dtype=torch.float32)
        if image_paths is not None:
            self.image_paths = list(image_paths)
        self._index = GalleryIndex(self.embeddings, metric="cosine", device=self.device)

    def load_embeddings(self, npy_path, json_path):
        import json
        self.set_embeddings(np.load(npy_path), json.load(open(json_path)))

    def search(self, query_embeddings, k=20):
        if self._index is None:
            raise ValueError("No database embeddings found. Please encode dataset first.")
        q = torch.as_tensor(np.asarray(query_embeddings) if not torch.is_tensor(query_embeddings) else query_embeddings,
                            dtype=torch.float32)
        if q.dim() == 1:
            q = q[None]
        return self._index.search(q, k=k)

    def retrieve_similar_images(self, query_embedding, k=20):
        """Retrieve the k most similar images: list of (path, similarity), best first."""
        sim, idx = self.search(query_embedding, k=k)
        sim, idx = sim[0].cpu().tolist(), idx[0].cpu().tolist()
        return [(self.image_paths[i] if self.image_paths is not None else i, s) for i, s in zip(idx, sim) if i >= 0]


def evaluate_queries(retrieval: ImageRetrieval, query_embeddings, query_names: Sequence[str],
                     ground_truth: Dict[str, dict], k: int = 20, ks=(5, 10, 20)) -> Dict[str, float]:
    """The notebook's evaluation cell (retrieval.ipynb:368-456) on top-k lists: cosine scoring,
    ranking, then MRR / MRR@k / Precision@k / AP / nDCG / Recall@k averaged over the queries found
    in ``ground_truth`` (``{query_name: {"patent_positives": [gallery file names]}}``).  AP, MRR
    and nDCG are their @k variants because only the top-k of the ranking is materialised."""
    names = [Path(p).name for p in retrieval.image_paths]
    row_of = {n: i for i, n in enumerate(names)}
    keep = [i for i, q in enumerate(query_names) if Path(q).name in ground_truth]
    if not keep:
        return {}
    q = torch.as_tensor(np.asarray(query_embeddings), dtype=torch.float32)[keep]
    _, idx = retrieval.search(q, k=k)
    pos_lists, n_pos = [], []
    for i in keep:
        pos = ground_truth[Path(query_names[i]).name]["patent_positives"]
        n_pos.append(len(set(pos)))
        pos_lists.append(sorted({row_of[p] for p in pos if p in row_of}))
    off, items = _csr_from_lists(pos_lists, idx.device)
    if items.numel() == 0:
        items = torch.zeros(1, dtype=torch.int64, device=idx.device)
    means, _ = ops.retrieval_metrics(idx, off, items, ks=ks, n_pos_total=torch.tensor(n_pos, dtype=torch.int32,
                                                                                    device=idx.device))
    return means


def full_ranking_metrics(query_rows: torch.Tensor, gallery_rows: torch.Tensor, pos_offsets: torch.Tensor,
                         pos_items: torch.Tensor, n_pos_total: Optional[torch.Tensor] = None, metric: str = "cosine",
                         c: float = 1.0, ks=(5, 10, 20), row_offset: int = 0, n_total: Optional[int] = None,
                         group=None, sharded: Optional[bool] = None):
    """The notebook's whole metric suite over the FULL ranking (retrieval.ipynb:383-456: MRR, MRR@k, Precision@k,
    AP, nDCG, Recall@k) without ranking anything: every one of them is a function of the RANKS of a query's
    positives, and a rank is a count -- ``hypret_rank_count`` (ranking convention: descending similarity /
    ascending distance, ties -> lower gallery index).  Works on a gallery row-shard per rank with ``sharded=True`` (two
    all-reduces, ``dist.full_ranking_ap``'s collectives).  ``query_rows`` / ``gallery_rows``: raw features (cosine) or points on
    the ball (hyperbolic).  Returns ``(means: dict, per_query [Q, 3+3*len(ks)] fp64)`` with the columns of
    ``ops.metric_names(ks)`` -- the layout ``io.evaluation_results`` turns into the reference's results JSON."""
    import torch.distributed as dist
    from .dist import _explicitly_sharded
    dev = query_rows.device
    Q = query_rows.shape[0]
    n_total = int(n_total) if n_total is not None else int(gallery_rows.shape[0])
    off = pos_offsets.to(dev, torch.int64)
    items = pos_items.to(dev, torch.int64)
    sharded = _explicitly_sharded(sharded, group)     # explicit, like dist.full_ranking_ap
    keys = ops.pair_keys(query_rows, gallery_rows, off, items, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(keys, op=dist.ReduceOp.SUM, group=group)
    counts, _ = ops.rank_count(query_rows, gallery_rows, off, items, keys, c, metric, idx_offset=row_offset)
    if sharded:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    seg = torch.repeat_interleave(torch.arange(Q, device=dev), off[1:] - off[:-1])
    ok = (items >= 0) & (items < n_total)
    seg, rank = seg[ok], (counts[ok, 0] + counts[ok, 1] + 1)
    n_pos = (off[1:] - off[:-1]).to(torch.float64) if n_pos_total is None else n_pos_total.to(dev, torch.float64)
    order = torch.argsort(seg * (n_total + 2) + rank)                 # by query, then by rank
    seg, rank = seg[order], rank[order].to(torch.float64)
    first = torch.zeros(Q, dtype=torch.int64, device=dev)
    first[1:] = torch.cumsum(torch.bincount(seg, minlength=Q), 0)[:-1]
    nth = (torch.arange(seg.numel(), device=dev) - first[seg] + 1).to(torch.float64)     # 1-based hit number

    def seg_sum(v):
        return torch.zeros(Q, dtype=torch.float64, device=dev).index_add_(0, seg, v)

    has = torch.bincount(seg, minlength=Q) > 0
    best = torch.full((Q,), float("inf"), dtype=torch.float64, device=dev).scatter_reduce_(0, seg, rank, "amin")
    safe_n = torch.where(n_pos > 0, n_pos, torch.ones_like(n_pos))
    # ideal DCG over |P| positions (positives missing from the gallery included, retrieval.ipynb:430-437)
    max_p = int(n_pos.max().item()) if Q else 0
    disc = torch.cumsum(1.0 / torch.log2(torch.arange(max_p, device=dev, dtype=torch.float64) + 2.0), 0)
    idcg = torch.where(n_pos > 0, disc[(n_pos.long() - 1).clamp(min=0)] if max_p else n_pos, torch.ones_like(n_pos))
    cols = [torch.where(has, 1.0 / best, torch.zeros_like(best)),
            torch.where(n_pos > 0, seg_sum(nth / rank) / safe_n, torch.zeros_like(n_pos)),
            torch.where(n_pos > 0, seg_sum(1.0 / torch.log2(rank + 1.0)) / idcg, torch.zeros_like(n_pos))]
    for k in ks:
        hits = seg_sum((rank <= k).to(torch.float64))
        cols.append(torch.where(has & (best <= k), 1.0 / best, torch.zeros_like(best)))
        cols.append(hits / k if k <= n_total else torch.zeros_like(hits))
        cols.append(torch.where(n_pos > 0, hits / safe_n, torch.zeros_like(n_pos)))
    per_query = torch.stack(cols, dim=1)
    names = ops.metric_names(ks)
    means = {n: (float(per_query[:, i].mean()) if Q else 0.0) for i, n in enumerate(names)}
    return means, per_query


def evaluate_test_set(gallery_embeddings, gallery_paths: Sequence[str], query_embeddings, query_names: Sequence[str],
                      ground_truth, results_path=None, key: str = "patent_positives", device=None):
    """The notebook's "Test" cell end to end (retrieval.ipynb:190-507): ground-truth JSON (path or dict) + cached
    gallery embeddings -> exact full-ranking metrics on the GPU -> the reference's results JSON (optional).
    Queries missing from the ground truth are skipped, as in the notebook."""
    from . import io as pio
    gt = pio.load_ground_truth(ground_truth) if not isinstance(ground_truth, dict) else ground_truth
    keep, off, items, n_tot = pio.positives_csr(gt, query_names, gallery_paths, key=key)
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    q = torch.as_tensor(np.asarray(query_embeddings), dtype=torch.float32)[keep].to(device)
    g = torch.as_tensor(np.asarray(gallery_embeddings), dtype=torch.float32).to(device)
    means, per_query = full_ranking_metrics(q, g, torch.from_numpy(off), torch.from_numpy(items),
                                            n_pos_total=torch.from_numpy(n_tot), metric="cosine")
    names = ops.metric_names((5, 10, 20))
    res = (pio.save_evaluation_results(results_path, per_query, names) if results_path is not None
           else pio.evaluation_results(per_query, names))
    return res
