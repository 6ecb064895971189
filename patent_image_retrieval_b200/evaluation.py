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
        pos_lists.append([int(p) for p in cand if 0 <= p < num_patents])
    offsets, items = _csr_from_lists(pos_lists, device)

    # ---- exact one-vs-all distances + sklearn-style AP: rank counting, no [Q,P] matrix, no per-query sync ----
    from .dist import full_ranking_ap
    if items.numel() == 0:
        return 0.0
    mean_ap, _, valid = full_ranking_ap(figure_embeddings.contiguous(), patents, offsets, items, c=c,
                                        metric="hyperbolic", n_total=num_patents, grouped_ties=True)
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
