"""On-disk formats of the reference on either side of the hot path, so real runs work unchanged.

    gallery cache        embeddings/{model_name}.npy  [N,D] f32  + embeddings/{model_name}.json  (list of image paths)
                         /root/reference/notebooks/retrieval.ipynb cell 1 (save: encode_dataset, load: load_embeddings)
    training data        training_data.npz: X_figures [F,512] (cast to f32 on load, src/train.py:1164),
                         Y_pos / Y_neg / implication / exclusion [*,2] int, optional positive_figure_pairs /
                         negative_figure_pairs [*,2];  label_offsets.json {patents, medium_cpcs, big_cpcs, main_cpcs}
                         /root/reference/src/train.py:3940-3984
    figure maps          figure_to_pos_patent {figure: patent index relative to the patent block} (src/train.py:1178-1208),
                         figure_to_pos_figures {figure: [figures of the same patent]}, symmetric (src/train.py:1143-1150)
    ground truth         {query file name: {"patent_positives": [gallery file names], "cpc_positives": [...]}}
                         /root/reference/notebooks/retrieval.ipynb cell 3 (`ground_truth[query_name]['patent_positives']`)
    evaluation results   {"query_wise_metrics": {...per-query lists...}, "summary_metrics": {"MRR", "MRR@5", ..., "Precision@20"}}
                         written with json.dump(indent=2), retrieval.ipynb cell 3

Host-side only (numpy / json): nothing here touches the GPU.
"""
from __future__ import annotations

import json
import os
from collections import defaultdict
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

PAIR_KEYS = ("Y_pos", "Y_neg", "implication", "exclusion")
OPTIONAL_PAIR_KEYS = ("positive_figure_pairs", "negative_figure_pairs")
LABEL_OFFSET_KEYS = ("patents", "medium_cpcs", "big_cpcs", "main_cpcs")


# ----------------------------------------------------------------------------- gallery cache
def save_gallery_cache(directory, model_name: str, embeddings, image_paths: Sequence[str]) -> Tuple[Path, Path]:
    """``np.save(f'embeddings/{model_name}.npy', embeddings)`` + the JSON path list (retrieval.ipynb cell 1)."""
    emb = np.asarray(embeddings)
    if emb.ndim != 2 or emb.shape[0] != len(image_paths):
        raise ValueError("embeddings must be [N,D] with one image path per row")
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    npy, js = d / f"{model_name}.npy", d / f"{model_name}.json"
    np.save(npy, emb)
    with open(js, "w") as f:
        json.dump([str(p) for p in image_paths], f)
    return npy, js


def load_gallery_cache(directory, model_name: str):
    """Returns ``(embeddings [N,D] float32, image_paths)`` or ``None`` when either file is missing -- the
    reference's ``load_embeddings`` silently leaves the index empty in that case."""
    d = Path(directory)
    npy, js = d / f"{model_name}.npy", d / f"{model_name}.json"
    if not (npy.exists() and js.exists()):
        return None
    emb = np.load(npy)
    with open(js) as f:
        paths = json.load(f)
    if emb.ndim != 2 or emb.shape[0] != len(paths):
        raise ValueError(f"{npy} has {emb.shape} rows but {js} lists {len(paths)} paths")
    return np.ascontiguousarray(emb, dtype=np.float32), list(paths)


# ----------------------------------------------------------------------------- training data
@dataclass
class TrainingData:
    X_figures: np.ndarray                       # [F,D] float32
    Y_pos: np.ndarray                           # [*,2] (figure, patent index relative to the patent block)
    Y_neg: np.ndarray
    implication: np.ndarray                     # [*,2] label -> label
    exclusion: np.ndarray
    label_offsets: Dict[str, int]
    positive_figure_pairs: Optional[np.ndarray] = None
    negative_figure_pairs: Optional[np.ndarray] = None
    extra: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def num_patents(self) -> int:
        """Patents are the first labels; the block ends where the next label type starts (src/train.py:4009)."""
        start = self.label_offsets["patents"]
        nxt = [v for v in self.label_offsets.values() if v > start]
        return (min(nxt) - start) if nxt else 0

    def label_num(self, num_main_cpcs: int = 9) -> int:
        """LABEL_NUM of src/train.py:4009-4016 (the count of the last label type is not stored on disk)."""
        o = self.label_offsets
        return ((o["medium_cpcs"] - o["patents"]) + (o["big_cpcs"] - o["medium_cpcs"]) +
                (o["main_cpcs"] - o["big_cpcs"]) + num_main_cpcs)


def _pairs(a) -> np.ndarray:
    a = np.asarray(a)
    if a.size == 0:
        return np.zeros((0, 2), dtype=np.int64)
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError(f"pair array must be [*,2], got {a.shape}")
    return a.astype(np.int64)


def save_training_data(directory, data: TrainingData) -> None:
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    arrays = {"X_figures": np.asarray(data.X_figures)}
    for k in PAIR_KEYS:
        arrays[k] = np.asarray(getattr(data, k), dtype=np.int32).reshape(-1, 2)
    for k in OPTIONAL_PAIR_KEYS:
        if getattr(data, k) is not None:
            arrays[k] = np.asarray(getattr(data, k), dtype=np.int32).reshape(-1, 2)
    arrays.update(data.extra)
    np.savez(d / "training_data.npz", **arrays)
    with open(d / "label_offsets.json", "w") as f:
        json.dump({k: int(v) for k, v in data.label_offsets.items()}, f)


def load_training_data(directory) -> TrainingData:
    """``training_data.npz`` + ``label_offsets.json`` of one prepared-data directory (src/train.py:3940-3984)."""
    d = Path(directory)
    with np.load(d / "training_data.npz") as z:
        keys = set(z.files)
        missing = [k for k in ("X_figures",) + PAIR_KEYS if k not in keys]
        if missing:
            raise KeyError(f"training_data.npz lacks {missing}")
        x = np.ascontiguousarray(z["X_figures"], dtype=np.float32)          # cast to f32 on load (train.py:1164)
        pairs = {k: _pairs(z[k]) for k in PAIR_KEYS}
        opt = {k: (_pairs(z[k]) if k in keys else None) for k in OPTIONAL_PAIR_KEYS}
        extra = {k: z[k] for k in keys - {"X_figures"} - set(PAIR_KEYS) - set(OPTIONAL_PAIR_KEYS)}
    with open(d / "label_offsets.json") as f:
        offsets = {k: int(v) for k, v in json.load(f).items()}
    for k in LABEL_OFFSET_KEYS:
        if k not in offsets:
            raise KeyError(f"label_offsets.json lacks '{k}'")
    return TrainingData(X_figures=x, label_offsets=offsets, extra=extra, **pairs, **opt)


def figure_to_pos_patent(Y_pos, num_figures: int, num_labels: int) -> Dict[int, int]:
    """src/train.py:1178-1208: last pair wins; pairs with an out-of-range figure or label index are skipped."""
    out: Dict[int, int] = {}
    for fig, pat in np.asarray(Y_pos).reshape(-1, 2).tolist():
        if 0 <= fig < num_figures and 0 <= pat < num_labels:
            out[int(fig)] = int(pat)
    return out


def figure_to_pos_figures(positive_figure_pairs, num_figures: Optional[int] = None) -> Dict[int, List[int]]:
    """src/train.py:1143-1150: symmetric adjacency lists of the positive figure pairs."""
    out: Dict[int, List[int]] = defaultdict(list)
    if positive_figure_pairs is None:
        return {}
    for a, b in np.asarray(positive_figure_pairs).reshape(-1, 2).tolist():
        if num_figures is not None and not (0 <= a < num_figures and 0 <= b < num_figures):
            continue
        out[int(a)].append(int(b))
        out[int(b)].append(int(a))
    return dict(out)


# ----------------------------------------------------------------------------- ground truth / results
def load_ground_truth(path) -> Dict[str, dict]:
    with open(path) as f:
        gt = json.load(f)
    if not isinstance(gt, dict):
        raise ValueError("ground truth must map query names to {'patent_positives': [...], ...}")
    return gt


def positives_csr(ground_truth: Dict[str, dict], query_names: Sequence[str], gallery_paths: Sequence[str],
                  key: str = "patent_positives"):
    """CSR of each query's positives as gallery ROW indices, as the notebook matches them: by file NAME
    (``Path(p).name``).  Returns ``(keep, offsets [len(keep)+1] int64, items int64, n_pos_total int32)`` where
    ``keep`` are the positions in ``query_names`` found in the ground truth (the others are skipped, as in the
    notebook) and ``n_pos_total`` counts ALL listed positives, in the gallery or not (Recall / AP denominators,
    retrieval.ipynb:411-443)."""
    row_of = {}
    for i, p in enumerate(gallery_paths):
        row_of.setdefault(Path(p).name, i)
    keep, lists, n_tot = [], [], []
    for qi, q in enumerate(query_names):
        entry = ground_truth.get(Path(q).name)
        if entry is None:
            continue
        pos = list(entry.get(key, []))
        keep.append(qi)
        n_tot.append(len(set(pos)))
        lists.append(sorted({row_of[Path(p).name] for p in pos if Path(p).name in row_of}))
    offsets = np.zeros(len(lists) + 1, dtype=np.int64)
    if lists:
        offsets[1:] = np.cumsum([len(x) for x in lists])
    items = np.array([v for x in lists for v in x], dtype=np.int64)
    return keep, offsets, items, np.array(n_tot, dtype=np.int32)


_QUERY_WISE = (("reciprocal_ranks", "mrr"), ("reciprocal_ranks@5", "mrr@5"), ("reciprocal_ranks@20", "mrr@20"),
               ("ap_scores", "ap"), ("ndcg_scores", "ndcg"), ("recall_5", "recall@5"), ("recall_10", "recall@10"),
               ("recall_20", "recall@20"), ("precision_5", "precision@5"), ("precision_10", "precision@10"),
               ("precision_20", "precision@20"))
_SUMMARY = (("MRR", "mrr"), ("MRR@5", "mrr@5"), ("MRR@20", "mrr@20"), ("mAP", "ap"), ("mNDCG", "ndcg"),
            ("Recall@5", "recall@5"), ("Recall@10", "recall@10"), ("Recall@20", "recall@20"),
            ("Precision@5", "precision@5"), ("Precision@10", "precision@10"), ("Precision@20", "precision@20"))


def evaluation_results(per_query, names: Sequence[str]) -> dict:
    """The ``detailed_results`` dict of the notebook's evaluation cell from the ``per_query`` matrix of
    ``ops.retrieval_metrics`` (columns named ``names`` = ``ops.metric_names((5, 10, 20))``)."""
    per = np.asarray(per_query.detach().cpu() if hasattr(per_query, "detach") else per_query, dtype=np.float64)
    col = {n: per[:, i] for i, n in enumerate(names)}
    missing = [m for _, m in _QUERY_WISE if m not in col]
    if missing:
        raise KeyError(f"metrics {missing} missing: evaluate with ks=(5, 10, 20)")
    return {"query_wise_metrics": {k: [float(v) for v in col[m]] for k, m in _QUERY_WISE},
            "summary_metrics": {k: float(np.mean(col[m])) if per.shape[0] else 0.0 for k, m in _SUMMARY}}


def save_evaluation_results(path, per_query, names: Sequence[str]) -> dict:
    res = evaluation_results(per_query, names)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(res, f, indent=2)
    return res
