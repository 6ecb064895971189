"""Generate tests/golden/*.npz.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Two kinds of vectors are written:

* ``ref_*``  -- produced by the REFERENCE'S OWN code or its own third-party calls, run here:
    - ``auxiliary.mean_average_precision`` imported from /root/reference/src/auxiliary.py:200-224
    - the metric code of notebooks/retrieval.ipynb cell 3 (helper defs :310-324 and the
      per-query block :391-443), extracted from the .ipynb JSON and exec'd verbatim
    - ``sklearn.metrics.pairwise.cosine_similarity`` + ``np.argsort(sim)[::-1]`` exactly as
      notebooks/retrieval.ipynb:368,383, and ``sklearn.metrics.average_precision_score``
      as src/train.py:3285
  These pin the oracle's cosine / metric functions.

* ``kat_*``  -- known-answer vectors MINTED BY THE ORACLE ITSELF in fp64 (geoopt is not
  available, so the hyperbolic arithmetic is PARITY UNPINNED; these only freeze the
  restatement against regressions and give the GPU tests fixed inputs).

Nothing under tests/ reads /root/reference at test time; only this script does.
"""
from __future__ import annotations

import json
import sys
import textwrap
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))


def notebook_sources():
    nb = json.load(open(REF / "notebooks" / "retrieval.ipynb"))
    src = "".join(nb["cells"][3]["source"])
    i0 = src.find("def calculate_mrr_at_k")
    i0 = src.rfind("\n", 0, i0) + 1
    i1 = src.find("# Evaluation metrics")
    i1 = src.rfind("\n", 0, i1) + 1
    helpers = textwrap.dedent(src[i0:i1])
    j0 = src.find("# Calculate MRR, MRR@5, and MRR@20")
    j0 = src.rfind("\n", 0, j0) + 1
    j1 = src.find("# Calculate final metrics")
    j1 = src.rfind("\n", 0, j1) + 1
    block = textwrap.dedent(src[j0:j1])
    return helpers, block


def run_notebook_metrics(ranked_names, positives_sets):
    helpers, block = notebook_sources()
    ns = {"np": np}
    exec(helpers, ns)
    names = ["ap_scores", "ndcg_scores", "recall_5", "recall_10", "recall_20", "reciprocal_ranks",
             "reciprocal_ranks_5", "reciprocal_ranks_20", "precision_5", "precision_10", "precision_20"]
    for n in names:
        ns[n] = []
    code = compile(block, "retrieval.ipynb:cell3", "exec")
    for retrieved_paths, positives in zip(ranked_names, positives_sets):
        ns["retrieved_paths"] = retrieved_paths
        ns["positives"] = positives
        exec(code, ns)
    return {n: np.asarray(ns[n], dtype=np.float64) for n in names}


def reference_code_with_shim():
    """Run the REFERENCE'S OWN ``models.py`` / ``train.py`` functions (imported from
    /root/reference/src) with geoopt replaced by this repo's shim -- geoopt itself is not
    installable.  This pins the reference's call order, reductions, sentinels and metric
    conventions around the (unpinned) hyperbolic arithmetic.  ``refshim_*`` keys.

    Patches needed to make the shipped code run at all:
      * ``models.dropout = 0.0``   (src/models.py:306 uses an undefined global ``dropout``)
      * ``self.temperature`` is never set in HyperbolicEmbeddingModel (src/models.py:725) -- that
        method is not exercised here
      * missing optional imports (torch_geometric, matplotlib, seaborn) are stubbed
    """
    import types
    from patent_image_retrieval_b200 import geoopt_shim
    geoopt_shim.install()
    for name in ["torch_geometric", "torch_geometric.utils", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.cm", "seaborn", "geoopt.optim.radam"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            def _ga(a):
                if a.startswith("__"):
                    raise AttributeError(a)
                return lambda *x, **k: None
            m.__getattr__ = _ga
            sys.modules[name] = m
    sys.path.insert(0, str(REF / "src"))
    import models as ref_models          # flips the default dtype to float64 (src/models.py:248-249)
    import train as ref_train
    ref_models.dropout = 0.0
    out = {}
    torch.manual_seed(11)
    # ---- projection head (models.py:291-318, 481-505, 803-807), fp64 as the reference really runs ----
    for c in (1.0, 0.5):
        tag = str(c).replace(".", "p")
        model = ref_models.FigureOnlyHyperbolicModel(32, 16, hidden_dims=[24], c=c, dropout_rate=0.3).eval()
        x = torch.randn(10, 32, dtype=torch.float32) * 0.5
        x[0] = 0
        with torch.no_grad():
            y = model(x)
        sd = model.state_dict()
        out[f"refshim_head_x_c{tag}"] = x.numpy()
        out[f"refshim_head_y_c{tag}"] = y.numpy()
        for key, v in sd.items():
            out[f"refshim_head_{key}_c{tag}"] = v.numpy()
    # ---- evaluate_retrieval (train.py:3108-3296) incl. multi-positive, out-of-range, sentinels ---------
    c = 2.0
    model = ref_models.HyperbolicEmbeddingModel(32, 16, label_num=60, hidden_dims=[24], c=c).eval()
    X = torch.randn(40, 32, dtype=torch.float32) * 0.5
    label_offsets = {"patents": 0, "medium_cpcs": 45, "big_cpcs": 55, "main_cpcs": 58}
    rng = np.random.default_rng(5)
    f2p = {}
    for i in range(40):
        r = rng.random()
        if r < 0.5:
            f2p[i] = int(rng.integers(0, 45))
        elif r < 0.8:
            f2p[i] = [int(v) for v in rng.choice(50, size=3, replace=False)]     # some >= 45: out of range
        elif r < 0.9:
            f2p[i] = [50, 57]                                                      # no valid positive
    eval_indices = list(range(40))
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m_ap = ref_train.evaluate_retrieval(model, X, eval_indices, f2p, label_offsets, "cpu", 16)
        empty = ref_train.evaluate_retrieval(model, X, [], f2p, label_offsets, "cpu", 16)
        noff = ref_train.evaluate_retrieval(model, X, eval_indices, f2p, {"cpcs": 3}, "cpu", 16)
    out["refshim_eval_map"] = np.asarray(m_ap)
    out["refshim_eval_empty"] = np.asarray(empty)
    out["refshim_eval_no_offset"] = np.asarray(noff)
    out["refshim_eval_X"] = X.numpy()
    for key, v in model.state_dict().items():
        out[f"refshim_eval_{key}"] = v.numpy()
    out["refshim_eval_f2p_json"] = np.frombuffer(json.dumps({str(k): v for k, v in f2p.items()}).encode(), dtype=np.uint8)
    # ---- in-batch losses (train.py:2291-2336, 1010-1045), autograd through the double loop ---------------
    k = torch.tensor([-0.5], dtype=torch.float64)
    a = pmath_pts(7, 12, 0.25).requires_grad_(True)
    p = pmath_pts(7, 12, 0.25).requires_grad_(True)
    loss = ref_train.hyperbolic_contrastive_loss(a, p, k, temperature=0.07)
    loss.backward()
    out["refshim_hcl_a"], out["refshim_hcl_p"] = a.detach().numpy(), p.detach().numpy()
    out["refshim_hcl_loss"] = loss.detach().numpy()
    out["refshim_hcl_da"], out["refshim_hcl_dp"] = a.grad.numpy(), p.grad.numpy()
    s_, pos, neg = pmath_pts(5, 12, 0.25), pmath_pts(5, 12, 0.25), pmath_pts(15, 12, 0.25)
    out["refshim_s2p_s"], out["refshim_s2p_pos"], out["refshim_s2p_neg"] = s_.numpy(), pos.numpy(), neg.numpy()
    out["refshim_s2p_loss"] = ref_train.sample_to_prototype_loss(s_, pos, neg, 3, k, margin=0.1).numpy()
    torch.set_default_dtype(torch.float32)
    return out


def pmath_pts(n, d, scale):
    from oracle import pmath
    k = torch.tensor(-0.5, dtype=torch.float64)
    return pmath.project(pmath.expmap0(torch.randn(n, d, dtype=torch.float64) * scale, k=k), k=k)


def main():
    rng = np.random.default_rng(1234)
    out = {}

    # ---- reference: cosine + argsort (retrieval.ipynb:368,383) ---------------------------------
    from sklearn.metrics.pairwise import cosine_similarity
    from sklearn.metrics import average_precision_score
    q = rng.standard_normal((12, 64)).astype(np.float32)
    g = rng.standard_normal((300, 64)).astype(np.float32)
    g[7] = 0.0                       # zero row: sklearn leaves it unnormalised
    sim = cosine_similarity(q, g)
    order = np.stack([np.argsort(sim[i])[::-1] for i in range(q.shape[0])])
    out["ref_cos_q"], out["ref_cos_g"], out["ref_cos_sim"], out["ref_cos_order"] = q, g, sim, order

    # ---- reference: notebook metric block (retrieval.ipynb:310-324, 391-443) ----------------------
    n_gallery, n_query = 300, 12
    gallery_names = [f"fig_{i:05d}.png" for i in range(n_gallery)]
    positives_idx = []
    for i in range(n_query):
        m = int(rng.integers(1, 9))
        pos = rng.choice(n_gallery + 20, size=m, replace=False)     # some positives are NOT in the gallery
        positives_idx.append(np.sort(pos))
    ranked_names = [[gallery_names[j] for j in order[i]] for i in range(n_query)]
    positives_sets = [set(f"fig_{j:05d}.png" for j in pos) for pos in positives_idx]
    nbm = run_notebook_metrics(ranked_names, positives_sets)
    for k, v in nbm.items():
        out["ref_nb_" + k] = v
    out["ref_nb_pos_offsets"] = np.cumsum([0] + [len(p) for p in positives_idx]).astype(np.int64)
    out["ref_nb_pos_items"] = np.concatenate(positives_idx).astype(np.int64)

    # ---- reference: sklearn AP (train.py:3285) incl. heavy ties; auxiliary.mean_average_precision ------
    sys.path.insert(0, str(REF / "src"))
    import auxiliary  # noqa: E402  (imports fine: torch/numpy/sklearn only)
    scores = np.round(rng.standard_normal((20, 200)), 1).astype(np.float32)    # rounded -> many ties
    target = (rng.random((20, 200)) < 0.05).astype(np.float32)
    target[:, 0] = 1.0
    out["ref_ap_scores"], out["ref_ap_target"] = scores, target
    out["ref_ap_values"] = np.asarray([average_precision_score(target[i], scores[i]) for i in range(20)])
    preds = torch.from_numpy(scores.T.copy())            # [batch, labels]
    tgts = torch.from_numpy(target.T.copy())
    out["ref_aux_map"] = np.asarray(auxiliary.mean_average_precision(preds, tgts))

    # ---- oracle-minted KATs for the hyperbolic arithmetic (PARITY UNPINNED) --------------------------
    from oracle import pmath, head
    torch.manual_seed(7)
    for c in (1.0, 0.5, 2.0):
        tag = str(c).replace(".", "p")
        u = torch.randn(16, 32, dtype=torch.float64) * 0.2
        u[0] = 0.0                                      # zero row -> clamp_min path
        u[1] *= 40.0                                    # far outside -> project clip
        k = torch.tensor(-c, dtype=torch.float64)
        x = pmath.project(pmath.expmap0(u, k=k), k=k)
        y = pmath.project(pmath.expmap0(torch.randn(16, 32, dtype=torch.float64) * 0.3, k=k), k=k)
        out[f"kat_u_c{tag}"] = u.numpy()
        out[f"kat_x_c{tag}"] = x.numpy()
        out[f"kat_y_c{tag}"] = y.numpy()
        out[f"kat_dist_c{tag}"] = pmath.dist(x[:, None, :], y[None, :, :], k=k).numpy()
        out[f"kat_dist0_c{tag}"] = pmath.dist0(x, k=k).numpy()
        out[f"kat_madd_c{tag}"] = pmath.mobius_add(x, y, k=k).numpy()
        w = torch.randn(24, 32, dtype=torch.float64) * 0.3
        out[f"kat_w_c{tag}"] = w.numpy()
        out[f"kat_matvec_c{tag}"] = pmath.mobius_matvec(w, x, k=k).numpy()
        out[f"kat_tanh_c{tag}"] = pmath.mobius_fn_apply(torch.tanh, x, k=k).numpy()
        b = pmath.expmap0(torch.randn(24, dtype=torch.float64) * 1e-3, k=k)
        out[f"kat_b_c{tag}"] = b.numpy()
        out[f"kat_mlin_c{tag}"] = head.mobius_linear(x, w, b, hyperbolic_input=True, k=k).numpy()
        out[f"kat_mlin_e_c{tag}"] = head.mobius_linear(u, w, b, hyperbolic_input=False, k=k).numpy()

    out.update(reference_code_with_shim())
    np.savez_compressed(HERE / "golden.npz", **out)
    print("wrote", HERE / "golden.npz", {k: v.shape for k, v in out.items() if k.startswith("ref_")})


if __name__ == "__main__":
    main()
