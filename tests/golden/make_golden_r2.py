"""Generate tests/golden/golden_r2.npz (round 2 additions).  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden_r2.py

``refshim2_*``: the REFERENCE'S OWN ``models.py`` (imported from /root/reference/src, geoopt replaced by this repo's shim
as in make_golden.py) -- ``HyperbolicEmbeddingModel.calculate_hierarchical_loss`` (src/models.py:550-604, with
``_hmi_insideness`` / ``_hmi_disjointedness`` :630-674) and ``calculate_reg_loss`` (:606-628): loss values and the
autograd gradients with respect to ``label_emb`` / the figure embeddings, plus the reference model's state-dict key list
(the shim's PoincareBall is an nn.Module with geoopt's ``isp_c`` parameter, so the keys are geoopt-shaped).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden  # noqa: E402  (sets up sys.path; provides the reference import with the shim)


def main():
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
    sys.path.insert(0, str(make_golden.REF / "src"))
    import models as ref_models          # flips the default dtype to float64 (src/models.py:248-249)
    ref_models.dropout = 0.0
    out = {}
    torch.manual_seed(23)
    c = 0.5
    model = ref_models.HyperbolicEmbeddingModel(32, 16, label_num=60, hidden_dims=[24], c=c)
    out["refshim2_state_dict_keys"] = np.frombuffer("\n".join(sorted(model.state_dict().keys())).encode(), dtype=np.uint8)
    with torch.no_grad():
        # spread the labels over the ball: some near the origin, some far out, three beyond the project clip
        scale = torch.linspace(0.3, 9.0, 60, dtype=torch.float64)[:, None]
        pts = model.label_emb.data * scale
        pts[57:] *= 40.0
        model.label_emb.data.copy_(pts)
    rng = np.random.default_rng(3)
    imp = torch.from_numpy(rng.integers(0, 60, size=(48, 2)).astype(np.int64))
    exc = torch.from_numpy(rng.integers(0, 60, size=(36, 2)).astype(np.int64))
    figs = (torch.randn(20, 16, dtype=torch.float64) * torch.linspace(0.05, 0.6, 20, dtype=torch.float64)[:, None])
    figs[17:] *= 30.0
    figs = model.ball.projx(figs).requires_grad_(True)
    inside, disjoint = model.calculate_hierarchical_loss(imp, exc)
    label_reg, instance_reg = model.calculate_reg_loss(figs)
    g_in, = torch.autograd.grad(inside, model.label_emb, retain_graph=True)
    g_dj, = torch.autograd.grad(disjoint, model.label_emb, retain_graph=True)
    g_lr, = torch.autograd.grad(label_reg, model.label_emb, retain_graph=True)
    g_ir, = torch.autograd.grad(instance_reg, figs)
    out.update({"refshim2_c": np.asarray(c), "refshim2_label_emb": model.label_emb.detach().numpy(),
                "refshim2_imp": imp.numpy(), "refshim2_exc": exc.numpy(), "refshim2_figs": figs.detach().numpy(),
                "refshim2_inside": inside.detach().numpy(), "refshim2_disjoint": disjoint.detach().numpy(),
                "refshim2_label_reg": label_reg.detach().numpy(), "refshim2_instance_reg": instance_reg.detach().numpy(),
                "refshim2_inside_grad": g_in.numpy(), "refshim2_disjoint_grad": g_dj.numpy(),
                "refshim2_label_reg_grad": g_lr.numpy(), "refshim2_instance_reg_grad": g_ir.numpy(),
                "refshim2_insideness": model._hmi_insideness(model.label_emb[imp[:, 0]], model.label_emb[imp[:, 1]]).detach().numpy(),
                "refshim2_disjointedness": model._hmi_disjointedness(model.label_emb[exc[:, 0]], model.label_emb[exc[:, 1]]).detach().numpy()})
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(HERE / "golden_r2.npz", **out)
    print("wrote", HERE / "golden_r2.npz", {k: (v.shape, float(v) if v.ndim == 0 else None) for k, v in out.items()})


if __name__ == "__main__":
    main()
