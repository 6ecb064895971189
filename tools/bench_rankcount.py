"""Matrix-free exact full-ranking metrics at C2 size (10k queries x 300k gallery, D=512): hypret_pair_keys +
hypret_rank_count + AP / metric suite from the counts, against the dense path (pairdist -> ap_full) on a slice."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from patent_image_retrieval_b200 import evaluation, ops, synth  # noqa: E402
from patent_image_retrieval_b200.dist import full_ranking_ap  # noqa: E402

Q, N, D, P = (int(x) for x in (sys.argv[1:5] + [10000, 300000, 512, 8][len(sys.argv) - 1:]))
gal, _, _ = ops.project_rows(synth.gaussian_features(N, D, seed=0, device="cuda"), 1.0, want_operand=False)
qry, _, _ = ops.project_rows(synth.gaussian_features(Q, D, seed=1, device="cuda"), 1.0, want_operand=False)
g = torch.Generator().manual_seed(2)
items = torch.randint(0, N, (Q * P,), generator=g).cuda()
off = torch.arange(0, Q * P + 1, P, dtype=torch.int64).cuda()
out = {}
for name, fn in (("full_ranking_ap (sklearn ties)", lambda: full_ranking_ap(qry, gal, off, items, grouped_ties=True)),
                 ("full_ranking_metrics (notebook suite)", lambda: evaluation.full_ranking_metrics(qry, gal, off, items, metric="hyperbolic"))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    out[name] = {"seconds": time.perf_counter() - t0}
out["fp32_fma_pairs"] = float(Q) * N * D
out["tflops_2ops_per_pair_element"] = 2.0 * Q * N * D / out["full_ranking_ap (sklearn ties)"]["seconds"] / 1e12
# dense path on a 512-query slice for comparison (the full [Q,N] matrix would be 12 GB)
qs = qry[:512]
t0 = time.perf_counter()
d = ops.pairdist(qs, gal, 1.0)
dense = ops.ap_full(-d, off[:513], items[:512 * P], grouped_ties=True)
torch.cuda.synchronize()
out["dense_512_queries_seconds"] = time.perf_counter() - t0
fused = full_ranking_ap(qs, gal, off[:513], items[:512 * P], grouped_ties=True)
out["fused_equals_dense_on_slice"] = bool(torch.allclose(fused[1], dense[1], rtol=1e-13, atol=1e-16))
print(json.dumps({"Q": Q, "N": N, "D": D, "positives_per_query": P, **out}))
