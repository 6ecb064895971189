import sys, torch
sys.path.insert(0, ".")
from patent_image_retrieval_b200 import ops, synth
from patent_image_retrieval_b200.geoopt_shim import pmath
def tm(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for d in (128, 256):
    n, c = 8192, 0.5
    k = torch.tensor([-c])
    mu = synth.gaussian_features(n, d, seed=2, scale=1.0, device="cuda")
    a = pmath.project(pmath.expmap0(mu + 0.1 * synth.gaussian_features(n, d, seed=3, scale=1.0, device="cuda"), k=k), k=k)
    p = pmath.project(pmath.expmap0(mu + 0.1 * synth.gaussian_features(n, d, seed=4, scale=1.0, device="cuda"), k=k), k=k)
    dm, rl, cl = ops.pairdist_ce_fwd(a, p, c, 1 / 0.07, False)
    asq, psq = ops.row_sqnorm(a), ops.row_sqnorm(p)
    w, rs, cs = ops.pairdist_ce_bwd(dm, asq, psq, c, rl, None, 1 / 0.07, 1.0, 0.0)
    print(d, "gram_dist", round(tm(lambda: ops.gram_dist(a, p, c)), 3),
          "ce_fwd tc", round(tm(lambda: ops.pairdist_ce_fwd(a, p, c, 1 / 0.07, False, tensor_cores=True)), 3),
          "ce_fwd cuda-core", round(tm(lambda: ops.pairdist_ce_fwd(a, p, c, 1 / 0.07, False, tensor_cores=False)), 3),
          "ce_bwd kernel", round(tm(lambda: ops.pairdist_ce_bwd(dm, asq, psq, c, rl, None, 1 / 0.07, 1.0, 0.0)), 3),
          "W@P", round(tm(lambda: w @ p), 3), "W.t@A", round(tm(lambda: w.t() @ a), 3))
    w3, rs3, cs3 = ops.pairdist_ce_bwd(dm, asq, psq, c, rl, None, 1 / 0.07, 1.0, 0.0, split=True)
    wp, wta = ops.split_products(w3, a, p)
    ref_wp, ref_wta = w.double() @ p.double(), w.double().t() @ a.double()
    print(d, "ce_bwd kernel (3 bf16 planes)", round(tm(lambda: ops.pairdist_ce_bwd(dm, asq, psq, c, rl, None, 1 / 0.07, 1.0, 0.0, split=True)), 3),
          "split products (both)", round(tm(lambda: ops.split_products(w3, a, p)), 3),
          "rel err split", float((wp - ref_wp).abs().max() / ref_wp.abs().max()), float((wta - ref_wta).abs().max() / ref_wta.abs().max()),
          "rel err fp32 gemm", float(((w @ p) - ref_wp).abs().max() / ref_wp.abs().max()))
