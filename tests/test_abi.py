"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol declared in
include/hypret.h.  No compute entry point is called (no GPU here) -- only the pure-host ones."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "hypret.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hypret_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from patent_image_retrieval_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 7
    for s in syms:
        assert hasattr(lib, s), f"libhypret.so does not export {s}"
        assert s in _lib.SIGNATURES, f"ctypes binding missing for {s}"
    assert sorted(_lib.SIGNATURES) == syms


def test_host_only_entry_points():
    from patent_image_retrieval_b200 import _lib, ops
    lib = _lib.load()
    assert lib.hypret_version() >= 100
    assert lib.hypret_strerror(0) == b"ok"
    assert b"invalid" in lib.hypret_strerror(-1)
    assert ops.operand_kpad(512) == 528 and ops.operand_kpad(768) == 784 and ops.operand_kpad(100) == 144


SCHED_CASES = [
    # Q, N, d, kprime, max_ctas
    (10000, 300000, 512, 16, 0),        # C2: 79 query tiles on 148 CTAs -> phase 1 + phase 2
    (1000, 10000, 2048, 16, 0),         # C1
    (100000, 1000000, 768, 32, 0),      # C3: full waves + tail
    (10000, 10_000_000, 512, 16, 0),    # C4 on one GPU
    (10000, 1_250_000, 512, 16, 0),     # C4 shard on 8 GPUs
    (1, 1, 128, 4, 0),
    (129, 257, 768, 32, 0),
    (900, 5000, 128, 8, 5),             # forced small grid: full waves + both phases
    (700, 3000, 64, 16, 7),
    (128 * 6, 256 * 9, 64, 16, 6),      # T == P exactly: only full waves
    (128 * 3, 256 * 40, 64, 16, 6),     # b == 0: phase 1 only
]


@pytest.mark.parametrize("Q,N,d,kp,cap", SCHED_CASES)
def test_strip_schedule_covers_every_tile_once(Q, N, d, kp, cap):
    """Every (query tile, gallery tile) pair is visited by exactly one strip, list slots of a
    query tile are distinct and < n_lists, a CTA has at most one strip per step, and the
    busiest CTA is within a few percent of the ideal share."""
    from patent_image_retrieval_b200 import ops
    p = ops.score_plan(Q, N, d, kp, cap)
    pair = p["pair"]
    T, G, P = p["n_qtiles"], p["n_gtiles"], p["grid"] // pair      # P scheduling units (CTAs or CTA pairs)
    TR = -(-T // pair)                                              # scheduling rows (query tiles or tile pairs)
    assert T == -(-Q // 128) and G == -(-N // 256)
    assert pair in (1, 2) and p["grid"] % pair == 0
    assert 2 <= p["stages"] <= 8 and p["smem_bytes"] <= 232448
    strips = ops.score_strips(Q, N, d, kp, cap)
    seen = {}
    per_cta = {}
    steps = set()
    for cta, step, qt, g0, g1, slot in strips:
        assert 0 <= qt < T and qt % pair == 0 and 0 <= g0 < g1 <= G and 0 <= slot < p["n_lists"] and 0 <= cta < P
        assert (cta, step) not in steps
        steps.add((cta, step))
        assert (qt, slot) not in seen
        seen[(qt, slot)] = (g0, g1)
        per_cta[cta] = per_cta.get(cta, 0) + (g1 - g0)
    if TR * G <= 200000:                  # exhaustive coverage check on the small cases
        cover = {}
        for (qt, slot), (g0, g1) in seen.items():
            for g in range(g0, g1):
                assert (qt, g) not in cover
                cover[(qt, g)] = 1
        assert len(cover) == TR * G
    else:                                 # interval check on the big ones
        by_qt = {}
        for (qt, slot), (g0, g1) in seen.items():
            by_qt.setdefault(qt, []).append((g0, g1))
        assert len(by_qt) == TR
        for qt, iv in by_qt.items():
            iv.sort()
            assert iv[0][0] == 0 and iv[-1][1] == G
            assert all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
    ideal = TR * G / P
    if ideal >= 64:
        assert max(per_cta.values()) <= 1.03 * ideal + 2
    # resident query tile only while the gallery ring stays >= 4 stages deep
    assert p["resident"] == 0 or p["stages"] >= 4


@pytest.mark.parametrize("Q,N,d,kp,min_lists", [(64, 20000, 768, 64, 4), (2000, 60000, 768, 64, 4),
                                                 (100000, 1000000, 768, 64, 4), (300, 3000, 128, 64, 3)])
def test_wide_topk_schedule_gives_every_query_enough_lists(Q, N, d, kp, min_lists):
    """Wide top-k (k > kprime): every strip is cut into sub-strips so that each query tile owns at least
    ``min_lists`` disjoint candidate lists which together still cover the gallery exactly once."""
    from patent_image_retrieval_b200 import ops
    p = ops.score_plan(Q, N, d, kp, 0, min_lists)
    pair, G = p["pair"], p["n_gtiles"]
    by_qt = {}
    slots = set()
    for cta, step, qt, g0, g1, slot in ops.score_strips(Q, N, d, kp, 0, min_lists):
        assert (qt, slot) not in slots and slot < p["n_lists"]
        slots.add((qt, slot))
        by_qt.setdefault(qt, []).append((g0, g1))
    assert len(by_qt) == -(-p["n_qtiles"] // pair)
    for qt, iv in by_qt.items():
        iv.sort()
        assert iv[0][0] == 0 and iv[-1][1] == G and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
        assert len(iv) >= min(min_lists, G)
    assert p["smem_bytes"] <= 232448 and p["stages"] >= 2


def test_c2_plan_numbers():
    from patent_image_retrieval_b200 import ops
    p = ops.score_plan(10000, 300000, 512, 16)
    # 79 query tiles -> 40 tile pairs on 74 CTA pairs: 1 strip each + 34 rows with a second strip, rest in phase 2
    assert (p["grid"], p["pair"], p["n_full"], p["tail_rows"], p["a"], p["b"]) == (148, 2, 0, 40, 1, 34)
    assert p["l1"] == 586 and p["rem_rows"] == 6 and p["m"] == 12 and p["l2"] == 49
    # k' = 16: lists in registers, two epilogue warpgroups -> two list slots per strip (13 strips per query at most)
    assert p["epi_groups"] == 2 and p["n_lists"] == 26 and p["resident"] == 1 and p["stages"] == 4
    assert ops.score_plan(10000, 300000, 512, 26)["epi_groups"] == 1
    with pytest.raises(RuntimeError):
        ops.score_plan(10, 10, 512, 65)          # kprime > 64


def test_cpu_tensor_is_rejected():
    import torch
    from patent_image_retrieval_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.project_rows(torch.zeros(4, 64))


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """sizeof / offsetof of the two structs that cross the ABI, as gcc lays them out from include/hypret.h, against
    the ctypes mirrors in _lib.py (a silent mismatch would corrupt the routing table or the score plan)."""
    import shutil
    import subprocess
    from patent_image_retrieval_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields_route = ["base", "n_ranks", "me", "ql"]
    fields_plan = [name for name, _ in _lib.ScorePlan._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "hypret.h"}"', "int main(void) {",
           '  printf("%zu\\n", sizeof(hypret_peer_route));']
    src += [f'  printf("%zu\\n", offsetof(hypret_peer_route, {f}));' for f in fields_route]
    src += ['  printf("%zu\\n", sizeof(hypret_score_plan_t));']
    src += [f'  printf("%zu\\n", offsetof(hypret_score_plan_t, {f}));' for f in fields_plan]
    src += ["  return 0;", "}"]
    c_file, exe = tmp_path / "layout.c", tmp_path / "layout"
    c_file.write_text("\n".join(src))
    subprocess.run([gcc, "-std=c99", "-o", str(exe), str(c_file)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.PeerRoute)] + [getattr(_lib.PeerRoute, f).offset for f in fields_route]
    want += [ctypes.sizeof(_lib.ScorePlan)] + [getattr(_lib.ScorePlan, f).offset for f in fields_plan]
    assert got == want


def test_default_list_sizes_and_certificate_candidates():
    """k <= 10 keeps the 16-slot register lists and certifies with k+14 candidates (the bound of the 24th best at k=10);
    11 <= k <= 26 uses k+14 slots (<= 32) directly; beyond that the 64-slot wide lists."""
    from patent_image_retrieval_b200.retrieval import default_kbound, default_kprime
    assert [default_kprime(k) for k in (1, 5, 10, 11, 18, 26, 27, 100, 128)] == [16, 16, 16, 25, 32, 32, 64, 64, 64]
    assert [default_kbound(k, default_kprime(k)) for k in (1, 5, 10, 11, 26, 100)] == [16, 19, 24, 25, 32, 64]
    for k in range(1, 27):
        kp, kb = default_kprime(k), default_kbound(k, default_kprime(k))
        assert k <= kp <= 32 and kp <= kb <= 32
    with pytest.raises(ValueError):
        default_kprime(129)


def test_argument_validation_happens_before_any_device_work():
    """Every entry point validates shapes / enums / pointers first and returns HYPRET_EINVAL (-1) or HYPRET_EUNSUPPORTED
    (-2) without touching a device, so the error behaviour of the C ABI is testable on a CPU-only box; an empty problem
    is HYPRET_OK.  (Valid arguments would go on to check_device() and fail with a CUDA error here: not called.)"""
    from patent_image_retrieval_b200 import _lib
    lib = _lib.load()
    P = 0x1000          # a non-NULL, 16-byte aligned dummy: validation never dereferences it
    EINVAL, EUNSUP, OK = -1, -2, 0
    # exact top-k: k > 32, bad metric, missing pointers; Q == 0 is an empty problem
    assert lib.hypret_exact_topk(P, P, P, 4, 100, 64, 1.0, 1, 33, 0, P, P, P, None, P, P, None) == EINVAL
    assert lib.hypret_exact_topk(P, P, P, 4, 100, 64, 1.0, 7, 10, 0, P, P, P, None, P, P, None) == EINVAL
    assert lib.hypret_exact_topk(P, P, P, 4, 100, 62, 1.0, 1, 10, 0, P, P, P, None, P, P, None) == EINVAL
    assert lib.hypret_exact_topk(P, P, None, 4, 100, 64, 1.0, 1, 10, 0, P, P, P, None, P, P, None) == EINVAL
    assert lib.hypret_exact_topk(P, P, P, 0, 100, 64, 1.0, 1, 10, 0, P, P, P, None, P, P, None) == OK
    assert lib.hypret_exact_topk(P, P, P, 4, 2 ** 31, 64, 1.0, 1, 10, 0, P, P, P, None, P, P, None) == EUNSUP
    # scoring with a shared bound: kbound < kprime, kbound > 32, 24-slot lists, no threshold workspace
    args = lambda kp, kb, ws: (P, 10, P, 10, 64, kp, kb, 1, 0, 0, P, P, ws, None, None, None)
    assert lib.hypret_score_topk_bound(*args(16, 8, P)) == EINVAL
    assert lib.hypret_score_topk_bound(*args(16, 40, P)) == EINVAL
    assert lib.hypret_score_topk_bound(*args(24, 32, P)) == EINVAL
    assert lib.hypret_score_topk_bound(*args(16, 24, None)) == EINVAL
    # certified rerank: k' > 32 is the wide path (unsupported with a certificate); ksel outside (k', 32]
    cert = lambda kp, ksel, k: (P, P, 4, 100, 64, 1.0, 1, P, P, None, 2, kp, ksel, k, 0, P, P, None, P, P, P, P, P, None,
                                None, None)
    assert lib.hypret_rerank_cert(*cert(64, 0, 10)) == EUNSUP
    assert lib.hypret_rerank_cert(*cert(16, 12, 10)) == EINVAL
    assert lib.hypret_rerank_cert(*cert(16, 40, 10)) == EINVAL
    # fused MobiusLinear layer: n_out % 16, n_out > 256, d_in % 4, curvature, n_project
    mg = lambda d_in, n_out, c, npj: (P, P, 8, d_in, n_out, None, None, c, 0, npj, None, P, None, None, None)
    assert lib.hypret_mobius_gemm(*mg(64, 24, 1.0, 1)) == EINVAL
    assert lib.hypret_mobius_gemm(*mg(64, 272, 1.0, 1)) == EINVAL
    assert lib.hypret_mobius_gemm(*mg(62, 32, 1.0, 1)) == EINVAL
    assert lib.hypret_mobius_gemm(*mg(64, 32, 0.0, 1)) == EINVAL
    assert lib.hypret_mobius_gemm(*mg(64, 32, 1.0, 3)) == EINVAL
    assert lib.hypret_mobius_gemm(P, P, 0, 64, 32, None, None, 1.0, 0, 1, None, P, None, None, None) == OK
    # its backward: gbias without bias, gxn without xsq
    assert lib.hypret_mobius_epilogue_bwd(P, 8, 32, None, None, 1.0, 0, 1, P, P, P, None, None) == EINVAL
    assert lib.hypret_mobius_epilogue_bwd(P, 8, 32, None, P, 1.0, 0, 1, P, P, None, P, None) == EINVAL
    assert lib.hypret_sgemm_strided(P, 1, 1, P, 1, 1, 4, 4, 4, P, None, P, None) == EINVAL      # row_scale without addend
    assert lib.hypret_sgemm_strided(P, 1, 1, P, 1, 1, 0, 4, 4, None, None, P, None) == OK
    # flash train_hyp: D > 128 / D % 16 belong to the generic path
    fl = lambda d: (P, P, P, P, P, P, 8, 8, d, 1.0, 10.0, P, P, None)
    assert lib.hypret_flash_lse(*fl(256)) == EINVAL
    assert lib.hypret_flash_lse(*fl(72)) == EINVAL
    assert lib.hypret_flag_compact(P, -1, P, P, P, None) == EINVAL
    assert lib.hypret_cert_merged(P, 4, 64, 1.0, 9, P, P, 10, P, P, P, P, None, None) == EINVAL
