"""ctypes binding of libhypret.so (include/hypret.h).

The product path has no CPU fallback: if the library cannot be built or loaded the
first call raises.  Every wrapper checks the integer status and raises
``RuntimeError(hypret_strerror(rc))``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p, POINTER,
                    Structure)

from . import build as _build

_LIB = None


class PeerRoute(Structure):
    """hypret_peer_route (include/hypret.h)."""
    _fields_ = [("base", c_void_p * 16), ("n_ranks", c_int32), ("me", c_int32), ("ql", c_int64)]


class ScorePlan(Structure):
    _fields_ = [
        ("n_qtiles", c_int32),
        ("n_gtiles", c_int32),
        ("n_lists", c_int32),
        ("grid", c_int32),
        ("stages", c_int32),
        ("resident", c_int32),
        ("smem_bytes", c_int32),
        ("n_full", c_int32),
        ("tail_rows", c_int32),
        ("a", c_int32),
        ("b", c_int32),
        ("l1", c_int32),
        ("rem_rows", c_int32),
        ("rem_g0", c_int32),
        ("m", c_int32),
        ("l2", c_int32),
        ("n_steps", c_int32),
        ("sub", c_int32),
        ("sub_tail", c_int32),
        ("pair", c_int32),
        ("epi_groups", c_int32),
    ]

    def asdict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


# symbol -> (restype, argtypes); must list every function include/hypret.h declares
SIGNATURES = {
    "hypret_strerror": (c_char_p, [c_int]),
    "hypret_version": (c_int, []),
    "hypret_operand_kpad": (c_int64, [c_int]),
    "hypret_project_rows": (c_int, [c_void_p, c_int64, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "hypret_project_rows_cert": (c_int, [c_void_p, c_int64, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "hypret_rerank_cert": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_exact_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_int, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_exact_topk_after": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_int,
                                        c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "hypret_peer_free": (c_int, [c_void_p]),
    "hypret_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "hypret_peer_close": (c_int, [c_void_p]),
    "hypret_peer_copy": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "hypret_project_rows_peers": (c_int, [c_void_p, c_int64, c_int, c_float, c_int, c_void_p, POINTER(c_void_p), c_int,
                                          c_void_p, c_void_p]),
    "hypret_peer_signal": (c_int, [POINTER(c_void_p), c_int, c_uint32, c_void_p]),
    "hypret_peer_wait": (c_int, [c_void_p, c_int, c_uint32, c_void_p, c_void_p]),
    "hypret_cand_select_route": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                                         POINTER(PeerRoute), c_int64, c_void_p]),
    "hypret_kth_smallest_route": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, POINTER(PeerRoute), c_int64,
                                          c_void_p]),
    "hypret_rerank_pruned_route": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int64, c_void_p, POINTER(PeerRoute), c_int64,
                                           c_int64, c_void_p]),
    "hypret_score_plan": (c_int, [c_int64, c_int64, c_int, c_int, c_int, c_int, POINTER(ScorePlan)]),
    "hypret_score_strip": (c_int, [POINTER(ScorePlan), c_int, c_int, POINTER(c_int32)]),
    "hypret_score_topk_bound": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_score_topk": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_rerank": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                              c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "hypret_row_sqnorm64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "hypret_rerank_pruned": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_cand_select": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hypret_kth_smallest": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "hypret_mobius_epilogue": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p]),
    "hypret_pairdist": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "hypret_pairdist_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_int,
                                    c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "hypret_pairdist_ce_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "hypret_pairdist_ce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p,
                                       c_float, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                       c_int64, c_int64, c_void_p]),
    "hypret_gram_kpad": (c_int64, [c_int]),
    "hypret_split3": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "hypret_gram_split": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hypret_gram_dist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                 c_float, c_void_p, c_void_p]),
    "hypret_neg_lse": (c_int, [c_void_p, c_int64, c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p]),
    "hypret_retrieval_metrics": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_int32),
                                         c_int, c_void_p, c_void_p, c_void_p]),
    "hypret_ap_full": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "hypret_pair_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_void_p]),
    "hypret_rank_count": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "hypret_ap_from_counts": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_mobius_gemm": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_mobius_epilogue_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_sgemm_strided": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_cert_merged": (c_int, [c_void_p, c_int64, c_int, c_float, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_flag_compact": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_lse_combine": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "hypret_sum_parts": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "hypret_flash_kpad": (c_int64, [c_int]),
    "hypret_flash_workspace": (c_int64, [c_int64, c_int64, c_int]),
    "hypret_flash_prep": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "hypret_flash_lse": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                 c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "hypret_flash_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float, c_float, c_float,
                                  c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "hypret_rowpair_dist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "hypret_rowpair_dist_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "hypret_hmi_pairs": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_float, c_float, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "hypret_dist0_reg": (c_int, [c_void_p, c_int64, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "hypret_radam_ball_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_float,
                                       c_float, c_float, c_float, c_int, c_void_p]),
    "hypret_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}


def load() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    try:
        # HYPRET_CHECKED=1: the library built with -DHYPRET_CHECKED (device-side asserts on every guarded index)
        path = _build.build(checked=os.environ.get("HYPRET_CHECKED", "0") == "1")
    except Exception as exc:  # no silent fallback: surface the build failure
        raise RuntimeError(f"libhypret.so is not built and cannot be built here: {exc}") from exc
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().hypret_strerror(rc)
        raise RuntimeError(f"{msg.decode() if msg else 'hypret error'} (rc={rc})")
