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


def test_score_plan_is_balanced_and_covers_gallery():
    from patent_image_retrieval_b200 import ops
    for (Q, N, d, kp) in [(10000, 300000, 512, 16), (1000, 10000, 2048, 16), (10000, 10_000_000, 512, 16),
                          (1, 1, 128, 4), (129, 257, 768, 32)]:
        p = ops.score_plan(Q, N, d, kp)
        assert p["n_qtiles"] == -(-Q // 128) and p["n_gtiles"] == -(-N // 256)
        assert p["n_splits"] * p["tiles_per_split"] >= p["n_gtiles"]
        assert (p["n_splits"] - 1) * p["tiles_per_split"] < p["n_gtiles"]       # no empty split
        assert 2 <= p["stages"] <= 8 and p["smem_bytes"] <= 232448
        assert p["resident"] == (1 if d <= 512 else 0)
    p = ops.score_plan(10000, 300000, 512, 16)
    items = p["n_qtiles"] * p["n_splits"]
    waves = -(-items // 148)
    assert items / (waves * 148) > 0.9           # wave quantisation loss < 10 %
    with pytest.raises(RuntimeError):
        ops.score_plan(10, 10, 512, 64)          # kprime > 32


def test_cpu_tensor_is_rejected():
    import torch
    from patent_image_retrieval_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.project_rows(torch.zeros(4, 64))
