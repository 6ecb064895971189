"""In-tree build of libhypret.so (hand-written sm_100a CUDA behind a C ABI).

``python -m patent_image_retrieval_b200.build`` or ``build()``.  nvcc cross-compiles
for sm_100a without a GPU; the resulting ``libhypret.so`` sits next to this file so
that it travels to the GPU box with the source tree (it is git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libhypret.so"
STAMP = PKG_DIR / ".libhypret.stamp"
OBJ_DIR = PKG_DIR / "build"
# the CHECKED build: the same sources with -DHYPRET_CHECKED, which turns every HYPRET_CHECK(...) in the kernels into a
# device-side assert on the index about to be used (compute-sanitizer is closed on this GPU pool; tests/test_gpu_checked.py
# runs the ragged-shape cases through this library: an out-of-range index traps instead of corrupting memory)
CHECKED_LIB_PATH = PKG_DIR / "libhypret_checked.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    # no --use_fast_math: the projection and rerank kernels are parity-critical
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_fresh(checked: bool = False) -> bool:
    lib, stamp = (CHECKED_LIB_PATH, PKG_DIR / ".libhypret_checked.stamp") if checked else (LIB_PATH, STAMP)
    return lib.exists() and stamp.exists() and stamp.read_text().strip() == _fingerprint("checked" if checked else "")


def _compile_one(nvcc: str, src: Path, obj: Path, verbose: bool, extra=()) -> str:
    cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "-shared"], *extra, "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name} ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    return res.stderr


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> Path:
    """One object per ``csrc/*.cu`` (compiled in parallel, recompiled only when that source, a header or the
    flags changed), linked into ``libhypret.so`` (``checked``: ``libhypret_checked.so``, device-side index asserts on)."""
    lib_path = CHECKED_LIB_PATH if checked else LIB_PATH
    stamp = PKG_DIR / ".libhypret_checked.stamp" if checked else STAMP
    obj_dir = PKG_DIR / "build_checked" if checked else OBJ_DIR
    extra = ("-DHYPRET_CHECKED",) if checked else ()
    if not force and is_fresh(checked):
        return lib_path
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libhypret.so (no CPU fallback exists)")
    from concurrent.futures import ThreadPoolExecutor
    obj_dir.mkdir(exist_ok=True)
    shared = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        shared.update(f.read_bytes())
    shared.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    jobs = []
    for src in sources():
        obj, tag = obj_dir / (src.stem + ".o"), obj_dir / (src.stem + ".tag")
        want = hashlib.sha256(shared.digest() + src.read_bytes()).hexdigest()
        if force or not obj.exists() or not tag.exists() or tag.read_text() != want:
            jobs.append((src, obj, tag, want))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        logs = list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], verbose, extra), jobs))
    for (src, obj, tag, want), log in zip(jobs, logs):
        tag.write_text(want)
        if verbose:
            print(f"== {src.name}\n{log}", file=sys.stderr)
    objs = [str(obj_dir / (s.stem + ".o")) for s in sources()]
    res = subprocess.run([nvcc, "-shared", "-o", str(lib_path), *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    stamp.write_text(_fingerprint("checked" if checked else ""))
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
