"""SASS evidence that the hot kernels run on the Blackwell paths: for every kernel of libhypret.so that issues tcgen05 /
TMA instructions, the per-mnemonic counts and the instruction lines themselves (UTCHMMA = tcgen05.mma kind::f16,
LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops).
    python tools/sass_summary.py            # writes profiles/sass_<kernel>.txt (cuobjdump -sass, no GPU needed)"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "patent_image_retrieval_b200" / "libhypret.so"
KEEP = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCBAR|UTCATOMSWS)\b")
COUNT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCBAR|UTCATOMSWS|SYNCS|MUFU\.\w+|HMMA|DFMA|"
                   r"FFMA|HFMA2|REDUX|ATOMS|ATOMG|RED|LDG\.\w+|STG\.\w+|LDS\.\w*|STS\.\w*|SHFL\.\w+)")
NAMES = {"score_topk_kernel": "score_topk", "gram_dist_kernel": "gram_dist", "flash_tile_kernel": "flash_tile",
         "mobius_gemm_kernel": "mobius_gemm"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    by_file = collections.defaultdict(list)
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        for key, short in NAMES.items():
            if key in mangled and KEEP.search(f):
                by_file[short].append((mangled, f))
    demangle = lambda m: subprocess.run(["c++filt", m], capture_output=True, text=True).stdout.strip() or m
    for short, items in by_file.items():
        out = [f"# cuobjdump -sass {LIB.name} (sm_100a), kernels matching '{short}': tcgen05 / TMA instruction lines",
               "# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load),",
               "# UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = tcgen05.alloc/dealloc, UTMAPF = TMA prefetch", ""]
        for mangled, f in items:
            lines = f.split("\n")
            n_instr = sum(1 for ln in lines if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", ln))
            counts = collections.Counter(m.group(1) for m in COUNT.finditer(f))
            out.append(f"## {demangle(mangled)[:200]}")
            out.append(f"   instructions: {n_instr}; " + ", ".join(f"{k} {v}" for k, v in sorted(counts.items())))
            for ln in lines:
                if KEEP.search(ln):
                    out.append("   " + re.sub(r"\s+", " ", ln.strip())[:160])
            out.append("")
        (ROOT / "profiles" / f"sass_{short}.txt").write_text("\n".join(out))
        print(f"profiles/sass_{short}.txt: {len(items)} kernels")


if __name__ == "__main__":
    main()
