"""CPU check of the compare-exchange network behind the scoring kernel's dense cold start
(csrc/score_topk.cu: sort16_pairs, reg_merge16).  The exchange pairs are hand-typed constants in the kernel source;
this test reads them back from the .cu file and proves, by the 0-1 principle, that they sort every input, and that the
half-cleaner + four bitonic stages of reg_merge16 return the 16 smallest of two sorted runs in order."""
import itertools
import re
from pathlib import Path

import numpy as np

SRC = Path(__file__).resolve().parents[1] / "patent_image_retrieval_b200" / "csrc" / "score_topk.cu"


def _pairs():
    text = SRC.read_text()
    body = text[text.index("void sort16_pairs"):]
    body = body[:body.index("#pragma unroll")]
    arrays = re.findall(r"constexpr int ([AB])\[63\] = \{([^}]*)\}", body)
    got = {name: [int(x) for x in vals.replace("\n", " ").split(",")] for name, vals in arrays}
    assert set(got) == {"A", "B"} and len(got["A"]) == 63 and len(got["B"]) == 63
    return list(zip(got["A"], got["B"]))


def test_sort16_network_sorts_every_input():
    pairs = _pairs()
    assert all(0 <= a < b < 16 for a, b in pairs)
    # 0-1 principle: a comparator network sorts all inputs iff it sorts all 2^16 binary inputs
    x = ((np.arange(1 << 16)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.int8)
    for a, b in pairs:                                    # compare-exchange: afterwards x[a] <= x[b]
        lo, hi = np.minimum(x[:, a], x[:, b]), np.maximum(x[:, a], x[:, b])
        x[:, a], x[:, b] = lo, hi
    assert bool((np.diff(x, axis=1) >= 0).all())


def test_merge16_keeps_the_sixteen_smallest_in_order():
    rng = np.random.default_rng(0)
    cases = [(np.sort(rng.integers(0, 40, 16)), np.sort(rng.integers(0, 40, 16))) for _ in range(2000)]
    # all 0-1 run pairs as well (17 x 17)
    for za, zb in itertools.product(range(17), range(17)):
        cases.append((np.array([0] * za + [1] * (16 - za)), np.array([0] * zb + [1] * (16 - zb))))
    for lst, new in cases:
        m = np.minimum(lst, new[::-1]).copy()             # half-cleaner against the reversed run: bitonic, the 16 smallest
        for j in (8, 4, 2, 1):
            for a in range(16):
                b = a ^ j
                if b > a and m[b] < m[a]:
                    m[a], m[b] = m[b], m[a]
        want = np.sort(np.concatenate([lst, new]))[:16]
        assert np.array_equal(m, want)
