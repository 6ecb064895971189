"""CPU oracle for the hyperbolic-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (Alvarodelamaza/patent-image-retrieval) ships no
tests, golden vectors or fixtures for this path, and the library that owns its
hyperbolic arithmetic (``geoopt``, un-pinned, imported at src/models.py:5,7 and
src/train.py:15,18,37) is neither vendored under /root/reference nor
installable here.  This package therefore *restates* geoopt's published
stereographic-model formulas (geoopt/manifolds/stereographic/math.py, most
plausibly v0.5.0) and drives them in the reference's own call order.  The parts
of the path that ARE runnable here (sklearn ``cosine_similarity`` /
``average_precision_score``, ``np.argsort``, the reference's
``auxiliary.mean_average_precision`` and the metric helpers of
notebooks/retrieval.ipynb) are pinned by golden vectors generated from the
reference itself -- see tests/golden/make_golden.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
path (``patent_image_retrieval_b200``) never does, and fails loudly when its
CUDA library is missing.
"""
from . import pmath, head, retrieval, contrastive  # noqa: F401
