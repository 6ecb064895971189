"""patent_image_retrieval_b200 -- B200-native hot path of Alvarodelamaza/patent-image-retrieval.

Poincare-ball projection -> query x gallery hyperbolic / cosine scoring -> top-k ->
recall@k / mAP, as hand-written sm_100a CUDA (tcgen05 / TMEM / TMA) behind a C ABI
(include/hypret.h), with the reference's Python signatures on top.

(The directory ``patent-image-retrieval_b200`` at the repo root is a symlink to this
package: a hyphen is not importable.)
"""
from . import ops, synth  # noqa: F401
from .retrieval import GalleryIndex, SearchPipeline, StageEvents, default_kprime  # noqa: F401

__all__ = ["ops", "synth", "GalleryIndex", "SearchPipeline", "StageEvents", "default_kprime"]
