"""``skimage.metrics`` stand-in: only ``structural_similarity`` with the arguments the reference uses
(src/dup/refine.py:52: ``structural_similarity(arr_a, arr_b, data_range=1.0)``)."""
from oracle.ref_py import structural_similarity  # noqa: F401

__all__ = ["structural_similarity"]
