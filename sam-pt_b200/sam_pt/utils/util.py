"""Subset of the reference's sam_pt/utils/util.py that the hot path writes into its outputs."""
from enum import IntEnum


class PointVisibilityType(IntEnum):
    """Codes stored in `visibilities` (reference sam_pt/utils/util.py:267-282; written at sam_pt.py:656,674,687-690)."""
    VISIBLE = 1
    INVISIBLE = 0
    REINIT_FAILED = -1
    OUTSIDE_FRAME = -2
    PATCH_NON_SIMILAR = -3
    REJECTED_AFTER_PATCH_WAS_NON_SIMILAR = -4
