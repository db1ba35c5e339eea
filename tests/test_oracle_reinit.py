"""CPU pins of the pieces `oracle/reinit_ref.py` restates from packages that are absent here (scikit-image): the Lab conversion of
the patch-similarity filter (reference sam_pt.py:645 `skimage.color.rgb2lab`).  Parity with skimage itself stays unpinned
(DESIGN.md §2); these tests pin the restatement to the published sRGB / D65 known answers and, loosely, to OpenCV's independent
implementation."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reinit_ref import rgb2lab  # noqa: E402


@pytest.mark.parametrize("rgb,lab", [((255, 0, 0), (53.2408, 80.0925, 67.2032)), ((0, 255, 0), (87.7347, -86.1827, 83.1793)),
                                     ((0, 0, 255), (32.2970, 79.1875, -107.8602)), ((255, 255, 255), (100.0, 0.0, 0.0)),
                                     ((0, 0, 0), (0.0, 0.0, 0.0)), ((128, 128, 128), (53.5850, 0.0, 0.0))])
def test_rgb2lab_known_answers(rgb, lab):
    """CIE L*a*b* of the sRGB primaries / white / black / mid grey under D65, 2-degree observer (published values; the 1e-2 slack
    covers the rounding of the 6-digit matrix and white point skimage uses)."""
    got = rgb2lab(np.array([[rgb]], dtype=np.uint8))[0, 0]
    assert np.abs(got - np.array(lab)).max() < 2e-2, (rgb, got)


def test_rgb2lab_against_opencv():
    """OpenCV's float RGB->Lab path is an independent implementation of the same conversion; it quantises through look-up tables, so
    the agreement is loose (0.5 on the 0..100 scale) -- enough to catch a wrong matrix, gamma branch, white point or channel order."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    a = rgb2lab(img)
    b = cv2.cvtColor(img.astype(np.float32) / 255.0, cv2.COLOR_RGB2LAB).astype(np.float64)
    assert np.abs(a - b).max() < 0.5
    assert np.abs(a - b).mean() < 0.1
