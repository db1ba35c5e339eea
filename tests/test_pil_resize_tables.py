"""CPU: the coefficient tables + integer passes of the resize kernel are bit-exact against PIL itself (the library the
reference's ResizeLongestSide.apply_image calls), for the up-scaling and antialiased down-scaling cases of BASELINE configs."""
import numpy as np
import pytest
from PIL import Image

from sampt_b200 import pil_resize


@pytest.mark.parametrize("hw,out", [((240, 320), (768, 1024)), ((480, 854), (576, 1024)), ((270, 480), (576, 1024)),
                                     ((1080, 1920), (576, 1024)), ((37, 53), (90, 41))])
def test_tables_bit_exact_vs_pil(hw, out):
    rng = np.random.default_rng(hw[0] * 7 + out[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((out[1], out[0]), resample=Image.BILINEAR))
    got = pil_resize.resize_reference_numpy(img, out)
    assert np.array_equal(got, ref)
