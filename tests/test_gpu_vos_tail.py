"""GPU parity of the fused VOS harness tail (reference sam_pt/vos_eval/eval.py:304-355) against the torch
restatement: uint8 index masks.  Index work: exact, except at pixels where two channel probabilities are within float32
rounding of each other (the oracle's `softmax` uses the CPU's vectorised exp, the kernel `expf`); the synthetic logits carry a
margin so that no such pixel exists, and the test asserts exact equality."""
import pytest
import torch

from oracle import vos_tail_ref
from sam_pt.vos_eval.postprocess import vos_index_masks

pytestmark = pytest.mark.gpu


def _case(M, T, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    logits = []
    for m in range(M):
        lo = torch.randn((T, 1, H // 8 + 1, W // 8 + 1), generator=g) * 6
        l = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[:, 0]
        l = l + 0.37 * (m + 1)                                  # break exact ties between objects
        logits.append(l)
    logits[-1][T - 1] = -float("inf")                            # an object whose points were all invisible on a frame
    gt = [(torch.rand((H * 2, W * 2), generator=g) > 0.6).float() for _ in range(M)]
    ti = [int(torch.randint(0, T - 1, (1,), generator=g)) for _ in range(M)]
    ti[0] = 0
    return logits, gt, ti


@pytest.mark.parametrize("M,T,H,W,resize,flip", [(1, 4, 48, 64, False, False), (3, 6, 60, 107, False, True),
                                                 (4, 5, 48, 85, True, False), (7, 3, 30, 54, True, True)])
def test_index_masks_match_reference_tail(M, T, H, W, resize, flip):
    logits, gt, ti = _case(M, T, H, W, seed=M * 7 + T)
    out_shape = (H * 2 + 3, W * 2 + 1) if resize else None
    ref = vos_tail_ref.vos_index_masks(logits, gt, ti, (H, W), out_shape, resize, flip)
    got = vos_index_masks([l.cuda() for l in logits], gt, ti, (H, W), out_shape, resize, flip).cpu()
    assert got.dtype == torch.uint8 and got.shape == ref.shape
    mism = (got != ref).float().mean().item()
    assert mism == 0.0, mism
    for i, t in enumerate(ti):                                   # the rules themselves
        assert (got[:t] != i + 1).all()


def test_full_size_c2_properties():
    """BASELINE C2 size (50 x 480 x 854, 1 object): size-independent properties -- nothing before the query frame, the query
    frame equals the (nearest-resized) ground truth, elsewhere mask == (logit > 0)."""
    g = torch.Generator().manual_seed(3)
    T, H, W = 50, 480, 854
    l = torch.randn((T, H, W), generator=g).cuda()
    gt = (torch.rand((H, W), generator=g) > 0.5).float()
    got = vos_index_masks([l], [gt], [7], (H, W))
    assert (got[:7] == 0).all()
    assert torch.equal(got[7].cpu(), gt.to(torch.uint8))
    assert torch.equal(got[8:], (l[8:] > 0).to(torch.uint8))
