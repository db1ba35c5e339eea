"""Weight names in libsampt_b200 are shared per device.  Several models of the same kind may live in one process (the test
suite itself, or a harness comparing SAM with HQ-SAM): whoever runs must see ITS weights, and a plain SAM decoder registered
after an HQ-SAM decoder must not inherit the HQ-only tensors."""
import pytest
import torch

from oracle import pips_ref, sam_ref
from sampt_b200 import synth

pytestmark = pytest.mark.gpu


def test_two_pips_models_alternate():
    from sam_pt.point_tracker.pips import Pips
    models = []
    for seed in (7201, 99):
        m = Pips(S=8, stride=4)
        m.load_state_dict(synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), seed)))
        models.append(m.cuda().eval())
    frames = synth.make_clip(8, 96, 128, seed=5)["frames"].cuda()
    q = torch.tensor([[0.0, 40.0, 30.0], [0.0, 90.0, 60.0]], device="cuda")

    def run(m):
        traj, _ = m.track(m.encode_frames(frames), q, 0.9, iters=6)
        return traj.cpu()

    a0, b0 = run(models[0]), run(models[1])
    a1, b1 = run(models[0]), run(models[1])
    assert (a0 - b0).abs().max() > 1e-3          # different weights really give different tracks
    assert torch.equal(a0, a1) and torch.equal(b0, b1)


def test_plain_sam_after_hq_sam_in_one_process():
    from sampt_b200 import factory
    from segment_anything.predictor import SamPredictor
    from segment_anything_hq.predictor import SamPredictor as HQPredictor
    cfg = sam_ref.VIT_TEST
    sd_plain = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg), 31))
    sd_hq = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg, hq=True), 47))
    plain = SamPredictor(factory.build_sam("vit_test", sd_plain).cuda())
    hq = HQPredictor(factory.build_sam("vit_test", sd_hq, hq=True).cuda())
    img = synth.make_clip(1, 96, 128, seed=3)["frames"][0].permute(1, 2, 0).numpy()
    pts = torch.tensor([[[300.0, 200.0], [600.0, 400.0]]], device="cuda")
    lab = torch.ones((1, 2), dtype=torch.int, device="cuda")

    def run(p):
        p.set_image(img)
        m, i, _ = p.predict_torch(pts, lab, None, None, False, True)
        return m.cpu(), i.cpu()

    m0, i0 = run(plain)
    mh, _ = run(hq)
    m1, i1 = run(plain)   # used to fail: "HQ decoder expects 6 output tokens, got 5" (stale hf_* tensors in the registry)
    mh2, _ = run(hq)
    assert torch.equal(m0, m1) and torch.equal(i0, i1)
    assert torch.equal(mh, mh2)
