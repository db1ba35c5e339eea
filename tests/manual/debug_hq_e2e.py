"""Bring-up aid: the HQ-SAM + PIPS tiny clip of tests/test_gpu_sam.py::test_hq_encoder_interm_and_e2e with per-frame diagnostics
(visibilities, logit differences, where masks disagree, the oracle's refinement-box margin).  Run on the GPU box:
    python tests/manual/debug_hq_e2e.py            (SAMPT_DECODER_TC=0 / SAMPT_VIT_PRECISION=.. select variants)"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sam-pt_b200"))

from oracle import pips_ref, sam_ref, sampt_ref  # noqa: E402
from sampt_b200 import factory, synth  # noqa: E402


def main():
    cfg = sam_ref.VIT_TEST
    sam_sd = synth.condition_sam(synth.make_state_dict(sam_ref.sam_state_dict_shapes(cfg, hq=True), 47))
    pips_sd = synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), 7201))
    ckpt = synth.write_pips_checkpoint_dir(pips_sd, os.path.join(tempfile.mkdtemp(), "pips"))
    video = synth.make_video_dict(3, 96, 128, 4)
    taps = {}
    ref = sampt_ref.sampt_forward(pips_sd, sam_ref.RefSamPredictor(sam_sd, cfg, hq=True), video, positive_points_per_mask=4,
                                  sam_iou_threshold=-1e9, taps=taps)
    model = factory.build_sam_pt("vit_test", sam_sd, ckpt, positive_points_per_mask=4, sam_iou_threshold=-1e9, hq=True)
    out = model(video)
    print("env", {k: v for k, v in os.environ.items() if k.startswith("SAMPT_")})
    print("traj max diff", float((out["trajectories"].cpu() - ref["trajectories"]).abs().max()))
    print("vis equal", torch.equal(out["visibilities"].cpu(), ref["visibilities"]))
    print("scores", out["scores"], ref["scores"])
    for f in range(3):
        a, b = out["logits"][0][f].cpu(), ref["logits"][0][f]
        d = (a - b).abs()
        diff = (a > 0) != (b > 0)
        print(f"frame {f}: max|dlogit| {float(d.max()):.4e} mean {float(d.mean()):.4e} at {divmod(int(d.argmax()), a.shape[1])}; "
              f"flipped {int(diff.sum())} of {int((b > 0).sum())}; box_margin {taps['box_margin'][(f, 0)]:.4f}; n_refine {taps['n_refine'][(f, 0)]}")
        for y, x in diff.nonzero().tolist():
            print(f"   flipped ({y},{x}): gpu {float(a[y, x]):.5f} oracle {float(b[y, x]):.5f}")


if __name__ == "__main__":
    main()
