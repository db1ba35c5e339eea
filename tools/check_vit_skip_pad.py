#!/usr/bin/env python
"""Hardware check of the experimental SAMPT_VIT_SKIP_PAD=1 path (csrc/vit_pipeline.cu): the first encode of a shape runs in
full and saves the image-independent padding rows; the second one runs the leading windowed blocks on the live windows only.
Both must give bit-identical features, and they must equal the features of a process WITHOUT the flag (pass --dump / --compare).

  SAMPT_VIT_SKIP_PAD=0 python tools/check_vit_skip_pad.py --dump /tmp/ref.pt
  SAMPT_VIT_SKIP_PAD=1 python tools/check_vit_skip_pad.py --compare /tmp/ref.pt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sam-pt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from sampt_b200 import factory, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dump")
    ap.add_argument("--compare")
    ap.add_argument("--vit", default="vit_b")
    args = ap.parse_args()
    from segment_anything.predictor import SamPredictor
    sam = factory.build_sam(args.vit)
    shapes = {k: tuple(v.shape) for k, v in sam.state_dict().items()}
    sam.load_state_dict(synth.condition_sam(synth.make_state_dict(shapes, 7202)))
    pred = SamPredictor(sam.cuda())
    out = {}
    for (H, W) in ((480, 854), (240, 320), (854, 480)):       # landscape (rows padded), 3:4, portrait (columns padded)
        frames = synth.make_clip(3, H, W, seed=5)["frames"].cuda()
        f1 = pred.encode_frames(frames)
        f2 = pred.encode_frames(frames)                         # with the flag: compacted run using the rows saved by f1
        f3 = pred.encode_frames(frames[:2])                     # another batch size
        torch.cuda.synchronize()
        assert torch.equal(f1, f2), (H, W, (f1 - f2).abs().max().item())
        assert torch.equal(f1[:2], f3), (H, W, (f1[:2] - f3).abs().max().item())
        out[f"{H}x{W}"] = f1.cpu()
        print(f"{H}x{W}: first / second / smaller-batch encodes identical; |features| max {f1.abs().max().item():.3f}")
    if args.dump:
        torch.save(out, args.dump)
    if args.compare:
        ref = torch.load(args.compare)
        for k, v in out.items():
            assert torch.equal(v, ref[k]), (k, (v - ref[k]).abs().max().item())
        print("identical to the reference dump")


if __name__ == "__main__":
    main()
