#!/usr/bin/env python
"""Launch each secondary hot kernel a few times in isolation (for `ncu --set full -k regex:...`)."""
import os
import sys
from ctypes import c_int

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sam-pt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sampt_b200 import native, factory, synth  # noqa: E402

dev = torch.device("cuda", 0)
ctx = native.get_context(dev)
L = native.lib()
# fp32 linear M=64 (PIPS mixer shapes)
for (M, N, K) in [(64, 2048, 512), (64, 512, 2048)]:
    x = torch.randn((M, K), device=dev); w = torch.randn((N, K), device=dev); b = torch.randn((N,), device=dev)
    y = torch.empty((M, N), device=dev)
    for _ in range(3):
        native.check(L.sampt_linear_f32(ctx.handle, native.ptr(x), c_int(K), native.ptr(w), c_int(K), native.ptr(b), native.ptr(None),
                                        c_int(0), native.ptr(y), c_int(N), c_int(M), c_int(N), c_int(K), c_int(0), native.stream_ptr()))
torch.cuda.synchronize()
# one ViT-H encode of 2 frames (attn_prep / attn_tc windowed + global / ln_rows) and one PIPS fnet of 2 frames (conv, inorm)
from segment_anything.predictor import SamPredictor  # noqa: E402
sam = factory.build_sam("vit_h").to(dev)
pred = SamPredictor(sam)
clip = synth.make_clip(2, 480, 854, seed=1)
frames = clip["frames"].to(dev)
for _ in range(2):
    pred.encode_frames(frames)
from sam_pt.point_tracker.pips import Pips  # noqa: E402
pm = Pips(S=8, stride=4).to(dev)
for _ in range(2):
    pm.encode_frames(frames)
torch.cuda.synchronize()
print("done")
