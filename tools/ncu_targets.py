#!/usr/bin/env python
"""The roofline kernels of bench.py launched in isolation, as the target of one `ncu --set full` capture:

  ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|pips_corr|attn_ws' -o gpurun_out/prof_r02_roofline \
      python tools/ncu_targets.py

(ViT-H mlp.lin1 GEMM, M=40960 N=5120 K=1280, first in the default fp8-corrected form [precision 6], then as three fp16 passes
[precision 4]; fused correlation gather, N=292 points, L2 flushed before every launch; the windowed and the global attention launch
of one 10-frame ViT-H batch).  profiles/extract_traffic.py turns the report's raw page into profiles/r02_roofline_traffic.json, which
bench.py reads for the `roofline.traffic` fields."""
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sam-pt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

bench.ROOFLINE_WARM, bench.ROOFLINE_REPS = 1, 2   # 2-3 launches per kernel: the --set full report stays below gpurun's 64 MiB
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
enc = SimpleNamespace(embed_dim=1280)
model = SimpleNamespace(sam_predictor=SimpleNamespace(model=SimpleNamespace(image_encoder=enc)))
if os.environ.get("NCU_TARGET") == "pips":
    # one PIPS tracker run on a small clip (8 points): target of `ncu -k regex:sgemm_skinny|mixer_token|pips_corr -c ..`
    from oracle import pips_ref  # noqa: E402  (shapes of the synthetic state-dict only)
    from sam_pt.point_tracker.pips import Pips  # noqa: E402
    from sampt_b200 import synth  # noqa: E402
    m = Pips(S=8, stride=4)
    m.load_state_dict(synth.condition_pips(synth.make_state_dict(pips_ref.pips_state_dict_shapes(), 7201)))
    m = m.cuda().eval()
    frames = synth.make_clip(8, 480, 854, seed=5)["frames"].cuda()
    q = torch.tensor([[0.0, 100.0 + 60 * i, 80.0 + 40 * i] for i in range(8)], device="cuda")
    traj, _ = m.track(m.encode_frames(frames), q, 0.9, iters=6)
    torch.cuda.synchronize()
    print(traj.shape)
    sys.exit(0)
print(bench.gemm_roofline(model, dev, SimpleNamespace(encoder_batch=10, precision=6)))
print(bench.gemm_roofline(model, dev, SimpleNamespace(encoder_batch=10, precision=4)))
print(bench.corr_roofline(dev))
print(bench.attn_roofline(dev))
