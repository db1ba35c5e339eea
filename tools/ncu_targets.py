#!/usr/bin/env python
"""The two roofline kernels of bench.py launched in isolation, as the target of one `ncu --set full` capture:

  ncu --set full --clock-control none -k regex:'gemm_tc_kernel|pips_corr' -c 8 -o gpurun_out/prof_roofline python tools/ncu_targets.py

(ViT-H mlp.lin1 GEMM, M=40960 N=5120 K=1280, 3 split passes; fused correlation gather, N=292 points, L2 flushed before every
launch).  profiles/extract_traffic.py turns the report's raw page into profiles/r01_roofline_traffic.json, which bench.py
reads for the `roofline.traffic` fields."""
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sam-pt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
enc = SimpleNamespace(embed_dim=1280)
model = SimpleNamespace(sam_predictor=SimpleNamespace(model=SimpleNamespace(image_encoder=enc)))
print(bench.gemm_roofline(model, dev, SimpleNamespace(encoder_batch=10, precision=3)))
print(bench.corr_roofline(dev))
print(bench.attn_roofline(dev))
