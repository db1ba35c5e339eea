#!/bin/bash
# tcgen05.mma issue cadence: one accumulator (dependent instructions) vs two alternating accumulators (tools/mma_cadence_probe.cu)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sam-pt_b200/csrc -o /tmp/mma_probe tools/mma_cadence_probe.cu > gpurun_out/c15_build.log 2>&1; echo "build rc=$?"
timeout 120 /tmp/mma_probe > gpurun_out/c15_mma_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/c15_mma_probe.log
