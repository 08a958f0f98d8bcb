#!/bin/bash
# timelines of the persistent attention kernels (trace build), new column-sum kernel
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py fwd > gpurun_out/r2_trace_fwd.txt 2>&1; tail -3 gpurun_out/r2_trace_fwd.txt
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py bwd > gpurun_out/r2_trace_bwd.txt 2>&1; tail -3 gpurun_out/r2_trace_bwd.txt
timeout 200 python -m pytest tests/test_rowops_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/kbench.py --only ln --json gpurun_out/r2_kbench_call9.json 2>&1 | grep -E "colsum|layernorm_fwd"
