#!/bin/bash
# attention forward with bulk-tensor-store epilogue + atomic item counter, persistent backward with the item counter
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests10.log 2>&1; tail -6 gpurun_out/r2_attn_tests10.log
timeout 300 python -m pytest tests/test_encoder_gpu.py tests/test_depth_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/kbench.py --only attn --json gpurun_out/r2_kbench_call10.json 2>&1 | tail -7
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py fwd > gpurun_out/r2_trace_fwd10.txt 2>&1; tail -2 gpurun_out/r2_trace_fwd10.txt
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py bwd > gpurun_out/r2_trace_bwd10.txt 2>&1; tail -2 gpurun_out/r2_trace_bwd10.txt
timeout 300 python tools/step_gaps.py > gpurun_out/r2_step_gaps.txt 2>&1; head -40 gpurun_out/r2_step_gaps.txt
timeout 600 python tools/ab.py sweep fwd1:MMFB_ATTN_FWD=1 bwd16:MMFB_ATTN_BWD=16 lnstream:MMFB_LN_BWD=stream --steps 16
