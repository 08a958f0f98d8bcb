#!/bin/bash
# attention forward v2 + persistent backward + streaming LayerNorm backward: parity, isolated timings, step A/B
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py -m gpu -q -rfE > gpurun_out/r2_variant_tests8.log 2>&1; tail -8 gpurun_out/r2_variant_tests8.log
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_encoder_gpu.py tests/test_rowops_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests8.log 2>&1; tail -5 gpurun_out/r2_attn_tests8.log
timeout 300 python tools/kbench.py --only attn,ln --json gpurun_out/r2_kbench_call8.json 2>&1 | tail -16
timeout 900 python tools/ab.py sweep fwd1:MMFB_ATTN_FWD=1 bwd16:MMFB_ATTN_BWD=16 lnstream:MMFB_LN_BWD=stream again: --steps 16
PROF_ONLY=attn timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_prof_attn8 python tools/prof_r2.py 2>&1 | tail -2
