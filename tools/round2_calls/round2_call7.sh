#!/bin/bash
# attention forward v2 (loader warp, probing MMA issuer, pipelined TMEM loads): parity, isolated timing, ncu, step
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests7.log 2>&1; tail -5 gpurun_out/r2_attn_tests7.log
timeout 300 python tools/kbench.py --only attn --json gpurun_out/r2_kbench_call7.json 2>&1 | tail -7
PROF_ONLY=attn timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn_fwd_pair -f -o gpurun_out/r2_prof_attn7 python tools/prof_r2.py 2>&1 | tail -2
timeout 600 python tools/ab.py sweep fwd1:MMFB_ATTN_FWD=1 again: --steps 16
