#!/bin/bash
# packed-fp32 LayerNorm backward: parity + isolated timing; ncu --set full of the attention kernels and the LN backward
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rowops_gpu.py tests/test_variants_gpu.py tests/test_encoder_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_ln_tests6.log 2>&1; tail -5 gpurun_out/r2_ln_tests6.log
timeout 300 python tools/kbench.py --only ln --json gpurun_out/r2_kbench_call6.json 2>&1 | tail -8
PROF_ONLY=attn,ln timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_prof_attn6 python tools/prof_r2.py 2>&1 | tail -3
