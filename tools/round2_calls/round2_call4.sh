#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEP > gpurun_out/r2_gpu_tests_full4.log 2>&1; tail -12 gpurun_out/r2_gpu_tests_full4.log
MMFB_PDL=0 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/kbench.py --json gpurun_out/r2_kbench_call4.json 2>&1 | tail -32
timeout 900 python tools/ab.py sweep pdl0:MMFB_PDL=0 nox2 again: pdl0b:MMFB_PDL=0 --steps 16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_step_call4.csv python bench.py --profile --steps 1 --warmup 3 --no-parity 2>&1 | tail -2
python tools/agg_launches.py gpurun_out/r2_launches_step_call4.csv 2>&1 | head -28
