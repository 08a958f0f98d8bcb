#!/bin/bash
# GPU call 2 of round 2: whole suite again after the fixes, isolated kernel timings for every staged variant, the launch list
# of one step with the default kernels.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEP > gpurun_out/r2_gpu_tests_full2.log 2>&1; tail -15 gpurun_out/r2_gpu_tests_full2.log
MMFB_STAGED_TESTS=1 timeout 200 python -m pytest tests/test_staged_gpu.py -m gpu -q -k "layernorm" 2>&1 | tail -4
MMFB_LN_BWD=tile timeout 300 python -m pytest tests/test_rowops_gpu.py tests/test_encoder_gpu.py tests/test_frontends_gpu.py -m gpu -q -x 2>&1 | tail -3
MMFB_LN_BWD=lean timeout 300 python -m pytest tests/test_rowops_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/kbench.py --json gpurun_out/r2_kbench_default.json 2>&1 | tail -45
for v in x2 pf pfx2; do
  MMFB_LIB=$PWD/mmf_b200/csrc/libmmfb200_$v.so timeout 300 python tools/kbench.py --only gemm --json gpurun_out/r2_kbench_$v.json 2>&1 | tail -14
done
MMFB_LIB=$PWD/mmf_b200/csrc/libmmfb200_x2.so timeout 300 python tools/kbench.py --only attn --json gpurun_out/r2_kbench_x2_attn.json 2>&1 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_step_default.csv python bench.py --profile --steps 1 --warmup 3 --no-parity 2>&1 | tail -3
python tools/agg_launches.py gpurun_out/r2_launches_step_default.csv 2>&1 | head -40
