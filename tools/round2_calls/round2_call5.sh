#!/bin/bash
# paired-tile attention forward + masked-chunk skipping in the backward: parity, isolated timings, step
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py -m gpu -q -x -rfEP > gpurun_out/r2_attn_tests5.log 2>&1; tail -15 gpurun_out/r2_attn_tests5.log
timeout 600 python -m pytest tests -m gpu -q -rfE > gpurun_out/r2_gpu_tests_full5.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_full5.log
timeout 300 python tools/kbench.py --only attn --json gpurun_out/r2_kbench_call5.json 2>&1 | tail -8
timeout 900 python tools/ab.py sweep fwd1:MMFB_ATTN_FWD=1 again: fwd1b:MMFB_ATTN_FWD=1 --steps 16
timeout 300 python bench.py > gpurun_out/r2_bench_call5.json 2> gpurun_out/r2_bench_call5.err; tail -c 1500 gpurun_out/r2_bench_call5.json
