#!/bin/bash
# dropout through mask look-up tables in the attention forward (packed pair AND) and backward (keep factor AND)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py tests/test_encoder_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests18.log 2>&1; tail -3 gpurun_out/r2_attn_tests18.log
for i in 1 2; do timeout 200 python tools/kbench.py --only attn --iters 30 > gpurun_out/r2_kbench_call18_$i.log 2>&1; grep -E "attention|timeout|FAILED" gpurun_out/r2_kbench_call18_$i.log | cut -c1-150; done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_call18.json 2> gpurun_out/r2_bench_call18.err; python tools/show_bench.py gpurun_out/r2_bench_call18.json
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
