#!/bin/bash
# new defaults (paired forward with bulk stores, persistent backward v2, streaming LayerNorm backward, bit-sliced keep-bits):
# whole GPU suite, bench, isolated timings, host synchronisations / idle time of a step
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfE > gpurun_out/r2_gpu_tests_full12.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_full12.log
timeout 400 python bench.py --optimizer > gpurun_out/r2_bench_call12.json 2> gpurun_out/r2_bench_call12.err; python tools/show_bench.py gpurun_out/r2_bench_call12.json
timeout 300 python tools/kbench.py --json gpurun_out/r2_kbench_call12.json > gpurun_out/r2_kbench_call12.log 2>&1; cat gpurun_out/r2_kbench_call12.log | cut -c1-200
timeout 300 python tools/step_gaps.py > gpurun_out/r2_step_gaps12.txt 2>&1; head -34 gpurun_out/r2_step_gaps12.txt | cut -c1-220
