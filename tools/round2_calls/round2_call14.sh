#!/bin/bash
# blocked forward (two key blocks, single read of S), graph fixes
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests14.log 2>&1; tail -6 gpurun_out/r2_attn_tests14.log
timeout 300 python -m pytest tests/test_encoder_gpu.py tests/test_depth_gpu.py tests/test_next_gpu.py -m gpu -q -x -rfE 2>&1 | tail -4
timeout 300 python tools/kbench.py --only attn --json gpurun_out/r2_kbench_call14.json > gpurun_out/r2_kbench_call14.log 2>&1; cat gpurun_out/r2_kbench_call14.log | cut -c1-200 | tail -9
MMFB_LIB=mmf_b200/csrc/libmmfb200_trace.so timeout 200 python tools/trace_attn.py fwd > gpurun_out/r2_trace_fwd14.txt 2>&1; tail -2 gpurun_out/r2_trace_fwd14.txt
timeout 400 python bench.py --workload mmbt --graph --no-cpu-baseline > gpurun_out/r2_bench_mmbt_graph.json 2> gpurun_out/r2_bench_mmbt_graph.err; python tools/show_bench.py gpurun_out/r2_bench_mmbt_graph.json; tail -2 gpurun_out/r2_bench_mmbt_graph.err
timeout 600 python tools/ab.py sweep fwd2:MMFB_ATTN_FWD=2 again: --steps 16
