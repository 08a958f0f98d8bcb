#!/bin/bash
# staged-barrier fix (per-stage o_staged / alternating kv_staged): stress with ragged padding; graph test with its full log
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_next_gpu.py -m gpu -q -x -rfE -k graphed > gpurun_out/r2_graph_test15.log 2>&1; grep -E "^E |passed|failed" gpurun_out/r2_graph_test15.log | head -12
timeout 300 python -m pytest tests/test_variants_gpu.py tests/test_attention_gpu.py -m gpu -q -x -rfE > gpurun_out/r2_attn_tests15.log 2>&1; tail -3 gpurun_out/r2_attn_tests15.log
for i in 1 2 3 4; do timeout 200 python tools/kbench.py --only attn --iters 40 > gpurun_out/r2_kbench_call15_$i.log 2>&1; grep -E "attention|timeout|FAILED" gpurun_out/r2_kbench_call15_$i.log | cut -c1-150; done
timeout 600 python tools/ab.py sweep fwdb:MMFB_ATTN_FWD=b again: --steps 16
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
