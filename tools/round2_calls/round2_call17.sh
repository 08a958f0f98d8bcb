#!/bin/bash
# GEMM epilogue: second half of the accumulator slice requested before the first half's math - tests, isolated timings, A/B
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/kbench.py --only gemm > gpurun_out/r2_kbench_call17_epf.log 2>&1; grep gemm gpurun_out/r2_kbench_call17_epf.log | cut -c1-150
MMFB_LIB=mmf_b200/csrc/libmmfb200_noepf.so timeout 300 python tools/kbench.py --only gemm > gpurun_out/r2_kbench_call17_noepf.log 2>&1; grep gemm gpurun_out/r2_kbench_call17_noepf.log | cut -c1-150
timeout 600 python tools/ab.py sweep noepf again: noepfb:MMFB_LIB=mmf_b200/csrc/libmmfb200_noepf.so --steps 16
