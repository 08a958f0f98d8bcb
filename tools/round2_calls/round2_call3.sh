#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_next_gpu.py -m gpu -q -rfEP -k "ce_rows or fused_masked" 2>&1 | tail -12
MMFB_LIB=$PWD/mmf_b200/csrc/libmmfb200_x2.so timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -o gpurun_out/r2_prof_x2 python tools/prof_r2.py 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
