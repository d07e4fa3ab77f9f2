#!/bin/bash
# usage: tools/run_checks.sh stage1 stage2 ...   (each stage under its own timeout; logs in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
for s in "$@"; do
  echo "=== stage $s ===" | tee -a gpurun_out/check.log
  timeout 300 python tools/gpu_check.py --stage $s 2>&1 | tail -60 | tee -a gpurun_out/check.log
  echo "exit: ${PIPESTATUS[0]}" | tee -a gpurun_out/check.log
done
