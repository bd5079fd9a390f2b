#!/bin/bash
# Round-2 certificate experiment: parity tests with the new kernel, then A/B against the previous build.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_cert.txt
bash tools/gpu_ab.sh "cfg2" "base cert cert16 cert30" 8589934592 2>&1 | tee gpurun_out/cert_ab.txt
bash tools/gpu_ab.sh "cfg3_m20000 cfg4 cfg5" "base cert" 4294967296 2>&1 | tee -a gpurun_out/cert_ab.txt
bash tools/gpu_ncu_ab.sh "cfg2" "base cert" 2>&1 | tee -a gpurun_out/cert_ab.txt
