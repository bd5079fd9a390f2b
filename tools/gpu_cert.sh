#!/bin/bash
# A/B: one-rounding binning + direct-scatter build against the committed kernel; parity first.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_bin1.txt
bash tools/gpu_ab.sh "cfg1 cfg4 cfg3 cfg5 cfg2" "head bin1" 17179869184 2>&1 | tee gpurun_out/bin1_ab.txt
