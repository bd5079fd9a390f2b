#!/bin/bash
# Round-2 certificate experiment: the other builds / workloads with the first certificate age at 64.
mkdir -p gpurun_out
bash tools/gpu_ab.sh "cfg3_m20000 cfg5 cfg4" "base qall64" 8589934592 2>&1 | tee gpurun_out/cert_ab4.txt
