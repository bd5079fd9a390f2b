#!/bin/bash
# A/B: partially unrolled tested steps (smaller hot code) against the committed kernel.
mkdir -p gpurun_out
bash tools/gpu_ab.sh "cfg2 cfg1 cfg4" "head tu8 tu4" 17179869184 2>&1 | tee gpurun_out/tested_unroll_ab.txt
