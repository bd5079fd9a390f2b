#!/bin/bash
# A/B of scheduler / sampler micro-optimisations against the committed kernel.
mkdir -p gpurun_out
bash tools/gpu_ab.sh "cfg1 cfg2 cfg4 cfg3" "head sched schedg" 17179869184 2>&1 | tee gpurun_out/sched_ab.txt
