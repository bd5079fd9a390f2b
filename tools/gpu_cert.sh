#!/bin/bash
# The two-per-lane sampler as the default: full parity run, then the tiled / fused workloads against
# the previous kernel (shared memory per warp grew by 512 B).
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_ab.sh "cfg3 cfg5" "head new" 17179869184 2>&1 | tee gpurun_out/gen2_tiled_ab.txt
