#!/bin/bash
# A/B: leaner tiled append against the committed kernel; tiled parity tests first.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "tile or tiled or fused or production or cfg3" 2>&1 | tail -3
bash tools/gpu_ab.sh "cfg3 cfg5 cfg3_m20000" "head tile1" 17179869184 2>&1 | tee gpurun_out/tile1_ab.txt
