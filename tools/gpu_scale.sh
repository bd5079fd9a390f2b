#!/bin/bash
# N-GPU bench (run under gpurun --gpus N): torchrun bench on cfg2 (default) and cfg3_m20000, CLI --gpus N parity
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_gpus.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-300; cut -c1-700 gpurun_out/bench_${N}gpu.json
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $N --steps 2 --warmup 1 --workload cfg3_m20000 > gpurun_out/bench_${N}gpu_cfg3_m20000.json 2>> gpurun_out/bench_${N}gpu.err
cut -c1-700 gpurun_out/bench_${N}gpu_cfg3_m20000.json
timeout -s KILL 300 cudabrot_b200/bin/cudabrot --gpus $N -w 2000 -h 2000 -m 2000 -c 20 --samples 1073741824 -o gpurun_out/multi.pgm 2>&1 | grep -E "passes|Max|error|fail"
timeout -s KILL 300 cudabrot_b200/bin/cudabrot --gpus 1 -w 2000 -h 2000 -m 2000 -c 20 --samples 1073741824 -o gpurun_out/single.pgm 2>&1 | grep -E "passes|Max"
cmp gpurun_out/multi.pgm gpurun_out/single.pgm && echo "PGM identical for $N GPUs vs 1 GPU"
rm -f gpurun_out/multi.pgm gpurun_out/single.pgm
