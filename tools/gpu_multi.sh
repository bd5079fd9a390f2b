#!/bin/bash
# Multi-GPU checks (run under gpurun --gpus N): in-process merge test, CLI --gpus, torchrun bench.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout -s KILL 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "multi_gpu or cli_end" 2>&1 | tail -5
timeout -s KILL 300 cudabrot_b200/bin/cudabrot --gpus $N -w 2000 -h 2000 -m 2000 -c 20 --samples 1073741824 -o gpurun_out/multi.pgm 2>&1 | tail -8
timeout -s KILL 300 cudabrot_b200/bin/cudabrot --gpus 1 -w 2000 -h 2000 -m 2000 -c 20 --samples 1073741824 -o gpurun_out/single.pgm 2>&1 | grep -E "passes|Max"
cmp gpurun_out/multi.pgm gpurun_out/single.pgm && echo "PGM identical for $N GPUs vs 1 GPU"
rm -f gpurun_out/multi.pgm gpurun_out/single.pgm
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err; cat gpurun_out/bench_${N}gpu.json
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 2 --warmup 1 --workload cfg3 > gpurun_out/bench_${N}gpu_cfg3.json 2>> gpurun_out/bench_${N}gpu.err
cat gpurun_out/bench_${N}gpu_cfg3.json
