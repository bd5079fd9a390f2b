timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tools/gpu_ab.sh "cfg2 cfg3_m20000 cfg5" "cur p4" 8589934592 2>&1 | grep -v "^ \|Traceback\|json\|File" | tee gpurun_out/ab_r2_8.txt
