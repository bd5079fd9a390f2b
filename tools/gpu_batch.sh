timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tools/gpu_ab.sh "cfg3_m20000 cfg3 cfg5 cfg2 cfg1 cfg4" "cur" 8589934592 2>&1 | grep -v "^ \|Traceback\|json\|File" | tee gpurun_out/ab_r2_7.txt
export BUDDHA_NO_CARRY=1
echo "BUDDHA_NO_CARRY=1:"; tools/gpu_ab.sh "cfg3_m20000 cfg3 cfg5" "cur" 8589934592 2>&1 | grep -v "^ \|Traceback\|json\|File" | tee -a gpurun_out/ab_r2_7.txt
unset BUDDHA_NO_CARRY
for wl in dense64 dense256 dense1k dense4k zoom_dense; do python bench.py --workload $wl --steps 3 --warmup 1 --skip-baselines --no-extras --samples-per-step 8589934592 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl: %.3e samples/s  %.3e increments/s  red frac %.2f  fp64 frac %.2f' % (d['value'], d['increments_per_s'], d['roofline_red']['frac'], d['roofline']['frac']))"; done | tee gpurun_out/dense.txt
