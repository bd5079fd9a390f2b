#!/bin/bash
# Selected ncu metrics of the render kernel for several library builds:
# tools/gpu_ncu_ab.sh "cfg1 cfg2" "base v22" [samples]
WLS=${1:-cfg2}; VARS=${2:-"base"}; N=${3:-1073741824}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warp_latency_issue_stalled_wait_per_warp_active.pct,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
mkdir -p gpurun_out
for wl in $WLS; do for v in $VARS; do
  BUDDHA_LIB=$PWD/tools/ab/$v.so BUDDHA_TILE_SERIAL=1 ncu --metrics $M --clock-control none -k regex:render_persistent -s 1 -c 1 --csv \
    --log-file gpurun_out/ncuab_${wl}_$v.csv timeout -s KILL 300 python bench.py --workload $wl --steps 1 --warmup 1 --skip-baselines --no-extras --samples-per-step $N > /dev/null 2>&1
  python - "$wl" "$v" <<'PY'
import csv, sys
wl, v = sys.argv[1:3]
rows = [r for r in csv.reader(l for l in open('gpurun_out/ncuab_%s_%s.csv' % (wl, v)) if l.startswith('"'))]
h = rows[0]; im, iv = h.index("Metric Name"), h.index("Metric Value")
d = {r[im]: r[iv] for r in rows[1:]}
def g(k):
    for n, x in d.items():
        if n.startswith(k): return float(x.replace(',', ''))
    return float('nan')
print("%s %-6s %.2f ms inst/cand %.2f fp64/cand %.2f issue %.1f%% fp64pipe %.1f%% warps %.1f%% regs %d | stalls/issue wait %.2f math %.2f notsel %.2f noinst %.2f ssb %.2f disp %.2f br %.2f lsb %.2f | bankconf %.3g" % (
    wl, v, g("gpu__time_duration")/1e6, g("smsp__inst_executed.sum")/2**30, g("sm__inst_executed_pipe_fp64")/2**30,
    g("smsp__issue_active"), g("sm__pipe_fp64_cycles_active"), g("sm__warps_active"), g("launch__registers_per_thread"),
    g("smsp__average_warps_issue_stalled_wait_per"), g("smsp__average_warps_issue_stalled_math"), g("smsp__average_warps_issue_stalled_not_sel"),
    g("smsp__average_warps_issue_stalled_no_inst"), g("smsp__average_warps_issue_stalled_short"), g("smsp__average_warps_issue_stalled_dispatch"),
    g("smsp__average_warps_issue_stalled_branch"), g("smsp__average_warps_issue_stalled_long"), g("l1tex__data_bank")))
PY
done; done
