"""Per-function / per-line breakdown of an ncu --import-source report (SASS rows only).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_src_breakdown.py src.csv [kernels.cuh as built] [n_candidates]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
src_path = sys.argv[2] if len(sys.argv) > 2 else '/root/repo/cudabrot_b200/csrc/buddha_kernels.cuh'
ncand = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0 ** 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr_i]
iL, iSass = 0, 3
iInst = h.index("Instructions Executed")
iSamp = h.index("# Samples")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_")]
src_lines = open(src_path).read().split('\n')
funcs = []
for n, l in enumerate(src_lines, 1):
    if l.startswith('__device__') or l.startswith('__global__') or l.startswith('render_') or l.startswith('orbit_drain'):
        mm = re.search(r'(\w+)\s*\(', l if '(' in l and not l.startswith('__global__ void __launch') else src_lines[n])
        if mm: funcs.append((n, mm.group(1)))
def func_of(line):
    name = '?'
    for n, f in funcs:
        if n <= line: name = f
    return name
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')
per_line = collections.Counter(); per_func = collections.Counter(); samp_func = collections.Counter()
fp_func = collections.Counter(); per_op = collections.Counter(); samp_line = collections.Counter()
per_op_func = collections.defaultdict(collections.Counter)
stall_tot = collections.Counter()
total = 0; cur_line = None
for r in rows[hdr_i + 1:]:
    if len(r) <= iInst: continue
    if r[iL]:
        try: cur_line = int(r[iL])
        except ValueError: pass
    if not r[2].startswith('0x'): continue     # source-only (aggregate) row
    sass = r[iSass].strip()
    try: inst = int(r[iInst] or 0)
    except ValueError: continue
    toks = sass.split()
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    f = func_of(cur_line or 0)
    per_line[cur_line] += inst; per_func[f] += inst; total += inst; per_op[op] += inst
    per_op_func[f][op] += inst
    if op in FP64: fp_func[f] += inst
    try: s = int(r[iSamp] or 0)
    except ValueError: s = 0
    samp_func[f] += s; samp_line[cur_line] += s
    for i, n in stall_cols:
        try: stall_tot[n] += int(r[i] or 0)
        except ValueError: pass
ts = sum(samp_func.values())
fp_total = sum(fp_func.values())
print("warp instr %d = %.2f per candidate (%.1f thread-instr); FP64 %.2f per candidate (%.1f thread-instr, %.1f %%)" %
      (total, total / ncand, 32 * total / ncand, fp_total / ncand, 32 * fp_total / ncand, 100 * fp_total / total))
print("%-26s %7s %7s %7s %8s" % ("function", "inst%", "fp64%", "other%", "samples%"))
for f, v in per_func.most_common(24):
    print("%-26s %6.2f%% %6.2f%% %6.2f%% %7.2f%%   %s" % (
        f, 100 * v / total, 100 * fp_func[f] / total, 100 * (v - fp_func[f]) / total, 100 * samp_func[f] / max(ts, 1),
        ", ".join("%s %.1f" % (o, 100 * c / total) for o, c in per_op_func[f].most_common(7))))
print("\nops: " + ", ".join("%s %.1f%%" % (o, 100 * c / total) for o, c in per_op.most_common(30)))
if stall_tot:
    st = sum(stall_tot.values())
    print("\nstall samples: " + ", ".join("%s %.1f%%" % (n[6:], 100 * c / st) for n, c in stall_tot.most_common(12)))
print()
for l, v in per_line.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    print("%5s %6.2f%% inst %6.2f%% smp  %s" % (l, 100 * v / total, 100 * samp_line[l] / max(ts, 1), src_lines[l - 1].strip()[:100] if l else ''))
