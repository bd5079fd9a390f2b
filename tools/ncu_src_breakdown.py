"""Per-source-line / per-phase instruction breakdown of an ncu --import-source report.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_src_breakdown.py src.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr_i]
iL, iS, iA = 0, 1, 2
iSass = 3
iInst = h.index("Instructions Executed")
iSamp = h.index("# Samples")
src_lines = open('/root/repo/cudabrot_b200/csrc/buddha_kernels.cuh').read().split('\n')
# function ranges by scanning the source for top-level definitions
funcs = []
for n, l in enumerate(src_lines, 1):
    m = re.match(r'^(?:__device__|__global__|template|struct|#define)\b.*?(\w+)\s*\(', l)
    m2 = re.match(r'^\w[\w\s\*&:<>]*\s(\w+)\(', l)
    if l.startswith('#define BUDDHA_ZSTEP'): funcs.append((n, 'ZSTEP'))
    elif l.startswith('__device__') or l.startswith('__global__') or l.startswith('render_') or l.startswith('orbit_drain') :
        # name on this or next line
        mm = re.search(r'(\w+)\s*\(', l if '(' in l and not l.startswith('__global__ void __launch') else src_lines[n])
        if mm: funcs.append((n, mm.group(1)))
def func_of(line):
    name = '?'
    for n, f in funcs:
        if n <= line: name = f
    return name
per_line = collections.Counter(); per_func = collections.Counter(); samp_func = collections.Counter()
per_op_func = collections.defaultdict(collections.Counter)
total = 0
cur_line = None
for r in rows[hdr_i + 1:]:
    if len(r) <= iInst: continue
    try: inst = int(r[iInst] or 0)
    except ValueError: continue
    if r[iL]: 
        try: cur_line = int(r[iL])
        except ValueError: pass
    if not r[iA]: continue   # source-only row
    per_line[cur_line] += inst; f = func_of(cur_line or 0); per_func[f] += inst; total += inst
    try: samp_func[f] += int(r[iSamp] or 0)
    except ValueError: pass
    op = r[iSass].split()[0] if r[iSass] else '?'
    if op.startswith('@'): op = r[iSass].split()[1]
    per_op_func[f][op.split('.')[0]] += inst
print("total warp instr", total)
ts = sum(samp_func.values())
for f, v in per_func.most_common():
    print("%-22s %6.2f%% inst  %6.2f%% samples   top ops: %s" % (f, 100 * v / total, 100 * samp_func[f] / max(ts, 1),
          ", ".join("%s %.1f%%" % (o, 100 * c / total) for o, c in per_op_func[f].most_common(6))))
print()
for l, v in per_line.most_common(40):
    print("%5s %6.2f%%  %s" % (l, 100 * v / total, src_lines[l - 1].strip()[:110] if l else ''))
