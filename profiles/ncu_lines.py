"""Attribute ncu warp-stall samples to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <kz_engine.cu as profiled> [top]
The report's SASS page is joined, instruction by instruction, with `nvdisasm --print-line-info-inline`
of a cubin rebuilt from the same source (same flags), because the CSV source page carries SASS only."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, src = sys.argv[1], os.path.abspath(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
cubin, dis, page = os.path.join(tmp, "k.cubin"), os.path.join(tmp, "k.dis"), os.path.join(tmp, "page.csv")
subprocess.check_call(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-cubin",
                       src, "-o", cubin], stderr=subprocess.DEVNULL)
open(dis, "w").write(subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout)
open(page, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:kz_step"],
                                     capture_output=True, text=True).stdout)
seq, pending, func, cur = [], [], None, None
base = os.path.basename(src)
for ln in open(dis):
    m = re.match(r'\.text\.(\S+):', ln)
    if m:
        func = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and func and 'kz_step_kernelILi1' in func.replace('<', 'I'):
        if pending:
            eng = [p for p in pending if p[0] == base]
            cur = eng[0][1] if eng else None
            pending = []
        seq.append(cur)
rows = list(csv.reader(open(page)))
h = rows[1]
data = [r for r in rows[2:] if len(r) >= len(h) and r[0] != h[0]]  # skip per-kernel header rows of further launches
ci = {n: i for i, n in enumerate(h)}
keys = ['stall_long_sb', 'stall_wait', 'stall_no_inst', 'stall_short_sb', 'stall_branch_resolving', 'stall_lg',
        'stall_mio', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_barrier', 'stall_dispatch', 'stall_drain']
agg = collections.defaultdict(collections.Counter)
# the page lists every profiled launch one after another: keep the first
assert len(data) >= len(seq), (len(seq), len(data))
data = data[:len(seq)]
for line, r in zip(seq, data):
    c = agg[line]
    c['samples'] += int(r[ci['# Samples']]); c['inst'] += int(r[ci['Instructions Executed']])
    for k in keys:
        c[k] += int(r[ci[k]])
tot = sum(c['samples'] for c in agg.values())
print("total samples", tot, " instructions", sum(c['inst'] for c in agg.values()))
allk = collections.Counter()
for c in agg.values():
    for k in keys: allk[k] += c[k]
print("stall mix:", {k.replace('stall_', ''): round(100 * v / tot, 1) for k, v in allk.most_common()})
text = open(src).read().split('\n')
for line, c in sorted(agg.items(), key=lambda x: -x[1]['samples'])[:top]:
    t = text[line - 1].strip()[:64] if line else ''
    print(f"{100*c['samples']/tot:5.1f}% L{line} inst={c['inst']/1e6:6.2f}M long={c['stall_long_sb']} wait={c['stall_wait']} "
          f"noinst={c['stall_no_inst']} short={c['stall_short_sb']} br={c['stall_branch_resolving']} lg={c['stall_lg']} mio={c['stall_mio']} | {t}")
