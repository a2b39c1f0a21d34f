"""Dev tool: join an ncu SASS source page with nvdisasm line info -> per-source-line and per-region
instruction / stall-sample shares.  usage: ncu_lines.py report.ncu-rep [kernel-substring]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else "search_kernelILi1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "fugu_b200/libfugu_gpu.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("fg_kernels")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
fn = None; line = None; funcs = collections.OrderedDict()
for l in dis.split("\n"):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); funcs[fn] = []; continue
    m = re.search(r'//## File ".*fg_kernels.cu", line (\d+)', l)
    if m: line = int(m.group(1)); continue
    if re.search(r'//## File "(.*)", line (\d+)', l): line = -1; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and fn: funcs[fn].append((line, m.group(2)))
name = [f for f in funcs if sub in f][0]
ins = funcs[name]
skip = os.environ.get("LAUNCH_SKIP", "0")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
if len(data) > len(ins):  # several kernels in the page: keep the rows of the requested one
    data = data[:len(ins)]
assert len(data) == len(ins), (len(data), len(ins))
iex = hdr.index('Instructions Executed'); smp = hdr.index('# Samples')
cols = {k: hdr.index(k) for k in ['stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_branch_resolving', 'stall_not_selected', 'stall_math', 'stall_mio', 'stall_lg']}
by = collections.defaultdict(lambda: collections.Counter())
for (ln, sass), r in zip(ins, data):
    c = by[ln]; c['inst'] += int(r[iex]); c['samp'] += int(r[smp])
    for k, i in cols.items(): c[k] += int(r[i])
ti = sum(c['inst'] for c in by.values()); ts = sum(c['samp'] for c in by.values())
src = open(os.path.join(root, "fugu_b200/csrc/fg_kernels.cu")).read().split("\n")
print(f"kernel {name}: {ti} warp-instructions, {ts} samples")
print("line   inst%  samp%  barr  lsb   ssb   wait  source")
for ln, c in sorted(by.items(), key=lambda kv: -kv[1]['samp'])[:int(os.environ.get("TOP", "45"))]:
    t = src[ln - 1].strip()[:90] if ln > 0 else "(inlined/other)"
    print(f"{ln:5d} {100*c['inst']/ti:6.1f} {100*c['samp']/ts:6.1f} {100*c['stall_barrier']/ts:5.1f} {100*c['stall_long_sb']/ts:5.1f} {100*c['stall_short_sb']/ts:5.1f} {100*c['stall_wait']/ts:5.1f}  {t}")

# ---- region aggregation (markers in the source) ----
marks = []
for i, t in enumerate(src, 1):
    m = re.search(r'//\s*(----[^-].*?----|\([ab]\).*|slot lookup.*|\.\.\. then.*|candidate bitmap for.*)$', t.strip())
    if m: marks.append((i, m.group(1)[:50]))
marks = [(1, 'helpers')] + marks + [(10**9, 'end')]
reg = collections.defaultdict(lambda: collections.Counter())
for ln, c in by.items():
    name_ = 'inlined/other'
    if ln > 0:
        for (a, n1), (b, _) in zip(marks, marks[1:]):
            if a <= ln < b: name_ = n1
    reg[name_].update(c)
print("\nregion                                             inst%  samp%  barr  lsb   ssb   wait")
for n1, c in sorted(reg.items(), key=lambda kv: -kv[1]['inst']):
    print(f"{n1:50s} {100*c['inst']/ti:6.1f} {100*c['samp']/ts:6.1f} {100*c['stall_barrier']/ts:5.1f} {100*c['stall_long_sb']/ts:5.1f} {100*c['stall_short_sb']/ts:5.1f} {100*c['stall_wait']/ts:5.1f}")
