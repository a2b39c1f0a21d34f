"""Dev tool: join an ncu SASS source page with nvdisasm line info -> per-source-line instruction / stall-sample
shares.  usage: ncu_lines.py report.ncu-rep kernel-substring [object-prefix]   (object-prefix: fg_lead | fg_kernels)"""
import collections, csv, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else "lead_kernelILi1"
obj = sys.argv[3] if len(sys.argv) > 3 else "fg_lead"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "fugu_b200/libfugu_gpu.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(obj)][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
fn = None; line = None; funcs = collections.OrderedDict()
for l in dis.split("\n"):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); funcs[fn] = []; continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and fn: funcs[fn].append((line, m.group(2)))
name = [f for f in funcs if sub in f][0]
ins = funcs[name]
skip = os.environ.get("LAUNCH_SKIP", "0")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
if len(data) > len(ins):
    data = data[:len(ins)]
assert len(data) == len(ins), (len(data), len(ins))
iex = hdr.index('Instructions Executed'); smp = hdr.index('# Samples')
cols = {k: hdr.index(k) for k in ['stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_branch_resolving', 'stall_not_selected', 'stall_math', 'stall_mio', 'stall_lg'] if k in hdr}
by = collections.defaultdict(lambda: collections.Counter())
for (ln, sass), r in zip(ins, data):
    c = by[ln]; c['inst'] += int(r[iex]); c['samp'] += int(r[smp])
    for k, i in cols.items(): c[k] += int(r[i])
ti = sum(c['inst'] for c in by.values()); ts = sum(c['samp'] for c in by.values())
srcs = {}
def text(ln):
    if not ln: return "(none)"
    f, n = ln
    if f not in srcs:
        p = os.path.join(root, "fugu_b200/csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else None
    return (srcs[f][n - 1].strip()[:100] if srcs[f] and n - 1 < len(srcs[f]) else "")
print(f"kernel {name}: {ti} warp-instructions, {ts} samples")
print("file:line            inst%  samp%  lsb   ssb   wait  source")
for ln, c in sorted(by.items(), key=lambda kv: -kv[1]['inst'])[:int(os.environ.get("TOP", "60"))]:
    tag = f"{ln[0][:12]}:{ln[1]}" if ln else "-"
    print(f"{tag:20s} {100*c['inst']/ti:6.1f} {100*c['samp']/ts:6.1f} {100*c['stall_long_sb']/ts:5.1f} {100*c['stall_short_sb']/ts:5.1f} {100*c['stall_wait']/ts:5.1f}  {text(ln)}")
