"""Per-source-line executed instruction counts (outermost frame) of a profiled kernel, for a line range.
usage: ncu_linecounts.py rep lib kernel_substr cells lo hi"""
import csv, collections, re, subprocess, sys, io, os, tempfile
rep, lib, kname, cells, lo, hi = sys.argv[1:7]
cells = float(cells); lo = int(lo); hi = int(hi)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
dis = ""
for c in [f for f in os.listdir(d) if "sm_100a" in f]:
    dis += subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, c)], capture_output=True, text=True).stdout
lines = dis.split("\n")
start = [i for i, l in enumerate(lines) if l.strip().startswith(".text.") and kname in l][0]
instrs = []; cur = ("?", 0)
for l in lines[start + 1:]:
    if l.strip().startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln = m.group(1).split("/")[-1], int(m.group(2))
        mm = list(re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)))
        if mm: f, ln = mm[-1].group(1).split("/")[-1], int(mm[-1].group(2))
        cur = (f, ln); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.+?);", l)
    if m: instrs.append((cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
data = [r for r in rows[2:ends[1]] if len(r) == len(hdr)]
iI = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for ((f, ln), txt), r in zip(instrs, data):
    if not f.startswith(os.environ.get("MW_SRC", "dycore_kernels")) or not (lo <= ln <= hi): continue
    mo = re.match(r'\s*(@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)', txt); op = mo.group(2) if mo else '?'
    a = agg[ln]; a[0] += int(r[iI]); a[1] += int(r[iSm]); a[2][op.split('.')[0]] += int(r[iI])
for ln, a in sorted(agg.items(), key=lambda t: t[0]):
    print(f"line {ln}: instr/cell {a[0]*32/cells:7.1f} samples {a[1]:6d} | " + ", ".join(f"{o} {32*c/cells:.0f}" for o, c in a[2].most_common(6)))
