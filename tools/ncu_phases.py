"""Attribute executed instructions (fp64 / other) and stall samples of a profiled kernel to source-line ranges of
the kernel body.  usage: ncu_phases.py rep.ncu-rep lib.so kernel_mangled_substr cells "name:lo-hi,name:lo-hi,..." """
import csv, collections, re, subprocess, sys, io, os, tempfile
rep, lib, kname, cells, spec = sys.argv[1:6]
cells = float(cells)
ranges = []
for it in spec.split(","):
    n, r = it.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
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
iI = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
assert len(data) == len(instrs), (len(data), len(instrs))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])
tot = [0, 0, 0]
for ((f, ln), txt), r in zip(instrs, data):
    name = "other"
    if f.startswith(os.environ.get("MW_SRC", "dycore_kernels")):
        for n, lo, hi in ranges:
            if lo <= ln <= hi: name = n; break
    else:
        name = "other:" + f
    n = int(r[iI]); s = int(r[iSm])
    mo = re.match(r'\s*(@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)', txt); op = mo.group(2) if mo else '?'
    isfp = op.split('.')[0] in ("DFMA", "DMUL", "DADD")
    a = agg[name]; a[0] += n; a[1] += n if isfp else 0; a[2] += s; a[4][op.split('.')[0]] += n
    tot[0] += n; tot[1] += n if isfp else 0; tot[2] += s
print(f"total warp-instr {tot[0]}  per cell: {tot[0]*32/cells:.0f} thread-instr, {tot[1]*32/cells:.0f} fp64;  samples {tot[2]}")
for name, a in sorted(agg.items(), key=lambda t: -t[1][0]):
    top = ", ".join(f"{o} {32*c/cells:.0f}" for o, c in a[4].most_common(9))
    print(f"{name:22s} instr/cell {a[0]*32/cells:7.0f} ({100*a[0]/tot[0]:5.1f}%)  fp64 {a[1]*32/cells:6.0f}  other {(a[0]-a[1])*32/cells:6.0f}  samples {100*a[2]/tot[2]:5.1f}%  | {top}")
