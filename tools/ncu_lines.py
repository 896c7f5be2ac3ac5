"""Attribute warp-stall samples / executed instructions of a profiled kernel to CUDA source lines.
usage: ncu_lines.py rep.ncu-rep lib.so kernel_mangled_substr"""
import csv, collections, re, subprocess, sys, io, os, tempfile
rep, lib, kname = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if "sm_100a" in f]
dis = ""
for c in cub:
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, c)], capture_output=True, text=True).stdout
lines = dis.split("\n")
start = None
for i, l in enumerate(lines):
    if l.strip().startswith(".text.") and kname in l: start = i; break
instrs = []; cur = ("?", 0, [])
for l in lines[start + 1:]:
    if l.strip().startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)): chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
        cur = chain; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.+?);", l)
    if m: instrs.append((cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
data = [r for r in rows[2:ends[1]] if len(r) == len(hdr)]
iI = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples")
assert len(data) == len(instrs), (len(data), len(instrs))
byline = collections.Counter(); bylineI = collections.Counter(); tot = 0; totI = 0
for (chain, txt), r in zip(instrs, data):
    s = int(r[iSm]); n = int(r[iI])
    outer = chain[-1] if isinstance(chain, list) else ("?", 0)   # outermost frame = line in the kernel body
    byline[outer] += s; bylineI[outer] += n; tot += s; totI += n
print("samples", tot, "instr", totI)
for (f, ln), s in sorted(byline.items(), key=lambda t: -t[1])[:40]:
    print(f"{100*s/tot:5.1f}% samples  {100*bylineI[(f,ln)]/totI:5.1f}% instr  {f}:{ln}")
