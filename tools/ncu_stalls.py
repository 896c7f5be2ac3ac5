"""Top SASS instructions by stall samples with their source line. usage: ncu_stalls.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
data = [r for r in rows[2:ends[1]] if len(r) == len(hdr)]
iS = hdr.index("Source"); iSm = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[iSm]) for r in data)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iSm]))[:n]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    prev = data[i-1][iS].strip()[:40] if i else ''
    print(f"{i:5d} {100*int(r[iSm])/tot:5.2f}%  x{int(r[iI]):9d}  {r[iS].strip()[:60]:60s} | " + ", ".join(f"{n}={v}" for v, n in st))
