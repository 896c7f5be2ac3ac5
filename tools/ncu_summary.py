"""Summarise an .ncu-rep (one kernel) into the numbers DESIGN.md / profiles/ quote.  Usage: ncu_summary.py rep [cells]"""
import csv, collections, re, subprocess, sys, io
rep = sys.argv[1]; cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
m = dict(zip(h, v)); units = dict(zip(h, u))
def g(k):
    return m.get(k, "n/a")
print("kernel:", g("Kernel Name"), " grid", g("Grid Size"), "block", g("Block Size"))
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in m: print(f"  {k} = {m[k]} {units.get(k,'')}")
st = {k: float(x) for k, x in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
print("  stalls per issue:", ", ".join(f"{k.split('stalled_')[1].replace('_per_issue_active.ratio','')}={x:.2f}" for k, x in sorted(st.items(), key=lambda t: -t[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
data = rows[2:ends[1]]
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed")
ops = collections.Counter(); tot = 0
for rr in data:
    if len(rr) <= max(iI, iS) or not rr[iI].strip().isdigit(): continue
    n = int(rr[iI]); mm = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', rr[iS]); op = mm.group(2) if mm else '?'
    ops[op] += n; tot += n
fp = sum(n for o, n in ops.items() if o.split('.')[0] in ("DFMA", "DMUL", "DADD"))
print(f"  warp-instructions {tot}, fp64 (DFMA+DMUL+DADD) {fp} = {100*fp/tot:.1f}%")
if cells: print(f"  per cell-stage: {tot*32/cells:.0f} thread-instr, {fp*32/cells:.0f} fp64")
print("  top opcodes:", ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(14)))
