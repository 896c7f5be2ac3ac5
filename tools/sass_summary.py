"""Per-kernel SASS opcode histogram of libmwb200.so (sm_100a cubins): the evidence that the Blackwell paths are real
(UTMALDG = TMA box loads, UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, SYNCS = mbarrier) -- profiles/sass_summary.txt.
usage: python tools/sass_summary.py [lib] > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "miniweatherml_b200", "libmwb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("libmwb200.so: cubins for", ", ".join(archs), "| kernels:", len(hist))
KEY = ["UTMALDG", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "HMMA", "DFMA", "DMUL", "DADD", "LDS", "STS", "SHFL", "LDG", "STG", "BAR", "MUFU"]
tot = collections.Counter()
for k, h in hist.items():
    n = sum(h.values())
    fam = collections.Counter()
    for op, c in h.items():
        fam[op.split(".")[0]] += c
    for f, c in fam.items():
        tot[f] += c
    key = ", ".join("%s %d" % (f, fam[f]) for f in KEY if fam.get(f))
    top = ", ".join("%s %d" % (o, c) for o, c in h.most_common(6))
    print("\n%s\n  instructions %d (%.1f KB) | %s\n  top: %s" % (k, n, n * 16 / 1024.0, key, top))
print("\nwhole library:", ", ".join("%s %d" % (f, tot[f]) for f in KEY if tot.get(f)))
print("legacy tensor path (HMMA / mma.sync): %d instructions; Hopper wgmma (HGMMA): %d" % (tot.get("HMMA", 0), tot.get("HGMMA", 0)))
