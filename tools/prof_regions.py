"""Dev tool: per-region summary of an ncu report's SASS page (samples, executed warp instructions, top stall reasons),
plus the waits on every mbarrier / named barrier and the shared-memory wavefronts by instruction class.
usage: python tools/prof_regions.py <report.ncu-rep> [lines per region=100]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}; d = rows[2:]
S, I, SRC = ix["# Samples"], ix["Instructions Executed"], ix["Source"]
cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[S]) for r in d); print("kernel", rows[0][1][:80]); print("total samples", tot, "warp instr %.1fM" % (sum(int(r[I]) for r in d) / 1e6))
for a in range(0, len(d), step):
    b = min(len(d), a + step); s = sum(int(r[S]) for r in d[a:b]); i = sum(int(r[I]) for r in d[a:b])
    t = {c: sum(int(r[ix[c]]) for r in d[a:b]) for c in cols}; ss = max(1, sum(t.values()))
    top = {k[6:]: round(100 * v / ss) for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:4]}
    print(f"{a:5d}-{b:5d} samples {s:6d} ({100*s/tot:4.1f}%) instr {i/1e6:7.1f}M {top}  {d[a][SRC].strip()[:44]}")
print("-- waits")
for i, r in enumerate(d):
    if "TRYWAIT" in r[SRC] or "BAR.SYNC" in r[SRC]:
        s = sum(int(x[S]) for x in d[i:i + 4])
        if s * 200 > tot: print(i, r[SRC].strip()[:70].ljust(70), s)
W = ix.get("L1 Wavefronts Shared")
if W is not None:
    agg = {}
    for r in d:
        w = int(r[W] or 0)
        if w:
            toks = r[SRC].split(); op = toks[1] if toks[0].startswith("@") else toks[0]
            agg[op] = agg.get(op, 0) + w
    print("-- shared wavefronts (M):", {k: round(v / 1e6, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
