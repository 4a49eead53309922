"""Dev tool: summarise `ncu --page source --csv --print-source sass` output: samples and executed
instructions per SASS line, grouped into contiguous hot regions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total samples", tot_s, "warp instr", tot_i, "sass lines", len(data))
agg = {c: sum(int(r[ix[c]]) for r in data) for c in stall_cols}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
mode = sys.argv[2] if len(sys.argv) > 2 else "top"
if mode == "top":
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]
    for i in sorted(top):
        r = data[i]
        st = {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]])}
        print(i, r[ix["Source"]].strip()[:70].ljust(70), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), st)
else:  # dump a range
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, b):
        r = data[i]
        st = {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]])}
        print(i, r[ix["Source"]].strip()[:80].ljust(80), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), st)
