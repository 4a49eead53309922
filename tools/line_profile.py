"""Dev tool: join `nvdisasm -g -c` line info with an ncu SASS-level CSV (tools/sass_hotspots.py input)
and print samples / executed warp instructions per CUDA source line of one kernel.
usage: line_profile.py <fused.dis> <kernel mangled substring> <ncu sass csv> <source file>"""
import csv, os, re, sys
dis, kname, ncsv, srcfile = sys.argv[1:5]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = None; inl = []; per_instr = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln, rest = m.group(1), int(m.group(2)), m.group(3)
        if "inlined at" in rest:
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', rest)
            cur = (f, ln, m2.group(1), int(m2.group(2)))
        else:
            cur = (f, ln, None, None)
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        per_instr.append(cur)
rows = list(csv.reader(open(ncsv))); hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}; d = rows[2:]
print("disasm instrs", len(per_instr), "ncu instrs", len(d))
src = open(srcfile).read().split("\n")
agg = {}
for k, r in enumerate(d):
    if k >= len(per_instr): break
    c = per_instr[k]
    if c is None: key = -1
    elif c[0].endswith(os.path.basename(srcfile)): key = c[1]
    elif c[2] and c[2].endswith(os.path.basename(srcfile)): key = c[3]
    else: key = -2
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[ix["# Samples"]]); a[1] += int(r[ix["Instructions Executed"]])
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
for key in sorted(agg):
    s, i = agg[key]
    if s * 200 < ts and i * 200 < ti: continue
    text = src[key - 1].strip()[:90] if key > 0 else "(other file / unknown)"
    print(f"{key:5d} samp {100*s/ts:5.1f}% instr {100*i/ti:5.1f}%  {text}")
