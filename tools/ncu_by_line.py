"""Aggregates an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) by CUDA source line (dev tool).
usage: ncu_by_line.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; cur = None; agg = {}; fname = None; src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iB = hdr.index("stall_barrier"); iL = hdr.index("stall_long_sb"); iSh = hdr.index("stall_short_sb"); continue
    if hdr is None: continue
    if r[0] != "":
        cur = (fname, int(r[0])); src[cur] = r[1]
        continue
    if cur is None or len(r) <= iS: continue
    a = agg.setdefault(cur, [0, 0, 0, 0, 0])
    for q, ix in enumerate((iS, iI, iB, iL, iSh)):
        try: a[q] += int(r[ix] or 0)
        except ValueError: pass
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print("samples", tot, "warp-insts", toti)
print("file:line  samples%  inst%  barrier% long_sb% short_sb%")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%s:%d %5.1f %5.1f | %4.1f %4.1f %4.1f | %s" % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / toti, 100 * a[2] / tot, 100 * a[3] / tot, 100 * a[4] / tot, src[k].strip()[:110]))
