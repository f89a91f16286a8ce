"""Per-source-line view of one kernel in an ncu source page (dev tool): samples, warp instructions, shared-memory wavefronts, global L1 tag
requests, top stall reasons.  usage: ncu_by_line2.py src.csv kernel_substring [top] [sort_column]"""
import csv, sys
csv.field_size_limit(1 << 30)
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40; sort = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hdr = None; cur = None; agg = {}; fname = None; src = {}; infn = False
cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Tag Requests Global", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_lg"]
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": infn = want in r[1]; continue
    if r[0] == "Line No": hdr = r; ix = [hdr.index(c) for c in cols]; continue
    if hdr is None or not infn: continue
    if r[0] != "":
        cur = (fname, int(r[0])); src[cur] = r[1]; continue
    if cur is None: continue
    a = agg.setdefault(cur, [0] * len(cols))
    for q, i in enumerate(ix):
        try: a[q] += int(r[i] or 0)
        except (ValueError, IndexError): pass
tot = [sum(a[q] for a in agg.values()) or 1 for q in range(len(cols))]
print("totals", dict(zip(cols, tot)))
print("file:line  samp% inst% smemWF% gTag% | bar long short wait mio math lg (% of samples)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][sort])[:top]:
    print("%s:%d %5.1f %5.1f %5.1f %5.1f | %4.1f %4.1f %4.1f %4.1f %4.1f %4.1f %4.1f | %s" % ((k[0], k[1]) + tuple(100 * a[q] / tot[q] for q in range(4)) + tuple(100 * a[q] / tot[0] for q in range(4, 11)) + (src[k].strip()[:90],)))
