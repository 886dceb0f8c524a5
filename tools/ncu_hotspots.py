"""tools/ncu_hotspots.py <source.csv from `ncu --page source --csv --print-source cuda,sass`> [top]:
instructions executed and stall samples per CUDA source line (all files of the report)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
agg = collections.OrderedDict()
fname = None
hdr = None
mode = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); src = 1
        continue
    if hdr is None or fname is None or not r[0].isdigit():
        continue
    try:
        n = int(r[ci] or 0); smp = int(r[cs] or 0)
    except Exception:
        continue
    if n == 0 and smp == 0:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[src].strip()[:110]])
    a[0] += n; a[1] += smp
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total inst", tot, "samples", tots)
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %18s:%-4d %s" % (100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(tots, 1), f, ln, v[2]))
