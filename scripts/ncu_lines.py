"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass` output.
usage: python scripts/ncu_lines.py src.csv [section_index] [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 5]
s = starts[sec]; e = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
hdr = rows[s]
iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
agg = collections.OrderedDict(); total = 0; tot_s = 0
for r in rows[s + 1:e]:
    if len(r) < len(hdr): continue
    try: n = int(r[iI]); sm = int(r[iS]); th = int(r[iT])
    except ValueError: continue
    key = (r[0], r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += n; a[1] += sm; a[2] += th
    total += n; tot_s += sm
print("total inst", total, "samples", tot_s)
for (ln, src), (n, sm, th) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>5} {100*n/total:5.1f}% inst {100*sm/max(tot_s,1):5.1f}% smp  lanes {th/max(n,1):4.1f}  {src}")
