"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and average duration, share.
usage: launch_list_by_kernel.py <ncu csv log> > profiles/<name>.csv"""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, im = hdr.index("Kernel Name"), hdr.index("Metric Value")
ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.defaultdict(lambda: [0, 0.0, 0, 0])
for r in rows[1:]:
    k = r[ik].replace("void ", "").split("(")[0]
    if "cub::" in k:
        k = k.split("<")[0]
    try:
        d = float(r[im].replace(",", ""))
    except ValueError:
        continue
    a = agg[k]
    a[0] += 1; a[1] += d
    try:
        a[2] = max(a[2], int(r[ig].strip("() ").split(",")[0])); a[3] = max(a[3], int(r[ib].strip("() ").split(",")[0]))
    except ValueError:
        pass
tot = sum(v[1] for v in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ns", "share", "avg_ns", "max_grid_x", "max_block_x"])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, v[0], round(v[1], 1), round(v[1] / tot, 4), round(v[1] / v[0], 2), v[2], v[3]])
