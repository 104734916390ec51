"""Join ncu's per-SASS-instruction counts with nvdisasm line info: instructions / stall samples per source line.
usage: ncu_by_line.py <ncu source csv> <nvdisasm -g -c output> <kernel name substring> [top]"""
import csv, re, sys
from collections import defaultdict
src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1) nvdisasm: offset -> (file, line), for the wanted function
off2line, cur, infn = {}, None, False
for ln in open(dis, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        infn = kname in m.group(1); cur = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > ie and r[ia].startswith("0x")]
base = int(data[0][ia], 16)
inst, samp, stalls = defaultdict(float), defaultdict(float), defaultdict(lambda: defaultdict(float))
for r in data:
    key = off2line.get(int(r[ia], 16) - base, ("?", 0))
    inst[key] += float(r[ie] or 0); samp[key] += float(r[isamp] or 0)
    for i, h in stall_cols:
        if i < len(r) and r[i]: stalls[key][h] += float(r[i])
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
print("by file:")
byf = defaultdict(lambda: [0.0, 0.0])
for k in inst: byf[k[0]][0] += inst[k]; byf[k[0]][1] += samp[k]
for f, (a, b) in byf.items(): print(f"  {f:24s} inst {100*a/ti:5.1f}%  samples {100*b/ts:5.1f}%")
print("top lines by stall samples:")
for k in sorted(samp, key=lambda k: -samp[k])[:top]:
    st = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
    print(f"  {k[0]:18s}:{k[1]:4d}  inst {100*inst[k]/ti:5.1f}%  samples {100*samp[k]/ts:5.1f}%  " + ", ".join(f"{h[6:]}={v:.0f}" for h, v in st))
