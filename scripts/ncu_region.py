"""Print SASS lines (address order) of an `ncu --page source --csv` dump with per-line samples; optional address window."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
seen = set(); out = []
for r in rows[2:]:
    if len(r) < 10: continue
    a = r[col['Address']]
    if a in seen: continue
    seen.add(a)
    try: ai = int(a, 16)
    except ValueError: continue
    out.append((ai, r[col['Source']], int(r[col['# Samples']] or 0), int(r[col['Instructions Executed']] or 0)))
base = out[0][0]; tot = 0
for ai, src, s, ex in out:
    off = ai - base
    if lo <= off <= hi:
        tot += s
        print(f"{off:5x} {s:5d} {ex:8d} {src[:100]}")
print("window samples", tot, "all", sum(o[2] for o in out))
