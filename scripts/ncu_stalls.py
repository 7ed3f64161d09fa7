"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = {h: 0 for h in stall_cols}
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2:
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    for h in stall_cols:
        try:
            tot[h] += int(r[col[h]] or 0)
        except (ValueError, IndexError):
            pass
    lines.append((samples, r[col["Address"]], r[col["Source"]], int(r[col["Instructions Executed"]] or 0),
                  {h: r[col[h]] for h in stall_cols if col[h] < len(r) and r[col[h]] not in ("", "0")}))
S = sum(tot.values())
print("total samples", sum(l[0] for l in lines), "stall sum", S)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {h:24s} {v:8d} {100.0 * v / max(S, 1):5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("hottest lines:")
for s, addr, src, ex, st in sorted(lines, key=lambda l: -l[0])[:n]:
    print(f"  {s:6d} {ex:9d} {src[:70]:70s} {st}")
