"""Developer helper: top stalled SASS lines of an ncu --page source --csv dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci['# Samples']]
tot = sum(float(r[ci['# Samples']] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -float(r[ci['# Samples']] or 0))[:n]:
    s = float(r[ci['# Samples']] or 0)
    top = sorted(((float(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{s:7.0f} {s / tot * 100:5.1f}% line{idx[id(r)]:5d} {r[ci['Source']].strip()[:70]:70s} {top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f}")
print('total samples', tot)
