"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean us, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    agg.setdefault((r[ki][:72], r[gi]), []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'grid':>14s} {'n':>5s} {'mean us':>9s} {'share':>7s}")
for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} {g:>14s} {len(v):5d} {sum(v) / len(v):9.1f} {sum(v) / tot * 100:6.1f}%")
print(f"total {tot:.0f} us over {sum(len(v) for v in agg.values())} launches (cold-cache, serialised: compare shares)")
