#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]; ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    u = r[ui]
    ms = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v if u in ("msecond", "ms") else v * 1e3
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"void |<unnamed>::|ab::|\(anonymous namespace\)::", "", name)
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: total kernel time {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
print("kernel,launches,total_ms,share,avg_us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1]:.3f},{v[1] / tot:.4f},{v[1] / v[0] * 1e3:.1f}")
