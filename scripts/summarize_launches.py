"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, data = None, []
for r in rows:
  if len(r) > 5 and r[0] == "ID": hdr = r; continue
  if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
agg = collections.defaultdict(lambda: [0, 0.0])
for d in data:
  name = re.sub(r"^void ", "", re.sub(r"\(.*", "", d["Kernel Name"]))
  name = re.sub(r"(mcba::|<unnamed>::)", "", name)
  v = float(d["Metric Value"].replace(",", ""))
  if d["Metric Unit"] in ("ns", "nsecond"): v /= 1e3
  elif d["Metric Unit"] in ("ms", "msecond"): v *= 1e3
  agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print(f"{k[:70]:70s} {v[0]:8d} {v[1]:10.1f} {v[1]/v[0]:8.2f} {100*v[1]/tot:5.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")
