"""Print one step's kernel sequence from an ncu launch-list CSV (gpu__time_duration.sum)."""
import csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        t = float(x["Metric Value"].replace(",", ""))
        if x["Metric Unit"] in ("ns", "nsecond"):
            t /= 1000
        rows.append((x["Kernel Name"], x["Grid Size"], t))
idx = [i for i, x in enumerate(rows) if "normalize" in x[0]]
a = idx[1] if len(idx) > 1 else idx[0]
b = idx[2] if len(idx) > 2 else len(rows)
tot = 0.0
for name, grid, t in rows[a:b]:
    tot += t
    print(name.split("::")[-1][:40].ljust(40), grid.ljust(14), f"{t:8.1f}")
print("total us", round(tot, 1))
