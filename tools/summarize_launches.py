"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share of device time."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1; c[1] += ns
total = sum(v[1] for v in tot.values())
print(f"{'kernel':110s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:110]:110s} {n:8d} {ns/1e6:10.3f} {ns/n/1e3:10.1f} {100*ns/total:6.2f}%")
print(f"{'TOTAL':110s} {sum(v[0] for v in tot.values()):8d} {total/1e6:10.3f}")
