"""Per-kernel table of an ncu launch list (gpu__time_duration.sum csv): launches of the LAST iteration of run_vol.py.
Usage: python scripts/launch_table.py gpurun_out/r2_launches_x.csv [iterations=3]"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v))
it = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = len(rows) // it
last = rows[-n:]
agg = collections.OrderedDict()
for k, v in last:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in last)
print(f"{len(last)} launches in the last iteration, {tot:.1f} us of kernel time")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:10.1f} us  {100 * v / tot:5.1f} %  x{c:<3d} {k}")
