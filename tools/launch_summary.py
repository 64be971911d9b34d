"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/launches_rNN.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]])
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name][0] += 1
    tot[name][1] += v
all_ms = sum(v[1] for v in tot.values())
print(f"launch list: {sys.argv[1]}  ({sum(v[0] for v in tot.values())} launches, {all_ms:.2f} ms of kernel time; "
      "per-launch times are cold-cache and serialised under ncu)")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>12s} {'avg ms':>10s} {'share':>8s}")
for name, (n, ms) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{name:60s} {n:9d} {ms:12.3f} {ms / n:10.4f} {100 * ms / all_ms:7.2f}%")
