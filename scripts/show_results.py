"""Print the numbers of the last GPU call (gpurun_out/): bench lines and ncu launch lists."""
import collections
import csv
import json
import os
import sys

GO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
for fn in sys.argv[1:] or ["bench_stride15.json", "bench_r01.json", "launches_small.csv", "launches_r01.csv"]:
    p = os.path.join(GO, fn)
    if not os.path.isfile(p) or os.path.getsize(p) == 0:
        continue
    if fn.endswith(".json"):
        d = json.load(open(p))
        print(fn, "value", round(d["value"], 3), "ms/step", round(d["ms_per_step"]), "clocks", d["clocks"], "launches", d["gpu_launches"])
        print("  solver", d["solver"], "roofline us", round(d["roofline"]["us_per_launch"], 1), "frac", round(d["roofline"]["frac"], 3))
        if d.get("e2e"):
            print("  e2e", round(d["e2e"]["value"], 3), "same-mesh", round(d["same_mesh_as_cpu_baseline"]["value"], 1),
                  "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("measured_at_sample"))
    else:
        rows = [r for r in csv.reader(open(p)) if len(r) > 10]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            k = r[ki].split("(")[0]
            agg[k][0] += 1
            agg[k][1] += v
        print(fn)
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"  {k:60s} {n:5d} {t / n / 1e3:8.1f} us")
