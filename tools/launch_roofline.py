"""Per-launch DRAM roofline table from an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
log (long format: one row per launch x metric).  Ranks launches by recoverable time = duration - bytes / HBM peak.

    python tools/launch_roofline.py gpurun_out/launch_dram.csv [hbm_gbs] > profiles/rNN_launch_roofline.txt
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6483.6
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = defaultdict(dict)
for r in rows[1:]:
    try:
        lid = int(r[ix["ID"]])
    except ValueError:
        continue
    d = launches[lid]
    d["name"] = r[ix["Kernel Name"]]
    d["grid"], d["block"] = r[ix["Grid Size"]], r[ix["Block Size"]]
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[m] = val * scale
tot = sum(d.get("gpu__time_duration.sum", 0) for d in launches.values())
print(f"# {len(launches)} launches, {tot:.1f} us total (serialised, cold-cache: compare SHARES); HBM peak {peak:.0f} GB/s")
print(f"# {'id':>4} {'us':>8} {'MB':>8} {'GB/s':>7} {'%peak':>6} {'recov_us':>8}  kernel")
out = []
for lid, d in sorted(launches.items()):
    us = d.get("gpu__time_duration.sum", 0.0)
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    gbs = by / (us * 1e-6) / 1e9 if us else 0.0
    rec = us - by / (peak * 1e9) * 1e6
    out.append((lid, us, by, gbs, rec, d["name"], d["grid"]))
    print(f"  {lid:4d} {us:8.1f} {by / 1e6:8.1f} {gbs:7.0f} {100 * gbs / peak:6.1f} {rec:8.1f}  {d['name'][:60]} {d['grid']}")
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for lid, us, by, gbs, rec, name, grid in out:
    k = name.split("(")[0]
    agg[k][0] += us
    agg[k][1] += by
    agg[k][2] += rec
    agg[k][3] += 1
print("\n## per kernel, by recoverable time (duration - DRAM bytes / peak)")
for k, (us, by, rec, n) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"  {rec:8.1f} us recoverable of {us:8.1f} us ({100 * us / tot:4.1f}%)  {n:3d}x  {by / 1e6:9.1f} MB  {by / (us * 1e-6) / 1e9 if us else 0:6.0f} GB/s  {k[:70]}")
