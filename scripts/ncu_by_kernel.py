"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list.
usage: python scripts/ncu_by_kernel.py gpurun_out/launches.csv [substring ...]   (substrings: also print those launches one by one)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, mi, ii, gi = (hdr.index(c) for c in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Grid Size"))
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[ii], {"k": r[ki].split("(")[0].replace("hq::", "").replace("void ", ""), "grid": r[gi]})
    try:
        d[r[mi]] = float(r[vi].replace(",", ""))
    except ValueError:
        pass
agg = collections.defaultdict(lambda: [0, 0.0])
for d in L.values():
    agg[d["k"]][0] += 1
    agg[d["k"]][1] += d["gpu__time_duration.sum"]
tot = sum(v[1] for v in agg.values())
print(f"{len(L)} launches, {tot / 1e3:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} {v[0]:4d} launches {v[1] / 1e3:9.1f} us {100 * v[1] / tot:5.1f} %")
for d in L.values():
    if any(s in d["k"] for s in sys.argv[2:]):
        t = d["gpu__time_duration.sum"]
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        print(f"  {d['k']:28s} grid {d['grid']:14s} {t / 1e3:8.1f} us {b / 1e6:8.1f} MB {b / t:7.0f} GB/s")
