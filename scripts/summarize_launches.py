"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list: per-kernel count, time, share, DRAM GB/s.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [last_n_launches]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    rows.append(r)
# group by launch ID
by_id = collections.OrderedDict()
for r in rows:
    k = r["ID"]
    d = by_id.setdefault(k, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[m] = v * mult
launches = list(by_id.values())
if len(sys.argv) > 2:
    launches = launches[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for d in launches:
    nm = re.sub(r"\(.*", "", d["name"])
    nm = re.sub(r"<.*", "", nm)
    a = agg.setdefault(nm, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
for nm, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    gbs = a[2] / (a[1] * 1e-6) / 1e9 if a[1] > 0 else 0
    print(f"{nm:48s} n={a[0]:4d} us={a[1]:10.1f} share={100*a[1]/tot:5.1f}%  dram={a[2]/1e6:9.1f} MB  {gbs:7.0f} GB/s")
print(f"total us {tot:.1f}, launches {len(launches)}")
