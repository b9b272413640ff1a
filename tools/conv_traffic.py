"""profiles/r02_conv_traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
of a bench.py run: DRAM bytes per conv_tc_kernel launch (bench.py's roofline.traffic) and the kernel's share of device time."""
import csv, json, re, sys
path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
per = {}
for row in csv.DictReader(lines):
    kid = row["ID"]
    d = per.setdefault(kid, {"name": row["Kernel Name"]})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if row["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[row["Metric Name"]] = v * mult
conv = [d for d in per.values() if "conv_tc_kernel" in d["name"]]
tot_us = sum(d.get("us", 0.0) for d in per.values())
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in conv) / max(len(conv), 1)
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in conv) / max(len(conv), 1)
res = {"kernel": "lb::conv_tc_kernel", "launches": len(conv), "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
       "traffic_bytes_per_launch": rd + wr, "avg_launch_us_under_ncu": sum(d.get("us", 0.0) for d in conv) / max(len(conv), 1),
       "share_of_device_time_under_ncu": sum(d.get("us", 0.0) for d in conv) / max(tot_us, 1e-9),
       "source": " ".join(sys.argv[3:]) or path}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
