#!/usr/bin/env python
"""Summarise an `ncu --csv --log-file` launch list: launches, total / average duration and share per kernel, DRAM bytes
when those metrics were collected.   python tools/ncu_summarise.py launches.csv [out.csv]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    name = name.replace("void ", "").replace("ancuts::", "")[-70:]
    m = r[idx["Metric Name"]]; v = float(r[idx["Metric Value"]].replace(",", "")); unit = r[idx["Metric Unit"]]
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0))
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a["n"] += 1; a["us"] += v
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        a["rd" if "read" in m else "wr"] += v
tot = sum(a["us"] for a in agg.values()) or 1.0
lines = ["kernel,launches,total_us,avg_us,share_pct,dram_read_gb,dram_write_gb"]
for k, a in sorted(agg.items(), key=lambda t: -t[1]["us"]):
    n = max(a["n"], 1)
    lines.append(f"{k},{a['n']},{a['us']:.1f},{a['us'] / n:.2f},{100 * a['us'] / tot:.1f},{a['rd'] / 1e9:.3f},{a['wr'] / 1e9:.3f}")
text = "\n".join(lines)
print(f"# launches {sum(a['n'] for a in agg.values())}, total {tot / 1e3:.1f} ms")
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# launches {sum(a['n'] for a in agg.values())}, total {tot / 1e3:.1f} ms\n" + text + "\n")
