"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/r02_launches.csv "<command that was profiled>" > profiles/r02_launches_summary.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = {}
for r in rows:
    name, val = r[4], float(r[-1])
    short = re.sub(r"\(.*", "", name).split("::")[-1]
    d = tot.setdefault(short, {"launches": 0, "total_ms": 0.0})
    d["launches"] += 1
    d["total_ms"] += val / 1e6
total = sum(d["total_ms"] for d in tot.values())
for d in tot.values():
    d["share"] = d["total_ms"] / total
print(json.dumps({"command": sys.argv[2] if len(sys.argv) > 2 else "", "total_ms": total, "kernels": tot}, indent=1))
