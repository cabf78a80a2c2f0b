"""Turns the ncu CSV of one eager step (gpu__time_duration + dram bytes per launch) into profiles/launches_rN_summary.txt
and profiles/conv_traffic_rN.json (average DRAM traffic per conv_gemm launch, read by bench.py's roofline block, stamped with the
commit it was measured at):   python profiles/traffic_summary.py launches.csv summary.txt traffic.json [commit]"""
import collections
import csv
import json
import sys

src, out_txt, out_json = sys.argv[1:4]
commit = sys.argv[4] if len(sys.argv) > 4 else None
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    per[(int(r["ID"]), r["Kernel Name"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (_, name), mt in sorted(per.items()):
    short = name.split("(")[0].replace("void ", "").replace("k2::<unnamed>::", "").replace("unnamed>::", "")
    fam = "conv_gemm" if short.startswith("conv_gemm") else short
    a = agg[fam]
    a[0] += 1
    a[1] += mt.get("gpu__time_duration.sum", 0.0)
    a[2] += mt.get("dram__bytes_read.sum", 0.0) + mt.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
with open(out_txt, "w") as f:
    f.write("kernel family                      launches   time_ms   share   dram_MB_total  dram_MB/launch\n")
    for fam, (n, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{fam:34s} {n:8d} {ns / 1e6:9.3f} {100 * ns / tot:6.1f}% {by / 1e6:13.1f} {by / 1e6 / n:12.2f}\n")
    f.write(f"total {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms (ncu: cold-cache, serialised)\n")
n, ns, by = agg["conv_gemm"]
json.dump({"dram_bytes_per_launch": by / n, "launches": n, "share_of_step_time": ns / tot, "commit": commit,
           "note": f"mean of dram__bytes_read.sum + dram__bytes_write.sum over the conv_gemm launches of one step (ncu, "
                   f"{src.split('/')[-1]} under profiles/); algorithmic bytes per launch are in profiles/README.md"},
          open(out_json, "w"))
print(open(out_txt).read())
