"""profiles/k1_traffic.json (read by bench.py as roofline.traffic) from the ncu CSV of the K1 launches of one step:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none -k regex:kv_attn --csv --log-file gpurun_out/k1_traffic.csv python profiles/run_step_for_ncu.py
    python profiles/k1_traffic_from_ncu.py gpurun_out/k1_traffic.csv profiles/r2/k1_traffic_metrics.csv > profiles/k1_traffic.json
"""
import collections
import csv
import json
import sys


def main(path, kept_as):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(dict)
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3}.get(u, 1)
        per[r["ID"]][r["Metric Name"]] = v * mul
        per[r["ID"]]["kernel"] = r["Kernel Name"]
    tot = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in per.values())
    us = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    by_kernel = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        k = d["kernel"].split("(")[0].replace("void ", "")
        by_kernel[k][0] += 1
        by_kernel[k][1] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        by_kernel[k][2] += d.get("gpu__time_duration.sum", 0)
    print(json.dumps({"dram_bytes_per_step": tot, "dram_bytes_per_launch_avg": tot / max(len(per), 1), "launches": len(per),
                      "ncu_time_us_sum": us,
                      "per_kernel": {k: {"launches": v[0], "dram_mb_per_launch": v[1] / v[0] / 1e6, "us_per_launch": v[2] / v[0]}
                                     for k, v in by_kernel.items()},
                      "source": f"{kept_as} (ncu dram__bytes_read.sum + dram__bytes_write.sum over the K1 launches of one eager "
                                "step, config 2, round 2)"}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
