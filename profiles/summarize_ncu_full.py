"""Per-launch table from an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`:
duration, tensor-pipe and issue-slot utilisation, DRAM bytes, L2 hit rate, registers, shared memory, achieved occupancy.

    ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > gpurun_out/step_full_raw.csv
    python profiles/summarize_ncu_full.py gpurun_out/step_full_raw.csv > profiles/r2/step_full_kernels.txt
"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1e-3),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("issue%", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    ("dramRdMB", "dram__bytes_read.sum", 1e-6),
    ("dramWrMB", "dram__bytes_write.sum", 1e-6),
    ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2hit%", "lts__t_sector_hit_rate.pct", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("smemKB", "launch__shared_mem_per_block_dynamic", 1e-3),
    ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("waves", "launch__waves_per_multiprocessor", 1),
]


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    for h, i in list(idx.items()):          # some metrics carry a section prefix ("FBSP.TriageCompute.dram__throughput...")
        if "." in h and h.split(".", 2)[-1] not in idx and h.count(".") >= 3:
            idx.setdefault(h.split(".", 2)[-1], i)
    name_i, grid_i = idx.get("Kernel Name"), idx.get("Grid Size")
    scale_unit = {"nsecond": 1.0, "ns": 1.0, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print(f"{'kernel':46s} {'grid':>14s} " + " ".join(f"{c[0]:>8s}" for c in COLS))
    for r in rows[2:]:
        if not r or name_i is None:
            continue
        nm = re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("l2d::", "").replace("<unnamed>::", "")
        out = []
        for label, metric, mul in COLS:
            i = idx.get(metric)
            v = num(r[i]) if i is not None and i < len(r) else None
            if v is None:
                out.append(f"{'-':>8s}")
                continue
            u = units[i] if i < len(units) else ""
            if label in ("us",):
                v = v * scale_unit.get(u, 1.0) * 1e-3      # -> microseconds
            elif label in ("dramRdMB", "dramWrMB", "smemKB"):
                v = v * scale_unit.get(u, 1.0) * mul
            out.append(f"{v:8.1f}")
        print(f"{nm[:46]:46s} {r[grid_i] if grid_i is not None else '':>14s} " + " ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
