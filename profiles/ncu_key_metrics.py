"""Key utilisation figures per launch from `ncu -i X.ncu-rep --page raw --csv`: pipes, issue slots, stall reasons, memory."""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
    "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main(path, contains=""):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    stall = [h for h in header if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")] or \
            [h for h in header if "warp_issue_stalled" in h and "per_warp_active" in h]
    stall2 = [h for h in header if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if contains and contains not in name:
            continue
        print("=" * 100)
        print(name[:120], "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
        for m in WANT:
            if m in idx:
                print(f"  {m:90s} {r[idx[m]]:>14s} {units[idx[m]]}")
        top = []
        for h in stall2 or stall:
            try:
                top.append((float(r[idx[h]].replace(",", "")), h))
            except ValueError:
                pass
        for v, h in sorted(top, reverse=True)[:10]:
            print(f"  {h:90s} {v:14.3f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
