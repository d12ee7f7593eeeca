"""Summarise an .ncu-rep (read offline with `ncu -i`): key raw metrics per kernel launch -> markdown/JSON."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_xu",
    "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "smsp__inst_executed_pipe_fp64", "sm__inst_executed.avg.per_cycle_elapsed",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k) for k in KEYS) or "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                d[h + (" [" + units[i] + "]" if units[i] else "")] = r[i]
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
