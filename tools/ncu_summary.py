"""Summarise one kernel of an .ncu-rep (ncu --set full) into the text format kept under profiles/.
usage: ncu_summary.py <file.ncu-rep> "<header line>" [kernel index in the report, default 0] > profiles/<name>.txt"""
import csv, subprocess, sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, header = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
    col = {h: i for i, h in enumerate(hdr)}
    print(header)
    for k in KEYS:
        if k in col:
            print("%-75s %s %s" % (k, vals[col[k]], units[col[k]]))
    print("\nwarp stall reasons (per issue-active cycle):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            v = float(vals[col[h]])
            if v >= 0.05:
                print("  %-85s %f" % (h, v))


if __name__ == "__main__":
    main()
