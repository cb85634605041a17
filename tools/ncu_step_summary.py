"""One JSON record per captured launch of an .ncu-rep (ncu --set full [+ --metrics sm__inst_executed_pipe_alu.sum]):
what bench.py's roofline needs from a profile of ONE step of the benchmark workload -- duration, DRAM bytes,
instructions, ALU-pipe instructions -- so that roofline.frac can be recomputed from committed data.
usage: ncu_step_summary.py <file.ncu-rep> <out.json> "<what was captured>" """
import csv, json, subprocess, sys

ALU_PEAK_PER_SM_CYCLE = 2.0  # ALU pipe: one warp instruction per 2 cycles per SM sub-partition, 4 sub-partitions (B300_MICROARCH.md "Pipe rates")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
    launches = []
    for r in rows[2:]:
        g = lambda k: num(r[col[k]]) if k in col else None
        u = lambda k: units[col[k]] if k in col else ""
        sms = g("launch__sm_count") or 148
        cyc = g("sm__cycles_active.avg")
        pct = g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
        alu_exact = g("sm__inst_executed_pipe_alu.sum")
        rec = dict(kernel=r[col["Kernel Name"]].replace("<unnamed>::", ""), grid=r[col["Grid Size"]], block=r[col["Block Size"]],
                   duration_us=g("gpu__time_duration.sum") * scale.get(u("gpu__time_duration.sum"), 1.0),
                   dram_read_bytes=g("dram__bytes_read.sum") * scale.get(u("dram__bytes_read.sum"), 1.0),
                   dram_write_bytes=g("dram__bytes_write.sum") * scale.get(u("dram__bytes_write.sum"), 1.0),
                   inst_executed=g("smsp__inst_executed.sum"), sm_cycles_active_avg=cyc, sm_count=sms,
                   alu_pipe_pct_of_peak_active=pct, issue_active_pct=g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   lsu_pipe_pct=g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                   tensor_pipe_pct=g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   registers=g("launch__registers_per_thread"), smem_dynamic_bytes=g("launch__shared_mem_per_block_dynamic"))
        if alu_exact is not None:
            rec["alu_pipe_inst"] = alu_exact
            rec["alu_pipe_inst_source"] = "sm__inst_executed_pipe_alu.sum"
        elif pct is not None and cyc is not None:
            rec["alu_pipe_inst"] = pct / 100.0 * ALU_PEAK_PER_SM_CYCLE * cyc * sms
            rec["alu_pipe_inst_source"] = "pct_of_peak_sustained_active x %.1f inst/clk/SM x sm__cycles_active.avg x SMs" % ALU_PEAK_PER_SM_CYCLE
        launches.append(rec)
    json.dump(dict(what=what, source=rep.split("/")[-1], alu_peak_inst_per_sm_cycle=ALU_PEAK_PER_SM_CYCLE, launches=launches),
              open(out, "w"), indent=1)
    for l in launches:
        print("%-60s %9.1f us  dram %7.1f MB  inst %.3g  alu %.3g (%s%%)" % (l["kernel"][:60], l["duration_us"],
              (l["dram_read_bytes"] + l["dram_write_bytes"]) / 1e6, l["inst_executed"], l.get("alu_pipe_inst", 0), l["alu_pipe_pct_of_peak_active"]))


if __name__ == "__main__":
    main()
