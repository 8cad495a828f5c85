"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/ (the .ncu-rep itself is scratch)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# summary of {rep} ({len(data)} launches; ncu --set full --clock-control none; per-launch values)"]
    name_col = hdr.index("Kernel Name")
    lines.append("kernel: " + " | ".join(sorted(set(r[name_col] for r in data))))
    for k in KEYS + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") or h in ("smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__pcsamp_sample_count")]:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k:80s} {units[i]:16s} " + "  ".join(r[i] for r in data))
    text = "\n".join(lines) + "\n"
    if out: open(out, "w").write(text)
    else: print(text)
if __name__ == "__main__":
    main(*sys.argv[1:])
