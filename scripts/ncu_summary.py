"""Compact per-launch summary of an .ncu-rep (needs `ncu` on PATH; no GPU): one CSV row per profiled launch with
the counters DESIGN.md quotes.    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv, subprocess, sys

KEEP = [
    ("gpu__time_duration.sum", "duration_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "ctas_per_sm_by_regs"),
    ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__cycles_active.avg", "smsp_cycles_active"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors"), ("lts__t_sectors_srcunit_tex_op_write.sum", "l2_write_sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "math_pipe_throttle", "wait", "not_selected", "barrier", "mio_throttle",
          "lg_throttle", "no_instruction", "dispatch_stall", "branch_resolving", "membar", "sleeping"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [k for _, k in KEEP] + ["stall_" + s for s in STALLS])
        for r in data:
            name = r[col["Kernel Name"]]
            vals = []
            for m, _ in KEEP:
                v = r[col[m]] if m in col else ""
                if m in col and units[col[m]] in ("Mbyte", "Gbyte", "Kbyte") and v:
                    v = "%.0f" % (float(v) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[col[m]]])
                if m == "gpu__time_duration.sum" and m in col and units[col[m]] != "us" and v:
                    v = "%.3f" % (float(v) * {"ns": 1e-3, "ms": 1e3, "s": 1e6}.get(units[col[m]], 1.0))
                vals.append(v)
            st = []
            for s in STALLS:
                m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
                st.append(r[col[m]] if m in col else "")
            w.writerow([name] + vals + st)
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
