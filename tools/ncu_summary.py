"""Summarise one kernel of an .ncu-rep (ncu --set full) into the text committed under profiles/.

usage: python tools/ncu_summary.py report.ncu-rep "command line that was profiled" > profiles/rNN_xxx.txt
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print("ncu --set full --clock-control none, kernel %s" % name)
        print("command: %s" % cmd)
        print()
        col = {h: i for i, h in enumerate(hdr)}
        for k in KEYS:
            if k in col:
                print("%-92s %-16s %s" % (k, units[col[k]], vals[col[k]]))
        for h in hdr:
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                v = vals[col[h]]
                try:
                    if float(v) >= 0.05:
                        print("%-92s %-16s %s" % (h, units[col[h]], v))
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main()
