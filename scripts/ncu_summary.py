"""Summarise `ncu --set full` reports for profiles/ (no GPU needed).
    python scripts/ncu_summary.py "<title>" <report.ncu-rep> ["<title>" <report> ...]
One block per report: the metrics DESIGN.md / bench.py's roofline refer to (time, DRAM traffic, tensor / FP64 pipes,
issue utilisation, occupancy, shared-memory conflicts, L2 hit rate)."""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main():
    args = sys.argv[1:]
    for title, rep in zip(args[0::2], args[1::2]):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("== %s  [%s, ncu --set full --clock-control none]" % (title, rep.rsplit("/", 1)[-1]))
            print("   kernel: %s" % d.get("Kernel Name", "?").split("(")[0])
            for k in KEYS:
                if k in d:
                    print("   %-82s %s %s" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
    main()
