"""one block per profiled launch of `ncu -i report --page raw --csv` files:  python scripts/ncu_raw_summary.py a.raw.csv [b.raw.csv ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            print(f"== {path}: empty")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {d.get('Kernel Name', '?').split('(')[0].replace('void ', '')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}  [{path.rsplit('/', 1)[-1]}, ncu --set full --clock-control none]")
            for k in KEYS:
                if k in d:
                    print(f"   {k:84s} {d[k]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    main()
