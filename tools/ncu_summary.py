"""Text summary of an `ncu --page raw --csv` export (the metrics DESIGN.md quotes), for profiles/."""
import csv
import sys

WANT = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    names = WANT + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio"))
    for w in names:
        if w not in hdr:
            continue
        x = vals[hdr.index(w)]
        try:
            if "stalled" in w and float(x) < 0.05:
                continue
        except ValueError:
            pass
        print(f"{w:100s} {units[hdr.index(w)]:18s} {x}")


if __name__ == "__main__":
    main(sys.argv[1])
