"""Print selected metrics of an ncu raw CSV export (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import re
import sys

PAT = re.compile(
    r"gpu__time_duration.sum|sm__throughput.avg.pct|smsp__inst_executed.sum$|smsp__thread_inst_executed.sum$"
    r"|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit"
    r"|smsp__issue_active.avg.pct|lts__t_sectors_op_red.sum|lts__t_sectors_op_atom.sum|lts__t_throughput.avg.pct"
    r"|local_op_(ld|st).sum$|dram__bytes_(read|write).sum$|thread_inst_executed_per_inst_executed"
    r"|pipe_(xu|fma|alu|fp64|lsu)_cycles_active.avg.pct|warp_issue_stalled_.*_per_warp_active.pct"
    r"|launch__grid_size|launch__block_size|sm__cycles_elapsed.max|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
    r"|gpu__dram_throughput.avg.pct|l1tex__throughput.avg.pct|launch__waves|sm__inst_executed_pipe_.*sum$"
    r"|lts__t_sector_hit_rate.pct|smsp__cycles_active.avg|launch__shared_mem|derived__smsp__sass_thread_inst_executed_op_.*_pred_on.sum$")


def main(path, kernel_filter=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        if kernel_filter and kernel_filter not in name:
            continue
        print("==", name[:100])
        for i, h in enumerate(hdr):
            if PAT.search(h) and vals[i] not in ("", "0", "0.00"):
                print("  %-90s %-14s %s" % (h, units[i], vals[i]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
