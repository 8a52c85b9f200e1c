#!/usr/bin/env python
"""Per-kernel table from an ncu capture: `ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv`.
One row per captured launch: duration, registers, grid, achieved occupancy, active lanes per warp instruction, IPC, issue-slot
utilisation, DRAM bytes / rate, L2 rate, hit rates.  The capture command is in the header of profiles/*_all_kernels_table.txt."""
import csv
import re
import sys

HBM_PEAK = 6444.4        # MEASURED_PEAKS.json hbm_gbs on this pool's B200s


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {k: h.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
                                   "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
                                   "sm__inst_executed.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
                                   "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
                                   "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed")}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    to_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print(f"{'kernel':48s} {'us':>8s} {'regs':>4s} {'grid':>7s} {'occ%':>5s} {'lanes':>5s} {'IPC':>5s} {'issue%':>6s} {'DRAM MB':>8s} {'DRAM GB/s':>9s} "
          f"{'DRAM%':>6s} {'L2 GB/s':>8s} {'L2%':>5s} {'L1hit':>5s} {'L2hit':>5s}")
    for r in rows[2:]:
        g = lambda k: num(r[col[k]])
        us = g("gpu__time_duration.sum") * to_us[units[col["gpu__time_duration.sum"]]]
        dram = (g("dram__bytes_read.sum") * to_b[units[col["dram__bytes_read.sum"]]] +
                g("dram__bytes_write.sum") * to_b[units[col["dram__bytes_write.sum"]]])
        l2 = g("lts__t_sectors.sum") * 32.0
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
        print(f"{name[-48:]:48s} {us:8.1f} {int(g('launch__registers_per_thread')):4d} {int(g('launch__grid_size')):7d} "
              f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} "
              f"{g('sm__inst_executed.avg.per_cycle_active'):5.2f} {g('sm__inst_issued.avg.pct_of_peak_sustained_active'):6.1f} {dram / 1e6:8.1f} "
              f"{dram / us / 1e3:9.1f} {100.0 * dram / us / 1e3 / HBM_PEAK:6.1f} {l2 / us / 1e3:8.1f} {g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
              f"{g('l1tex__t_sector_hit_rate.pct'):5.1f} {g('lts__t_sector_hit_rate.pct'):5.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
