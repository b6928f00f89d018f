#!/usr/bin/env python3
"""One line per capture of tools/ncu_capture_r02.sh (gpurun_out/r02_ncu_*_raw.csv): duration, DRAM bytes and rate, the busiest
pipe, occupancy limits.  usage: ncu_summary_r02.py DIR"""
import csv
import glob
import os
import sys

KEYS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rd MB", 1e-6), ("dram__bytes_write.sum", "wr MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1)]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(d):
    for path in sorted(glob.glob(os.path.join(d, "r02_ncu_*_raw.csv"))):
        rows = list(csv.reader(open(path)))
        try:
            hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        except StopIteration:
            print(os.path.basename(path), ": no kernel captured")
            continue
        names, units = rows[hdr], rows[hdr + 1]
        for r in rows[hdr + 2:]:
            if len(r) < len(names):
                continue
            m = dict(zip(names, r))
            u = dict(zip(names, units))
            out = [os.path.basename(path)[8:-8], m["Kernel Name"].split("(")[0][:28]]
            t_us = rd = wr = None
            for k, label, scale in KEYS:
                v = num(m.get(k, ""))
                if v is None:
                    continue
                unit = u.get(k, "")
                if k == "gpu__time_duration.sum":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
                    t_us = v
                elif k.startswith("dram__bytes"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
                    if "read" in k:
                        rd = v
                    else:
                        wr = v
                out.append(f"{label} {v:.1f}" if abs(v) < 1e6 else f"{label} {v:.3g}")
            if t_us and rd is not None and wr is not None:
                out.append(f"DRAM GB/s {(rd + wr) / t_us * 1e3 / 1e3:.0f}")
            print(" | ".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
