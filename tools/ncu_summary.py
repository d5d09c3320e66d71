#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` capture (reads the .ncu-rep through `ncu -i ... --page raw --csv`) and the
time-weighted tensor-pipe utilisation over the captured launches.

    python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep > profiles/r02_ncu_conv_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "launch__registers_per_thread",
]


def to_float(v, unit):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * scale.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    print(f"# {rep}: {len(data)} launches; per launch: kernel, grid, and the metrics below (ncu --set full --clock-control none)")
    tw_active = tw_elapsed = t_total = 0.0
    dram = 0.0
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        print("----")
        print(f"  {'Kernel Name':74s} {name}")
        print(f"  {'Grid Size':74s} {r[col['Grid Size']]}")
        t = None
        for m in METRICS:
            if m not in col or r[col[m]] in ("", "n/a"):
                continue
            print(f"  {m:74s} {r[col[m]]} {units[col[m]]}")
        t = to_float(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        a = float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        e = float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])
        tw_active += a * t
        tw_elapsed += e * t
        t_total += t
        dram += to_float(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
            to_float(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    print("====")
    print(f"launches {len(data)}, total {t_total:.1f} us")
    print(f"time-weighted sm__pipe_tensor_cycles_active: {tw_active / t_total:.1f} % of sustained-active, "
          f"{tw_elapsed / t_total:.1f} % of elapsed")
    print(f"mean dram bytes (read + write) per launch: {dram / len(data):.0f}")


if __name__ == "__main__":
    main()
