#!/usr/bin/env python
"""Time-weighted tensor-pipe utilisation per kernel class from an ncu metrics CSV of one training step:

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'conv_v2|wgrad_kernel' \\
        -c 130 --csv --log-file gpurun_out/tensor_active_step.csv python tools/profile_step.py --steps 1 --batch 32
    python tools/tensor_active.py gpurun_out/tensor_active_step.csv > profiles/r02_tensor_active_step.txt
"""
import collections
import csv
import re
import sys

ACT = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
ELA = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    L = collections.OrderedDict()
    for d in csv.DictReader(lines):
        L.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"]})[d["Metric Name"]] = (
            float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
    us = lambda v, u: v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    print("# every conv_v2_kernel / wgrad_kernel launch of ONE b=32 bf16 training step; classes: s1 = 3x3 stride 1 (slab mode),")
    print("# fullK = stride-2 / ConvT / their dgrads, c16 = 16-channel-input layers (chunk 32), generic = kEpi -1 (NCHW outputs, 16-row dgrad)")
    agg, inst = collections.OrderedDict(), collections.OrderedDict()
    for d in L.values():
        t = us(*d["gpu__time_duration.sum"])
        a, e = d[ACT][0], d[ELA][0]
        name = d["name"]
        m = re.search(r"<(.*)>", name)
        args = m.group(1).split(", ") if m else []
        if "wgrad" in name:
            cls = "wgrad grid " + d["grid"]
        elif len(args) >= 6:
            if "-1" in args[5]:
                cls = "conv generic"
            elif args[4] == "1":
                cls = "conv fullK"
            elif args[1] == "32":
                cls = "conv c16"
            else:
                cls = f"conv s1 chunk{args[1]} nt{args[2]}"
        else:
            cls = "conv generic"
        for table, key in ((agg, cls), (inst, (m.group(1) if m else name[-40:]) if "wgrad" not in name else cls)):
            g = table.setdefault(key, [0, 0.0, 0.0, 0.0])
            g[0] += 1
            g[1] += t
            g[2] += a * t
            g[3] += e * t
    for cls, g in agg.items():
        print(f"{cls:30s} n={g[0]:3d} {g[1]:8.1f} us  tensor active {g[2] / g[1]:5.1f} % (of elapsed {g[3] / g[1]:5.1f} %)")

    def tot(pred):
        s = [g for c, g in agg.items() if pred(c)]
        T = sum(g[1] for g in s)
        return sum(g[0] for g in s), T, sum(g[2] for g in s) / T, sum(g[3] for g in s) / T
    print()
    for label, pred in (("all 3x3 stride-1 conv launches (slab mode, >= 64 channels)", lambda c: c.startswith("conv s1")),
                        ("all conv_v2_kernel launches", lambda c: c.startswith("conv")),
                        ("all wgrad_kernel launches", lambda c: c.startswith("wgrad"))):
        n, T, a, e = tot(pred)
        print(f"{label:60s} n={n:3d} {T:8.1f} us  time-weighted tensor active {a:5.1f} % (of elapsed {e:5.1f} %)")
    print("\n# per instantiation <dtype, chunk bytes, taps per item, pair, full-K, kEpi (1 mask | 2 residual | 4 out1 | 8 out2)>")
    for k, g in sorted(inst.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} n={g[0]:2d} {g[1]:8.1f} us {g[1] / g[0]:7.1f} each  tensor active {g[2] / g[1]:5.1f} %")


if __name__ == "__main__":
    main()
