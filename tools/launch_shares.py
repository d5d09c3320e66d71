#!/usr/bin/env python
"""Kernel shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): time per kernel family and per
conv / wgrad instantiation.   python tools/launch_shares.py gpurun_out/launches.csv STEPS > profiles/r02_launch_shares.txt"""
import collections
import csv
import re
import sys


def main():
    path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    fam, inst = collections.OrderedDict(), collections.OrderedDict()
    tot = 0.0
    for d in rows:
        k, t = d["Kernel Name"], float(d["Metric Value"]) / 1e3
        tot += t
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", k))
        short = re.sub(r"<.*", "", name)
        a = fam.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += t
        if "conv_v2" in k or "conv_igemm" in k or "wgrad_kernel" in k:
            m = re.search(r"<(.*)>", name)
            key = (short.split("::")[-1], m.group(1) if m else "", d["Grid Size"])
            b = inst.setdefault(key, [0, 0.0])
            b[0] += 1
            b[1] += t
    print(f"# {path}: {len(rows)} launches over {steps} training steps (b=32, bf16), total {tot:.1f} us "
          f"= {tot / steps:.1f} us per step (serialised under ncu, cold caches: compare SHARES, not absolute times)")
    for k, a in sorted(fam.items(), key=lambda x: -x[1][1]):
        print(f"{k[-48:]:48s} n/step={a[0] / steps:6.1f} {a[1] / steps:9.1f} us/step {100 * a[1] / tot:5.1f}%")
    print("\n# conv / wgrad instantiations: (kernel, template arguments <dtype, chunk bytes, taps per item, pair, full-K, kEpi>, grid)")
    for k, a in sorted(inst.items(), key=lambda x: -x[1][1]):
        print(f"{str(k):96s} n/step={a[0] / steps:5.1f} {a[1] / steps:8.1f} us/step {a[1] / a[0]:7.1f} us each")


if __name__ == "__main__":
    main()
