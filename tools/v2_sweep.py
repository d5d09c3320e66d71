#!/usr/bin/env python
"""Sweep the v2 tuning knobs (tiles per job, taps per weight stage, K chunk) at the main layer shapes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tools.conv_bench import run  # noqa: E402

n = 16
for c, h in ((96, 128), (192, 64), (288, 32), (64, 128)):
    for epi in ("out2", "resid_dual"):
        for impl in (3, 4):
            for p in (1, 2, 4):
                for nt in (3, 9):
                    if c == 288 and impl == 4:
                        continue
                    run(c, c, h, n, epi, tune=dict(p=p, nt=nt), impls=(impl,), check=False)
