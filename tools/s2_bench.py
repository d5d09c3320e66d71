#!/usr/bin/env python
"""Strided / transposed kinds and the C=64 layers: v1 vs v2 single vs v2 pair (tuning aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from virnet_b200 import ops  # noqa: E402
from tools.conv_bench import run  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
run(96, 192, 128, n, "dual", kind=ops.VK_CONV3X3_S2)
run(192, 288, 64, n, "dual", kind=ops.VK_CONV3X3_S2)
run(96, 192, 128, n, "dual", kind=ops.VK_CONV3X3_S2, tune=dict(p=1), impls=(3, 4))
run(192, 288, 64, n, "dual", kind=ops.VK_CONV3X3_S2, tune=dict(p=1), impls=(3, 4))
run(64, 64, 128, n, "out2")
run(64, 64, 128, n, "mask")
