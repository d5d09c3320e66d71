#!/usr/bin/env python
"""Role-stall breakdown of the persistent conv kernel on the strided kinds (full-K mode) and the C=64 layers."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from virnet_b200 import ops  # noqa: E402
from tools.v2_timing import run  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
run(96, 192, 128, n, "dual", kind=ops.VK_CONV3X3_S2, impl=2)
run(192, 288, 64, n, "dual", kind=ops.VK_CONV3X3_S2, impl=2)
run(64, 64, 128, n, "out2", impl=2)
run(64, 64, 128, n, "mask", impl=2)
run(96, 96, 128, n, "out2", impl=2)
