"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The reference is plain Python with no packaging for the hot path (`pip install /root/reference` only
finds the vendored gradual_warmup_lr scheduler), so the "install" is a file copy of the packages the hot
path imports.  Nothing under baseline/_ref/ is product code or test-required: it is the comparator that
`bench.py --impl reference` (CPU) and bench.py's `gpu_comparator` leg (eager PyTorch / cuDNN on the same
B200) time, as BASELINE.md §5.1/§5.6 prescribe.  Everything falls back to the oracle port when it is absent.

    python baseline/install_ref.py        # needs /root/reference (build container only)
"""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
PARTS = ["networks", "loss", "utils", "ResizeRight", "datasets", "configs", "scripts", "gradual_warmup_lr",
         "dnd_submission_py", "train_denoising_syn.py", "train_denoising_real.py", "train_SISR.py", "LICENSE"]


def install(verbose: bool = True) -> bool:
    if not (SRC / "networks" / "VIRNet.py").exists():
        if verbose:
            print("baseline/install_ref.py: /root/reference absent; keeping", DST if DST.exists() else "nothing")
        return DST.exists()
    DST.mkdir(parents=True, exist_ok=True)
    for p in PARTS:
        s, d = SRC / p, DST / p
        if not s.exists():
            continue
        if s.is_dir():
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    if verbose:
        n = sum(1 for _ in DST.rglob("*.py"))
        print(f"baseline/install_ref.py: staged {n} reference .py files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
