"""Import the UNMODIFIED reference (/root/reference) in the build container.

The reference imports a few packages that are absent here at module import time
(`thop` via utils/util_net.py:7; `lpips`, `matplotlib.pyplot`, `skimage` via
utils/util_image.py:9-14).  None of them is used on the hot path, so they are replaced
by empty stub modules.  This helper is only used by tools/gen_golden.py and by tests that
are skipped when /root/reference is absent (it never exists on the GPU box).
"""
import sys
import types
from pathlib import Path

_CANDIDATES = (Path("/root/reference"), Path(__file__).resolve().parents[1] / "baseline" / "_ref")


def _find_root() -> Path:
    for c in _CANDIDATES:
        if (c / "networks" / "VIRNet.py").exists():
            return c
    return _CANDIDATES[0]


# /root/reference in the build container; the staged copy baseline/_ref (baseline/install_ref.py, git-ignored)
# on the GPU box, where it is only ever used as the timed comparator of bench.py
REF_ROOT = _find_root()


def available() -> bool:
    return (REF_ROOT / "networks" / "VIRNet.py").exists()


def _stub(name: str, **attrs):
    if name in sys.modules:
        return
    try:
        __import__(name)
        return
    except Exception:  # noqa: BLE001
        pass
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def import_reference():
    """Returns (networks.VIRNet module, loss.ELBO_simple module)."""
    if not available():
        raise RuntimeError("the reference is not present (/root/reference or baseline/_ref)")
    _stub("thop", profile=lambda *a, **k: (0, 0))
    _stub("lpips")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    import numpy as np

    def _as_ubyte(x):                      # skimage.img_as_ubyte for float images in [0, 1]
        x = np.asarray(x)
        return x if x.dtype == np.uint8 else np.clip(np.round(x * 255.0), 0, 255).astype(np.uint8)

    _stub("skimage", img_as_ubyte=_as_ubyte,
          img_as_float32=lambda x: np.asarray(x, dtype=np.float32) / (255.0 if np.asarray(x).dtype == np.uint8 else 1.0),
          img_as_float64=lambda x: np.asarray(x, dtype=np.float64) / (255.0 if np.asarray(x).dtype == np.uint8 else 1.0))
    _stub("skimage.metrics")
    _stub("skimage.color")
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import importlib
    vir = importlib.import_module("networks.VIRNet")
    elbo = importlib.import_module("loss.ELBO_simple")
    return vir, elbo
