"""The C-ABI library must load on a CPU-only box and export every symbol the header declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    text = (ROOT / "include" / "virnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|int64_t|uint32_t|uint64_t|char\s*\*|const char\s*\*)\s+\**(vk_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("vk_conv_igemm", "vk_conv_wgrad", "vk_elbo_denoise", "vk_adam_clip_step", "vk_pack_input"):
        assert must in names


def test_library_loads_and_exports_all_symbols():
    from virnet_b200 import lib
    if not lib.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    handle = lib.load()
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in include/virnet_b200.h but not exported"
    assert set(lib.exported_symbols()) == set(declared_functions())
    assert b"sm_100a" in handle.vk_version()


def test_struct_layouts_match():
    from virnet_b200 import lib
    h = lib.load()
    assert h.vk_sizeof_conv_args() == ctypes.sizeof(lib.vk_conv_args)
    assert h.vk_sizeof_wgrad_args() == ctypes.sizeof(lib.vk_wgrad_args)
    assert ctypes.sizeof(lib.vk_pack_desc) == 48
    assert ctypes.sizeof(lib.vk_adam_group) == 24


def test_bad_arguments_are_rejected_without_a_gpu():
    from virnet_b200 import lib
    h = lib.load()
    assert h.vk_conv_igemm(None, None) == -1
    a = lib.vk_conv_args()
    assert h.vk_conv_igemm(ctypes.byref(a), None) == -1
    g = lib.vk_wgrad_args()
    assert h.vk_conv_wgrad(ctypes.byref(g), None) == -1
