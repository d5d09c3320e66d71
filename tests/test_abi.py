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


def test_header_is_valid_c99_and_struct_layouts_match_ctypes(tmp_path):
    """include/virnet_b200.h must be consumable by a plain C compiler (the boundary is a C ABI, not C++), and every
    struct a caller fills must have the size and field offsets the ctypes mirror in virnet_b200/lib.py uses."""
    import shutil
    import subprocess
    from virnet_b200 import lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = ["vk_conv_args", "vk_wgrad_args", "vk_elbo_sisr_args", "vk_sft_desc", "vk_adam_group", "vk_pack_desc",
               "vk_unpack_desc", "vk_extra_src", "vk_sft_apply_args", "vk_sft_apply_bwd_args"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "virnet_b200.h"', 'int main(void) {']
    for name in structs:
        ct = getattr(lib, name)
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in ct._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi_check.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_check"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for name in structs:
        ct = getattr(lib, name)
        assert int(out[name]) == ctypes.sizeof(ct), name
        for field, _ in ct._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(ct, field).offset, f"{name}.{field}"


def declared_prototypes():
    """{name: [parameter declarations]} of every function prototype in the header."""
    text = (ROOT / "include" / "virnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"^\s*(?:const\s+)?(?:int|int64_t|uint32_t|uint64_t|char\s*\*|const char\s*\*)\s+\**(vk_\w+)\s*\(([^;{]*?)\)\s*;",
                         text, flags=re.M | re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params in ([""], ["void"]) else params
    return out


def test_ctypes_signatures_follow_the_header():
    """Every prototype of include/virnet_b200.h must have a ctypes signature in virnet_b200/lib.py with the same number
    of parameters and the same parameter classes (pointer / 32-bit / 64-bit integer / float / double): a prototype that
    changes without its binding would pass garbage to the kernels."""
    from virnet_b200 import lib

    def cls_of_decl(decl):
        if "*" in decl:
            return "ptr"
        ty = decl.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int32_t": "i32", "uint32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "i64", "float": "f32",
                "double": "f64"}[ty]

    def cls_of_ctype(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
            return "ptr"
        return {ctypes.c_int32: "i32", ctypes.c_uint32: "i32", ctypes.c_int: "i32", ctypes.c_int64: "i64",
                ctypes.c_uint64: "i64", ctypes.c_float: "f32", ctypes.c_double: "f64"}[t]

    protos = declared_prototypes()
    assert len(protos) >= 40, len(protos)
    missing = sorted(set(protos) - set(lib._SIGNATURES))
    assert not missing, f"header functions without a ctypes signature: {missing}"
    for name, params in protos.items():
        _, args = lib._SIGNATURES[name]
        assert len(args) == len(params), (name, len(args), len(params))
        got = [cls_of_ctype(t) for t in args]
        want = [cls_of_decl(p) for p in params]
        assert got == want, (name, got, want)
