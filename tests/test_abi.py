"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fg.h declares, struct layouts match, and the no-device behaviour is an error (there is
no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from film_grain_b200 import build
    build.build()
    from film_grain_b200 import _lib
    return _lib.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "fg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from film_grain_b200 import _lib
    declared = _header_functions()
    assert len(declared) >= 15
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (fg_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_lib.ABI.keys())


def test_struct_layout_matches_header(lib, tmp_path):
    from film_grain_b200 import FgParams, FgStats
    src = tmp_path / "sz.c"
    src.write_text('#include "fg.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(fg_params),sizeof(fg_stats),offsetof(fg_params,seed),offsetof(fg_params,radius_log_mu),'
                   'offsetof(fg_params,row_begin));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals == [C.sizeof(FgParams), C.sizeof(FgStats), FgParams.seed.offset, FgParams.radius_log_mu.offset,
                    FgParams.row_begin.offset]


def test_no_device_is_an_error_not_a_fallback(lib):
    import film_grain_b200 as fg
    if fg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fg.GpuError) as e:
        fg.Context(0)
    assert e.value.code == -4
    assert lib.fg_error_string(-4) == b"no usable CUDA device"


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "film_grain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("the oracle", "").replace("cpu oracle", "") or f == "fg_zig_tables.h", (dirpath, f)
