"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vfmreg_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from vfm_registration_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_all()
    return _lib.LIB_PATH


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"VFMREG_API\s+[\w\s\*]+?\b(vfmreg_\w+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vfmreg_b200.h but not exported"


def test_python_binding_covers_header(lib_path):
    from vfm_registration_b200 import _lib
    assert set(_lib.declared_symbols()) == set(_declared())
    lib = _lib.load()
    assert lib.vfmreg_version() == 200


def test_struct_layouts_match_header(lib_path, tmp_path):
    """sizeof/offsetof of the ABI structs as the C compiler sees them == the ctypes mirrors."""
    from vfm_registration_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vfmreg_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(vfmreg_register_params), offsetof(vfmreg_register_params, inlier_thresh), sizeof(vfmreg_register_result),'
                   'offsetof(vfmreg_register_result, fitness), sizeof(vfmreg_camera), offsetof(vfmreg_camera, subsample));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.RegisterParams), _lib.RegisterParams.inlier_thresh.offset, ctypes.sizeof(_lib.RegisterResult),
            _lib.RegisterResult.fitness.offset, ctypes.sizeof(_lib.Camera), _lib.Camera.subsample.offset]
    assert got == want


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vfm_registration_b200 as v
    with pytest.raises(v.VfmRegError, match="no CPU fallback"):
        v.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vfm_registration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports the oracle"
                assert "oracle/_build" not in txt and "liboracle" not in txt, f"{f} links the oracle"
