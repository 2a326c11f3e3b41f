"""The C-ABI library loads without a GPU and exports every symbol include/feellgood_b200.h
declares; without a device the constructors fail loudly (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import cases
from feellgood_b200 import capi

HEADER = os.path.join(cases.ROOT, "include", "feellgood_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(gpu_lib):
    names = _declared()
    assert len(names) >= 35
    missing = [s for s in names if not hasattr(gpu_lib, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == names          # the ctypes binding tracks the header


def test_exports_are_plain_c(gpu_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(_declared()) <= exported
    # nothing from the oracle is linked into the product
    assert not [s for s in exported if s.startswith("fgo_") or s.startswith("fgref_")]
    deps = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "fg_oracle" not in deps and "fgref" not in deps


def test_no_device_fails_loudly(gpu_lib):
    import torch
    if torch.cuda.is_available():
        return                                    # covered by the gpu tests on a B200 box
    case = cases.small_cuboid()
    from feellgood_b200 import LinAlgebra, Settings
    s = Settings([capi.tet_prm(**r) for r in case.tet_regions],
                 [capi.tri_prm(**r) for r in case.tri_regions])
    try:
        LinAlgebra(s, case.mesh)
    except capi.FgError as e:
        assert e.code == -2 and "no CUDA device" in str(e)
    else:
        raise AssertionError("LinAlgebra construction must fail without a CUDA device")
    assert gpu_lib.fg_version() >= 100


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors = sizeof of the C structs (compiled with gcc)."""
    code = r'''
#include <stdio.h>
#include "feellgood_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(fg_tet_prm), sizeof(fg_tri_prm),
 sizeof(fg_mesh), sizeof(fg_params), sizeof(fg_step_result), sizeof(fg_iter_result));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.dirname(HEADER), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe], text=True).split()]
    assert sizes == [C.sizeof(capi.TetPrm), C.sizeof(capi.TriPrm), C.sizeof(capi.CMesh),
                     C.sizeof(capi.CParams), C.sizeof(capi.StepResult), C.sizeof(capi.IterResult)]


def test_every_entry_point_is_documented():
    """INTEGRATION.md maps every exported entry point to the reference call it replaces."""
    doc = open(os.path.join(cases.ROOT, "INTEGRATION.md")).read()
    missing = [s_ for s_ in _declared() if s_ not in doc]
    assert not missing, missing
