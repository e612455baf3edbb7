"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/blance_b200.h declares; without a CUDA device the compute entry points
fail loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import blance_b200
from blance_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "blance_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(blance_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build.lib_path())
    names = declared_functions()
    assert set(names) == set(api.EXPORTS), (names, api.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layout_matches_header():
    # sizes as laid out by the C compiler for include/blance_b200.h (checked by compiling a probe)
    import subprocess
    import tempfile
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "blance_b200.h"
    int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(blance_plan_in), offsetof(blance_plan_in, state_priority),
                            offsetof(blance_plan_in, n_rules), offsetof(blance_plan_in, engine), sizeof(blance_plan_out)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(probe)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "p")], check=True)
        out = subprocess.run([os.path.join(d, "p")], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    size_in, off_sp, off_nr, off_eng, size_out = map(int, out)
    assert ctypes.sizeof(api.PlanIn) == size_in
    assert api.PlanIn.state_priority.offset == off_sp
    assert api.PlanIn.n_rules.offset == off_nr
    assert api.PlanIn.engine.offset == off_eng
    assert ctypes.sizeof(api.PlanOut) == size_out


def _have_gpu():
    lib = api.capi()
    ctx = ctypes.c_void_p()
    st = lib.blance_ctx_create(ctypes.byref(ctx), -1)
    if st == 0:
        lib.blance_ctx_destroy(ctx)
    return st == 0


def test_no_cpu_fallback_without_a_device():
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    lib = api.capi()
    ctx = ctypes.c_void_p()
    assert lib.blance_ctx_create(ctypes.byref(ctx), -1) == -3          # BLANCE_ERR_CUDA
    assert b"no CPU fallback" in lib.blance_last_error(None)
    with pytest.raises(blance_b200.BlanceError):
        blance_b200.PlanNextMapEx({}, {"0": {}}, ["a"], [], ["a"], {"primary": (0, 1)})
    with pytest.raises(blance_b200.BlanceError):
        blance_b200.CalcPartitionMoves(["primary"], {"primary": ["a"]}, {"primary": ["b"]}, False)
