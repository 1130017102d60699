"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
validates arguments, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from qcdmet_b200 import _lib, workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qcdmet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    return sorted({n for n in names if n.startswith("qcdmet_") or n == "rhf_response"})


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert "rhf_response" in syms and "qcdmet_jacobian" in syms and len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/qcdmet_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "python binding table and header disagree"
    assert b"sm_100a" in lib.qcdmet_version()


def test_argument_validation():
    lib = _lib.load()
    w = wl.sweep(8)
    handle = ctypes.c_void_p()
    hs = w.h1start.copy()
    bad_row = w.h1row.copy()
    bad_row[0] = 99
    rc = lib.qcdmet_plan_create(ctypes.byref(handle), -1, 8, w.nterms, 4, _lib._ptr(hs), _lib._ptr(bad_row),
                                _lib._ptr(w.h1col), len(w.row_start) - 1, _lib._ptr(w.row_start), None,
                                _lib._ptr(w.row_idx))
    assert rc == -1 and b"out of range" in lib.qcdmet_last_error()
    rc = lib.qcdmet_plan_create(ctypes.byref(handle), -1, 8, w.nterms, 8, _lib._ptr(hs), _lib._ptr(w.h1row),
                                _lib._ptr(w.h1col), len(w.row_start) - 1, _lib._ptr(w.row_start), None,
                                _lib._ptr(w.row_idx))
    assert rc == -1 and b"numPairs" in lib.qcdmet_last_error()
    assert lib.qcdmet_plan_compute(None) == -1
    assert lib.qcdmet_plan_destroy(None) == 0


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.qcdmet_device_count() > 0:
        pytest.skip("a CUDA device is present")
    w = wl.sweep(8)
    with pytest.raises(_lib.QcdmetError) as ei:
        _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    h0 = np.array(w.h0, copy=True).reshape(-1)
    out = np.ones(w.norb * w.norb * w.nterms)
    rc = lib.rhf_response(w.norb, w.nterms, w.num_pairs, _lib._ptr(w.h1start), _lib._ptr(w.h1row),
                          _lib._ptr(w.h1col), _lib._ptr(h0), _lib._ptr(out))
    assert rc == -2
    assert (out == 1.0).all() and np.array_equal(h0, w.h0.reshape(-1))   # nothing written


def test_missing_library_fails_loudly():
    """No silent fallback: a missing libqcdmet_b200.so is an OSError at load time."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['QCDMET_B200_LIB'] = '/nonexistent/libqcdmet_b200.so'; "
            "from qcdmet_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept OSError as e:\n    print('OSError', 'no CPU fallback' in str(e))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "OSError True", out.stderr[-500:]


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "jacobian_c_abi")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "jacobian_c_abi.c"), "-o", exe, "-L", os.path.join(ROOT, "qcdmet_b200"),
           "-lqcdmet_b200", "-Wl,-rpath," + os.path.join(ROOT, "qcdmet_b200"), "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    """include/qcdmet_b200.h compiles as C (no C++/torch types) and the library links from a C program;
    without a device the program reports the library's error instead of computing anything."""
    import subprocess
    exe = _build_c_example(tmp_path)
    out = subprocess.run([exe, "16"], capture_output=True, text=True, timeout=120)
    if _lib.load().qcdmet_device_count() == 0:
        assert out.returncode == 1 and "no CPU fallback" in out.stderr
    else:
        assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_program_computes_jacobian(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    for n in ("16", "200"):
        out = subprocess.run([exe, n], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "block asymmetry 0.0e+00" in out.stdout
