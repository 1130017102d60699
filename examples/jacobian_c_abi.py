#!/usr/bin/env python
"""The fused Jacobian through the bare C ABI (ctypes), the way INTEGRATION.md binds it from QC-DMET:
one call replaces rhf_response + the Python assembly loop.  Needs a CUDA device.

    python examples/jacobian_c_abi.py [norb]
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qcdmet_b200 import workloads   # noqa: E402  (synthetic inputs only)

norb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = workloads.sweep(norb)
lib = ctypes.CDLL(os.environ.get("QCDMET_B200_LIB", os.path.join(ROOT, "qcdmet_b200", "libqcdmet_b200.so")))
lib.qcdmet_jacobian.restype = ctypes.c_int
lib.qcdmet_last_error.restype = ctypes.c_char_p
vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
Q, P = w.nresid, w.nterms
J = np.empty((Q, P), dtype=np.float64, order='F')
eig = np.empty(norb)
H0 = np.ascontiguousarray(w.h0)
t0 = time.time()
rc = lib.qcdmet_jacobian(ctypes.c_int(norb), ctypes.c_int(P), ctypes.c_int(w.num_pairs), vp(w.h1start), vp(w.h1row),
                         vp(w.h1col), vp(H0), ctypes.c_int(len(w.row_start) - 1), vp(w.row_start), None,
                         vp(w.row_idx), ctypes.c_int(0), ctypes.c_int(P), vp(J), None, vp(eig))
assert rc == 0, lib.qcdmet_last_error().decode()
print(f"Norb={norb}: (Q, P) = {J.shape} Jacobian in {time.time() - t0:.3f} s (plan creation + H2D + syevd + kernels + D2H)")
print(f"HOMO-LUMO gap = {eig[w.num_pairs] - eig[w.num_pairs - 1]:.4f},  |J|_max = {np.abs(J).max():.4f}")
