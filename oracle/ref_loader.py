"""ctypes loader for the UNMODIFIED reference kernel compiled into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/response_oracle.py header).  Mirrors how the
reference binds its library (src/qcdmethelper.py:23-24,102-109): ints by value as c_int,
buffers as void pointers from numpy.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_LIBS = {}


def ref_path(flavour: str = "mkl") -> str:
    return os.path.join(_REF_DIR, f"libqcdmet_ref_{flavour}.so")


def available(flavour: str = "mkl") -> bool:
    return os.path.exists(ref_path(flavour))


def load(flavour: str = "mkl"):
    """flavour: 'mkl' (MKL embedded in libtorch_cpu; needs `import torch` first) or 'openblas'."""
    if flavour in _LIBS:
        return _LIBS[flavour]
    if flavour == "mkl":
        import torch  # noqa: F401  (loads libtorch_cpu.so and its dependencies: the MKL symbols)
    else:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    lib = ctypes.CDLL(ref_path(flavour))
    lib.rhf_response.restype = None
    _LIBS[flavour] = lib
    return lib


def rhf_response(norb, num_pairs, h1start, h1row, h1col, h0, flavour: str = "mkl"):
    """Call the compiled reference.  Returns (resp[P,N,N], one_rdm[N,N]).

    Guards the reference's own limits: 32-bit output index (lib/rhf_response.cc:106).
    """
    lib = load(flavour)
    nterms = len(h1start) - 1
    if norb * norb * max(nterms, 1) >= 2 ** 31:
        raise ValueError("reference kernel overflows its int index at N^2*P >= 2^31")
    h1start = np.ascontiguousarray(h1start, dtype=np.int32)
    h1row = np.ascontiguousarray(h1row, dtype=np.int32)
    h1col = np.ascontiguousarray(h1col, dtype=np.int32)
    if h1row.size == 0:  # keep a valid pointer for Nterms == 0
        h1row = np.zeros(1, dtype=np.int32)
        h1col = np.zeros(1, dtype=np.int32)
    buf = np.array(np.asarray(h0, dtype=np.float64).reshape(norb * norb), dtype=np.float64, copy=True)
    out = np.ones(max(norb * norb * nterms, 1), dtype=np.float64)
    lib.rhf_response(ctypes.c_int(norb), ctypes.c_int(nterms), ctypes.c_int(num_pairs),
                     h1start.ctypes.data_as(ctypes.c_void_p),
                     h1row.ctypes.data_as(ctypes.c_void_p),
                     h1col.ctypes.data_as(ctypes.c_void_p),
                     buf.ctypes.data_as(ctypes.c_void_p),
                     out.ctypes.data_as(ctypes.c_void_p))
    resp = out[:norb * norb * nterms].reshape((nterms, norb, norb), order="C")
    return resp, buf.reshape(norb, norb)
