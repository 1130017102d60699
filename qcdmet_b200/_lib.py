"""ctypes binding of libqcdmet_b200.so (include/qcdmet_b200.h).

Same mechanism the reference uses for libqcdmet (src/qcdmethelper.py:23-24: ctypes.CDLL, which
releases the GIL for the duration of every call).  The library path is $QCDMET_B200_LIB, else
the in-tree build next to this file.  There is NO fallback: a missing library or a missing
CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCDMET_B200_LIB", os.path.join(_HERE, "libqcdmet_b200.so"))

KERNEL_AUTO, KERNEL_GENERAL, KERNEL_DMMA, KERNEL_DMMA_NOSYM, KERNEL_FACTORISED = 0, 1, 2, 3, 4
LM_NORMAL, LM_QR = 0, 1

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); mirrors include/qcdmet_b200.h one to one
SIGNATURES = {
    "rhf_response": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_jacobian": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_jacobian_from_orbitals": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                     ctypes.c_void_p]),
    "qcdmet_shutdown": (ctypes.c_int, []),
    "qcdmet_plan_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "qcdmet_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_set_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "qcdmet_plan_set_kernel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "qcdmet_plan_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_set_output_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]),
    "qcdmet_plan_export_jacobian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_attach_peers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "qcdmet_plan_attach_peer_pointers": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "qcdmet_plan_detach_peers": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_set_h0": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_set_orbitals": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_compute": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_jacobian_device": (ctypes.c_void_p, [ctypes.c_void_p]),
    "qcdmet_plan_get_jacobian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_get_dblocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_get_eigenvalues": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_get_orbitals": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_get_full_response": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_get_one_rdm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_construct_baths": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p]),
    "qcdmet_plan_lm_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "qcdmet_plan_embed_hubbard": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_set_lm_method": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "qcdmet_plan_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_num_residuals": (ctypes.c_longlong, [ctypes.c_void_p]),
    "qcdmet_plan_kernel_used": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_symmetric_used": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_split_used": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_flops": (ctypes.c_double, [ctypes.c_void_p]),
    "qcdmet_plan_kernel_time_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_schedule_info": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_auto_kernel": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_contract_schedule": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_set_devices": (ctypes.c_int, [ctypes.c_int]),
    "qcdmet_get_devices": (ctypes.c_int, []),
    "qcdmet_plan_set_devices": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "qcdmet_plan_num_devices": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "qcdmet_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_orbitals_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                                   ctypes.POINTER(ctypes.c_void_p)]),
    "qcdmet_plan_commit_orbitals": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_plan_get_jacobian_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "qcdmet_plan_timings": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qcdmet_plan_jacobian_checksum": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_int)]),
    "qcdmet_plan_info": (ctypes.c_char_p, [ctypes.c_void_p]),
    "qcdmet_plan_flush_l2": (ctypes.c_int, [ctypes.c_void_p]),
    "qcdmet_measure_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]),
    "qcdmet_last_error": (ctypes.c_char_p, []),
    "qcdmet_version": (ctypes.c_char_p, []),
    "qcdmet_device_count": (ctypes.c_int, []),
}

_lib = None


class QcdmetError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libqcdmet_b200 error {code}: {message}")
        self.code = code


def load():
    """Load the shared library (once).  Raises OSError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                      f"g.build()'` or `make -C qcdmet_b200/csrc` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise QcdmetError(rc, load().qcdmet_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def as_i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def schedule_info(norb, num_pairs, h1start, h1row, h1col, row_start, row_idx=None):
    """Host-only: how the DMMA kernel would schedule these lists (dict); no device needed."""
    lib = load()
    hs, hr, hc, rs, ri = as_i32(h1start), as_i32(h1row), as_i32(h1col), as_i32(row_start), as_i32(row_idx)
    if hr.size == 0:
        hr = hc = np.zeros(1, dtype=np.int32)
    info = np.zeros(10, dtype=np.int32)
    check(lib.qcdmet_schedule_info(int(norb), len(hs) - 1, int(num_pairs), _ptr(hs), _ptr(hr), _ptr(hc),
                                   len(rs) - 1, _ptr(rs), _ptr(ri), _ptr(info)))
    keys = ("eligible", "rows_per_fragment", "nj", "components", "row_tiles", "symmetric", "split", "internal_columns",
            "block_pairs", "column_tiles")
    out = dict(zip(keys, (int(v) for v in info)))
    out["why"] = lib.qcdmet_last_error().decode() if not out["eligible"] else ""
    return out


def auto_kernel(norb, num_pairs, h1start, h1row, h1col, row_start, row_idx=None):
    """Host-only: (kernel id AUTO would select, bytes of site panels + tiled operands for the factorised path)."""
    lib = load()
    hs, hr, hc, rs, ri = as_i32(h1start), as_i32(h1row), as_i32(h1col), as_i32(row_start), as_i32(row_idx)
    if hr.size == 0:
        hr = hc = np.zeros(1, dtype=np.int32)
    kernel = ctypes.c_int(-1)
    nbytes = ctypes.c_longlong(0)
    check(lib.qcdmet_auto_kernel(int(norb), len(hs) - 1, int(num_pairs), _ptr(hs), _ptr(hr), _ptr(hc), len(rs) - 1,
                                 _ptr(rs), _ptr(ri), ctypes.byref(kernel), ctypes.byref(nbytes)))
    return kernel.value, nbytes.value


def contract_schedule(row_start):
    """Host-only: tiles of the factorised path's contraction kernel ->
    list of (slot_rows, strips[8 warps][<= 3] of (b1, b2, n2))."""
    lib = load()
    rs = as_i32(row_start)
    counts = np.zeros(3, dtype=np.int32)
    check(lib.qcdmet_contract_schedule(len(rs) - 1, _ptr(rs), _ptr(counts), None, None, None))
    tile_info = np.zeros(4 * max(int(counts[0]), 1), dtype=np.int32)
    slot_row = np.zeros(max(int(counts[1]), 1), dtype=np.int32)
    strips = np.zeros(max(int(counts[2]), 1), dtype=np.int32)
    check(lib.qcdmet_contract_schedule(len(rs) - 1, _ptr(rs), _ptr(counts), _ptr(tile_info), _ptr(slot_row), _ptr(strips)))
    assert int(counts[2]) == 24 * int(counts[0])
    tiles = []
    for t in range(int(counts[0])):
        sb, ns, tb, nstr = (int(x) for x in tile_info[4 * t:4 * t + 4])
        per_warp = []
        for w in range(8):
            ent = [int(x) for x in strips[tb + 3 * w:tb + 3 * w + 3] if x >= 0]
            per_warp.append([(x & 0xFF, (x >> 8) & 0xFF, x >> 16) for x in ent])
        assert sum(len(p) for p in per_warp) == nstr
        tiles.append((slot_row[sb:sb + ns].copy(), per_warp))
    return tiles


def set_devices(n):
    """Process-wide default number of GPUs per plan (plans created afterwards with device=-1); -1 = all."""
    check(load().qcdmet_set_devices(int(n)))


def get_devices():
    return int(load().qcdmet_get_devices())


def measure_fp64_peak(device=-1, seconds=0.3):
    """Sustained DMMA.8x8x4 rate of the device in TFLOP/s (the roofline denominator, measured in-process)."""
    out = ctypes.c_double(0.0)
    check(load().qcdmet_measure_fp64_peak(int(device), float(seconds), ctypes.byref(out)))
    return out.value


class _PinnedHolder:
    """Owns one cudaHostAlloc block; freed when the last numpy view of it is collected."""

    def __init__(self, nbytes):
        self.lib = load()
        self.ptr = self.lib.qcdmet_host_alloc(int(nbytes))
        if not self.ptr:
            raise QcdmetError(-4, self.lib.qcdmet_last_error().decode())

    def __del__(self):
        try:
            if self.ptr:
                self.lib.qcdmet_host_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except (AttributeError, TypeError):
            pass


def pinned_empty(shape, order="F"):
    """float64 numpy array in page-locked host memory (cudaHostAlloc): device<->host copies reach it at the
    full PCIe rate of every GPU.  The memory lives as long as any view of the array."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    holder = _PinnedHolder(8 * max(n, 1))
    carr = (ctypes.c_double * max(n, 1)).from_address(holder.ptr)
    carr._holder = holder                       # numpy keeps carr alive, carr keeps the allocation alive
    return np.frombuffer(carr, dtype=np.float64, count=n).reshape(shape, order=order)


class Plan:
    """Thin RAII wrapper over qcdmet_plan_* (device-resident perturbation list + fragment rows).

    devices: number of GPUs the plan spans inside this process (None: the library default, i.e. 1 unless
    qcdmet_set_devices / $QCDMET_B200_DEVICES says otherwise and device == -1)."""

    def __init__(self, norb, num_pairs, h1start, h1row, h1col, row_start, rows=None, row_idx=None,
                 device=-1, devices=None):
        self.lib = load()
        self.norb, self.num_pairs = int(norb), int(num_pairs)
        self._keep = [as_i32(h1start), as_i32(h1row), as_i32(h1col), as_i32(row_start), as_f64(rows),
                      as_i32(row_idx)]
        hs, hr, hc, rs, r, ri = self._keep
        self.nterms = len(hs) - 1
        self.nfrag = len(rs) - 1
        if hr.size == 0:
            hr = hc = np.zeros(1, dtype=np.int32)
        self.handle = ctypes.c_void_p()
        check(self.lib.qcdmet_plan_create(ctypes.byref(self.handle), device, self.norb, self.nterms,
                                          self.num_pairs, _ptr(hs), _ptr(hr), _ptr(hc), self.nfrag,
                                          _ptr(rs), _ptr(r), _ptr(ri)))
        self.nresid = int(self.lib.qcdmet_plan_num_residuals(self.handle))
        self.col_begin, self.col_end = 0, self.nterms
        self._pinned_out = None
        if devices is not None:
            self.set_devices(devices)

    def set_devices(self, n):
        check(self.lib.qcdmet_plan_set_devices(self.handle, int(n)))

    @property
    def num_devices(self):
        return int(self.lib.qcdmet_plan_num_devices(self.handle))

    def close(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.qcdmet_plan_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except (AttributeError, TypeError):      # interpreter shutdown: the library object is already gone
            pass

    __del__ = close

    def set_columns(self, begin, end):
        check(self.lib.qcdmet_plan_set_columns(self.handle, int(begin), int(end)))
        self.col_begin, self.col_end = int(begin), int(end)

    def set_kernel(self, kernel):
        check(self.lib.qcdmet_plan_set_kernel(self.handle, int(kernel)))

    def set_stream(self, stream_ptr):
        check(self.lib.qcdmet_plan_set_stream(self.handle, ctypes.c_void_p(stream_ptr)))

    def set_output_device(self, ptr, capacity):
        """Jacobian slice goes to caller-owned device memory (ptr: int address, capacity in doubles)."""
        check(self.lib.qcdmet_plan_set_output_device(self.handle, ctypes.c_void_p(ptr), int(capacity)))

    def export_jacobian(self):
        """64-byte CUDA IPC handle of this plan's full (Q x P) Jacobian buffer."""
        buf = ctypes.create_string_buffer(64)
        check(self.lib.qcdmet_plan_export_jacobian(self.handle, buf))
        return buf.raw

    def attach_peers(self, rank, handles):
        """handles: list of the 64-byte IPC handles of all ranks (own entry ignored)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        check(self.lib.qcdmet_plan_attach_peers(self.handle, len(handles), int(rank), blob))

    def attach_peer_pointers(self, rank, ptrs):
        arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])
        check(self.lib.qcdmet_plan_attach_peer_pointers(self.handle, len(ptrs), int(rank), arr))

    def detach_peers(self):
        check(self.lib.qcdmet_plan_detach_peers(self.handle))

    def set_h0(self, h0):
        h0 = as_f64(h0)
        assert h0.size == self.norb * self.norb
        check(self.lib.qcdmet_plan_set_h0(self.handle, _ptr(h0)))

    def set_orbitals(self, c, e):
        """c: (N, N) with eigenvector k in column k (numpy convention); e ascending."""
        c_colmajor = np.asfortranarray(np.asarray(c, dtype=np.float64))
        e = as_f64(e)
        check(self.lib.qcdmet_plan_set_orbitals(self.handle, c_colmajor.ctypes.data_as(ctypes.c_void_p),
                                                _ptr(e)))

    def compute(self):
        check(self.lib.qcdmet_plan_compute(self.handle))

    def sync(self):
        check(self.lib.qcdmet_plan_sync(self.handle))

    def jacobian(self, out=None, pinned=False):
        """(Q, ncols) float64, Fortran order: the memory of the reference's returned array.

        out: caller's array (any host memory).  pinned=True (and out None): the plan's own page-locked
        buffer is reused -- the returned array is overwritten by the next call with pinned=True."""
        ncols = self.col_end - self.col_begin
        if out is None and pinned:
            if self._pinned_out is None or self._pinned_out.shape != (self.nresid, ncols):
                self._pinned_out = pinned_empty((self.nresid, ncols), order="F")
            out = self._pinned_out
        elif out is None:
            out = np.empty((self.nresid, ncols), dtype=np.float64, order="F")
        assert out.flags.f_contiguous and out.shape == (self.nresid, ncols)
        check(self.lib.qcdmet_plan_get_jacobian(self.handle, _ptr(out)))
        return out

    def jacobian_columns(self, c0, c1, out=None):
        """Columns [c0, c1) of the computed slice, (Q, c1 - c0) Fortran order."""
        if out is None:
            out = np.empty((self.nresid, c1 - c0), dtype=np.float64, order="F")
        assert out.flags.f_contiguous and out.shape == (self.nresid, c1 - c0)
        check(self.lib.qcdmet_plan_get_jacobian_columns(self.handle, int(c0), int(c1), _ptr(out)))
        return out

    def timings(self):
        """dict of the last step's breakdown in ms (None = not measured)."""
        ms = np.zeros(5, dtype=np.float64)
        check(self.lib.qcdmet_plan_timings(self.handle, _ptr(ms)))
        keys = ("h2d_ms", "syevd_ms", "panels_ms", "kernel_ms", "d2h_ms")
        return {k: (float(v) if v >= 0 else None) for k, v in zip(keys, ms)}

    def checksum(self):
        """(sum, sum of squares, xor of the bit patterns, all devices identical) of the Jacobian on the device."""
        out = np.zeros(2, dtype=np.float64)
        bits = ctypes.c_ulonglong(0)
        same = ctypes.c_int(1)
        check(self.lib.qcdmet_plan_jacobian_checksum(self.handle, _ptr(out), ctypes.byref(bits), ctypes.byref(same)))
        return float(out[0]), float(out[1]), int(bits.value), bool(same.value)

    def flush_l2(self):
        check(self.lib.qcdmet_plan_flush_l2(self.handle))

    @property
    def info(self):
        return self.lib.qcdmet_plan_info(self.handle).decode()

    def orbitals_device(self):
        """(address of C, address of e) on the plan's (first) device."""
        c, e = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.lib.qcdmet_plan_orbitals_device(self.handle, ctypes.byref(c), ctypes.byref(e)))
        return c.value, e.value

    def commit_orbitals(self):
        check(self.lib.qcdmet_plan_commit_orbitals(self.handle))

    def lm_step(self, resid, lam):
        """Damped normal-equation step from the device-resident Jacobian.  Returns (step, max|J^T r|, g.step)."""
        resid = as_f64(resid)
        assert resid.size == self.nresid
        step = np.empty(self.col_end - self.col_begin, dtype=np.float64)
        info = np.zeros(2, dtype=np.float64)
        check(self.lib.qcdmet_plan_lm_step(self.handle, _ptr(resid), float(lam), _ptr(step), _ptr(info)))
        return step, float(info[0]), float(info[1])

    def set_lm_method(self, method):
        """LM_NORMAL (equilibrated normal equations + Cholesky) or LM_QR (QR of [J; sqrt(lambda) D])."""
        check(self.lib.qcdmet_plan_set_lm_method(self.handle, int(method)))

    def gradient(self, resid):
        """2 J^T resid from the device-resident Jacobian."""
        resid = as_f64(resid)
        assert resid.size == self.nresid
        grad = np.empty(self.col_end - self.col_begin, dtype=np.float64)
        check(self.lib.qcdmet_plan_gradient(self.handle, _ptr(resid), _ptr(grad)))
        return grad

    def dblocks(self):
        out = np.empty(self.nresid, dtype=np.float64)
        check(self.lib.qcdmet_plan_get_dblocks(self.handle, _ptr(out)))
        return out

    def one_rdm(self):
        """Mean-field 1-RDM 2 C_o C_o^T of the plan's orbitals (what construct1RDM_loc returns)."""
        out = np.empty((self.norb, self.norb), dtype=np.float64)
        check(self.lib.qcdmet_plan_get_one_rdm(self.handle, _ptr(out)))
        return out

    def construct_baths(self, imp_start, imp_idx, num_bath_req=None, threshold=1e-13):
        """Schmidt baths of all fragments from the plan's orbitals (qcdmet_plan_construct_baths).  Returns
        (num_bath [F], occupations [S] in rank order, bath [S, Norb]: row imp_start[f] + k = k-th bath orbital)."""
        imp_start, imp_idx = as_i32(imp_start), as_i32(imp_idx)
        nfrag = imp_start.size - 1
        req = as_i32(num_bath_req)
        total = int(imp_start[-1])
        num_bath = np.zeros(nfrag, dtype=np.int32)
        occ = np.zeros(total, dtype=np.float64)
        bath = np.zeros((total, self.norb), dtype=np.float64)
        check(self.lib.qcdmet_plan_construct_baths(self.handle, nfrag, _ptr(imp_start), _ptr(imp_idx), _ptr(req),
                                                   float(threshold), _ptr(num_bath), _ptr(occ), _ptr(bath)))
        return num_bath, occ, bath

    def embed_hubbard(self, hopping, hubbard_u, imp_start, imp_idx, num_bath_req=None, threshold=1e-13, want_tei=True,
                      want_bath=True):
        """Baths and embedded Hubbard Hamiltonians of all fragments (qcdmet_plan_embed_hubbard).  Returns a dict with
        num_bath [F], occ [S], bath [S, Norb] or None, core_electrons [F] and per-fragment lists oei / fock (n x n) and
        tei (n x n x n x n or None), n = impurity + kept bath orbitals."""
        imp_start, imp_idx = as_i32(imp_start), as_i32(imp_idx)
        nfrag = imp_start.size - 1
        m = np.diff(imp_start)
        req = m.copy() if num_bath_req is None else np.minimum(np.asarray(num_bath_req, dtype=np.int64), m)
        cap = (m + req).astype(np.int64)
        off2 = np.concatenate([[0], np.cumsum(cap ** 2)])
        off4 = np.concatenate([[0], np.cumsum(cap ** 4)])
        total = int(imp_start[-1])
        num_bath = np.zeros(nfrag, dtype=np.int32)
        occ = np.zeros(total, dtype=np.float64)
        bath = np.zeros((total, self.norb), dtype=np.float64) if want_bath else None
        oei = np.zeros(int(off2[-1]), dtype=np.float64)
        fock = np.zeros(int(off2[-1]), dtype=np.float64)
        tei = np.zeros(int(off4[-1]), dtype=np.float64) if want_tei else None
        core = np.zeros(nfrag, dtype=np.float64)
        h = as_f64(hopping)
        check(self.lib.qcdmet_plan_embed_hubbard(self.handle, _ptr(h), float(hubbard_u), nfrag, _ptr(imp_start), _ptr(imp_idx),
                                                 _ptr(as_i32(num_bath_req)), float(threshold), _ptr(num_bath), _ptr(occ),
                                                 _ptr(bath), _ptr(oei), _ptr(fock), _ptr(tei), _ptr(core)))
        out = {"num_bath": num_bath, "occ": occ, "bath": bath, "core_electrons": core, "oei": [], "fock": [], "tei": []}
        for f in range(nfrag):
            c, n = int(cap[f]), int(m[f] + num_bath[f])
            out["oei"].append(oei[off2[f]:off2[f + 1]].reshape(c, c)[:n, :n].copy())
            out["fock"].append(fock[off2[f]:off2[f + 1]].reshape(c, c)[:n, :n].copy())
            out["tei"].append(None if tei is None else tei[off4[f]:off4[f + 1]].reshape(c, c, c, c)[:n, :n, :n, :n].copy())
        return out

    def eigenvalues(self):
        out = np.empty(self.norb, dtype=np.float64)
        check(self.lib.qcdmet_plan_get_eigenvalues(self.handle, _ptr(out)))
        return out

    def orbitals(self):
        out = np.empty((self.norb, self.norb), dtype=np.float64, order="F")
        check(self.lib.qcdmet_plan_get_orbitals(self.handle, _ptr(out)))
        return out

    def device_ptr(self):
        return self.lib.qcdmet_plan_jacobian_device(self.handle)

    @property
    def kernel_used(self):
        return self.lib.qcdmet_plan_kernel_used(self.handle)

    @property
    def symmetric_used(self):
        return bool(self.lib.qcdmet_plan_symmetric_used(self.handle))

    @property
    def split_used(self):
        return bool(self.lib.qcdmet_plan_split_used(self.handle))

    @property
    def launches(self):
        return self.lib.qcdmet_plan_launches(self.handle)

    @property
    def flops(self):
        return self.lib.qcdmet_plan_flops(self.handle)

    def kernel_time_ms(self):
        ms = ctypes.c_double(-1.0)
        check(self.lib.qcdmet_plan_kernel_time_ms(self.handle, ctypes.byref(ms)))
        return ms.value
