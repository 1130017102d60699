#!/usr/bin/env python
"""Fits with fragments of more than 4 orbitals (the shapes of examples/polyyne.py and examples/sn2.py in
QC-DMET): which kernel the library selects, what the site panels of the factorised path cost in device
memory, and -- when a CUDA device is present -- the Jacobian itself.

    python examples/large_fragments.py [polyyne|sn2]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qcdmet_b200 import _lib, workloads   # noqa: E402  (synthetic inputs only)

name = sys.argv[1] if len(sys.argv) > 1 else "sn2"
w = {"polyyne": workloads.polyyne_like, "sn2": workloads.sn2_like}[name]()
kernel, panel_bytes = _lib.auto_kernel(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.row_idx)
names = {_lib.KERNEL_GENERAL: "general", _lib.KERNEL_DMMA: "fused DMMA", _lib.KERNEL_FACTORISED: "factorised (panels + contraction)"}
print(f"{w.name}: Norb={w.norb} Nocc={w.num_pairs} fragments={w.meta['frag_sizes']} -> P={w.nterms} Q={w.nresid}")
print(f"kernel: {names[kernel]}; site panels and tiled operands: {panel_bytes / 1e6:.1f} MB; "
      f"contraction tiles per column: {len(_lib.contract_schedule(w.row_start))}")
if _lib.load().qcdmet_device_count() == 0:
    sys.exit("no CUDA device: the library has no CPU fallback (selection and schedule above are host-only queries)")
plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
plan.set_h0(w.h0)
plan.compute(); plan.sync()                       # first call: module load
t0 = time.time()
plan.compute(); plan.sync()
t1 = time.time()
jac = plan.jacobian()
print(f"Jacobian {jac.shape} in {1e3 * (t1 - t0):.2f} ms (kernels only, {plan.launches} launches, "
      f"{plan.flops / (t1 - t0) * 1e-12:.1f} TFLOP/s executed), |J|_max = {np.abs(jac).max():.4f}")
off = 0
for a, b in zip(w.row_start[:-1], w.row_start[1:]):
    r = int(b - a)
    blk = jac[off:off + r * r].reshape(r, r, -1, order="F")
    assert np.array_equal(blk, blk.transpose(1, 0, 2)), "fragment blocks are exactly symmetric"
    off += r * r
plan.close()
