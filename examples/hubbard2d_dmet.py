#!/usr/bin/env python
"""2-D Hubbard DMET, 2 x 2 impurities on an l x l lattice (the shape of BASELINE.json configs[1]; cf. examples/hubbard.py
of QC-DMET).  Everything around the impurity solver runs on the GPU(s): the umat-fit Jacobian and residual share one
eigensolve per Levenberg-Marquardt point, and the Schmidt baths + embedded Hamiltonians of all fragments come from one
device call per DMET iteration (embed_on_gpu) instead of one (N-4) x (N-4) eigh per fragment.  The 8-orbital embedded
problems are solved by the FCI stand-in (Davidson route).

    python examples/hubbard2d_dmet.py [l] [U] [n_gpus]        # n_gpus > 1: one plan spans several GPUs, same calls
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcdmet_b200 import _lib, dmet, localintegrals_hubbard as lh, workloads   # noqa: E402

l = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hubbard_u = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
if len(sys.argv) > 3:
    _lib.set_devices(int(sys.argv[3]))          # what $QCDMET_B200_DEVICES does from the job script
nsites = l * l

# PBC x anti-PBC hopping, sites ordered plaquette by plaquette: orbitals 4k .. 4k+3 are one 2 x 2 impurity
myInts = lh.localintegrals_hubbard(workloads.hubbard2d_hopping(l, l), hubbard_u, nsites)
impurityClusters = []
for k in range(nsites // 4):
    v = np.zeros(nsites, dtype=int)
    v[4 * k:4 * k + 4] = 1
    impurityClusters.append(v)

theDMET = dmet.dmet(myInts, impurityClusters, True, method='ED', SCmethod='LSTSQ', fitImpBath=True)
theDMET.relaxation = 0.5            # the impurity+bath fit oscillates without damping on these lattices
theDMET.embed_on_gpu = True         # baths + embedded integrals of all fragments on the device
t0 = time.time()
energy = theDMET.doselfconsistent(verbose=True)
print(f"RHF energy / site  = {myInts.fullEhf / nsites:.8f}")
print(f"DMET energy / site = {energy / nsites:.8f}   ({theDMET.iterations} iterations, {time.time() - t0:.1f} s, "
      f"{_lib.get_devices()} GPU(s) per plan)")
print(f"Time cf func = {theDMET.time_func:.3f} s, cf grad = {theDMET.time_grad:.3f} s, dmet ed = {theDMET.time_ed:.3f} s, "
      f"dmet cf = {theDMET.time_cf:.3f} s")
print("The u-matrix (first impurity) =")
print(theDMET.umat[:4, :4])
