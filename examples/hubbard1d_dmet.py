#!/usr/bin/env python
"""1-D Hubbard DMET with the reference's workflow (cf. examples/hubbard.py of QC-DMET): build the lattice
integrals, tile the chain with 2-site impurities, run the self-consistent loop.  The umat fit uses the GPU
Jacobian (libqcdmet_b200); the impurity problems are solved by the dense-FCI stand-in.  Needs a CUDA device.

    python examples/hubbard1d_dmet.py [nsites] [U]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcdmet_b200 import dmet, localintegrals_hubbard as lh   # noqa: E402

nsites = int(sys.argv[1]) if len(sys.argv) > 1 else 24
hubbard_u = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
cluster = 2

myInts = lh.localintegrals_hubbard(lh.hubbard_chain(nsites), hubbard_u, nsites)   # half filling
impurityClusters = []
for k in range(nsites // cluster):
    v = np.zeros(nsites, dtype=int)
    v[cluster * k:cluster * (k + 1)] = 1
    impurityClusters.append(v)

isTranslationInvariant = True
theDMET = dmet.dmet(myInts, impurityClusters, isTranslationInvariant, method='ED', SCmethod='LSTSQ', fitImpBath=True)
energy = theDMET.doselfconsistent(verbose=True)
print(f"RHF energy / site  = {myInts.fullEhf / nsites:.8f}")
print(f"DMET energy / site = {energy / nsites:.8f}   ({theDMET.iterations} iterations)")
print(f"Time cf func = {theDMET.time_func:.3f} s, cf grad = {theDMET.time_grad:.3f} s, dmet ed = {theDMET.time_ed:.3f} s, "
      f"dmet cf = {theDMET.time_cf:.3f} s")
print("The u-matrix =")
print(theDMET.umat[:cluster, :cluster])
