"""DMET-loop parity harness: the same Python-3 loop, lattice integrals and dense-FCI
stand-in solver driven once by the oracle's Jacobian (CPU) and once by the GPU path."""
import numpy as np

from oracle import response_oracle as ro
from qcdmet_b200 import dmet as dmet_mod, localintegrals_hubbard as lh


class OracleDmet(dmet_mod.dmet):
    """dmet whose Jacobian comes from the CPU oracle (restatement of src/dmet.py:477-525 on top of
    lib/rhf_response.cc) instead of libqcdmet_b200."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.gpu_residual = False          # reference's host route for the residual: no device in this loop

    def rdm_differences_derivative(self, newumatflat):
        rs, rows, idx = self._fragment_rows()
        h0 = self.ints.loc_rhf_fock() + self.flat2square(newumatflat)
        e, c = ro.eigh_sorted(h0)
        hp = self.helper
        return ro.jacobian_from_orbitals(c, e, hp.numPairs, hp.H1start, hp.H1row, hp.H1col, rs, rows, idx)


def chain_clusters(nsites, m):
    out = []
    for k in range(nsites // m):
        v = np.zeros(nsites, dtype=int)
        v[m * k:m * (k + 1)] = 1
        out.append(v)
    return out


CASES = {
    # name: (sites, U, cluster size, translation invariant, fitImpBath, SCmethod)
    "chain12_U4_TI_bath": (12, 4.0, 2, True, True, "LSTSQ"),
    "chain12_U4_imp_only": (12, 4.0, 2, False, False, "LSTSQ"),
    "chain8_U2_bath": (8, 2.0, 2, False, True, "LSTSQ"),
    "chain12_U6_bfgs": (12, 6.0, 2, False, False, "BFGS"),
    "chain16_U4_single_site": (16, 4.0, 1, False, False, "LSTSQ"),
}


def build(case, cls):
    n, u, m, ti, bath, sc = CASES[case]
    ints = lh.localintegrals_hubbard(lh.hubbard_chain(n), u, n)
    return cls(ints, chain_clusters(n, m), ti, method='ED', SCmethod=sc, fitImpBath=bath)
