"""Small dense full-CI impurity solver with the reference's solver signature.

Stand-in for src/chemps2.py (PyCheMPS2 FCI, absent here): same arguments, same return values and the
same impurity-energy formula (src/chemps2.py:119-125), for active spaces of a few orbitals.  Used
identically on both sides of the DMET-loop parity harness; not part of the hot path.
"""
from __future__ import annotations

import itertools

import numpy as np

_CACHE = {}


def _strings(norb, nel):
    return [sum(1 << k for k in occ) for occ in itertools.combinations(range(norb), nel)]


def _excitation_ops(norb, nel):
    """E_ij = a_i^dagger a_j on the space of nel-electron strings: dense (norb, norb, dim, dim)."""
    key = (norb, nel)
    if key in _CACHE:
        return _CACHE[key]
    strs = _strings(norb, nel)
    index = {s: k for k, s in enumerate(strs)}
    dim = len(strs)
    ops = np.zeros((norb, norb, dim, dim))
    for col, s in enumerate(strs):
        for j in range(norb):
            if not (s >> j) & 1:
                continue
            s1 = s ^ (1 << j)
            sign_j = (-1) ** bin(s1 & ((1 << j) - 1)).count("1")
            for i in range(norb):
                if (s1 >> i) & 1:
                    continue
                s2 = s1 | (1 << i)
                sign_i = (-1) ** bin(s1 & ((1 << i) - 1)).count("1")
                ops[i, j, index[s2], col] += sign_i * sign_j
    _CACHE[key] = ops
    return ops


def solve(CONST, OEI, FOCK, TEI, Norb, Nel, Nimp, chempot_imp=0.0, printoutput=False):
    """Ground state of  H = sum_ij T_ij E_ij + 1/2 sum_ijkl (ij|kl) (E_ij E_kl - delta_jk E_il),
    T = FOCK (minus chempot_imp on the impurity diagonal, src/chemps2.py:45-47), TEI in chemists' notation.
    Returns (ImpurityEnergy, OneRDM) exactly as src/chemps2.py:119-125 defines them."""
    assert Nel % 2 == 0 and Norb <= 6, "dense FCI stand-in: at most 6 active orbitals"
    na = Nel // 2
    ea = _excitation_ops(Norb, na)                # alpha == beta string space
    da = ea.shape[2]
    dim = da * da
    eye = np.eye(da)
    # spin-summed excitation operators on the product space |alpha> x |beta>
    E = np.empty((Norb, Norb, dim, dim))
    for i in range(Norb):
        for j in range(Norb):
            E[i, j] = np.kron(ea[i, j], eye) + np.kron(eye, ea[i, j])
    T = np.array(FOCK, dtype=float, copy=True)
    if chempot_imp != 0.0:
        T[np.arange(Nimp), np.arange(Nimp)] -= chempot_imp
    teff = T - 0.5 * np.einsum('ikkj->ij', TEI)
    H = np.einsum('ij,ijab->ab', teff, E)
    EE = E.reshape(Norb * Norb, dim, dim)
    V = np.asarray(TEI, dtype=float).reshape(Norb * Norb, Norb * Norb)
    VE = np.tensordot(V, EE, axes=(1, 0))          # sum_kl (ij|kl) E_kl
    H += 0.5 * np.einsum('xab,xbc->ac', EE, VE)
    H = 0.5 * (H + H.T)
    w, v = np.linalg.eigh(H)
    gs = v[:, 0]
    # 2-RDM in chemists' notation: G[i,j,k,l] = <E_ij E_kl> - delta_jk <E_il>
    Egs = np.einsum('xab,b->xa', EE, gs)           # E_kl |gs>
    Et_gs = np.einsum('xba,b->xa', EE, gs)         # E_ij^T |gs> = E_ji |gs>
    one = np.einsum('a,xa->x', gs, Egs).reshape(Norb, Norb)
    two = np.einsum('xa,ya->xy', Et_gs, Egs).reshape(Norb, Norb, Norb, Norb)
    for j in range(Norb):
        two[:, j, j, :] -= one
    OneRDM = np.einsum('ijkk->ij', two) / (Nel - 1)
    ImpurityEnergy = CONST
    ImpurityEnergy += 0.5 * np.einsum('ij,ij->', OneRDM[:Nimp, :], OEI[:Nimp, :] + FOCK[:Nimp, :])
    ImpurityEnergy += 0.5 * np.einsum('ijkl,ijkl->', two[:Nimp, :, :, :], TEI[:Nimp, :, :, :])
    return (ImpurityEnergy, OneRDM)
