"""CPU checker for the bath / embedding row of SURVEY.md 8 (f-4).  TEST INFRASTRUCTURE ONLY: imported by tests/ and
nothing else; the product path (qcdmet_b200/) never touches it.

Restates, step by step and with the reference's own (slow) formulation, what the reference does per fragment:

    construct_bath      src/qcdmethelper.py:124-168   eigh of the (N - m) x (N - m) environment block of the 1-RDM
    fock_from_core      src/localintegrals_hubbard.py:71-78    loc_rhf_fock_bis
    embedded_oei        src/localintegrals_hubbard.py:84-87    dmet_oei
    embedded_fock       src/localintegrals_hubbard.py:89-92    dmet_fock
    embedded_tei        src/localintegrals_hubbard.py:105-116  dmet_tei (the O(n^4 N) loop, kept as a loop)
    core_density        src/dmet.py:209-220                    occupation rounding, loc2dmet diag(core occupations) loc2dmet^T

PARITY UNPINNED for this row: the reference's Python layer is Python 2 (print statements, integer division) and ships no
tests or golden vectors for these functions, so it cannot be run here; unlike the response kernel
(oracle/_ref, compiled from the reference's own source) this file is pinned only by the invariants the reference
itself asserts (orthonormal loc2dmet, :164) and by the Schmidt-decomposition properties checked in
tests/test_bath_oracle.py.
"""
from __future__ import annotations

import numpy as np


def construct_bath(one_rdm, impurity_flags, num_bath, threshold=1e-13):
    """(numBathOrbs, loc2dmet, coreOccupations) as src/qcdmethelper.py:124-168 returns them.

    loc2dmet columns: impurity unit vectors | bath | pure environment by decreasing occupation."""
    one_rdm = np.asarray(one_rdm, dtype=float)
    flags = np.asarray(impurity_flags).reshape(-1)
    ntot = flags.size
    # :126-132  environment block through the outer-product mask of (1 - impurity flags)
    envflag = 1 - flags
    mask = np.outer(envflag, envflag) == 1
    nenv = int(np.sum(envflag))
    env_rdm = one_rdm[mask].reshape(nenv, nenv)
    nimp = int(np.sum(flags))
    # :137-140  natural orbitals of the environment; key = max(-n, n - 2), ascending: occupations closest to 1 first
    occ, vec = np.linalg.eigh(env_rdm)
    key = np.maximum(-occ, occ - 2.0)
    idx = key.argsort()
    keep = int(np.sum(-key[idx] > threshold))                # :140  entangled orbitals only
    num_bath = min(keep, int(num_bath))                      # :143
    occ, vec = occ[idx], vec[:, idx]
    # :146-151  the remaining orbitals by decreasing occupation
    rest_occ, rest_vec = occ[num_bath:], vec[:, num_bath:]
    order = (-rest_occ).argsort()
    vec = np.concatenate([vec[:, :num_bath], rest_vec[:, order]], axis=1)
    core = np.concatenate([np.zeros(nimp + num_bath), rest_occ[order]])
    # :153-161  zero columns in front for the impurity, then one zero row per impurity orbital with a 1 in its column
    full = np.zeros((ntot, ntot))
    env_rows = np.flatnonzero(envflag)
    for k, row in enumerate(env_rows):
        full[row, nimp:] = vec[k]
    for counter, row in enumerate(np.flatnonzero(flags)):
        full[row, counter] = 1.0
    # :164  the reference's own assertion
    assert np.linalg.norm(full.T @ full - np.eye(ntot)) < 1e-12
    return num_bath, full, core


def core_density(loc2dmet, core_occupations, core_cutoff=0.01):
    """src/dmet.py:209-220: occupations within core_cutoff of 0 / 2 are rounded, anything else is a badly chosen bath
    (the reference prints a complaint and stops on an assertion; the checker raises)."""
    occ = np.array(core_occupations, dtype=float)
    for cnt in range(occ.size):
        if occ[cnt] < core_cutoff:
            occ[cnt] = 0.0
        elif occ[cnt] > 2.0 - core_cutoff:
            occ[cnt] = 2.0
        else:
            raise ValueError("bath orbital with occupation %g left in the environment" % occ[cnt])
    return loc2dmet @ np.diag(occ) @ loc2dmet.T, occ


def fock_from_core(hopping, hubbard_u, dm_loc):
    """src/localintegrals_hubbard.py:71-78: hopping + diag(U/2 * DM[x, x])"""
    n = hopping.shape[0]
    jk = np.zeros(n)
    for orb in range(n):
        jk[orb] = 0.5 * hubbard_u * dm_loc[orb, orb]
    return np.asarray(hopping, dtype=float) + np.diag(jk)


def embedded_oei(hopping, loc2dmet, num_active):
    """src/localintegrals_hubbard.py:84-87"""
    c = loc2dmet[:, :num_active]
    return c.T @ np.asarray(hopping, dtype=float) @ c


def embedded_fock(hopping, hubbard_u, loc2dmet, num_active, core_dm):
    """src/localintegrals_hubbard.py:89-92"""
    c = loc2dmet[:, :num_active]
    return c.T @ fock_from_core(hopping, hubbard_u, core_dm) @ c


def embedded_tei(hubbard_u, loc2dmet, num_active):
    """src/localintegrals_hubbard.py:105-116, as the loop it is (small cases only)"""
    n = loc2dmet.shape[0]
    tei = np.zeros((num_active,) * 4)
    for a in range(num_active):
        for b in range(num_active):
            for c in range(num_active):
                for d in range(num_active):
                    value = 0.0
                    for x in range(n):
                        value += loc2dmet[x, a] * loc2dmet[x, b] * loc2dmet[x, c] * loc2dmet[x, d]
                    tei[a, b, c, d] = hubbard_u * value
    return tei
