"""Small full-CI impurity solver with the reference's solver signature.

Stand-in for src/chemps2.py (PyCheMPS2 FCI, absent here): same arguments, same return values and the
same impurity-energy formula (src/chemps2.py:119-125).  Used identically on both sides of the DMET-loop
parity harness; not part of the hot path.

Two routes, chosen by the size of the active space:
  * <= 6 orbitals: the Hamiltonian as a dense matrix on the determinant space, numpy eigh;
  * 7 .. 10 orbitals (the 2x2-impurity + 4-bath problem of the 2-D Hubbard fits has 8: dimension
    C(8,4)^2 = 4900): the ground state by Davidson iteration.  The wave function is a matrix Psi[alpha, beta]
    over alpha and beta occupation strings and H is applied as sums of Kronecker products,
        H Psi = Ha Psi + Psi Ha^T + sum_ij A_ij Psi (sum_kl (ij|kl) A_kl)^T,
    with A_ij the (sparse) single-spin excitation operators; no dim x dim matrix is ever built.
"""
from __future__ import annotations

import itertools

import numpy as np

_CACHE = {}
DENSE_MAX_ORBITALS = 6


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


def _energy_and_rdm(CONST, OEI, FOCK, TEI, Nel, Nimp, one, two):
    """src/chemps2.py:119-125: 1-RDM from the 2-RDM, impurity energy from the impurity rows."""
    OneRDM = np.einsum('ijkk->ij', two) / (Nel - 1)
    ImpurityEnergy = CONST
    ImpurityEnergy += 0.5 * np.einsum('ij,ij->', OneRDM[:Nimp, :], OEI[:Nimp, :] + FOCK[:Nimp, :])
    ImpurityEnergy += 0.5 * np.einsum('ijkl,ijkl->', two[:Nimp, :, :, :], TEI[:Nimp, :, :, :])
    return (ImpurityEnergy, OneRDM)


def solve(CONST, OEI, FOCK, TEI, Norb, Nel, Nimp, chempot_imp=0.0, printoutput=False):
    """Ground state of  H = sum_ij T_ij E_ij + 1/2 sum_ijkl (ij|kl) (E_ij E_kl - delta_jk E_il),
    T = FOCK (minus chempot_imp on the impurity diagonal, src/chemps2.py:45-47), TEI in chemists' notation.
    Returns (ImpurityEnergy, OneRDM) exactly as src/chemps2.py:119-125 defines them."""
    assert Nel % 2 == 0 and Norb <= 10, "FCI stand-in: at most 10 active orbitals"
    T = np.array(FOCK, dtype=float, copy=True)
    if chempot_imp != 0.0:
        T[np.arange(Nimp), np.arange(Nimp)] -= chempot_imp
    TEI = np.asarray(TEI, dtype=float)
    if Norb <= DENSE_MAX_ORBITALS:
        one, two = _ground_state_dense(T, TEI, Norb, Nel)
    else:
        one, two = _ground_state_davidson(T, TEI, Norb, Nel)
    return _energy_and_rdm(CONST, np.asarray(OEI, dtype=float), np.asarray(FOCK, dtype=float), TEI, Nel, Nimp, one, two)


# ------------------------------------------------------------------------------------------
# dense route
# ------------------------------------------------------------------------------------------
def _ground_state_dense(T, TEI, Norb, Nel):
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
    teff = T - 0.5 * np.einsum('ikkj->ij', TEI)
    H = np.einsum('ij,ijab->ab', teff, E)
    EE = E.reshape(Norb * Norb, dim, dim)
    V = TEI.reshape(Norb * Norb, Norb * Norb)
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
    return one, two


# ------------------------------------------------------------------------------------------
# Davidson route
# ------------------------------------------------------------------------------------------
class _KronHamiltonian:
    """H on Psi[alpha, beta] as Kronecker sums (alpha and beta string spaces are the same: na == nb)."""

    def __init__(self, T, TEI, Norb, Nel):
        n2 = Norb * Norb
        A = _excitation_ops(Norb, Nel // 2).reshape(n2, -1)          # (n2, da * da), row x = (i, j)
        self.da = da = int(round(np.sqrt(A.shape[1])))
        self.n2 = n2
        self.A = A.reshape(n2, da, da)
        V = TEI.reshape(n2, n2)
        teff = T - 0.5 * np.einsum('ikkj->ij', TEI)
        VA = (V @ A).reshape(n2, da, da)                             # sum_kl (ij|kl) A_kl
        # same-spin part: one-body + 1/2 sum (ij|kl) A_ij A_kl
        self.Hs = np.einsum('x,xab->ab', teff.reshape(-1), self.A) + 0.5 * np.einsum('xab,xbc->ac', self.A, VA)
        # opposite-spin part: sum_x A_x Psi VA_x^T, as two GEMMs over the stacked operators
        # rows (x, a); the excitation operators are sparse (<= one entry per row): CSR when scipy is there
        self.Astack = self.A.reshape(n2 * da, da)
        try:
            from scipy import sparse
            self.Astack = sparse.csr_matrix(self.Astack)
        except ImportError:
            pass
        self.VAstack = VA.transpose(1, 0, 2).reshape(da, n2 * da)    # [c, (x, b)] = VA_x[c, b]
        self.diag = (np.diag(self.Hs)[:, None] + np.diag(self.Hs)[None, :]
                     + np.einsum('xa,xb->ab', np.einsum('xaa->xa', self.A), np.einsum('xbb->xb', VA)))

    def apply(self, psi):
        da, n2 = self.da, self.n2
        out = self.Hs @ psi + psi @ self.Hs.T
        X = np.asarray(self.Astack @ psi).reshape(n2, da, da)        # X_x = A_x Psi
        # sum_x X_x VA_x^T : [a, (x, b)] @ [(x, b), c]
        out += X.transpose(1, 0, 2).reshape(da, n2 * da) @ self.VAstack.T
        return out

    def excitations(self, psi):
        """E_kl Psi for every (k, l): (n2, da * da)."""
        da, n2 = self.da, self.n2
        left = np.asarray(self.Astack @ psi).reshape(n2, da, da)     # A_kl Psi
        right = np.einsum('ab,xcb->xac', psi, self.A)                # Psi A_kl^T
        return (left + right).reshape(n2, da * da)


def _davidson(ham, guess, tol=1e-11, max_iter=600, max_space=28, keep=4):
    """Lowest eigenpair by Davidson iteration with the diagonal preconditioner and thick restarts (the `keep`
    lowest Ritz vectors survive a restart).  Deterministic: fixed start vector, fixed arithmetic order."""
    da = ham.da
    diag = ham.diag.reshape(-1)
    vs, ws = [], []

    def add(vec, hvec=None):
        for _ in range(2):                                           # Gram-Schmidt, twice
            for b in vs:
                vec = vec - b * np.dot(b, vec)
        nrm = np.linalg.norm(vec)
        if nrm < 1e-13:
            return False
        vec = vec / nrm
        vs.append(vec)
        ws.append(ham.apply(vec.reshape(da, da)).reshape(-1))
        return True

    add(np.array(guess, dtype=float, copy=True))
    theta, u = None, None
    for _ in range(max_iter):
        Vm, Wm = np.array(vs), np.array(ws)
        small = Vm @ Wm.T
        small = 0.5 * (small + small.T)
        th, s = np.linalg.eigh(small)
        theta, s0 = th[0], s[:, 0]
        u = s0 @ Vm
        hu = s0 @ Wm
        r = hu - theta * u
        if np.linalg.norm(r) < tol:
            break
        if len(vs) >= max_space:                                     # thick restart
            k = min(keep, len(vs))
            newv, neww = s[:, :k].T @ Vm, s[:, :k].T @ Wm
            vs[:], ws[:] = list(newv), list(neww)
        denom = diag - theta
        denom = np.where(np.abs(denom) < 1e-4, 1e-4, denom)
        if not add(r / denom) and not add(r):
            break
    else:
        raise RuntimeError("FCI Davidson iteration did not converge")
    u = u / np.linalg.norm(u)
    if u[np.argmax(np.abs(u))] < 0:
        u = -u
    return theta, u


def _ground_state_davidson(T, TEI, Norb, Nel):
    """Work in the eigenbasis of the one-body matrix: there the aufbau determinant is a good start vector and
    the diagonal a good preconditioner; the reduced density matrices are rotated back at the end."""
    eps, U = np.linalg.eigh(0.5 * (T + T.T))
    teiR = np.einsum('pqrs,pi,qj,rk,sl->ijkl', TEI, U, U, U, U, optimize=True)
    ham = _KronHamiltonian(np.diag(eps), teiR, Norb, Nel)
    # aufbau determinant (strings are ordered: index 0 = lowest orbitals) plus a small fixed component along every
    # determinant, so that no spin / spatial symmetry sector is excluded from the search
    dim = ham.da * ham.da
    guess = 1e-2 * np.random.default_rng(20240).standard_normal(dim)
    guess[0] += 1.0
    _, gs = _davidson(ham, guess)
    M = ham.excitations(gs.reshape(ham.da, ham.da))                  # M[(k, l)] = E_kl |gs>
    one = (M @ gs).reshape(Norb, Norb)
    # <E_ij E_kl> = (E_ji gs) . (E_kl gs)
    Mt = M.reshape(Norb, Norb, -1).transpose(1, 0, 2).reshape(Norb * Norb, -1)
    two = (Mt @ M.T).reshape(Norb, Norb, Norb, Norb)
    for j in range(Norb):
        two[:, j, j, :] -= one
    one = U @ one @ U.T
    two = np.einsum('ijkl,pi,qj,rk,sl->pqrs', two, U, U, U, U, optimize=True)
    return one, two
