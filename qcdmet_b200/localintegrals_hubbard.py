"""Hubbard-model integral provider with the interface of the reference's localintegrals_hubbard
(src/localintegrals_hubbard.py:23-116), restated for Python 3 and without the dense N^4 ERI tensor
(:54-56, 880 GB for the 24 x 24 lattice) or the O(n^4 N) Python loop of dmet_tei (:105-116).

Host-side helper for the "next" rows of SURVEY.md 8(f): it feeds the umat-fit path and the DMET-loop
parity harness; it is not part of the measured hot path.
"""
from __future__ import annotations

import numpy as np


class localintegrals_hubbard:

    def __init__(self, hopping, HubbardU, Nelectrons):
        self.activeOEI = np.asarray(hopping, dtype=float)
        self.Norbs = self.activeOEI.shape[0]
        self.Nelec = int(Nelectrons)
        self.HubbardU = float(HubbardU)

        eigvals, eigvecs = np.linalg.eigh(self.activeOEI)
        idx = eigvals.argsort()
        eigvals, eigvecs = eigvals[idx], eigvecs[:, idx]
        assert self.Nelec % 2 == 0
        numPairs = self.Nelec // 2
        assert eigvals[numPairs] - eigvals[numPairs - 1] > 1e-8      # closed shell (:38)

        self.ao2loc = np.eye(self.Norbs)
        self.TI_OK = True
        self.origDMloc = 2 * np.dot(eigvecs[:, :numPairs], eigvecs[:, :numPairs].T)
        self.origJKloc = np.diag(0.5 * self.HubbardU * np.diag(self.origDMloc))
        self.fullEhf = np.einsum('ij,ij->', self.activeOEI + 0.5 * self.origJKloc, self.origDMloc)
        self.activeFOCK = self.activeOEI + self.origJKloc
        self.ERIinMEM = False      # the on-site ERI is never materialised

    def const(self):
        return 0.0

    def loc_oei(self):
        return self.activeOEI

    def loc_rhf_fock(self):
        return self.activeFOCK

    def loc_rhf_fock_bis(self, DMloc):
        return self.activeOEI + np.diag(0.5 * self.HubbardU * np.diag(DMloc))

    def loc_tei(self):
        raise NotImplementedError("the dense N^4 on-site ERI (src/localintegrals_hubbard.py:54-56) is not built")

    def dmet_oei(self, loc2dmet, numActive):
        c = loc2dmet[:, :numActive]
        return np.dot(np.dot(c.T, self.activeOEI), c)

    def dmet_fock(self, loc2dmet, numActive, coreDMloc):
        c = loc2dmet[:, :numActive]
        return np.dot(np.dot(c.T, self.loc_rhf_fock_bis(coreDMloc)), c)

    def dmet_init_guess_rhf(self, loc2dmet, numActive, numPairs, Nimp, chempot_imp):
        c = loc2dmet[:, :numActive]
        Fock_small = np.dot(np.dot(c.T, self.activeFOCK), c)
        if chempot_imp != 0.0:
            Fock_small[np.arange(Nimp), np.arange(Nimp)] -= chempot_imp
        eigvals, eigvecs = np.linalg.eigh(Fock_small)
        eigvecs = eigvecs[:, eigvals.argsort()]
        return 2 * np.dot(eigvecs[:, :numPairs], eigvecs[:, :numPairs].T)

    def dmet_tei(self, loc2dmet, numAct):
        c = loc2dmet[:, :numAct]
        return self.HubbardU * np.einsum('xi,xj,xk,xl->ijkl', c, c, c, c, optimize=True)


def hubbard_chain(nsites, t=1.0, antiperiodic=True):
    """1-D hopping matrix; anti-periodic closure makes half filling closed-shell for nsites % 4 == 0
    (the trick at examples/hubbard.py:40-41)."""
    h = np.zeros((nsites, nsites))
    for i in range(nsites - 1):
        h[i, i + 1] = h[i + 1, i] = -t
    h[0, nsites - 1] = h[nsites - 1, 0] = t if antiperiodic else -t
    return h
