"""Mean-field helper with QC-DMET's interface, backed by libqcdmet_b200.so.

Mirrors class qcdmethelper of the reference (src/qcdmethelper.py:26-168): same constructor and
method names, argument order and return shapes, for the RHF-response path

    construct1RDM_response(doSCF, umat_loc, NOrotation)        src/qcdmethelper.py:84-112
    construct1RDM_loc_response(doSCF, umat, list_H1)           README.md:57 (older name)
    convertH1sparse()                                          src/qcdmethelper.py:48-65

plus the fused entry the Jacobian assembly uses instead of materialising N x N x P:

    response_jacobian(umat_loc, row_start, rows=None, row_idx=None)  ->  (Q, P) Fortran-ordered

All device work goes through the ctypes C ABI (include/qcdmet_b200.h).  No CPU fallback exists
for the response; construct1RDM_loc / construct1RDM_base stay host numpy exactly as in the
reference (they are not part of the response path, SURVEY.md 8a a14 "next").
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class qcdmethelper:

    def __init__(self, theLocalIntegrals, list_H1, altcf, minFunc):
        self.locints = theLocalIntegrals
        assert self.locints.Nelec % 2 == 0
        self.numPairs = self.locints.Nelec // 2          # src/qcdmethelper.py:32
        self.altcf = altcf
        self.minFunc = None
        if self.altcf:
            assert minFunc == 'OEI' or minFunc == 'FOCK_INIT'
            self.minFunc = minFunc

        # list_H1: dense list of N x N 0/1 matrices (reference form, src/dmet.py:118-159), or a
        # pre-built CSR triple (H1start, H1row, H1col): the dense list is 8 N^2 P bytes.
        if isinstance(list_H1, tuple) and len(list_H1) == 3:
            self.list_H1 = None
            self.H1start, self.H1row, self.H1col = (np.ascontiguousarray(a, dtype=ctypes.c_int)
                                                    for a in list_H1)
        else:
            self.list_H1 = list_H1
            self.H1start, self.H1row, self.H1col = self.convertH1sparse()
        self.Nterms = len(self.H1start) - 1
        self._plans = {}
        self.reuse_output = False       # response_jacobian returns a view of the plan's pinned buffer

    # ---------------------------------------------------------------- src/qcdmethelper.py:48-65
    def convertH1sparse(self):
        H1start, H1row, H1col = [0], [], []
        for H1 in self.list_H1:
            rowco, colco = np.where(np.asarray(H1) == 1)
            H1row.extend(int(r) for r in rowco)
            H1col.extend(int(c) for c in colco)
            H1start.append(len(H1row))
        return (np.array(H1start, dtype=ctypes.c_int), np.array(H1row, dtype=ctypes.c_int),
                np.array(H1col, dtype=ctypes.c_int))

    # ---------------------------------------------------------------- src/qcdmethelper.py:67-82
    def construct1RDM_loc(self, doSCF, umat_loc):
        if self.altcf and self.minFunc == 'OEI':
            OEI = self.locints.loc_oei() + umat_loc
        else:
            OEI = self.locints.loc_rhf_fock() + umat_loc
        if doSCF:
            raise NotImplementedError("doSCF=True needs the reference's pyscf-based rhf module "
                                      "(src/rhf.py); it is hard-coded False in src/dmet.py:44")
        return self.construct1RDM_base(OEI, self.numPairs)

    # ---------------------------------------------------------------- src/qcdmethelper.py:114-122
    def construct1RDM_base(self, OEI, myNumPairs):
        eigenvals, eigenvecs = np.linalg.eigh(OEI)
        idx = eigenvals.argsort()
        eigenvecs = eigenvecs[:, idx]
        return 2 * np.dot(eigenvecs[:, :myNumPairs], eigenvecs[:, :myNumPairs].T)

    # ---------------------------------------------------------------- src/qcdmethelper.py:124-168
    def constructbath(self, OneDM, impurityOrbs, numBathOrbs, threshold=1e-13):
        """Schmidt bath from the environment block of the mean-field 1-RDM.  Returns
        (numBathOrbs, loc2dmet [N x N: impurity | bath | environment], coreOccupations).

        Ordering rule of the reference (:139-152): environment natural orbitals are ranked by how far their
        occupation n is from being empty or doubly occupied, d(n) = min(n, 2 - n) (largest first); the first
        numBathOrbs of them (at most those with d > threshold) are the bath, the rest are sorted by decreasing
        occupation (doubly occupied core first, empty virtuals last)."""
        flags = np.abs(np.asarray(impurityOrbs).reshape(-1)).astype(int)
        imp, env = np.flatnonzero(flags), np.flatnonzero(flags == 0)
        nimp, ntot = imp.size, flags.size
        occ, vec = np.linalg.eigh(np.asarray(OneDM)[np.ix_(env, env)])
        depth = np.minimum(occ, 2.0 - occ)                    # distance from an integer occupation
        rank = np.argsort(-depth, kind="stable")
        numBathOrbs = min(int(np.count_nonzero(depth > threshold)), int(numBathOrbs))
        bath, rest = rank[:numBathOrbs], rank[numBathOrbs:]
        rest = rest[np.argsort(-occ[rest], kind="stable")]
        order = np.concatenate([bath, rest])
        coreOccupations = np.concatenate([np.zeros(nimp + numBathOrbs), occ[rest]])
        loc2dmet = np.zeros((ntot, ntot))
        loc2dmet[imp, np.arange(nimp)] = 1.0
        loc2dmet[np.ix_(env, nimp + np.arange(env.size))] = vec[:, order]
        assert np.linalg.norm(loc2dmet.T @ loc2dmet - np.eye(ntot)) < 1e-12
        return (numBathOrbs, loc2dmet, coreOccupations)

    def constructbaths(self, umat_loc, impurityOrbsList, numBathOrbsList=None, threshold=1e-13, cache_key=None,
                       want_one_rdm=True):
        """Baths of ALL fragments in one device call, from the orbitals of Fock + umat -- the
        same eigensolve the residual and the Jacobian at this umat use.  Replaces the loop over
        constructbath(construct1RDM_loc(doSCF, umat), impurityOrbs, numBathOrbs) of src/dmet.py:194-204, whose
        numpy eigh of the (N - m) x (N - m) environment block per fragment becomes one m x m eigenproblem per
        fragment and one DGEMM (kernels_bath.cu).

        Returns (OneRDM or None, [(numBathOrbs, loc2dmet_active, bathOccupations), ...]) with
        loc2dmet_active = loc2dmet[:, :numImpOrbs + numBathOrbs] of the reference (impurity | bath columns; the
        pure-environment columns are never formed) and bathOccupations the environment occupations of ALL
        m candidates in rank order (closest to 1 first, src/qcdmethelper.py:139-140)."""
        imps = [np.flatnonzero(np.abs(np.asarray(v).reshape(-1)).astype(int)) for v in impurityOrbsList]
        imp_start = np.concatenate([[0], np.cumsum([i.size for i in imps])]).astype(np.int32)
        imp_idx = (np.concatenate(imps) if imps else np.zeros(0)).astype(np.int32)
        # any plan of this helper has the orbitals: use (and create if needed) the impurity-row plan
        plan = self._plan_with_orbitals(umat_loc, imp_start, None, imp_idx, None, cache_key)
        req = None if numBathOrbsList is None else np.asarray(numBathOrbsList, dtype=np.int32)
        num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx, req, threshold)
        n = self.locints.Norbs
        out = []
        for f, imp in enumerate(imps):
            nb = int(num_bath[f])
            active = np.zeros((n, imp.size + nb))
            active[imp, np.arange(imp.size)] = 1.0
            active[:, imp.size:] = bath[imp_start[f]:imp_start[f] + nb].T
            out.append((nb, active, occ[imp_start[f]:imp_start[f + 1]].copy()))
        return (plan.one_rdm() if want_one_rdm else None), out

    def embed_hubbard(self, umat_loc, impurityOrbsList, numBathOrbsList=None, threshold=1e-13, cache_key=None):
        """constructbaths + the embedded Hamiltonians of every fragment in one device call, for integral providers
        with an on-site interaction (localintegrals_hubbard: attributes activeOEI, HubbardU).  Replaces, per
        fragment, constructbath and dmet_oei / dmet_fock / dmet_tei (src/localintegrals_hubbard.py:88-116).
        Returns the dict of _lib.Plan.embed_hubbard plus 'active': loc2dmet[:, :n] per fragment."""
        imps = [np.flatnonzero(np.abs(np.asarray(v).reshape(-1)).astype(int)) for v in impurityOrbsList]
        imp_start = np.concatenate([[0], np.cumsum([i.size for i in imps])]).astype(np.int32)
        imp_idx = np.concatenate(imps).astype(np.int32)
        plan = self._plan_with_orbitals(umat_loc, imp_start, None, imp_idx, None, cache_key)
        req = None if numBathOrbsList is None else np.asarray(numBathOrbsList, dtype=np.int32)
        out = plan.embed_hubbard(self.locints.activeOEI, self.locints.HubbardU, imp_start, imp_idx, req, threshold)
        n = self.locints.Norbs
        out["active"] = []
        for f, imp in enumerate(imps):
            nb = int(out["num_bath"][f])
            active = np.zeros((n, imp.size + nb))
            active[imp, np.arange(imp.size)] = 1.0
            active[:, imp.size:] = out["bath"][imp_start[f]:imp_start[f] + nb].T
            out["active"].append(active)
        out["imp_start"] = imp_start
        return out

    def _oei(self, doSCF, umat_loc, NOrotation):
        if doSCF:
            raise NotImplementedError("doSCF=True is outside the response hot path (src/dmet.py:44)")
        OEI = self.locints.loc_rhf_fock() + umat_loc          # src/qcdmethelper.py:87
        if NOrotation is not None:                            # src/qcdmethelper.py:98-99
            OEI = np.dot(np.dot(NOrotation.T, OEI), NOrotation)
        return np.ascontiguousarray(OEI, dtype=ctypes.c_double)

    # ---------------------------------------------------------------- src/qcdmethelper.py:84-112
    def construct1RDM_response(self, doSCF, umat_loc, NOrotation):
        """Full response, shape (Nterms, Norbs, Norbs) -- only sensible when 8 N^2 P bytes fit
        in host memory; dmet.rdm_differences_derivative uses response_jacobian instead."""
        N = self.locints.Norbs
        OEI = self._oei(doSCF, umat_loc, NOrotation).reshape(N * N)
        rdm_deriv_rot = np.ones([N * N * self.Nterms], dtype=ctypes.c_double)
        lib = _lib.load()
        h1row = self.H1row if self.H1row.size else np.zeros(1, dtype=ctypes.c_int)
        h1col = self.H1col if self.H1col.size else np.zeros(1, dtype=ctypes.c_int)
        rc = lib.rhf_response(ctypes.c_int(N), ctypes.c_int(self.Nterms), ctypes.c_int(self.numPairs),
                              self.H1start.ctypes.data_as(ctypes.c_void_p),
                              h1row.ctypes.data_as(ctypes.c_void_p),
                              h1col.ctypes.data_as(ctypes.c_void_p),
                              OEI.ctypes.data_as(ctypes.c_void_p),
                              rdm_deriv_rot.ctypes.data_as(ctypes.c_void_p))
        _lib.check(rc)
        return rdm_deriv_rot.reshape((self.Nterms, N, N), order='C')

    def construct1RDM_loc_response(self, doSCF, umat, list_H1=None):
        """README-era signature (README.md:57); list_H1 now lives in the constructor."""
        if list_H1 is not None and list_H1 is not self.list_H1:
            other = qcdmethelper(self.locints, list_H1, self.altcf, self.minFunc)
            return other.construct1RDM_response(doSCF, umat, None)
        return self.construct1RDM_response(doSCF, umat, None)

    # ---------------------------------------------------------------- fused path
    def _plan(self, row_start, rows, row_idx):
        key = (np.asarray(row_start).tobytes(),
               None if rows is None else (rows.shape, hash(rows.tobytes())),
               None if row_idx is None else np.asarray(row_idx).tobytes())
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 4:
                for old in self._plans.values():
                    old.close()
                self._plans.clear()
            plan = _lib.Plan(self.locints.Norbs, self.numPairs, self.H1start, self.H1row, self.H1col,
                             row_start, rows, row_idx)
            self._plans[key] = plan
        return plan

    def _plan_with_orbitals(self, umat_loc, row_start, rows, row_idx, NOrotation, cache_key=None):
        """Plan whose device orbitals are those of Fock + umat.  The Levenberg-Marquardt driver evaluates
        the residual and the Jacobian at the same umat (src/dmet.py:631): the eigensolve (one cuSOLVER
        syevd) is shared between the two calls instead of being repeated as in the reference
        (numpy eigh in construct1RDM_base + dsyev in rhf_response)."""
        plan = self._plan(row_start, rows, row_idx)
        if cache_key is not None:
            # cheap identity of (Fock + umat): e.g. the bytes of the flat parameter vector (dmet passes it), so
            # that large problems neither compare nor copy an N x N matrix per call
            key = ("key", cache_key, NOrotation is not None)
            cached = getattr(plan, "_oei_cache", None)
            if isinstance(cached, tuple) and cached == key:
                return plan
            oei = self._oei(False, umat_loc, NOrotation)
        else:
            oei = self._oei(False, umat_loc, NOrotation)
            cached = getattr(plan, "_oei_cache", None)
            if isinstance(cached, np.ndarray) and np.array_equal(cached, oei):
                return plan
            key = np.array(oei, copy=True)
        plan._oei_cache = None              # a failed set_h0 leaves no orbitals behind: never reuse the old key
        plan.set_h0(oei)
        plan._oei_cache = key
        self.num_eigensolves = getattr(self, "num_eigensolves", 0) + 1
        return plan

    def response_jacobian(self, umat_loc, row_start, rows=None, row_idx=None, NOrotation=None,
                          want_dblocks=False, cache_key=None):
        """Fragment blocks of every dD_p, written directly as Jacobian columns.

        Replaces construct1RDM_response + the assembly loop of src/dmet.py:496-520.
        Returns a Fortran-ordered (Q, Nterms) float64 array (and the mean-field 1-RDM blocks)."""
        plan = self._plan_with_orbitals(umat_loc, row_start, rows, row_idx, NOrotation, cache_key)
        plan.compute()
        # reuse_output: the plan's page-locked buffer (full PCIe rate, no 8 Q P byte page-touch per call); the
        # returned array is then overwritten by the next call -- scipy's leastsq / lmder copies it at once
        jac = plan.jacobian(pinned=self.reuse_output)
        if want_dblocks:
            return jac, plan.dblocks()
        return jac

    def response_jacobian_device(self, umat_loc, row_start, rows=None, row_idx=None, NOrotation=None, cache_key=None):
        """Compute the Jacobian and leave it on the device (for Plan.lm_step); returns the plan."""
        plan = self._plan_with_orbitals(umat_loc, row_start, rows, row_idx, NOrotation, cache_key)
        plan.compute()
        return plan

    def construct1RDM_blocks(self, umat_loc, row_start, rows=None, row_idx=None, NOrotation=None, cache_key=None):
        """Fragment blocks R_f (2 C_o C_o^T) R_f^T of the mean-field 1-RDM, flattened like the residual
        (column-major per fragment): the model part of dmet.rdm_differences (src/dmet.py:416-421,432),
        from the device eigensolve shared with response_jacobian."""
        return self._plan_with_orbitals(umat_loc, row_start, rows, row_idx, NOrotation, cache_key).dblocks()
