"""The umat-fit side of QC-DMET's `dmet` class, with the reference's names and signatures.

Mirrors src/dmet.py for everything on the correlation-potential (umat) fit path:

    __init__ / testclusters / make_imp_size          src/dmet.py:28-116
    makelist_H1 (CSR form) / make_mask               src/dmet.py:118-177
    flat2square / square2flat                        src/dmet.py:560-587
    rdm_differences                                  src/dmet.py:395-439
    rdm_differences_derivative                       src/dmet.py:477-525   <- fused GPU Jacobian
    costfunction / costfunction_derivative           src/dmet.py:361-363,381-388
    verify_gradient / hessian_eigenvalues            src/dmet.py:527-558
    fit_umat (the SCmethod dispatch of doselfconsistent, src/dmet.py:626-636)

The impurity solvers (PyCheMPS2 / pyscf wrappers, src/chemps2.py, src/pyscf_*.py) and the
integral providers are outside this hot path: `doexact` accepts any callable with
the reference's solver signature, and `imp_1RDM` / `dmetOrbs` can be set directly.
"""
from __future__ import annotations

import time

import numpy as np

from . import qcdmethelper


class dmet:

    def __init__(self, theInts, impurityClusters, isTranslationInvariant, method='ED', SCmethod='LSTSQ',
                 fitImpBath=True, use_constrained_opt=False, solver=None):
        if isTranslationInvariant:
            assert theInts.TI_OK
        assert method in ('ED', 'CC', 'MP2')
        assert SCmethod in ('LSTSQ', 'BFGS', 'NONE')

        self.ints = theInts
        self.Norb = self.ints.Norbs
        self.impClust = impurityClusters
        self.umat = np.zeros([self.Norb, self.Norb], dtype=float)
        self.relaxation = 0.0

        self.NI_hack = False
        self.method = method
        self.doSCF = False
        self.TransInv = isTranslationInvariant
        self.SCmethod = SCmethod
        self.CC_E_TYPE = 'LAMBDA'
        self.BATH_ORBS = None
        self.fitImpBath = fitImpBath
        self.doDET = False
        self.doDET_NO = False
        self.NOrotation = None
        self.altcostfunc = use_constrained_opt
        self.minFunc = None
        if self.altcostfunc:
            self.minFunc = 'FOCK_INIT'
            assert self.fitImpBath is False
            assert self.SCmethod in ('BFGS', 'NONE')
        self.solver = solver

        self.print_u = True
        self.print_rdm = True

        allOne = self.testclusters()
        if not allOne:
            assert self.TransInv is False

        self.energy = 0.0
        self.imp_1RDM = []
        self.dmetOrbs = []
        self.imp_size = self.make_imp_size()
        self.mu_imp = 0.0
        self.mask = self.make_mask()
        self.helper = qcdmethelper.qcdmethelper(self.ints, self.makelist_H1(), self.altcostfunc, self.minFunc)

        self.time_ed = 0.0
        self.time_cf = 0.0
        self.time_func = 0.0
        self.time_grad = 0.0

    # ------------------------------------------------------------------ src/dmet.py:95-116
    def testclusters(self):
        quicktest = np.zeros([self.Norb], dtype=int)
        for item in self.impClust:
            quicktest += np.abs(item)
        assert np.all(quicktest >= 0)
        assert np.all(quicktest <= 1)
        return bool(np.all(quicktest == 1))

    def make_imp_size(self):
        maxiter = 1 if self.TransInv else len(self.impClust)
        return np.array([int(np.sum(np.abs(self.impClust[c]))) for c in range(maxiter)])

    # ------------------------------------------------------------------ src/dmet.py:118-159
    def makelist_H1(self):
        """Perturbation list in CSR form (H1start, H1row, H1col).

        Same parameter order as the reference's dense list (fragment by fragment, row <= col inside
        the diagonal block; TI: the element replicated in every cell) and the same element order
        inside one perturbation as np.where on the dense matrix (src/qcdmethelper.py:54).  The
        dense N x N matrices themselves (8 N^2 P bytes) are never built."""
        H1start, H1row, H1col = [0], [], []

        def emit(pairs):
            for r, c in sorted(set(pairs)):
                H1row.append(r)
                H1col.append(c)
            H1start.append(len(H1row))

        if self.TransInv:
            localsize = int(self.imp_size[0])
            for row in range(localsize):
                for col in ([row] if self.doDET else range(row, localsize)):
                    pairs = []
                    for jumper in range(self.Norb // localsize):
                        js = localsize * jumper
                        pairs += [(js + row, js + col), (js + col, js + row)]
                    emit(pairs)
        else:
            jumpsquare = 0
            for localsize in self.imp_size:
                localsize = int(localsize)
                for row in range(localsize):
                    for col in ([row] if self.doDET else range(row, localsize)):
                        emit([(jumpsquare + row, jumpsquare + col), (jumpsquare + col, jumpsquare + row)])
                jumpsquare += localsize
        return (np.array(H1start, dtype=np.int32), np.array(H1row, dtype=np.int32),
                np.array(H1col, dtype=np.int32))

    # ------------------------------------------------------------------ src/dmet.py:161-177
    def make_mask(self):
        themask = np.zeros([self.Norb, self.Norb], dtype=bool)
        jump = 0
        for localsize in self.imp_size:
            localsize = int(localsize)
            for row in range(localsize):
                for col in ([row] if self.doDET else range(row, localsize)):
                    themask[jump + row, jump + col] = True
            jump += localsize
        return themask

    # ------------------------------------------------------------------ src/dmet.py:560-587
    def flat2square(self, umatflat):
        umatsquare = np.zeros([self.Norb, self.Norb], dtype=float)
        umatsquare[self.mask] = umatflat
        umatsquare = umatsquare.T.copy()
        umatsquare[self.mask] = umatflat
        if self.TransInv:
            size = int(self.imp_size[0])
            for it in range(1, self.Norb // size):
                umatsquare[it * size:(it + 1) * size, it * size:(it + 1) * size] = umatsquare[0:size, 0:size]
        if self.NOrotation is not None:
            umatsquare = np.dot(np.dot(self.NOrotation, umatsquare), self.NOrotation.T)
        return umatsquare

    def square2flat(self, umatsquare):
        umatsquare_bis = np.array(umatsquare, copy=True)
        if self.NOrotation is not None:
            umatsquare_bis = np.dot(np.dot(self.NOrotation.T, umatsquare_bis), self.NOrotation)
        return umatsquare_bis[self.mask]

    # ------------------------------------------------------------------ fragment rows
    def _fragment_rows(self):
        """Residual rows per fragment: dmetOrbs[f]^T (fitImpBath, src/dmet.py:502) or the impurity
        site indices flatnonzero(impClust[f]) (src/dmet.py:509)."""
        nfit = len(self.imp_size)
        if self.fitImpBath:
            assert len(self.dmetOrbs) >= nfit, "fitImpBath needs dmetOrbs (run doexact first)"
            blocks = [np.ascontiguousarray(np.asarray(self.dmetOrbs[f], dtype=float).T) for f in range(nfit)]
            sizes = [b.shape[0] for b in blocks]
            rows = np.ascontiguousarray(np.concatenate(blocks, axis=0))
            return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), rows, None
        idx = [np.flatnonzero(self.impClust[f]) for f in range(nfit)]
        row_idx = np.concatenate(idx).astype(np.int32)
        row_start = np.concatenate([[0], np.cumsum([len(i) for i in idx])]).astype(np.int32)
        return row_start, None, row_idx

    # ------------------------------------------------------------------ src/dmet.py:395-439
    def rdm_differences(self, newumatflat):
        start_func = time.time()
        newumatsquare_loc = self.flat2square(newumatflat)
        OneRDM_loc = self.helper.construct1RDM_loc(self.doSCF, newumatsquare_loc)
        pieces = []
        for count in range(len(self.imp_size)):
            if self.fitImpBath:
                mf_1RDM = np.dot(np.dot(self.dmetOrbs[count].T, OneRDM_loc), self.dmetOrbs[count])
                ed_1RDM = self.imp_1RDM[count]
            else:
                idx = np.flatnonzero(self.impClust[count])
                mf_1RDM = (OneRDM_loc[:, idx])[idx, :]
                m = int(self.imp_size[count])
                ed_1RDM = self.imp_1RDM[count][:m, :m]
            theerror = mf_1RDM - ed_1RDM
            pieces.append(np.reshape(theerror, theerror.size, order='F'))
        errors = np.concatenate(pieces)
        self.time_func += time.time() - start_func
        return errors

    # ------------------------------------------------------------------ src/dmet.py:477-525
    def rdm_differences_derivative(self, newumatflat):
        """(Q, P) Jacobian of rdm_differences.  One fused GPU call instead of the N x N x P
        response + P*F Python slicing iterations of the reference."""
        start_grad = time.time()
        newumatsquare_loc = self.flat2square(newumatflat)
        row_start, rows, row_idx = self._fragment_rows()
        gradient = self.helper.response_jacobian(newumatsquare_loc, row_start, rows=rows, row_idx=row_idx,
                                                 NOrotation=self.NOrotation)
        self.time_grad += time.time() - start_grad
        return gradient

    # ------------------------------------------------------------------ src/dmet.py:361-363,381-388
    def costfunction(self, newumatflat):
        return np.linalg.norm(self.rdm_differences(newumatflat)) ** 2

    def costfunction_derivative(self, newumatflat):
        errors = self.rdm_differences(newumatflat)
        error_derivs = self.rdm_differences_derivative(newumatflat)
        return 2.0 * np.dot(error_derivs.T, errors)

    # ------------------------------------------------------------------ src/dmet.py:527-539
    def verify_gradient(self, umatflat):
        gradient = self.costfunction_derivative(umatflat)
        cost_reference = self.costfunction(umatflat)
        gradientbis = np.zeros([len(gradient)], dtype=float)
        stepsize = 1e-7
        for cnt in range(len(gradient)):
            umatbis = np.array(umatflat, copy=True)
            umatbis[cnt] += stepsize
            gradientbis[cnt] = (self.costfunction(umatbis) - cost_reference) / stepsize
        return np.linalg.norm(gradient - gradientbis), np.linalg.norm(gradient)

    # ------------------------------------------------------------------ src/dmet.py:541-558
    def hessian_eigenvalues(self, umatflat):
        stepsize = 1e-7
        gradient_reference = self.costfunction_derivative(umatflat)
        hessian = np.zeros([len(umatflat), len(umatflat)], dtype=float)
        for cnt in range(len(umatflat)):
            gradient = np.array(umatflat, copy=True)
            gradient[cnt] += stepsize
            gradient = self.costfunction_derivative(gradient)
            hessian[:, cnt] = (gradient - gradient_reference) / stepsize
        hessian = 0.5 * (hessian + hessian.T)
        return np.sort(np.linalg.eigvalsh(hessian))

    # ------------------------------------------------------------------ src/dmet.py:626-636
    def fit_umat(self):
        """The umat solve of one DMET iteration (the SCmethod dispatch of doselfconsistent)."""
        from scipy import optimize
        start_cf = time.time()
        if self.SCmethod == 'LSTSQ':
            result = optimize.leastsq(self.rdm_differences, self.square2flat(self.umat),
                                      Dfun=self.rdm_differences_derivative, factor=0.1)
            self.umat = self.flat2square(result[0])
        elif self.SCmethod == 'BFGS':
            result = optimize.minimize(self.costfunction, self.square2flat(self.umat),
                                       jac=self.costfunction_derivative, options={'disp': False})
            self.umat = self.flat2square(result.x)
        # Remove arbitrary chemical potential shifts (src/dmet.py:636)
        self.umat = self.umat - np.eye(self.umat.shape[0]) * np.average(np.diag(self.umat))
        self.time_cf += time.time() - start_cf
        return self.umat

    def transform_ed_1rdm(self):
        result = np.zeros([self.umat.shape[0], self.umat.shape[0]], dtype=float)
        squarejumper = 0
        for count in range(len(self.imp_1RDM)):
            localsize = int(self.imp_size[count])
            result[squarejumper:squarejumper + localsize, squarejumper:squarejumper + localsize] = \
                self.imp_1RDM[count][:localsize, :localsize]
            squarejumper += localsize
        return result

    def onedm_solution_rhf(self):
        return self.helper.construct1RDM_loc(self.doSCF, self.umat)
