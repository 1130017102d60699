"""The umat-fit side of QC-DMET's `dmet` class, with the reference's names and signatures.

Mirrors src/dmet.py for everything on the correlation-potential (umat) fit path:

    __init__ / testclusters / make_imp_size          src/dmet.py:28-116
    makelist_H1 (CSR form) / make_mask               src/dmet.py:118-177
    flat2square / square2flat                        src/dmet.py:560-587
    rdm_differences                                  src/dmet.py:395-439
    rdm_differences_derivative                       src/dmet.py:477-525   <- fused GPU Jacobian
    costfunction / costfunction_derivative           src/dmet.py:361-363,381-388
    verify_gradient / hessian_eigenvalues            src/dmet.py:527-558
    constructNOrotation                              src/dmet.py:343-359
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
        self.bath_on_gpu = False          # doexact: baths of all fragments from one device call (helper.constructbaths)
        self.embed_on_gpu = False         # ... and, for on-site (Hubbard) integrals, the embedded Hamiltonians too
        self._bath_cache = None
        self.solver_threads = 1            # > 1: the embedded problems of one doexact pass are solved in a thread pool
        self.gpu_residual = True      # residual blocks from the device eigensolve shared with the Jacobian
        self.fit_on_gpu = False       # LSTSQ: Levenberg-Marquardt with every O(Q P^2) step on the device
        self.lm_method = 0            # _lib.LM_NORMAL (normal equations + Cholesky) or _lib.LM_QR (QR of [J; sqrt(lambda) D])

        self.print_u = True
        self.print_rdm = True
        self.print_hessian = False    # src/dmet.py:622-623: Hessian spectrum of the cost function per DMET iteration
        self.last_hessian_eigenvalues = None
        self.NOvecs = None

        allOne = self.testclusters()
        if not allOne:
            assert self.TransInv is False

        self.energy = 0.0
        self.imp_1RDM = []
        self.dmetOrbs = []
        self.imp_size = self.make_imp_size()
        self.mu_imp = 0.0
        self.mask = self.make_mask()
        self._mask_rc = None
        self._imp_rows = None
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

    # ------------------------------------------------------------------ src/dmet.py:179-341
    def doexact(self, chempot_imp=0.0):
        """One pass over the fragments: bath construction, embedded Hamiltonian, impurity solver, energy and
        electron count.  Host orchestration around third-party solvers in the reference ("next" row f-1 of
        SURVEY.md 8); the solver is any callable with the signature of src/chemps2.py:27 (default: the
        dense-FCI stand-in qcdmet_b200.fci_solver.solve).  The DET / CASCI / incomplete-tiling branches
        (src/dmet.py:259-337) need pyscf and are not restated."""
        maxiter = 1 if self.TransInv else len(self.impClust)
        device_baths = device_emb = None
        if self.embed_on_gpu:
            # baths + embedded one- and two-electron integrals of every fragment from one device call
            assert hasattr(self.ints, "HubbardU"), "embed_on_gpu needs an on-site (Hubbard) integral provider"
            key = np.asarray(self.umat).tobytes()
            if self._bath_cache is None or self._bath_cache[0] != key:
                req = None if self.BATH_ORBS is None else [self.BATH_ORBS[c] for c in range(maxiter)]
                self._bath_cache = (key, self.helper.embed_hubbard(self.umat, self.impClust[:maxiter], req,
                                                                   cache_key=self._umat_key(self.square2flat(self.umat))))
            device_emb = self._bath_cache[1]
        elif self.bath_on_gpu:
            # one eigensolve (shared with the residual / Jacobian at this umat) and one batched bath call per umat;
            # the Newton iteration on the chemical potential re-enters doexact with the same umat
            key = np.asarray(self.umat).tobytes()
            if self._bath_cache is None or self._bath_cache[0] != key:
                req = None if self.BATH_ORBS is None else [self.BATH_ORBS[c] for c in range(maxiter)]
                self._bath_cache = (key, self.helper.constructbaths(self.umat, self.impClust[:maxiter], req,
                                                                    cache_key=self._umat_key(self.square2flat(self.umat))))
            OneRDM, device_baths = self._bath_cache[1]
        elif not self.embed_on_gpu:
            OneRDM = self.helper.construct1RDM_loc(self.doSCF, self.umat)
        self.energy = 0.0
        self.imp_1RDM = []
        self.dmetOrbs = []
        solver = self.solver
        if solver is None:
            from . import fci_solver
            solver = fci_solver.solve
        remainingOrbs = np.ones([len(self.impClust[0])], dtype=float)
        problems = []
        for counter in range(maxiter):
            impurityOrbs = np.abs(self.impClust[counter])
            numImpOrbs = int(np.sum(impurityOrbs))
            numBathOrbs = numImpOrbs if self.BATH_ORBS is None else self.BATH_ORBS[counter]
            core_cutoff = 0.01 if self.BATH_ORBS is None else 0.5
            if device_emb is not None:
                numBathOrbs = int(device_emb["num_bath"][counter])
                r0, r1 = device_emb["imp_start"][counter], device_emb["imp_start"][counter + 1]
                left = device_emb["occ"][r0 + numBathOrbs:r1]
                if np.any((left > core_cutoff) & (left < 2.0 - core_cutoff)):
                    raise AssertionError("Bad DMET bath orbital selection: bath orbital with occupation "
                                         f"{left} would go into the environment")
                Norb_in_imp = numImpOrbs + numBathOrbs
                self.dmetOrbs.append(device_emb["active"][counter])
                Nelec_in_imp = int(round(self.ints.Nelec - device_emb["core_electrons"][counter]))
                problems.append((0.0, device_emb["oei"][counter], device_emb["fock"][counter], device_emb["tei"][counter],
                                 Norb_in_imp, Nelec_in_imp, numImpOrbs, chempot_imp))
                remainingOrbs -= impurityOrbs
                continue
            if device_baths is not None:
                # loc2dmet[:, :imp + bath] and the core density without the pure-environment orbitals: for an
                # idempotent D/2, D_env,env = sum_bath occ_k b_k b_k^T + (core 1-RDM)
                numBathOrbs, loc2dmet, cand_occ = device_baths[counter]
                left = cand_occ[numBathOrbs:]
                if np.any((left > core_cutoff) & (left < 2.0 - core_cutoff)):
                    raise AssertionError("Bad DMET bath orbital selection: bath orbital with occupation "
                                         f"{left} would go into the environment")
                env = np.flatnonzero(impurityOrbs == 0)
                bath = loc2dmet[:, numImpOrbs:]
                core1RDM_loc = np.zeros_like(OneRDM)
                core1RDM_loc[np.ix_(env, env)] = (OneRDM - (bath * cand_occ[:numBathOrbs]) @ bath.T)[np.ix_(env, env)]
                core_electrons = float(np.trace(core1RDM_loc))
            else:
                numBathOrbs, loc2dmet, core1RDM_dmet = self.helper.constructbath(OneRDM, impurityOrbs, numBathOrbs)
                for cnt in range(len(core1RDM_dmet)):
                    if core1RDM_dmet[cnt] < core_cutoff:
                        core1RDM_dmet[cnt] = 0.0
                    elif core1RDM_dmet[cnt] > 2.0 - core_cutoff:
                        core1RDM_dmet[cnt] = 2.0
                    else:
                        raise AssertionError("Bad DMET bath orbital selection: bath orbital with occupation "
                                             f"{core1RDM_dmet[cnt]} would go into the environment")
                core1RDM_loc = np.dot(np.dot(loc2dmet, np.diag(core1RDM_dmet)), loc2dmet.T)
                core_electrons = float(np.sum(core1RDM_dmet))
            Norb_in_imp = numImpOrbs + numBathOrbs
            Nelec_in_imp = int(round(self.ints.Nelec - core_electrons))
            self.dmetOrbs.append(loc2dmet[:, :Norb_in_imp])
            assert Norb_in_imp <= self.Norb
            dmetOEI = self.ints.dmet_oei(loc2dmet, Norb_in_imp)
            dmetFOCK = self.ints.dmet_fock(loc2dmet, Norb_in_imp, core1RDM_loc)
            dmetTEI = self.ints.dmet_tei(loc2dmet, Norb_in_imp)
            problems.append((0.0, dmetOEI, dmetFOCK, dmetTEI, Norb_in_imp, Nelec_in_imp, numImpOrbs, chempot_imp))
            remainingOrbs -= impurityOrbs
        # the embedded problems are independent: solve them side by side (results are collected in fragment
        # order, so the energy sum does not depend on the number of threads)
        if self.solver_threads > 1 and len(problems) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=self.solver_threads) as pool:
                solutions = list(pool.map(lambda args: solver(*args), problems))
        else:
            solutions = [solver(*args) for args in problems]
        for IMP_energy, IMP_1RDM in solutions:
            self.energy += IMP_energy
            self.imp_1RDM.append(IMP_1RDM)
        Nelectrons = 0.0
        for counter in range(maxiter):
            m = int(self.imp_size[counter])
            Nelectrons += np.trace(self.imp_1RDM[counter][:m, :m])
        if self.TransInv:
            Nelectrons = Nelectrons * len(self.impClust)
            self.energy = self.energy * len(self.impClust)
            remainingOrbs[:] = 0
        assert np.all(remainingOrbs == 0), "incomplete impurity tilings need the reference's pyscf branch"
        self.energy += self.ints.const()
        return Nelectrons

    # ------------------------------------------------------------------ src/dmet.py:589-666
    def numeleccostfunction(self, chempot_imp):
        return self.doexact(chempot_imp) - self.ints.Nelec

    def doselfconsistent(self, max_iterations=50, verbose=False):
        from scipy import optimize
        iteration = 0
        u_diff = 1.0
        convergence_threshold = 1e-5
        while u_diff > convergence_threshold and iteration < max_iterations:
            iteration += 1
            umat_old = np.array(self.umat, copy=True)
            start_ed = time.time()
            self.mu_imp = optimize.newton(self.numeleccostfunction, self.mu_imp)
            self.time_ed += time.time() - start_ed
            # src/dmet.py:622-623 runs hessian_eigenvalues here on every iteration (P + 1 Jacobians for a
            # debug print); kept behind a flag, off by default
            if self.print_hessian and self.SCmethod != 'NONE' and not self.altcostfunc:
                self.last_hessian_eigenvalues = self.hessian_eigenvalues(self.square2flat(self.umat))
                if verbose:
                    print("   Hessian eigenvalues (min, max) =", self.last_hessian_eigenvalues[[0, -1]])
            if self.SCmethod != 'NONE':
                self.fit_umat()
            u_diff = np.linalg.norm(umat_old - self.umat)
            self.umat = self.relaxation * umat_old + (1.0 - self.relaxation) * self.umat
            if verbose:
                print(f"DMET iteration {iteration}: energy = {self.energy:.10f}  mu = {self.mu_imp:.8f}  "
                      f"|du| = {u_diff:.3e}")
            if self.SCmethod == 'NONE':
                u_diff = 0.1 * convergence_threshold
        self.iterations = iteration
        return self.energy

    # ------------------------------------------------------------------ src/dmet.py:560-587
    def flat2square(self, umatflat):
        """umat = sum_p x_p H1_p: parameter p sits at the masked (row <= col) position p of the row-major
        scan and at its mirror image; TI repeats cell 0 along the diagonal; an optional natural-orbital
        rotation R maps the parameter basis to the lattice basis (R u R^T)."""
        x = np.asarray(umatflat, dtype=float)
        if self._mask_rc is None:
            self._mask_rc = np.nonzero(self.mask)       # row-major order == the order of makelist_H1
        rows, cols = self._mask_rc
        u = np.zeros((self.Norb, self.Norb))
        u[rows, cols] = x
        u[cols, rows] = x
        if self.TransInv:
            m = int(self.imp_size[0])
            ncell = self.Norb // m
            u = np.kron(np.eye(ncell), u[:m, :m])
            if ncell * m < self.Norb:
                full = np.zeros((self.Norb, self.Norb))
                full[:ncell * m, :ncell * m] = u
                u = full
        if self.NOrotation is not None:
            u = self.NOrotation @ u @ self.NOrotation.T
        return u

    def square2flat(self, umatsquare):
        """Inverse of flat2square on its range: the masked upper-triangle entries in the parameter basis."""
        u = np.asarray(umatsquare, dtype=float)
        if self.NOrotation is not None:
            u = self.NOrotation.T @ u @ self.NOrotation
        return np.array(u[self.mask], copy=True)

    # ------------------------------------------------------------------ src/dmet.py:343-359
    def constructNOrotation(self, NOvecs=None):
        """Block-diagonal orthogonal matrix holding the natural-orbital vectors of every fitted impurity
        (identity on orbitals outside all impurities; TI: the first block repeated in every cell)."""
        vecs = self.NOvecs if NOvecs is None else NOvecs
        rot = np.eye(self.Norb)
        pos = 0
        for count, m in enumerate(int(x) for x in self.imp_size):
            rot[pos:pos + m, pos:pos + m] = np.asarray(vecs[count], dtype=float)
            pos += m
        if self.TransInv:
            m = int(self.imp_size[0])
            for cell in range(1, self.Norb // m):
                rot[cell * m:(cell + 1) * m, cell * m:(cell + 1) * m] = rot[:m, :m]
        return rot

    # ------------------------------------------------------------------ fragment rows
    def _umat_key(self, umatflat):
        """Identity of Fock + umat(x) for the helper's eigensolve cache: the parameter vector itself (the Fock
        matrix and the NO rotation are fixed while a fit runs; dmetOrbs changes select another plan)."""
        nor = None if self.NOrotation is None else hash(np.asarray(self.NOrotation).tobytes())
        return (np.asarray(umatflat, dtype=float).tobytes(), nor, id(self.ints))

    def _fragment_rows(self):
        """Residual rows per fragment: dmetOrbs[f]^T (fitImpBath, src/dmet.py:502) or the impurity
        site indices flatnonzero(impClust[f]) (src/dmet.py:509)."""
        nfit = len(self.imp_size)
        if not self.fitImpBath:
            if self._imp_rows is None:
                idx = [np.flatnonzero(self.impClust[f]) for f in range(nfit)]
                self._imp_rows = (np.concatenate([[0], np.cumsum([len(i) for i in idx])]).astype(np.int32), None,
                                  np.concatenate(idx).astype(np.int32))
            return self._imp_rows
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
        """Residual of the fit.  With gpu_residual (default) the mean-field blocks come from the device
        eigensolve that the Jacobian at the same umat reuses; otherwise the reference's
        host route (numpy eigh of the full matrix, src/qcdmethelper.py:114-122) is followed line by line."""
        if self.gpu_residual and not self.altcostfunc:
            # E_GAP: unlike numpy's eigh on the host route, the device path refuses a vanishing HOMO-LUMO gap
            # (the response at that umat is undefined anyway, lib/rhf_response.cc:66-71)
            start_func = time.time()
            row_start, rows, row_idx = self._fragment_rows()
            # lattice basis, as src/dmet.py:399-421 (flat2square already carries the NO rotation of umat); the
            # Jacobian call below passes NOrotation, so the two only share the eigensolve when it is None
            mf = self.helper.construct1RDM_blocks(self.flat2square(newumatflat), row_start, rows=rows,
                                                  row_idx=row_idx, NOrotation=None, cache_key=self._umat_key(newumatflat))
            pieces = []
            for count in range(len(self.imp_size)):
                if self.fitImpBath:
                    ed_1RDM = self.imp_1RDM[count]
                else:
                    m = int(self.imp_size[count])
                    ed_1RDM = self.imp_1RDM[count][:m, :m]
                pieces.append(np.reshape(ed_1RDM, ed_1RDM.size, order='F'))
            errors = mf - np.concatenate(pieces)
            self.time_func += time.time() - start_func
            return errors
        return self.rdm_differences_host(newumatflat)

    def rdm_differences_host(self, newumatflat):
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

    # ------------------------------------------------------------------ src/dmet.py:441-475
    def rdm_differences_bis(self, newumatflat):
        """Masked (upper-triangle) impurity-block residual of the constrained-optimisation variant
        (use_constrained_opt=True).  Host numpy as in the reference: this variant never calls the response."""
        start_func = time.time()
        newumatsquare_loc = self.flat2square(newumatflat)
        OneRDM_loc = self.helper.construct1RDM_loc(self.doSCF, newumatsquare_loc)
        pieces = []
        jumpc = 0
        for count in range(len(self.imp_size)):
            m = int(self.imp_size[count])
            idx = np.flatnonzero(self.impClust[count])
            theerror = (OneRDM_loc[:, idx])[idx, :] - self.imp_1RDM[count][:m, :m]
            mask_t = self.mask[jumpc:jumpc + m, jumpc:jumpc + m]
            pieces.append(theerror[mask_t])          # boolean-mask order (row-major), as src/dmet.py:470
            jumpc += m
        self.time_func += time.time() - start_func
        return np.concatenate(pieces)

    # ------------------------------------------------------------------ src/dmet.py:365-379,390-393
    def alt_costfunction(self, newumatflat):
        newumatsquare_loc = self.flat2square(newumatflat)
        OneRDM_loc = self.helper.construct1RDM_loc(self.doSCF, newumatsquare_loc)
        errors_sq = self.flat2square(self.rdm_differences_bis(newumatflat))
        base = self.ints.loc_oei() if self.minFunc == 'OEI' else self.ints.loc_rhf_fock()
        e_fun = np.trace(np.dot(base, OneRDM_loc))
        e_cstr = np.sum(newumatsquare_loc * errors_sq)
        return -e_fun - e_cstr

    def alt_costfunction_derivative(self, newumatflat):
        return -self.rdm_differences_bis(newumatflat)

    # ------------------------------------------------------------------ src/dmet.py:477-525
    def rdm_differences_derivative(self, newumatflat):
        """(Q, P) Jacobian of rdm_differences.  One fused GPU call instead of the N x N x P
        response + P*F Python slicing iterations of the reference."""
        start_grad = time.time()
        newumatsquare_loc = self.flat2square(newumatflat)
        row_start, rows, row_idx = self._fragment_rows()
        gradient = self.helper.response_jacobian(newumatsquare_loc, row_start, rows=rows, row_idx=row_idx,
                                                 NOrotation=self.NOrotation, cache_key=self._umat_key(newumatflat))
        self.time_grad += time.time() - start_grad
        return gradient

    # ------------------------------------------------------------------ src/dmet.py:361-363,381-388
    def costfunction(self, newumatflat):
        return np.linalg.norm(self.rdm_differences(newumatflat)) ** 2

    def costfunction_derivative(self, newumatflat):
        """2 J^T r (src/dmet.py:381-388).  The Jacobian stays on the device; only the P-vector comes back."""
        errors = self.rdm_differences(newumatflat)
        start_grad = time.time()
        row_start, rows, row_idx = self._fragment_rows()
        plan = self.helper.response_jacobian_device(self.flat2square(newumatflat), row_start, rows=rows,
                                                    row_idx=row_idx, NOrotation=self.NOrotation,
                                                    cache_key=self._umat_key(newumatflat))
        thegradient = plan.gradient(errors)
        self.time_grad += time.time() - start_grad
        return thegradient

    # ------------------------------------------------------------------ src/dmet.py:527-539
    def _unit_steps(self, x, h):
        """x + h e_k for every parameter k (one row each)."""
        x = np.asarray(x, dtype=float)
        return x[None, :] + h * np.eye(x.size)

    def verify_gradient(self, umatflat, stepsize=1e-7):
        """Forward finite differences of the cost function against the analytic 2 J^T r.
        Returns (|analytic - numeric|, |analytic|); the reference prints the two norms."""
        analytic = self.costfunction_derivative(umatflat)
        f0 = self.costfunction(umatflat)
        numeric = np.array([(self.costfunction(xk) - f0) / stepsize for xk in self._unit_steps(umatflat, stepsize)])
        return float(np.linalg.norm(analytic - numeric)), float(np.linalg.norm(analytic))

    # ------------------------------------------------------------------ src/dmet.py:541-558
    def hessian_eigenvalues(self, umatflat, stepsize=1e-7):
        """Spectrum of the symmetrised forward-difference Hessian of the cost function: P + 1 analytic
        gradients, each of which is one fused Jacobian + one device-side J^T r (no Jacobian leaves HBM)."""
        g0 = self.costfunction_derivative(umatflat)
        cols = [(self.costfunction_derivative(xk) - g0) / stepsize for xk in self._unit_steps(umatflat, stepsize)]
        hess = np.stack(cols, axis=1) if cols else np.zeros((0, 0))
        return np.linalg.eigvalsh(0.5 * (hess + hess.T))

    # ------------------------------------------------------------------ SURVEY.md 8 f-3
    def leastsq_gpu(self, x0, xtol=1e-13, gtol=1e-13, max_iter=100, verbose=False):
        """Levenberg-Marquardt minimisation of |rdm_differences(x)|^2 with every O(Q P^2) step on the device:
        the Jacobian never leaves HBM, J^T J / J^T r come from cuBLAS, the damped normal equations from a
        cuSOLVER Cholesky (qcdmet_plan_lm_step).  Stands in for scipy.optimize.leastsq / MINPACK lmder
        (src/dmet.py:631) when the (Q, P) Jacobian is too large to ship to the host every iteration.
        An accepted trial point costs one eigensolve (shared by its residual and its Jacobian); a rejected
        one costs one eigensolve and one Cholesky."""
        from ._lib import QcdmetError
        row_start, rows, row_idx = self._fragment_rows()
        x = np.array(x0, dtype=float, copy=True)
        r = self.rdm_differences(x)
        cost = float(np.dot(r, r))
        lam = 1e-3
        self.lm_iterations = 0
        for it in range(max_iter):
            start_grad = time.time()
            plan = self.helper.response_jacobian_device(self.flat2square(x), row_start, rows=rows, row_idx=row_idx,
                                                        NOrotation=self.NOrotation, cache_key=self._umat_key(x))
            self.time_grad += time.time() - start_grad
            plan.set_lm_method(self.lm_method)
            self.lm_iterations += 1
            accepted = False
            for _ in range(40):
                try:
                    step, gmax, gdots = plan.lm_step(r, lam)
                except QcdmetError as exc:
                    if exc.code != -3:
                        raise
                    lam *= 10.0
                    continue
                if gmax < gtol:
                    return x
                x_new = x + step
                r_new = self.rdm_differences(x_new)
                cost_new = float(np.dot(r_new, r_new))
                if verbose:
                    print(f"   LM {it}: cost {cost:.6e} -> {cost_new:.6e}  lambda {lam:.1e}  |step| {np.abs(step).max():.2e}")
                if cost_new <= cost:
                    small = np.abs(step).max() < xtol * (1.0 + np.abs(x).max())
                    x, r, cost = x_new, r_new, cost_new
                    lam = max(lam / 10.0, 1e-15)
                    accepted = True
                    if small:
                        return x
                    break
                lam *= 10.0
            if not accepted:
                return x
        return x

    # ------------------------------------------------------------------ src/dmet.py:626-636
    def fit_umat(self):
        """The umat solve of one DMET iteration (the SCmethod dispatch of doselfconsistent)."""
        from scipy import optimize
        start_cf = time.time()
        if self.altcostfunc and self.SCmethod == 'BFGS':        # src/dmet.py:627-629
            result = optimize.minimize(self.alt_costfunction, self.square2flat(self.umat),
                                       jac=self.alt_costfunction_derivative, options={'disp': False})
            self.umat = self.flat2square(result.x)
        elif self.SCmethod == 'LSTSQ' and self.fit_on_gpu:
            self.umat = self.flat2square(self.leastsq_gpu(self.square2flat(self.umat)))
        elif self.SCmethod == 'LSTSQ':
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
