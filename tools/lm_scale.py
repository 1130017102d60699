#!/usr/bin/env python
"""Device Levenberg-Marquardt (dmet.leastsq_gpu) at the sizes the unit tests do not reach.

  python tools/lm_scale.py north_star   N = 4000, P = 10^4, Q = 16 000: the targets are the mean-field blocks at a known
                                        umat x*, so the fit must return to x* (up to the chemical-potential gauge);
                                        prints the time per LM iteration
  python tools/lm_scale.py hubbard24    24 x 24 lattice (staggered potential), 2 x 2 impurities, non-TI, P = 1 440:
                                        device LM (both routes) against scipy.optimize.leastsq (MINPACK)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcdmet_b200 import _lib, dmet as dmet_mod, localintegrals_hubbard as lh, workloads as wl   # noqa: E402


class FockInts:
    """Minimal integral provider for the fit path (src/qcdmethelper.py:30-32,71-81,87): a fixed Fock matrix."""

    def __init__(self, fock, nelec):
        self.Norbs, self.Nelec, self.ERIinMEM, self.TI_OK = fock.shape[0], nelec, False, True
        self._fock = fock

    def loc_rhf_fock(self):
        return self._fock

    def loc_oei(self):
        return self._fock


def gauge(u):
    return u - np.eye(u.shape[0]) * np.average(np.diag(u))


def north_star():
    w = wl.sweep(4000)
    d = dmet_mod.dmet(FockInts(w.h0, 2 * w.num_pairs), w.imp_clust, False, method='ED', SCmethod='LSTSQ', fitImpBath=False)
    p = d.helper.Nterms
    rng = np.random.default_rng(1)
    x_true = 2e-3 * rng.standard_normal(p)
    # targets: the mean-field impurity blocks at umat(x_true)
    d.imp_1RDM = [np.zeros((4, 4)) for _ in range(len(w.imp_clust))]
    r_true = d.rdm_differences(x_true)                       # = model(x_true) - 0
    q = 0
    for f in range(len(w.imp_clust)):
        d.imp_1RDM[f] = r_true[q:q + 16].reshape(4, 4, order="F").copy()
        q += 16
    assert np.abs(d.rdm_differences(x_true)).max() < 1e-12
    x0 = np.zeros(p)
    print(f"north star: P = {p}, Q = {r_true.size}, cost at x0 = {d.costfunction(x0):.3e}")
    for method, name in ((_lib.LM_NORMAL, "normal equations + Cholesky"),):
        d.lm_method = method
        t0 = time.time()
        x = d.leastsq_gpu(x0, max_iter=12)
        dt = time.time() - t0
        err = np.abs(gauge(d.flat2square(x)) - gauge(d.flat2square(x_true))).max()
        print(f"  {name}: {d.lm_iterations} LM iterations in {dt:.1f} s ({dt / d.lm_iterations:.2f} s per iteration: one eigensolve + "
              f"Jacobian + J^T J + Cholesky, plus one eigensolve per trial point), cost {d.costfunction(x):.3e}, "
              f"max |umat - umat*| (gauge removed) = {err:.2e}")


def hubbard24():
    from scipy import optimize
    l = 24
    n = l * l
    stagger = np.array([0.4 if (k % 4) in (0, 3) else -0.4 for k in range(n)])
    ints = lh.localintegrals_hubbard(wl.hubbard2d_hopping(l, l) + np.diag(stagger), 4.0, n)
    clusters = []
    for k in range(n // 4):
        v = np.zeros(n, dtype=int)
        v[4 * k:4 * k + 4] = 1
        clusters.append(v)
    d = dmet_mod.dmet(ints, clusters, False, method='ED', SCmethod='LSTSQ', fitImpBath=False)
    mf = ints.origDMloc
    d.imp_1RDM = []
    for f in range(n // 4):
        blk = mf[4 * f:4 * f + 4, 4 * f:4 * f + 4].copy()
        blk[~np.eye(4, dtype=bool)] *= 0.8
        d.imp_1RDM.append(blk)
    x0 = np.zeros(d.helper.Nterms)
    print(f"hubbard 24 x 24: P = {x0.size}, Q = {d.rdm_differences(x0).size}, cost at x0 = {d.costfunction(x0):.3e}")
    t0 = time.time()
    ref = optimize.leastsq(d.rdm_differences, x0, Dfun=d.rdm_differences_derivative, factor=0.1,
                           ftol=1e-15, xtol=1e-15, gtol=0.0, full_output=True)
    t_ref = time.time() - t0
    print(f"  MINPACK (scipy.optimize.leastsq, GPU Jacobian): {ref[2]['nfev']} function evaluations in {t_ref:.1f} s, "
          f"cost {d.costfunction(ref[0]):.3e}")
    for method, name in ((_lib.LM_NORMAL, "normal equations + Cholesky"), (_lib.LM_QR, "QR of [J; sqrt(lambda) D]")):
        d.lm_method = method
        t0 = time.time()
        x = d.leastsq_gpu(x0)
        dt = time.time() - t0
        err = np.abs(gauge(d.flat2square(x)) - gauge(d.flat2square(ref[0]))).max()
        print(f"  device LM, {name}: {d.lm_iterations} iterations in {dt:.2f} s, cost {d.costfunction(x):.3e}, "
              f"max |umat - umat_MINPACK| (gauge removed) = {err:.2e}")


if __name__ == "__main__":
    {"north_star": north_star, "hubbard24": hubbard24}[sys.argv[1] if len(sys.argv) > 1 else "hubbard24"]()
