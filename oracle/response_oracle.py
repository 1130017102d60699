"""CPU oracle for the RHF 1-RDM linear response and the umat-fit Jacobian.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (qcdmet_b200/) may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs do, and only as the checker / the timed CPU baseline.

This is a numpy *restatement* (written from the mathematics, not a copy) of

  * lib/rhf_response.cc:32-119      -> rhf_response()          (full N x N x P response)
  * src/qcdmethelper.py:48-65       -> convert_h1_sparse()     (dense H1 list -> CSR triple)
  * src/qcdmethelper.py:84-112      -> construct1rdm_response()
  * src/qcdmethelper.py:114-122     -> construct1rdm_base()
  * src/dmet.py:105-116             -> make_imp_size()
  * src/dmet.py:118-159             -> makelist_h1() / makelist_h1_sparse()
  * src/dmet.py:161-177             -> make_mask()
  * src/dmet.py:560-587             -> flat2square() / square2flat()
  * src/dmet.py:395-439             -> rdm_differences()
  * src/dmet.py:477-525             -> rdm_differences_derivative() / jacobian_from_response()

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md 4, 8c).  The oracle is pinned against the UNMODIFIED reference kernel compiled
into oracle/_ref/ (oracle/build_ref.sh) by tests/test_oracle.py, and against the
fixtures under tests/golden/ that oracle/make_golden.py generated from that compiled
reference in the build container.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# lib/rhf_response.cc
# --------------------------------------------------------------------------------------
def eigh_sorted(h0: np.ndarray):
    """Ascending eigen-decomposition (lib/rhf_response.cc:42-52, dsyev 'V','U')."""
    e, c = np.linalg.eigh(np.asarray(h0, dtype=np.float64))
    idx = np.argsort(e, kind="stable")
    return e[idx], c[:, idx]


def response_from_orbitals(c, e, num_pairs, h1start, h1row, h1col):
    """dD_p for every perturbation from given orbitals (lib/rhf_response.cc:66-108).

    Returns an array of shape (P, N, N) where out[p, col, row] == rdm_deriv[row + N*(col + N*p)]
    (the memory of the reference output, viewed C-order as src/qcdmethelper.py:111 does).
    Every dD_p is symmetric so the index order does not matter numerically.
    """
    c = np.asarray(c, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    n = c.shape[0]
    no = int(num_pairs)
    co, cv = c[:, :no], c[:, no:]
    # lib/rhf_response.cc:66-71   temp[v,o] = -1/(e_v - e_o)
    w = -1.0 / (e[no:, None] - e[None, :no])
    nterms = len(h1start) - 1
    out = np.empty((nterms, n, n), dtype=np.float64)
    for p in range(nterms):
        rows = np.asarray(h1row[h1start[p]:h1start[p + 1]], dtype=np.int64)
        cols = np.asarray(h1col[h1start[p]:h1start[p + 1]], dtype=np.int64)
        # lib/rhf_response.cc:82-90   Z = w * sum_e Cv[row_e,:]^T Co[col_e,:]
        z = w * (cv[rows, :].T @ co[cols, :])
        # lib/rhf_response.cc:97,100  work = 2 Cv Z Co^T
        work = 2.0 * (cv @ z) @ co.T
        # lib/rhf_response.cc:104-108 dD = work + work^T
        out[p] = work + work.T
    return out


def rhf_response(norb, nterms, num_pairs, h1start, h1row, h1col, h0):
    """Restatement of extern "C" rhf_response (lib/rhf_response.cc:32).

    Returns (rdm_deriv[P,N,N], one_rdm[N,N]); the reference overwrites H0 with the 1-RDM
    (lib/rhf_response.cc:57-64) -- returned here instead of mutating the input.
    """
    h0 = np.asarray(h0, dtype=np.float64).reshape(norb, norb)
    e, c = eigh_sorted(h0)
    co = c[:, :num_pairs]
    one_rdm = 2.0 * co @ co.T
    assert len(h1start) == nterms + 1
    return response_from_orbitals(c, e, num_pairs, h1start, h1row, h1col), one_rdm


# --------------------------------------------------------------------------------------
# src/qcdmethelper.py
# --------------------------------------------------------------------------------------
def convert_h1_sparse(list_h1):
    """src/qcdmethelper.py:48-65: np.where(H1 == 1) in row-major order -> CSR triple (int32)."""
    h1start, h1row, h1col = [0], [], []
    for h1 in list_h1:
        r, c = np.where(np.asarray(h1) == 1)
        h1row.extend(r.tolist())
        h1col.extend(c.tolist())
        h1start.append(len(h1row))
    return (np.array(h1start, dtype=np.int32), np.array(h1row, dtype=np.int32),
            np.array(h1col, dtype=np.int32))


def construct1rdm_base(oei, num_pairs):
    """src/qcdmethelper.py:114-122."""
    e, c = eigh_sorted(oei)
    return 2.0 * c[:, :num_pairs] @ c[:, :num_pairs].T


def construct1rdm_response(fock, umat_loc, num_pairs, h1start, h1row, h1col, no_rotation=None):
    """src/qcdmethelper.py:84-112 with doSCF=False (hard-coded in src/dmet.py:44)."""
    oei = np.asarray(fock, dtype=np.float64) + np.asarray(umat_loc, dtype=np.float64)
    if no_rotation is not None:
        oei = no_rotation.T @ oei @ no_rotation
    n = oei.shape[0]
    resp, _ = rhf_response(n, len(h1start) - 1, num_pairs, h1start, h1row, h1col, oei)
    return resp


# --------------------------------------------------------------------------------------
# src/dmet.py : parameter ordering
# --------------------------------------------------------------------------------------
def make_imp_size(imp_clust, trans_inv):
    """src/dmet.py:105-116."""
    nmax = 1 if trans_inv else len(imp_clust)
    return np.array([int(np.sum(np.abs(imp_clust[k]))) for k in range(nmax)], dtype=np.int64)


def makelist_h1_sparse(norb, imp_size, trans_inv, do_det=False):
    """src/dmet.py:118-159 + src/qcdmethelper.py:48-65 without the dense N x N intermediates.

    Element order inside one perturbation equals np.where's row-major order of the dense
    H1 (sorted by row, then column), so the CSR triple is identical to the reference's.
    """
    h1start, h1row, h1col = [0], [], []

    def emit(pairs):
        pairs = sorted(set(pairs))
        for r, c in pairs:
            h1row.append(r)
            h1col.append(c)
        h1start.append(len(h1row))

    if trans_inv:
        m = int(imp_size[0])
        ncell = norb // m
        for row in range(m):
            cols = [row] if do_det else range(row, m)
            for col in cols:
                pairs = []
                for j in range(ncell):
                    pairs.append((j * m + row, j * m + col))
                    pairs.append((j * m + col, j * m + row))
                emit(pairs)
    else:
        jump = 0
        for m in imp_size:
            m = int(m)
            for row in range(m):
                cols = [row] if do_det else range(row, m)
                for col in cols:
                    emit([(jump + row, jump + col), (jump + col, jump + row)])
            jump += m
    return (np.array(h1start, dtype=np.int32), np.array(h1row, dtype=np.int32),
            np.array(h1col, dtype=np.int32))


def makelist_h1(norb, imp_size, trans_inv, do_det=False):
    """Dense list of N x N int matrices exactly as src/dmet.py:118-159 builds it (small N only)."""
    s, r, c = makelist_h1_sparse(norb, imp_size, trans_inv, do_det)
    out = []
    for p in range(len(s) - 1):
        h1 = np.zeros((norb, norb), dtype=int)
        h1[r[s[p]:s[p + 1]], c[s[p]:s[p + 1]]] = 1
        out.append(h1)
    return out


def make_mask(norb, imp_size, do_det=False):
    """src/dmet.py:161-177."""
    mask = np.zeros((norb, norb), dtype=bool)
    jump = 0
    for m in imp_size:
        m = int(m)
        for row in range(m):
            for col in ([row] if do_det else range(row, m)):
                mask[jump + row, jump + col] = True
        jump += m
    return mask


def flat2square(umatflat, norb, mask, imp_size, trans_inv, no_rotation=None):
    """src/dmet.py:560-580."""
    u = np.zeros((norb, norb), dtype=np.float64)
    u[mask] = umatflat
    u = u.T.copy()
    u[mask] = umatflat
    if trans_inv:
        m = int(imp_size[0])
        for it in range(1, norb // m):
            u[it * m:(it + 1) * m, it * m:(it + 1) * m] = u[0:m, 0:m]
    if no_rotation is not None:
        u = no_rotation @ u @ no_rotation.T
    return u


def square2flat(umatsquare, mask, no_rotation=None):
    """src/dmet.py:582-587."""
    u = np.array(umatsquare, dtype=np.float64, copy=True)
    if no_rotation is not None:
        u = no_rotation.T @ u @ no_rotation
    return u[mask]


# --------------------------------------------------------------------------------------
# src/dmet.py : residual and Jacobian
# --------------------------------------------------------------------------------------
def jacobian_from_response(resp, imp_clust, imp_size, fit_imp_bath, dmet_orbs=None):
    """src/dmet.py:484-520: assemble the (Q, P) Jacobian from the (P, N, N) response."""
    nfrag = len(imp_size)
    cols = []
    for p in range(resp.shape[0]):
        pieces = []
        for f in range(nfrag):
            if fit_imp_bath:
                b = np.asarray(dmet_orbs[f], dtype=np.float64)
                loc = b.T @ resp[p] @ b
            else:
                idx = np.flatnonzero(imp_clust[f])
                loc = resp[p][:, idx][idx, :]
            pieces.append(np.reshape(loc, loc.size, order="F"))
        cols.append(np.concatenate(pieces) if pieces else np.zeros(0))
    return np.array(cols).T


def rdm_differences_derivative(fock, umatflat, num_pairs, imp_clust, trans_inv, fit_imp_bath,
                               dmet_orbs=None):
    """src/dmet.py:477-525 (doDET False, NOrotation None)."""
    norb = fock.shape[0]
    imp_size = make_imp_size(imp_clust, trans_inv)
    mask = make_mask(norb, imp_size)
    umat = flat2square(umatflat, norb, mask, imp_size, trans_inv)
    s, r, c = makelist_h1_sparse(norb, imp_size, trans_inv)
    resp = construct1rdm_response(fock, umat, num_pairs, s, r, c)
    return jacobian_from_response(resp, imp_clust, imp_size, fit_imp_bath, dmet_orbs)


def rdm_differences(fock, umatflat, num_pairs, imp_clust, trans_inv, fit_imp_bath, imp_1rdm,
                    dmet_orbs=None):
    """src/dmet.py:395-439 (doDET False)."""
    norb = fock.shape[0]
    imp_size = make_imp_size(imp_clust, trans_inv)
    mask = make_mask(norb, imp_size)
    umat = flat2square(umatflat, norb, mask, imp_size, trans_inv)
    d = construct1rdm_base(np.asarray(fock) + umat, num_pairs)
    pieces = []
    for f in range(len(imp_size)):
        if fit_imp_bath:
            b = np.asarray(dmet_orbs[f], dtype=np.float64)
            err = b.T @ d @ b - imp_1rdm[f]
        else:
            idx = np.flatnonzero(imp_clust[f])
            m = int(imp_size[f])
            err = d[:, idx][idx, :] - imp_1rdm[f][:m, :m]
        pieces.append(np.reshape(err, err.size, order="F"))
    return np.concatenate(pieces)


def jacobian_from_orbitals(c, e, num_pairs, h1start, h1row, h1col, row_start, rows=None,
                           row_idx=None):
    """Fragment-block Jacobian straight from (C, e) (SURVEY.md 8a 'mathematical statement').

    rows   : (S, N) dense projector rows (B_f^T stacked) or None
    row_idx: (S,) site indices (impurity-only) used when rows is None
    Column-major blocks per fragment; result (Q, P).  O(P * S * nv * no): small cases only.
    """
    c = np.asarray(c, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    no = int(num_pairs)
    co, cv = c[:, :no], c[:, no:]
    if rows is not None:
        pv, po = rows @ cv, rows @ co
    else:
        pv, po = cv[row_idx, :], co[row_idx, :]
    w = 1.0 / (e[None, :no] - e[no:, None])
    nterms = len(h1start) - 1
    nfrag = len(row_start) - 1
    q = int(sum((row_start[f + 1] - row_start[f]) ** 2 for f in range(nfrag)))
    jac = np.empty((q, nterms), dtype=np.float64)
    for p in range(nterms):
        rr = np.asarray(h1row[h1start[p]:h1start[p + 1]], dtype=np.int64)
        cc = np.asarray(h1col[h1start[p]:h1start[p + 1]], dtype=np.int64)
        z = w * (cv[rr, :].T @ co[cc, :])
        y = pv @ z
        off = 0
        for f in range(nfrag):
            a0, a1 = row_start[f], row_start[f + 1]
            g = y[a0:a1] @ po[a0:a1].T
            blk = 2.0 * (g + g.T)
            jac[off:off + blk.size, p] = blk.reshape(-1, order="F")
            off += blk.size
    return jac
