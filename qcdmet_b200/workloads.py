"""Synthetic workloads of the shapes named in BASELINE.json / SURVEY.md 8(d).

Host-side input generators only (numpy); used by bench.py, __graft_entry__.smoke() and the
tests so that every leg (GPU path, oracle, compiled reference) sees the same seeded inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    norb: int
    num_pairs: int
    h0: np.ndarray                 # (N, N) symmetric "Fock + umat"
    imp_clust: list                # list of length-N 0/1 int vectors (src/dmet.py impurityClusters)
    trans_inv: bool
    fit_imp_bath: bool
    h1start: np.ndarray
    h1row: np.ndarray
    h1col: np.ndarray
    row_start: np.ndarray          # (F+1,) fragment row offsets
    row_idx: np.ndarray | None     # (S,) site index per row (impurity-only)
    rows: np.ndarray | None        # (S, N) dense projector rows (fitImpBath), row-major
    dmet_orbs: list | None = None  # per-fragment (N, nb_f) embedding orbitals when fit_imp_bath
    meta: dict = field(default_factory=dict)

    @property
    def nterms(self) -> int:
        return len(self.h1start) - 1

    @property
    def nrows(self) -> int:
        return int(self.row_start[-1])

    @property
    def nresid(self) -> int:
        d = np.diff(self.row_start)
        return int(np.sum(d * d))


def spectrum_matrix(norb: int, num_pairs: int, seed: int, gap_half: float = 0.1):
    """H0 = U diag(e) U^T with a controlled HOMO-LUMO gap.

    e_occ ~ U(-1,-gap_half), e_virt ~ U(+gap_half,+1), U = Q factor of a standard-normal matrix.
    Returns (h0, e_sorted, U_sorted).
    """
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((norb, norb))
    u, _ = np.linalg.qr(g)
    eo = np.sort(rng.uniform(-1.0, -gap_half, size=num_pairs))
    ev = np.sort(rng.uniform(gap_half, 1.0, size=norb - num_pairs))
    e = np.concatenate([eo, ev])
    h0 = (u * e[None, :]) @ u.T
    h0 = 0.5 * (h0 + h0.T)
    return h0, e, u


def _clusters_from_sizes(norb, sizes):
    clusters, jump = [], 0
    for m in sizes:
        v = np.zeros(norb, dtype=int)
        v[jump:jump + m] = 1
        clusters.append(v)
        jump += m
    assert jump <= norb
    return clusters


def _h1_sparse(norb, imp_size, trans_inv):
    """Parameter ordering of src/dmet.py:118-159 in CSR form (same element order as
    src/qcdmethelper.py:48-65 produces from the dense list)."""
    h1start, h1row, h1col = [0], [], []

    def emit(pairs):
        for r, c in sorted(set(pairs)):
            h1row.append(r)
            h1col.append(c)
        h1start.append(len(h1row))

    if trans_inv:
        m = int(imp_size[0])
        for row in range(m):
            for col in range(row, m):
                pairs = []
                for j in range(norb // m):
                    pairs += [(j * m + row, j * m + col), (j * m + col, j * m + row)]
                emit(pairs)
    else:
        jump = 0
        for m in imp_size:
            m = int(m)
            for row in range(m):
                for col in range(row, m):
                    emit([(jump + row, jump + col), (jump + col, jump + row)])
            jump += m
    return (np.array(h1start, dtype=np.int32), np.array(h1row, dtype=np.int32),
            np.array(h1col, dtype=np.int32))


def make_workload(name, norb, num_pairs, frag_sizes, *, trans_inv=False, fit_imp_bath=False,
                  nterms=None, seed=None, h0=None, gap_half=0.1):
    """Generic builder: contiguous fragments of the given sizes starting at orbital 0."""
    if seed is None:
        seed = 20240 + norb
    if h0 is None:
        h0, _, _ = spectrum_matrix(norb, num_pairs, seed, gap_half)
    clusters = _clusters_from_sizes(norb, frag_sizes)
    imp_size = np.array(frag_sizes[:1] if trans_inv else frag_sizes, dtype=np.int64)
    hs, hr, hc = _h1_sparse(norb, imp_size, trans_inv)
    if nterms is not None and nterms < len(hs) - 1:
        hs = hs[:nterms + 1].copy()
        hr = hr[:hs[-1]].copy()
        hc = hc[:hs[-1]].copy()
    nfit = 1 if trans_inv else len(frag_sizes)
    rows = row_idx = dmet_orbs = None
    if fit_imp_bath:
        # embedding orbitals from the mean-field 1-RDM: impurity sites + Schmidt bath
        # (src/qcdmethelper.py:124-168 shape: N x (imp + bath), orthonormal columns)
        e, c = np.linalg.eigh(h0)
        d = 2.0 * c[:, :num_pairs] @ c[:, :num_pairs].T
        dmet_orbs, blocks, sizes = [], [], []
        for f in range(nfit):
            imp = np.flatnonzero(clusters[f])
            env = np.flatnonzero(clusters[f] == 0)
            m = len(imp)
            occ, vec = np.linalg.eigh(d[np.ix_(env, env)])
            order = np.argsort(np.maximum(-occ, occ - 2.0))[:m]
            b = np.zeros((norb, 2 * m))
            b[imp, np.arange(m)] = 1.0
            b[np.ix_(env, m + np.arange(m))] = vec[:, order]
            dmet_orbs.append(b)
            blocks.append(b.T.copy())
            sizes.append(2 * m)
        rows = np.ascontiguousarray(np.concatenate(blocks, axis=0))
        row_start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    else:
        idx = [np.flatnonzero(clusters[f]) for f in range(nfit)]
        row_idx = np.concatenate(idx).astype(np.int32)
        row_start = np.concatenate([[0], np.cumsum([len(i) for i in idx])]).astype(np.int32)
    return Workload(name=name, norb=norb, num_pairs=num_pairs, h0=h0, imp_clust=clusters,
                    trans_inv=trans_inv, fit_imp_bath=fit_imp_bath, h1start=hs, h1row=hr, h1col=hc,
                    row_start=row_start, row_idx=row_idx, rows=rows, dmet_orbs=dmet_orbs,
                    meta={"seed": seed, "frag_sizes": list(map(int, frag_sizes))})


def sweep(norb: int, nterms: int | None = None, m: int = 4) -> Workload:
    """Primary synthetic sweep: no = N/2, 4-site fragments tiling all N,
    impurity-only rows, P_max = 2.5 N (N=4000 -> 10^4), seed = 20240 + N."""
    assert norb % m == 0
    w = make_workload(f"sweep_N{norb}", norb, norb // 2, [m] * (norb // m), nterms=nterms)
    w.name = f"sweep_N{norb}_P{w.nterms}"
    return w


def hubbard2d_hopping(lx: int, ly: int, t: float = 1.0, bc=(1.0, -1.0), cluster=(2, 2)):
    """Nearest-neighbour hopping on an lx x ly torus, sites ordered cluster by cluster
    (2x2 impurity = orbitals 0..3), PBC x anti-PBC by default so half filling is closed-shell
    (cf. examples/hubbard.py:40-41 for the 1-D anti-PBC trick)."""
    cx, cy = cluster
    assert lx % cx == 0 and ly % cy == 0
    index = {}
    k = 0
    for bx in range(lx // cx):
        for by in range(ly // cy):
            for ix in range(cx):
                for iy in range(cy):
                    index[(bx * cx + ix, by * cy + iy)] = k
                    k += 1
    n = lx * ly
    h = np.zeros((n, n))
    for x in range(lx):
        for y in range(ly):
            i = index[(x, y)]
            for dx, dy in ((1, 0), (0, 1)):
                x2, y2 = x + dx, y + dy
                sgn = 1.0
                if x2 == lx:
                    x2, sgn = 0, sgn * bc[0]
                if y2 == ly:
                    y2, sgn = 0, sgn * bc[1]
                j = index[(x2, y2)]
                h[i, j] += -t * sgn
                h[j, i] += -t * sgn
    return h


def hubbard24(trans_inv: bool = False, fit_imp_bath: bool = False, l: int = 24, u_shift: float = 0.0):
    """2-D Hubbard l x l half filling, 2x2 impurities (BASELINE.json configs[1])."""
    n = l * l
    h0 = hubbard2d_hopping(l, l)
    if u_shift:
        h0 = h0 + u_shift * np.eye(n)
    w = make_workload(f"hubbard{l}x{l}" + ("_TI" if trans_inv else ""), n, n // 2, [4] * (n // 4),
                      trans_inv=trans_inv, fit_imp_bath=fit_imp_bath, h0=h0)
    return w


def polyyne_like():
    """Shape of examples/polyyne.py with cc-pVDZ: N=168, no=36, 12 fragments of 14."""
    return make_workload("polyyne_like", 168, 36, [14] * 12, seed=20240 + 168)


def sn2_like():
    """Shape of examples/sn2.py:74: N=321, no=58, unit sizes [52, 24 x 10, 29]."""
    return make_workload("sn2_like", 321, 58, [52] + [24] * 10 + [29], seed=20240 + 321)


def hydrogen_like():
    """Shape of examples/hydrogen.py: N=10, no=5, TI with 2-site cells, fitImpBath (Q=16, P=3)."""
    return make_workload("hydrogen_like", 10, 5, [2] * 5, trans_inv=True, fit_imp_bath=True,
                         seed=20240 + 10)
