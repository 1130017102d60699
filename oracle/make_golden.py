"""Generate tests/golden/*.npz from the UNMODIFIED reference kernel compiled in oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference to have been
compiled by oracle/build_ref.sh):   python oracle/make_golden.py
Each fixture stores the inputs (H0, CSR perturbation list, fragment rows) and the compiled
reference's outputs (full response, 1-RDM), plus the Jacobian assembled from that response by
the rule of src/dmet.py:496-520.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, response_oracle as ro   # noqa: E402
from qcdmet_b200 import workloads as wl                # noqa: E402


def subsample(w, idx, name):
    """Keep the perturbations idx of a workload (golden files store the full N x N response per column)."""
    hs, hr, hc = [0], [], []
    for p in idx:
        a, b = w.h1start[p], w.h1start[p + 1]
        hr += list(w.h1row[a:b]); hc += list(w.h1col[a:b]); hs.append(len(hr))
    w.h1start, w.h1row, w.h1col = (np.array(hs, dtype=np.int32), np.array(hr, dtype=np.int32),
                                   np.array(hc, dtype=np.int32))
    w.name = name
    return w


def cases():
    # fragments of more than 8 rows (factorised path of the library): packed 14-row fragments, impurity+bath
    # rows of a 10- and a 6-orbital impurity, one 100-row fragment (cut into slice pairs by the contraction)
    w = wl.make_workload("frag14_N42", 42, 15, [14, 14, 14])
    yield subsample(w, range(0, w.nterms, 7), "frag14_N42")
    w = wl.make_workload("frag_bath_N40", 40, 17, [10, 6], fit_imp_bath=True)
    yield subsample(w, range(0, w.nterms, 2), "frag_bath_N40")
    w = wl.make_workload("bigfrag_N104", 104, 40, [100])
    yield subsample(w, [0, 1, 100, 2500, 5000, 5049], "bigfrag_N104")
    yield wl.hydrogen_like()
    yield wl.sweep(24)
    yield wl.make_workload("mixed_N30", 30, 11, [5, 3, 7])
    yield wl.make_workload("mixed_bath_N26", 26, 9, [3, 2, 4], fit_imp_bath=True)
    yield wl.hubbard24(trans_inv=True, l=6)
    yield wl.hubbard24(trans_inv=False, l=6)
    yield wl.make_workload("partial_N40", 40, 20, [4, 4, 6])        # impurities do not tile the system
    yield wl.make_workload("ti_bathless_N32", 32, 14, [4] * 8, trans_inv=True)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only_missing = "--only-missing" in sys.argv
    for w in cases():
        if only_missing and os.path.exists(os.path.join(out_dir, f"{w.name}.npz")):
            continue
        resp, one_rdm = ref_loader.rhf_response(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0, "mkl")
        resp_ob, _ = ref_loader.rhf_response(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.h0, "openblas")
        assert np.abs(resp - resp_ob).max() < 1e-11, "MKL and OpenBLAS builds of the reference disagree"
        imp_size = ro.make_imp_size(w.imp_clust, w.trans_inv)
        jac = ro.jacobian_from_response(resp, w.imp_clust, imp_size, w.fit_imp_bath, w.dmet_orbs)
        path = os.path.join(out_dir, f"{w.name}.npz")
        np.savez_compressed(
            path, norb=w.norb, num_pairs=w.num_pairs, h0=w.h0, h1start=w.h1start, h1row=w.h1row,
            h1col=w.h1col, row_start=w.row_start,
            row_idx=w.row_idx if w.row_idx is not None else np.zeros(0, dtype=np.int32),
            rows=w.rows if w.rows is not None else np.zeros((0, 0)),
            trans_inv=w.trans_inv, fit_imp_bath=w.fit_imp_bath,
            imp_clust=np.array(w.imp_clust), resp=resp, one_rdm=one_rdm, jac=jac)
        print(f"{path}: N={w.norb} P={w.nterms} Q={jac.shape[0]}  {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
