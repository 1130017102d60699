import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["name"] = os.path.basename(path)[:-4]
    g["norb"] = int(g["norb"])
    g["num_pairs"] = int(g["num_pairs"])
    g["trans_inv"] = bool(g["trans_inv"])
    g["fit_imp_bath"] = bool(g["fit_imp_bath"])
    g["rows"] = g["rows"] if g["rows"].size else None
    g["row_idx"] = g["row_idx"] if g["rows"] is None else None
    return g


def golden_ids():
    return [os.path.basename(p)[:-4] for p in GOLDEN]


def ref_available():
    from oracle import ref_loader
    return ref_loader.available("mkl")


def ref_jacobian_columns(w, cols, chunk=256):
    """Jacobian columns `cols` of workload w from the UNMODIFIED reference kernel compiled into oracle/_ref
    (its own dsyev, its own response loop), assembled by the rule of src/dmet.py:496-520.  The reference is
    called on chunks of columns that respect its 32-bit output index (lib/rhf_response.cc:106)."""
    from oracle import ref_loader
    cols = np.asarray(cols, dtype=np.int64)
    n = w.norb
    chunk = max(1, min(chunk, (2 ** 31 - 1) // (n * n) - 1))
    if w.rows is None:
        ra, rb = [], []
        for f in range(len(w.row_start) - 1):
            idx = w.row_idx[w.row_start[f]:w.row_start[f + 1]]
            a, b = np.meshgrid(idx, idx, indexing="ij")
            ra.append(a.reshape(-1, order="F"))
            rb.append(b.reshape(-1, order="F"))
        ra, rb = np.concatenate(ra), np.concatenate(rb)
    out = np.empty((w.nresid, len(cols)))
    for c0 in range(0, len(cols), chunk):
        part = cols[c0:c0 + chunk]
        lens = (w.h1start[part + 1] - w.h1start[part]).astype(np.int64)
        hs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        take = np.concatenate([np.arange(w.h1start[c], w.h1start[c + 1]) for c in part])
        resp, _ = ref_loader.rhf_response(n, w.num_pairs, hs, w.h1row[take], w.h1col[take], w.h0, "mkl")
        if w.rows is None:
            for k in range(len(part)):
                out[:, c0 + k] = resp[k][ra, rb]
        else:
            # fitImpBath: B_f^T dD_p B_f for every fragment, one GEMM pair per fragment over the chunk
            off = 0
            for f in range(len(w.row_start) - 1):
                b = w.rows[w.row_start[f]:w.row_start[f + 1]]                 # (nb, N) = B_f^T
                nb = b.shape[0]
                loc = np.einsum("an,pnm,bm->pab", b, resp, b, optimize=True)   # (chunk, nb, nb)
                out[off:off + nb * nb, c0:c0 + len(part)] = loc.transpose(0, 2, 1).reshape(len(part), nb * nb).T
                off += nb * nb
    return out
