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
