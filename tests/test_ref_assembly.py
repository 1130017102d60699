"""CPU check of the checker used by the config-size GPU tests and by bench.py's parity gate: Jacobian columns
assembled from the compiled reference's full response must equal the oracle's fragment-block Jacobian."""
import numpy as np
import pytest

from oracle import response_oracle as ro
from qcdmet_b200 import workloads as wl
from util import ref_available, ref_jacobian_columns

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("case", ["sweep", "bath", "ti_bath", "ragged"])
def test_reference_assembly_matches_oracle(case):
    if case == "sweep":
        w = wl.sweep(48)
    elif case == "bath":
        w = wl.make_workload("bath", 40, 18, [4, 2, 6, 4], fit_imp_bath=True)
    elif case == "ti_bath":
        w = wl.hubbard24(trans_inv=True, fit_imp_bath=True, l=6)
    else:
        w = wl.make_workload("ragged", 37, 15, [5, 9, 3])
    e, c = ro.eigh_sorted(w.h0)
    full = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    cols = np.unique(np.linspace(0, w.nterms - 1, 7).round().astype(int))
    got = ref_jacobian_columns(w, cols, chunk=3)
    assert np.abs(got - full[:, cols]).max() <= 1e-10


def test_bench_assembly_and_spread_columns():
    import bench
    w = wl.sweep(32)
    cols = bench.spread_columns(w.nterms, 9)
    assert cols[0] == 0 and cols[-1] == w.nterms - 1 and len(np.unique(cols)) == len(cols)
    info = bench.cpu_reference_rate(w, cols, keep_response=True)
    jref = bench.jacobian_columns_from_response(w, info["response"])
    e, c = ro.eigh_sorted(w.h0)
    full = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert np.abs(jref - full[:, cols]).max() <= 1e-10
    assert info["value"] > 0 and info["kind"] == "reference"
