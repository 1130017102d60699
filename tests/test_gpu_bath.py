"""Batched bath construction on the device (qcdmet_plan_construct_baths) against the host
restatement of src/qcdmethelper.py:124-168 (numpy eigh of the environment block, one fragment at a time)."""
import time

import numpy as np
import pytest

from oracle import response_oracle as ro
from qcdmet_b200 import _lib, dmet as dmet_mod, localintegrals_hubbard as lh, qcdmethelper as qh, workloads as wl
from dmet_harness import CASES, OracleDmet, build, chain_clusters

pytestmark = pytest.mark.gpu


def host_bath(one_rdm, imp_idx, nbath, threshold=1e-13):
    flags = np.zeros(one_rdm.shape[0], dtype=int)
    flags[imp_idx] = 1
    helper = qh.qcdmethelper.__new__(qh.qcdmethelper)       # constructbath uses no state
    nb, loc2dmet, core = qh.qcdmethelper.constructbath(helper, one_rdm, flags, nbath, threshold)
    return nb, loc2dmet, core


def plan_for(h0, num_pairs, imp_start, imp_idx):
    n = h0.shape[0]
    h1start = np.array([0, 1], dtype=np.int32)             # any perturbation list: the baths only need the orbitals
    plan = _lib.Plan(n, num_pairs, h1start, np.array([0], dtype=np.int32), np.array([0], dtype=np.int32),
                     np.asarray(imp_start, dtype=np.int32), None, np.asarray(imp_idx, dtype=np.int32))
    plan.set_h0(h0)
    return plan


def random_problem(n, num_pairs, sizes, seed):
    h0, _, _ = wl.spectrum_matrix(n, num_pairs, seed)
    rng = np.random.default_rng(seed + 1)
    perm = rng.permutation(n)
    imp_start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    imp_idx = perm[:imp_start[-1]].astype(np.int32)        # disjoint, unordered on purpose inside a fragment
    for f in range(len(sizes)):
        imp_idx[imp_start[f]:imp_start[f + 1]].sort()
    return h0, imp_start, imp_idx


@pytest.mark.parametrize("n,num_pairs,sizes", [
    (40, 17, [1, 2, 5, 4]),
    (97, 41, [13, 7, 8, 1, 3]),                 # odd sizes: padded Jacobi index
    (150, 60, [52, 24, 29]),                    # sn2-like unit sizes (examples/sn2.py:74)
    (64, 32, [4] * 16),                         # tiling of all orbitals
    (60, 10, [20, 13, 4]),                      # more impurity orbitals than occupied orbitals: rank-deficient P_II
    (60, 49, [20, 13, 4]),                      # ... than virtual orbitals: eigenvalues 1 of P_II
])
def test_baths_match_host_constructbath(n, num_pairs, sizes):
    h0, imp_start, imp_idx = random_problem(n, num_pairs, sizes, seed=11 + n)
    plan = plan_for(h0, num_pairs, imp_start, imp_idx)
    num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx)
    one = plan.one_rdm()
    c, e = plan.orbitals(), plan.eigenvalues()
    assert np.abs(one - 2 * c[:, :num_pairs] @ c[:, :num_pairs].T).max() < 1e-12
    for f, m in enumerate(sizes):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        nb_ref, loc2dmet, core = host_bath(one, idx, m)
        want = min(m, n - num_pairs, num_pairs)             # entangled orbitals of a generic state
        assert nb_ref == num_bath[f] == want
        got = bath[imp_start[f]:imp_start[f] + nb_ref].T    # (N, nb)
        ref = loc2dmet[:, m:m + nb_ref]
        overlap = np.abs(np.sum(got * ref, axis=0))
        assert np.abs(overlap - 1.0).max() < 1e-10          # same orbitals up to the sign of each column
        sign = np.sign(np.sum(got * ref, axis=0))
        assert np.abs(got * sign - ref).max() < 1e-9
        assert np.abs(got[idx]).max() == 0.0                # no weight on the impurity
        occ_ref = np.diag(ref.T @ one @ ref)
        assert np.abs(occ[imp_start[f]:imp_start[f] + nb_ref] - occ_ref).max() < 1e-11
        assert np.abs(bath[imp_start[f] + nb_ref:imp_start[f + 1]]).max(initial=0.0) == 0.0
    plan.close()


def test_requested_bath_count_and_threshold():
    # fragment 1 holds a site that is decoupled from everything else: its orbital is not entangled with the
    # environment and the reference drops the corresponding bath orbital (src/qcdmethelper.py:141-144)
    n, num_pairs = 24, 10
    h0, _, _ = wl.spectrum_matrix(n, num_pairs, 5)
    h0[7, :] = 0.0
    h0[:, 7] = 0.0
    h0[7, 7] = -5.0                                          # doubly occupied, isolated
    imp_start = np.array([0, 3, 6], dtype=np.int32)
    imp_idx = np.array([1, 2, 3, 6, 7, 8], dtype=np.int32)
    plan = plan_for(h0, num_pairs, imp_start, imp_idx)
    one = plan.one_rdm()
    num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx, num_bath_req=[2, 3])
    assert list(num_bath) == [2, 2]                          # request honoured; unentangled orbital dropped
    for f in range(2):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        nb_ref, loc2dmet, _ = host_bath(one, idx, [2, 3][f])
        assert nb_ref == num_bath[f]
        got, ref = bath[imp_start[f]:imp_start[f] + nb_ref].T, loc2dmet[:, 3:3 + nb_ref]
        assert np.abs(np.abs(np.sum(got * ref, axis=0)) - 1.0).max() < 1e-10
    assert abs(occ[imp_start[1] + 2]) < 1e-12 or abs(occ[imp_start[1] + 2] - 2.0) < 1e-12
    with pytest.raises(_lib.QcdmetError):
        plan.construct_baths(np.array([0, 65], dtype=np.int32), np.arange(65, dtype=np.int32) % n)
    with pytest.raises(_lib.QcdmetError):
        plan.construct_baths(np.array([0, 2], dtype=np.int32), np.array([3, 3], dtype=np.int32))
    plan.close()


def test_hubbard24_all_fragments_in_one_call():
    """configs[1]: 24 x 24 lattice, 144 impurities of 2 x 2 sites.  The lattice symmetry makes bath occupations
    degenerate, so orbitals are compared through the projector on the bath space and the embedded 1-RDM."""
    w = wl.hubbard24(l=24)
    n, num_pairs = w.norb, w.num_pairs
    imp_start, imp_idx = w.row_start, w.row_idx
    plan = plan_for(w.h0, num_pairs, imp_start, imp_idx)
    plan.construct_baths(imp_start, imp_idx)                 # warm-up (allocations, cuBLAS)
    t0 = time.perf_counter()
    num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx)
    t_dev = time.perf_counter() - t0
    one = plan.one_rdm()
    assert np.all(num_bath == 4)
    t_host = 0.0
    for f in (0, 71, 143):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        t0 = time.perf_counter()
        nb_ref, loc2dmet, _ = host_bath(one, idx, 4)
        t_host += time.perf_counter() - t0
        got, ref = bath[imp_start[f]:imp_start[f] + 4].T, loc2dmet[:, 4:8]
        assert np.abs(got.T @ got - np.eye(4)).max() < 1e-10
        assert np.abs(got @ got.T - ref @ ref.T).max() < 1e-8
        emb = np.concatenate([np.eye(n)[:, idx], got], axis=1)
        d_emb = emb.T @ one @ emb
        assert abs(np.trace(d_emb) - 8.0) < 1e-9            # half-filled 8-orbital embedded problem
        assert np.abs(np.sort(occ[imp_start[f]:imp_start[f + 1]]) - np.sort(np.diag(ref.T @ one @ ref))).max() < 1e-9
    print(f"\nbaths of 144 fragments (N = 576): device {t_dev * 1e3:.2f} ms for all, host {t_host / 3 * 1e3:.1f} ms per fragment")
    assert t_dev < t_host                                    # all 144 on the device against 3 on the host
    plan.close()


def test_north_star_shape_1000_fragments():
    """N = 4000, 1000 impurities of 4 orbitals: every bath in one call; the host route (eigh of a 3996 x 3996
    block per fragment) is run for one fragment only."""
    w = wl.sweep(4000, nterms=8)
    n, num_pairs = w.norb, w.num_pairs
    imp_start, imp_idx = w.row_start, w.row_idx
    assert len(imp_start) - 1 == 1000
    plan = plan_for(w.h0, num_pairs, imp_start, imp_idx)
    plan.construct_baths(imp_start[:3], imp_idx[:8])
    t0 = time.perf_counter()
    num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx)
    t_dev = time.perf_counter() - t0
    assert np.all(num_bath == 4)
    c = plan.orbitals()
    one = 2 * c[:, :num_pairs] @ c[:, :num_pairs].T
    f = 517
    idx = imp_idx[imp_start[f]:imp_start[f + 1]]
    t0 = time.perf_counter()
    nb_ref, loc2dmet, _ = host_bath(one, idx, 4)
    t_host = time.perf_counter() - t0
    got, ref = bath[imp_start[f]:imp_start[f] + 4].T, loc2dmet[:, 4:8]
    assert np.abs(np.abs(np.sum(got * ref, axis=0)) - 1.0).max() < 1e-9
    # every fragment: orthonormal baths, no weight on the impurity, Schmidt property D b = occ b on the environment
    b = bath.reshape(1000, 4, n)
    gram = np.einsum("fkn,fln->fkl", b, b)
    assert np.abs(gram - np.eye(4)).max() < 1e-10
    for f in (0, 999):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        env = np.setdiff1d(np.arange(n), idx)
        bb = b[f]
        resid = one[np.ix_(env, env)] @ bb[:, env].T - bb[:, env].T * occ[imp_start[f]:imp_start[f] + 4]
        assert np.abs(resid).max() < 1e-10
    print(f"\nbaths of 1000 fragments (N = 4000): device {t_dev * 1e3:.1f} ms for all (256 MB to the host included), "
          f"host {t_host:.2f} s per fragment")
    plan.close()


@pytest.mark.parametrize("case", ["chain12_U4_imp_only", "chain8_U2_bath", "chain12_U4_TI_bath"])
def test_dmet_loop_with_device_baths(case):
    """The whole self-consistency loop with bath_on_gpu: same converged energy and umat as the host-bath loop."""
    ref = build(case, dmet_mod.dmet)
    e_ref = ref.doselfconsistent()
    dev = build(case, dmet_mod.dmet)
    dev.bath_on_gpu = True
    e_dev = dev.doselfconsistent()
    assert dev.iterations == ref.iterations
    assert abs(e_dev - e_ref) <= 1e-8
    assert np.abs(dev.umat - ref.umat).max() <= 1e-7


def test_fit_imp_bath_jacobian_from_device_baths():
    """fitImpBath rows built from the device baths give the Jacobian of the host-bath rows (oracle), block by
    block up to the sign of the bath columns."""
    n, num_pairs, sizes = 48, 20, [4] * 6
    h0, imp_start, imp_idx = random_problem(n, num_pairs, sizes, seed=77)
    plan = plan_for(h0, num_pairs, imp_start, imp_idx)
    num_bath, occ, bath = plan.construct_baths(imp_start, imp_idx)
    one = plan.one_rdm()
    plan.close()
    clusters = []
    for f in range(len(sizes)):
        v = np.zeros(n, dtype=int)
        v[imp_idx[imp_start[f]:imp_start[f + 1]]] = 1
        clusters.append(v)
    rows_dev, rows_ref, signs = [], [], []
    for f, m in enumerate(sizes):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        _, loc2dmet, _ = host_bath(one, idx, m)
        got = bath[imp_start[f]:imp_start[f] + m].T
        sg = np.sign(np.sum(got * loc2dmet[:, m:2 * m], axis=0))
        rows_dev.append(np.concatenate([np.eye(n)[idx], got.T]))
        rows_ref.append(loc2dmet[:, :2 * m].T)
        signs.append(np.concatenate([np.ones(m), sg]))
    row_start = np.concatenate([[0], np.cumsum([r.shape[0] for r in rows_dev])]).astype(np.int32)
    # perturbations: E_ij + E_ji inside each impurity (the non-TI list of src/dmet.py:118-159, by site index)
    hs, hr, hc = [0], [], []
    for f, m in enumerate(sizes):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        for a in range(m):
            for b in range(a, m):
                if a == b:
                    hr.append(idx[a]); hc.append(idx[a])
                else:
                    hr += [idx[a], idx[b]]; hc += [idx[b], idx[a]]
                hs.append(len(hr))
    hs, hr, hc = (np.asarray(x, dtype=np.int32) for x in (hs, hr, hc))
    gpu = _lib.Plan(n, num_pairs, hs, hr, hc, row_start, np.concatenate(rows_dev), None)
    gpu.set_h0(h0)
    gpu.compute()
    jac = gpu.jacobian()
    c, e = gpu.orbitals(), gpu.eigenvalues()
    gpu.close()
    ref = ro.jacobian_from_orbitals(c, e, num_pairs, hs, hr, hc, row_start, np.concatenate(rows_ref), None)
    off = 0
    for f, sg in enumerate(signs):
        nb = sg.size
        scale = np.outer(sg, sg).reshape(-1, order="F")          # residual (a, b) of the block, column-major
        assert np.abs(jac[off:off + nb * nb] * scale[:, None] - ref[off:off + nb * nb]).max() <= 1e-10
        off += nb * nb


@pytest.mark.parametrize("l,u", [(8, 4.0), (24, 2.0)])
def test_embedded_hubbard_hamiltonians_match_host_transforms(l, u):
    """qcdmet_plan_embed_hubbard against localintegrals_hubbard.dmet_oei / dmet_fock / dmet_tei evaluated on the host
    with the SAME (device) bath orbitals: every fragment of the l x l lattice, 2 x 2 impurities."""
    n = l * l
    hop = wl.hubbard2d_hopping(l, l)
    ints = lh.localintegrals_hubbard(hop, u, n)
    w = wl.hubbard24(l=l)
    imp_start, imp_idx = w.row_start, w.row_idx
    plan = plan_for(ints.loc_rhf_fock(), n // 2, imp_start, imp_idx)
    plan.embed_hubbard(hop, u, imp_start[:2], imp_idx[:4])            # warm-up
    t0 = time.perf_counter()
    emb = plan.embed_hubbard(hop, u, imp_start, imp_idx)
    t_dev = time.perf_counter() - t0
    one = plan.one_rdm()
    nfrag = len(imp_start) - 1
    t_host = 0.0
    for f in ([0, 7, nfrag - 1] if l > 8 else range(nfrag)):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        nb = int(emb["num_bath"][f])
        bath = emb["bath"][imp_start[f]:imp_start[f] + nb].T
        active = np.concatenate([np.eye(n)[:, idx], bath], axis=1)
        occ = emb["occ"][imp_start[f]:imp_start[f] + nb]
        env = np.setdiff1d(np.arange(n), idx)
        core = np.zeros((n, n))
        core[np.ix_(env, env)] = (one - (bath * occ) @ bath.T)[np.ix_(env, env)]
        t0 = time.perf_counter()
        oei, fock, tei = ints.dmet_oei(active, 4 + nb), ints.dmet_fock(active, 4 + nb, core), ints.dmet_tei(active, 4 + nb)
        t_host += time.perf_counter() - t0
        assert np.abs(emb["oei"][f] - oei).max() <= 1e-11
        assert np.abs(emb["fock"][f] - fock).max() <= 1e-11
        assert np.abs(emb["tei"][f] - tei).max() <= 1e-11
        assert abs(emb["core_electrons"][f] - np.trace(core)) <= 1e-9
        assert int(round(n - emb["core_electrons"][f])) == 8
    print(f"\nembedding of {nfrag} fragments (N = {n}): device {t_dev * 1e3:.2f} ms for all (baths included); host transforms "
          f"{t_host * 1e3:.1f} ms for the checked fragments alone")
    plan.close()


@pytest.mark.parametrize("case", ["chain12_U4_imp_only", "chain12_U4_TI_bath"])
def test_dmet_loop_with_device_embedding(case):
    """The whole loop with baths AND embedded integrals from the device: same converged energy and umat."""
    ref = build(case, dmet_mod.dmet)
    e_ref = ref.doselfconsistent()
    dev = build(case, dmet_mod.dmet)
    dev.embed_on_gpu = True
    e_dev = dev.doselfconsistent()
    assert dev.iterations == ref.iterations
    assert abs(e_dev - e_ref) <= 1e-8
    assert np.abs(dev.umat - ref.umat).max() <= 1e-7


def oracle_embedding_problem():
    """Disordered chain: no degenerate occupations, so every bath orbital is unique up to its sign."""
    n, num_pairs, u = 20, 10, 3.0
    rng = np.random.default_rng(5)
    noise = 0.3 * rng.standard_normal((n, n))
    hop = lh.hubbard_chain(n) + 0.5 * (noise + noise.T)
    h0 = hop + np.diag(0.5 * rng.standard_normal(n))
    imp_start = np.array([0, 2, 7], dtype=np.int32)
    imp_idx = np.array([3, 4, 10, 11, 12, 15, 17], dtype=np.int32)
    return n, num_pairs, u, hop, h0, imp_start, imp_idx


def check_embedding_against_bath_oracle(emb, one, n, u, hop, imp_start, imp_idx):
    """emb: what Plan.embed_hubbard returns.  Everything is recomputed by oracle/bath_oracle.py from the 1-RDM alone
    (the reference's route: eigh of the environment block, src/qcdmethelper.py:124-168; src/dmet.py:204-226;
    src/localintegrals_hubbard.py:84-116) and compared after fixing the sign of each bath orbital."""
    from oracle import bath_oracle as bo
    for f in range(len(imp_start) - 1):
        idx = imp_idx[imp_start[f]:imp_start[f + 1]]
        m = idx.size
        flags = np.zeros(n, dtype=int)
        flags[idx] = 1
        nb, loc2dmet, core = bo.construct_bath(one, flags, m)
        assert nb == emb["num_bath"][f]
        act = m + nb
        got = emb["bath"][imp_start[f]:imp_start[f] + nb].T
        ref = loc2dmet[:, m:act]
        dots = np.sum(got * ref, axis=0)
        assert np.abs(np.abs(dots) - 1.0).max() <= 1e-10
        sv = np.concatenate([np.ones(m), np.sign(dots)])
        assert np.abs(got * sv[m:] - ref).max() <= 1e-9
        occ_ref = np.diag(ref.T @ one @ ref)
        assert np.abs(emb["occ"][imp_start[f]:imp_start[f] + nb] - occ_ref).max() <= 1e-10
        core_dm, _ = bo.core_density(loc2dmet, core)
        s2 = np.outer(sv, sv)
        assert np.abs(emb["oei"][f] * s2 - bo.embedded_oei(hop, loc2dmet, act)).max() <= 1e-10
        assert np.abs(emb["fock"][f] * s2 - bo.embedded_fock(hop, u, loc2dmet, act, core_dm)).max() <= 1e-10
        s4 = np.einsum("a,b,c,d->abcd", sv, sv, sv, sv)
        assert np.abs(emb["tei"][f] * s4 - bo.embedded_tei(u, loc2dmet, act)).max() <= 1e-10
        assert abs(emb["core_electrons"][f] - np.trace(core_dm)) <= 1e-9


def test_device_embedding_matches_bath_oracle_end_to_end():
    """Baths, occupations and embedded oei / fock / tei of the device path against the independent CPU checker that
    follows the reference's own route from the 1-RDM (no quantity of the device path enters the checker's side)."""
    n, num_pairs, u, hop, h0, imp_start, imp_idx = oracle_embedding_problem()
    plan = plan_for(h0, num_pairs, imp_start, imp_idx)
    emb = plan.embed_hubbard(hop, u, imp_start, imp_idx)
    one = plan.one_rdm()
    e, c = np.linalg.eigh(h0)
    assert np.abs(one - 2.0 * c[:, :num_pairs] @ c[:, :num_pairs].T).max() <= 1e-12
    check_embedding_against_bath_oracle(emb, one, n, u, hop, imp_start, imp_idx)
    plan.close()
