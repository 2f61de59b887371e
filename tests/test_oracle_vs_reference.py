"""CPU, build container only: the oracle and the MPO data builders against the UNMODIFIED
reference imported live from /root/reference (skipped where that path is absent, e.g. on
the GPU box -- there the committed golden fixtures take over, tests/test_oracle.py)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import dmrg_oracle as orc
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present")
quiet = lambda: contextlib.redirect_stdout(io.StringIO())


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


@pytest.mark.parametrize("prm", [(0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5), (0.35, 0., 1., 0., 2. / 3., 0., -0.5),
                                 (0.2, 0.8, 0.1, 0.9, 0.4, 0.6, 0.7)])
@pytest.mark.parametrize("N", [2, 3, 6, 11])
def test_mpo_builders_equal_reference(ref, prm, N):
    from pydmrg_b200 import mpo as mine
    a = ref.asep.return_mpo(N, prm)
    b = mine.asep_mpo(N, prm)
    assert len(a) == len(b) == 1
    for Wa, Wb in zip(a[0], b[0]):
        assert Wa.shape == Wb.shape and np.array_equal(Wa, Wb)
    a = ref.asep.curr_mpo(N, prm)
    b = mine.asep_curr_mpo(N, prm)
    for Wa, Wb in zip(a[0], b[0]):
        assert Wa.shape == Wb.shape and np.array_equal(Wa, Wb)
    ap = ref.asep.return_mpo(N, prm, periodic=True)
    bp = mine.asep_mpo(N, prm, periodic=True)
    assert len(ap) == len(bp)
    for opa, opb in zip(ap, bp):
        for Wa, Wb in zip(opa, opb):
            assert (Wa is None and Wb is None) or np.array_equal(Wa, Wb)
    with quiet():
        ct = ref.mpo_tools.mpo_conj_trans(ref.asep.return_mpo(N, prm))
    for Wa, Wb in zip(ct[0], mine.mpo_conj_trans(mine.asep_mpo(N, prm))[0]):
        assert np.array_equal(Wa, Wb)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matvec_and_env_random(ref, seed):
    rng = np.random.default_rng(seed)
    d, Dl, Dr = 2, int(rng.integers(1, 12)), int(rng.integers(1, 12))
    W1, W2 = ref.asep.return_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
    w = 6
    L, R = crand(rng, Dl, w, Dl), crand(rng, Dr, w, Dr)
    x1, x2 = crand(rng, d, Dl, Dr), crand(rng, d, d, Dl, Dr)
    Hf, _ = ref.diag_tools.make_ham_func_oneSite([x1], [[W1]], [[L, R]], 0)
    assert np.allclose(Hf(x1.ravel()), orc.heff_onesite(x1.ravel(), [L], [W1], [R], x1.shape), rtol=0, atol=1e-12 * np.abs(x1).max() * Dl * Dr * 36)
    Hf2, _ = ref.diag_tools.make_ham_func_twoSite([np.zeros((d, Dl, 3)), np.zeros((d, 3, Dr))], [[W1, W2]], [[L, R]], 0)
    y2 = orc.heff_twosite(x2.ravel(), [L], [W1], [W2], [R], x2.shape)
    assert np.linalg.norm(Hf2(x2.ravel()) - y2) <= 1e-13 * np.linalg.norm(y2)
    Hd = ref.diag_tools.calc_ham_oneSite([x1], [[W1]], [[L, R]], 0)
    assert np.allclose(Hd, orc.heff_dense_onesite([L], [W1], [R], x1.shape))


def test_one_site_run_matches_reference(ref):
    prm = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)
    mpo = ref.asep.return_mpo(8, prm)
    np.random.seed(11)
    with quiet():
        r = ref.dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=3, returnEntSpec=True)
    np.random.seed(11)
    E, EE, EEs, _, _ = orc.run_dmrg_onesite(mpo, 8, tol=1e-10, max_iter=3)
    assert abs(E - r[0]) < 1e-8 * abs(r[0])
    assert abs(EE - r[1]) < 1e-7
    assert np.allclose(np.sort(EEs)[::-1][:4], np.sort(r[3])[::-1][:4], atol=1e-7)


def test_full_contract_matches_reference(ref):
    rng = np.random.default_rng(4)
    N = 5
    dims = [1, 2, 4, 4, 2, 1]
    mps = [crand(rng, 2, dims[i], dims[i + 1]) for i in range(N)]
    lmps = [crand(rng, 2, dims[i], dims[i + 1]) for i in range(N)]
    mpo = ref.asep.return_mpo(N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    with quiet():
        a = ref.contract.full_contract(mpo=mpo, mps=[mps], lmps=[lmps])
        b = ref.contract.full_contract(mps=[mps])
    assert abs(a - orc.full_contract(mpo, mps, lmps)) < 1e-12 * abs(a)
    assert abs(b - orc.full_contract(None, mps)) < 1e-12 * abs(b)


def test_saved_states_round_trip_between_reference_and_product(ref, tmp_path):
    """saved_states plug in unchanged (north-star): a state written by the reference's save_mps (HDF5 layout of
    tools/mps_tools.py:364-373, NPZ layout :355-362) is read by the product's load_mps and vice versa."""
    from pydmrg_b200 import mps as mps_tools
    rng = np.random.default_rng(0)
    mpsL = [[rng.standard_normal(s) + 1j * rng.standard_normal(s) for s in ((2, 1, 2), (2, 2, 4), (2, 4, 2), (2, 2, 1))]
            for _ in range(2)]
    for fmt in ("hdf5", "npz"):
        a, b = str(tmp_path / ("ref_" + fmt)), str(tmp_path / ("prod_" + fmt))
        ref.mps_tools.save_mps(mpsL, a, gaugeSite=2, fformat=fmt)
        got, gs = mps_tools.load_mps(a, fformat=fmt)
        assert int(gs) == 2 and all(np.array_equal(x, y) for s in range(2) for x, y in zip(got[s], mpsL[s]))
        mps_tools.save_mps(mpsL, b, gaugeSite=3, fformat=fmt)
        got, gs = ref.mps_tools.load_mps(b, fformat=fmt)
        assert int(gs) == 3 and all(np.array_equal(x, y) for s in range(2) for x, y in zip(got[s], mpsL[s]))


def test_orthonormalize_states_equals_reference(ref):
    """tools/mps_tools.py:392-430 (used by full_contract(orth=True) and the scan scripts' post-processing)."""
    from pydmrg_b200 import mps as mps_tools
    rng = np.random.default_rng(2)
    shapes = ((2, 1, 2), (2, 2, 4), (2, 4, 2), (2, 2, 1))
    shared = [crand(rng, *s) for s in shapes]
    mpsL = [[t.copy() for t in shared] for _ in range(2)]
    for st in range(2):
        mpsL[st][2] = crand(rng, *shapes[2])                       # only the centre tensors differ between the states
    with quiet():
        a = ref.mps_tools.orthonormalize_states([[t.copy() for t in M] for M in mpsL], gSite=2)
    b = mps_tools.orthonormalize_states([[t.copy() for t in M] for M in mpsL], gSite=2)
    for st in range(2):
        for x, y in zip(a[st], b[st]):
            assert np.array_equal(x, y)
    c0, c1 = b[0][2].ravel(), b[1][2].ravel()
    assert abs(np.vdot(c0, c1)) < 1e-13 and abs(np.vdot(c0, c0) - 1) < 1e-13
