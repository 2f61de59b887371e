"""The reference's own test-suite (pydmrg/tests.py -> pydmrg/tests/asep/*.py), restated test by
test against the product's API -- same calls, same tolerances, seeded where the reference draws unseeded parameters --
and executed on the CPU over the emulated device entry points of tests/emu_device.py.

    arnoldiCheck.py            run_dmrg(alg='exact') vs alg='arnoldi'                         atol = rtol = 1e-4
    davidsonMultipleCheck.py   'exact' vs 'davidson', nStates = 2                             1e-4
    multipleBondDim.py         chi schedule [2,4,6]: fresh run vs restart from saved files    np.isclose default
    contractCheck.py           sweep energy vs <r|H|r>/<r|r>, <l|H|l>/<l|l>, <l|H|r>/<l|r>    np.isclose default
    currentCheck.py            current operator vs finite-difference dE/ds, ds = 1e-4         atol = 1e-3

(The 2-D tests need the 2-D ASEP builder, which is outside the hot path; tests.py:60-71 errors as written.)"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

N, MBD = 10, 10


@pytest.fixture
def E(monkeypatch):
    import emu_device
    emu_device.install(monkeypatch.setattr)
    from pydmrg_b200 import davidson
    davidson.release_workspaces()
    yield emu_device
    davidson.release_workspaces()


def ham_params(seed):
    return tuple(np.random.default_rng(seed).random(7))          # the reference draws np.random.rand() x 7, unseeded


def test_arnoldiCheck(E):
    """tests/asep/arnoldiCheck.py (tests.py:6-10)."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(N, ham_params(7))
    np.random.seed(0)
    E1 = dmrg.run_dmrg(mpo, mbd=MBD, alg="exact", ctx=ctx, recycle=False)[0]
    np.random.seed(0)
    E2 = dmrg.run_dmrg(mpo, mbd=MBD, alg="arnoldi", ctx=ctx, recycle=False, native=False)[0]
    assert np.isclose(E1, E2, atol=1e-4, rtol=1e-4)


def test_davidsonMultipleCheck(E):
    """tests/asep/davidsonMultipleCheck.py (tests.py:12-16): exact vs davidson with nStates = 2."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(N, ham_params(8))
    np.random.seed(1)
    E1 = dmrg.run_dmrg(mpo, mbd=MBD, alg="exact", nStates=2, ctx=ctx)[0]
    np.random.seed(1)
    E2 = dmrg.run_dmrg(mpo, mbd=MBD, alg="davidson", nStates=2, ctx=ctx, native=False)[0]
    assert np.isclose(E1, E2, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("alg", ["exact", "davidson", "arnoldi"])
def test_multipleBondDim(E, alg, tmp_path):
    """tests/asep/multipleBondDim.py (tests.py:18-29): the chi schedule [2,4,6] run fresh and restarted from its own
    saved files gives the same energies at every chi."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(N, ham_params(9))
    fname = str(tmp_path / ("tests_" + alg))
    kw = dict(mbd=np.array([2, 4, 6]), alg=alg, fname=fname, nStates=2, ctx=ctx, native=False)
    np.random.seed(2)
    Ea = dmrg.run_dmrg(mpo, **kw)[0]
    Eb = dmrg.run_dmrg(mpo, initGuess=fname, **kw)[0]
    assert Ea.shape == (3,) and Eb.shape == (3,)
    # the reference compares the last chi (tests.py:24-29 index [-1]); the solver floor of a restarted stage is 1e-7
    assert np.isclose(Ea[-1], Eb[-1])
    assert np.all(np.real(Ea[1:]) >= np.real(Ea[:-1]) - 1e-6)   # largest-Re(E) is variational in chi


def test_contractCheck(E, tmp_path):
    """tests/asep/contractCheck.py (tests.py:31-37): the sweep energy equals <r|H|r>/<r|r>, <l|H|l>/<l|l> and
    <l|H|r>/<l|r> contracted from the SAVED states (file names as arguments, as the reference's scripts do)."""
    from pydmrg_b200 import contract, dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(N, ham_params(10))
    fname = str(tmp_path / "current")
    np.random.seed(3)
    E0 = dmrg.run_dmrg(mpo, mbd=MBD, alg="davidson", nStates=1, fname=fname, calcLeftState=True, ctx=ctx, native=False)[0]
    r, l = fname + "_mbd0", fname + "_mbd0_left"
    e1 = contract.full_contract(mpo=mpo, mps=r, ctx=ctx) / contract.full_contract(mps=r, ctx=ctx)
    e2 = contract.full_contract(mpo=mpo, lmps=l, ctx=ctx) / contract.full_contract(lmps=l, ctx=ctx)
    e3 = contract.full_contract(mpo=mpo, mps=r, lmps=l, ctx=ctx) / contract.full_contract(mps=r, lmps=l, ctx=ctx)
    # H r = E r and H^dagger l = conj(E) l, so all three quotients equal E (tests.py:33-37 compares all of them)
    assert np.isclose(E0, e1) and np.isclose(E0, e2) and np.isclose(E0, e3)
    assert abs(e3 - E0) < 1e-6 * abs(E0)


def test_currentCheck(E, tmp_path):
    """tests/asep/currentCheck.py (tests.py:51-58): the current from the current MPO between the left and the right
    eigenstate equals the finite-difference derivative dE/ds (ds = 1e-4) to 1e-3."""
    from pydmrg_b200 import contract, dmrg, mpo as M
    ctx = E.Context("c128")
    prm = np.array(ham_params(11))
    ds = 1e-4
    fname = str(tmp_path / "current")
    np.random.seed(4)
    dmrg.run_dmrg(M.asep_mpo(N, tuple(prm)), mbd=MBD, alg="davidson", nStates=1, fname=fname, calcLeftState=True, ctx=ctx,
                  native=False, res_tol=1e-9)
    r, l = fname + "_mbd0", fname + "_mbd0_left"
    cur = contract.full_contract(mpo=M.asep_curr_mpo(N, tuple(prm)), mps=r, lmps=l, ctx=ctx) / contract.full_contract(mps=r, lmps=l, ctx=ctx)
    Es = []
    for sign in (-1.0, 1.0):
        p = prm.copy(); p[-1] += sign * ds
        np.random.seed(4)
        Es.append(dmrg.run_dmrg(M.asep_mpo(N, tuple(p)), mbd=MBD, alg="davidson", nStates=1, ctx=ctx, native=False, res_tol=1e-9)[0])
    dEds = (Es[1] - Es[0]) / (2 * ds)
    assert np.isclose(np.real(cur), np.real(dEds), atol=1e-3)
