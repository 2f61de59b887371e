"""CPU: the NumPy oracle against the committed golden fixtures (outputs of the unmodified
reference, tests/golden/make_golden.py) -- pins the oracle without needing /root/reference."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc

CASES = ["a", "b", "c", "d"]
TOL = 1e-13  # pure sums of products


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


@pytest.mark.parametrize("mode", ["blas", "einsum"])
@pytest.mark.parametrize("c", CASES)
def test_heff_onesite(golden, c, mode):
    g = golden["kernels"]
    x = g[c + "_x1"]
    y = orc.heff_onesite(x.ravel(), [g[c + "_L"]], [g[c + "_W1"]], [g[c + "_R"]], x.shape, mode)
    assert rel(y, g[c + "_y1"]) < TOL


@pytest.mark.parametrize("mode", ["blas", "einsum"])
@pytest.mark.parametrize("c", CASES)
def test_heff_twosite(golden, c, mode):
    g = golden["kernels"]
    x = g[c + "_x2"]
    y = orc.heff_twosite(x.ravel(), [g[c + "_L"]], [g[c + "_W1"]], [g[c + "_W2"]], [g[c + "_R"]], x.shape, mode)
    assert rel(y, g[c + "_y2"]) < TOL


@pytest.mark.parametrize("mode", ["blas", "einsum"])
@pytest.mark.parametrize("c", CASES)
def test_env_updates(golden, c, mode):
    g = golden["kernels"]
    A, W, L, R = g[c + "_A"], g[c + "_W1"], g[c + "_L"], g[c + "_R"]
    assert rel(orc.env_update_right(A, W, L, None, mode), g[c + "_envR_out"]) < TOL
    assert rel(orc.env_update_left(A, W, R, None, mode), g[c + "_envL_out"]) < TOL
    assert rel(orc.env_update_right(A, W, L, g[c + "_Abra"], mode), g[c + "_envR_bra_out"]) < TOL
    assert rel(orc.env_update_left(A, None, R, g[c + "_Abra"], mode), g[c + "_envL_none_out"]) < TOL


@pytest.mark.parametrize("c", CASES)
def test_dense_matches_matvec(golden, c):
    g = golden["kernels"]
    x = g[c + "_x1"]
    H = orc.heff_dense_onesite([g[c + "_L"]], [g[c + "_W1"]], [g[c + "_R"]], x.shape)
    assert rel(-H @ x.ravel(), g[c + "_y1"]) < TOL
    x2 = g[c + "_x2"]
    H2 = orc.heff_dense_twosite([g[c + "_L"]], [g[c + "_W1"]], [g[c + "_W2"]], [g[c + "_R"]], x2.shape)
    assert rel(-H2 @ x2.ravel(), g[c + "_y2"]) < TOL


@pytest.mark.parametrize("c", CASES)
def test_svd_truncate(golden, c):
    g = golden["kernels"]
    x2 = g[c + "_x2"]
    mbd = int(g[c + "_svd_mbd"])
    A, B, S, EE, EEs = orc.svd_truncate_twosite(x2.ravel(), x2.shape, mbd)
    assert np.allclose(S, g[c + "_svd_S"], rtol=1e-12, atol=0)
    assert abs(EE - g[c + "_svd_EE"]) < 1e-12
    assert np.allclose(EEs, g[c + "_svd_EEs"], rtol=1e-10, atol=1e-14)
    # tensors only up to gauge: compare projectors
    Ar, Br = g[c + "_svd_A"], g[c + "_svd_B"]
    assert A.shape == Ar.shape and B.shape == Br.shape
    d, Dl, k = A.shape
    Pa = lambda T: (lambda m: m @ m.conj().T)(T.swapaxes(0, 1).reshape(Dl * d, k))
    assert rel(Pa(A), Pa(Ar)) < 1e-10


def test_davidson_restatement(golden):
    g = golden["davidson"]
    L, R, W, x0 = g["L"], g["R"], g["W"], g["x0"]
    stats = {}
    mv = lambda x: orc.heff_onesite(x, [L], [W], [R], x0.shape)
    e, xs = orc.davidson_nosym(mv, [x0.ravel()], nroots=1, stats=stats)
    # same algorithm, same arithmetic order up to BLAS vs einsum rounding
    assert abs(e[0] - g["eval"]) < 1e-9 * abs(g["eval"])
    assert abs(stats["n_matvec"] - int(g["n_matvec"])) <= max(3, int(0.05 * int(g["n_matvec"])))
    # eigenvector up to phase
    v, vr = xs[0], g["evec"]
    assert abs(abs(np.vdot(v, vr)) / (np.linalg.norm(v) * np.linalg.norm(vr)) - 1) < 1e-10
    # Ritz value vs dense eigenvalue: the reference's own floor (residual <= 1e-7)
    assert abs(e[0] - g["exact_min_real"]) < 1e-6


def test_config0_tasep(golden):
    """BASELINE.json configs[0]: TASEP N=10 chi=10 s=-0.5, one-site finite DMRG."""
    from oracle.dmrg_oracle import run_dmrg_onesite
    g = golden["runs"]
    mpo = _asep_mpo(10, g["cfg0_params"])
    np.random.seed(0)
    E, EE, EEs, _, _ = run_dmrg_onesite(mpo, 10, tol=1e-8, max_iter=10)
    assert abs(E - g["cfg0_davidson_E"]) < 1e-8 * abs(g["cfg0_davidson_E"])
    assert abs(E - g["cfg0_exact_E"]) < 2e-7
    np.random.seed(0)
    E2, EE2, _, _, _ = run_dmrg_onesite(mpo, 10, tol=1e-8, max_iter=10, alg="exact")
    assert abs(E2 - g["cfg0_exact_E"]) < 1e-10 * abs(g["cfg0_exact_E"])


def test_small_runs_vs_ed(golden):
    g = golden["runs"]
    for tag in ("tasep6", "asep6"):
        mpo = _asep_mpo(6, g[tag + "_params"])
        np.random.seed(3)
        E, EE, EEs, mps, _ = orc.run_dmrg_onesite(mpo, 8, tol=1e-10, max_iter=6)
        assert abs(E - g[tag + "_dmrg_E"]) < 2e-8
        assert abs(E - g[tag + "_ed_top"]) < 5e-8
        assert abs(EE - g[tag + "_dmrg_EE"]) < 1e-6
        # two-site finite sweep (composition) has the same fixed point
        np.random.seed(3)
        E2 = orc.run_dmrg_twosite(mpo, 8, n_sweeps=3)[0]
        assert abs(E2 - g[tag + "_ed_top"]) < 5e-8


def test_full_contract(golden):
    g = golden["runs"]
    for tag in ("tasep6", "asep6"):
        mps = [g[tag + "_mps%d" % i] for i in range(6)]
        mpo = _asep_mpo(6, g[tag + "_params"])
        den = orc.full_contract(None, mps)
        e = orc.full_contract(mpo, mps) / den
        assert abs(e - g[tag + "_energy_rr"]) < 1e-12
        cur = orc.full_contract(_asep_curr_mpo(6, g[tag + "_params"]), mps) / den
        assert abs(cur - g[tag + "_current_rr"]) < 1e-12


def test_idmrg(golden):
    g = golden["runs"]
    mpo4 = _asep_mpo(4, (0.2, 0.8, 0.1, 0.9, 0.4, 0.6, -0.5))
    np.random.seed(5)
    E, EE, EEs, S, N, _, _ = orc.run_idmrg(mpo4, 10, max_iter=20)
    assert abs(E - g["idmrg_E"]) < 1e-6 * abs(g["idmrg_E"])
    assert abs(EE - g["idmrg_EE"]) < 1e-6


def test_heisenberg_mpo_two_sites():
    mpo = orc.heisenberg_mpo(2, sign=1.0)
    H = np.einsum("abij,bckl->ikjl", mpo[0][0], mpo[0][1]).reshape(4, 4)
    ev = np.sort(np.linalg.eigvalsh(H))
    assert np.allclose(ev, [-0.75, 0.25, 0.25, 0.25])


# ---- model builders shared by tests (data only) ---------------------------------------
def _asep_mpo(N, prm):
    from pydmrg_b200.mpo import asep_mpo
    return asep_mpo(N, tuple(float(v) for v in prm))


def _asep_curr_mpo(N, prm):
    from pydmrg_b200.mpo import asep_curr_mpo
    return asep_curr_mpo(N, tuple(float(v) for v in prm))


# ---------------------------------------------------------------------------------------
# warm-started truncation (oracle/recycle_oracle.py): the algorithmic claim, checked on the CPU
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,N,chi", [("heisenberg", 16, 12), ("asep", 14, 10)])
def test_recycled_truncation_reaches_the_exact_svd_fixed_point(model, N, chi):
    """Two-site sweeps that warm-start the truncation from each bond's previous bases (one ordered subspace
    step + a Rayleigh-Ritz window at the cut) converge to the same energy (1e-10 relative) and centre-bond
    entropy (1e-8) as sweeps that run the reference's fresh SVD at every bond, with chi truncating."""
    from oracle import recycle_oracle as R
    from oracle import dmrg_oracle as orc
    if model == "heisenberg":
        mpo = orc.heisenberg_mpo(N)
    else:
        from pydmrg_b200 import mpo as M
        mpo = M.asep_mpo(N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    st = {}
    exact = R.run_twosite(mpo, chi, 6, recycle=False)
    rec = R.run_twosite(mpo, chi, 6, recycle=True, stats=st)
    assert st.get("n_recycled", 0) > 0
    (E0, EE0), (E1, EE1) = exact[-1], rec[-1]
    assert abs(E1 - E0) < 1e-10 * abs(E0)
    assert abs(EE1 - EE0) < 1e-8


def test_recycle_basis_oracle_properties():
    """The restated subspace step: orthonormal rows, C = P X, window rows orthogonal and sorted, exact when
    started from the exact bases."""
    from oracle import recycle_oracle as R
    rng = np.random.default_rng(2)
    a, c, k, p, b = 48, 40, 16, 4, 4
    U, _ = np.linalg.qr(rng.standard_normal((a, a)) + 1j * rng.standard_normal((a, a)))
    V, _ = np.linalg.qr(rng.standard_normal((c, c)) + 1j * rng.standard_normal((c, c)))
    s = np.exp(-0.3 * np.arange(c))
    X = (U[:, :c] * s) @ V.conj().T
    W = V[:, :k + p].conj().T
    Pt, Ct, S = R.recycle_basis(X, X.T, W, k, b)
    assert np.max(np.abs(Pt @ Pt.conj().T - np.eye(k + p))) < 1e-13
    assert np.max(np.abs(Ct - Pt @ X)) < 1e-13
    assert np.max(np.abs(S - s[:k + p])) < 1e-12
    G = Ct[k - b:] @ Ct[k - b:].conj().T
    assert np.max(np.abs(G - np.diag(np.diag(G)))) < 1e-14


# ---------------------------------------------------------------------------------------
# multi-state targeting (nStates = 2)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["asep6", "asep8"])
def test_two_states_onesite_matches_reference_run(golden, tag):
    """oracle/multistate_oracle.py (state-averaged RDM, dmrg.py:56-66; two-root Davidson) against the
    unmodified reference's run_dmrg(nStates=2) on the same seed -- the restatement follows the reference
    operation by operation, so it lands on the reference's numbers far inside its 1e-7 solver floor -- and
    against dense ED for E and the gap."""
    from oracle import multistate_oracle as ms
    g = golden["multistate"]
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    mpo = _asep_mpo(N, g["params"])
    np.random.seed(seed)
    E, EE, gap, EEs, mpsL, env, Eall = ms.run_dmrg_onesite_states(mpo, mbd, 2, tol=1e-10, max_iter=6)
    assert abs(E - g[tag + "_E"]) < 1e-10 * abs(g[tag + "_E"])
    assert abs(gap - g[tag + "_gap"]) < 1e-9
    assert abs(EE - g[tag + "_EE"]) < 1e-8
    assert np.max(np.abs(np.sort(np.real(EEs)) - np.sort(np.real(g[tag + "_EEs"])))) < 1e-8
    ed = g[tag + "_ed"]
    assert abs(E - ed[0]) < 1e-7 * abs(ed[0]) and abs(gap - (ed[0] - ed[1])) < 1e-7
    # every state carries the same isometries away from the gauge site
    for site in range(N):
        if site != N // 2 + 1:
            assert np.array_equal(mpsL[0][site], mpsL[1][site])


@pytest.mark.parametrize("tag", ["asep6", "asep8"])
def test_two_states_twosite_restatement(golden, tag):
    """Two-site sweeps targeting two states (the product's extension, restated in NumPy): same two leading
    eigenvalues as ED / the reference's one-site run when chi is exact, and the centre-bond entropy of the
    target state equals that of a single-state two-site run."""
    from oracle import multistate_oracle as ms
    g = golden["multistate"]
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    mpo = _asep_mpo(N, g["params"])
    np.random.seed(seed)
    E, EE, EEs, S, mpsL, env = ms.run_dmrg_twosite_states(mpo, mbd, 2, n_sweeps=4)
    ed = g[tag + "_ed"]
    assert abs(E[0] - ed[0]) < 1e-7 * abs(ed[0]) and abs(E[1] - ed[1]) < 1e-7 * abs(ed[0])
    assert abs((E[0] - E[1]) - g[tag + "_gap"]) < 1e-7
    np.random.seed(seed)
    E1, EE1 = orc.run_dmrg_twosite(mpo, mbd, n_sweeps=4)[:2]
    assert abs(E1 - E[0]) < 1e-7 * abs(E1) and abs(EE1 - EE) < 1e-6


def test_twosite_state_averaged_truncation_properties():
    """rdm_truncate_twosite_states: shared isometry on the side left behind, exact projection of every state,
    optimal for the state-averaged density matrix."""
    from oracle import multistate_oracle as ms
    rng = np.random.default_rng(0)
    shape = (2, 2, 6, 5)
    psis = [rng.standard_normal(120) + 1j * rng.standard_normal(120) for _ in range(2)]
    for direction in ("right", "left"):
        As, Bs, Sn, EE, EEs = ms.rdm_truncate_twosite_states(psis, shape, 7, direction)
        for s in range(2):
            m = psis[s].reshape(shape).transpose(2, 0, 1, 3).reshape(12, 10)
            Am = As[s].transpose(1, 0, 2).reshape(12, 7)
            Bm = Bs[s].transpose(1, 0, 2).reshape(7, 10)
            iso = Am if direction == "right" else Bm.conj().T
            assert np.allclose(iso.conj().T @ iso, np.eye(7), atol=1e-13)
            proj = iso @ iso.conj().T @ m if direction == "right" else m @ iso @ iso.conj().T
            assert np.allclose(Am @ Bm, proj, atol=1e-13)
        assert As[0] is As[1] if direction == "right" else Bs[0] is Bs[1]


# ---------------------------------------------------------------------------------------
# left + right eigenstates and observables <l|O|r>/<l|r> (BASELINE configs[2]); chi schedules through saved states
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["tasep8", "asep8"])
def test_left_and_right_eigenstates_match_reference_run(golden, tag):
    """The reference's run_dmrg(calcLeftState=True) -- a second DMRG on mpo_conj_trans(mpo), dmrg.py:338,448-488 --
    against the oracle's two runs: eigenvalue, both entropies, and the current / energy <l|O|r>/<l|r>
    (tools/contract.py:5-44) that the scan scripts report.  The reference's left run starts from its converged
    right state (aliased lists, dmrg.py:371), the oracle's from the seeded start; converged quantities agree to
    the reference's solver floor (||r|| <= 1e-7)."""
    from pydmrg_b200.mpo import mpo_conj_trans
    g = golden["leftright"]
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    prm = g[tag + "_params"]
    mpo = _asep_mpo(N, prm)
    np.random.seed(seed)
    E, EE, EEs, mps_r, _ = orc.run_dmrg_onesite(mpo, mbd, tol=1e-10, max_iter=6)
    np.random.seed(seed)
    El, EEl, EEsl, mps_l, _ = orc.run_dmrg_onesite(mpo_conj_trans(mpo), mbd, tol=1e-10, max_iter=6)
    scale = abs(g[tag + "_E"])
    assert abs(E - g[tag + "_E"]) < 1e-7 * scale and abs(E - g[tag + "_ed_top"]) < 1e-7 * scale
    assert abs(El - np.conj(E)) < 1e-7 * scale
    assert abs(EE - g[tag + "_EE"]) < 1e-6 and abs(EEl - g[tag + "_EEl"]) < 1e-6
    den = orc.full_contract(None, mps_r, mps_l)
    cur = orc.full_contract(_asep_curr_mpo(N, prm), mps_r, mps_l) / den
    ene = orc.full_contract(mpo, mps_r, mps_l) / den
    assert abs(cur - g[tag + "_current_lr"]) < 1e-6 * max(1.0, abs(g[tag + "_current_lr"]))
    assert abs(ene - g[tag + "_energy_lr"]) < 1e-7 * scale and abs(ene - g[tag + "_ed_top"]) < 1e-7 * scale


def test_chi_schedule_through_saved_states(golden):
    """run_dmrg(mbd=[4, 8, 16], fname=...): every stage loads the previous state (centre on gaugeSiteSave),
    zero-pads it (mps_tools.py:244-279) and sweeps from the loaded gauge site (dmrg.py:352-358, 248-259)."""
    g = golden["leftright"]
    N, seed = (int(v) for v in g["sched_cfg"])
    mpo = _asep_mpo(N, g["asep8_params"])
    np.random.seed(seed)
    mps, gs = None, 0
    for i, chi in enumerate((4, 8, 16)):
        if mps is not None:
            mps = orc.increase_mbd(mps, chi)
        E, EE, EEs, mps, _ = orc.run_dmrg_onesite(mpo, chi, tol=1e-10, max_iter=4, mps=mps, gauge_site_load=gs)
        gs = N // 2 + 1
        assert abs(E - g["sched_E"][i]) < 2e-7 * abs(g["sched_E"][i])
        assert abs(EE - g["sched_EE"][i]) < 1e-5


def test_increase_mbd_equals_product_builder():
    from pydmrg_b200 import mps as mps_tools
    for N in (7, 8):
        np.random.seed(N)
        a = orc.random_mps(N, 4)
        b = [t.copy() for t in a]
        pa = orc.increase_mbd(a, 16)
        pb = mps_tools.increase_mbd(b, 16)
        assert all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(pa, pb))
