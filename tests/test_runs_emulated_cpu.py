"""The product's run API (pydmrg_b200.dmrg / idmrg / eigs / env / truncate / contract / mps) executed END TO END
on the CPU: the device entry points are replaced by the torch-CPU emulations of tests/emu_device.py, everything above
them -- sweeps, steps, local-solver driver, Python Davidson loop, truncation bookkeeping, state averaging, chi
schedules, save / load, result packing -- is the product's own code.  The runs are compared with the fixtures the
UNMODIFIED reference produced (tests/golden/*.npz) and with dense ED.

What this covers: the host layer, every round, without a GPU.  What it does not: the CUDA kernels (``-m gpu``)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

PRM = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)


@pytest.fixture
def E(monkeypatch):
    import emu_device
    emu_device.install(monkeypatch.setattr)
    from pydmrg_b200 import davidson
    davidson.release_workspaces()
    yield emu_device
    davidson.release_workspaces()


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


# ---- the emulations themselves against the reference's outputs ----------------------------------------------------
@pytest.mark.parametrize("c", ["a", "b", "c", "d"])
def test_emulated_entry_points_match_reference_kernels(E, golden, c):
    from pydmrg_b200 import device as D
    g = golden["kernels"]
    ctx = E.Context("c128")
    L, R, W1, W2 = g[c + "_L"], g[c + "_R"], g[c + "_W1"], g[c + "_W2"]
    x1, x2 = g[c + "_x1"], g[c + "_x2"]
    d, Dl, Dr = x1.shape
    plan = D.HeffPlan(1, d, Dl, Dr, [(ctx.dev(L), D.combine_onesite(W1, L.shape[1], R.shape[1], d, np.complex128), ctx.dev(R))], None)
    y = torch.empty(x1.size, dtype=torch.complex128)
    plan.apply(ctx.dev(x1.ravel()), y, -1.0)
    assert rel(y.numpy(), g[c + "_y1"]) < 1e-13
    plan = D.HeffPlan(1, d * d, Dl, Dr, [(ctx.dev(L), D.combine_twosite(W1, W2, np.complex128), ctx.dev(R))], None)
    y = torch.empty(x2.size, dtype=torch.complex128)
    plan.apply(ctx.dev(x2.ravel()), y, -1.0)
    assert rel(y.numpy(), g[c + "_y2"]) < 1e-13
    # ket-range and block-mask restrictions add up to the whole operator (the two sharded modes)
    parts = torch.zeros_like(y)
    for k0, kn in ((0, Dl // 2), (Dl // 2, Dl - Dl // 2)):
        yp = torch.empty_like(y)
        D.HeffPlan(1, d * d, Dl, Dr, [(ctx.dev(L), D.combine_twosite(W1, W2, np.complex128), ctx.dev(R))], None, k_range=(k0, kn)).apply(
            ctx.dev(x2.ravel()), yp, -1.0)
        parts += yp
    assert rel(parts.numpy(), g[c + "_y2"]) < 1e-13
    A = ctx.dev(g[c + "_A"])
    assert rel(D.env_update_right(A, W1, ctx.dev(L), None, None).numpy(), g[c + "_envR_out"]) < 1e-13
    assert rel(D.env_update_left(A, W1, ctx.dev(R), None, None).numpy(), g[c + "_envL_out"]) < 1e-13
    assert rel(D.env_update_right(A, W1, ctx.dev(L), ctx.dev(g[c + "_Abra"]), None).numpy(), g[c + "_envR_bra_out"]) < 1e-13
    assert rel(D.env_update_left(A, None, ctx.dev(R), ctx.dev(g[c + "_Abra"]), None).numpy(), g[c + "_envL_none_out"]) < 1e-13


# ---- run_dmrg ------------------------------------------------------------------------------------------------------
def test_config0_one_site_run(E, golden):
    """BASELINE configs[0]: TASEP N=10 chi=10 s=-0.5, the reference's own one-site algorithm through the product's
    run_dmrg: Davidson against the reference's Davidson run (its 1e-7 solver floor), 'exact' against its exact run."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["runs"]
    mpo = M.asep_mpo(10, tuple(float(v) for v in g["cfg0_params"]))
    ctx = E.Context("c128")
    np.random.seed(0)
    out = dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10, ctx=ctx, native=False, returnEntSpec=True)
    assert abs(out[0] - g["cfg0_davidson_E"]) < 1e-7 * abs(g["cfg0_davidson_E"])
    assert abs(out[0] - g["cfg0_exact_E"]) < 1e-7 * abs(g["cfg0_exact_E"])
    assert abs(out[1] - g["cfg0_davidson_EE"]) < 1e-6
    np.random.seed(0)
    out = dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10, alg="exact", ctx=ctx, recycle=False)
    assert abs(out[0] - g["cfg0_exact_E"]) < 1e-10 * abs(g["cfg0_exact_E"])
    assert abs(out[1] - g["cfg0_exact_EE"]) < 1e-8


@pytest.mark.parametrize("twoSite", [False, True])
def test_small_runs_vs_ed_and_reference(E, golden, twoSite):
    from pydmrg_b200 import contract, dmrg, mpo as M
    g = golden["runs"]
    ctx = E.Context("c128")
    for tag in ("tasep6", "asep6"):
        prm = tuple(float(v) for v in g[tag + "_params"])
        mpo = M.asep_mpo(6, prm)
        np.random.seed(3)
        out = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=6, ctx=ctx, native=False, twoSite=twoSite, res_tol=1e-11,
                            returnEntSpec=True, returnState=True)
        E0, EE, gap, EEs, mps = out
        assert abs(E0 - g[tag + "_ed_top"]) < 1e-9 * abs(g[tag + "_ed_top"])
        assert abs(E0 - g[tag + "_dmrg_E"]) < 1e-7 * abs(g[tag + "_dmrg_E"])
        if not twoSite:
            assert abs(EE - g[tag + "_dmrg_EE"]) < 1e-6
        # observables on the returned state (tests/asep/currentCheck.py:29-44)
        cur = contract.full_contract(mpo=M.asep_curr_mpo(6, prm), mps=mps, ctx=ctx) / contract.full_contract(mps=mps, ctx=ctx)
        assert abs(cur - g[tag + "_current_rr"]) < 1e-6 * max(1.0, abs(g[tag + "_current_rr"]))


@pytest.mark.parametrize("twoSite", [False, True])
@pytest.mark.parametrize("tag", ["asep6", "asep8"])
def test_two_states_runs(E, golden, tag, twoSite):
    """nStates = 2 (every production script of the reference): E and the gap against the reference's run and ED,
    with the reference's one-site algorithm and with two-site sweeps (state-averaged truncation)."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["multistate"]
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    mpo = M.asep_mpo(N, tuple(float(v) for v in g["params"]))
    ctx = E.Context("c128")
    np.random.seed(seed)
    out = dmrg.run_dmrg(mpo, mbd=mbd, tol=1e-10, maxIter=6, nStates=2, ctx=ctx, native=False, twoSite=twoSite, res_tol=1e-10)
    ed = g[tag + "_ed"]
    assert abs(out[0] - ed[0]) < 1e-8 * abs(ed[0])
    assert abs(out[2] - (ed[0] - ed[1])) < 1e-7 * abs(ed[0])
    assert abs(out[2] - g[tag + "_gap"]) < 1e-7 * abs(ed[0])
    if not twoSite:
        assert abs(out[1] - g[tag + "_EE"]) < 1e-6


@pytest.mark.parametrize("tag", ["tasep8", "asep8"])
def test_left_and_right_eigenstates_run(E, golden, tag):
    """run_dmrg(calcLeftState=True) against the reference's run (tests/golden/leftright.npz): eigenvalue, both
    entropies, current and energy <l|O|r>/<l|r> through the product's full_contract."""
    from pydmrg_b200 import contract, dmrg, mpo as M
    g = golden["leftright"]
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    prm = tuple(float(v) for v in g[tag + "_params"])
    mpo = M.asep_mpo(N, prm)
    ctx = E.Context("c128")
    np.random.seed(seed)
    E0, EE, gap, mps = dmrg.run_dmrg(mpo, mbd=mbd, tol=1e-12, maxIter=6, res_tol=1e-11, calcLeftState=True, returnState=True,
                                     ctx=ctx, native=False)
    r, l = mps
    scale = abs(g[tag + "_ed_top"])
    assert abs(E0 - g[tag + "_ed_top"]) < 1e-9 * scale and abs(E0 - g[tag + "_E"]) < 1e-7 * scale
    assert abs(EE[0] - g[tag + "_EE"]) < 1e-6 and abs(EE[1] - g[tag + "_EEl"]) < 1e-6
    den = contract.full_contract(mps=r, lmps=l, ctx=ctx)
    cur = contract.full_contract(mpo=M.asep_curr_mpo(N, prm), mps=r, lmps=l, ctx=ctx) / den
    ene = contract.full_contract(mpo=mpo, mps=r, lmps=l, ctx=ctx) / den
    assert abs(cur - g[tag + "_current_lr"]) < 1e-6 * max(1.0, abs(g[tag + "_current_lr"]))
    assert abs(ene - g[tag + "_ed_top"]) < 1e-8 * scale


def test_chi_schedule_with_saved_states(E, golden, tmp_path):
    """mbd=[4, 8, 16] continued through files (dmrg.py:352-358): arrays over chi in the reference's return packing."""
    from pydmrg_b200 import dmrg, mpo as M, mps as mps_tools
    g = golden["leftright"]
    N, seed = (int(v) for v in g["sched_cfg"])
    mpo = M.asep_mpo(N, tuple(float(v) for v in g["asep8_params"]))
    ctx = E.Context("c128")
    np.random.seed(seed)
    fname = str(tmp_path / "st")
    Evec, EEvec, gapvec = dmrg.run_dmrg(mpo, mbd=[4, 8, 16], tol=1e-10, maxIter=4, fname=fname, ctx=ctx, native=False)
    assert Evec.shape == (3,) and np.all(np.isnan(gapvec))
    assert np.max(np.abs(Evec - g["sched_E"]) / np.abs(g["sched_E"])) < 2e-7
    assert np.max(np.abs(EEvec - g["sched_EE"])) < 1e-5
    mpsL, gs = mps_tools.load_mps(fname + "_mbd2")
    assert int(gs) == N // 2 + 1 and len(mpsL[0]) == N and max(max(t.shape) for t in mpsL[0]) == 16


def test_idmrg_run(E, golden):
    from pydmrg_b200 import idmrg, mpo as M
    g = golden["runs"]
    ctx = E.Context("c128")
    mpo4 = M.asep_mpo(4, (0.2, 0.8, 0.1, 0.9, 0.4, 0.6, -0.5))
    np.random.seed(5)
    out = idmrg.run_idmrg(mpo4, mbd=10, maxIter=20, returnEntSpec=True, ctx=ctx, native=False)
    assert abs(out[0] - g["idmrg_E"]) < 1e-6 * abs(g["idmrg_E"])
    assert abs(out[1] - g["idmrg_EE"]) < 1e-6


def test_f64_context_heisenberg_two_site(E):
    """dtype='f64' host path (real Ritz pairs forced, real tensors end to end) against the complex oracle run."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import dmrg, mpo as M
    mpo = M.heisenberg_mpo(10)
    ctx = E.Context("f64")
    np.random.seed(2)
    out = dmrg.run_dmrg(mpo, mbd=16, tol=1e-10, maxIter=4, twoSite=True, dtype="f64", ctx=ctx, native=False, res_tol=1e-10)
    np.random.seed(2)
    Eo = orc.run_dmrg_twosite(mpo, 16, n_sweeps=5)[0]
    assert abs(out[0] - Eo) < 1e-8 * abs(Eo) and abs(np.imag(out[0])) == 0


# ---- gauge utilities ----------------------------------------------------------------------------------
def _dense_state(mps):
    v = np.asarray(mps[0])[:, 0, :]                              # (d, D1)
    for t in mps[1:]:
        v = np.einsum("...a,pab->...pb", v, np.asarray(t))
    return v[..., 0].reshape(-1)


def test_gauge_utilities(E):
    """gauge.move_gauge_* / make_mps_right / calc_entanglement_all (tools/mps_tools.py:36-96,
    tools/aux/process_states.py:5-27): the state is unchanged, the sites left behind are isometries, and the entropy of
    every bond equals the one of the dense bipartition."""
    from pydmrg_b200 import gauge, mps as mps_tools
    ctx = E.Context("c128")
    N = 7
    np.random.seed(4)
    mps = [t + 0.3j * np.random.rand(*t.shape) for t in mps_tools.create_rand_mps(N, 6)]
    psi0 = _dense_state(mps)
    right = gauge.make_mps_right([t.copy() for t in mps], ctx=ctx)
    assert isinstance(right[0], np.ndarray)
    assert np.allclose(_dense_state(right), psi0, atol=1e-14 * np.abs(psi0).max() * 100)
    for t in right[1:]:
        assert np.allclose(np.einsum("pab,pcb->ac", t, t.conj()), np.eye(t.shape[1]), atol=1e-12)
    ref = mps_tools.make_mps_right([t.copy() for t in mps])      # the host restatement of the reference's routine
    assert np.allclose(np.abs(np.vdot(_dense_state(ref), _dense_state(right))), np.vdot(psi0, psi0).real, rtol=1e-12)
    moved = gauge.move_gauge([t.copy() for t in right], 0, 4, ctx=ctx)
    assert np.allclose(_dense_state(moved), psi0, atol=1e-12 * np.abs(psi0).max())
    for t in moved[:4]:
        assert np.allclose(np.einsum("pab,pac->bc", t.conj(), t), np.eye(t.shape[2]), atol=1e-12)
    back = gauge.move_gauge(moved, 4, 1, ctx=ctx)
    assert np.allclose(_dense_state(back), psi0, atol=1e-12 * np.abs(psi0).max())
    EE = gauge.calc_entanglement_all([[t.copy() for t in right]], gSite=0, ctx=ctx)
    psi = psi0 / np.linalg.norm(psi0)
    for b in range(N - 1):
        s = np.linalg.svd(psi.reshape(2 ** (b + 1), -1), compute_uv=False)
        p = s[s > 1e-15] ** 2
        assert abs(EE[0, b] - np.sum(-p * np.log2(p))) < 1e-10


def test_preconditioned_two_site_run(E, golden):
    """run_dmrg(..., precond='diag'): same energy as ED / the unpreconditioned run, fewer mat-vecs from a cold start."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["multistate"]
    mpo = M.asep_mpo(8, tuple(float(v) for v in g["params"]))
    ctx = E.Context("c128")
    res = {}
    for pc in (None, "diag"):
        np.random.seed(1)
        st = {}
        out = dmrg.run_dmrg(mpo, mbd=16, tol=1e-10, maxIter=3, twoSite=True, ctx=ctx, native=False, res_tol=1e-10, stats=st,
                            precond=pc)
        res[pc] = (out[0], st["n_matvec"])
    ed = g["asep8_ed"][0]
    assert abs(res["diag"][0] - ed) < 1e-9 * abs(ed) and abs(res[None][0] - ed) < 1e-9 * abs(ed)
    assert res["diag"][1] < res[None][1]


# ---- the rest of the run API ("kept verbatim") ------------------------------------------------------------
def _iso_left(t):
    return np.allclose(np.einsum("pab,pac->bc", np.conj(t), t), np.eye(t.shape[2]), atol=1e-10)


def _iso_right(t):
    return np.allclose(np.einsum("pab,pcb->ac", t, np.conj(t)), np.eye(t.shape[1]), atol=1e-10)


def test_gauge_site_save_and_restart_from_returned_state(E, golden):
    """gaugeSiteSave (dmrg.py:272-291): the returned state is left-canonical before that site and right-canonical
    after it; restarting from (returned state, gaugeSite) and from the returned environment (initGuess / initEnv,
    returnState / returnEnv) reproduces the converged energy at once."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["runs"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["asep6_params"]))
    ctx = E.Context("c128")
    np.random.seed(3)
    E0, EE, gap, mps, env = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=6, ctx=ctx, native=False, res_tol=1e-11, gaugeSiteSave=2,
                                          returnState=True, returnEnv=True)
    assert isinstance(mps[0][0], np.ndarray) and isinstance(env[0][0], np.ndarray) and len(env[0]) == 7
    assert all(_iso_left(t) for t in mps[0][:2]) and all(_iso_right(t) for t in mps[0][3:])
    assert abs(E0 - g["asep6_ed_top"]) < 1e-9 * abs(g["asep6_ed_top"])
    st = {}
    E1 = dmrg.run_dmrg(mpo, mbd=8, tol=1e-8, maxIter=3, ctx=ctx, native=False, res_tol=1e-11, initGuess=(mps, 2), stats=st)[0]
    assert abs(E1 - E0) < 1e-10 * abs(E0)
    E2 = dmrg.run_dmrg(mpo, mbd=8, tol=1e-8, maxIter=3, ctx=ctx, native=False, res_tol=1e-11, initGuess=(mps, 2), initEnv=env)[0]
    assert abs(E2 - E0) < 1e-10 * abs(E0)


def test_restart_from_files_and_schedule_continuation(E, golden, tmp_path):
    """initGuess as a file stem (dmrg.py:376-389): stage 0 loads <stem>_mbd0, later stages <stem>_mbd{k-1} padded."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["runs"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["asep6_params"]))
    ctx = E.Context("c128")
    np.random.seed(3)
    a = str(tmp_path / "a")
    Ea = dmrg.run_dmrg(mpo, mbd=[2, 4], tol=1e-10, maxIter=4, ctx=ctx, native=False, fname=a)[0]
    Eb = dmrg.run_dmrg(mpo, mbd=[4, 8], tol=1e-10, maxIter=4, ctx=ctx, native=False, res_tol=1e-11, initGuess=a)[0]
    assert Ea.shape == (2,) and Eb.shape == (2,)
    assert abs(Eb[1] - g["asep6_ed_top"]) < 1e-9 * abs(g["asep6_ed_top"])
    assert Ea[1].real <= Eb[1].real + 1e-9 and Ea[0].real <= Ea[1].real + 1e-9        # variational in chi (largest Re E)


def test_two_states_options(E, golden):
    """targetState, orthonormalize (bool and gap threshold, diag_tools.py:275-286), preserveState, minIter."""
    from pydmrg_b200 import dmrg, mpo as M
    g = golden["multistate"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["params"]))
    ed = g["asep6_ed"]
    ctx = E.Context("c128")
    kw = dict(mbd=8, tol=1e-10, maxIter=5, nStates=2, ctx=ctx, native=False, res_tol=1e-10)
    np.random.seed(6)
    out1 = dmrg.run_dmrg(mpo, targetState=1, **kw)
    assert abs(out1[0] - ed[1]) < 1e-7 * abs(ed[0]) and abs(out1[2] - (ed[0] - ed[1])) < 1e-7
    for orth in (True, 10.0, 1e-12):
        np.random.seed(6)
        out = dmrg.run_dmrg(mpo, orthonormalize=orth, **kw)
        assert abs(out[0] - ed[0]) < 1e-7 * abs(ed[0]) and abs(out[2] - (ed[0] - ed[1])) < 1e-6
    # preserveState keeps the previous tensor whenever no eigenvector overlaps it by 0.98 (diag_tools.py:93-97).  With
    # two states the reference then indexes a second vector that no longer exists (IndexError, dmrg.py:288); with one
    # state it runs -- reference value for this seed: E = 0.8905626478 (run here through oracle/ref_loader.py)
    np.random.seed(6)
    out = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=5, nStates=1, preserveState=True, ctx=ctx, native=False)
    assert abs(out[0] - 0.8905626478039578) < 1e-7
    st = {}
    np.random.seed(6)
    dmrg.run_dmrg(mpo, mbd=8, tol=1.0, maxIter=10, minIter=3, ctx=ctx, native=False, stats=st)
    st2 = {}
    np.random.seed(6)
    dmrg.run_dmrg(mpo, mbd=8, tol=1.0, maxIter=10, minIter=0, ctx=ctx, native=False, stats=st2)
    assert st["n_matvec"] > st2["n_matvec"]                     # minIter forces more sweeps than the loose tol needs


def test_periodic_chain(E):
    """Periodic ASEP (extra two-site string operators with None entries, mpo/asep.py:77-94): one-site and two-site
    runs against dense diagonalisation of the sum of all MPO terms."""
    from pydmrg_b200 import dmrg, mpo as M
    N = 6
    mpo = M.asep_mpo(N, (0.0, 0.0, 0.3, 0.7, 0.0, 0.0, -0.4), periodic=True)
    assert len(mpo) > 1 and any(W is None for W in mpo[1])
    H = 0
    for op in mpo:
        T = None
        for W in op:
            Wm = np.eye(2).reshape(1, 1, 2, 2) if W is None else W
            T = Wm[0] if T is None else np.einsum("a...,abij->b...ij", T, Wm)
        T = T[0]
        idx = list(range(0, 2 * N, 2)) + list(range(1, 2 * N, 2))
        H = H + T.transpose(idx).reshape(2 ** N, 2 ** N)
    ev = np.linalg.eigvals(H)
    top = ev[np.argmax(ev.real)]
    ctx = E.Context("c128")
    for two in (False, True):
        np.random.seed(2)
        out = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=6, ctx=ctx, native=False, res_tol=1e-11, twoSite=two)
        assert abs(out[0] - top) < 1e-8 * max(1.0, abs(top)), (two, out[0], top)


def test_two_site_run_with_warm_started_truncation(E):
    """run_dmrg(twoSite=True) at chi = 64 on 16 sites -- the 128 x 128 cuts take the projective route and, under the
    start policy, the warm-started one -- against the same run with ``recycle=False``: the whole host flow
    (sweeps, Python Davidson loop, truncation bookkeeping) with nothing stubbed above the device entry points."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.heisenberg_mpo(16)
    res = {}
    for rec in (False, True):
        np.random.seed(2)
        st = {}
        out = dmrg.run_dmrg(mpo, mbd=64, tol=0.0, maxIter=3, twoSite=True, ctx=ctx, native=False, res_tol=1e-10, stats=st, recycle=rec)
        res[rec] = (out[0], out[1], st)
    assert res[True][2].get("n_recycled", 0) > 0 and res[False][2].get("n_recycled", 0) == 0
    assert res[True][2].get("n_exact_trunc", 0) >= 8                      # centre bond: both directions, every sweep
    assert abs(res[True][0] - res[False][0]) < 1e-10 * abs(res[False][0])
    assert abs(res[True][1] - res[False][1]) < 1e-8


@pytest.mark.parametrize("N", [3, 4, 5, 7])
def test_short_and_odd_chains(E, N):
    """Edge cases of the chain length (odd N has its own branch in the reference's allocation code,
    mps_tools.py:217, env_tools.py:19): one-site and two-site runs against dense diagonalisation."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(N, PRM)
    T = mpo[0][0][0]
    for W in mpo[0][1:]:
        T = np.einsum("a...,abij->b...ij", T, W)
    idx = list(range(0, 2 * N, 2)) + list(range(1, 2 * N, 2))
    ev = np.linalg.eigvals(T[0].transpose(idx).reshape(2 ** N, 2 ** N))
    top = ev[np.argmax(ev.real)]
    for two in (False, True):
        np.random.seed(1)
        out = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=4, ctx=ctx, native=False, res_tol=1e-11, twoSite=two)
        assert abs(out[0] - top) < 1e-10 * abs(top)


def test_two_site_chain_of_two_sites(E):
    """N = 2: one bond.  Two-site sweeps solve it exactly; the reference's one-site driver cannot run it (its default
    gaugeSiteSave = N/2 + 1 = 2 steps past the last site, dmrg.py:326 -> IndexError), and neither does this one."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(2, PRM)
    T = np.einsum("abij,bckl->acikjl", mpo[0][0], mpo[0][1])[0, 0].reshape(4, 4)
    ev = np.linalg.eigvals(T)
    np.random.seed(1)
    out = dmrg.run_dmrg(mpo, mbd=4, tol=1e-10, maxIter=2, ctx=ctx, native=False, res_tol=1e-11, twoSite=True)
    assert abs(out[0] - ev[np.argmax(ev.real)]) < 1e-12
    with pytest.raises(IndexError):
        dmrg.run_dmrg(mpo, mbd=4, tol=1e-10, maxIter=2, ctx=ctx, native=False)


def test_gauge_shortcut_for_unchanged_bonds(E):
    """run_dmrg(twoSite=True, gauge_shortcut=True): bonds whose local solve stops in its first cycle are re-gauged from
    the centre tensor instead of truncated again -- same energies, entropy and mat-vec count as the plain run."""
    from pydmrg_b200 import dmrg, mpo as M
    ctx = E.Context("c128")
    mpo = M.asep_mpo(14, PRM)
    res = {}
    for gs in (False, True):
        np.random.seed(2)
        st = {}
        out = dmrg.run_dmrg(mpo, mbd=32, tol=0.0, maxIter=6, twoSite=True, ctx=ctx, native=False, res_tol=1e-9, stats=st,
                            gauge_shortcut=gs)
        res[gs] = (out[0], out[1], st)
    assert res[True][2].get("n_gauge_moves", 0) > 50 and res[False][2].get("n_gauge_moves", 0) == 0
    assert abs(res[True][0] - res[False][0]) < 1e-11 * abs(res[False][0])
    assert abs(res[True][1] - res[False][1]) < 1e-9
    assert abs(res[True][2]["n_matvec"] - res[False][2]["n_matvec"]) <= 0.02 * res[False][2]["n_matvec"]
    # one-site sweeps ignore the option
    np.random.seed(2)
    dmrg.run_dmrg(mpo, mbd=8, tol=1e-6, maxIter=1, ctx=ctx, native=False, gauge_shortcut=True)


def test_full_contract_with_orthonormalised_states(E, golden):
    """full_contract(orth=True) (tools/contract.py:16-19 -> orthonormalize_states, mps_tools.py:392-430) on the two
    states returned by a nStates = 2 run: orthonormal after the call.  (scipy.linalg.orth returns a basis of the SPAN,
    so the individual states are mixed -- the reference's own comment at mps_tools.py:416 says "something is wrong
    here"; the product reproduces the reference's arrays bit for bit, tests/test_oracle_vs_reference.py.)"""
    from pydmrg_b200 import contract, dmrg, mpo as M
    g = golden["multistate"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["params"]))
    ctx = E.Context("c128")
    np.random.seed(6)
    gs = 4
    E0, EE, gap, mps = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=5, nStates=2, ctx=ctx, native=False, res_tol=1e-10,
                                     returnState=True, gaugeSiteSave=gs)
    bra = [[np.conj(t) for t in Mst] for Mst in mps]
    kw = dict(mps=mps, lmps=bra, orth=True, gSite=gs, glSite=gs, ctx=ctx)
    n00 = contract.full_contract(state=0, lstate=0, **kw)
    n01 = contract.full_contract(state=1, lstate=0, **kw)
    assert abs(n00 - 1) < 1e-12 and abs(n01) < 1e-12
    e0 = contract.full_contract(mpo=mpo, state=0, lstate=0, **kw) / n00
    e1 = contract.full_contract(mpo=mpo, state=1, lstate=1, **kw)
    ed = g["asep6_ed"]
    lo, hi = min(ed[0].real, ed[1].real), max(ed[0].real, ed[1].real)
    assert lo - 1e-6 < e0.real < hi + 1e-6 and lo - 1e-6 < e1.real < hi + 1e-6    # both live in the two-state span


def test_two_site_resume_from_one_site_checkpoint(E, golden, tmp_path):
    """A one-site checkpoint (centre saved at N//2+1, the reference's default gaugeSiteSave) resumed with twoSite=True,
    and initGuess=(mps, g): the centre is moved to site 0 before the blocks are built, so the very first two-site
    sweep already sits at the converged energy (maxIter=0: one sweep).  Before the fix the blocks F[1..g] were left
    blocks used as right ones and the first sweep solved wrong local problems."""
    from pydmrg_b200 import dmrg, mpo as M, mps as mps_tools
    g = golden["runs"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["asep6_params"]))
    ctx = E.Context("c128")
    np.random.seed(3)
    stem = str(tmp_path / "one")
    E0, _, _, mps = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=6, ctx=ctx, native=False, res_tol=1e-11, fname=stem, returnState=True)
    assert mps_tools.load_mps(stem + "_mbd0")[1] == 6 // 2 + 1
    st = {}
    E1 = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=0, ctx=ctx, native=False, res_tol=1e-11, initGuess=stem, twoSite=True, stats=st)[0]
    assert abs(E1 - E0) < 1e-10 * abs(E0)
    # a converged start costs about one mat-vec cycle per bond (2 x 5 bonds); wrong blocks cost hundreds
    assert st["n_matvec"] <= 40, st["n_matvec"]
    E2 = dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=0, ctx=ctx, native=False, res_tol=1e-11, initGuess=(mps, 4), twoSite=True)[0]
    assert abs(E2 - E0) < 1e-10 * abs(E0)


@pytest.mark.parametrize("twoSite", [False, True])
def test_in_memory_guess_continues_through_the_chi_schedule(E, golden, twoSite):
    """initGuess as an in-memory state with several mbd stages: stage k > 0 continues from stage k-1 (it used to
    restart every stage from the original guess, so the schedule did no useful work)."""
    from pydmrg_b200 import dmrg, mpo as M, mps as mps_tools
    g = golden["runs"]
    mpo = M.asep_mpo(6, tuple(float(v) for v in g["asep6_params"]))
    ctx = E.Context("c128")
    np.random.seed(5)
    guess = mps_tools.make_all_mps_right(mps_tools.create_all_mps(6, 2, 1))
    st = {}
    Es = dmrg.run_dmrg(mpo, mbd=[2, 4, 8], tol=1e-10, maxIter=[4, 4, 0], ctx=ctx, native=False, res_tol=1e-11, initGuess=guess,
                       twoSite=twoSite, stats=st)[0]
    # the last stage runs ONE sweep only: it can only sit at the ED value if it started from the chi=4 result
    assert abs(Es[2] - g["asep6_ed_top"]) < (1e-9 if twoSite else 1e-6) * abs(g["asep6_ed_top"])
    assert Es[0].real <= Es[1].real + 1e-9 <= Es[2].real + 2e-9
