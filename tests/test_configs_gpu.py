"""GPU parity tests on the BASELINE.json configurations themselves ("parity on the BASELINE configs, on the
device"): the benchmarked chi=1024 mat-vec against the oracle, the warm-started truncation where chi TRUNCATES, and
configs [1], [2], [3] at sizes the NumPy oracle finishes in tens of seconds, sweep by sweep.

Tolerances: contractions 1e-13; energies 1e-10 relative where the oracle's own solver floor allows it (Hermitian: the
error of a Ritz value is O(||r||^2), the reference stops at ||r|| <= 1e-7), 1e-7 relative against the oracle's
NON-Hermitian runs (O(||r||)); entanglement entropy / spectrum 1e-8; observables <l|O|r>/<l|r> 1e-8."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TASEP = (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, -0.5)          # bench.py workload (configs[2] operator, W rows 3,4 empty)
TASEP_ACTIVE = (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, 0.5)    # s > 0: current-enhanced, entangled state
ASEP = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)


@pytest.fixture(scope="module")
def P():
    import types
    from pydmrg_b200 import contract, core, device, dmrg, env, mpo, mps, scan, truncate
    return types.SimpleNamespace(dmrg=dmrg, env=env, core=core, mpo=mpo, mps=mps, truncate=truncate, D=device,
                                 contract=contract, scan=scan)


# ---------------------------------------------------------------------------------------------------------------
# (a) the benchmarked problem itself
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prm", [TASEP, ASEP], ids=["tasep_index_list", "asep_dense_w"])
def test_chi1024_two_site_matvec_vs_oracle(P, prm):
    """bench.py's step -- the chi=1024, d=2, w=6 complex128 two-site H_eff mat-vec (128-row tiles, 8 m-tiles, slab
    index lists) -- against the oracle's restatement of diag_tools.py:146-162 on the same seeded inputs: 1e-13."""
    from oracle import dmrg_oracle as orc
    chi = 1024
    ctx = P.core.Context("c128")
    W1, W2 = P.mpo.asep_mpo(6, prm)[0][2:4]
    rng = np.random.default_rng(0)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    L, R, x = cr(chi, 6, chi), cr(chi, 6, chi), cr(2, 2, chi, chi)
    plan = P.D.HeffPlan(P.D.C128, 4, chi, chi, [(ctx.dev(L), P.D.combine_twosite(W1, W2, np.complex128), ctx.dev(R))], ctx.ws)
    xd = ctx.dev(x.reshape(-1))
    yd = torch.empty_like(xd)
    plan.apply(xd, yd, -1.0)
    assert P.D.lib.pdmrg_last_gemm_path() == 1                  # the TMA / DMMA tile kernel, not the generic one
    ref = orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas")
    err = np.linalg.norm(yd.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-13, err
    plan.close()


def test_chi1024_block_update_vs_oracle(P):
    """update_envR at chi=1024 (the other chi^3 kernel sequence of a bond step) against the oracle: 1e-13."""
    from oracle import dmrg_oracle as orc
    chi = 1024
    ctx = P.core.Context("c128")
    W = P.mpo.asep_mpo(6, ASEP)[0][2]
    rng = np.random.default_rng(1)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    L, A = cr(chi, 6, chi), cr(2, chi, chi)
    out = P.D.env_update_right(ctx.dev(A), W.astype(np.complex128), ctx.dev(L), None, ctx.ws)
    ref = orc.env_update_right(A, W, L, None, "blas")
    err = np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-13, err


# ---------------------------------------------------------------------------------------------------------------
# (b) warm-started truncation with chi truncating, against fresh decompositions and against the oracle's SVD sweeps
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["heisenberg_f64", "tasep_c128"])
def test_truncating_sweeps_recycled_vs_exact_vs_oracle(P, model):
    """The warm-started truncation where chi TRUNCATES HARD, against an oracle without a solver floor: Heisenberg N=14 /
    entangled TASEP (s=+0.5) N=14 at chi=8 -- discarded weight 1e-7..1e-3 at the centre (asserted) -- with the
    projective / warm-started path switched on for these small cuts (Context(projective_min=8); production default 128)
    and every local problem of the ORACLE solved by dense diagonalisation (alg='exact', diag_tools.py:193-212), so its
    fixed point is exact.  Product with recycle=True, with recycle=False, and the oracle's plain-SVD sweeps
    (mps_tools.py:98-120) from the same seeded start: E 1e-10 relative, EE and the entanglement spectrum 1e-8
    (the north-star tolerances)."""
    from oracle import dmrg_oracle as orc
    N, chi, sweeps = 14, 8, 8
    if model == "heisenberg_f64":
        mpo, dtype = P.mpo.heisenberg_mpo(N), "f64"
    else:
        mpo, dtype = P.mpo.asep_mpo(N, TASEP_ACTIVE), "c128"
    np.random.seed(4)
    start = orc.make_mps_right(orc.random_mps(N, chi))
    Eo, EEo, EEso, So, _, _, hist = orc.run_dmrg_twosite(mpo, chi, n_sweeps=sweeps, alg="exact", mps=[t.copy() for t in start])
    assert abs(hist[-1] - hist[-2]) < 1e-12 * abs(Eo), "oracle not at its fixed point"
    res = {}
    for rec in (False, True):
        ctx = P.core.Context(dtype, projective_min=8)
        st = {}
        out = P.dmrg.run_dmrg(mpo, mbd=chi, tol=0.0, maxIter=sweeps - 1, twoSite=True, ctx=ctx, returnEntSpec=True, stats=st,
                              recycle=rec, res_tol=1e-12, initGuess=[[t.copy() for t in start]])
        res[rec] = (out[0], float(np.real(out[1])), np.sort(np.asarray(out[3]))[::-1], st)
    (E0, EE0, s0, st0), (E1, EE1, s1, st1) = res[False], res[True]
    assert st1.get("n_recycled", 0) > 0 and st0.get("n_recycled", 0) == 0
    assert st1.get("disc_max", 0.0) > 1e-8, "chi must truncate in this test: %r" % st1.get("disc_max")
    scale = abs(Eo)
    assert abs(E1 - E0) < 1e-10 * scale, (E1, E0)
    assert abs(E0 - Eo) < 1e-10 * scale and abs(E1 - Eo) < 1e-10 * scale, (E0, E1, Eo)
    assert abs(EE1 - EE0) < 1e-8 and abs(EE0 - EEo) < 1e-8 and abs(EE1 - EEo) < 1e-8, (EE0, EE1, EEo)
    so = np.sort(np.asarray(EEso))[::-1]
    assert np.max(np.abs(s1 - s0)) < 1e-8 and np.max(np.abs(s0 - so)) < 1e-8 and np.max(np.abs(s1 - so)) < 1e-8


# ---------------------------------------------------------------------------------------------------------------
# (c) configs[2]: non-Hermitian left + right eigenstates, two-site, entangled TASEP
# ---------------------------------------------------------------------------------------------------------------
def test_config2_left_right_two_site_entangled(P):
    """configs[2] at a size the oracle affords (TASEP alpha=0.35, beta=2/3, s=+0.5 -- the current-enhanced side, EE ~ 0.7;
    N=24, chi=32): run_dmrg(calcLeftState=True, twoSite=True) against the oracle's two runs (dmrg.py:405,451: the left
    state is the right state of mpo_conj_trans) and the observables of dmrg.py:448-465 / tests/asep/contractCheck.py:
    <l|H|r>/<l|r> = E to 1e-8, current <l|J|r>/<l|r> vs the oracle's contraction of ITS states."""
    from oracle import dmrg_oracle as orc
    N, chi, sweeps = 24, 32, 5
    prm = TASEP_ACTIVE
    mpo = P.mpo.asep_mpo(N, prm)
    mpol = P.mpo.mpo_conj_trans(mpo)
    np.random.seed(7)
    start = orc.make_mps_right(orc.random_mps(N, chi))
    Er, EEr, _, _, mr, _, _ = orc.run_dmrg_twosite(mpo, chi, n_sweeps=sweeps, mps=[t.copy() for t in start])
    El, EEl, _, _, ml, _, _ = orc.run_dmrg_twosite(mpol, chi, n_sweeps=sweeps, mps=[t.copy() for t in start])
    cur_mpo = P.mpo.asep_curr_mpo(N, prm)
    den_o = orc.full_contract(None, mr, ml)
    cur_o = orc.full_contract(cur_mpo, mr, ml) / den_o
    guess = [[t.copy() for t in start]]
    E, EE, gap, mps = P.dmrg.run_dmrg(mpo, mbd=chi, tol=0.0, maxIter=sweeps - 1, twoSite=True, calcLeftState=True, returnState=True,
                                      initGuess=[guess, [[t.copy() for t in start]]], res_tol=1e-11)
    r, l = mps
    assert EE[0] > 0.3, "state must be entangled: EE = %r" % (EE,)
    scale = abs(Er)
    assert abs(E - Er) < 1e-7 * scale and abs(np.conj(El) - Er) < 1e-6 * scale
    assert abs(EE[0] - EEr) < 1e-6 and abs(EE[1] - EEl) < 1e-6
    den = P.contract.full_contract(mps=r, lmps=l)
    ene = P.contract.full_contract(mpo=mpo, mps=r, lmps=l) / den
    cur = P.contract.full_contract(mpo=cur_mpo, mps=r, lmps=l) / den
    assert abs(ene - E) < 1e-8 * scale, (ene, E)
    assert abs(cur - cur_o) < 1e-6 * max(1.0, abs(cur_o)), (cur, cur_o)
    # d psi / d s = <J> (Hellmann-Feynman on the tilted generator): the current is the s-derivative of E
    h = 1e-4
    Es = []
    for ds in (-h, h):
        out = P.dmrg.run_dmrg(P.mpo.asep_mpo(N, prm[:6] + (prm[6] + ds,)), mbd=chi, tol=0.0, maxIter=2, twoSite=True,
                              initGuess=[[np.array(t) for t in r[0]]], res_tol=1e-11)
        Es.append(out[0])
    assert abs((Es[1] - Es[0]) / (2 * h) - cur) < 1e-5 * max(1.0, abs(cur))


# ---------------------------------------------------------------------------------------------------------------
# (d) configs[1] and [3], sweep by sweep at a chi the oracle affords
# ---------------------------------------------------------------------------------------------------------------
def test_config1_heisenberg_sweep_by_sweep(P):
    """configs[1] (Heisenberg ground-state sweeps, real f64 path) at N=30, chi=32: the energy after EVERY sweep and the
    final entropy / spectrum against the oracle's two-site sweeps from the same seeded start."""
    from oracle import dmrg_oracle as orc
    N, chi, sweeps = 30, 32, 4
    mpo = P.mpo.heisenberg_mpo(N)
    np.random.seed(11)
    start = orc.make_mps_right(orc.random_mps(N, chi))
    Eo, EEo, EEso, _, _, _, hist = orc.run_dmrg_twosite(mpo, chi, n_sweeps=sweeps, mps=[t.copy() for t in start])
    ctx = P.core.Context("f64")
    mpsL = [[ctx.dev(t) for t in start]]
    mpsL[0][0] = mpsL[0][0] / torch.linalg.vector_norm(mpsL[0][0])
    F = P.env.calc_env(mpsL[0], mpo, chi, gaugeSite=0, ctx=ctx)
    for sw in range(sweeps):
        E, mpsL, F, EE, EEs, S = P.dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, res_tol=1e-11)
        assert abs(E - hist[sw]) < 1e-9 * abs(hist[sw]), (sw, E, hist[sw])
    assert abs(E - Eo) < 1e-10 * abs(Eo)
    # entropies are first order in the state error and the oracle's Davidson stops at ||r|| <= 1e-7 (the reference's
    # rule): 1e-6 here; the 1e-8 comparison against an oracle WITHOUT that floor is
    # test_truncating_sweeps_recycled_vs_exact_vs_oracle
    assert abs(EE - EEo) < 1e-6
    a, b = np.sort(np.asarray(EEs))[::-1], np.sort(np.asarray(EEso))[::-1]
    assert np.max(np.abs(a[:16] - b[:16])) < 1e-6


def test_config3_scan_points_vs_oracle(P):
    """configs[3] (ASEP s-field scan, one independent two-site run per s) at N=20, chi=32 for three s values on both sides of the
    s=0 transition (the points next to it are covered against ED in test_scan_across_the_transition_vs_ed), through scan.run_scan (chi ramp, cold starts) against the oracle's runs: the reference's solver
    floor (1e-7 relative, non-Hermitian) and 1e-6 on the entropy."""
    from oracle import dmrg_oracle as orc
    N, chi = 20, 32
    s_vals = [-0.3, 0.2, 0.4]
    mpo_of = lambda s: P.mpo.asep_mpo(N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    np.random.seed(3)
    res = P.scan.run_scan(mpo_of, s_vals, chi, ramp=(8,), maxIter=5, tol=1e-11, res_tol=1e-10)
    for i, s in enumerate(s_vals):
        np.random.seed(3)
        Eo, EEo = orc.run_dmrg_twosite(mpo_of(s), chi, n_sweeps=5)[:2]
        assert abs(res["E"][i] - Eo) < 1e-7 * max(1.0, abs(Eo)), (s, res["E"][i], Eo)
        assert abs(res["EE"][i] - EEo) < 1e-6, (s, res["EE"][i], EEo)


# ---------------------------------------------------------------------------------------------------------------
# scans across the s = 0 transition against dense ED; concurrent workers; thick restart; device Arnoldi
# ---------------------------------------------------------------------------------------------------------------
def dense_from_mpo(mpo):
    """2^N x 2^N matrix of an MPO list (sum over its terms; None = identity), physical index of site 0 slowest."""
    N = len(mpo[0])
    tot = 0
    for op in mpo:
        cur = np.ones((1, 1, 1), dtype=np.complex128)
        for site in range(N):
            W = op[site]
            if W is None:
                W = np.einsum("ab,ij->abij", np.eye(cur.shape[0]), np.eye(2))
            cur = np.einsum("aij,abkl->bikjl", cur, W).reshape(W.shape[1], cur.shape[1] * 2, cur.shape[2] * 2)
        tot = tot + cur[0]
    return tot


S_ACROSS = [-0.12, -0.06, -0.02, 0.01, 0.03, 0.08, 0.15]


@pytest.fixture(scope="module")
def ed_across():
    from pydmrg_b200 import mpo as M
    out = []
    for s in S_ACROSS:
        w = np.linalg.eigvals(dense_from_mpo(M.asep_mpo(10, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, s))))
        w = w[np.argsort(-w.real)]
        out.append((w[0].real, w[1].real))
    return out


def test_scan_across_the_transition_vs_ed(P, ed_across):
    """configs[3] in miniature against dense ED (N=10, chi=32 = exact): seven s values across s=0, where the two leading
    eigenvalues of the tilted generator approach each other.  (1) cold starts, dynamic schedule, 2 concurrent workers
    (threads + streams on one GPU), thick restart; (2) the same with one worker and the reference's restart -- E to 1e-9."""
    mpo_of = lambda s: P.mpo.asep_mpo(10, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    for opts in (dict(workers=2, restart_keep=4), dict(workers=1)):
        st = {}
        res = P.scan.run_scan(mpo_of, S_ACROSS, 32, ramp=(8,), maxIter=6, tol=1e-12, res_tol=1e-11, schedule="dynamic", seed=11,
                              stats=st, **opts)
        assert sorted(res["points"]) == list(range(len(S_ACROSS)))
        for i, s in enumerate(S_ACROSS):
            assert abs(res["E"][i] - ed_across[i][0]) < 1e-9 * max(1.0, abs(ed_across[i][0])), (opts, s, res["E"][i], ed_across[i])
        if opts.get("restart_keep"):
            assert st.get("n_thick_restarts", 0) > 0


def test_continuation_scan_two_states_vs_ed(P, ed_across):
    """The reference's scan mode (scripts/asep/current/bondDimConv.py:19-27): serial continuation, every point
    warm-started from the previous point's states, nStates=2 so that both branches of the transition stay in the
    targeted space -- E and the gap E0 - E1 against dense ED at every s (one-site sweeps, the reference's algorithm;
    chi = 32 is exact at N=10)."""
    mpo_of = lambda s: P.mpo.asep_mpo(10, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    np.random.seed(5)
    prev, Es, gaps = None, [], []
    for s in S_ACROSS:
        out = P.dmrg.run_dmrg(mpo_of(s), mbd=32, tol=1e-11, maxIter=8, nStates=2, initGuess=prev, returnState=True, res_tol=1e-11)
        Es.append(out[0]); gaps.append(out[2])
        prev = (out[-1], 10 // 2 + 1)
    for i, s in enumerate(S_ACROSS):
        e0, e1 = ed_across[i]
        assert abs(Es[i] - e0) < 1e-8 * max(1.0, abs(e0)), (s, Es[i], e0)
        assert abs(gaps[i] - (e0 - e1)) < 1e-6 * max(1.0, abs(e0)), (s, gaps[i], e0 - e1)
    res = P.scan.run_scan(mpo_of, S_ACROSS, 32, ramp=(), maxIter=8, tol=1e-11, res_tol=1e-11, continuation=True, twoSite=False,
                          nStates=2)
    for i, s in enumerate(S_ACROSS):
        assert abs(res["E"][i] - ed_across[i][0]) < 1e-8 * max(1.0, abs(ed_across[i][0])), (s, res["E"][i], ed_across[i])


@pytest.mark.parametrize("dtype", ["c128", "f64"])
def test_thick_restart_and_device_arnoldi_on_a_local_problem(P, dtype):
    """A physical two-site local problem with a nearly degenerate leading pair (ASEP next to s=0) / the Heisenberg chain:
    the reference's restart, thick restarts (C++ loop and Python loop) and device-resident thick-restarted Arnoldi
    (alg='arnoldi') against dense diagonalisation of the same operator (alg='exact')."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import davidson as dav, eigs
    N, chi = 12, 16
    mpo = P.mpo.heisenberg_mpo(N) if dtype == "f64" else P.mpo.asep_mpo(N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, 0.01))
    c = P.core.Context(dtype)
    np.random.seed(5)
    mps = orc.make_mps_right(orc.random_mps(N, chi))
    env = orc.build_env(mps, mpo, chi, gauge_site=0, mode="blas")
    site = N // 2 - 1
    M = [[c.dev(np.real(t) if dtype == "f64" else t) for t in mps]]
    F = [[c.dev(np.real(t) if dtype == "f64" else t) for t in Fm] for Fm in env]
    Eex, _, _ = eigs.calc_eigs(M, mpo, F, site, 1, alg="exact", oneSite=False, ctx=c)
    res = {}
    for tag, kw in (("reference", {}), ("thick", {"restart_keep": 4}), ("thick_py", {"restart_keep": 4, "native": False}),
                    ("thick_wide", {"restart_keep": 6, "max_space": 24})):
        st = {}
        E, v, _ = eigs.calc_eigs(M, mpo, F, site, 1, alg="davidson", oneSite=False, ctx=c, stats=st, res_tol=1e-10, **kw)
        res[tag] = (E, st["n_matvec"], st.get("n_thick_restarts", 0))
        assert abs(E - Eex) < 1e-9 * max(1.0, abs(Eex)), (tag, E, Eex)
    if res["reference"][1] > 40:                                 # the problem needed restarts at all
        assert res["thick"][2] > 0 and res["thick_py"][2] > 0
        assert res["thick"][1] <= res["reference"][1] + 2, res
    st = {}
    Ea, va, _ = eigs.calc_eigs(M, mpo, F, site, 1, alg="arnoldi", oneSite=False, ctx=c, stats=st)
    assert abs(Ea - Eex) < 1e-5 * max(1.0, abs(Eex)), (Ea, Eex)                       # ARPACK's tol = 1e-5 (diag_tools.py:225)
    Eh, _, _ = eigs.calc_eigs(M, mpo, F, site, 1, alg="arnoldi_host", oneSite=False, ctx=c)
    assert abs(Ea - Eh) < 1e-5 * max(1.0, abs(Eex))
    torch.manual_seed(3)
    M2 = M + [[t + 0.3 * torch.randn_like(t) for t in M[0]]]                          # a second, different guess state
    E2, _, _ = eigs.calc_eigs(M2, mpo, F, site, 2, alg="arnoldi", oneSite=False, ctx=c)
    E2x, _, _ = eigs.calc_eigs(M2, mpo, F, site, 2, alg="exact", oneSite=False, ctx=c)
    # the second state may be one member of a complex-conjugate pair (either member is a valid answer): compare Re and |Im|
    fold = lambda e: np.asarray([complex(z.real, abs(z.imag)) for z in np.asarray(e)])
    assert np.max(np.abs(fold(E2) - fold(E2x))) < 1e-4 * max(1.0, abs(Eex)), (E2, E2x)
