"""GPU tests of the features above the kernels: two-site multi-state sweeps, state-averaged truncation, left + right
eigenstates against the reference's fixture, column-restricted block updates (csrc/env_cols.cu), gauge utilities, the
diagonal preconditioner, multi-pass projection, gauge shortcut / solver options.  First B200 run: round 2
(profiles/gpu_tests_r02.log), all green; they run unconditionally under -m gpu."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRM = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)


@pytest.fixture(scope="module")
def P():
    import pydmrg_b200 as pkg
    from pydmrg_b200 import contract, core, device, dmrg, mpo, parallel, truncate
    return pkg


@pytest.mark.parametrize("tag", ["asep6", "asep8"])
def test_two_site_two_states_gap(P, tag):
    """run_dmrg(nStates=2, twoSite=True): E and the gap against the reference's one-site nStates=2 run and
    dense ED (tests/golden/multistate.npz, generated from the unmodified reference)."""
    g = np.load(os.path.join(GOLD, "multistate.npz"))
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    mpo = P.mpo.asep_mpo(N, tuple(g["params"]))
    np.random.seed(seed)
    out = P.dmrg.run_dmrg(mpo, mbd=mbd, tol=1e-10, maxIter=6, nStates=2, twoSite=True, res_tol=1e-10)
    E, EE, gap = out[0], out[1], out[2]
    ed = g[tag + "_ed"]
    assert abs(E - ed[0]) < 1e-7 * abs(ed[0])
    assert abs(gap - (ed[0] - ed[1])) < 1e-6 * abs(ed[0])
    assert abs(gap - g[tag + "_gap"]) < 1e-6 * abs(ed[0])


@pytest.mark.parametrize("direction", ["right", "left"])
@pytest.mark.parametrize("dtype", ["c128", "f64"])
def test_state_averaged_twosite_truncation(P, direction, dtype):
    """truncate.rdm_truncate_twosite_states on the device against its NumPy restatement."""
    from oracle import multistate_oracle as ms
    c = P.core.Context(dtype)
    rng = np.random.default_rng(11)
    shape = (2, 2, 40, 36)
    k = 48
    n = 4 * 40 * 36
    rnd = (lambda: rng.standard_normal(n) + 1j * rng.standard_normal(n)) if dtype == "c128" else (lambda: rng.standard_normal(n))
    psis = [rnd() for _ in range(2)]
    As, Bs, Sn, EE, EEs = P.truncate.rdm_truncate_twosite_states([c.dev(p) for p in psis], shape, k, c, direction)
    Ao, Bo, Sno, EEo, _ = ms.rdm_truncate_twosite_states(psis, shape, k, direction)
    for s in range(2):
        A, B = c.host(As[s]), c.host(Bs[s])
        Am, Bm = A.transpose(1, 0, 2).reshape(80, k), B.transpose(1, 0, 2).reshape(k, 72)
        Amo, Bmo = Ao[s].transpose(1, 0, 2).reshape(80, k), Bo[s].transpose(1, 0, 2).reshape(k, 72)
        iso = Am if direction == "right" else Bm.conj().T
        assert np.max(np.abs(iso.conj().T @ iso - np.eye(k))) < 1e-12
        assert np.max(np.abs(Am @ Bm - Amo @ Bmo)) < 1e-10
    assert abs(EE - EEo) < 1e-10 and np.max(np.abs(Sn - Sno)) < 1e-10


@pytest.mark.parametrize("tag", ["tasep8", "asep8"])
def test_left_right_run_matches_reference_fixture(P, tag):
    """run_dmrg(calcLeftState=True) against tests/golden/leftright.npz (the unmodified reference's run; states read
    back from its saved files): E, both entropies, current and energy <l|O|r>/<l|r> (north-star: 1e-8 observables on
    converged states; the reference itself stops at ||r|| <= 1e-7, hence 1e-6 against its numbers and 1e-8 against ED)."""
    g = np.load(os.path.join(GOLD, "leftright.npz"))
    N, mbd, seed = (int(v) for v in g[tag + "_cfg"])
    prm = tuple(float(v) for v in g[tag + "_params"])
    mpo = P.mpo.asep_mpo(N, prm)
    np.random.seed(seed)
    E, EE, gap, mps = P.dmrg.run_dmrg(mpo, mbd=mbd, tol=1e-12, maxIter=6, res_tol=1e-11, calcLeftState=True, returnState=True)
    r, l = mps
    scale = abs(g[tag + "_ed_top"])
    assert abs(E - g[tag + "_ed_top"]) < 1e-9 * scale and abs(E - g[tag + "_E"]) < 1e-7 * scale
    assert abs(EE[0] - g[tag + "_EE"]) < 1e-6 and abs(EE[1] - g[tag + "_EEl"]) < 1e-6
    den = P.contract.full_contract(mps=r, lmps=l)
    cur = P.contract.full_contract(mpo=P.mpo.asep_curr_mpo(N, prm), mps=r, lmps=l) / den
    ene = P.contract.full_contract(mpo=mpo, mps=r, lmps=l) / den
    assert abs(cur - g[tag + "_current_lr"]) < 1e-6 * max(1.0, abs(g[tag + "_current_lr"]))
    assert abs(ene - g[tag + "_ed_top"]) < 1e-8 * scale


@pytest.mark.parametrize("dtype", ["c128", "f64"])
@pytest.mark.parametrize("dims", [(2, 96, 80), (2, 64, 64), (2, 40, 200)])
def test_column_restricted_block_updates(P, dtype, dims):
    """pdmrg_env_update_right_cols / _left_cols (csrc/env_cols.cu): the ket-column slices computed separately add
    up to the unrestricted update (validated entry points) and leave every other entry of the output untouched."""
    import torch
    from pydmrg_b200 import device as D
    c = P.core.Context(dtype)
    rng = np.random.default_rng(21)
    d, Dl, Dr = dims
    rnd = (lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if dtype == "c128" else (lambda *s: rng.standard_normal(s))
    W = P.mpo.asep_mpo(6, PRM)[0][2].astype(np.complex128 if dtype == "c128" else np.float64)
    w = W.shape[0]
    A = c.dev(rnd(d, Dl, Dr))
    for direction, blk, n in (("right", c.dev(rnd(Dl, w, Dl)), Dr), ("left", c.dev(rnd(Dr, w, Dr)), Dl)):
        full = (D.env_update_right if direction == "right" else D.env_update_left)(A, W, blk, None, c.ws)
        out = torch.full_like(full, float("nan"))
        cuts = [0, 16, 16, min(48, n - 8), n]                 # one empty range
        for k0, k1 in zip(cuts[:-1], cuts[1:]):
            D.env_update_cols(direction, A, W, blk, None, k0, k1 - k0, c.ws, out=out)
        torch.cuda.synchronize()
        assert not bool(torch.isnan(torch.view_as_real(out) if out.is_complex() else out).any())
        err = (out - full).abs().max().item() / full.abs().max().item()
        assert err < 1e-13, (direction, err)


def test_gauge_utilities_on_device(P):
    """gauge.make_mps_right / move_gauge / calc_entanglement_all on the device against the host (NumPy) routines."""
    from pydmrg_b200 import gauge, mps as mps_tools
    c = P.core.Context("c128")
    N = 10
    np.random.seed(8)
    mps = [t + 0.2j * np.random.rand(*t.shape) for t in mps_tools.create_rand_mps(N, 24)]

    def dense(m):
        v = np.asarray(m[0])[:, 0, :]
        for t in m[1:]:
            v = np.einsum("...a,pab->...pb", v, np.asarray(t))
        return v[..., 0].reshape(-1)

    psi0 = dense(mps)
    right = gauge.make_mps_right([t.copy() for t in mps], ctx=c)
    assert np.allclose(dense(right), psi0, atol=1e-12 * np.abs(psi0).max())
    for t in right[1:]:
        assert np.allclose(np.einsum("pab,pcb->ac", t, t.conj()), np.eye(t.shape[1]), atol=1e-12)
    moved = gauge.move_gauge([c.dev(t) for t in right], 0, 6, ctx=c)          # device in, device out
    assert all(hasattr(t, "is_cuda") and t.is_cuda for t in moved)
    assert np.allclose(dense([c.host(t) for t in moved]), psi0, atol=1e-12 * np.abs(psi0).max())
    EE = gauge.calc_entanglement_all([[t.copy() for t in right]], gSite=0, ctx=c)
    psi = psi0 / np.linalg.norm(psi0)
    for b in range(N - 1):
        s = np.linalg.svd(psi.reshape(2 ** (b + 1), -1), compute_uv=False)
        p = s[s > 1e-15] ** 2
        assert abs(EE[0, b] - np.sum(-p * np.log2(p))) < 1e-9


@pytest.mark.parametrize("dtype", ["c128", "f64"])
def test_diagonal_preconditioner_kernels_and_loops(P, dtype):
    """pdmrg_heff_diag against the diagonal of the dense operator, pdmrg_vec_precond against NumPy, and both Davidson
    loops with diag=...: same eigenvalue as without, fewer mat-vecs on a physical local problem."""
    import torch
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import davidson as dav, device as D
    c = P.core.Context(dtype)
    N, chi = 10, 16
    mpo = P.mpo.heisenberg_mpo(N) if dtype == "f64" else P.mpo.asep_mpo(N, PRM)
    np.random.seed(5)
    mps = orc.make_mps_right(orc.random_mps(N, chi))
    env = orc.build_env(mps, mpo, chi, gauge_site=0, mode="blas")
    site = N // 2 - 1
    M = [[c.dev(np.real(t) if dtype == "f64" else t) for t in mps]]
    F = [[c.dev(np.real(t) if dtype == "f64" else t) for t in Fm] for Fm in env]
    plan, shape = P.eigs.make_plan(M[0], mpo, F, site, False, c)
    diag = plan.diag(-1.0)
    Ls = [Fm[site] for Fm in env]; Rs = [Fm[site + 2] for Fm in env]
    dense = -orc.heff_dense_twosite(Ls, [op[site] for op in mpo], [op[site + 1] for op in mpo], Rs, shape)
    assert np.max(np.abs(c.host(diag) - np.diag(dense))) < 1e-12 * np.abs(dense).max()
    # the element-wise kernel
    ops = D.VecOps(c.code, plan.n, c.device)
    rng = np.random.default_rng(0)
    r = rng.standard_normal(plan.n) + (1j * rng.standard_normal(plan.n) if dtype == "c128" else 0)
    theta = -1.3 + (0.2j if dtype == "c128" else 0)
    rd = c.dev(r)
    n2 = torch.zeros(1, dtype=torch.float64, device=c.device)
    ops.precond(rd, diag, theta, 0.1, n2)
    den = np.diag(dense) - theta
    den = np.where(np.abs(den) < 0.1, 0.1 * den / np.abs(den), den)
    assert np.max(np.abs(c.host(rd) - r / den)) < 1e-13 * np.abs(r / den).max()
    assert abs(n2.item() - np.linalg.norm(r / den) ** 2) < 1e-12 * n2.item()
    # both loops, with and without the preconditioner
    guess = P.eigs.two_site_guess(M[0][site], M[0][site + 1], c).reshape(-1)
    res = {}
    for native in (False, True):
        for use in (False, True):
            st = {}
            e, _ = dav.davidson(lambda x, y: plan.apply(x, y, -1.0), [guess.clone()], plan.n, c.code, c.device, nroots=1, stats=st,
                                plan=plan, native=native, res_tol=1e-10, diag=diag if use else None)
            res[(native, use)] = (e[0], st["n_matvec"])
    ref = res[(False, False)][0]
    for key, (e, nmv) in res.items():
        assert abs(e - ref) < 1e-9 * abs(ref), key
    # the preconditioner must not cost mat-vecs (it saves them on cold starts; this guess is already close, and on the
    # B200 the c128 ASEP problem took 97 vs 96 — within one cycle)
    for native in (False, True):
        assert res[(native, True)][1] <= res[(native, False)][1] + 2, res


@pytest.mark.parametrize("dtype", ["c128", "f64"])
@pytest.mark.parametrize("m", [0, 5, 16, 17, 18, 28, 40])
def test_project_with_more_than_sixteen_vectors(P, dtype, m):
    """pdmrg_vec_project in passes of 16 basis vectors (subspaces of nroots >= 2 or max_space > 12):
    r <- (r - sum_i c_i X_i) / sqrt(scale), norm2 = ||r||^2 of the result."""
    import torch
    from pydmrg_b200 import device as D
    c = P.core.Context(dtype)
    n = 3001
    rng = np.random.default_rng(m)
    rnd = (lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if dtype == "c128" else (lambda *s: rng.standard_normal(s))
    X, r, cf = rnd(max(m, 1), n), rnd(n), rnd(max(m, 1))
    ops = D.VecOps(c.code, n, c.device)
    Xd, rd, cd_ = c.dev(X), c.dev(r), c.dev(cf)
    scale = torch.tensor([2.5], dtype=torch.float64, device=c.device)
    n2 = torch.zeros(1, dtype=torch.float64, device=c.device)
    ops.project(Xd, m, cd_, rd, scale_dev=scale, norm2_out=n2)
    ref = (r - (cf[:m] @ X[:m] if m else 0)) / np.sqrt(2.5)
    assert np.max(np.abs(c.host(rd) - ref)) < 1e-12 * np.abs(ref).max()
    assert abs(n2.item() - np.linalg.norm(ref) ** 2) < 1e-11 * np.linalg.norm(ref) ** 2


@pytest.mark.parametrize("model", ["heisenberg_f64", "asep_c128"])
def test_gauge_shortcut_and_solver_options(P, model):
    """run_dmrg(twoSite=True) with gauge_shortcut / max_space / precond against the plain run on the device: same
    energy (1e-10) and entropy (1e-8); the shortcut is taken, the larger subspace and the preconditioner do not cost
    mat-vecs."""
    if model == "heisenberg_f64":
        mpo, dtype, chi = P.mpo.heisenberg_mpo(18), "f64", 64
    else:
        mpo, dtype, chi = P.mpo.asep_mpo(18, PRM), "c128", 64
    res = {}
    for tag, opts in (("plain", {}), ("shortcut", {"gauge_shortcut": True}), ("space", {"max_space": 24}), ("precond", {"precond": "diag"})):
        np.random.seed(2)
        st = {}
        out = P.dmrg.run_dmrg(mpo, mbd=chi, tol=0.0, maxIter=6, twoSite=True, dtype=dtype, stats=st, res_tol=1e-10, **opts)
        res[tag] = (out[0], out[1], st)
    E0, EE0 = res["plain"][0], res["plain"][1]
    for tag in ("shortcut", "space", "precond"):
        assert abs(res[tag][0] - E0) < 1e-10 * abs(E0), tag
        assert abs(res[tag][1] - EE0) < 1e-8, tag
    assert res["shortcut"][2].get("n_gauge_moves", 0) > 0
    assert res["space"][2]["n_matvec"] <= 1.05 * res["plain"][2]["n_matvec"]
    assert res["precond"][2]["n_matvec"] <= 1.05 * res["plain"][2]["n_matvec"]
