"""CPU run of the HOST logic of the moving-centre truncation (pydmrg_b200/truncate.py): the device entry points
are replaced by small torch-CPU emulations of their documented semantics, so the control flow (exact first
visits, start policy, warm-started step, bookkeeping of the spare vectors and conjugations, tensor layouts)
executes without a GPU.  The CUDA kernels themselves are covered by the ``-m gpu`` tests."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


from emu_device import _Ctx, _permute4, _gemm_nt, _left_basis, _svd, _recycle_basis, _heevd, install  # noqa: F401


@pytest.fixture
def T(monkeypatch):
    from pydmrg_b200 import truncate
    install(monkeypatch.setattr)
    return truncate


def _psi_and_m(rng, Dl, Dr, decay):
    rows, cols = 2 * Dl, 2 * Dr
    U, _ = np.linalg.qr(rng.standard_normal((rows, rows)) + 1j * rng.standard_normal((rows, rows)))
    V, _ = np.linalg.qr(rng.standard_normal((cols, cols)) + 1j * rng.standard_normal((cols, cols)))
    q = min(rows, cols)
    s = np.exp(-decay * np.arange(q)); s /= np.linalg.norm(s)
    m = (U[:, :q] * s) @ V[:, :q].conj().T
    psi = m.reshape(Dl, 2, 2, Dr).transpose(1, 2, 0, 3).reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(psi)), m, s


def _mat(A, B):
    d1, Dl, k = A.shape
    d2, _, Dr = B.shape
    Am = A.numpy().transpose(1, 0, 2).reshape(Dl * d1, k)
    Bm = B.numpy().transpose(1, 0, 2).reshape(k, d2 * Dr)
    return Am, Bm


def test_truncate_twosite_visits_and_start_policy(T):
    """right (exact, first visit) -> left (exact: truncating state, policy wants two exact visits) -> right
    (warm-started) -> left (warm-started): every step returns the isometry on the side left behind and the
    exact projection of psi as centre; the warm-started steps reproduce the optimal discarded weight."""
    rng = np.random.default_rng(0)
    ctx = _Ctx()
    Dl = Dr = 64                                     # rows = cols = 128: the projective path is active
    chi = 96
    psi, m, s = _psi_and_m(rng, Dl, Dr, 0.12)        # discarded weight ~1e-10 at k = 96: "truncating"
    shape = (2, 2, Dl, Dr)
    cache, stats = {}, {}
    w_opt = float(np.sum(s[chi:] ** 2))
    A = B = None
    for visit, direction in enumerate(["right", "left", "right", "left"]):
        warm = None if A is None else (A, B)
        A, B, Sn, EE, EEs = T.truncate_twosite(psi, shape, chi, ctx, direction, recycle=cache, site=3, warm=warm, stats=stats)
        Am, Bm = _mat(A, B)
        iso = Am if direction == "right" else Bm.conj().T
        assert np.max(np.abs(iso.conj().T @ iso - np.eye(chi))) < 1e-12
        proj = iso @ (iso.conj().T @ m) if direction == "right" else (m @ iso) @ iso.conj().T
        assert np.max(np.abs(Am @ Bm - proj)) < 1e-12
        w = np.linalg.norm(m) ** 2 - np.linalg.norm(Am @ Bm) ** 2
        assert abs(w - w_opt) < 1e-13
        # the next visit of this bond is seeded with the spare rows of the side that was left behind
        other = "left" if direction == "right" else "right"
        assert tuple(cache[(3, other)].shape) == (T.spare_count(chi, 128, 128), 128)
    assert stats.get("n_exact_trunc") == 2 and stats.get("n_recycled") == 2
    assert cache[("state", 3)]["exact_visits"] == 2


def test_truncate_twosite_immediate_recycling_when_nothing_is_discarded(T):
    rng = np.random.default_rng(1)
    ctx = _Ctx()
    Dl = Dr = 64
    chi = 96
    psi, m, s = _psi_and_m(rng, Dl, Dr, 0.6)         # spectrum below 1e-16 long before k = 96
    cache, stats = {}, {}
    A, B, *_ = T.truncate_twosite(psi, (2, 2, Dl, Dr), chi, ctx, "right", recycle=cache, site=0, warm=None, stats=stats)
    A, B, *_ = T.truncate_twosite(psi, (2, 2, Dl, Dr), chi, ctx, "left", recycle=cache, site=0, warm=(A, B), stats=stats)
    assert stats.get("n_exact_trunc") == 1 and stats.get("n_recycled") == 1
    Am, Bm = _mat(A, B)
    assert np.max(np.abs(Am @ Bm - m)) < 1e-12


def test_full_rank_split_logic(T):
    rng = np.random.default_rng(2)
    ctx = _Ctx()
    for Dl, Dr, direction in ((32, 100, "right"), (100, 32, "right"), (32, 100, "left"), (100, 32, "left")):
        psi, m, s = _psi_and_m(rng, Dl, Dr, 0.05)
        A, B, Sn, EE, EEs = T.full_rank_split(psi, (2, 2, Dl, Dr), ctx, direction)
        Am, Bm = _mat(A, B)
        assert np.max(np.abs(Am @ Bm - m)) < 1e-12
        iso = Am if direction == "right" else Bm.conj().T
        assert np.max(np.abs(iso.conj().T @ iso - np.eye(iso.shape[1]))) < 1e-12


def test_two_site_sweeps_host_flow_with_emulated_device(T, monkeypatch):
    """dmrg.twoSiteSweep end to end on the CPU: local solves and block updates come from the NumPy oracle, the
    truncation runs the product's host logic over the emulated device entry points.  Checks the plumbing the GPU
    tests cannot see from the outside: bonds are decomposed exactly until the start policy allows the warm start,
    the centre bond stays exact, and recycled sweeps land on the energy of sweeps that never recycle."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import dmrg, device as D
    mpo = orc.heisenberg_mpo(16)          # chi = 64: bonds 6, 7 (centre), 8 are 128 x 128 cuts, the rest below 128
    ctx = _Ctx()
    ctx.tdtype = torch.complex128

    def calc_eigs(mpsL, W, F, site, nStates, alg="davidson", oneSite=True, ctx=None, stats=None, **kw):
        A0, B0 = mpsL[0][site].numpy(), mpsL[0][site + 1].numpy()
        shape = (A0.shape[0], B0.shape[0], A0.shape[1], B0.shape[2])
        Ls = [Fm[site].numpy() for Fm in F]; Rs = [Fm[site + 2].numpy() for Fm in F]
        mv = lambda x: orc.heff_twosite(x, Ls, [op[site] for op in W], [op[site + 1] for op in W], Rs, shape, "blas")
        guess = np.einsum("ijk,lkm->iljm", A0, B0).reshape(-1)
        E, vecs, _ = orc.local_eigs(mv, [guess], 1, site, len(mpsL[0]), alg="davidson", check=False)
        return E[0], torch.from_numpy(np.ascontiguousarray(vecs[:, :1].astype(np.complex128))), None

    monkeypatch.setattr(dmrg, "calc_eigs", calc_eigs)
    monkeypatch.setattr(D, "env_update_right", lambda A, W, L, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_right(A.numpy(), W, L.numpy(), None, "blas").astype(np.complex128))))
    monkeypatch.setattr(D, "env_update_left", lambda B, W, R, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_left(B.numpy(), W, R.numpy(), None, "blas").astype(np.complex128))))
    ctx.ws = None
    ctx.dev = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex128)))
    res = {}
    for rec in (False, True):
        np.random.seed(3)
        mps = orc.make_mps_right(orc.random_mps(16, 64))
        env = orc.build_env(mps, mpo, 64, gauge_site=0, mode="blas")
        mpsL = [[ctx.dev(t) for t in mps]]
        F = [[ctx.dev(t) for t in Fm] for Fm in env]
        cache = {} if rec else None
        stats_all = []
        for sw in range(3):
            st = {}
            E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, 64, ctx=ctx, stats=st, recycle=cache)
            stats_all.append(st)
        res[rec] = (complex(E), float(EE), stats_all)
    (E0, EE0, _), (E1, EE1, st1) = res[False], res[True]
    assert abs(E1 - E0) < 1e-10 * abs(E0) and abs(EE1 - EE0) < 1e-8
    assert st1[0].get("n_recycled", 0) + st1[1].get("n_recycled", 0) + st1[2].get("n_recycled", 0) > 0
    assert all(s.get("n_exact_trunc", 0) >= 2 for s in st1)          # the centre bond, both directions, every sweep


# ---------------------------------------------------------------------------------------
# several target states, two-site
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("direction", ["right", "left"])
@pytest.mark.parametrize("spectrum", [True, False])
def test_state_averaged_twosite_truncation_logic(T, direction, spectrum):
    """truncate.rdm_truncate_twosite_states over the emulated entry points against its NumPy restatement
    (oracle/multistate_oracle.py): shared isometry, per-state centres = exact projections, same spectrum."""
    from oracle import multistate_oracle as ms
    rng = np.random.default_rng(4)
    ctx = _Ctx()
    shape = (2, 2, 6, 5)
    k = 7
    psis = [rng.standard_normal(120) + 1j * rng.standard_normal(120) for _ in range(2)]
    As, Bs, Sn, EE, EEs = T.rdm_truncate_twosite_states([torch.from_numpy(p.copy()) for p in psis], shape, k, ctx, direction,
                                                        spectrum=spectrum)
    Ao, Bo, Sno, EEo, EEso = ms.rdm_truncate_twosite_states(psis, shape, k, direction)
    for s in range(2):
        Am, Bm = _mat(As[s], Bs[s])
        Amo = Ao[s].transpose(1, 0, 2).reshape(12, k)
        Bmo = Bo[s].transpose(1, 0, 2).reshape(k, 10)
        iso = Am if direction == "right" else Bm.conj().T
        assert np.max(np.abs(iso.conj().T @ iso - np.eye(k))) < 1e-12
        assert np.max(np.abs(Am @ Bm - Amo @ Bmo)) < 1e-12          # gauge-invariant: the projected wave-function
    assert (As[0] is As[1]) if direction == "right" else (Bs[0] is Bs[1])
    if spectrum:
        assert abs(EE - EEo) < 1e-12 and np.max(np.abs(Sn - Sno)) < 1e-12
    else:
        assert np.isfinite(EE) and abs(np.sum(Sn ** 2) - 1) < 1e-12


def test_two_site_two_state_sweeps_host_flow(T, monkeypatch):
    """dmrg.twoSiteSweep(nStates=2) on the CPU (local solves and block updates from the oracle, the
    state-averaged truncation through the product's host logic): both energies and the gap land on the NumPy
    restatement's and on dense ED (fixture from the reference's tools/ed.py)."""
    import os
    from oracle import dmrg_oracle as orc, multistate_oracle as ms
    from pydmrg_b200 import dmrg, device as D, mpo as M
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "multistate.npz"))
    N, chi, seed = (int(v) for v in g["asep8_cfg"])
    mpo = M.asep_mpo(N, tuple(g["params"]))
    ctx = _Ctx()
    ctx.ws = None
    ctx.dev = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex128)))

    def calc_eigs(mpsL, W, F, site, nStates, alg="davidson", oneSite=True, ctx=None, stats=None, **kw):
        A0, B0 = mpsL[0][site].numpy(), mpsL[0][site + 1].numpy()
        shape = (A0.shape[0], B0.shape[0], A0.shape[1], B0.shape[2])
        Ls = [Fm[site].numpy() for Fm in F]; Rs = [Fm[site + 2].numpy() for Fm in F]
        mv = lambda x: orc.heff_twosite(x, Ls, [op[site] for op in W], [op[site + 1] for op in W], Rs, shape, "blas")
        guesses = [np.einsum("ijk,lkm->iljm", mpsL[s][site].numpy(), mpsL[s][site + 1].numpy()).reshape(-1) for s in range(nStates)]
        E, vecs, _ = orc.local_eigs(mv, guesses, nStates, site, len(mpsL[0]), alg="davidson", edge_preserve_state=False)
        return E, torch.from_numpy(np.ascontiguousarray(vecs.astype(np.complex128))), None

    monkeypatch.setattr(dmrg, "calc_eigs", calc_eigs)
    monkeypatch.setattr(D, "env_update_right", lambda A, W, L, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_right(A.numpy(), W, L.numpy(), None, "blas").astype(np.complex128))))
    monkeypatch.setattr(D, "env_update_left", lambda B, W, R, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_left(B.numpy(), W, R.numpy(), None, "blas").astype(np.complex128))))
    np.random.seed(seed)
    states = ms.random_states(N, chi, 2)
    env = orc.build_env(states[0], mpo, chi, gauge_site=0, mode="blas")
    mpsL = [[ctx.dev(t) for t in st] for st in states]
    F = [[ctx.dev(t) for t in Fm] for Fm in env]
    for sw in range(4):
        E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, nStates=2)
    np.random.seed(seed)
    Eo, EEo = ms.run_dmrg_twosite_states(mpo, chi, 2, n_sweeps=4)[:2]
    ed = g["asep8_ed"]
    assert abs(E[0] - ed[0]) < 1e-7 * abs(ed[0]) and abs(E[1] - ed[1]) < 1e-7 * abs(ed[0])
    assert abs(E[0] - Eo[0]) < 1e-9 and abs(E[1] - Eo[1]) < 1e-9 and abs(EE - EEo) < 1e-7
    for site in range(1, N):                                   # isometries are shared, the centre (site 0) is not
        assert mpsL[0][site] is mpsL[1][site] or torch.equal(mpsL[0][site], mpsL[1][site])


def test_recycled_weight_check_and_rdm_spectrum_length():
    """(a) a warm-started step whose discarded weight exceeds twice that of the bond's last exact decomposition
    is rejected (the caller redoes the bond exactly); (b) the one-site RDM spectrum has min(d*Dl, Dr) entries, like the
    reference's calc_ent_right / calc_ent_left (dmrg.py:30-47), not d*Dl."""
    from pydmrg_b200 import truncate as T
    cache = {}
    T.note_exact_visit(cache, 5, (64, 64, 32), 1e-9)
    assert T.recycled_weight_ok(cache, 5, 1.5e-9) and not T.recycled_weight_ok(cache, 5, 3e-9)
    assert T.recycled_weight_ok(cache, 6, 1.0)                              # no exact visit on record: nothing to compare with
    T.note_exact_visit(cache, 7, (64, 64, 32), 0.0)
    assert T.recycled_weight_ok(cache, 7, 5e-14) and not T.recycled_weight_ok(cache, 7, 1e-12)
    w = np.sort(np.concatenate([np.zeros(12), np.array([0.5, 0.3, 0.15, 0.05])]))     # ascending, as heevd returns them
    EE, spec, S = T._spectrum_from_rdm(w, 4)
    assert len(spec) == 4 and len(S) == 4 and abs(np.sum(S ** 2) - 1) < 1e-14
    p = np.array([0.5, 0.3, 0.15, 0.05])
    assert abs(EE - np.sum(-p * np.log2(p))) < 1e-14
