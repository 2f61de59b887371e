"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py).  Run in the build container:

    python tests/golden/make_golden.py

The reference holds no golden vectors of its own (its tests are unseeded differential
checks), so these seeded outputs of the reference ARE the pins.
Every array here is produced by a reference function named in the key's comment.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

ref = ref_loader.load()
quiet = lambda: contextlib.redirect_stdout(io.StringIO())


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def kernels_fixture():
    """Component-level: reference Hfun (one/two-site), update_envR/L, renorm_inf on seeded
    random complex tensors (SURVEY 8d synthetic definition), ragged bond dims included."""
    out = {}
    cases = [  # name, d, Dl, Dr, asep params
        ("a", 2, 8, 8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)),
        ("b", 2, 5, 7, (0.35, 0., 1., 0., 2. / 3., 0., -0.5)),
        ("c", 2, 16, 12, (0.2, 0.8, 0.1, 0.9, 0.4, 0.6, 0.3)),
        ("d", 2, 1, 2, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)),
    ]
    for ci, (name, d, Dl, Dr, prm) in enumerate(cases):
        rng = np.random.default_rng(100 + ci)
        mpo = ref.asep.return_mpo(6, prm)
        W1, W2 = mpo[0][2], mpo[0][3]
        w = W1.shape[0]
        out[name + "_params"] = np.array(prm)
        out[name + "_W1"], out[name + "_W2"] = W1, W2
        # ---- one-site Hfun: tools/diag_tools.py:108-125
        L, R = crand(rng, Dl, w, Dl), crand(rng, Dr, w, Dr)
        x1 = crand(rng, d, Dl, Dr)
        M = [x1]
        Hfun, _ = ref.diag_tools.make_ham_func_oneSite(M, [[W1]], [[L, R]], 0)
        out[name + "_L"], out[name + "_R"], out[name + "_x1"] = L, R, x1
        out[name + "_y1"] = Hfun(x1.ravel())
        # ---- two-site Hfun: tools/diag_tools.py:146-162  (env slots 0 and 2 around sites 0,1)
        x2 = crand(rng, d, d, Dl, Dr)
        Dm = 4
        M2 = [np.zeros((d, Dl, Dm)), np.zeros((d, Dm, Dr))]
        Hfun2, _ = ref.diag_tools.make_ham_func_twoSite(M2, [[W1, W2]], [[L, R]], 0)
        out[name + "_x2"] = x2
        out[name + "_y2"] = Hfun2(x2.ravel())
        # ---- update_envR / update_envL: tools/env_tools.py:40-50, 28-38 (default bra = conj)
        A = crand(rng, d, Dl, Dr)
        F = [[L, None]]
        with quiet():
            F = ref.env_tools.update_envR([A], [[W1]], F, 0)
        out[name + "_A"], out[name + "_envR_out"] = A, F[0][1]
        F = [[None, R]]
        with quiet():
            F = ref.env_tools.update_envL([A], [[W1]], F, 0)
        out[name + "_envL_out"] = F[0][0]
        # explicit (non-conjugate) bra, as used for <l|O|r> (tools/contract.py:38-39)
        Abra = crand(rng, d, Dl, Dr)
        F = [[L, None]]
        F = ref.env_tools.update_envR([A], [[W1]], F, 0, Ml=[Abra])
        out[name + "_Abra"], out[name + "_envR_bra_out"] = Abra, F[0][1]
        F = [[None, R]]
        F = ref.env_tools.update_envL([A], [[None]], F, 0, Ml=[Abra])   # None MPO entry = identity
        out[name + "_envL_none_out"] = F[0][0]
        # ---- renorm_inf: tools/mps_tools.py:98-120
        mbd = max(2, min(Dl, Dr))
        mps = [[np.zeros((d, Dl, Dm)), np.zeros((d, Dm, Dr))]]
        env = [[np.zeros((Dl, w, Dl)), np.zeros((Dr, w, Dr))]]
        mps, EE, EEs, S = ref.mps_tools.renorm_inf(mps, [[W1, W2]], env, x2.reshape(-1, 1).copy(), mbd)
        out[name + "_svd_mbd"] = np.array(mbd)
        out[name + "_svd_S"], out[name + "_svd_EE"], out[name + "_svd_EEs"] = S, np.array(EE), EEs
        out[name + "_svd_A"], out[name + "_svd_B"] = mps[0][0], mps[0][1]
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels.npz", len(out), "arrays")


def davidson_fixture():
    """Reference Davidson (tools/la_tools.py:315-557 via tools/diag_tools.py:243-290) on a
    seeded one-site problem; records eigenvalue, residual and iteration stats."""
    out = {}
    rng = np.random.default_rng(7)
    d, Dl, Dr = 2, 6, 6
    mpo = ref.asep.return_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    W = mpo[0][2]
    w = W.shape[0]
    L, R = crand(rng, Dl, w, Dl) * 0.3, crand(rng, Dr, w, Dr) * 0.3
    x0 = crand(rng, d, Dl, Dr)
    M = [np.zeros((d, 1, Dl)), x0, np.zeros((d, Dr, 1))]
    F = [[None, L, R, None]]
    count = [0]
    Hfun, precond = ref.diag_tools.make_ham_func_oneSite(M, [[None, W, None]], F, 1)
    def counted(x):
        count[0] += 1
        return Hfun(x)
    vals, vec = ref.la_tools.eig(counted, [x0.ravel()], precond, nroots=1, pick=ref.diag_tools.pick_eigs,
                                 follow_state=False, tol=1e-16, max_cycle=1000)
    H = ref.diag_tools.calc_ham_oneSite(M, [[None, W, None]], F, 1)
    exact = np.linalg.eigvals(-H)
    out["L"], out["R"], out["W"], out["x0"] = L, R, W, x0
    out["eval"] = np.array(vals)
    out["evec"] = vec
    out["n_matvec"] = np.array(count[0])
    out["exact_min_real"] = exact[np.argmin(exact.real)]
    E, vecs, ovlp = ref.diag_tools.calc_eigs(M and [M], [[None, W, None]], F, 1, 1, alg="davidson")
    out["calc_eigs_E"] = np.array(E)
    np.savez_compressed(os.path.join(HERE, "davidson.npz"), **out)
    print("davidson.npz: eval", vals, "n_mv", count[0], "exact", out["exact_min_real"])


def runs_fixture():
    """End-to-end reference runs (run_dmrg dmrg.py:309, run_idmrg idmrg.py:133, ED tools/ed.py:4)."""
    out = {}
    # config 0: TASEP N=10 chi=10 s=-0.5 (BASELINE.json configs[0]); params = SURVEY 8c choice
    prm = (0.35, 0., 1., 0., 2. / 3., 0., -0.5)
    for alg in ("davidson", "exact"):
        np.random.seed(0)
        mpo = ref.asep.return_mpo(10, prm)
        with quiet():
            r = ref.dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10, alg=alg, returnEntSpec=True)
        out["cfg0_%s_E" % alg], out["cfg0_%s_EE" % alg], out["cfg0_%s_EEs" % alg] = np.array(r[0]), np.array(r[1]), r[3]
    out["cfg0_params"] = np.array(prm)
    # ED known answers (chi exact)
    for tag, prm6 in (("tasep6", prm), ("asep6", (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))):
        mpo = ref.asep.return_mpo(6, prm6)
        u, _ = ref.ed.ed(mpo)
        out[tag + "_ed_top"] = np.array(u[0])
        np.random.seed(3)
        with quiet():
            r = ref.dmrg.run_dmrg(mpo, mbd=8, tol=1e-10, maxIter=6, returnEntSpec=True, returnState=True)
        out[tag + "_dmrg_E"], out[tag + "_dmrg_EE"], out[tag + "_dmrg_EEs"] = np.array(r[0]), np.array(r[1]), r[3]
        out[tag + "_params"] = np.array(prm6)
        # observables on the returned state: current via curr_mpo (tests/asep/currentCheck.py:29-44)
        mps = r[4]
        cmpo = ref.asep.curr_mpo(6, prm6)
        with quiet():
            num = ref.contract.full_contract(mps=mps, mpo=cmpo)
            den = ref.contract.full_contract(mps=mps)
            hnum = ref.contract.full_contract(mps=mps, mpo=mpo)
        out[tag + "_current_rr"] = np.array(num / den)
        out[tag + "_energy_rr"] = np.array(hnum / den)
        for i, M in enumerate(mps[0]):
            out[tag + "_mps%d" % i] = M
    # a slightly larger run, one-site: N=12 chi=16
    prm12 = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)
    np.random.seed(1)
    mpo = ref.asep.return_mpo(12, prm12)
    with quiet():
        r = ref.dmrg.run_dmrg(mpo, mbd=16, tol=1e-10, maxIter=3, returnEntSpec=True)
    out["asep12_E"], out["asep12_EE"], out["asep12_EEs"] = np.array(r[0]), np.array(r[1]), r[3]
    # (calcLeftState needs h5py file round trips in the reference -- dmrg.py:455 builds fname+'_left' --
    #  so the left-eigenstate path is checked against <l|H|r>/<l|r> instead, tests/test_sweeps_gpu.py)
    # iDMRG: idmrg.py:133
    np.random.seed(5)
    mpo4 = ref.asep.return_mpo(4, (0.2, 0.8, 0.1, 0.9, 0.4, 0.6, -0.5))
    with quiet():
        r = ref.idmrg.run_idmrg(mpo4, mbd=10, maxIter=20, returnEntSpec=True)
    out["idmrg_E"], out["idmrg_EE"], out["idmrg_EEs"] = np.array(r[0]), np.array(r[1]), r[3]
    np.savez_compressed(os.path.join(HERE, "runs.npz"), **out)
    for k in sorted(out):
        if out[k].size <= 2:
            print(k, out[k])


def multistate_fixture():
    """Reference run_dmrg with nStates=2 (dmrg.py:309; state-averaged RDM :56-66, block Davidson with two roots,
    gap = E[0] - E[1] :295-296) on seeded starts, plus the two leading ED eigenvalues (tools/ed.py:4)."""
    out = {}
    prm = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)
    for tag, N, mbd, seed in (("asep6", 6, 8, 6), ("asep8", 8, 16, 4)):
        mpo = ref.asep.return_mpo(N, prm)
        u, _ = ref.ed.ed(mpo)
        np.random.seed(seed)
        with quiet():
            r = ref.dmrg.run_dmrg(mpo, mbd=mbd, tol=1e-10, maxIter=6, nStates=2, returnEntSpec=True)
        out[tag + "_E"], out[tag + "_EE"], out[tag + "_gap"], out[tag + "_EEs"] = np.array(r[0]), np.array(r[1]), np.array(r[2]), r[3]
        out[tag + "_ed"] = np.array(u[:2])
        out[tag + "_cfg"] = np.array([N, mbd, seed])
    out["params"] = np.array(prm)
    np.savez_compressed(os.path.join(HERE, "multistate.npz"), **out)
    for k in sorted(out):
        if out[k].size <= 3:
            print(k, out[k])


def leftright_fixture():
    """Reference run_dmrg(calcLeftState=True) (dmrg.py:338, 448-488: second DMRG on mpo_conj_trans(mpo); needs
    fname, i.e. the HDF5 round trip -- served by the h5py mini-shim of oracle/ref_loader.py where h5py is absent),
    observables <l|O|r>/<l|r> through tools/contract.py:5-44 as the scan scripts compute them
    (scripts/asep/current/multipleStates.py:89-101), and a chi schedule continued through saved states."""
    import tempfile
    out = {}
    cases = (("tasep8", 8, (0.35, 0., 1., 0., 2. / 3., 0., -0.5), 9), ("asep8", 8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5), 10))
    for tag, N, prm, seed in cases:
        mpo = ref.asep.return_mpo(N, prm)
        cmpo = ref.asep.curr_mpo(N, prm)
        tmp = tempfile.mkdtemp()
        np.random.seed(seed)
        with quiet():
            r = ref.dmrg.run_dmrg(mpo, mbd=16, tol=1e-10, maxIter=6, fname=os.path.join(tmp, "st"), calcLeftState=True,
                                  returnState=True, returnEntSpec=True)
            # The two lists returned in memory are ONE aliased object (dmrg.py:371 passes mpsList itself to
            # make_all_mps_right, so the left run overwrites the right state); the files written after each run
            # hold the two states, and they are what the scan scripts contract.
            assert r[4][0] is r[4][1]
            mps, _ = ref.mps_tools.load_mps(os.path.join(tmp, "st_mbd0"))
            mpsl, _ = ref.mps_tools.load_mps(os.path.join(tmp, "st_mbd0_left"))
            num = ref.contract.full_contract(mps=mps, mpo=cmpo, lmps=mpsl)
            den = ref.contract.full_contract(mps=mps, lmps=mpsl)
            hnum = ref.contract.full_contract(mps=mps, mpo=mpo, lmps=mpsl)
        u, _ = ref.ed.ed(mpo)
        out[tag + "_E"], out[tag + "_EE"], out[tag + "_EEl"] = np.array(r[0]), np.array(r[1][0]), np.array(r[1][1])
        out[tag + "_EEs"], out[tag + "_EEsl"] = r[3][0], r[3][1]
        out[tag + "_current_lr"], out[tag + "_energy_lr"] = np.array(num / den), np.array(hnum / den)
        out[tag + "_ed_top"] = np.array(u[0])
        out[tag + "_params"] = np.array(prm)
        out[tag + "_cfg"] = np.array([N, 16, seed])
    # chi schedule [4, 8, 16] continued through the saved states (dmrg.py:352-358, mps_tools.py:244-279)
    prm = cases[1][2]
    mpo = ref.asep.return_mpo(8, prm)
    tmp = tempfile.mkdtemp()
    np.random.seed(11)
    with quiet():
        r = ref.dmrg.run_dmrg(mpo, mbd=[4, 8, 16], tol=1e-10, maxIter=4, fname=os.path.join(tmp, "st"))
    out["sched_E"], out["sched_EE"] = np.array(r[0]), np.array(r[1])
    out["sched_cfg"] = np.array([8, 11])
    np.savez_compressed(os.path.join(HERE, "leftright.npz"), **out)
    for k in sorted(out):
        if out[k].size <= 3:
            print(k, out[k])


if __name__ == "__main__":
    if "multistate" in sys.argv[1:] or "leftright" in sys.argv[1:]:
        if "multistate" in sys.argv[1:]:
            multistate_fixture()
        if "leftright" in sys.argv[1:]:
            leftright_fixture()
        sys.exit(0)
    kernels_fixture()
    davidson_fixture()
    runs_fixture()
    multistate_fixture()
    leftright_fixture()
