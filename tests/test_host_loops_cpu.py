"""CPU runs of the product's HOST loops over emulated device entry points (tests/emu_device.py):

* the Python Davidson loop of pydmrg_b200/davidson.py (the specification the C++ ``pdmrg_davidson`` mirrors) on
  the reference's seeded problem (tests/golden/davidson.npz);
* sharded two-site sweeps (``ShardSpec``) on TWO gloo ranks: sharded mat-vec with one all-reduce, consensus
  broadcasts of the Davidson control scalars, column-sharded block updates -- with rank-dependent rounding noise
  injected, so that a control-flow divergence between the ranks would show up as a hang or a mismatch.

The CUDA kernels themselves are covered by the ``-m gpu`` tests."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def test_python_davidson_loop_on_the_reference_problem(monkeypatch, golden):
    """Same eigenvalue as the reference's Davidson (1e-9), same mat-vec count within a few cycles, eigenvector equal
    up to a phase; with a restart in between (232 mat-vecs > max_space)."""
    from emu_device import install
    from oracle import dmrg_oracle as orc
    install(monkeypatch.setattr)
    from pydmrg_b200 import davidson as dav
    dav.release_workspaces()
    g = golden["davidson"]
    L, R, W, x0 = g["L"], g["R"], g["W"], g["x0"]
    n = x0.size

    def matvec(x, y):
        y.copy_(torch.from_numpy(orc.heff_onesite(x.numpy(), [L], [W], [R], x0.shape)))

    st = {}
    e, vecs = dav.davidson(matvec, [torch.from_numpy(x0.ravel().copy())], n, 1, torch.device("cpu"), nroots=1, stats=st,
                           native=False)
    assert abs(e[0] - g["eval"]) < 1e-9 * abs(g["eval"])
    assert abs(st["n_matvec"] - int(g["n_matvec"])) <= max(3, int(0.05 * int(g["n_matvec"])))
    v, vr = vecs[0].numpy(), g["evec"]
    assert abs(abs(np.vdot(v, vr)) / (np.linalg.norm(v) * np.linalg.norm(vr)) - 1) < 1e-10
    # the consensus hook sees the small arrays twice per cycle and may not change the result when it is the identity
    calls = []
    dav.release_workspaces()
    e2, _ = dav.davidson(matvec, [torch.from_numpy(x0.ravel().copy())], n, 1, torch.device("cpu"), nroots=1, native=False,
                         consensus=lambda t: calls.append(t.numel()))
    assert e2[0] == e[0] and len(calls) == 2 * st["cycles"]
    dav.release_workspaces()


# ---------------------------------------------------------------------------------------------------------------
def _sweep_worker(rank, world, port, q, noise):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    torch.set_num_threads(1)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_device
    from emu_device import _Ctx
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import device as D, dmrg, eigs, parallel
    emu_device.install(setattr)
    N, chi = 12, 32
    mpo = orc.heisenberg_mpo(N)
    eps = noise * rank                                           # rank-dependent "last bits"
    calls = [0]                                                  # sharded mat-vecs = all-reduces issued by this rank

    class Plan:
        def __init__(self, n):
            self.n = n

        def close(self):
            pass

    def site_tensors(M, W, F, site):
        A0, B0 = M[site], M[site + 1]
        shape = (A0.shape[0], B0.shape[0], A0.shape[1], B0.shape[2])
        Ls = [Fm[site].numpy() for Fm in F]
        Rs = [Fm[site + 2].numpy() for Fm in F]
        return shape, Ls, [op[site] for op in W], [op[site + 1] for op in W], Rs

    def make_plan(M, W, F, site, oneSite, ctx, block_masks=None):
        shape, Ls, W1, W2, Rs = site_tensors(M, W, F, site)
        plan = Plan(int(np.prod(shape)))
        plan.apply = lambda x, y, sign: y.copy_(torch.from_numpy(-sign * orc.heff_twosite(x.numpy(), Ls, W1, W2, Rs, shape, "blas")))
        return plan, shape

    class Sharded:
        """ShardedHeff(mode='ket'): this rank's part of the mat-vec = H applied to x restricted to its range of the
        left ket index, then one all-reduce."""

        def __init__(self, M, W, F, site):
            self.shape, self.Ls, self.W1, self.W2, self.Rs = site_tensors(M, W, F, site)
            self.k0, self.kn = parallel.ket_chunks(self.shape[2], world)[rank]
            self.plan = Plan(int(np.prod(self.shape)))

        def apply(self, x, y, sign):
            calls[0] += 1
            xm = np.zeros(self.shape, complex)
            xm[:, :, self.k0:self.k0 + self.kn, :] = x.numpy().reshape(self.shape)[:, :, self.k0:self.k0 + self.kn, :]
            part = -sign * orc.heff_twosite(xm.reshape(-1), self.Ls, self.W1, self.W2, self.Rs, self.shape, "blas") * (1 + eps)
            y.copy_(torch.from_numpy(part))
            if world > 1:
                dist.all_reduce(torch.view_as_real(y))
            return y

    class Spec(parallel.ShardSpec):
        def heff(self, ctx, M, W, F, site, oneSite):
            return Sharded(M, W, F, site)

    def env_cols(direction, T, W, blk, Tbra, k0, kn, ws, out=None):
        f = orc.env_update_right if direction == "right" else orc.env_update_left
        full = f(T.numpy(), W, blk.numpy(), None, "blas").astype(complex) * (1 + eps)
        res = np.zeros_like(full)
        res[:, :, k0:k0 + kn] = full[:, :, k0:k0 + kn]
        return torch.from_numpy(res)

    eigs.make_plan = make_plan
    eigs.two_site_guess = lambda A, B, ctx: torch.einsum("ijk,lkm->iljm", A, B).contiguous()
    D.env_update_cols = env_cols
    D.env_update_right = lambda A, W, L, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_right(A.numpy(), W, L.numpy(), None, "blas").astype(complex) * (1 + eps)))
    D.env_update_left = lambda B, W, R, bra, ws: torch.from_numpy(np.ascontiguousarray(
        orc.env_update_left(B.numpy(), W, R.numpy(), None, "blas").astype(complex) * (1 + eps)))
    ctx = _Ctx()
    ctx.code, ctx.npdtype, ctx.ws = 1, np.complex128, None
    ctx.dev = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex128)))
    np.random.seed(3)
    mps = orc.make_mps_right(orc.random_mps(N, chi))
    env = orc.build_env(mps, mpo, chi, gauge_site=0, mode="blas")
    mpsL = [[ctx.dev(t) for t in mps]]
    F = [[ctx.dev(t) for t in Fm] for Fm in env]
    spec = Spec(rank, world, dist if world > 1 else None, min_dim=32) if world > 1 else None
    st = {}
    for sw in range(2):
        E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, stats=st, recycle=None, shard=spec, native=False,
                                                   res_tol=1e-9)
    if os.environ.get("PDMRG_TEST_DUMP") and rank == 0:
        import json
        json.dump({"res_history": st.get("res_history"), "n_matvec": st.get("n_matvec"), "cycles": st.get("cycles")},
                  open(os.environ["PDMRG_TEST_DUMP"] + ".w%d.n%g.json" % (world, noise), "w"))
    q.put((rank, complex(E), float(EE), st.get("n_sharded_solves", 0), st.get("n_sharded_env", 0), calls[0]))
    if world > 1:
        dist.destroy_process_group()


def _run(world, noise, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_sweep_worker, args=(r, world, port, q, noise)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict((r[0], r[1:]) for r in (q.get(timeout=600) for _ in ps))
    for p in ps:
        p.join(timeout=60)
    return res


def test_sharded_two_site_sweeps_stay_in_lockstep_on_two_gloo_ranks():
    """Two ranks run two sharded sweeps of a Heisenberg N=12 chi=32 chain; rank 1's mat-vec parts and block updates
    carry a relative 1e-13 perturbation (stand-in for different reduction orders).  Both ranks must finish (matched
    collectives: same cycle counts), report bit-identical E, and agree with the unsharded sweeps."""
    base = 32500 + (os.getpid() % 1500)
    single = _run(1, 0.0, base)[0]
    two = _run(2, 1e-13, base + 1)
    assert two[0][0] == two[1][0]                               # bit-identical energy on both ranks
    # every all-reduce of one rank met its counterpart: same number of sharded mat-vecs, solves and block updates
    # (the unsharded edge bonds run redundantly and may take a cycle more or less on a rank -- no collective there)
    assert two[0][2:] == two[1][2:] and two[0][2] > 0 and two[0][3] > 0 and two[0][4] > 0
    # rank 0's eigenpairs are adopted after EVERY solve, so 1e-13 noise stays 1e-13 (before that rule a flipped
    # convergence test on one rank left the ranks 1e-9 apart and cost 40 % more mat-vecs and 1e-8 in E)
    assert abs(two[0][0] - single[0]) < 1e-11 * abs(single[0])
    assert abs(two[0][1] - single[1]) < 1e-10


# ---------------------------------------------------------------------------------------------------------------
# s-field scan (BASELINE configs[3]) on two gloo ranks
# ---------------------------------------------------------------------------------------------------------------
def _scan_worker(rank, world, port, q, continuation):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_device
    emu_device.install(setattr)
    from pydmrg_b200 import mpo as M, scan
    s_vals = np.linspace(-0.5, 0.5, 5)
    mpo_of = lambda s: M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    np.random.seed(7 + rank)
    res = scan.run_scan(mpo_of, s_vals, 8, rank, world, dist, ramp=(4,), tol=1e-10, maxIter=4, continuation=continuation,
                        ctx=emu_device.Context("c128"), native=False, res_tol=1e-10)
    q.put((rank, res["E"], res["EE"], res["points"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("continuation", [False, True])
def test_s_scan_on_two_gloo_ranks(continuation):
    """pydmrg_b200.scan.run_scan: the points are dealt to two ranks (every second point for cold starts, contiguous
    blocks with warm starts inside a block for ``continuation``), no data-path collective, one gather; every rank
    ends with the full curve, which matches dense diagonalisation of the tilted generator at every s."""
    import torch.multiprocessing as mp
    from pydmrg_b200 import mpo as M
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 1500) + int(continuation)
    ps = [ctx.Process(target=_scan_worker, args=(r, 2, port, q, continuation)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict((r[0], r[1:]) for r in (q.get(timeout=600) for _ in ps))
    for p in ps:
        p.join(timeout=60)
    s_vals = np.linspace(-0.5, 0.5, 5)
    exact = []
    for s in s_vals:
        op = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))[0]
        T = op[0][0]
        for W in op[1:]:
            T = np.einsum("a...,abij->b...ij", T, W)
        T = T[0]
        idx = list(range(0, 12, 2)) + list(range(1, 12, 2))
        ev = np.linalg.eigvals(T.transpose(idx).reshape(64, 64))
        exact.append(ev[np.argmax(ev.real)])
    exact = np.array(exact)
    assert sorted(res[0][2] + res[1][2]) == [0, 1, 2, 3, 4]
    assert res[0][2] == ([0, 1, 2] if continuation else [0, 2, 4])
    for rank in (0, 1):
        assert np.array_equal(res[rank][0], res[0][0])
        assert np.max(np.abs(res[rank][0] - exact)) < 1e-9          # the eigenvalue at s = 0 is exactly 0: absolute


def test_larger_davidson_subspace_needs_fewer_matvecs(monkeypatch, golden):
    """``max_space`` (solver option of run_dmrg / calc_eigs): same eigenpair, fewer restarts."""
    from emu_device import install
    from oracle import dmrg_oracle as orc
    install(monkeypatch.setattr)
    from pydmrg_b200 import davidson as dav
    g = golden["davidson"]
    L, R, W, x0 = g["L"], g["R"], g["W"], g["x0"]

    def matvec(x, y):
        y.copy_(torch.from_numpy(orc.heff_onesite(x.numpy(), [L], [W], [R], x0.shape)))

    out = {}
    for ms in (12, 24):
        dav.release_workspaces()
        st = {}
        e, vecs = dav.davidson(matvec, [torch.from_numpy(x0.ravel().copy())], x0.size, 1, torch.device("cpu"), nroots=1, stats=st,
                               native=False, max_space=ms)
        out[ms] = (e[0], st["n_matvec"])
    dav.release_workspaces()
    assert abs(out[24][0] - out[12][0]) < 1e-7 * abs(out[12][0])
    assert out[24][1] < out[12][1]


def test_diagonal_preconditioner_same_eigenpair_fewer_matvecs(monkeypatch, golden):
    """davidson(diag=...): t = r / (diag(A) - theta) in the correction step.  On the reference's seeded problem the
    eigenpair is the same (the fixed point does not depend on the preconditioner) and the mat-vec count drops; the
    emulated diagonal equals the diagonal of the dense operator."""
    from emu_device import install, HeffPlan, Context
    from oracle import dmrg_oracle as orc
    D = install(monkeypatch.setattr)
    from pydmrg_b200 import davidson as dav
    g = golden["davidson"]
    L, R, W, x0 = g["L"], g["R"], g["W"], g["x0"]
    ctx = Context("c128")
    d, Dl, Dr = x0.shape
    plan = HeffPlan(1, d, Dl, Dr, [(ctx.dev(L), D.combine_onesite(W, L.shape[1], R.shape[1], d, np.complex128), ctx.dev(R))], None)
    diag = plan.diag(-1.0)
    dense = -orc.heff_dense_onesite([L], [W], [R], x0.shape)           # matvec = -H
    assert np.allclose(diag.numpy(), np.diag(dense), atol=1e-13 * np.abs(dense).max())
    out = {}
    for use in (False, True):
        dav.release_workspaces()
        st = {}
        e, vecs = dav.davidson(lambda x, y: plan.apply(x, y, -1.0), [ctx.dev(x0.ravel())], x0.size, 1, torch.device("cpu"), nroots=1,
                               stats=st, native=False, res_tol=1e-10, diag=diag if use else None)
        out[use] = (e[0], st["n_matvec"], vecs[0].numpy())
    dav.release_workspaces()
    assert abs(out[True][0] - out[False][0]) < 1e-9 * abs(out[False][0])
    assert abs(out[True][0] - g["exact_min_real"]) < 1e-9 * abs(g["exact_min_real"])
    assert abs(abs(np.vdot(out[True][2], out[False][2])) - 1) < 1e-8
    assert out[True][1] < out[False][1]


def test_sharded_sweep_check_script_dry_run_on_two_gloo_ranks():
    """tests/mgpu_sharded_sweep_check.py (the multi-GPU check of sharded sweeps) executed as it is on two gloo ranks
    over the emulated device entry points (PDMRG_CHECK_EMULATE=1): ASEP c128 and Heisenberg f64, ket and bond mode,
    real ShardedHeff / ShardSpec / env_update code paths -- the script must report ok."""
    import json
    import subprocess
    port = 36500 + (os.getpid() % 1500)
    env = dict(os.environ, PDMRG_CHECK_EMULATE="1", PDMRG_CHECK_CHI="32", PDMRG_CHECK_N="12", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(HERE, "mgpu_sharded_sweep_check.py")],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["world"] == 2
    for key in ("asep_ket", "asep_bond", "heisenberg_ket", "heisenberg_bond"):
        assert rep[key]["sharded_solves"] > 0 and rep[key]["sharded_block_updates"] > 0 and rep[key]["same_on_all_ranks"]
        assert rep[key]["dE"] < 1e-9


def test_python_loop_thick_restart_stagnation_and_relative_tolerance(monkeypatch, golden):
    """The Python Davidson loop (the readable specification of csrc/davidson.cu) with the round-2 solver options over the
    emulated vector kernels: thick restarts reach the reference's eigenpair with fewer mat-vecs, ARPACK's relative rule
    stops earlier, the stagnation stop ends a solve that makes no progress."""
    from emu_device import install
    from oracle import dmrg_oracle as orc
    install(monkeypatch.setattr)
    from pydmrg_b200 import davidson as dav
    g = golden["davidson"]
    L, R, W, x0 = g["L"], g["R"], g["W"], g["x0"]

    def matvec(x, y):
        y.copy_(torch.from_numpy(orc.heff_onesite(x.numpy(), [L], [W], [R], x0.shape)))

    def run(**kw):
        dav.release_workspaces()
        st = {}
        e, vecs = dav.davidson(matvec, [torch.from_numpy(x0.ravel().copy())], x0.size, 1, torch.device("cpu"), nroots=1, stats=st,
                               native=False, **kw)
        return e[0], st, vecs[0].numpy()

    e0, st0, _ = run(res_tol=1e-9, max_cycle=3000)
    e1, st1, v1 = run(res_tol=1e-9, max_cycle=3000, restart_keep=4)
    assert abs(e1 - e0) < 1e-8 * abs(e0) and st1["n_thick_restarts"] > 0 and st1["n_matvec"] < st0["n_matvec"]
    A = -orc.heff_dense_onesite([L], [W], [R], x0.shape)
    assert np.linalg.norm(A @ v1 - e1 * v1) < 1e-8 * np.linalg.norm(v1)
    e2, st2, _ = run(res_tol=1e-5, tol_relative=True, restart_keep=4)
    assert st2["n_matvec"] < st1["n_matvec"] and abs(e2 - e0) < 1e-4 * abs(e0)
    # the stagnation stop ends a solve early and reports the residual it stopped at; on this NON-NORMAL operator the residual
    # of the restarted iteration is not monotone and has long plateaus, so the stop fires far from convergence -- the
    # reason why the scan driver and the bench do not use it
    e3, st3, v3 = run(res_tol=1e-15, max_cycle=3000, stagnation=200)
    assert st3["cycles"] < 3000 and np.isfinite(e3)
    r3 = np.linalg.norm(A @ v3 - e3 * v3) / np.linalg.norm(v3)
    assert abs(r3 - st3["last_residual"]) < 1e-6 * max(1.0, r3) and r3 > 1e-9
    dav.release_workspaces()


def test_scan_self_check_and_per_point_seeds():
    """scan.scgf_shape_ok flags a point that sits off the convex curve; run_scan's per-point seeds accept a sequence."""
    from pydmrg_b200.scan import scgf_shape_ok
    s = np.linspace(-0.5, 0.5, 21)
    E = 2.0 * np.exp(-3.0 * s) - 0.26                                       # convex, decreasing (the shape of configs[3])
    assert scgf_shape_ok(s, E) and scgf_shape_ok(s[::-1], E[::-1])
    bad = E.copy(); bad[7] -= 0.3                                         # a 10 % dip, the size seen in the scans
    assert not scgf_shape_ok(s, bad)
    worse = E.copy(); worse[12] += 0.2                                      # not even monotone
    assert not scgf_shape_ok(s, worse)
