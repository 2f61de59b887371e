"""CPU: host-side logic of the N > 1 paths -- scan-point blocks, MPO-bond block partition --
including a world_size-2 gloo run (no GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scan_blocks_cover_everything():
    from pydmrg_b200.parallel import scan_block
    for n, world in ((64, 8), (50, 8), (7, 3), (3, 8)):
        seen = []
        for r in range(world):
            lo, hi = scan_block(n, r, world)
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_scan_indices_interleaved_and_gather():
    """Interleaved assignment covers every point once, is balanced to within one point, and gather_scan puts
    each rank's values back at its own indices."""
    from pydmrg_b200.parallel import gather_scan, scan_indices
    for n, world in ((64, 8), (50, 8), (7, 3), (3, 8)):
        owned = [scan_indices(n, r, world, interleave=True) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        assert [scan_indices(n, r, world) for r in range(world)][0][0] == 0
    mine = scan_indices(7, 1, 3, interleave=True)
    full = gather_scan([float(i) for i in mine], 7, 1, 3, None, interleave=True)     # no dist: only own slots filled
    assert [full[i].real for i in mine] == [float(i) for i in mine] and full[0] == 0


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_block_partition_is_a_partition(world):
    from pydmrg_b200 import mpo as M
    from pydmrg_b200.device import combine_twosite
    from pydmrg_b200.parallel import block_pattern, partition_blocks, partition_cost
    W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
    Wc = combine_twosite(W1, W2, np.float64)
    pat = block_pattern(Wc, 6, 6, 4)
    assert pat.sum() == 11                      # (j,0) for all j and (5,o) for all o
    masks = partition_blocks(pat, world)
    tot = np.zeros_like(pat, dtype=int)
    for m in masks:
        tot += m
    assert np.array_equal(tot, pat.astype(int))  # every non-zero block owned exactly once
    per, ideal = partition_cost(masks)
    assert max(per) <= 12 and ideal >= 1.0
    assert max(per) <= {1: 12, 2: 7, 4: 4, 8: 3}[world]      # the optimum for the ASEP star pattern
    # linearity: sum over ranks of the masked operators is the operator
    Wr = Wc.reshape(6, 4, 6, 4)
    acc = np.zeros_like(Wr)
    for m in masks:
        acc += Wr * m[:, None, :, None]
    assert np.array_equal(acc, Wr)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pydmrg_b200.parallel import gather_scan, scan_block
    n = 11
    lo, hi = scan_block(n, rank, world)
    local = [complex(i, -i) for i in range(lo, hi)]
    out = gather_scan(local, n, rank, world, dist)
    q.put((rank, out))
    dist.destroy_process_group()


def test_gather_scan_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    expect = np.array([complex(i, -i) for i in range(11)])
    for rank, out in res:
        assert np.array_equal(out, expect)


def _queue_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import time
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pydmrg_b200.parallel import WorkQueue, gather_scan_dynamic
    order = [5, 0, 3, 1, 4, 2, 6]
    got = []
    for idx in WorkQueue(order, rank, world, dist):
        got.append(idx)
        time.sleep(0.3 if (rank == 0 and len(got) == 1) else 0.01)      # rank 0's first point is an expensive one
    out = gather_scan_dynamic(got, [complex(i, rank) for i in got], len(order), world, dist)
    # a second queue in the same process group must start from a fresh counter
    again = list(WorkQueue([0, 1], rank, world, dist))
    dist.barrier()
    q.put((rank, got, out, again))
    dist.destroy_process_group()


def test_dynamic_work_queue_gloo_world2():
    """parallel.WorkQueue: every point is taken exactly once, in the given order, and the rank that sits on an
    expensive point takes fewer of them; results are assembled on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_queue_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = {r: (got, out, again) for r, got, out, again in (q.get(timeout=120) for _ in ps)}
    for p in ps:
        p.join(timeout=60)
    assert sorted(res[0][0] + res[1][0]) == list(range(7))
    assert len(res[0][0]) < len(res[1][0])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(np.real(res[0][1]), np.arange(7))
    assert sorted(res[0][2] + res[1][2]) == [0, 1]


def test_work_queue_single_process():
    from pydmrg_b200.parallel import WorkQueue
    assert list(WorkQueue([2, 0, 1], 0, 1)) == [2, 0, 1]


# ---------------------------------------------------------------------------------------
# right / left eigenstates on different ranks
# ---------------------------------------------------------------------------------------
def _oracle_runner(mpo, mbd=8, tol=1e-10, maxIter=4, returnEntSpec=False, returnState=False, returnEnv=False, **kw):
    """Stand-in for dmrg.run_dmrg with its return packing, computed by the NumPy oracle (CPU)."""
    from oracle import dmrg_oracle as orc
    np.random.seed(kw.get("seed", 5))
    E, EE, EEs, mps, env = orc.run_dmrg_onesite(mpo, int(np.atleast_1d(mbd)[-1]), tol=tol, max_iter=maxIter)
    out = [E, EE, None]
    if returnEntSpec:
        out.append(EEs)
    if returnState:
        out.append([mps])
    if returnEnv:
        out.append(env)
    return out


def _lr_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pydmrg_b200 import mpo as M
    from pydmrg_b200.parallel import run_dmrg_left_right
    mpo = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    out = run_dmrg_left_right(mpo, rank, world, dist, runner=_oracle_runner, mbd=8, returnEntSpec=True, returnState=True)
    q.put((rank, out))
    dist.destroy_process_group()


def test_left_right_eigenstates_on_two_ranks_gloo():
    """Rank 0 solves for the right eigenstate, rank 1 for the left one (on the conjugate-transposed MPO); every
    rank ends with the reference's calcLeftState packing; <l|H|r>/<l|r> reproduces the eigenvalue."""
    import torch.multiprocessing as mp
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_lr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    mpo = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    right = _oracle_runner(mpo, mbd=8, returnEntSpec=True, returnState=True)
    left = _oracle_runner(M.mpo_conj_trans(mpo), mbd=8, returnEntSpec=True, returnState=True)
    for rank in (0, 1):
        E, EEpair, gap, EEsPair, mpsPair = res[rank]
        assert E == right[0] and gap is None
        assert EEpair == [right[1], left[1]]
        assert np.array_equal(EEsPair[0], right[3]) and np.array_equal(EEsPair[1], left[3])
        for a, b in zip(mpsPair[0][0], right[4][0]):
            assert np.array_equal(a, b)
        for a, b in zip(mpsPair[1][0], left[4][0]):
            assert np.array_equal(a, b)
    # physics: the left state is the left eigenvector of the same eigenvalue
    E, _, _, _, (mr, ml) = res[0]
    lmps = [np.conj(t) for t in ml[0]]
    num = orc.full_contract(mpo, mr[0], lmps)
    den = orc.full_contract(None, mr[0], lmps)
    assert abs(num / den - E) < 1e-7 * abs(E)
    assert abs(left[0] - np.conj(right[0])) < 1e-7 * abs(E)


def test_left_rank_uses_the_reference_file_names(tmp_path):
    """The left rank of run_dmrg_left_right reads / writes '<name>_mbd{i}_left' (dmrg.py:423,436,476): the names of the
    reference and of run_dmrg(calcLeftState=True), not '<name>_left_mbd{i}'."""
    from pydmrg_b200.parallel import run_dmrg_left_right
    seen = {}

    class FakeDist:
        @staticmethod
        def all_gather_object(out, obj):
            out[0] = out[1] = obj

    def runner(mpo, **kw):
        seen.update(kw)
        return [1.0, 0.0, None]

    run_dmrg_left_right([[None]], 1, 2, FakeDist, runner=runner, fname=str(tmp_path / "st"), initGuess=str(tmp_path / "g"), mbd=4)
    assert seen["fname"] == str(tmp_path / "st") and seen["initGuess"] == str(tmp_path / "g")
    assert seen["name_suffix"] == "_left"


def test_left_right_pair_detection():
    from pydmrg_b200.parallel import _is_pair
    t = np.zeros((2, 1, 2))
    mpsList = [[t, t, t]]
    env2 = [[t, t], [t, t]]                      # two MPO terms: a list of length two that is NOT a pair
    assert _is_pair((mpsList, mpsList)) and _is_pair([env2, env2])
    assert not _is_pair(mpsList) and not _is_pair(env2) and not _is_pair((mpsList, 3)) and not _is_pair(None)


# ---------------------------------------------------------------------------------------
# sharded sweeps: wiring of calc_eigs(shard=...) (the kernels are stubbed; no GPU)
# ---------------------------------------------------------------------------------------
def test_calc_eigs_routes_large_bonds_through_the_sharded_matvec(monkeypatch):
    import torch
    from pydmrg_b200 import eigs, parallel

    class FakePlan:
        def __init__(self, n):
            self.n, self.closed = n, False

        def apply(self, x, y, sign):
            y.copy_(x * sign)

        def close(self):
            self.closed = True

    class FakeSharded:
        def __init__(self, n):
            self.plan, self.calls = FakePlan(n), 0

        def apply(self, x, y, sign):
            self.calls += 1
            y.copy_(2 * x * sign)
            return y

    class Spec(parallel.ShardSpec):
        def __init__(self, min_dim):
            self.rank, self.world, self.dist, self.mode, self.min_dim = 0, 2, None, "ket", min_dim
            self.made, self.agreed = [], 0

        def heff(self, ctx, M, W, F, site, oneSite):
            sh = FakeSharded(M[site].numel() * M[site + 1].shape[0] * M[site + 1].shape[2] // M[site].shape[2])
            self.made.append(sh)
            return sh

        def consensus(self, t):
            self.agreed += 1
            return t

        def consensus_host(self, a, ctx):
            self.host_agreed = getattr(self, "host_agreed", 0) + 1
            return a

    class Ctx:
        code, device, tdtype, npdtype = 1, torch.device("cpu"), torch.complex128, np.complex128

        def dev(self, a):
            return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.complex128))

    seen = {}

    def fake_davidson(matvec, x0, n, code, device, nroots=1, plan=None, native=None, consensus=None, **kw):
        y = torch.empty_like(x0[0])
        matvec(x0[0], y)
        seen.update(plan=plan, native=native, consensus=consensus)
        if consensus is not None:
            consensus(y)
        return np.array([-1.0 + 0j]), [x0[0] / x0[0].norm()]

    def fake_make_plan(M, W, F, site, oneSite, ctx, block_masks=None):
        d1, Dl, _ = M[site].shape
        d2, _, Dr = M[site + 1].shape
        return FakePlan(d1 * d2 * Dl * Dr), (d1, d2, Dl, Dr)

    monkeypatch.setattr(eigs, "davidson", fake_davidson)
    monkeypatch.setattr(eigs, "make_plan", fake_make_plan)
    monkeypatch.setattr(eigs, "two_site_guess", lambda A, B, ctx: torch.einsum("ijk,lkm->iljm", A, B).contiguous())
    monkeypatch.setattr(eigs, "check_overlap", lambda guess, vecs, E, preserveState, ctx: (np.atleast_1d(E), vecs, 1.0))
    rng = np.random.default_rng(0)
    mk = lambda *s: torch.from_numpy(rng.standard_normal(s) + 0j)
    mps = [[mk(2, 4, 8), mk(2, 8, 8), mk(2, 8, 8), mk(2, 8, 4)]]
    ctx = Ctx()
    spec = Spec(min_dim=8)
    E, V, _ = eigs.calc_eigs(mps, [[None] * 4], [[None] * 5], 1, 1, oneSite=False, ctx=ctx, shard=spec, return_device=True)
    assert len(spec.made) == 1 and spec.made[0].calls == 1 and spec.made[0].plan.closed     # Dl = Dr = 8: sharded
    assert seen["plan"] is None and seen["native"] is False and seen["consensus"] is not None
    assert spec.agreed == 3 and spec.host_agreed == 1          # start vector, control scalars, eigenvector; eigenvalue
    assert E == 1.0 and tuple(V.shape) == (2 * 2 * 8 * 8, 1)
    E, V, _ = eigs.calc_eigs(mps, [[None] * 4], [[None] * 5], 0, 1, oneSite=False, ctx=ctx, shard=spec, return_device=True)
    assert len(spec.made) == 1 and seen["consensus"] is None and seen["plan"] is not None   # Dl = 4 < min_dim: local plan
    assert spec.agreed == 4 and spec.host_agreed == 2          # ... whose eigenpair is still taken from rank 0


def test_shard_spec_keeps_every_rank_a_ket_range():
    from pydmrg_b200.parallel import ShardSpec, ket_chunks
    s = ShardSpec(3, 8, None, min_dim=64)
    assert s.min_dim == 128 and all(n > 0 for _, n in ket_chunks(s.min_dim, 8))
