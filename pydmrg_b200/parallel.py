"""Multi-GPU decomposition of the hot path (one process per GPU, torch.distributed / NCCL).

Two paths shard:

* parameter / s-field scans: independent DMRG runs, one contiguous block of scan points per
  rank, no data-path collective (results gathered at the end) -- ``scan_block`` /
  ``gather_scan``;
* the H_eff mat-vec at large chi, split over the MPO bond: the non-zero (j, o) blocks of the
  local operator are partitioned over ranks; each rank computes its partial y with only the
  L[:, j, :] / R[:, o, :] slabs it touches, followed by ONE all-reduce(sum) of y over
  NVLink per mat-vec -- ``partition_blocks`` / ``ShardedHeff``.
"""
import numpy as np


def scan_block(n_points, rank, world):
    """Contiguous block [lo, hi) of scan points owned by ``rank`` (warm-start continuation is
    kept inside a block)."""
    base, rem = divmod(n_points, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scan_indices(n_points, rank, world, interleave=False):
    """Scan points owned by ``rank``: the contiguous block of scan_block, or -- for cold-started scans, whose
    cost per point varies smoothly (3x) along the scan -- every ``world``-th point, which balances the ranks."""
    if interleave:
        return list(range(rank, n_points, world))
    lo, hi = scan_block(n_points, rank, world)
    return list(range(lo, hi))


def gather_scan(local_values, n_points, rank, world, dist=None, interleave=False):
    """All ranks receive the full array of scan results (complex128)."""
    out = np.zeros(n_points, dtype=np.complex128)
    out[scan_indices(n_points, rank, world, interleave)] = np.asarray(local_values, dtype=np.complex128)
    if world == 1 or dist is None:
        return out
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.view_as_real(torch.from_numpy(out)).to(dev).contiguous()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return torch.view_as_complex(t.cpu()).numpy()


class WorkQueue:
    """Dynamic distribution of independent scan points: ``order`` (the same list on every rank, most expensive
    first if known) is handed out one point at a time -- a rank takes the next point when it is free.  The only shared
    state is ONE atomic counter in torch.distributed's key-value store (``Store.add``, served by rank 0's TCPStore: a
    few bytes per point, nothing on the data path).  The cost of a scan point varies by more than 10x (the gap of the
    tilted generator closes at s = 0), which static blocks or strides cannot balance."""

    _calls = 0

    def __init__(self, order, rank, world, dist=None, store=None):
        self.order, self.rank, self.world = list(order), rank, world
        WorkQueue._calls += 1                                    # every rank builds its queues in the same order
        self.key = "pdmrg_workqueue_%d" % WorkQueue._calls
        self.local = 0
        self.store = None
        import threading
        self._lock = threading.Lock()
        if world > 1:
            self.store = store if store is not None else dist.distributed_c10d._get_default_store()

    def __iter__(self):
        return self

    def __next__(self):
        with self._lock:                                         # several worker threads of one rank share the queue
            if self.store is None:
                pos = self.local
                self.local += 1
            else:
                pos = int(self.store.add(self.key, 1)) - 1
        if pos >= len(self.order):
            raise StopIteration
        return self.order[pos]


def gather_scan_dynamic(indices, local_values, n_points, world, dist=None):
    """Full result array on every rank from (index, value) pairs held by the ranks (one object gather at the end)."""
    out = np.zeros(n_points, dtype=np.complex128)
    pairs = [(int(i), complex(v)) for i, v in zip(indices, local_values)]
    if world > 1 and dist is not None:
        everyone = [None] * world
        dist.all_gather_object(everyone, pairs)
        pairs = [p for sub in everyone for p in sub]
    for i, v in pairs:
        out[i] = v
    return out


def run_dmrg_left_right(mpo, rank, world, dist, runner=None, **kwargs):
    """``run_dmrg(..., calcLeftState=True)`` with the right and the left eigenstate computed on different
    ranks.  The reference runs them one after the other as two independent DMRGs -- the
    second on ``mpo_conj_trans(mpo)``, started from a copy of the same initial state (dmrg.py:338, 369-373,
    405, 451) -- so they shard with no data-path collective: even ranks take the right state, odd ranks the
    left one, and one object gather at the end assembles the reference's return packing
    ``[E, [EE, EEl], gap, ([EEs, EEsl]), ([mps, mpsl]), ([env, envl])]`` (dmrg.py:491-508; E and gap are those
    of the right run, as in the reference) on EVERY rank.

    Both ranks must draw the same initial state: seed NumPy identically before the call (the reference draws
    one more, discarded, random MPS before copying; it has no influence on either run).  ``initGuess`` /
    ``initEnv`` given as a ``(right, left)`` pair are split between the sides.  File-based continuation
    (``fname`` / a string ``initGuess``) keeps the reference's naming: the left side uses ``<name>_left``.
    ``runner`` defaults to ``pydmrg_b200.dmrg.run_dmrg`` (injectable for CPU tests of the distribution logic)."""
    if runner is None:
        from .dmrg import run_dmrg as runner
    from .mpo import mpo_conj_trans
    if world < 2:
        return runner(mpo, calcLeftState=True, **kwargs)
    kw = dict(kwargs)
    kw.pop("calcLeftState", None)
    flags = [bool(kw.get(k, False)) for k in ("returnEntSpec", "returnState", "returnEnv")]
    side = rank % 2                                              # 0: right eigenstate, 1: left eigenstate
    for key in ("initGuess", "initEnv"):
        v = kw.get(key)
        if isinstance(v, tuple) and len(v) == 2 and isinstance(v[1], (int, np.integer)) and _is_pair(v[0]):
            kw[key] = (v[0][side], v[1])                         # ((right, left), gaugeSite)
        elif _is_pair(v):
            kw[key] = v[side]
    if side == 1:
        # the left rank reads / writes '<name>_mbd{i}_left', the reference's names (dmrg.py:423,436,476) and those of
        # run_dmrg(calcLeftState=True), so checkpoints are interchangeable between the drivers
        kw["name_suffix"] = kw.get("name_suffix", "") + "_left"
    out = None
    if rank < 2:                                                 # ranks beyond the first pair have nothing to do
        out = runner(mpo if side == 0 else mpo_conj_trans(mpo), **kw)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    right, left = gathered[0], gathered[1]
    res = [right[0], [right[1], left[1]], right[2]]
    pos = 3
    for on in flags:
        if on:
            res.append([right[pos], left[pos]])
            pos += 1
    return res


def _is_pair(v):
    """(right, left) pair of MPS lists [state][site] / environments [mpoInd][k], as opposed to a single one."""
    try:
        return (isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(side, (list, tuple)) for side in v)
                and all(isinstance(side[0], (list, tuple)) and np.ndim(side[0][0]) == 3 for side in v))
    except Exception:
        return False


def block_pattern(Wc, wL, wR, P):
    """Boolean (wL, wR) matrix: which (j, o) blocks of the combined local operator are non-zero."""
    B = np.abs(np.asarray(Wc).reshape(wL, P, wR, P)).max(axis=(1, 3)) > 0
    return B


def partition_blocks(pattern, world):
    """Balanced partition of the non-zero (j, o) blocks over ``world`` ranks.

    The cost of a rank is the number of chi^3 slab GEMMs it must run: |{o}| stage-A slabs plus
    |{j}| stage-C slabs of the blocks it owns.  Greedy seeding followed by a local search (single
    block moves and pair swaps that lower the (max, sum of squares) cost).  Returns a list of
    ``world`` uint8 masks of shape (wL, wR) that partition ``pattern``."""
    wL, wR = pattern.shape
    blocks = [(j, o) for j in range(wL) for o in range(wR) if pattern[j, o]]
    rowc, colc = pattern.sum(axis=1), pattern.sum(axis=0)
    blocks.sort(key=lambda b: -(rowc[b[0]] + colc[b[1]]))
    owner = {}

    def cost_of(bl):
        return len({b[0] for b in bl}) + len({b[1] for b in bl}) if bl else 0

    def groups():
        g = [[] for _ in range(world)]
        for b, r in owner.items():
            g[r].append(b)
        return g

    def score(g):
        c = [cost_of(x) for x in g]
        return (max(c), sum(v * v for v in c))

    for b in blocks:                                   # greedy seed
        g = groups()
        best = min(range(world), key=lambda r: (cost_of(g[r] + [b]), cost_of(g[r] + [b]) - cost_of(g[r]), r))
        owner[b] = best
    improved = True
    while improved:                                    # local search
        improved = False
        cur = score(groups())
        for b in blocks:
            r0 = owner[b]
            for r in range(world):
                if r == r0:
                    continue
                owner[b] = r
                sc = score(groups())
                if sc < cur:
                    cur, improved, r0 = sc, True, r
                else:
                    owner[b] = r0
        for a in blocks:
            for b in blocks:
                if owner[a] == owner[b]:
                    continue
                owner[a], owner[b] = owner[b], owner[a]
                sc = score(groups())
                if sc < cur:
                    cur, improved = sc, True
                else:
                    owner[a], owner[b] = owner[b], owner[a]
    # small patterns (the MPOs of the reference have <= ~20 blocks): exact branch and bound on the
    # maximum cost, seeded with the local-search result, ranks filled in canonical order
    best = {"score": score(groups()), "owner": dict(owner)}
    budget = [200000]
    rows = [set() for _ in range(world)]
    cols = [set() for _ in range(world)]
    assign = {}

    def rec(i, used):
        if budget[0] <= 0:
            return
        budget[0] -= 1
        cur_max = max((len(rows[r]) + len(cols[r])) for r in range(world))
        if cur_max > best["score"][0]:
            return
        if i == len(blocks):
            sc = (cur_max, sum((len(rows[r]) + len(cols[r])) ** 2 for r in range(world)))
            if sc < best["score"]:
                best["score"], best["owner"] = sc, dict(assign)
            return
        j, o = blocks[i]
        for r in range(min(used + 1, world)):
            addj, addo = j not in rows[r], o not in cols[r]
            if addj:
                rows[r].add(j)
            if addo:
                cols[r].add(o)
            assign[(j, o)] = r
            rec(i + 1, max(used, r + 1))
            if addj:
                rows[r].discard(j)
            if addo:
                cols[r].discard(o)
        assign.pop((j, o), None)

    if len(blocks) <= 24:
        rec(0, 0)
    owner = best["owner"]
    masks = []
    for r in range(world):
        m = np.zeros((wL, wR), dtype=np.uint8)
        for (j, o), rr in owner.items():
            if rr == r:
                m[j, o] = 1
        masks.append(m)
    return masks


def partition_cost(masks):
    """Slab-GEMM count per rank and the ideal speed-up over the unsharded count."""
    per = [int(m.any(axis=1).sum() + m.any(axis=0).sum()) for m in masks]
    tot = np.zeros_like(masks[0])
    for m in masks:
        tot |= m
    full = int(tot.any(axis=1).sum() + tot.any(axis=0).sum())
    return per, full / max(max(per), 1)


def ket_chunks(Dl, parts, align=16):
    """Split [0, Dl) into ``parts`` contiguous ranges whose boundaries are multiples of ``align``
    (keeps every operand of the shard TMA-aligned); trailing ranks may get an empty range."""
    per = -(-Dl // parts)
    per = -(-per // align) * align
    out = []
    for r in range(parts):
        lo = min(Dl, r * per)
        hi = min(Dl, lo + per)
        out.append((lo, hi - lo))
    return out


class ShardedHeff:
    """Sharded mat-vec with one all-reduce: y = allreduce_sum( H_r x ).

    mode='bond' : H_r = this rank's (j, o) blocks of the MPO bond (the environment slabs a rank touches
                  shrink with the rank count -- the memory-sizing mode of SURVEY 8e; the speed-up is
                  limited by the block structure, e.g. 4x for the ASEP star pattern);
    mode='ket'  : H_r = all blocks restricted to a contiguous range of the left ket index (every rank does
                  1/n of both GEMM stages: balanced for any rank count, environments replicated);
    mode='auto' : 'ket'."""

    def __init__(self, code, P, Dl, Dr, ops, workspace, rank, world, dist, mode="bond"):
        from . import device as D
        self.dist, self.world = dist, world
        self.mode = "ket" if mode == "auto" else mode
        self.partition = []
        if self.mode == "bond":
            masks = []
            for (L, Wc, R) in ops:
                pat = block_pattern(Wc, L.shape[1], R.shape[1], P)
                parts = partition_blocks(pat, world)
                self.partition.append(parts)
                masks.append(parts[rank])
            self.plan = D.HeffPlan(code, P, Dl, Dr, ops, workspace, block_masks=masks)
        elif self.mode == "ket":
            self.k_range = ket_chunks(Dl, world)[rank]
            self.plan = D.HeffPlan(code, P, Dl, Dr, ops, workspace, k_range=self.k_range)
        else:
            raise ValueError("mode must be 'bond', 'ket' or 'auto'")
        self.n = self.plan.n
        self.local_flops = self.plan.flops_sparse

    def apply(self, x, y, sign=-1.0):
        import torch
        self.plan.apply(x, y, sign)
        if self.world > 1:
            self.dist.all_reduce(torch.view_as_real(y) if y.is_complex() else y, op=self.dist.ReduceOp.SUM)
        return y


class ShardSpec:
    """How a sweep shards its mat-vecs (passed as ``run_dmrg(..., shard=ShardSpec(rank, world, dist))``):
    every rank runs the whole sweep on replicated tensors; the local eigenproblems of bonds with
    min(Dl, Dr) >= ``min_dim`` use a ShardedHeff in ``mode`` ('ket': balanced for any MPO; 'bond': blocks of
    the MPO bond) -- one NCCL all-reduce of y per mat-vec; with ``shard_env`` the block updates of those bonds
    are split over the ket index of the new block (one all-reduce of the block).  The truncation stays replicated."""

    def __init__(self, rank, world, dist, mode="ket", min_dim=256, shard_env=True, fused=True):
        self.rank, self.world, self.dist, self.mode = int(rank), int(world), dist, mode
        self.min_dim = max(int(min_dim), 16 * self.world)       # every rank must own a TMA-aligned ket range
        self.shard_env = bool(shard_env)
        # fused (ket mode, one MPO term): the local solve runs in the C++ Davidson loop on every rank, each mat-vec with
        # the exchange fused into the kernels over peer memory (pdmrg_davidson_sharded) -- no NCCL call and no control
        # broadcast inside a solve; otherwise the Python loop with one NCCL all-reduce per mat-vec
        self.fused = bool(fused) and mode == "ket"
        self._peer = None

    def fused_heff(self, ctx, M, W, F, site, oneSite):
        """This rank's ket-split shard of the local operator bound to the sweep's cached peer communicator, or None when
        the fused path does not apply (several MPO terms)."""
        from .eigs import _site_ops
        if not getattr(self, "fused", False) or len(W) != 1 or getattr(getattr(ctx, "device", None), "type", "cuda") != "cuda":
            return None                                          # several MPO terms, or the CPU dry run of the host logic
        ops, P, shape = _site_ops(M, W, F, site, oneSite, ctx)
        fz = FusedShardedHeff(ctx.code, P, shape[-2], shape[-1], ops, ctx.ws, self.rank, self.world, self.dist, mode="ket",
                              comm=self._peer)
        if fz._own_comm:                                         # first large bond, or a larger one: keep the new buffers
            if self._peer is not None:
                self._peer.close()
            self._peer = fz.peer
            fz._own_comm = False
        return fz

    def heff(self, ctx, M, W, F, site, oneSite):
        from .eigs import _site_ops
        ops, P, shape = _site_ops(M, W, F, site, oneSite, ctx)
        return ShardedHeff(ctx.code, P, shape[-2], shape[-1], ops, ctx.ws, self.rank, self.world, self.dist, mode=self.mode)

    def env_update(self, direction, T, W, blk, ctx):
        """Block update split over the ket index of the NEW block: every rank computes the columns
        it owns (1/world of both GEMM stages, csrc/env_cols.cu) into a zero-initialised full-size block; one
        all-reduce assembles it on every rank."""
        import torch
        from . import device as D
        n = T.shape[2] if direction == "right" else T.shape[1]
        k0, kn = ket_chunks(n, self.world)[self.rank]
        out = D.env_update_cols(direction, T, W, blk, None, k0, kn, ctx.ws)
        self.dist.all_reduce(torch.view_as_real(out) if out.is_complex() else out, op=self.dist.ReduceOp.SUM)
        return out

    def consensus_host(self, a, ctx):
        """Rank 0's copy of a small host array (eigenvalues) on every rank."""
        import torch
        if self.world <= 1:
            return a
        a = np.ascontiguousarray(a, dtype=np.complex128)
        t = torch.view_as_real(torch.from_numpy(a.reshape(-1).copy())).to(ctx.device)
        self.dist.broadcast(t, src=0)
        return torch.view_as_complex(t.cpu().contiguous()).numpy().reshape(a.shape)

    def wants_env(self, *dims):
        return self.world > 1 and self.shard_env and min(dims) >= self.min_dim

    def consensus(self, t):
        """Overwrite a small device tensor by rank 0's copy (control scalars, eigenvectors)."""
        import torch
        if self.world > 1:
            self.dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, src=0)
        return t


class PeerComm:
    """Peer-memory communicator of the fused sharded mat-vec (csrc/comm.cu): three buffers per rank -- the slab buffer Z the
    peers' GEMM epilogues store into, the staging buffer of the reduced y, the flags -- exchanged ONCE through CUDA IPC.
    Sized in bytes, so one communicator serves every bond of a sweep whose buffers fit (parallel.ShardSpec caches it)."""

    def __init__(self, z_bytes, y_bytes, rank, world, dist=None):
        import ctypes
        import torch
        from ._lib import check, lib
        self.rank, self.world = rank, world
        self.z_bytes, self.y_bytes = int(z_bytes), int(y_bytes)
        self._lib = lib
        sizes = [self.z_bytes, self.y_bytes, (2 * world + 1) * 8 + 256]
        self._own, handles = [], []
        for nb in sizes:
            ptr = ctypes.c_void_p(0)
            hbuf = ctypes.create_string_buffer(64)
            check(lib.pdmrg_ipc_alloc(nb, ctypes.byref(ptr), hbuf), "pdmrg_ipc_alloc")
            self._own.append(ptr)
            handles.append(hbuf.raw)
        if world > 1:
            everyone = [None] * world
            dist.all_gather_object(everyone, handles)
        else:
            everyone = [handles]
        self._opened = []
        tables = []
        for k in range(3):
            arr = (ctypes.c_void_p * world)()
            for q in range(world):
                if q == rank:
                    arr[q] = self._own[k].value
                else:
                    ptr = ctypes.c_void_p(0)
                    check(lib.pdmrg_ipc_open(everyone[q][k], ctypes.byref(ptr)), "pdmrg_ipc_open")
                    self._opened.append(ptr)
                    arr[q] = ptr.value
            tables.append(arr)
        self.handle = ctypes.c_void_p(0)
        check(lib.pdmrg_comm_create(ctypes.byref(self.handle), rank, world, tables[0], tables[1], tables[2]), "pdmrg_comm_create")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def fits(self, z_bytes, y_bytes):
        return self.handle is not None and z_bytes <= self.z_bytes and y_bytes <= self.y_bytes

    def check(self):
        """Raise if a flag wait of this rank timed out (a peer died or diverged)."""
        import ctypes
        from . import device as D
        from ._lib import check
        err = ctypes.c_ulonglong(0)
        check(self._lib.pdmrg_comm_error(self.handle, ctypes.byref(err), D._stream()), "pdmrg_comm_error")
        if err.value:
            raise RuntimeError("sharded mat-vec: a peer did not signal epoch %d within the time-out (rank %d)" % (err.value, self.rank))

    def close(self):
        import torch
        if getattr(self, "handle", None):
            torch.cuda.synchronize()
            self._lib.pdmrg_comm_destroy(self.handle)
            self.handle = None
            for p in self._opened:
                self._lib.pdmrg_ipc_close(p)
            for p in self._own:
                self._lib.pdmrg_ipc_free(p)
            self._opened, self._own = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FusedShardedHeff:
    """MPO-bond-sharded mat-vec whose exchange is done by the kernels themselves over peer memory
    (CUDA IPC + NVLink loads/stores): stage-C GEMM with a scatter epilogue -> flag -> reduce + all-gather
    by peer stores -> flag -> copy.  No NCCL on the data path; ``dist`` is only used once, to ship the
    64-byte IPC handles and the slab masks between the ranks.  world == 1 runs the same kernels on
    the local buffers.  ``comm``: an existing PeerComm with large enough buffers (ket mode)."""

    def __init__(self, code, P, Dl, Dr, ops, workspace, rank, world, dist=None, mode="bond", comm=None):
        import ctypes
        from . import device as D
        from ._lib import check, lib
        assert len(ops) == 1, "one MPO term per fused plan"
        assert mode in ("bond", "ket")
        self.rank, self.world, self.mode = rank, world, mode
        L, Wc, R = ops[0]
        if mode == "bond":
            pat = block_pattern(Wc, L.shape[1], R.shape[1], P)
            self.partition = [partition_blocks(pat, world)]
            self.plan = D.HeffPlan(code, P, Dl, Dr, ops, workspace, block_masks=[self.partition[0][rank]])
        else:
            # ket split: stage C runs as one long-K GEMM whose scatter epilogue reduce-scatters the partial y
            self.partition = None
            self.plan = D.HeffPlan(code, P, Dl, Dr, ops, workspace, k_range=ket_chunks(Dl, world)[rank])
        self.n = self.plan.n
        self._lib, self._check, self._ct = lib, check, ctypes
        zb, yb = int(lib.pdmrg_comm_z_bytes(self.plan._plan, world)), int(lib.pdmrg_comm_y_bytes(self.plan._plan))
        self._own_comm = comm is None or not comm.fits(zb, yb)
        self.peer = PeerComm(zb, yb, rank, world, dist) if self._own_comm else comm
        self._comm = self.peer.handle
        mask = ctypes.c_uint(0)
        check(lib.pdmrg_heff_needed_j(self.plan._plan, 0, ctypes.byref(mask)), "pdmrg_heff_needed_j")
        if world > 1 and mode == "bond":
            everyone = [None] * world
            dist.all_gather_object(everyone, int(mask.value))
        else:
            everyone = [int(mask.value)] * world
        self._jmask = (ctypes.c_uint * world)(*everyone)

    def apply(self, x, y, sign=-1.0):
        from . import device as D
        pl = self.plan
        ws = pl.workspace.get(pl.ws_bytes)
        if self.mode == "ket":
            self._check(self._lib.pdmrg_heff_apply_fused_ket(pl._plan, self._comm, pl._Lp, pl._Rp, D._ptr(x), D._ptr(y), float(sign),
                                                             D._ptr(ws), ws.numel(), D._stream()), "pdmrg_heff_apply_fused_ket")
            return y
        self._check(self._lib.pdmrg_heff_apply_fused(pl._plan, self._comm, pl._Lp, pl._Rp, D._ptr(x), D._ptr(y), float(sign),
                                                     self._jmask, D._ptr(ws), ws.numel(), D._stream()), "pdmrg_heff_apply_fused")
        return y

    def close(self):
        if getattr(self, "_own_comm", False) and getattr(self, "peer", None) is not None:
            self.peer.close()
        self.peer = None
        self._comm = None
        if getattr(self, "plan", None) is not None:
            self.plan.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
