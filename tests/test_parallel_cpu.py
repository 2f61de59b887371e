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
