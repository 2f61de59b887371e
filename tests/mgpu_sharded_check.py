"""Run under torchrun with N >= 2 ranks (one per GPU): checks the two bond-sharded mat-vec paths."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    from pydmrg_b200 import core, device as D, mpo as M, parallel
    ctx = core.Context("c128")
    chi = int(os.environ.get("PDMRG_CHECK_CHI", "256"))
    W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.complex128)
    g = torch.Generator(device="cuda"); g.manual_seed(11)          # identical operands on every rank
    rnd = lambda *s: torch.randn(*s, dtype=torch.complex128, device="cuda", generator=g)
    L, R, x = rnd(chi, 6, chi), rnd(chi, 6, chi), rnd(4 * chi * chi)
    full = D.HeffPlan(D.C128, 4, chi, chi, [(L, Wc, R)], ctx.ws)
    y_ref = torch.empty_like(x); full.apply(x, y_ref)
    nccl = parallel.ShardedHeff(D.C128, 4, chi, chi, [(L, Wc, R)], ctx.ws, rank, world, dist)
    y_n = torch.empty_like(x); nccl.apply(x, y_n)
    ket = parallel.ShardedHeff(D.C128, 4, chi, chi, [(L, Wc, R)], ctx.ws, rank, world, dist, mode="ket")
    y_k = torch.empty_like(x); ket.apply(x, y_k)
    fused = parallel.FusedShardedHeff(D.C128, 4, chi, chi, [(L, Wc, R)], ctx.ws, rank, world, dist)
    y_f = torch.empty_like(x)
    for _ in range(4):
        fused.apply(x, y_f)
    fket = parallel.FusedShardedHeff(D.C128, 4, chi, chi, [(L, Wc, R)], ctx.ws, rank, world, dist, mode="ket")
    y_fk = torch.empty_like(x)
    for _ in range(4):
        fket.apply(x, y_fk)
    torch.cuda.synchronize()
    e_fk = ((y_fk - y_ref).norm() / y_ref.norm()).item()
    chk2 = torch.view_as_real(y_fk).clone()
    ref2 = chk2.clone(); dist.broadcast(ref2, 0)
    same_k = bool(torch.equal(chk2, ref2))
    e_n = ((y_n - y_ref).norm() / y_ref.norm()).item()
    e_f = ((y_f - y_ref).norm() / y_ref.norm()).item()
    e_k = ((y_k - y_ref).norm() / y_ref.norm()).item()
    # bit-identical across ranks
    chk = torch.view_as_real(y_f).clone()
    ref0 = chk.clone(); dist.broadcast(ref0, 0)
    same = bool(torch.equal(chk, ref0))
    # timing
    def timeit(f, n=10):
        for _ in range(3): f()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    t_full = timeit(lambda: full.apply(x, y_ref))
    t_nccl = timeit(lambda: nccl.apply(x, y_n))
    t_fused = timeit(lambda: fused.apply(x, y_f))
    t_ket = timeit(lambda: ket.apply(x, y_k))
    t_fket = timeit(lambda: fket.apply(x, y_fk))
    # stress: back-to-back applies with no host synchronisation in between (the exchange flags and the
    # GEMM's shared-memory ring are the only ordering); every batch must reproduce the reference
    worst = 0.0
    for _ in range(int(os.environ.get("PDMRG_CHECK_BATCHES", "12"))):
        for _ in range(4):
            fused.apply(x, y_f)
        for _ in range(4):
            fket.apply(x, y_fk)
        torch.cuda.synchronize()
        worst = max(worst, ((y_f - y_ref).norm() / y_ref.norm()).item(), ((y_fk - y_ref).norm() / y_ref.norm()).item())
    ok = e_n < 1e-13 and e_f < 1e-13 and e_k < 1e-13 and e_fk < 1e-13 and same and same_k and worst < 1e-13
    if rank == 0:
        print("chi %d world %d: err bond/nccl %.2e bond/fused %.2e ket/nccl %.2e ket/fused %.2e bit-identical %s %s stress %.2e | "
              "ms unsharded %.3f bond+nccl %.3f bond+fused %.3f ket+nccl %.3f ket+fused %.3f"
              % (chi, world, e_n, e_f, e_k, e_fk, same, same_k, worst, t_full, t_nccl, t_fused, t_ket, t_fket), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    fused.close()
    fket.close()
    dist.barrier()
    if rank == 0 and flag.item() == 1:
        print("SHARDED_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
