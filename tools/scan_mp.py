"""Throughput of P concurrent scan workers sharing ONE GPU (development probe): the small-chi stages of a cold
scan point are launch-latency bound, so independent points might overlap on the same device.

Measured (B200, no MPS daemon): P=1 535 points/hour, P=2 381, P=4 385, P=8 386 -- separate processes are
time-sliced by the driver at a granularity far above a Davidson cycle, so sharing a GPU between processes
LOSES throughput.  One rank per GPU it is (concurrency inside one process would need a re-entrant library)."""
import sys, time, os
import multiprocessing as mp
import numpy as np

def worker(idx, npts, q, barrier):
    import torch
    sys.path.insert(0, ".")
    from pydmrg_b200 import mpo as M, scan
    s_vals = np.linspace(-0.5, 0.5, 64)[idx * npts:(idx + 1) * npts]
    mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    np.random.seed(idx)
    scan.run_scan(mpo_of, s_vals[:1], 64, maxIter=0, tol=1e-8)        # warm-up: library init, allocator
    barrier.wait()
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_vals, 256, maxIter=3, tol=1e-8)
    torch.cuda.synchronize()
    q.put((idx, time.perf_counter() - t0, [float(np.real(e)) for e in res["E"]]))

if __name__ == "__main__":
    mp.set_start_method("spawn")
    npts = 2
    for P in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
        q = mp.Queue(); barrier = mp.Barrier(P)
        procs = [mp.Process(target=worker, args=(i, npts, q, barrier)) for i in range(P)]
        for p in procs: p.start()
        outs = [q.get() for _ in procs]
        for p in procs: p.join()
        wall = max(o[1] for o in outs)
        print("P=%d workers on one GPU: %d points in %.2f s -> %.0f s-points/hour/GPU" % (P, P * npts, wall, P * npts / wall * 3600), flush=True)
