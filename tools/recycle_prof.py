"""Kernel-level breakdown of one warm-started truncation step (development aid; torch.profiler)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import core, truncate
from torch.profiler import profile, ProfilerActivity

for dtype, chi in (("c128", 1024), ("c128", 256), ("f64", 256)):
    ctx = core.Context(dtype)
    d1 = d2 = 2; Dl = Dr = chi; k = chi
    rows, cols = Dl * d1, d2 * Dr
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rnd = lambda *s: torch.randn(*s, dtype=ctx.tdtype, device="cuda", generator=g)
    U, _ = torch.linalg.qr(rnd(rows, rows)); V, _ = torch.linalg.qr(rnd(cols, cols))
    s = torch.exp(-torch.arange(rows, device="cuda", dtype=torch.float64) * (30.0 / rows)).to(ctx.tdtype)
    m = (U * s) @ V.mH
    psi = m.reshape(Dl, d1, d2, Dr).permute(1, 2, 0, 3).reshape(-1).contiguous()
    p = truncate.spare_count(k, rows, cols)
    Wfull = V[:, : k + p].mH.contiguous()
    warm = Wfull[:k].reshape(k, d2, Dr).permute(1, 0, 2).contiguous()
    ext = Wfull[k:].contiguous()
    f = lambda: truncate.recycled_truncate(psi, (d1, d2, Dl, Dr), k, ctx, "right", warm, ext)
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5 * 1e3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        f(); torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in ev) / 1e3
    print("== %s chi=%d: wall %.2f ms per step, GPU kernel time %.2f ms, %d kernel launches" % (dtype, chi, wall, tot, sum(e.count for e in ev)))
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]:
        print("   %8.3f ms  x%-4d %s" % (e.device_time_total / 1e3, e.count, e.key[:110]))
