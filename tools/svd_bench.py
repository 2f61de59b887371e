"""Timing of the truncation back-ends (development aid): gesvd vs the Gram route, heevd alone."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import device as D

def t(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

ws = D.Workspace(torch.device("cuda"))
for n in [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024, 2048]:
    for dt in (torch.complex128, torch.float64):
        a = torch.randn(n, n, dtype=dt, device="cuda")
        u, s, vh = torch.linalg.svd(a)
        r = {}
        for name, decay in (("slow", 8.0), ("fast", 36.0)):
            s = torch.exp(-torch.arange(n, device="cuda", dtype=torch.float64) * (decay / n))
            m = ((u * s.to(dt)) @ vh).contiguous()
            r["gesvd_" + name] = t(lambda: D.svd(m.clone(), ws, 0), 2)
            r["gram_" + name] = t(lambda: D.svd(m.clone(), ws, 2, keep=n // 2))
            r["gram_full_" + name] = t(lambda: D.svd(m.clone(), ws, 2))
        h = (m @ m.conj().T).contiguous()
        r["heevd"] = t(lambda: D.heevd(h.clone(), ws))
        r["torch.qr_half"] = t(lambda: torch.linalg.qr(m[:, : n // 2].contiguous()))
        print(n, str(dt).split(".")[-1], {k: round(v, 2) for k, v in r.items()}, flush=True)
