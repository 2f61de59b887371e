"""Timing of the truncation back-ends at the chi=1024 two-site size (2048 x 2048 complex128)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import device as D

def t(f, n=2):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ws = D.Workspace(torch.device("cuda"))
for n in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]:
    for dt in (torch.complex128, torch.float64):
        a = torch.randn(n, n, dtype=dt, device="cuda")
        # decaying spectrum like a physical wave-function
        u, s, vh = torch.linalg.svd(a)
        s = torch.exp(-torch.arange(n, device="cuda", dtype=torch.float64) * (30.0 / n))
        a = (u * s.to(dt)) @ vh
        r = {}
        r["gesvd"] = t(lambda: D.svd(a.clone(), ws, 0))
        if n <= 1024:
            r["gesvdj"] = t(lambda: D.svd(a.clone(), ws, 1), 1)
        h = a @ a.conj().T
        r["heevd"] = t(lambda: D.heevd(h.clone(), ws))
        r["torch.svd"] = t(lambda: torch.linalg.svd(a), 1)
        r["torch.eigh"] = t(lambda: torch.linalg.eigh(h), 1)
        r["torch.qr"] = t(lambda: torch.linalg.qr(a), 1)
        print(n, dt, {k: round(v, 1) for k, v in r.items()}, flush=True)
