"""Timing of Householder QR (cuSOLVER via torch) for tall graded matrices (development aid)."""
import sys, time, torch
def t(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for dt in (torch.complex128, torch.float64):
    for n, k in ((512, 256), (512, 320), (1024, 640), (2048, 1024), (2048, 1280), (4096, 2560)):
        a = torch.randn(n, k, dtype=dt, device="cuda")
        r = {"qr": t(lambda: torch.linalg.qr(a)),
             "geqrf": t(lambda: torch.geqrf(a)),
             "chol_k": t(lambda: torch.linalg.cholesky((a.mH @ a))),
             "eigh_k/4": t(lambda: torch.linalg.eigh((a[:, : k // 4].mH @ a[:, : k // 4]))),
             "svd_kxn": t(lambda: torch.linalg.svd(a, full_matrices=False), 2)}
        print(str(dt).split(".")[-1], n, k, {q: round(v, 2) for q, v in r.items()}, flush=True)
