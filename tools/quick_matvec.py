"""Quick timing of the two-site H_eff mat-vec and its GEMM stages (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import device as D, mpo as M

def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ws = D.Workspace(torch.device("cuda"))
for chi in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    for name, dt, code in (("c128", torch.complex128, D.C128), ("f64", torch.float64, D.F64)):
        W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
        Wc = D.combine_twosite(W1, W2, D.np_dtype(code))
        L = torch.randn(chi, 6, chi, dtype=dt, device="cuda")
        R = torch.randn(chi, 6, chi, dtype=dt, device="cuda")
        plan = D.HeffPlan(code, 4, chi, chi, [(L, Wc, R)], ws)
        x = torch.randn(4 * chi * chi, dtype=dt, device="cuda"); y = torch.empty_like(x)
        ms = timeit(lambda: plan.apply(x, y))
        print(f"chi={chi} {name}: matvec {ms:.3f} ms  dense {plan.flops_dense/ms*1e-9:.2f} TFLOP/s  sparse-count {plan.flops_sparse/ms*1e-9:.2f}", flush=True)
        # stage A alone
        T = torch.empty(chi * 6, 4 * chi, dtype=dt, device="cuda")
        f = lambda: D.gemm_nt(R, x, T, M=chi * 6, N=4 * chi, K=chi, lda=chi, ldb=chi, ldc=4 * chi)
        ms = timeit(f)
        fl = (8 if code == D.C128 else 2) * chi * 6 * 4 * chi * chi
        print(f"   stageA-like GEMM {chi*6}x{4*chi}x{chi}: {ms:.3f} ms {fl/ms*1e-9:.2f} TFLOP/s")
        U = torch.randn(4, chi, 6 * chi, dtype=dt, device="cuda")
        f = lambda: D.gemm_nt(L, U, y, M=chi, N=chi, K=6 * chi, batch=4, lda=6 * chi, ldb=6 * chi, ldc=chi, strideA=0, strideB=chi * 6 * chi, strideC=chi * chi)
        ms = timeit(f)
        print(f"   stageC-like GEMM 4x {chi}x{chi}x{6*chi}: {ms:.3f} ms {fl/ms*1e-9:.2f} TFLOP/s")
        del plan
# cuBLAS comparison
for n in (4096,):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda"); b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    ms = timeit(lambda: torch.matmul(a, b))
    print(f"cuBLAS zgemm {n}: {8*n**3/ms*1e-9:.2f} TFLOP/s")
    c = D.gemm_nt(a, b)
    ms = timeit(lambda: D.gemm_nt(a, b, c, M=n, N=n, K=n, lda=n, ldb=n, ldc=n))
    print(f"ours  zgemm {n}: {8*n**3/ms*1e-9:.2f} TFLOP/s")
