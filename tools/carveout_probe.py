"""Is the ~10-20 us fixed cost per kernel at chi=256 the shared-memory carve-out flipping between the 192 KB GEMM and
the small kernels?  Times (CUDA events) the same GEMM back to back and interleaved with a small kernel."""
import sys, torch
sys.path.insert(0, ".")
from pydmrg_b200 import device as D
torch.manual_seed(0)
M, N, K, nb = 256, 1024, 256, 5
A = torch.randn(nb, M, K, dtype=torch.float64, device="cuda"); B = torch.randn(nb, N, K, dtype=torch.float64, device="cuda")
C = torch.empty(nb, M, N, dtype=torch.float64, device="cuda")
s = torch.randn(256, 256, dtype=torch.float64, device="cuda"); t = torch.empty_like(s)
gemm = lambda: D.gemm_nt(A, B, C, M=M, N=N, K=K, batch=nb, lda=K, ldb=K, ldc=N, strideA=M * K, strideB=N * K, strideC=M * N)
small = lambda: D.permute4(s, t, (256, 256), (1, 256))
def timeit(f, n=200):
    for _ in range(10): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mode in (0, 1):
    D.lib.pdmrg_set_streamk_mode(mode)
    print("streamk", mode, "gemm back-to-back %.1f us" % timeit(gemm), "| small kernel alone %.1f us" % timeit(small),
          "| gemm+small interleaved %.1f us" % timeit(lambda: (gemm(), small())), flush=True)
for K2 in (64, 128, 256, 512, 1024, 2048):
    A2 = torch.randn(nb, M, K2, dtype=torch.float64, device="cuda"); B2 = torch.randn(nb, N, K2, dtype=torch.float64, device="cuda")
    for mode in (0, 1):
        D.lib.pdmrg_set_streamk_mode(mode)
        g2 = lambda: D.gemm_nt(A2, B2, C, M=M, N=N, K=K2, batch=nb, lda=K2, ldb=K2, ldc=N, strideA=M * K2, strideB=N * K2, strideC=M * N)
        us = timeit(g2)
        print("K=%d streamk %d: %.1f us, %.1f TFLOP/s" % (K2, mode, us, 2.0 * M * N * K2 * nb / us * 1e-6), flush=True)
