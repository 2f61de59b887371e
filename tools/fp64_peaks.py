"""Measure the FP64 roofline denominators on this box: cuBLAS Dgemm / Zgemm through
torch.matmul (library GEMMs, used only as the measured peak), burst and sustained.
Writes gpurun_out/fp64_peaks.json.  Not part of the product path."""
import json, os, time, sys
import torch

def bench(dtype, n, reps=5, sustain_s=0.0):
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = (8.0 if dtype.is_complex else 2.0) * n ** 3
    best = 0.0
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
    sus = None
    if sustain_s > 0:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time(); k = 0
        e0.record()
        while time.time() - t0 < sustain_s:
            for _ in range(4):
                torch.matmul(a, b, out=c); k += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        sus = flop * k / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    return best, sus

out = {"gpu": torch.cuda.get_device_name(0)}
for name, dt, n in [("dgemm_4096", torch.float64, 4096), ("dgemm_8192", torch.float64, 8192),
                    ("zgemm_4096", torch.complex128, 4096), ("zgemm_8192", torch.complex128, 8192)]:
    b, s = bench(dt, n, sustain_s=3.0 if n == 8192 else 0.0)
    out[name + "_tflops_burst"] = b
    if s is not None:
        out[name + "_tflops_sustained"] = s
    print(name, b, s, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
