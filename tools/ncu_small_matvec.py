"""ncu target: a few two-site mat-vecs at chi = 256 / 512 (f64 Heisenberg w=5, c128 ASEP w=6) for per-kernel GPU durations."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import core, device as D, mpo as M
for tag, code, dt, mk in (("f64", D.F64, torch.float64, lambda: M.heisenberg_mpo(6)), ("c128", D.C128, torch.complex128, lambda: M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)))):
    ctx = core.Context(tag)
    W1, W2 = mk()[0][2:4]
    w = W1.shape[0]
    for chi in (256, 512):
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        rnd = lambda *s: torch.randn(*s, dtype=dt, device="cuda", generator=g)
        L, R, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(4 * chi * chi)
        y = torch.empty_like(x)
        plan = D.HeffPlan(code, 4, chi, chi, [(L, D.combine_twosite(W1, W2, D.np_dtype(code)), R)], ctx.ws)
        for _ in range(3):
            plan.apply(x, y)
        torch.cuda.synchronize()
        plan.close()
