"""Where does one cold scan point (ASEP N=50 chi=256) spend its time?  (development aid)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import dmrg, mpo as M, scan
mpo = M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
# capping the Davidson cycles per bond in the ramp (max_cycle=4..8) was tried and rejected: the final stage
# then needs 15-40x more mat-vecs and can settle on the wrong branch of the non-Hermitian generator
for rep, rc in enumerate([None, None]):
    np.random.seed(0)
    guess = None
    tt = time.perf_counter()
    for chi in (16, 64, 256):
        last = chi == 256
        st = {"profile": True}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = dmrg.run_dmrg(mpo, mbd=chi, tol=1e-8, maxIter=3 if last else 0, initGuess=guess, twoSite=True, returnState=True,
                            stats=st, **({} if last else dict(res_tol=1e-5, **({"max_cycle": rc} if rc else {}))))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        nxt = {16: 64, 64: 256, 256: 256}[chi]
        guess = out[-1]                      # two-site stages are not zero-padded: the bonds grow on their own
        t2 = time.perf_counter()
        print("rep", rep, "chi", chi, "run_dmrg %.3f s  grow %.3f s" % (t1 - t0, t2 - t1),
              {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k != "profile"}, flush=True)
    print("   total %.3f s  E %.12f EE %.10f" % (time.perf_counter() - tt, out[0].real, out[1]))
