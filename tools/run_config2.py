"""BASELINE configs[2] for real on one B200: TASEP N=100 (alpha=0.35, beta=2/3, s=+0.5: current-enhanced, entangled),
non-Hermitian LEFT + RIGHT eigenstates, two-site sweeps through the chi schedule [16, 64, 256, 1024] (stages continued
through saved states, as the reference does), then the observables of dmrg.py:448-465 / tests/asep/contractCheck.py on
the chi=1024 states:  <l|H|r>/<l|r> = E (north-star: 1e-8),  current <l|J|r>/<l|r>  vs  dE/ds by central differences."""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import contract, core, dmrg, mpo as M

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 100
prm = (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, 0.5)
mpo = M.asep_mpo(N, prm)
ctx = core.Context("c128")
sched = [c for c in (16, 64, 256) if c < chi] + [chi]
tmp = tempfile.mkdtemp()
np.random.seed(0)
st = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
E, EE, gap, EEs, mps = dmrg.run_dmrg(mpo, mbd=sched, tol=1e-9, maxIter=2, fname=os.path.join(tmp, "c2"), twoSite=True, calcLeftState=True,
                                     returnState=True, returnEntSpec=True, ctx=ctx, stats=st)
torch.cuda.synchronize(); t_run = time.perf_counter() - t0
r, l = mps
t0 = time.perf_counter()
den = contract.full_contract(mps=r, lmps=l, ctx=ctx)
ene = contract.full_contract(mpo=mpo, mps=r, lmps=l, ctx=ctx) / den
cur = contract.full_contract(mpo=M.asep_curr_mpo(N, prm), mps=r, lmps=l, ctx=ctx) / den
rr = contract.full_contract(mpo=mpo, mps=r, ctx=ctx) / contract.full_contract(mps=r, ctx=ctx)
torch.cuda.synchronize(); t_obs = time.perf_counter() - t0
# dE/ds at a chi the derivative runs can afford (the state needs ~chi = 128 for 1e-12 weights): continuation from the right state
h, small = 1e-3, min(chi, 128)
Es = []
for ds in (-h, h):
    out = dmrg.run_dmrg(M.asep_mpo(N, prm[:6] + (prm[6] + ds,)), mbd=small, tol=1e-11, maxIter=3, twoSite=True, ctx=ctx,
                        initGuess=[[np.array(t) for t in r[0]]], res_tol=1e-10)
    Es.append(out[0])
dEds = (Es[1] - Es[0]) / (2 * h)
E = np.atleast_1d(E)
Ef = E[-1]
rec = {"config": "TASEP N=100 alpha=0.35 beta=2/3 s=+0.5, two-site, calcLeftState=True, mbd=%s" % sched, "E_by_stage": [complex(e).real for e in E],
       "E_imag_last": complex(Ef).imag, "EE_right_by_stage": [float(np.real(e)) for e in np.atleast_1d(EE if len(sched) > 1 else EE[0])],
       "run_sec": t_run, "n_matvec": st.get("n_matvec"), "n_recycled": st.get("n_recycled"), "n_exact_trunc": st.get("n_exact_trunc"),
       "observables_sec": t_obs,
       "lHr_over_lr": [complex(ene).real, complex(ene).imag], "abs_lHr_minus_E_rel": abs(ene - Ef) / abs(Ef),
       "rHr_over_rr": [complex(rr).real, complex(rr).imag], "abs_rHr_minus_E_rel": abs(rr - Ef) / abs(Ef),
       "current_lJr_over_lr": [complex(cur).real, complex(cur).imag], "dE_ds_central_difference": [complex(dEds).real, complex(dEds).imag],
       "abs_current_minus_dEds_rel": abs(cur - dEds) / max(1.0, abs(cur)), "max_bond": int(max(t.shape[2] for t in r[0]))}
print(json.dumps(rec), flush=True)
