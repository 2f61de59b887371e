"""CPU experiment (NumPy, on top of the oracle): how many mat-vecs would a diagonal (Jacobi-Davidson style)
preconditioner  t = r / (diag(H_eff) - theta)  save over the identity preconditioner the reference uses
(tools/diag_tools.py:137-139)?  Two-site sweeps of an ASEP chain from a random start; prints mat-vecs per sweep.

Measured here (N=20; mat-vecs of sweeps 1, 2, 3):
  s=-0.5 chi=8 : identity 3427 / 1876 / 1816    diagonal 2134 / 1269 / 1241
  s=-0.5 chi=16: identity 3708 /  637 /  548    diagonal 1957 /  518 /  488
  s=+0.2 chi=8 : identity 2027 /   46 /   38    diagonal 1106 /   39 /   38
  s=+0.2 chi=16: identity 1141 /   38 /   38    diagonal  605 /   38 /   38
i.e. the cold sweeps that dominate a scan point need 1.6-1.9x fewer cycles; energies agree within the solver floor.
diag(H_eff)[p1,p2,k,s] = sum_{j,o} L[k,j,k] (sum_l W1[j,l,p1,p1] W2[l,o,p2,p2]) R[s,o,s] is an O(d^2 w^2 chi^2) kernel.
Not in the product yet (needs a device kernel for the diagonal and the division).

    python tests/experiments/precond_experiment.py

(lives under tests/ because it runs on the oracle, which only test code may import)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.linalg as sla
from oracle import dmrg_oracle as orc
from pydmrg_b200 import mpo as M

def davidson(matvec, x0, diag=None, nroots=1, max_space=12, lindep=1e-14, max_cycle=1000, mode="dav"):
    max_space = max_space + 3*nroots
    fresh=True; n_mv=0; heff=None
    xs=[];ax=[];space=0
    for cyc in range(max_cycle):
        if fresh:
            xs,ax,space=[],[],0
            xt=orc._gram_schmidt_list(x0)
        axt=[matvec(v) for v in xt]; n_mv+=len(xt)
        xs.extend(xt); ax.extend(axt)
        rnow=len(xt); head,space=space,space+rnow
        if heff is None: heff=np.zeros((max_space+nroots,)*2,complex)
        for i in range(rnow):
            for k in range(rnow): heff[head+k,head+i]=np.vdot(xt[k],axt[i])
        for i in range(head):
            for k in range(rnow):
                heff[head+k,i]=np.vdot(xt[k],ax[i]); heff[i,head+k]=np.vdot(xs[i],axt[k])
        w,v=sla.eig(heff[:space,:space])
        idx=np.argsort(w.real); e=w[idx][:nroots]; v=v[:,idx][:,:nroots]
        x0=[sum(v[i,k]*xs[i] for i in range(space)) for k in range(nroots)]
        ax0=[sum(v[i,k]*ax[i] for i in range(space)) for k in range(nroots)]
        xt=[ax0[k]-e[k]*x0[k] for k in range(nroots)]
        keep=[]
        for k in range(nroots):
            r=xt[k]
            if np.vdot(r,r).real>lindep:
                if diag is not None:
                    den=diag-e[k]
                    den=np.where(np.abs(den)<1e-3, 1e-3*np.exp(1j*np.angle(den+1e-30)), den)
                    r=r/den
                keep.append(r/np.linalg.norm(r))
        xt=keep
        for i in range(space):
            for r in xt: r-=xs[i]*np.vdot(xs[i],r)
        keep=[]
        for r in xt:
            nrm=np.linalg.norm(r)
            if nrm**2>lindep: keep.append(r/nrm)
        xt=keep
        if not xt: break
        fresh = space+nroots>max_space
    return e,x0,n_mv

def sweep(N,chi,prm,use_diag,nsw=3,seed=0):
    mpo=M.asep_mpo(N,prm); np.random.seed(seed)
    mps=orc.make_mps_right(orc.random_mps(N,chi)); env=orc.build_env(mps,mpo,chi,gauge_site=0)
    tot=[]
    for sw in range(nsw):
        nmv=0
        for site,direction in [(s,'right') for s in range(N-1)]+[(s,'left') for s in range(N-2,-1,-1)]:
            A0,B0=mps[site],mps[site+1]; d1,Dl,_=A0.shape; d2,_,Dr=B0.shape; shape=(d1,d2,Dl,Dr)
            Ls=[F[site] for F in env]; Rs=[F[site+2] for F in env]; W1s=[op[site] for op in mpo]; W2s=[op[site+1] for op in mpo]
            mv=lambda x: orc.heff_twosite(x,Ls,W1s,W2s,Rs,shape,"blas")
            guess=np.einsum("ijk,lkm->iljm",A0,B0).reshape(-1).astype(complex)
            diag=None
            if use_diag:
                L,R,W1,W2=Ls[0],Rs[0],W1s[0],W2s[0]
                Ld=np.einsum('kjk->kj',L); Rd=np.einsum('sos->so',R)
                Wd=np.einsum('jlmm,lopp->mpjo',W1,W2)      # diag in physical indices
                diag=-np.einsum('kj,mpjo,so->mpks',Ld,Wd,Rd).reshape(-1)     # matvec is -H
            e,x,k=davidson(mv,[guess],diag=diag); nmv+=k
            psi=x[0]
            A,B,S,EE,EEs=orc.svd_truncate_twosite(psi,shape,chi); nrm=np.linalg.norm(psi)
            if direction=='right':
                mps[site]=A; mps[site+1]=B*(S*nrm)[None,:,None]
                for m,op in enumerate(mpo): env[m][site+1]=orc.env_update_right(A,op[site],env[m][site],None,"blas")
            else:
                mps[site]=A*(S*nrm)[None,None,:]; mps[site+1]=B
                for m,op in enumerate(mpo): env[m][site+1]=orc.env_update_left(B,op[site+1],env[m][site+2],None,"blas")
        tot.append((nmv,-e[0]))
    return tot

for s in (-0.5, 0.2):
    prm=(0.5,0.5,0.1,0.9,0.5,0.5,s)
    for chi in (8,16):
        t=time.time(); a=sweep(20,chi,prm,False); b=sweep(20,chi,prm,True)
        print('s',s,'chi',chi,'identity',[(n,'%.10f'%E.real) for n,E in a],'diag',[(n,'%.10f'%E.real) for n,E in b],'%.0fs'%(time.time()-t))
