"""Truncation / renormalisation on the device.

T2  two-site SVD   -- renorm_inf, pydmrg/tools/mps_tools.py:98-120 (+ calc_entanglement :9-15)
T1  one-site RDM   -- renormalizeR / renormalizeL, pydmrg/dmrg.py:49-136 (+ calcRDM :18-28)

Site tensors are defined only up to a unitary on the cut bond (the reference even applies a
LAPACK-dependent rotation through sla.orth, dmrg.py:76,120); parity is asserted on
gauge-invariant quantities (Schmidt values, EE, projectors), never element-wise.
"""
import numpy as np
import torch

from . import device as D
from .core import default_context, is_dev, like


def calc_entanglement(S):
    """tools/mps_tools.py:9-15 on the host (<= chi numbers).  Returns EE, EEspec, S_normalised."""
    S = np.asarray(S, dtype=np.float64)
    S = S / np.sqrt(np.dot(S, S))
    p = S * S
    with np.errstate(divide="ignore", invalid="ignore"):
        spec = -p * np.log2(p)
    spec = np.where(p > 0, spec, 0.0)
    return float(np.sum(spec)), spec, S


def svd_truncate(psi, shape, mbd, ctx, absorb=None):
    """psi (d1,d2,Dl,Dr) device vector -> A (d1,Dl,k), B (d2,k,Dr), S_normalised (host), EE, EEs.

    ``absorb`` = None (iDMRG: S dropped, mps_tools.py:110-119), 'right' (B <- diag(S) B) or
    'left' (A <- A diag(S)); the absorbed S is rescaled to keep ||psi|| after truncation."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    m = ctx.empty(rows, cols)
    # m[(k,n),(q,s)] = psi[n,q,k,s]      (mps_tools.py:105-107)
    D.permute4(psi, m, (Dl, d1, d2, Dr), (Dr, d2 * Dl * Dr, Dl * Dr, 1))
    kfull = min(rows, cols)
    U, S, Vh, info = D.svd(m, ctx.svd_ws, ctx.svd_method_for(rows, cols), keep=min(int(mbd), kfull))
    Sh = S.cpu().numpy()                                        # sync; also validates the factorisation
    if int(info.item()) != 0:
        raise ArithmeticError("cuSOLVER SVD did not converge (info=%d)" % int(info.item()))
    k = min(int(mbd), kfull)
    EE, EEs, Sn = calc_entanglement(Sh[:k])
    scale = None
    if absorb is not None:
        # ||psi|| (the Gram route returns only the kept singular values; Davidson vectors are unit norm
        # up to rounding, and the absorbed scale only sets the norm of the next guess)
        full = max(np.sqrt(np.dot(Sh, Sh)), np.sqrt(np.dot(Sh[:k], Sh[:k])))
        scale = torch.from_numpy(np.ascontiguousarray(Sn * full)).to(ctx.device)
    A = ctx.empty(d1, Dl, k)
    B = ctx.empty(d2, k, Dr)
    # A[n,kl,a] = U[(kl,n),a]    (mps_tools.py:109-112);  B[q,a,s] = Vh[a,(q,s)]   (:113-115)
    D.permute4(U, A, (d1, Dl, k), (kfull, d1 * kfull, 1), scale=scale if absorb == "left" else None, scale_axis=2)
    D.permute4(Vh, B, (d2, k, Dr), (Dr, cols, 1), scale=scale if absorb == "right" else None, scale_axis=1)
    return A, B, Sn, EE, EEs


def eig_truncate(psi, shape, mbd, ctx, direction):
    """Projective two-site truncation for a MOVING centre (finite sweeps): only the isometry of the
    side being left behind is needed, so ONE Hermitian eigen-decomposition of the reduced density
    matrix replaces the SVD (same construction as the reference's one-site renormalizeR/L,
    dmrg.py:49-136, applied to the two-site wave-function):

      right:  rho = m m^H,  A = top-k eigenvectors (exact isometry),  new centre = A^H m  (= S Vh)
      left :  rho = m^H m,  B = top-k eigenvectors^H,                 new centre = m B^H  (= U S)

    with m[(Dl,d1),(d2,Dr)] = psi (mps_tools.py:105-107).  A (A^H m) is the exact projection of psi
    on the kept subspace; Schmidt weights p_i = lambda_i / sum(lambda) are accurate to eps uniformly.
    Returns A (d1,Dl,k), B (d2,k,Dr), S_normalised, EE, EEs like svd_truncate."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    k = min(int(mbd), rows, cols)
    m = ctx.empty(rows, cols)
    D.permute4(psi, m, (Dl, d1, d2, Dr), (Dr, d2 * Dl * Dr, Dl * Dr, 1))
    mT = ctx.empty(cols, rows)
    D.permute4(m, mT, (cols, rows), (1, cols))
    X = m if direction == "right" else mT                       # left: eigenvectors of mT mT^H are conj(v_a)
    n = X.shape[0]
    res = D.left_basis(X, ctx.svd_ws, keep=k)
    if res is None:                                             # cuSOLVER did not converge: LAPACK-grade route
        return svd_truncate(psi, shape, mbd, Context_with(ctx, D.SVD_GESVD), absorb=direction)
    Ut, lam_dev, info = res
    lam = np.clip(lam_dev.cpu().numpy()[:k], 0.0, None)
    EE, EEs, Sn = calc_entanglement(np.sqrt(lam))
    A = ctx.empty(d1, Dl, k)
    B = ctx.empty(d2, k, Dr)
    if direction == "right":
        # A[n,kl,a] = u_a[kl*d1+n];  centre C[a,c] = sum_x conj(u_a[x]) m[x,c]
        D.permute4(Ut, A, (d1, Dl, k), (1, d1, n))
        Vc = ctx.empty(k, rows)
        D.permute4(Ut, Vc, (k, rows), (n, 1), conj=True)
        C = D.gemm_nt(Vc, mT)
        D.permute4(C, B, (d2, k, Dr), (Dr, cols, 1))
    else:
        # eigenvector rows w_a = conj(v_a):  B[q,a,s] = Vh[a,(q,s)] = w_a[(q,s)];  centre = m V = m conj(W)^T
        D.permute4(Ut, B, (d2, k, Dr), (Dr, n, 1))
        C = D.gemm_nt(m, Ut, M=rows, N=k, K=cols, lda=cols, ldb=n, ldc=k, conjB=True)      # (rows x k)
        D.permute4(C, A, (d1, Dl, k), (k, d1 * k, 1))
    return A, B, Sn, EE, EEs


def Context_with(ctx, svd_method):
    """Shallow view of a context with another SVD back-end (shares workspaces)."""
    import copy
    c = copy.copy(ctx)
    c.svd_method = svd_method
    return c


def truncate_twosite(psi, shape, mbd, ctx, direction):
    """Truncation used by the finite two-site sweep: projective (one heevd) for large cuts, full SVD
    (cuSOLVER gesvd) for small ones or when an explicit SVD back-end was requested."""
    d1, d2, Dl, Dr = shape
    if ctx.svd_method < 0 and min(Dl * d1, d2 * Dr) >= 128:
        return eig_truncate(psi, shape, mbd, ctx, direction)
    return svd_truncate(psi, shape, mbd, ctx, absorb=direction)


def renorm_inf(mpsList, mpoList, envList, vecs, mbd, ctx=None):
    """Drop-in for tools/mps_tools.py:98-120 (nStates handled state by state)."""
    ctx = ctx or default_context()
    n1, n2 = mpoList[0][0].shape[2], mpoList[0][1].shape[2]
    n3, n4 = envList[0][0].shape[0], envList[0][1].shape[0]
    proto = mpsList[0][0]
    V = ctx.dev(vecs)
    EE = EEs = S = None
    for state in range(len(mpsList)):
        psi = V[:, state].contiguous() if V.dim() == 2 else V
        A, B, S, EE, EEs = svd_truncate(psi, (n1, n2, n3, n4), mbd, ctx)
        mpsList[state][0] = like(A, proto)
        mpsList[state][1] = like(B, proto)
    return mpsList, EE, EEs, S


# ---------------------------------------------------------------------------------------
# one-site reduced-density-matrix route
# ---------------------------------------------------------------------------------------
def _spectrum_from_rdm(w, keep):
    """Schmidt spectrum of the cut from RDM eigenvalues (ascending from heevd): the reference takes
    the singular values of the whole (d*Dl, Dr) matrix (dmrg.py:30-37), i.e. sqrt of every eigenvalue."""
    lam = np.clip(w[::-1], 0.0, None)                           # descending, >= 0
    return calc_entanglement(np.sqrt(lam))


def rdm_renormalize_right(psis, shape, nxts, ctx, n_avg=None):
    """renormalizeR, dmrg.py:49-91: equal-weight RDM over ``psis`` (state averaging, :56-66),
    eigenvectors (Hermitian solver instead of the reference's general eig -- the RDM is
    Hermitian PSD), top-Dr kept; A <- eigenvectors; neighbour <- (A^dagger psi) . M[site+1] (:90).
    Returns A, [new neighbours], EE, EEs."""
    d, Dl, Dr = shape
    nn = Dl * d
    n_avg = len(psis) if n_avg is None else n_avg
    Ms = []
    rho = ctx.empty(nn, nn)
    for i in range(n_avg):
        Mi = ctx.empty(nn, Dr)
        D.permute4(psis[i], Mi, (Dl, d, Dr), (Dr, Dl * Dr, 1))   # rows (Dl,d): swapaxes(0,1), dmrg.py:21-22
        D.gemm_nt(Mi, Mi, rho, M=nn, N=nn, K=Dr, lda=Dr, ldb=Dr, ldc=nn, conjB=True, alpha=1.0 / n_avg, beta=int(i > 0))
        Ms.append(Mi)
    w, vec, info = D.heevd(rho, ctx.svd_ws)
    wh = w.cpu().numpy()
    if int(info.item()) != 0:
        raise ArithmeticError("cuSOLVER heevd failed (info=%d)" % int(info.item()))
    keep = min(Dr, nn)
    # entanglement of the target state alone (dmrg.py:52) -- for one state it is the RDM spectrum
    if n_avg == 1:
        EE, EEs, _ = _spectrum_from_rdm(wh, keep)
    else:
        EE, EEs = _entropy_of(Ms[0], ctx)
    # top-`keep` eigenvectors are the LAST rows of vec; A[n,kl,a] = vec[nn-1-a, kl*d+n]
    A = ctx.empty(d, Dl, keep)
    last = vec[nn - 1]
    D.permute4(last, A, (d, Dl, keep), (1, d, -nn))
    Vc = ctx.empty(keep, nn)                                    # conj of the kept eigenvectors, row a
    D.permute4(last, Vc, (keep, nn), (-nn, 1), conj=True)
    new_next = []
    for s, nxt in enumerate(nxts):
        Mi = Ms[min(n_avg - 1, s)]
        Mt = ctx.empty(Dr, nn)
        D.permute4(Mi, Mt, (Dr, nn), (1, Dr))
        C1 = D.gemm_nt(Vc, Mt)                                  # C1[a,k] = sum_x conj(vec_a[x]) M[x,k]
        dn, Dk, Dj = nxt.shape
        Nt = ctx.empty(dn, Dj, Dk)
        D.permute4(nxt, Nt, (dn, Dj, Dk), (Dk * Dj, 1, Dj))
        out = ctx.empty(dn, keep, Dj)
        D.gemm_nt(C1, Nt, out, M=keep, N=Dj, K=Dk, batch=dn, lda=Dk, ldb=Dk, ldc=Dj, strideA=0,
                  strideB=Dj * Dk, strideC=keep * Dj)
        new_next.append(out)
    return A, new_next, EE, EEs


def rdm_renormalize_left(psis, shape, prvs, ctx, n_avg=None):
    """renormalizeL, dmrg.py:93-136.  Returns B, [new previous-site tensors], EE, EEs."""
    d, Dl, Dr = shape
    nn = d * Dr
    n_avg = len(psis) if n_avg is None else n_avg
    Mts = []
    rho = ctx.empty(nn, nn)
    for i in range(n_avg):
        Mt = ctx.empty(nn, Dl)                                  # Mt[(n,s),kl] = psi[n,kl,s]
        D.permute4(psis[i], Mt, (d, Dr, Dl), (Dl * Dr, 1, Dr))
        # rho[j,k] = sum_i M[i,j] conj(M[i,k])    (dmrg.py:28)
        D.gemm_nt(Mt, Mt, rho, M=nn, N=nn, K=Dl, lda=Dl, ldb=Dl, ldc=nn, conjB=True, alpha=1.0 / n_avg, beta=int(i > 0))
        Mts.append(Mt)
    w, vec, info = D.heevd(rho, ctx.svd_ws)
    wh = w.cpu().numpy()
    if int(info.item()) != 0:
        raise ArithmeticError("cuSOLVER heevd failed (info=%d)" % int(info.item()))
    keep = min(Dl, nn)
    if n_avg == 1:
        EE, EEs, _ = _spectrum_from_rdm(wh, keep)
    else:
        EE, EEs = _entropy_of(Mts[0], ctx)
    last = vec[nn - 1]
    # B[n,a,s] = vec_a[(n,s)]    (dmrg.py:121-126)
    B = ctx.empty(d, keep, Dr)
    D.permute4(last, B, (d, keep, Dr), (Dr, -nn, 1))
    Vc = ctx.empty(keep, nn)
    D.permute4(last, Vc, (keep, nn), (-nn, 1), conj=True)
    new_prev = []
    for s, prv in enumerate(prvs):
        Mt = Mts[min(n_avg - 1, s)]
        Mk = ctx.empty(Dl, nn)                                  # M[kl,(n,s)]
        D.permute4(Mt, Mk, (Dl, nn), (1, Dl))
        Ct = D.gemm_nt(Vc, Mk)                                  # Ct[a,k] = sum_x conj(vec_a[x]) psi[k,x]
        dp, Di, Dk = prv.shape
        out = ctx.empty(dp, Di, keep)
        D.gemm_nt(prv, Ct, out, M=dp * Di, N=keep, K=Dk, lda=Dk, ldb=Dk, ldc=keep)   # dmrg.py:135
        new_prev.append(out)
    return B, new_prev, EE, EEs


def _entropy_of(Mmat, ctx):
    """Entanglement of one state from its own singular values (dmrg.py:30-47)."""
    a = Mmat.clone()
    _, S, _, info = D.svd(a, ctx.svd_ws, ctx.svd_method_for(*a.shape))
    EE, EEs, _ = calc_entanglement(S.cpu().numpy())
    return EE, EEs
