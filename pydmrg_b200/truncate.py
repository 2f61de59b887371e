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
    first = ctx.svd_method_for(rows, cols)
    # cuSOLVER's QR-iteration gesvd occasionally reports non-converged superdiagonals (info > 0) on small
    # rank-deficient cuts (seen on the N=200 Heisenberg ramp); the one-sided Jacobi gesvdj, then the Gram
    # route, take over before giving up
    for method in [first] + [mth for mth in (D.SVD_GESVDJ, D.SVD_GRAM) if mth != first]:
        mm = m if method == D.SVD_GRAM else m.clone()           # gesvd / gesvdj destroy their input
        U, S, Vh, info = D.svd(mm, ctx.svd_ws, method, keep=min(int(mbd), kfull))
        Sh = S.cpu().numpy()                                    # sync; also validates the factorisation
        if int(info.item()) == 0 and np.all(np.isfinite(Sh)):
            break
    else:
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


RECYCLE_FRACTION = 8        # spare vectors p = window depth b = k // RECYCLE_FRACTION (at least 4)
RECYCLE_MIN_SPARE = 4
RECYCLE_SAFE_DISCARD = 1e-11   # below this discarded weight a bond is warm-started from its 2nd visit on, else from its 3rd


def window_depth(k):
    return max(RECYCLE_MIN_SPARE, k // RECYCLE_FRACTION)


def spare_count(k, rows, cols):
    """Number of spare basis vectors carried beyond the kept k for the recycled truncation."""
    return max(0, min(window_depth(k), min(rows, cols) - k))


def recycle_allowed(cache, site, shape_key):
    """A bond may be warm-started once it has been decomposed exactly AT ITS CURRENT SHAPE either once with
    (numerically) nothing discarded, or twice.  Bases taken from a state that is still far from converged and
    truncated leave a slowly decaying lag in the kept subspace (oracle/recycle_oracle.py: Heisenberg N=20,
    chi=32 lags the exact-SVD sweeps by 1e-10 for six sweeps when recycling starts at the second visit, and by
    1e-15 when it starts at the third)."""
    st = cache.get(("state", site))
    if st is None or st["shape"] != shape_key:
        return False
    return st["disc"] < RECYCLE_SAFE_DISCARD or st["exact_visits"] >= 2


def recycled_weight_ok(cache, site, disc_now):
    """After a warm-started step: the weight it discarded must not exceed what the bond's last exact decomposition
    discarded by more than a factor 2 (+1e-13 of rounding) -- a state that moved between two visits can leave the
    carried subspace suboptimal without any sign in the orthogonality defect; the caller then redoes the bond exactly."""
    st = cache.get(("state", site))
    if st is None:
        return True
    return disc_now <= 2.0 * st["disc"] + 1e-13


def note_exact_visit(cache, site, shape_key, disc):
    st = cache.get(("state", site))
    if st is None or st["shape"] != shape_key:
        st = {"shape": shape_key, "exact_visits": 0, "disc": 1.0}
    st["exact_visits"] += 1
    st["disc"] = float(disc)
    cache[("state", site)] = st


def eig_truncate(psi, shape, mbd, ctx, direction, spare=0, extras=None):
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
    res = D.left_basis(X, ctx.svd_ws, keep=k + spare)
    if res is None:                                             # cuSOLVER did not converge: LAPACK-grade route
        return svd_truncate(psi, shape, mbd, Context_with(ctx, D.SVD_GESVD), absorb=direction)
    Ut, lam_dev, info = res
    if extras is not None and spare > 0:
        # the next `spare` eigenvectors seed the recycled truncation of this bond's next visit: rows u_a
        # (left vectors) after a right-moving step, rows conj(v_a) (right vectors) after a left-moving one
        extras["ext"] = Ut[k:k + spare].clone()
    lam_all = np.clip(lam_dev.cpu().numpy(), 0.0, None)
    lam = lam_all[:k]
    if extras is not None:
        tot = float(lam_all.sum())
        extras["disc"] = float(lam_all[k:].sum() / tot) if tot > 0 else 0.0      # discarded weight of this cut
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


def recycled_truncate(psi, shape, mbd, ctx, direction, warm, ext, method=None, weight_check=None):
    """Projective truncation warm-started from the bond's previous bases (csrc/solver.cu
    pdmrg_recycle_basis; same outputs as eig_truncate plus the new spare rows, or None when it cannot be
    used).  ``warm`` is the site tensor holding the previous basis of the side that is NOT being left
    behind: moving right the right-canonical B0 (d2,k,Dr), whose rows b_i^H span the old right space
    (T = psi B0^H, P = conj(left vectors), C = P^H psi = new centre); moving left the centre-form
    A0 (d1,Dl,k) = U S (T^T = psi^H A0, P = right vectors, C^T = psi P = new centre)."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    k = min(int(mbd), rows, cols)
    p = ext.shape[0]
    K = k + p
    m = ctx.empty(rows, cols)
    D.permute4(psi, m, (Dl, d1, d2, Dr), (Dr, d2 * Dl * Dr, Dl * Dr, 1))
    mT = ctx.empty(cols, rows)
    D.permute4(m, mT, (cols, rows), (1, cols))
    right = direction == "right"
    a, c, X1, X2 = (rows, cols, m, mT) if right else (cols, rows, mT, m)
    if K > min(a, c) or ext.shape[1] != c:
        return None
    Wt = ctx.empty(K, c)
    if right:
        D.permute4(warm, Wt, (k, d2, Dr), (Dr, k * Dr, 1))          # Wt[i,(q,s)] = B0[q,i,s]
    else:
        D.permute4(warm, Wt, (k, Dl, d1), (1, k, Dl * k))           # Wt[i,(l,n)] = A0[n,l,i]
    Wt[k:].copy_(ext)
    b = min(k, window_depth(k))
    Pt, Ct, Sh, ok, used = D.recycle_basis(X1, X2, Wt, k, b, ctx.svd_ws, method=method)
    if not ok:
        return None
    EE, EEs, Sn = calc_entanglement(Sh[:k])
    if weight_check is not None:
        # captured weight = sum of the squared row norms of the new centre / ||psi||^2 (exact for any rotation of the rows)
        tot = float(torch.linalg.vector_norm(psi).item()) ** 2
        disc = max(0.0, 1.0 - float(np.dot(Sh[:k], Sh[:k])) / tot) if tot > 0 else 0.0
        if not weight_check(disc):
            return None
    A = ctx.empty(d1, Dl, k)
    B = ctx.empty(d2, k, Dr)
    if right:
        D.permute4(Pt, A, (d1, Dl, k), (1, d1, a), conj=True)       # A[n,l,i] = conj(P[i,(l,n)])
        D.permute4(Ct, B, (d2, k, Dr), (Dr, c, 1))                  # B[q,i,s] = C[i,(q,s)]
    else:
        D.permute4(Pt, B, (d2, k, Dr), (Dr, a, 1), conj=True)       # B[q,i,s] = conj(P[i,(q,s)])
        D.permute4(Ct, A, (d1, Dl, k), (1, d1, c))                  # A[n,l,i] = C[i,(l,n)]
    new_ext = ctx.empty(p, a)
    D.permute4(Pt[k:], new_ext, (p, a), (a, 1), conj=True)
    return A, B, Sn, EE, EEs, new_ext, used


def full_rank_split(psi, shape, ctx, direction):
    """Bond step that discards nothing (k = min(rows, cols), the bonds near the chain ends): any
    orthonormal basis of the full row / column space is optimal, so no spectrum is needed.  If the side
    left behind already has dimension k its isometry is the identity; otherwise one ordered Householder
    QR (pdmrg_recycle_basis with the identity as "previous basis" and no window).  The returned S are
    the row norms of the centre, NOT Schmidt values (entropies are only reported at the centre bond)."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    k = min(rows, cols)
    m = ctx.empty(rows, cols)
    D.permute4(psi, m, (Dl, d1, d2, Dr), (Dr, d2 * Dl * Dr, Dl * Dr, 1))
    right = direction == "right"
    A = ctx.empty(d1, Dl, k)
    B = ctx.empty(d2, k, Dr)
    eye = lambda n: torch.eye(n, dtype=ctx.tdtype, device=ctx.device)
    if right and k == rows:
        D.permute4(eye(rows), A, (d1, Dl, k), (k, d1 * k, 1))           # A[n,l,i] = delta((l,n), i)
        D.permute4(m, B, (d2, k, Dr), (Dr, cols, 1))                    # B[q,i,s] = m[i,(q,s)]
        S = torch.linalg.vector_norm(m, dim=1)
    elif (not right) and k == cols:
        D.permute4(eye(cols), B, (d2, k, Dr), (Dr, cols, 1))            # B[q,i,s] = delta(i, (q,s))
        D.permute4(m, A, (d1, Dl, k), (cols, d1 * cols, 1))             # A[n,l,i] = m[(l,n), i]
        S = torch.linalg.vector_norm(m, dim=0)
    else:
        mT = ctx.empty(cols, rows)
        D.permute4(m, mT, (cols, rows), (1, cols))
        a, c, X1, X2 = (rows, cols, m, mT) if right else (cols, rows, mT, m)
        Pt, Ct, Sq, ok, _ = D.recycle_basis(X1, X2, eye(c), k, 0, ctx.svd_ws, method=D.RECYCLE_HOUSEHOLDER)
        if not ok:
            return None
        S = torch.from_numpy(Sq)
        if right:
            D.permute4(Pt, A, (d1, Dl, k), (1, d1, a), conj=True)
            D.permute4(Ct, B, (d2, k, Dr), (Dr, c, 1))
        else:
            D.permute4(Pt, B, (d2, k, Dr), (Dr, a, 1), conj=True)
            D.permute4(Ct, A, (d1, Dl, k), (1, d1, c))
    Sh = S.cpu().numpy()
    if not np.all(np.isfinite(Sh)) or not np.any(Sh > 0):
        return None
    EE, EEs, Sn = calc_entanglement(Sh[:k])
    return A, B, Sn, EE, EEs


def Context_with(ctx, svd_method):
    """Shallow view of a context with another SVD back-end (shares workspaces)."""
    import copy
    c = copy.copy(ctx)
    c.svd_method = svd_method
    return c


def truncate_twosite(psi, shape, mbd, ctx, direction, recycle=None, site=None, warm=None, exact=False, stats=None):
    """Truncation used by the finite two-site sweep: projective (one heevd) for large cuts, full SVD
    (cuSOLVER gesvd) for small ones or when an explicit SVD back-end was requested.

    ``recycle`` (a dict owned by the sweep driver) switches on the warm-started variant: every exact
    step also stores the next few singular vectors of the bond, and the following visit of the same
    bond (``site``, opposite direction) starts from them and from ``warm`` = (mps[site], mps[site+1])
    instead of diagonalising the reduced density matrix again.  ``exact`` forces the eigen-decomposition
    (used at the bond whose spectrum is reported)."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    if not (ctx.svd_method < 0 and min(rows, cols) >= getattr(ctx, "projective_min", 128)):
        return svd_truncate(psi, shape, mbd, ctx, absorb=direction)
    if recycle is None:
        return eig_truncate(psi, shape, mbd, ctx, direction)
    right = direction == "right"
    k = min(int(mbd), rows, cols)
    p = spare_count(k, rows, cols)
    other = "left" if right else "right"
    ext = recycle.get((site, direction))
    shape_key = (rows, cols, k)
    if not exact and k == min(rows, cols):
        r = full_rank_split(psi, shape, ctx, direction)
        if r is not None:
            recycle.pop((site, other), None)
            if stats is not None:
                stats["n_full_rank"] = stats.get("n_full_rank", 0) + 1
            return r
    if not exact and p > 0 and ext is not None and warm is not None and recycle_allowed(recycle, site, shape_key):
        top = warm[1] if right else warm[0]
        want = (d2, k, Dr) if right else (d1, Dl, k)
        if is_dev(top) and tuple(top.shape) == want and top.dtype == ctx.tdtype and ext.dtype == ctx.tdtype \
                and tuple(ext.shape) == (p, cols if right else rows):
            # a bond where the Cholesky-QR2 variant had to be redone with Householder QR keeps using the latter
            r = recycled_truncate(psi, shape, mbd, ctx, direction, top.contiguous(), ext, method=recycle.get(("qr", site)),
                                  weight_check=lambda disc: recycled_weight_ok(recycle, site, disc))
            if r is None and stats is not None:
                stats["n_recycle_rejected"] = stats.get("n_recycle_rejected", 0) + 1
            if r is not None:
                recycle[(site, other)] = r[5]
                if r[6] == D.RECYCLE_HOUSEHOLDER:
                    recycle[("qr", site)] = D.RECYCLE_HOUSEHOLDER
                    if stats is not None:
                        stats["n_householder"] = stats.get("n_householder", 0) + 1
                if stats is not None:
                    stats["n_recycled"] = stats.get("n_recycled", 0) + 1
                return r[:5]
    extras = {}
    out = eig_truncate(psi, shape, mbd, ctx, direction, spare=p, extras=extras)
    if "ext" in extras:
        recycle[(site, other)] = extras["ext"]
        note_exact_visit(recycle, site, shape_key, extras.get("disc", 1.0))
    else:
        recycle.pop((site, other), None)
    if stats is not None:
        stats["n_exact_trunc"] = stats.get("n_exact_trunc", 0) + 1
        stats["disc_max"] = max(stats.get("disc_max", 0.0), extras.get("disc", 0.0))     # largest discarded weight seen
    return out


def rdm_truncate_twosite_states(psis, shape, mbd, ctx, direction, spectrum=True):
    """Two-site truncation for SEVERAL target states: the reference's state-averaged reduced
    density matrix (dmrg.py:56-66 / 100-110) built from the two-site wave-functions psi_s (d1,d2,Dl,Dr),

      right:  rho = 1/S sum_s m_s m_s^H,  A = top-k eigenvectors (shared by all states),  B_s = A^H m_s
      left :  rho = 1/S sum_s m_s^H m_s,  B = top-k eigenvectors^H (shared),              A_s = m_s B^H

    with m_s[(Dl,d1),(d2,Dr)] (mps_tools.py:105-107): one Hermitian eigen-decomposition, every state keeps its
    own centre tensor = the exact projection of psi_s on the kept subspace.  ``spectrum``: EE / EEs / S are
    those of the target state psi_0 alone (dmrg.py:52), truncated to k and normalised like renorm_inf
    (mps_tools.py:116-118); otherwise (bonds whose spectrum is not reported) they come from the averaged RDM.
    Returns ([A_s], [B_s], S, EE, EEs).  NumPy restatement: oracle/multistate_oracle.py."""
    d1, d2, Dl, Dr = shape
    rows, cols = Dl * d1, d2 * Dr
    k = min(int(mbd), rows, cols)
    right = direction == "right"
    nS = len(psis)
    n = rows if right else cols
    ms, mTs = [], []
    rho = ctx.empty(n, n)
    for i, psi in enumerate(psis):
        m = ctx.empty(rows, cols)
        D.permute4(psi, m, (Dl, d1, d2, Dr), (Dr, d2 * Dl * Dr, Dl * Dr, 1))
        mT = ctx.empty(cols, rows)
        D.permute4(m, mT, (cols, rows), (1, cols))
        X, kk = (m, cols) if right else (mT, rows)              # left: eigenvectors of mT mT^H are conj(v_a)
        D.gemm_nt(X, X, rho, M=n, N=n, K=kk, lda=kk, ldb=kk, ldc=n, conjB=True, alpha=1.0 / nS, beta=int(i > 0))
        ms.append(m)
        mTs.append(mT)
    if spectrum:
        _, S0, _, _ = D.svd(ms[0].clone(), ctx.svd_ws, ctx.svd_method_for(rows, cols))
        EE, EEs, Sn = calc_entanglement(S0.cpu().numpy()[:k])
    w, vec, info = D.heevd(rho, ctx.svd_ws)
    wh = w.cpu().numpy()
    if int(info.item()) != 0:
        raise ArithmeticError("cuSOLVER heevd failed (info=%d)" % int(info.item()))
    if not spectrum:
        EE, EEs, Sn = calc_entanglement(np.sqrt(np.clip(wh[::-1][:k], 0.0, None)))
    last = vec[n - 1]                                           # eigenvectors ascending: the top k are the LAST rows
    As, Bs = [], []
    if right:
        A = ctx.empty(d1, Dl, k)
        D.permute4(last, A, (d1, Dl, k), (1, d1, -n))           # A[p,l,a] = u_a[l*d1+p]
        Vc = ctx.empty(k, rows)
        D.permute4(last, Vc, (k, rows), (-n, 1), conj=True)
        for mT in mTs:
            C = D.gemm_nt(Vc, mT)                               # C[a,c] = sum_x conj(u_a[x]) m[x,c]
            B = ctx.empty(d2, k, Dr)
            D.permute4(C, B, (d2, k, Dr), (Dr, cols, 1))
            As.append(A)
            Bs.append(B)
    else:
        B = ctx.empty(d2, k, Dr)
        D.permute4(last, B, (d2, k, Dr), (Dr, -n, 1))           # B[q,a,s] = w_a[(q,s)],  w_a = conj(v_a)
        Wd = ctx.empty(k, cols)
        D.permute4(last, Wd, (k, cols), (-n, 1))
        for m in ms:
            C = D.gemm_nt(m, Wd, conjB=True)                    # C[x,a] = sum_c m[x,c] v_a[c]
            A = ctx.empty(d1, Dl, k)
            D.permute4(C, A, (d1, Dl, k), (k, d1 * k, 1))
            As.append(A)
            Bs.append(B)
    return As, Bs, Sn, EE, EEs


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
    """Schmidt spectrum of the cut from RDM eigenvalues (ascending from heevd): the reference takes the
    min(d*Dl, Dr) singular values of the (d*Dl, Dr) matrix (dmrg.py:30-37) = sqrt of the ``keep`` largest RDM
    eigenvalues (the remaining d*Dl - keep are zero up to rounding and are not part of its spectrum)."""
    lam = np.clip(w[::-1][:keep], 0.0, None)                    # descending, >= 0
    return calc_entanglement(np.sqrt(lam))


def _range_basis(X, ctx):
    """Orthonormal basis of the column space of X (nn, c) and the coefficients C = Q^H X, without a spectrum:
    returns (Pt (keep, nn) with rows conj(q_a), C (keep, c), row norms of C) or None.  For ONE state the
    reference's one-site renormalisation (RDM eigenvectors, dmrg.py:64-81) keeps Dr vectors of a rank-<=Dr
    matrix: nothing is discarded, every orthonormal basis of the range is the same isometry up to the gauge on
    the bond, so an ordered QR (pdmrg_recycle_basis with the identity as previous basis) replaces the
    (d*chi)^3 eigen-decomposition."""
    nn, c = X.shape
    if nn <= c:
        Pt = torch.eye(nn, dtype=ctx.tdtype, device=ctx.device)
        return Pt, X, torch.linalg.vector_norm(X, dim=1).cpu().numpy()
    Xt = ctx.empty(c, nn)
    D.permute4(X, Xt, (c, nn), (1, c))
    Pt, Ct, Sh, ok, _ = D.recycle_basis(X, Xt, torch.eye(c, dtype=ctx.tdtype, device=ctx.device), c, 0, ctx.svd_ws)
    if not ok or not np.any(Sh > 0):
        return None
    return Pt, Ct, Sh


def rdm_renormalize_right(psis, shape, nxts, ctx, n_avg=None, fast=False):
    """renormalizeR, dmrg.py:49-91: equal-weight RDM over ``psis`` (state averaging, :56-66),
    eigenvectors (Hermitian solver instead of the reference's general eig -- the RDM is
    Hermitian PSD), top-Dr kept; A <- eigenvectors; neighbour <- (A^dagger psi) . M[site+1] (:90).
    Returns A, [new neighbours], EE, EEs."""
    d, Dl, Dr = shape
    nn = Dl * d
    n_avg = len(psis) if n_avg is None else n_avg
    if fast and n_avg == 1:
        # single state, spectrum not needed at this site: QR instead of the RDM eigen-decomposition
        Mi = ctx.empty(nn, Dr)
        D.permute4(psis[0], Mi, (Dl, d, Dr), (Dr, Dl * Dr, 1))
        rb = _range_basis(Mi, ctx)
        if rb is not None:
            Pt, C1, Sh = rb
            keep = min(Dr, nn)
            EE, EEs, _ = calc_entanglement(Sh[:keep])             # row norms of C, NOT Schmidt values
            A = ctx.empty(d, Dl, keep)
            D.permute4(Pt, A, (d, Dl, keep), (1, d, nn), conj=True)   # A[n,kl,a] = q_a[kl*d+n] = conj(Pt[a, kl*d+n])
            new_next = []
            for nxt in nxts:
                dn, Dk, Dj = nxt.shape
                Nt = ctx.empty(dn, Dj, Dk)
                D.permute4(nxt, Nt, (dn, Dj, Dk), (Dk * Dj, 1, Dj))
                out = ctx.empty(dn, keep, Dj)
                D.gemm_nt(C1, Nt, out, M=keep, N=Dj, K=Dk, batch=dn, lda=Dk, ldb=Dk, ldc=Dj, strideA=0,
                          strideB=Dj * Dk, strideC=keep * Dj)
                new_next.append(out)
            return A, new_next, EE, EEs
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


def rdm_renormalize_left(psis, shape, prvs, ctx, n_avg=None, fast=False):
    """renormalizeL, dmrg.py:93-136.  Returns B, [new previous-site tensors], EE, EEs."""
    d, Dl, Dr = shape
    nn = d * Dr
    n_avg = len(psis) if n_avg is None else n_avg
    if fast and n_avg == 1:
        Mt = ctx.empty(nn, Dl)
        D.permute4(psis[0], Mt, (d, Dr, Dl), (Dl * Dr, 1, Dr))
        rb = _range_basis(Mt, ctx)
        if rb is not None:
            Pt, Ct, Sh = rb
            keep = min(Dl, nn)
            EE, EEs, _ = calc_entanglement(Sh[:keep])             # row norms of C, NOT Schmidt values
            B = ctx.empty(d, keep, Dr)
            D.permute4(Pt, B, (d, keep, Dr), (Dr, nn, 1), conj=True)   # B[n,a,s] = q_a[(n,s)]
            new_prev = []
            for prv in prvs:
                dp, Di, Dk = prv.shape
                out = ctx.empty(dp, Di, keep)
                D.gemm_nt(prv, Ct, out, M=dp * Di, N=keep, K=Dk, lda=Dk, ldb=Dk, ldc=keep)   # dmrg.py:135
                new_prev.append(out)
            return B, new_prev, EE, EEs
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
