"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of PyDMRG's sweep hot path.

This file is the *checker* for the CUDA path in ``pydmrg_b200``; it is never imported by
the product package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it.

Parity pin: every function below is checked against the UNMODIFIED reference (imported
through ``oracle/ref_loader.py``) by ``tests/test_oracle_vs_reference.py`` and against the
committed fixtures in ``tests/golden/*.npz`` (generated from the reference by
``tests/golden/make_golden.py``).  The reference ships no golden vectors of its own
(SURVEY.md section 4), so those seeded reference outputs are the pins.

All ``file:line`` citations are relative to ``/root/reference/pydmrg``.

Layouts (kept identical to the reference):
  mps[site]  : (d, Dl, Dr)                     tools/mps_tools.py:216
  mpo[site]  : (wl, wr, d_out, d_in) or None   mpo/asep.py:61-66, tools/diag_tools.py:113-119
  env[k]     : (D_bra, w, D_ket), k = 0..N     tools/env_tools.py:6-26
"""
import numpy as np
import scipy.linalg as sla

def _ident_w(d=2, w=1):
    # A None MPO entry acts as the identity on the physical index and passes the MPO bond
    # straight through (tools/diag_tools.py:113-115, tools/env_tools.py:31-33, 44-45):
    # W[a,b,i,j] = delta_ab delta_ij, with w taken from the neighbouring block.
    return np.einsum("ab,ij->abij", np.eye(w), np.eye(d))


# --------------------------------------------------------------------------------------
# H1 / H2 : effective-Hamiltonian mat-vec
# --------------------------------------------------------------------------------------
def heff_onesite(x, Ls, Ws, Rs, shape, mode="blas"):
    """-H_eff x for the one-site problem.  tools/diag_tools.py:108-125.

    y[o,p,i] = - sum_mpo sum L[p,n,m] W[n,j,o,l] R[i,j,k] x[l,m,k]
    Ls[m], Rs[m]: left/right blocks (Dl,w,Dl) / (Dr,w,Dr); Ws[m]: (w,w,d,d) or None.
    ``mode='einsum'`` issues the reference's own three un-optimised einsum calls (same
    cost profile: single-threaded c_einsum); ``mode='blas'`` uses tensordot (threaded GEMM).
    """
    d, Dl, Dr = shape
    xr = np.reshape(x, shape)
    acc = np.zeros(shape, dtype=np.complex128)
    for L, W, R in zip(Ls, Ws, Rs):
        if mode == "einsum":
            t1 = np.einsum("ijk,lmk->ijlm", R, xr)                      # :117
            if W is None:
                acc += np.einsum("pnm,inom->opi", L, t1)                # :115
            else:
                t2 = np.einsum("njol,ijlm->noim", W, t1)                # :118
                acc += np.einsum("pnm,noim->opi", L, t2)                # :119
        else:
            Wm = _ident_w(d, R.shape[1]) if W is None else W
            t1 = np.tensordot(xr, R, axes=([2], [2]))                   # (l,m,i,j)
            t2 = np.tensordot(Wm, t1, axes=([1, 3], [3, 0]))            # (n,o,m,i)
            acc += np.tensordot(t2, L, axes=([0, 2], [1, 2])).transpose(0, 2, 1)  # (o,i,p)->(o,p,i)
    return -acc.reshape(-1)


def heff_twosite(x, Ls, W1s, W2s, Rs, shape, mode="blas"):
    """-H_eff x for the two-site problem.  tools/diag_tools.py:146-162.

    x: (d1,d2,Dl,Dr) flattened C-order; y[m,p,i,r] = -sum L[i,j,k] W1[j,l,m,n] W2[l,o,p,q] R[r,o,s] x[n,q,k,s].
    """
    xr = np.reshape(x, shape)
    acc = np.zeros(shape, dtype=np.complex128)
    for L, W1, W2, R in zip(Ls, W1s, W2s, Rs):
        if W1 is None:
            continue                                                    # :150-151 (unsupported -> skipped)
        if mode == "einsum":
            s1 = np.einsum("ros,nqks->ronqk", R, xr)                    # :153
            s2 = np.einsum("lopq,ronqk->lprnk", W2, s1)                 # :154
            s3 = np.einsum("jlmn,lprnk->jmprk", W1, s2)                 # :155
            acc += np.einsum("ijk,jmprk->mpir", L, s3)                  # :156
        else:
            s1 = np.tensordot(xr, R, axes=([3], [2]))                   # (n,q,k,r,o)
            s2 = np.tensordot(W2, s1, axes=([1, 3], [4, 1]))            # (l,p,n,k,r)
            s3 = np.tensordot(W1, s2, axes=([1, 3], [0, 2]))            # (j,m,p,k,r)
            acc += np.tensordot(L, s3, axes=([1, 2], [0, 3])).transpose(1, 2, 0, 3)  # (i,m,p,r)->(m,p,i,r)
    return -acc.reshape(-1)


def heff_dense_onesite(Ls, Ws, Rs, shape):
    """Dense one-site H (NOT negated).  tools/diag_tools.py:30-41."""
    d, Dl, Dr = shape
    dim = d * Dl * Dr
    H = np.zeros((dim, dim), dtype=np.complex128)
    for L, W, R in zip(Ls, Ws, Rs):
        Wm = _ident_w(d, R.shape[1]) if W is None else W
        t = np.einsum("lmin,kmq->linkq", Wm, R)
        H += np.einsum("jlp,linkq->ijknpq", L, t).reshape(dim, dim)
    return H


def heff_dense_twosite(Ls, W1s, W2s, Rs, shape):
    """Dense two-site H (NOT negated).  tools/diag_tools.py:43-63."""
    d1, d2, Dl, Dr = shape
    dim = d1 * d2 * Dl * Dr
    H = np.zeros((dim, dim), dtype=np.complex128)
    for L, W1, W2, R in zip(Ls, W1s, W2s, Rs):
        W1m = _ident_w(d1) if W1 is None else W1
        W2m = _ident_w(d2) if W2 is None else W2
        t1 = np.einsum("lopq,ros->lpqrs", W2m, R)
        t2 = np.einsum("jlmn,lpqrs->jmnpqrs", W1m, t1)
        H += np.einsum("ijk,jmnpqrs->mpirnqks", L, t2).reshape(dim, dim)
    return H


# --------------------------------------------------------------------------------------
# E1 / E2 / E0 : environment (block) tensors
# --------------------------------------------------------------------------------------
def env_update_right(A, W, L, Abra=None, mode="blas"):
    """Grow the LEFT block over one site (the reference calls this update_envR).
    tools/env_tools.py:40-50.   F'[k,m,q] = sum L[j,l,p] Abra[i,j,k] W[l,m,i,n] A[n,p,q];
    Abra defaults to conj(A) (:41)."""
    if Abra is None:
        Abra = np.conj(A)
    d = A.shape[0]
    if mode == "einsum":
        t1 = np.einsum("jlp,ijk->lpik", L, Abra)                        # :47
        if W is None:
            return np.einsum("npq,mpnk->kmq", A, t1)                    # :45
        t2 = np.einsum("lmin,lpik->mpnk", W, t1)                        # :48
        return np.einsum("npq,mpnk->kmq", A, t2)                        # :49
    Wm = _ident_w(d, L.shape[1]) if W is None else W
    t1 = np.tensordot(L, Abra, axes=([0], [1]))                         # (l,p,i,k)
    t2 = np.tensordot(Wm, t1, axes=([0, 2], [0, 2]))                    # (m,n,p,k)
    return np.tensordot(t2, A, axes=([1, 2], [0, 1])).transpose(1, 0, 2)  # (m,k,q)->(k,m,q)


def env_update_left(B, W, R, Bbra=None, mode="blas"):
    """Grow the RIGHT block over one site (reference: update_envL).
    tools/env_tools.py:28-38.   F'[x,y,a] = sum B[e,a,f] R[c,d,f] W[y,d,b,e] Bbra[b,x,c]."""
    if Bbra is None:
        Bbra = np.conj(B)
    d = B.shape[0]
    if mode == "einsum":
        t1 = np.einsum("eaf,cdf->eacd", B, R)                           # :35
        if W is None:
            return np.einsum("bacy,bxc->xya", t1, Bbra)                 # :33
        t2 = np.einsum("eacd,ydbe->acyb", t1, W)                        # :36
        return np.einsum("acyb,bxc->xya", t2, Bbra)                     # :37
    Wm = _ident_w(d, R.shape[1]) if W is None else W
    t1 = np.tensordot(B, R, axes=([2], [2]))                            # (e,a,c,d)
    t2 = np.tensordot(t1, Wm, axes=([0, 3], [3, 1]))                    # (a,c,y,b)
    return np.tensordot(Bbra, t2, axes=([0, 2], [3, 1])).transpose(0, 2, 1)  # (x,a,y)->(x,y,a)


def alloc_env(mps, mpo, mbd):
    """Zero-filled environment lists with the reference's shapes.  tools/env_tools.py:6-26."""
    N = len(mps)
    envs = []
    for op in mpo:
        w = 1 if op[0] is None else op[0].shape[1]
        F = [np.array([[[1.0]]])]
        for site in range(N // 2):
            D = min(2 ** (site + 1), mbd)
            F.append(np.zeros((D, w, D)))
        if N % 2 == 1:
            D = min(2 ** (N // 2 + 1), mbd)
            F.append(np.zeros((D, w, D)))
        for site in range(N // 2 - 1, 0, -1):
            D = min(2 ** site, mbd)
            F.append(np.zeros((D, w, D)))
        F.append(np.array([[[1.0]]]))
        envs.append(F)
    return envs


def build_env(mps, mpo, mbd, mps_bra=None, gauge_site=0, mode="blas"):
    """Full environment build around ``gauge_site``.  tools/env_tools.py:79-89."""
    N = len(mps)
    envs = alloc_env(mps, mpo, mbd)
    for m, op in enumerate(mpo):
        for site in range(N - 1, gauge_site, -1):
            bra = None if mps_bra is None else mps_bra[site]
            envs[m][site] = env_update_left(mps[site], op[site], envs[m][site + 1], bra, mode)
        for site in range(gauge_site):
            bra = None if mps_bra is None else mps_bra[site]
            envs[m][site + 1] = env_update_right(mps[site], op[site], envs[m][site], bra, mode)
    return envs


def full_contract(mpo, mps, lmps=None, mode="blas"):
    """<l|O|r> for single states (lists of site tensors).  tools/contract.py:5-44.
    ``lmps`` holds the BRA tensors as used (default conj(mps), contract.py:12-13);
    mpo=None contracts the overlap (contract.py:33-34)."""
    N = len(mps)
    if lmps is None:
        lmps = [np.conj(M) for M in mps]
    if mpo is None:
        mpo = [[None] * N]
    total = 0
    for op in mpo:
        R = np.array([[[1.0]]])
        for site in range(N - 1, -1, -1):
            R = env_update_left(mps[site], op[site], R, lmps[site], mode)
        total = total + R[0, 0, 0]
    return total


# --------------------------------------------------------------------------------------
# K : restarted non-symmetric Davidson
# --------------------------------------------------------------------------------------
def pick_lowest_real(w, v, nroots):
    """Sort Ritz pairs by ascending real part.  tools/diag_tools.py:175-179."""
    idx = np.argsort(np.real(w))
    return w[idx], v[:, idx], idx


def _gram_schmidt_list(vs):
    """Classical Gram-Schmidt over a list.  tools/la_tools.py:586-598."""
    out = [vs[0] / np.sqrt(np.real(np.vdot(vs[0], vs[0])))]
    for v in vs[1:]:
        vi = v.astype(np.complex128).copy()
        for u in out:
            vi -= u * np.vdot(u, vi)
        out.append(vi / np.sqrt(np.real(np.vdot(vi, vi))))
    return out


def davidson_nosym(matvec, x0, nroots=1, max_space=12, lindep=1e-14, tol=1e-16,
                   max_cycle=1000, stats=None):
    """Restatement of davidson_nosym1, tools/la_tools.py:402-557, with the arguments the
    DMRG driver passes (tools/diag_tools.py:260: identity preconditioner, pick_eigs,
    follow_state=False, tol=1e-16, max_cycle=1000).

    Notes that matter for parity:
      * the convergence flags are never set (:485-488), so the loop ends when every new
        trial vector is pruned by ``lindep`` (||r||^2 <= 1e-14, :505-541) or max_cycle;
      * max_space is widened to max_space + 3*nroots (:412);
      * restart ("fresh start") from the current Ritz vectors when
        space + nroots > max_space (:544), then orth(x0) (:428).
    Returns (eigvals[:nroots], list of Ritz vectors).  ``stats`` (dict) receives the
    number of mat-vecs and cycles.
    """
    if not isinstance(x0, list):
        x0 = [x0]
    max_space = max_space + 3 * nroots
    heff = None
    fresh = True
    n_mv = 0
    e = None
    xs, ax, space, xt = [], [], 0, None
    cyc = 0
    for cyc in range(max_cycle):
        if fresh:
            xs, ax, space = [], [], 0
            xt = _gram_schmidt_list(x0)                                  # :428
        elif len(xt) > 1:
            xt = _gram_schmidt_list(xt)[:40]                             # :431-432
        axt = [matvec(v) for v in xt]                                    # :433
        n_mv += len(xt)
        xs.extend(xt)
        ax.extend(axt)
        rnow = len(xt)
        head, space = space, space + rnow
        if heff is None:
            heff = np.zeros((max_space + nroots, max_space + nroots), dtype=np.complex128)
        for i in range(rnow):                                            # :449-457
            for k in range(rnow):
                heff[head + k, head + i] = np.vdot(xt[k], axt[i])
        for i in range(head):
            for k in range(rnow):
                heff[head + k, i] = np.vdot(xt[k], ax[i])
                heff[i, head + k] = np.vdot(xs[i], axt[k])
        w, v = sla.eig(heff[:space, :space])                             # :461
        e, v, _ = pick_lowest_real(w, v, nroots)
        e = e[:nroots]
        v = v[:, :nroots]
        nr = e.shape[0]
        x0 = [sum(v[i, k] * xs[i] for i in reversed(range(space))) for k in range(nr)]   # :469, :615-632
        ax0 = [sum(v[i, k] * ax[i] for i in reversed(range(space))) for k in range(nr)]  # :470
        xt = [ax0[k] - e[k] * x0[k] for k in range(nr)]                  # :479-482
        dx_norm = [np.sqrt(np.real(np.vdot(r, r))) for r in xt]
        keep = []
        for k in range(nr):                                              # :505-522 (conv[k] is always False)
            if dx_norm[k] ** 2 > lindep:
                r = xt[k]                                                # identity preconditioner, diag_tools.py:137-139
                keep.append(r * (1.0 / np.sqrt(np.real(np.vdot(r, r)))))
        xt = keep
        for i in range(space):                                           # :524-527
            for r in xt:
                r -= xs[i] * np.vdot(xs[i], r)
        keep = []
        for r in xt:                                                     # :528-537
            nrm = np.sqrt(np.real(np.vdot(r, r)))
            if nrm ** 2 > lindep:
                keep.append(r * (1.0 / nrm))
        xt = keep
        if len(xt) == 0:                                                 # :539-541
            break
        fresh = space + nroots > max_space                               # :544
    if stats is not None:
        stats["n_matvec"] = n_mv
        stats["cycles"] = cyc + 1
    return e, x0


# --------------------------------------------------------------------------------------
# D : local eigenproblem driver
# --------------------------------------------------------------------------------------
def check_overlap(guess, vecs, E, preserve_state=False):
    """State-following logic.  tools/diag_tools.py:71-104 (printStates branch omitted)."""
    if vecs.ndim == 1:
        vecs = vecs.reshape(-1, 1)
    nv = vecs.shape[1]
    matched = False
    for j in range(nv):
        if np.abs(np.dot(guess, np.conj(vecs[:, j]))) > 0.98:
            matched = True
            if j != 0 and preserve_state:
                tmp = vecs[:, j].copy()
                vecs[:, j] = vecs[:, 0]
                vecs[:, 0] = tmp
                E[j], E[0] = E[0], E[j]
    if preserve_state and not matched:
        vecs = np.asarray(guess).reshape(-1, 1)                          # :93-97 keep the old guess
    return E, vecs, np.abs(np.dot(guess, np.conj(vecs[:, 0])))


def local_eigs(matvec, guesses, n_states, site, n_sites, alg="davidson", dense_H=None,
               preserve_state=False, edge_preserve_state=True, check=True, stats=None):
    """calc_eigs (tools/diag_tools.py:292-313) for 'davidson' (:243-290) and 'exact'
    (:193-212).  ``matvec`` is -H x; ``dense_H`` (un-negated) is needed for 'exact'.
    Returns (E, vecs (n, k), ovlp)."""
    n = guesses[0].size
    if alg == "davidson":
        k = min(n_states, n - 1)
        vals, vecso = davidson_nosym(matvec, [g.astype(np.complex128) for g in guesses[:k]],
                                     nroots=k, stats=stats)
        order = np.argsort(np.real(vals))
        E = -vals[order[:k]]
        vecs = np.zeros((n, k), dtype=np.complex128)
        for i in range(min(k, len(order))):
            vecs[:, i] = vecso[order[i]]
    elif alg == "exact":
        vals, vv = sla.eig(dense_H)
        order = np.argsort(vals)[::-1]
        E = vals[order[:n_states]]
        vecs = vv[:, order[:n_states]]
    else:
        raise ValueError(alg)
    if ((site == 0) or (site == n_sites - 1)) and edge_preserve_state:
        preserve_state = True
    if check:
        E, vecs, ovlp = check_overlap(guesses[0], vecs, E, preserve_state)
    else:
        ovlp = None
    return E, vecs, ovlp


# --------------------------------------------------------------------------------------
# T2 / T1 : truncation
# --------------------------------------------------------------------------------------
def entanglement(S):
    """tools/mps_tools.py:9-15.  Returns (EE, EEspec, S_normalised)."""
    S = S / np.sqrt(np.dot(S, np.conj(S)))
    p = S * np.conj(S)
    with np.errstate(divide="ignore", invalid="ignore"):
        spec = -p * np.log2(p)
    return np.sum(spec), spec, S


def svd_truncate_twosite(psi, shape, mbd):
    """renorm_inf, tools/mps_tools.py:98-120: psi (d1,d2,Dl,Dr) -> A (d1,Dl,k), B (d2,k,Dr),
    S truncated to mbd and L2-normalised; S is NOT absorbed into A or B."""
    d1, d2, Dl, Dr = shape
    m = np.reshape(psi, shape).transpose(2, 0, 1, 3).reshape(Dl * d1, d2 * Dr)   # :105-107
    # NB the reference reshapes to (n3*n1, n4*n2) (:107) -- same element count; rows are (Dl,d1),
    # columns (d2,Dr) in memory order.
    U, S, Vh = np.linalg.svd(m, full_matrices=False)
    k = min(mbd, S.shape[0])
    A = U.reshape(Dl, d1, -1)[:, :, :k].swapaxes(0, 1)                  # :110-112
    B = Vh.reshape(-1, d2, Dr)[:k].swapaxes(0, 1)                       # :113-115
    EE, EEs, Sn = entanglement(S[:k])
    return A, B, Sn, EE, EEs


def _rdm(M, direction):
    """dmrg.py:18-28."""
    d, Dl, Dr = M.shape
    if direction == "right":
        m = M.swapaxes(0, 1).reshape(Dl * d, Dr)
        return m @ m.conj().T
    m = M.swapaxes(0, 1).reshape(Dl, d * Dr)
    return m.T @ m.conj()


def rdm_renormalize_right(psi, shape, nxt):
    """renormalizeR for one state, dmrg.py:49-91.  Returns A (d,Dl,Dr), new neighbour, EE, EEs."""
    d, Dl, Dr = shape
    v = np.reshape(psi, shape)
    S = np.linalg.svd(v.reshape(d * Dl, Dr), compute_uv=False)           # dmrg.py:30-37 (entropy only)
    EE, EEs, _ = entanglement(S)
    vals, vecs = np.linalg.eig(_rdm(v, "right"))                         # :68
    inds = np.argsort(vals)[::-1][:Dr]
    vecs = sla.orth(vecs[:, inds])                                       # :76
    A = vecs.reshape(Dl, d, Dr).swapaxes(0, 1)                           # :80-81
    nxt_new = np.einsum("lmn,lmk,ikj->inj", np.conj(A), v, nxt)          # :90
    return A, nxt_new, EE, EEs


def rdm_renormalize_left(psi, shape, prv):
    """renormalizeL for one state, dmrg.py:93-136."""
    d, Dl, Dr = shape
    v = np.reshape(psi, shape)
    S = np.linalg.svd(v.swapaxes(0, 1).reshape(Dl, d * Dr), compute_uv=False)   # dmrg.py:39-47
    EE, EEs, _ = entanglement(S)
    vals, vecs = np.linalg.eig(_rdm(v, "left"))                          # :112
    inds = np.argsort(vals)[::-1][:Dl]
    vecs = sla.orth(vecs[:, inds]).T                                     # :120-121
    B = vecs.reshape(Dl, d, Dr).swapaxes(0, 1)                           # :125-126
    prv_new = np.einsum("ijk,lkm,lnm->ijn", prv, v, np.conj(B))          # :135
    return B, prv_new, EE, EEs


# --------------------------------------------------------------------------------------
# S : steps and drivers (single target state)
# --------------------------------------------------------------------------------------
def make_mps_right(mps):
    """Right-canonicalise in place by SVD from the right.  tools/mps_tools.py:36-46."""
    N = len(mps)
    for i in range(N - 1, 0, -1):
        m = mps[i].swapaxes(0, 1)
        n1, n2, n3 = m.shape
        U, s, V = np.linalg.svd(m.reshape(n1, n2 * n3), full_matrices=False)
        mps[i] = V.reshape(n1, n2, n3).swapaxes(0, 1)
        mps[i - 1] = np.einsum("klj,ji,i->kli", mps[i - 1], U, s)
    return mps


def random_mps(N, mbd, d=2, const=1e-2):
    """create_rand_mps, tools/mps_tools.py:211-221 (draw order identical so that
    np.random.seed(k) gives the same initial state as the reference)."""
    M = []
    for i in range(N // 2):
        M.append(const * np.random.rand(d, min(d ** i, mbd), min(d ** (i + 1), mbd)))
    if N % 2 == 1:
        i = N // 2 - 1
        M.append(const * np.random.rand(d, min(d ** (i + 1), mbd), min(d ** (i + 1), mbd)))
    for i in range(N // 2 - 1, -1, -1):
        M.append(const * np.random.rand(d, min(d ** (i + 1), mbd), min(d ** i, mbd)))
    return M


def onesite_step(mps, mpo, env, site, direction, alg="davidson", mode="blas", stats=None):
    """rightStep / leftStep for nStates=1.  dmrg.py:138-148, 174-184."""
    N = len(mps)
    shape = mps[site].shape
    Ls = [F[site] for F in env]
    Rs = [F[site + 1] for F in env]
    Ws = [op[site] for op in mpo]
    mv = lambda x: heff_onesite(x, Ls, Ws, Rs, shape, mode)
    dense = heff_dense_onesite(Ls, Ws, Rs, shape) if alg == "exact" else None
    E, vecs, _ = local_eigs(mv, [mps[site].reshape(-1)], 1, site, N, alg=alg, dense_H=dense, stats=stats)
    psi = vecs[:, 0]
    if direction == "right":
        A, nxt, EE, EEs = rdm_renormalize_right(psi, shape, mps[site + 1])
        mps[site], mps[site + 1] = A, nxt
        for m, op in enumerate(mpo):
            env[m][site + 1] = env_update_right(A, op[site], env[m][site], None, mode)
    else:
        B, prv, EE, EEs = rdm_renormalize_left(psi, shape, mps[site - 1])
        mps[site], mps[site - 1] = B, prv
        for m, op in enumerate(mpo):
            env[m][site] = env_update_left(B, op[site], env[m][site + 1], None, mode)
    return E[0], EE, EEs


def increase_mbd(mps, mbd, d=2):
    """Zero-pad a state to the bond profile of a larger chi.  tools/mps_tools.py:244-279 (open chain)."""
    N = len(mps)
    out = []
    for site, M in enumerate(mps):
        if site < N // 2:                                                # :248-252
            sz1, sz2 = min(d ** site, mbd), min(d ** (site + 1), mbd)
        elif N % 2 == 1 and site == N // 2:                              # :253-258
            sz1 = sz2 = min(d ** site, mbd)
        else:                                                            # :259-264
            sz1, sz2 = min(d ** (N - site), mbd), min(d ** (N - site - 1), mbd)
        _, ny, nz = M.shape
        out.append(np.pad(M.astype(np.complex128), ((0, 0), (0, sz1 - ny), (0, sz2 - nz))))
    return out


def run_dmrg_onesite(mpo, mbd, tol=1e-5, max_iter=10, min_iter=0, alg="davidson",
                     gauge_site_save=None, mps=None, mode="blas", stats=None, gauge_site_load=0):
    """run_dmrg for a single chi, nStates=1, fresh random start.  dmrg.py:309-508 with
    run_sweeps :238-307, rightSweep/leftSweep :150-208, checkConv :210-223.
    Call np.random.seed(k) first to reproduce the reference's initial MPS.  ``mps`` + ``gauge_site_load``: continue
    from a saved state whose centre sits on that site (dmrg.py:352-358, 391, run_sweeps :248-259)."""
    N = len(mpo[0])
    if gauge_site_save is None:
        gauge_site_save = N // 2 + 1
    if mps is None:
        mps = make_mps_right(random_mps(N, mbd))
    env = build_env(mps, mpo, mbd, gauge_site=gauge_site_load, mode=mode)
    it, E_prev, cont = 0, 0, True
    E = EE = EEs = None
    if gauge_site_load != 0:                                             # dmrg.py:248-259 (a loaded state)
        for site in range(gauge_site_load, N - 1):
            onesite_step(mps, mpo, env, site, "right", alg, mode, stats)
        for site in range(N - 1, 0, -1):
            onesite_step(mps, mpo, env, site, "left", alg, mode, stats)
    while cont:
        for site in range(0, N - 1):
            e, ee, ees = onesite_step(mps, mpo, env, site, "right", alg, mode, stats)
            if site == N // 2:
                E, EE, EEs = e, ee, ees
        for site in range(N - 1, 0, -1):
            e, ee, ees = onesite_step(mps, mpo, env, site, "left", alg, mode, stats)
            if site == N // 2:
                E, EE, EEs = e, ee, ees
        if abs(E - E_prev) < tol and it > min_iter:
            cont = False
        elif it > max_iter - 1:
            cont = False
        else:
            it += 1
            E_prev = E
    if gauge_site_save != 0:
        _E = None
        for site in range(0, gauge_site_save):
            e, ee, ees = onesite_step(mps, mpo, env, site, "right", alg, mode, stats)
            if site == N // 2:
                _E, _EE, _EEs = e, ee, ees
        site = gauge_site_save
        shape = mps[site].shape
        Ls = [F[site] for F in env]; Rs = [F[site + 1] for F in env]; Ws = [op[site] for op in mpo]
        mv = lambda x: heff_onesite(x, Ls, Ws, Rs, shape, mode)
        dense = heff_dense_onesite(Ls, Ws, Rs, shape) if alg == "exact" else None
        _, vecs, _ = local_eigs(mv, [mps[site].reshape(-1)], 1, site, N, alg=alg, dense_H=dense, stats=stats)
        mps[site] = vecs[:, 0].reshape(shape)
        if _E is not None:
            E, EE, EEs = _E, _EE, _EEs
    return E, EE, EEs, mps, env


def twosite_step(mps, mpo, env, site, direction, mbd, alg="davidson", mode="blas", stats=None):
    """Finite two-site bond step on sites (site, site+1): the composition the north-star
    names of the reference's own pieces --
      two-site Hfun  tools/diag_tools.py:146-162  (guess: einsum('ijk,lkm->iljm'), :257)
      SVD truncation tools/mps_tools.py:98-120    (renorm_inf)
      block update   tools/env_tools.py:40-50 / 28-38.
    The singular values are pushed into the site that becomes the new centre
    (right: B <- diag(S) B, left: A <- A diag(S)) so that the next bond's guess is the
    current wave-function (renorm_inf drops S because iDMRG re-pads its guess instead)."""
    A0, B0 = mps[site], mps[site + 1]
    d1, Dl, _ = A0.shape
    d2, _, Dr = B0.shape
    shape = (d1, d2, Dl, Dr)
    Ls = [F[site] for F in env]
    Rs = [F[site + 2] for F in env]
    W1s = [op[site] for op in mpo]
    W2s = [op[site + 1] for op in mpo]
    mv = lambda x: heff_twosite(x, Ls, W1s, W2s, Rs, shape, mode)
    guess = np.einsum("ijk,lkm->iljm", A0, B0).reshape(-1)
    dense = heff_dense_twosite(Ls, W1s, W2s, Rs, shape) if alg == "exact" else None
    E, vecs, _ = local_eigs(mv, [guess], 1, site, len(mps), alg=alg, dense_H=dense, check=False, stats=stats)
    psi = vecs[:, 0]
    A, B, S, EE, EEs = svd_truncate_twosite(psi, shape, mbd)
    nrm = np.sqrt(np.real(np.vdot(psi, psi)))
    if direction == "right":
        mps[site] = A
        mps[site + 1] = B * (S * nrm)[None, :, None]
        for m, op in enumerate(mpo):
            env[m][site + 1] = env_update_right(A, op[site], env[m][site], None, mode)
    else:
        mps[site] = A * (S * nrm)[None, None, :]
        mps[site + 1] = B
        for m, op in enumerate(mpo):
            env[m][site + 1] = env_update_left(B, op[site + 1], env[m][site + 2], None, mode)
    return E[0], EE, EEs, S


def run_dmrg_twosite(mpo, mbd, n_sweeps=2, alg="davidson", mps=None, mode="blas", stats=None):
    """Finite two-site DMRG: right sweep over bonds 0..N-2 then left sweep N-2..0,
    ``n_sweeps`` times; E / EE recorded at the centre bond (site N//2-1, N//2)."""
    N = len(mpo[0])
    if mps is None:
        mps = make_mps_right(random_mps(N, mbd))
    env = build_env(mps, mpo, mbd, gauge_site=0, mode=mode)
    # two-site bonds can grow beyond the one-site allocation; environments are replaced
    E = EE = EEs = S = None
    hist = []
    for sw in range(n_sweeps):
        for site in range(0, N - 1):
            e, ee, ees, s = twosite_step(mps, mpo, env, site, "right", mbd, alg, mode, stats)
            if site == N // 2 - 1:
                E, EE, EEs, S = e, ee, ees, s
        for site in range(N - 2, -1, -1):
            e, ee, ees, s = twosite_step(mps, mpo, env, site, "left", mbd, alg, mode, stats)
            if site == N // 2 - 1:
                E, EE, EEs, S = e, ee, ees, s
        hist.append(E)
    return E, EE, EEs, S, mps, env, hist


def idmrg_iter(mps, mpo2, env, mbd, alg="davidson", mode="blas", stats=None):
    """single_iter, idmrg.py:44-54 (nStates=1): two-site solve on the 2-slot env, SVD
    truncation, both block updates (tools/env_tools.py:52-68), bond-growth zero padding
    (idmrg.py:36-42)."""
    A0, B0 = mps
    d1, Dl, _ = A0.shape
    d2, _, Dr = B0.shape
    shape = (d1, d2, Dl, Dr)
    Ls = [F[0] for F in env]; Rs = [F[1] for F in env]
    W1s = [op[0] for op in mpo2]; W2s = [op[1] for op in mpo2]
    mv = lambda x: heff_twosite(x, Ls, W1s, W2s, Rs, shape, mode)
    guess = np.einsum("ijk,lkm->iljm", A0, B0).reshape(-1)
    dense = heff_dense_twosite(Ls, W1s, W2s, Rs, shape) if alg == "exact" else None
    E, vecs, _ = local_eigs(mv, [guess], 1, 0, 2, alg=alg, dense_H=dense,
                            edge_preserve_state=False, check=(alg != "exact"), stats=stats)
    A, B, S, EE, EEs = svd_truncate_twosite(vecs[:, 0], shape, mbd)
    for m, op in enumerate(mpo2):
        newR = env_update_left(B, op[1], env[m][1], None, mode)
        newL = env_update_right(A, op[0], env[m][0], None, mode)
        env[m][0], env[m][1] = newL, newR
    d = d1
    n1, n2, n3 = A.shape
    A = np.pad(A, ((0, 0), (0, min(mbd, n3) - n2), (0, min(mbd, n3 * d) - n3)))
    B = np.pad(B, ((0, 0), (0, min(mbd, n3 * d) - n3), (0, min(mbd, n3) - n2)))
    return E[0], EE, EEs, S, [A, B], env


def run_idmrg(mpo4, mbd, tol=1e-2, max_iter=1000, min_iter=10, alg="davidson", mode="blas", stats=None):
    """run_idmrg / run_iters for one chi, nStates=1.  idmrg.py:84-131, 133-296.
    ``mpo4`` is the reference's 4-site MPO: edge = sites (0,3), bulk = sites (1,2) (:16-34)."""
    edge = [[op[0], op[3]] for op in mpo4]
    bulk = [[op[1], op[2]] for op in mpo4]
    mps = random_mps(2, mbd)
    env = [[np.ones((1, 1, 1), dtype=np.complex128), np.ones((1, 1, 1), dtype=np.complex128)] for _ in mpo4]
    E, EE, EEs, S, mps, env = idmrg_iter(mps, edge, env, mbd, alg, mode, stats)
    N = 4
    it, E_prev, cont = 0, E, True
    while cont:
        E, EE, EEs, S, mps, env = idmrg_iter(mps, bulk, env, mbd, alg, mode, stats)
        if abs(E / N - E_prev) < tol and it > min_iter:
            cont = False
        elif it > max_iter - 1:
            cont = False
        else:
            it += 1
            E_prev = E / N
        N += 2
    return E, EE, EEs, S, N - 2, mps, env


# --------------------------------------------------------------------------------------
# Model MPOs in the reference's layout (data builders used by tests / bench only)
# --------------------------------------------------------------------------------------
def heisenberg_mpo(N, J=1.0, Jz=1.0, h=0.0, sign=-1.0):
    """XXZ chain H = sum J/2 (S+S- + S-S+) + Jz SzSz - h Sz in the reference's MPO layout
    (wl, wr, d_out, d_in), w = 5, same corner convention as mpo/asep.py:45-66 (last row /
    first column carry the running operator).  The reference has no spin model;
    because the reference maximises Re(eig) the ground state is obtained from ``sign=-1``
    (MPO of -H)."""
    Sp = np.array([[0., 1.], [0., 0.]]); Sm = Sp.T.copy()
    Sz = np.array([[0.5, 0.], [0., -0.5]]); I = np.eye(2); z = np.zeros((2, 2))
    W = np.array([[I, z, z, z, z],
                  [Sm, z, z, z, z],
                  [Sp, z, z, z, z],
                  [Sz, z, z, z, z],
                  [-h * sign * Sz, sign * J / 2 * Sp, sign * J / 2 * Sm, sign * Jz * Sz, I]])
    mpo = []
    for site in range(N):
        if site == 0:
            mpo.append(W[-1:, :].copy())
        elif site == N - 1:
            mpo.append(W[:, :1].copy())
        else:
            mpo.append(W.copy())
    return [mpo]
