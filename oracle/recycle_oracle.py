"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the warm-started ("recycled") two-site truncation of
``pydmrg_b200`` (csrc/solver.cu ``pdmrg_recycle_basis``, ``truncate.recycled_truncate``).

The reference truncates every bond with a fresh SVD (tools/mps_tools.py:98-120).  The product replaces that,
from a bond's second visit on, by one step of ordered subspace iteration started from the bond's previous
bases plus an exact Rayleigh-Ritz rotation restricted to a window around the cut.  This file restates that
algorithm on the CPU on top of ``oracle.dmrg_oracle`` (same local solver, same block updates), so that the
claim "recycled sweeps reach the same energies and entropies as exact-SVD sweeps" is checked without a GPU
(tests/test_oracle.py) and the CUDA implementation has a line-by-line checker.

Conventions as in dmrg_oracle: psi (d1,d2,Dl,Dr) -> m[(Dl,d1),(d2,Dr)] (mps_tools.py:105-107).
"""
import numpy as np

from . import dmrg_oracle as O

FRACTION = 8
MIN_SPARE = 4
SAFE_DISCARD = 1e-11      # below this discarded weight a bond is recycled from its second visit on, else from its third


def spare_count(k, rows, cols):
    return max(0, min(max(MIN_SPARE, k // FRACTION), min(rows, cols) - k))


def recycle_basis(X1, X2, Wt, k, b):
    """T = Wt X1^H (K x a); ordered QR of the rows of T -> Pt (rows orthonormal); Ct = Pt X2^T;
    rows [k-b, K) of Pt and Ct rotated by the Rayleigh-Ritz vectors of the window, sorted by weight.
    Returns Pt, Ct, row norms of Ct."""
    T = Wt @ X1.conj().T
    Q, _ = np.linalg.qr(T.T)                  # Householder QR keeps the column (= row of T) order
    Pt = Q.T.copy()
    Ct = Pt @ X2.T
    w0 = k - b
    if Ct.shape[0] - w0 > 1:
        Cw = Ct[w0:]
        lam, V = np.linalg.eigh(-(Cw @ Cw.conj().T))      # ascending of -H = descending weights
        Vh = V.conj().T
        Ct[w0:] = Vh @ Cw
        Pt[w0:] = Vh @ Pt[w0:]
    return Pt, Ct, np.linalg.norm(Ct, axis=1)


def matrix_of(psi, shape):
    d1, d2, Dl, Dr = shape
    return np.reshape(psi, shape).transpose(2, 0, 1, 3).reshape(Dl * d1, d2 * Dr)


def recycled_truncate(psi, shape, mbd, direction, warm, ext):
    """Same outputs as the product's truncate.recycled_truncate: A (d1,Dl,k), B (d2,k,Dr) with the centre on the
    side the sweep moves to, row norms of the centre, and the spare rows for the bond's next visit."""
    d1, d2, Dl, Dr = shape
    m = matrix_of(psi, shape)
    rows, cols = m.shape
    k = min(mbd, rows, cols)
    p = ext.shape[0]
    b = min(k, max(MIN_SPARE, k // FRACTION))
    if direction == "right":
        Wt = np.concatenate([warm.transpose(1, 0, 2).reshape(k, cols), ext], axis=0)      # rows b_i^H, then spares
        Pt, Ct, S = recycle_basis(m, m.T, Wt, k, b)
        A = Pt[:k].conj().T.reshape(Dl, d1, k).swapaxes(0, 1)                             # A = conj(P)^T
        B = Ct[:k].reshape(k, d2, Dr).swapaxes(0, 1)
    else:
        Wt = np.concatenate([warm.transpose(2, 1, 0).reshape(k, rows), ext], axis=0)      # rows (s_i u_i)^T
        Pt, Ct, S = recycle_basis(m.T, m, Wt, k, b)
        B = Pt[:k].conj().reshape(k, d2, Dr).swapaxes(0, 1)                               # B rows = conj(P)
        A = Ct[:k].T.reshape(Dl, d1, k).swapaxes(0, 1)
    return A, B, S[:k], Pt[k:].conj()


def exact_truncate(psi, shape, mbd, direction):
    """The reference's SVD truncation with the singular values pushed to the new centre, plus the spare
    singular vectors that seed the recycled step of the bond's next visit."""
    d1, d2, Dl, Dr = shape
    m = matrix_of(psi, shape)
    rows, cols = m.shape
    U, S, Vh = np.linalg.svd(m, full_matrices=False)
    k = min(mbd, S.shape[0])
    p = spare_count(k, rows, cols)
    A = U[:, :k].reshape(Dl, d1, k).swapaxes(0, 1)
    B = Vh[:k].reshape(k, d2, Dr).swapaxes(0, 1)
    if direction == "right":
        B = B * S[:k][None, :, None]
        ext = U[:, k:k + p].T.copy()                   # left vectors (non-conjugated rows): next LEFT-moving visit
    else:
        A = A * S[:k][None, None, :]
        ext = Vh[k:k + p].copy()                       # rows conj(v_a)^T: next RIGHT-moving visit
    tot = float(np.sum(S ** 2))
    disc = float(np.sum(S[k:] ** 2) / tot) if tot > 0 else 0.0
    return A, B, S[:k], ext, disc


def recycle_allowed(cache, site, shape_key):
    """Policy shared with the product (truncate.recycle_allowed): a bond may be warm-started once it has been
    decomposed exactly at its current shape either once with (numerically) nothing discarded, or twice --
    bases taken from a state that is still far from converged AND truncated leave a slowly decaying lag."""
    st = cache.get(("state", site))
    if st is None or st["shape"] != shape_key:
        return False
    return st["disc"] < SAFE_DISCARD or st["exact_visits"] >= 2


def note_exact_visit(cache, site, shape_key, disc):
    st = cache.get(("state", site))
    if st is None or st["shape"] != shape_key:
        st = {"shape": shape_key, "exact_visits": 0, "disc": 1.0}
    st["exact_visits"] += 1
    st["disc"] = disc
    cache[("state", site)] = st


def twosite_step(mps, mpo, env, site, direction, mbd, cache, exact=False, stats=None):
    A0, B0 = mps[site], mps[site + 1]
    d1, Dl, _ = A0.shape
    d2, _, Dr = B0.shape
    shape = (d1, d2, Dl, Dr)
    Ls = [F[site] for F in env]; Rs = [F[site + 2] for F in env]
    W1s = [op[site] for op in mpo]; W2s = [op[site + 1] for op in mpo]
    mv = lambda x: O.heff_twosite(x, Ls, W1s, W2s, Rs, shape, "blas")
    guess = np.einsum("ijk,lkm->iljm", A0, B0).reshape(-1)
    E, vecs, _ = O.local_eigs(mv, [guess], 1, site, len(mps), alg="davidson", check=False, stats=stats)
    psi = vecs[:, 0]
    rows, cols = Dl * d1, d2 * Dr
    k = min(mbd, rows, cols)
    p = spare_count(k, rows, cols)
    other = "left" if direction == "right" else "right"
    ext = None if cache is None else cache.get((site, direction))
    warm = B0 if direction == "right" else A0
    want = (d2, k, Dr) if direction == "right" else (d1, Dl, k)
    shape_key = (rows, cols, k)
    if (not exact and ext is not None and p > 0 and warm.shape == want
            and ext.shape == (p, cols if direction == "right" else rows) and recycle_allowed(cache, site, shape_key)):
        A, B, S, new_ext = recycled_truncate(psi, shape, mbd, direction, warm, ext)
        if stats is not None:
            stats["n_recycled"] = stats.get("n_recycled", 0) + 1
    else:
        A, B, S, new_ext, disc = exact_truncate(psi, shape, mbd, direction)
        if cache is not None:
            note_exact_visit(cache, site, shape_key, disc)
    if cache is not None:
        cache[(site, other)] = new_ext
    mps[site], mps[site + 1] = A, B
    if direction == "right":
        for m_, op in enumerate(mpo):
            env[m_][site + 1] = O.env_update_right(A, op[site], env[m_][site], None, "blas")
    else:
        for m_, op in enumerate(mpo):
            env[m_][site + 1] = O.env_update_left(B, op[site + 1], env[m_][site + 2], None, "blas")
    return E[0], S


def run_twosite(mpo, mbd, n_sweeps, recycle, seed=0, stats=None):
    """Two-site sweeps; with ``recycle`` every bond is decomposed exactly on its first visit and at the centre
    bond, warm-started otherwise.  Returns per sweep (E, EE) at the centre bond after the left half-sweep."""
    np.random.seed(seed)
    N = len(mpo[0])
    mps = O.make_mps_right(O.random_mps(N, mbd))
    env = O.build_env(mps, mpo, mbd, gauge_site=0, mode="blas")
    cache = {} if recycle else None
    out = []
    c = N // 2 - 1
    for _ in range(n_sweeps):
        for s in range(N - 1):
            twosite_step(mps, mpo, env, s, "right", mbd, cache, exact=(s == c), stats=stats)
        E = S = None
        for s in range(N - 2, -1, -1):
            e, sv = twosite_step(mps, mpo, env, s, "left", mbd, cache, exact=(s == c), stats=stats)
            if s == c:
                E, S = e, sv
        EE, _, _ = O.entanglement(S)
        out.append((complex(E), float(EE)))
    return out
