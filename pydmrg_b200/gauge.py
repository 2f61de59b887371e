"""Gauge utilities on the device: moving the orthogonality centre of an MPS by SVD, right-
canonicalisation and the entanglement profile -- the same functions, argument order and layouts as
pydmrg/tools/mps_tools.py:36-96 and pydmrg/tools/aux/process_states.py:5-27, composed from the C-ABI entry points
the sweep already uses (pdmrg_svd, pdmrg_permute4 with its row scaling, the DMMA pdmrg_gemm_nt).

"Like in, like out": a list of NumPy site tensors is moved to the device, processed and returned as NumPy; a list
of CUDA tensors stays on the device.  ``mps`` is ONE state (a list of (d, Dl, Dr) tensors), as in the reference."""
import numpy as np

from . import device as D
from .core import default_context, like
from .truncate import calc_entanglement


def _svd(a, ctx):
    """Thin SVD of a device matrix with the fallback chain of truncate.svd_truncate.  Returns U, S (device), Vh."""
    rows, cols = a.shape
    first = ctx.svd_method_for(rows, cols)
    info = None
    for method in [first] + [m for m in (D.SVD_GESVDJ, D.SVD_GRAM) if m != first]:
        mm = a if method == D.SVD_GRAM else a.clone()           # gesvd / gesvdj destroy their input
        U, S, Vh, info = D.svd(mm, ctx.svd_ws, method)
        Sh = S.cpu().numpy()
        if int(info.item()) == 0 and np.all(np.isfinite(Sh)):
            return U, S, Vh, Sh
    raise ArithmeticError("cuSOLVER SVD did not converge (info=%d)" % int(info.item()))


def move_gauge_right(mps, site, returnEE=False, ctx=None):
    """tools/mps_tools.py:64-73: mps[site] (d,Dl,Dr) -> left-isometry u of its (d*Dl, Dr) matrix, and
    mps[site+1] <- diag(s) v mps[site+1]."""
    ctx = ctx or default_context()
    proto = mps[site]
    M = ctx.dev(mps[site])
    d, Dl, Dr = M.shape
    U, S, Vh, Sh = _svd(M.reshape(d * Dl, Dr), ctx)
    k = S.numel()
    nxt = ctx.dev(mps[site + 1])
    dn, Dk, Dj = nxt.shape
    SV = ctx.empty(k, Dr)
    D.permute4(Vh, SV, (k, Dr), (Dr, 1), scale=S, scale_axis=0)            # diag(s) v
    Nt = ctx.empty(dn, Dj, Dk)
    D.permute4(nxt, Nt, (dn, Dj, Dk), (Dk * Dj, 1, Dj))
    out = ctx.empty(dn, k, Dj)
    D.gemm_nt(SV, Nt, out, M=k, N=Dj, K=Dk, batch=dn, lda=Dk, ldb=Dk, ldc=Dj, strideA=0, strideB=Dj * Dk, strideC=k * Dj)
    mps[site] = like(U.reshape(d, Dl, k), proto)
    mps[site + 1] = like(out, proto)
    if returnEE:
        return mps, calc_entanglement(Sh)[0]
    return mps


def move_gauge_left(mps, site, returnEE=False, ctx=None):
    """tools/mps_tools.py:75-85: mps[site] -> right-isometry v of its (Dl, d*Dr) matrix, and
    mps[site-1] <- mps[site-1] u diag(s)."""
    ctx = ctx or default_context()
    proto = mps[site]
    M = ctx.dev(mps[site])
    d, Dl, Dr = M.shape
    a = ctx.empty(Dl, d * Dr)
    D.permute4(M, a, (Dl, d, Dr), (Dr, Dl * Dr, 1))                        # swapaxes(0,1) + reshape
    U, S, Vh, Sh = _svd(a, ctx)
    k = S.numel()
    B = ctx.empty(d, k, Dr)
    D.permute4(Vh, B, (d, k, Dr), (Dr, d * Dr, 1))
    prv = ctx.dev(mps[site - 1])
    dp, Di, Dk = prv.shape
    USt = ctx.empty(k, Dl)
    D.permute4(U, USt, (k, Dl), (1, k), scale=S, scale_axis=0)             # (u diag(s))^T
    out = ctx.empty(dp, Di, k)
    D.gemm_nt(prv, USt, out, M=dp * Di, N=k, K=Dk, lda=Dk, ldb=Dk, ldc=k)
    mps[site] = like(B, proto)
    mps[site - 1] = like(out, proto)
    if returnEE:
        return mps, calc_entanglement(Sh)[0]
    return mps


def move_gauge(mps, init_site, fin_site, ctx=None):
    """tools/mps_tools.py:87-96."""
    if init_site < fin_site:
        for site in range(init_site, fin_site):
            mps = move_gauge_right(mps, site, ctx=ctx)
    else:
        for site in range(init_site, fin_site, -1):
            mps = move_gauge_left(mps, site, ctx=ctx)
    return mps


def make_mps_right(M, ctx=None):
    """Right-canonical form by SVD from the right (tools/mps_tools.py:36-46) on the device: the prologue of a run
    started directly at a large chi spends its time here (N SVDs of chi x d*chi)."""
    for i in range(len(M) - 1, 0, -1):
        M = move_gauge_left(M, i, ctx=ctx)
    return M


def calc_entanglement_all(mpsList, gSite=0, ctx=None, orth=False):
    """Entanglement entropy of every bond and every state, tools/aux/process_states.py:5-27: the centre is moved to the
    right end and swept back, the entropy of bond (site-1, site) is read off each SVD.  ``orth``: orthonormalise the
    states' centre tensors first (mps_tools.orthonormalize_states, :14-15; host arrays).  Returns EE (nStates, N-1);
    the states are left right-canonical."""
    if orth:
        from . import mps as mps_tools
        mpsList = mps_tools.orthonormalize_states(mpsList, gSite=gSite)
    nStates, N = len(mpsList), len(mpsList[0])
    EE = np.zeros((nStates, N - 1))
    for state in range(nStates):
        mps = mpsList[state]
        for site in range(int(gSite), N - 1):
            mps = move_gauge_right(mps, site, ctx=ctx)
        for site in range(N - 1, 0, -1):
            mps, ee = move_gauge_left(mps, site, returnEE=True, ctx=ctx)
            EE[state, site - 1] = ee
        mpsList[state] = mps
    return EE
