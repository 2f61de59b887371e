"""Environment (block) tensors on the device -- same function names, argument order and list
layouts as pydmrg/tools/env_tools.py.  ``F[mpoInd][k]`` has shape (D_bra, w, D_ket);
``F[.][site]`` is the left block of ``site`` and ``F[.][site+1]`` its right block."""
import numpy as np

from . import device as D
from .core import default_context, is_dev, like


def alloc_env(M, W, mbd, ctx=None):
    """Shapes of tools/env_tools.py:6-26 (zero blocks, unit boundary blocks)."""
    N = len(M)
    dev_out = is_dev(M[0])
    ctx = ctx or default_context()

    def mk(a):
        return ctx.dev(a) if dev_out else a.astype(ctx.npdtype)

    env_lst = []
    for op in W:
        w = 1 if op[0] is None else op[0].shape[1]
        F = [mk(np.ones((1, 1, 1)))]
        for site in range(N // 2):
            Dm = min(2 ** (site + 1), mbd)
            F.append(mk(np.zeros((Dm, w, Dm))))
        if N % 2 == 1:
            Dm = min(2 ** (N // 2 + 1), mbd)
            F.append(mk(np.zeros((Dm, w, Dm))))
        for site in range(N // 2 - 1, 0, -1):
            Dm = min(2 ** site, mbd)
            F.append(mk(np.zeros((Dm, w, Dm))))
        F.append(mk(np.ones((1, 1, 1))))
        env_lst.append(F)
    return env_lst


def update_envR(M, W, F, site, Ml=None, ctx=None):
    """Grow the left block over ``site``: F[.][site+1] <- (F[.][site], M[site], W[.][site]).
    tools/env_tools.py:40-50.  ``Ml`` = bra tensors (default conj(M), :41 -- conjugated inside
    the pack kernel instead of deep-copying the whole MPS as the reference does)."""
    ctx = ctx or default_context()
    A = ctx.dev(M[site])
    Abra = None if Ml is None else ctx.dev(Ml[site])
    for m in range(len(W)):
        out = D.env_update_right(A, W[m][site], ctx.dev(F[m][site]), Abra, ctx.ws)
        F[m][site + 1] = like(out, M[site])
    return F


def update_envL(M, W, F, site, Ml=None, ctx=None):
    """Grow the right block over ``site``: F[.][site] <- (F[.][site+1], M[site], W[.][site]).
    tools/env_tools.py:28-38."""
    ctx = ctx or default_context()
    B = ctx.dev(M[site])
    Bbra = None if Ml is None else ctx.dev(Ml[site])
    for m in range(len(W)):
        out = D.env_update_left(B, W[m][site], ctx.dev(F[m][site + 1]), Bbra, ctx.ws)
        F[m][site] = like(out, M[site])
    return F


def calc_env(M, W, mbd, Ml=None, gaugeSite=0, ctx=None):
    """Full environment build around ``gaugeSite``.  tools/env_tools.py:79-89."""
    N = len(M)
    env_lst = alloc_env(M, W, mbd, ctx)
    for site in range(N - 1, gaugeSite, -1):
        env_lst = update_envL(M, W, env_lst, site, Ml=Ml, ctx=ctx)
    for site in range(gaugeSite):
        env_lst = update_envR(M, W, env_lst, site, Ml=Ml, ctx=ctx)
    return env_lst


def calc_env_inf(M, W, mbd, Ml=None, gaugeSite=0, ctx=None):
    """Two-slot [L, R] environment of iDMRG, both 1x1x1 ones.  tools/env_tools.py:70-77."""
    ctx = ctx or default_context()
    dev_out = is_dev(M[0][0]) if isinstance(M[0], (list, tuple)) else is_dev(M[0])
    one = np.ones((1, 1, 1), dtype=ctx.npdtype)
    return [[ctx.dev(one) if dev_out else one.copy(), ctx.dev(one) if dev_out else one.copy()] for _ in W]


def update_env_inf(mps, mpo, env, mpsl=None, ctx=None):
    """iDMRG: absorb B into the right block and A into the left block.
    tools/env_tools.py:52-68 (the reference juggles a temporary centre slot; net effect below)."""
    ctx = ctx or default_context()
    A, B = ctx.dev(mps[0]), ctx.dev(mps[1])
    Abra = None if mpsl is None else ctx.dev(mpsl[0])
    Bbra = None if mpsl is None else ctx.dev(mpsl[1])
    for m in range(len(mpo)):
        newR = D.env_update_left(B, mpo[m][1], ctx.dev(env[m][1]), Bbra, ctx.ws)
        newL = D.env_update_right(A, mpo[m][0], ctx.dev(env[m][0]), Abra, ctx.ws)
        env[m][0] = like(newL, mps[0])
        env[m][1] = like(newR, mps[0])
    return env
