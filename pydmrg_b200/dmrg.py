"""Finite-chain DMRG driver on the B200 -- keeps the ``run_dmrg`` API, argument meaning,
return packing and step-function seam of pydmrg/dmrg.py, and replaces what runs inside a
step (local eigensolve -> renormalise -> one block update) with the device path.

    rightStep / leftStep      dmrg.py:138-148, 174-184   (the drop-in boundary)
    rightSweep / leftSweep    dmrg.py:150-172, 186-208
    checkConv                 dmrg.py:210-223
    run_sweeps                dmrg.py:238-307
    run_dmrg                  dmrg.py:309-508

The reference's finite driver is ONE-site with reduced-density-matrix renormalisation
(SURVEY 0.2); that path is reproduced (``twoSite=False``, default, as in the reference).
``twoSite=True`` runs the two-site sweep named by the north-star: two-site H_eff
(tools/diag_tools.py:146-162) + SVD truncation (tools/mps_tools.py:98-120) + block update,
i.e. the composition of the reference's own iDMRG pieces on a finite chain.

MPS tensors and blocks stay resident in HBM for the whole calculation; NumPy lists cross the
boundary only on entry (initial guess / loaded state) and exit (returnState / returnEnv /
save_mps), so ``mpo`` builders, ``saved_states`` and ``initGuess`` / ``initEnv`` keep the
reference's types.
"""
import numpy as np
import torch

from . import device as D
from . import mps as mps_tools
from .core import Context, default_context, is_dev, like
from .eigs import calc_eigs
from .env import calc_env, update_envL, update_envR
from .mpo import mpo_conj_trans
from .truncate import (rdm_renormalize_left, rdm_renormalize_right, rdm_truncate_twosite_states, svd_truncate,
                       truncate_twosite)

VERBOSE = 0


def _log(level, msg):
    if VERBOSE > level:
        print(msg)


class _Tick:
    """Optional wall-clock profile of a step (stats['profile'] = True): synchronises, so it is off
    by default."""

    def __init__(self, on):
        self.on = bool(on)
        if self.on:
            import time
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def lap(self, stats, key):
        if self.on:
            import time
            torch.cuda.synchronize()
            now = time.perf_counter()
            stats[key] = stats.get(key, 0.0) + now - self.t
            self.t = now


# ---------------------------------------------------------------------------------------
# renormalisation wrappers with the reference's signatures
# ---------------------------------------------------------------------------------------
def renormalizeR(mpsL, v, site, nStates=1, targetState=0, ctx=None, fast=False):
    """dmrg.py:49-91.  ``fast``: single state and the entanglement spectrum of this site is not needed --
    ordered QR instead of the RDM eigen-decomposition (same isometry up to the bond gauge)."""
    ctx = ctx or default_context()
    proto = mpsL[0][site]
    shape = tuple(proto.shape)
    V = ctx.dev(v)
    if V.dim() == 1:
        V = V.unsqueeze(1)
    n_avg = min(nStates, V.shape[1])
    psis = [V[:, i].contiguous() for i in range(n_avg)]
    nxts = [ctx.dev(mpsL[s][site + 1]) for s in range(nStates)]
    A, new_next, EE, EEs = rdm_renormalize_right(psis, shape, nxts, ctx, n_avg, fast=fast)
    for s in range(nStates):
        mpsL[s][site] = like(A, proto)
        mpsL[s][site + 1] = like(new_next[s], proto)
    return mpsL, EE, EEs


def renormalizeL(mpsL, v, site, nStates=1, targetState=0, ctx=None, fast=False):
    """dmrg.py:93-136 (``fast`` as in renormalizeR)."""
    ctx = ctx or default_context()
    proto = mpsL[0][site]
    shape = tuple(proto.shape)
    V = ctx.dev(v)
    if V.dim() == 1:
        V = V.unsqueeze(1)
    n_avg = min(nStates, V.shape[1])
    psis = [V[:, i].contiguous() for i in range(n_avg)]
    prvs = [ctx.dev(mpsL[s][site - 1]) for s in range(nStates)]
    B, new_prev, EE, EEs = rdm_renormalize_left(psis, shape, prvs, ctx, n_avg, fast=fast)
    for s in range(nStates):
        mpsL[s][site] = like(B, proto)
        mpsL[s][site - 1] = like(new_prev[s], proto)
    return mpsL, EE, EEs


def _merge_stats(stats, loc):
    """Add the counters of one local solve to the caller's statistics."""
    if stats is None:
        return
    for key, val in loc.items():
        if key == "last_residual":
            stats[key] = val
        elif isinstance(val, list):
            stats.setdefault(key, []).extend(val)
        else:
            stats[key] = stats.get(key, 0) + val


# ---------------------------------------------------------------------------------------
# steps (the drop-in seam)
# ---------------------------------------------------------------------------------------
def rightStep(mpsL, W, F, site, nStates=1, alg="davidson", preserveState=False, orthonormalize=False,
              ctx=None, stats=None, fast_renorm=False, **solver):
    """dmrg.py:138-148: solve at ``site``, renormalise to the right, grow the left block.
    ``fast_renorm`` (set by the sweeps for the sites whose EE is not reported): see renormalizeR."""
    ctx = ctx or default_context()
    E, v, ovlp = calc_eigs(mpsL, W, F, site, nStates, alg=alg, preserveState=preserveState,
                           orthonormalize=orthonormalize, ctx=ctx, stats=stats, **solver)
    mpsL, EE, EEs = renormalizeR(mpsL, v, site, nStates=nStates, ctx=ctx, fast=fast_renorm and nStates == 1)
    F = update_envR(mpsL[0], W, F, site, ctx=ctx)
    return E, mpsL, F, EE, EEs


def leftStep(mpsL, W, F, site, nStates=1, alg="davidson", preserveState=False, orthonormalize=False,
             ctx=None, stats=None, fast_renorm=False, **solver):
    """dmrg.py:174-184."""
    ctx = ctx or default_context()
    E, v, ovlp = calc_eigs(mpsL, W, F, site, nStates, alg=alg, preserveState=preserveState,
                           orthonormalize=orthonormalize, ctx=ctx, stats=stats, **solver)
    mpsL, EE, EEs = renormalizeL(mpsL, v, site, nStates=nStates, ctx=ctx, fast=fast_renorm and nStates == 1)
    F = update_envL(mpsL[0], W, F, site, ctx=ctx)
    return E, mpsL, F, EE, EEs


def twoSiteStep(mpsL, W, F, site, direction, mbd, nStates=1, alg="davidson", ctx=None, stats=None, recycle=None,
                exact_trunc=False, **solver):
    """Two-site bond step on (site, site+1): two-site local solve (tools/diag_tools.py:146-162,
    guess :257), SVD truncation to ``mbd`` (tools/mps_tools.py:98-120), block update
    (tools/env_tools.py:40-50 / 28-38).  The singular values go to the new centre site.
    Returns (E, mpsL, F, EE, EEs, S).

    ``gauge_shortcut=True`` (solver option, off by default): a bond whose local solve stops in its first cycle -- the
    state did not change, the normal case in the last sweeps of a run -- is re-gauged from its centre tensor (one
    ordered QR) instead of being truncated again; the centre bond, whose spectrum is reported, never takes it."""
    ctx = ctx or default_context()
    solver = dict(solver)
    shortcut = bool(solver.pop("gauge_shortcut", False))
    proto = mpsL[0][site]
    d1, Dl, _ = proto.shape
    d2, _, Dr = mpsL[0][site + 1].shape
    prof = stats is not None and stats.get("profile")
    tick = _Tick(prof)
    if nStates != 1:
        # several target states: block Davidson on the two-site problem, then the reference's
        # state-averaged RDM (dmrg.py:56-66) gives ONE isometry for all states; each keeps its own centre.
        # As in the reference's one-site steps the blocks are built from mpsL[0] (dmrg.py:147,183).
        E, v, _ = calc_eigs(mpsL, W, F, site, nStates, alg=alg, oneSite=False, edgePreserveState=False, ctx=ctx,
                            stats=stats, return_device=True, **solver)
        tick.lap(stats, "t_eig")
        n_avg = min(nStates, v.shape[1])
        psis = [v[:, i].contiguous() for i in range(n_avg)]
        As, Bs, S, EE, EEs = rdm_truncate_twosite_states(psis, (d1, d2, Dl, Dr), mbd, ctx, direction, spectrum=exact_trunc)
        tick.lap(stats, "t_svd")
        for s in range(nStates):
            mpsL[s][site] = like(As[min(s, n_avg - 1)], proto)
            mpsL[s][site + 1] = like(Bs[min(s, n_avg - 1)], proto)
        A, B = As[0], Bs[0]
    else:
        loc = {}
        E, v, _ = calc_eigs(mpsL, W, F, site, 1, alg=alg, oneSite=False, edgePreserveState=False, ctx=ctx,
                            stats=loc, return_device=True, **solver)
        _merge_stats(stats, loc)
        tick.lap(stats, "t_eig")
        kmid = mpsL[0][site].shape[2]
        if shortcut and not exact_trunc and alg == "davidson" and loc.get("cycles") == 1 and kmid <= mbd:
            # The solver stopped in its first cycle: the Ritz vector IS the (normalised) guess A_c . B, whose Schmidt
            # decomposition across this bond follows from the centre tensor alone -- a gauge move (ordered QR of a
            # (d*D, k) matrix, the one-site renormalisation of dmrg.py:49-136 without a spectrum) instead of the
            # truncation of the (d*Dl, d*Dr) wave-function.  Same state, same blocks up to the bond gauge.
            if direction == "right":
                A, nxt, EE, EEs = rdm_renormalize_right([ctx.dev(mpsL[0][site]).reshape(-1)], (d1, Dl, kmid),
                                                        [ctx.dev(mpsL[0][site + 1])], ctx, 1, fast=True)
                B = nxt[0]
            else:
                B, prv, EE, EEs = rdm_renormalize_left([ctx.dev(mpsL[0][site + 1]).reshape(-1)], (d2, kmid, Dr),
                                                       [ctx.dev(mpsL[0][site])], ctx, 1, fast=True)
                A = prv[0]
            S = None
            if stats is not None:
                stats["n_gauge_moves"] = stats.get("n_gauge_moves", 0) + 1
        else:
            psi = v[:, 0].contiguous()
            A, B, S, EE, EEs = truncate_twosite(psi, (d1, d2, Dl, Dr), mbd, ctx, direction, recycle=recycle, site=site,
                                                warm=(mpsL[0][site], mpsL[0][site + 1]), exact=exact_trunc, stats=stats)
        tick.lap(stats, "t_svd")
        mpsL[0][site] = like(A, proto)
        mpsL[0][site + 1] = like(B, proto)
    shard = solver.get("shard")
    if direction == "right":
        split = shard is not None and shard.wants_env(A.shape[1], A.shape[2])
        for m in range(len(W)):
            if split:
                out = shard.env_update("right", A, W[m][site], ctx.dev(F[m][site]), ctx)
            else:
                out = D.env_update_right(A, W[m][site], ctx.dev(F[m][site]), None, ctx.ws)
            F[m][site + 1] = like(out, proto)
    else:
        split = shard is not None and shard.wants_env(B.shape[1], B.shape[2])
        for m in range(len(W)):
            if split:
                out = shard.env_update("left", B, W[m][site + 1], ctx.dev(F[m][site + 2]), ctx)
            else:
                out = D.env_update_left(B, W[m][site + 1], ctx.dev(F[m][site + 2]), None, ctx.ws)
            F[m][site + 1] = like(out, proto)
    if split and stats is not None:
        stats["n_sharded_env"] = stats.get("n_sharded_env", 0) + 1
    tick.lap(stats, "t_env")
    return E, mpsL, F, EE, EEs, S


# ---------------------------------------------------------------------------------------
# sweeps
# ---------------------------------------------------------------------------------------
def rightSweep(mpsL, W, F, iterCnt, nStates=1, alg="davidson", preserveState=False, startSite=None,
               endSite=None, orthonormalize=False, ctx=None, stats=None, fast_renorm=False, **solver):
    """dmrg.py:150-172."""
    N = len(mpsL[0])
    startSite = 0 if startSite is None else startSite
    endSite = N - 1 if endSite is None else endSite
    Ereturn = EE = EEs = None
    _log(1, "Right Sweep {}".format(iterCnt))
    for site in range(startSite, endSite):
        E, mpsL, F, _EE, _EEs = rightStep(mpsL, W, F, site, nStates, alg=alg, preserveState=preserveState,
                                          orthonormalize=orthonormalize, ctx=ctx, stats=stats,
                                          fast_renorm=fast_renorm and site != N // 2, **solver)
        _log(2, "\tEnergy at Site {}: {}".format(site, E))
        if site == N // 2:
            Ereturn, EE, EEs = E, _EE, _EEs
    return Ereturn, mpsL, F, EE, EEs


def leftSweep(mpsL, W, F, iterCnt, nStates=1, alg="davidson", preserveState=False, startSite=None,
              endSite=None, orthonormalize=False, ctx=None, stats=None, fast_renorm=False, **solver):
    """dmrg.py:186-208."""
    N = len(mpsL[0])
    startSite = N - 1 if startSite is None else startSite
    endSite = 0 if endSite is None else endSite
    Ereturn = EE = EEs = None
    _log(1, "Left Sweep {}".format(iterCnt))
    for site in range(startSite, endSite, -1):
        E, mpsL, F, _EE, _EEs = leftStep(mpsL, W, F, site, nStates, alg=alg, preserveState=preserveState,
                                         orthonormalize=orthonormalize, ctx=ctx, stats=stats,
                                         fast_renorm=fast_renorm and site != N // 2, **solver)
        _log(2, "\tEnergy at Site {}: {}".format(site, E))
        if site == N // 2:
            Ereturn, EE, EEs = E, _EE, _EEs
    return Ereturn, mpsL, F, EE, EEs


def twoSiteSweep(mpsL, W, F, mbd, alg="davidson", ctx=None, stats=None, sites=None, recycle=None, nStates=1, **solver):
    """One full two-site sweep: bonds 0..N-2 moving right, then N-2..0 moving left.
    E / EE / S are recorded at the centre bond (N//2-1, N//2).

    ``recycle``: a dict kept by the caller across sweeps of the SAME state and bond dimension; when
    given, bonds whose previous bases are cached are truncated by the warm-started subspace step
    (truncate.recycled_truncate) instead of a fresh eigen-decomposition.  The centre bond is always
    decomposed exactly, so the reported entanglement spectrum never depends on the cache."""
    N = len(mpsL[0])
    E = EE = EEs = S = None
    for site in range(0, N - 1):
        e, mpsL, F, ee, ees, s = twoSiteStep(mpsL, W, F, site, "right", mbd, nStates=nStates, alg=alg, ctx=ctx, stats=stats,
                                             recycle=recycle, exact_trunc=(site == N // 2 - 1), **solver)
        if site == N // 2 - 1:
            E, EE, EEs, S = e, ee, ees, s
    for site in range(N - 2, -1, -1):
        e, mpsL, F, ee, ees, s = twoSiteStep(mpsL, W, F, site, "left", mbd, nStates=nStates, alg=alg, ctx=ctx, stats=stats,
                                             recycle=recycle, exact_trunc=(site == N // 2 - 1), **solver)
        if site == N // 2 - 1:
            E, EE, EEs, S = e, ee, ees, s
    return E, mpsL, F, EE, EEs, S


def checkConv(E_prev, E, tol, iterCnt, maxIter, minIter, nStates=1, targetState=0, EE=None, EEspec=(None,)):
    """dmrg.py:210-223."""
    if nStates != 1:
        E = E[targetState]
    if (np.abs(E - E_prev) < tol) and (iterCnt > minIter):
        cont, conv = False, True
    elif iterCnt > maxIter - 1:
        cont, conv = False, False
    else:
        iterCnt += 1
        E_prev = E
        cont, conv = True, False
    return cont, conv, E_prev, iterCnt


def printResults(converged, E, EE, EEspec, gap):
    """dmrg.py:225-236."""
    _log(1, "#" * 75)
    _log(0, ("Converged at E = {}" if converged else "Convergence not acheived, E = {}").format(E))
    _log(2, "\tGap = {}".format(gap))
    _log(2, "\tEntanglement Entropy  = {}".format(EE))
    _log(1, "#" * 75)


def run_sweeps(mpsL, W, F, initGuess=None, maxIter=0, minIter=None, tol=1e-10, fname=None, nStates=1,
               targetState=0, alg="davidson", preserveState=False, gaugeSiteLoad=0, gaugeSiteSave=0,
               returnState=False, returnEnv=False, returnEntSpec=False, orthonormalize=False,
               ctx=None, stats=None, twoSite=False, mbd=None, recycle=True, **solver):
    """dmrg.py:238-307.  ``mpsL`` / ``F`` may be NumPy lists (they are moved to the device).
    ``recycle``: two-site sweeps after the first reuse each bond's previous singular bases
    (truncate.truncate_twosite); False decomposes every bond from scratch."""
    ctx = ctx or default_context()
    shortcut = bool(solver.pop("gauge_shortcut", False))         # two-site only (twoSiteStep)
    host_in = not is_dev(mpsL[0][0])
    mpsL = [[ctx.dev(t) for t in M] for M in mpsL]
    F = [[ctx.dev(t) for t in Fm] for Fm in F]
    # The gauge site carries the norm of the whole state; for the random start of a long chain it is ~1e-195
    # (N=200: entries 1e-2 per site, mps_tools.py:216-235) and its SQUARE underflows in the solver's first
    # normalisation (the reference divides by zero there too).  The norm of an eigenvector guess is arbitrary:
    # rescale it once, max-abs first so that nothing under- or overflows.
    for M in mpsL:
        t = M[gaugeSiteLoad]
        t = t / torch.clamp(t.abs().max(), min=1e-300)
        M[gaugeSiteLoad] = t / torch.linalg.vector_norm(t)
    kw = dict(nStates=nStates, alg=alg, preserveState=preserveState, orthonormalize=orthonormalize, ctx=ctx,
              stats=stats, fast_renorm=bool(recycle), **solver)
    cont, iterCnt, E_prev = True, 0, 0
    conv = False
    S = None
    if twoSite:
        if mbd is None:
            mbd = mps_tools.maxBondDim([[t for t in mpsL[0]]])
        cache = {} if (recycle and nStates == 1) else None      # the warm-started truncation follows ONE state
        while cont:
            E, mpsL, F, EE, EEs, S = twoSiteSweep(mpsL, W, F, mbd, alg=alg, ctx=ctx, stats=stats, recycle=cache,
                                                  nStates=nStates, gauge_shortcut=shortcut, **solver)
            cont, conv, E_prev, iterCnt = checkConv(E_prev, E, tol, iterCnt, maxIter, minIter, nStates=nStates,
                                                    targetState=targetState)
        gaugeSiteSave = 0      # after a left sweep the centre is on site 0
    else:
        if gaugeSiteLoad != 0:                                  # dmrg.py:248-259
            E, mpsL, F, EE, EEs = rightSweep(mpsL, W, F, iterCnt, startSite=gaugeSiteLoad, **kw)
            E, mpsL, F, EE, EEs = leftSweep(mpsL, W, F, iterCnt, **kw)
        while cont:                                             # dmrg.py:260-271
            E, mpsL, F, EE, EEs = rightSweep(mpsL, W, F, iterCnt, **kw)
            E, mpsL, F, EE, EEs = leftSweep(mpsL, W, F, iterCnt, **kw)
            cont, conv, E_prev, iterCnt = checkConv(E_prev, E, tol, iterCnt, maxIter, minIter, nStates=nStates,
                                                    targetState=targetState)
        if gaugeSiteSave != 0:                                  # dmrg.py:272-291
            _E, mpsL, F, _EE, _EEs = rightSweep(mpsL, W, F, iterCnt + 1, endSite=gaugeSiteSave, **kw)
            _, v, _ = calc_eigs(mpsL, W, F, gaugeSiteSave, nStates, alg=alg, preserveState=preserveState,
                                orthonormalize=orthonormalize, ctx=ctx, stats=stats, **solver)
            shp = mpsL[0][gaugeSiteSave].shape
            for state in range(nStates):
                mpsL[state][gaugeSiteSave] = v[:, min(state, v.shape[1] - 1)].reshape(shp).contiguous()
            if _E is not None:
                E, EE, EEs = _E, _EE, _EEs
    if fname is not None:
        mps_tools.save_mps([[Context.host(t) for t in M] for M in mpsL], fname, gaugeSite=gaugeSiteSave)
    gap = (E[0] - E[1]) if nStates != 1 else None
    if hasattr(E, "__len__") and np.ndim(E) > 0:
        E = E[targetState]
    printResults(conv, E, EE, EEs, gap)
    output = [E, EE, gap]
    if returnEntSpec:
        output.append(EEs)
    if returnState:
        output.append([[Context.host(t) for t in M] for M in mpsL] if host_in else mpsL)
    if returnEnv:
        output.append([[Context.host(t) for t in Fm] for Fm in F] if host_in else F)
    return output


def run_dmrg(mpo, initEnv=None, initGuess=None, mbd=[2, 4, 8, 16], tol=1e-5, maxIter=10, minIter=0, fname=None,
             nStates=1, targetState=0, constant_mbd=False, alg="davidson", preserveState=False,
             gaugeSiteSave=None, returnState=False, returnEnv=False, returnEntSpec=False, orthonormalize=False,
             calcLeftState=False, *, dtype="c128", twoSite=False, ctx=None, stats=None, pad_bonds=None, name_suffix="",
             **solver):
    """Same call signature and return packing as pydmrg/dmrg.py:309-315 / :491-508.
    Keyword-only extensions: ``dtype`` ('c128' as the reference, or 'f64' for real problems),
    ``twoSite`` (two-site SVD sweeps), ``stats`` (dict: mat-vec / cycle counters), solver
    options (``res_tol``, ``lindep``, ``max_cycle``), ``recycle`` (run_sweeps) and ``pad_bonds``:
    whether a state carried to a larger ``mbd`` is zero-padded to it first (mps_tools.py:244-279).
    One-site sweeps cannot grow a bond, so they always pad (the reference's behaviour); two-site sweeps
    grow every bond by up to a factor d per visit on their own, so by default they do NOT pad -- the
    first sweeps of a chi stage then run on the small tensors instead of on zero-filled large ones.
    ``name_suffix``: appended AFTER the ``_mbd{i}`` part of every state file read or written (``'_left'`` lets a run
    on ``mpo_conj_trans(mpo)`` use the reference's names for the left eigenstate, dmrg.py:423,436,476).
    An in-memory ``initGuess`` seeds the first ``mbd`` stage only; later stages continue from the previous stage's
    state (the in-memory equivalent of the reference's save / load / increase_all_mbd between stages)."""
    ctx = ctx or default_context(dtype)
    N = len(mpo[0])
    if pad_bonds is None:
        pad_bonds = not twoSite
    grow = (lambda mL, chi: mps_tools.increase_all_mbd(mL, chi)) if pad_bonds else (lambda mL, chi: mL)
    if gaugeSiteSave is None:
        gaugeSiteSave = N // 2 + 1
    mbd = np.atleast_1d(np.asarray(mbd))
    nb = len(mbd)
    tol = np.ones(nb) * tol if not hasattr(tol, "__len__") else np.asarray(tol)
    maxIter = np.ones(nb) * maxIter if not hasattr(maxIter, "__len__") else np.asarray(maxIter)
    minIter = np.ones(nb) * minIter if not hasattr(minIter, "__len__") else np.asarray(minIter)
    assert len(tol) == nb and len(maxIter) == nb and len(minIter) == nb
    mpol = mpo_conj_trans(mpo) if calcLeftState else None       # dmrg.py:338
    Evec = np.zeros(nb, dtype=np.complex128)
    EEvec = np.zeros(nb, dtype=np.complex128)
    gapvec = np.zeros(nb, dtype=np.complex128)
    EEvecl = np.zeros(nb, dtype=np.complex128)
    mpsList = mpslList = env = envl = EEs = EEsl = None
    E = EE = gap = EEl = None
    for mbdInd, mbdi in enumerate(mbd):
        mbdi = int(mbdi)
        _log(1, "Starting Calc for MBD = {}".format(mbdi))
        # ---- initial MPS (dmrg.py:350-389) ----------------------------------------------
        if initGuess is None:
            if mbdInd != 0 and fname is not None:
                mpsList, gSite = mps_tools.load_mps(fname + "_mbd" + str(mbdInd - 1) + name_suffix)
                mpsList = grow(mpsList, mbdi)
                if calcLeftState:
                    mpslList, glSite = mps_tools.load_mps(fname + "_mbd" + str(mbdInd - 1) + name_suffix + "_left")
                    mpslList = grow(mpslList, mbdi)
            else:
                mpsList = mps_tools.make_all_mps_right(mps_tools.create_all_mps(N, mbdi, nStates))
                gSite = 0
                if calcLeftState:                                # dmrg.py:369-373
                    mps_tools.create_all_mps(N, mbdi, nStates)   # drawn and discarded, as the reference does
                    mpslList = mps_tools.make_all_mps_right([[t.copy() for t in M] for M in mpsList])
                    glSite = 0
        elif isinstance(initGuess, str):
            idx = mbdInd if mbdInd == 0 else mbdInd - 1
            mpsList, gSite = mps_tools.load_mps(initGuess + "_mbd" + str(idx) + name_suffix)
            if calcLeftState:
                mpslList, glSite = mps_tools.load_mps(initGuess + "_mbd" + str(idx) + name_suffix + "_left")
            if mbdInd != 0:
                mpsList = grow(mpsList, mbdi)
                if calcLeftState:
                    mpslList = grow(mpslList, mbdi)
        elif mbdInd != 0:
            # in-memory guess: it seeded stage 0; this stage continues from the previous stage's result, whose
            # centre sits where run_sweeps left it (site 0 after two-site sweeps, gaugeSiteSave after one-site ones)
            gSite = 0 if twoSite else gaugeSiteSave
            mpsList = grow([[Context.host(t) for t in M] for M in mpsList], mbdi)
            if calcLeftState:
                glSite = gSite
                mpslList = grow([[Context.host(t) for t in M] for M in mpslList], mbdi)
        else:
            # an in-memory mpsList (lists of ndarrays), gauge at site 0 -- or (mpsList, gaugeSite)
            guess, gSite = (initGuess if (isinstance(initGuess, tuple) and len(initGuess) == 2
                                          and np.ndim(initGuess[1]) == 0) else (initGuess, 0))
            mpsList = [[np.array(t) for t in M] for M in (guess[0] if calcLeftState else guess)]
            mpsList = grow(mpsList, mbdi)
            if calcLeftState:
                mpslList = grow([[np.array(t) for t in M] for M in guess[1]], mbdi)
                glSite = gSite
        gSite = int(gSite)
        mpsList = [[ctx.dev(t) for t in M] for M in mpsList]
        if calcLeftState:
            glSite = int(glSite)
            mpslList = [[ctx.dev(t) for t in M] for M in mpslList]
        if twoSite:
            # two-site sweeps start at bond (0, 1): a state whose centre is elsewhere (a one-site checkpoint saved at
            # N//2+1, the reference's default, or initGuess=(mps, g)) is re-gauged to site 0 first -- otherwise the
            # blocks F[1..g] built by calc_env(gaugeSite=g) are LEFT blocks where the sweep expects right ones
            from .gauge import move_gauge
            if gSite != 0:
                mpsList = [move_gauge(M, gSite, 0, ctx=ctx) for M in mpsList]
                gSite = 0
            if calcLeftState and glSite != 0:
                mpslList = [move_gauge(M, glSite, 0, ctx=ctx) for M in mpslList]
                glSite = 0
        # ---- environment (dmrg.py:391-397) ----------------------------------------------
        if initEnv is None:
            env = calc_env(mpsList[0], mpo, mbdi, gaugeSite=gSite, ctx=ctx)
            if calcLeftState:
                envl = calc_env(mpslList[0], mpol, mbdi, gaugeSite=glSite, ctx=ctx)
        else:
            env = initEnv
            if calcLeftState:
                env, envl = initEnv[0], initEnv[1]
        fname_mbd = None if fname is None else fname + "_mbd" + str(mbdInd) + name_suffix
        common = dict(maxIter=maxIter[mbdInd], minIter=minIter[mbdInd], tol=tol[mbdInd], nStates=nStates, alg=alg,
                      targetState=targetState, gaugeSiteSave=gaugeSiteSave, preserveState=preserveState,
                      returnState=True, returnEnv=True, returnEntSpec=True, orthonormalize=orthonormalize,
                      ctx=ctx, stats=stats, twoSite=twoSite, mbd=mbdi, **solver)
        _log(0, "Calculating Right Eigenstate")
        out = run_sweeps(mpsList, mpo, env, fname=fname_mbd, gaugeSiteLoad=gSite, **common)
        E, EE, gap, EEs, mpsList, env = out
        Evec[mbdInd], EEvec[mbdInd] = E, EE
        gapvec[mbdInd] = gap if gap is not None else np.nan
        if calcLeftState:                                        # dmrg.py:448-488
            _log(0, "Calculating Left Eigenstate")
            out = run_sweeps(mpslList, mpol, envl, fname=None if fname_mbd is None else fname_mbd + "_left",
                             gaugeSiteLoad=glSite, **common)
            _, EEl, _, EEsl, mpslList, envl = out
            EEvecl[mbdInd] = EEl
    # ---- result packing (dmrg.py:490-508) ----------------------------------------------------
    host = lambda L: [[Context.host(t) for t in M] for M in L]
    mps_out, env_out = host(mpsList), host(env)
    if calcLeftState:
        EE = [EE, EEl]
        EEs = [EEs, EEsl]
        mps_out = [mps_out, host(mpslList)]
        env_out = [env_out, host(envl)]
    output = [E, EE, gap] if nb == 1 else [Evec, EEvec, gapvec]
    if returnEntSpec:
        output.append(EEs)
    if returnState:
        output.append(mps_out)
    if returnEnv:
        output.append(env_out)
    return output
