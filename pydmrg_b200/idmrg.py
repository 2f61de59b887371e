"""Infinite-DMRG driver on the B200 -- ``run_idmrg`` API and ``single_iter`` seam of
pydmrg/idmrg.py (single_iter :44-54, run_iters :84-131, run_idmrg :133-296).  This is where
the reference's two-site mat-vec (tools/diag_tools.py:146-162) and SVD truncation
(tools/mps_tools.py:98-120) live."""
import numpy as np
import torch

from . import mps as mps_tools
from .core import Context, default_context, is_dev, like
from .eigs import calc_eigs
from .env import calc_env_inf, update_env_inf
from .truncate import renorm_inf

VERBOSE = 0


def return_bulk_mpo(mpo):
    """idmrg.py:16-24."""
    return [[op[1], op[2]] for op in mpo]


def return_edge_mpo(mpo):
    """idmrg.py:26-34."""
    return [[op[0], op[3]] for op in mpo]


def make_next_guess(mps, mbd=10, d=2, ctx=None):
    """Zero-pad the two site tensors to the grown bond.  idmrg.py:36-42."""
    ctx = ctx or default_context()
    for state in range(len(mps)):
        A, B = mps[state][0], mps[state][1]
        n1, n2, n3 = A.shape
        pa = (min(mbd, n3) - n2, min(mbd, n3 * d) - n3)
        if is_dev(A):
            mps[state][0] = torch.nn.functional.pad(A, (0, pa[1], 0, pa[0])).contiguous()
            mps[state][1] = torch.nn.functional.pad(B, (0, pa[0], 0, pa[1])).contiguous()
        else:
            mps[state][0] = np.pad(A, ((0, 0), (0, pa[0]), (0, pa[1])), "constant")
            mps[state][1] = np.pad(B, ((0, 0), (0, pa[1]), (0, pa[0])), "constant")
    return mps


def single_iter(N, mps, mpo, env, nStates, alg="davidson", mbd=10, ctx=None, stats=None, **solver):
    """idmrg.py:44-54: two-site solve -> SVD truncation -> both block updates -> padded guess."""
    ctx = ctx or default_context()
    E, vecs, _ = calc_eigs(mps, mpo, env, 0, nStates, alg=alg, oneSite=False, edgePreserveState=False, ctx=ctx,
                           stats=stats, **solver)
    mps, EE, EEs, S = renorm_inf(mps, mpo, env, vecs, mbd, ctx=ctx)
    env = update_env_inf(mps[0], mpo, env, mpsl=None, ctx=ctx)
    mps = make_next_guess(mps, mbd=mbd, ctx=ctx)
    return N, E, EE, EEs, mps, env, S


def checkConv(E_prev, E, tol, iterCnt, maxIter=100, minIter=None, nStates=1, targetState=0):
    """idmrg.py:69-82."""
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


def run_iters(mps, mpo, env, mbd=10, maxIter=100, minIter=None, tol=1e-10, fname=None, nStates=1, targetState=0,
              alg="davidson", preserveState=False, returnState=False, returnEnv=False, returnEntSpec=False,
              orthonormalize=False, ctx=None, stats=None, **solver):
    """idmrg.py:84-131 (returns the final mps/env where the reference names undefined
    variables, :128-130)."""
    ctx = ctx or default_context()
    host_in = not is_dev(mps[0][0])
    mps = [[ctx.dev(t) for t in M] for M in mps]
    env = [[ctx.dev(t) for t in F] for F in env]
    mpoEdge, mpoBulk = return_edge_mpo(mpo), return_bulk_mpo(mpo)
    N, E, EE, EEs, mps, env, S = single_iter(2, mps, mpoEdge, env, nStates, alg=alg, mbd=mbd, ctx=ctx, stats=stats, **solver)
    cont, iterCnt, E_prev = True, 0, E
    conv = False
    N += 2
    while cont:
        N, E, EE, EEs, mps, env, S = single_iter(N, mps, mpoBulk, env, nStates, alg=alg, mbd=mbd, ctx=ctx,
                                                 stats=stats, **solver)
        if VERBOSE > 1:
            print("N={}\tEnergy = {:f}\tdiff={:f}\tEE={:f}".format(N, np.real(E / N), np.real(np.abs(E / N - E_prev)), EE))
        cont, conv, E_prev, iterCnt = checkConv(E_prev, E / N, tol, iterCnt, maxIter, minIter, nStates=nStates,
                                                targetState=targetState)
        N += 2
    if fname is not None:
        mps_tools.save_mps([[Context.host(t) for t in M] for M in mps], fname, gaugeSite=0)
    gap = (E[0] - E[1]) if nStates != 1 else None
    if hasattr(E, "__len__") and np.ndim(E) > 0:
        E = E[targetState]
    output = [E, EE, gap]
    if returnEntSpec:
        output.append(EEs)
    if returnState:
        output.append([[Context.host(t) for t in M] for M in mps] if host_in else mps)
    if returnEnv:
        output.append([[Context.host(t) for t in F] for F in env] if host_in else env)
    return output


def run_idmrg(mpo, initEnv=None, initGuess=None, mbd=[2, 4, 8, 16], tol=1e-2, maxIter=1000, minIter=10, fname=None,
              nStates=1, targetState=0, alg="davidson", preserveState=False, edgePreserveState=False,
              returnState=False, returnEnv=False, returnEntSpec=False, orthonormalize=False, calcLeftState=False,
              *, dtype="c128", ctx=None, stats=None, **solver):
    """Same call signature / return packing as pydmrg/idmrg.py:133-139, :282-296.
    ``mpo`` is the reference's 4-site MPO (edge = sites 0,3; bulk = sites 1,2)."""
    ctx = ctx or default_context(dtype)
    if calcLeftState:
        raise NotImplementedError("run_idmrg(calcLeftState=True) references undefined names in the reference "
                                  "(idmrg.py:240) and never ran; use run_dmrg for left eigenstates")
    mbd = np.atleast_1d(np.asarray(mbd))
    nb = len(mbd)
    tol = np.ones(nb) * tol if not hasattr(tol, "__len__") else np.asarray(tol)
    maxIter = np.ones(nb) * maxIter if not hasattr(maxIter, "__len__") else np.asarray(maxIter)
    minIter = np.ones(nb) * minIter if not hasattr(minIter, "__len__") else np.asarray(minIter)
    Evec = np.zeros(nb, dtype=np.complex128)
    EEvec = np.zeros(nb, dtype=np.complex128)
    gapvec = np.zeros(nb, dtype=np.complex128)
    E = EE = gap = EEs = mpsList = env = None
    for mbdInd, mbdi in enumerate(mbd):
        mbdi = int(mbdi)
        if initGuess is None:
            mps = mps_tools.create_all_mps(2, mbdi, nStates)       # idmrg.py:172
        else:
            mps, _ = mps_tools.load_mps(initGuess + "_mbd" + str(mbdInd))
        env = calc_env_inf(mps, mpo, mbdi, ctx=ctx) if initEnv is None else initEnv
        fname_mbd = None if fname is None else fname + "_mbd" + str(mbdInd)
        out = run_iters(mps, mpo, env, mbd=mbdi, maxIter=maxIter[mbdInd], minIter=minIter[mbdInd], tol=tol[mbdInd],
                        fname=fname_mbd, nStates=nStates, alg=alg, targetState=targetState,
                        preserveState=preserveState, returnState=True, returnEnv=True, returnEntSpec=True,
                        orthonormalize=orthonormalize, ctx=ctx, stats=stats, **solver)
        E, EE, gap, EEs, mpsList, env = out
        Evec[mbdInd], EEvec[mbdInd] = E, EE
        gapvec[mbdInd] = gap if gap is not None else np.nan
    output = [E, EE, gap] if nb == 1 else [Evec, EEvec, gapvec]
    if returnEntSpec:
        output.append(EEs)
    if returnState:
        output.append(mpsList)
    if returnEnv:
        output.append(env)
    return output
