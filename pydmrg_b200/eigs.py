"""Local eigenproblem of a DMRG step on the device -- the replacement of
pydmrg/tools/diag_tools.py (make_ham_func :106-171, calc_eigs :292-313, calc_eigs_davidson
:243-290, calc_eigs_exact :193-212, calc_eigs_arnoldi :214-241, check_overlap :71-104).

Conventions kept from the reference:
  * ``Hfun(x)`` returns  -H_eff x  (diag_tools.py:125,162); the solver looks for the smallest
    real part and the sign is flipped back, so the state with the LARGEST Re(E) is targeted;
  * one-site guess = current site tensor (:255), two-site guess = einsum('ijk,lkm->iljm') (:257);
  * for one root ``E`` is a scalar and ``vecs`` has shape (n, 1) after check_overlap;
  * at the chain ends ``preserveState`` is forced on (:288) and, if no eigenvector overlaps
    the guess by more than 0.98, the old guess is kept (:93-97).
"""
import numpy as np
import scipy.linalg as sla
import torch

from . import device as D
from .core import default_context, is_dev, like
from .davidson import davidson


def _site_ops(M, W, F, site, oneSite, ctx):
    """Collect (L, Wc, R) per MPO term and the local problem shape."""
    if oneSite:
        d, Dl, Dr = M[site].shape
        P = d
        shape = (d, Dl, Dr)
    else:
        d1, Dl, _ = M[site].shape
        d2, _, Dr = M[site + 1].shape
        P = d1 * d2
        shape = (d1, d2, Dl, Dr)
    ops = []
    for m in range(len(W)):
        L = ctx.dev(F[m][site])
        if oneSite:
            R = ctx.dev(F[m][site + 1])
            Wc = D.combine_onesite(W[m][site], L.shape[1], R.shape[1], shape[0], ctx.npdtype)
        else:
            # finite chains: right block of the bond is F[site+2]; the iDMRG two-slot list
            # [L, R] (tools/env_tools.py:74-75) is addressed as F[site+1] by the reference
            R = ctx.dev(F[m][site + 2] if len(F[m]) > site + 2 else F[m][site + 1])
            W1, W2 = W[m][site], W[m][site + 1]
            # None = identity passing the MPO bond through, as in the one-site code
            # (diag_tools.py:113-115).  The reference's two-site Hfun skips such terms (:150-151,
            # "Adapt to No operator"), which drops wrap-around strings; here they are applied.
            if W1 is None:
                W1 = np.einsum("ab,ij->abij", np.eye(L.shape[1]), np.eye(shape[0]))
            if W2 is None:
                W2 = np.einsum("ab,ij->abij", np.eye(R.shape[1]), np.eye(shape[1]))
            Wc = D.combine_twosite(W1, W2, ctx.npdtype)
        ops.append((L, Wc, R))
    return ops, P, shape


def make_plan(M, W, F, site, oneSite=True, ctx=None, block_masks=None):
    ctx = ctx or default_context()
    ops, P, shape = _site_ops(M, W, F, site, oneSite, ctx)
    Dl, Dr = shape[-2], shape[-1]
    return D.HeffPlan(ctx.code, P, Dl, Dr, ops, ctx.ws, block_masks=block_masks), shape


def make_ham_func(M, W, F, site, usePrecond=False, debug=False, oneSite=True, ctx=None):
    """Returns (Hfun, precond) with the reference's closure contract (diag_tools.py:106-171):
    ``Hfun(x)`` maps a flat vector (ndarray -> ndarray, or CUDA tensor -> CUDA tensor) to -H x."""
    ctx = ctx or default_context()
    plan, shape = make_plan(M, W, F, site, oneSite, ctx)

    stage = {}

    def Hfun(x):
        if is_dev(x):
            xd = ctx.dev(x).reshape(-1)
            yd = torch.empty_like(xd)
            plan.apply(xd, yd, -1.0)
            return yd
        # host vector in, host vector out (the reference's contract).  The device buffers are kept by the closure; the
        # result is a FRESH ndarray on page-locked memory from torch's caching host allocator (solvers keep every image
        # they are handed, so a staging buffer must never be reused) -- D2H at link speed and no extra host copy.  A
        # page-locked input (e.g. a vector this closure returned) goes up asynchronously as well.
        xh = torch.from_numpy(np.ascontiguousarray(np.asarray(x).reshape(-1), dtype=ctx.npdtype))
        if "x" not in stage:
            stage["x"] = torch.empty(plan.n, dtype=ctx.tdtype, device=ctx.device)
            stage["y"] = torch.empty_like(stage["x"])
        stage["x"].copy_(xh, non_blocking=True)
        plan.apply(stage["x"], stage["y"], -1.0)
        yh = torch.empty(plan.n, dtype=ctx.tdtype, pin_memory=True)
        yh.copy_(stage["y"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return yh.numpy()

    def precond(dx, e, x0):                                   # diag_tools.py:137-139 (identity)
        return dx

    Hfun.plan = plan
    return Hfun, precond


def two_site_guess(A, B, ctx):
    """theta[i,l,j,m] = sum_k A[i,j,k] B[l,k,m]   (diag_tools.py:257) as d1 batched NT GEMMs."""
    d1, Dl, Dm = A.shape
    d2, _, Dr = B.shape
    Bt = ctx.empty(d2, Dr, Dm)
    D.permute4(B, Bt, (d2, Dr, Dm), (Dm * Dr, 1, Dr))
    theta = ctx.empty(d1, d2, Dl, Dr)
    for i in range(d1):
        D.gemm_nt(A[i], Bt, theta[i], M=Dl, N=Dr, K=Dm, batch=d2, lda=Dm, ldb=Dm, ldc=Dr,
                  strideA=0, strideB=Dr * Dm, strideC=Dl * Dr)
    return theta


def check_overlap(guess, vecs, E, preserveState, ctx):
    """diag_tools.py:71-104 with device vectors.  vecs: list of device vectors."""
    ops = D.VecOps(ctx.code, guess.numel(), ctx.device) if not hasattr(ctx, "_ovl_ops") or ctx._ovl_ops.n != guess.numel() else ctx._ovl_ops
    ctx._ovl_ops = ops
    slab = torch.stack(vecs) if len(vecs) > 1 else vecs[0].unsqueeze(0)
    ov = np.abs(ops.dots(slab, len(vecs), guess)[:len(vecs)].cpu().numpy())      # |<v_j|guess>|
    matched = False
    E = np.atleast_1d(np.asarray(E)).copy()
    for j in range(len(vecs)):
        if ov[j] > 0.98:
            matched = True
            if j != 0 and preserveState:
                vecs[0], vecs[j] = vecs[j], vecs[0]
                E[0], E[j] = E[j], E[0]
                ov[0], ov[j] = ov[j], ov[0]
    if preserveState and not matched:
        vecs = [guess.clone()]
        ov = np.abs(ops.dots(guess.unsqueeze(0), 1, guess)[:1].cpu().numpy())
    return E, vecs, float(ov[0])


def calc_eigs(mpsL, W, F, site, nStates, alg="davidson", preserveState=False, edgePreserveState=True,
              orthonormalize=False, oneSite=True, ctx=None, stats=None, res_tol=None, lindep=1e-14,
              fixed_cycles=None, return_device=None, plan=None, native=None, max_cycle=1000, shard=None, max_space=12,
              precond=None, precond_floor=0.1, restart_keep=0, arnoldi_tol=1e-5, stagnation=0):
    """calc_eigs, diag_tools.py:292-313.  Returns (E, vecs, ovlp): E scalar (nStates == 1) or
    array; vecs (n, k) -- NumPy when the MPS was NumPy, CUDA tensor otherwise.

    ``shard`` (parallel.ShardSpec): every rank runs the same sweep on replicated tensors and the mat-vec of
    bonds with min(Dl, Dr) >= shard.min_dim is split over the ranks (one all-reduce of y per
    mat-vec).  The Davidson control scalars and the final eigenvectors are taken from rank 0, so the ranks
    cannot drift apart.  ``max_space``: Davidson subspace before a restart (la_tools.py:318 default 12, widened by
    3 per root, :412); a larger one trades vector passes for fewer mat-vecs in cold sweeps.  ``precond='diag'``:
    diagonal Davidson preconditioner built from diag(H_eff) (davidson.davidson) instead of the reference's identity.
    ``restart_keep`` > 0: thick restart of the Davidson subspace; ``stagnation`` > 0: stop a solve whose residual has not
    been halved within that many cycles (both davidson.davidson; opt-in, beyond the reference's algorithm).

    ``alg='arnoldi'`` (calc_eigs_arnoldi, diag_tools.py:214-241: ARPACK ``eigs(k=nStates, which='SR', tol=1e-5, v0=guess)``)
    runs DEVICE-RESIDENT thick-restarted Arnoldi: the same Krylov kernels and C++ loop with the identity preconditioner
    (expansion by residuals = the Krylov space of v0), ARPACK's subspace size ncv = max(2k+1, 20), a thick restart that keeps
    the wanted + k + 2 Ritz vectors (Krylov-Schur, mathematically the implicit restart ARPACK performs) and ARPACK's
    relative stopping rule ||r|| <= tol |theta|.  ``alg='arnoldi_host'`` keeps SciPy's ARPACK on the host driving the device
    mat-vec (one D2H + H2D of the vector per mat-vec) for cross-checks."""
    ctx = ctx or default_context()
    own_plan = plan is None
    sharded = None
    if own_plan:
        plan, shape = make_plan(mpsL[0], W, F, site, oneSite, ctx)
        if shard is not None and shard.world > 1 and alg == "davidson" and min(shape[-2], shape[-1]) >= shard.min_dim:
            sharded = shard.heff(ctx, mpsL[0], W, F, site, oneSite)
    n = plan.n
    dev_in = is_dev(mpsL[0][site]) if return_device is None else return_device
    guesses = []
    nS = min(nStates, n - 1) if n > 1 else 1
    for s in range(nS):
        if oneSite:
            guesses.append(ctx.dev(mpsL[s][site]).reshape(-1))
        else:
            guesses.append(two_site_guess(ctx.dev(mpsL[s][site]), ctx.dev(mpsL[s][site + 1]), ctx).reshape(-1))
    if precond not in (None, "diag"):
        raise ValueError("precond must be None or 'diag'")
    pc = dict(diag=plan.diag(-1.0), precond_floor=precond_floor) if (precond == "diag" and alg == "davidson") else {}
    fused = None
    if alg == "davidson" and sharded is not None and nS <= 4:
        fused = shard.fused_heff(ctx, mpsL[0], W, F, site, oneSite)
    if fused is not None:
        # C++ loop on every rank over the fused peer-memory mat-vec: identical start vectors + bit-identical images +
        # fixed-order reductions = identical control flow on all ranks, with no collective inside the solve
        for gk in guesses:
            shard.consensus(gk)
        # (the preconditioner ``pc`` holds the diagonal of the WHOLE operator from the unsharded plan, identical on every
        # rank -- the shard's own diagonal covers its ket range only and would make the ranks diverge)
        vals, vecs = davidson(None, guesses, n, ctx.code, ctx.device, nroots=nS, lindep=lindep, res_tol=res_tol,
                              fixed_cycles=fixed_cycles, stats=stats, plan=fused.plan, native=True, max_cycle=max_cycle,
                              max_space=max_space, restart_keep=restart_keep, stagnation=stagnation, comm=fused._comm, **pc)
        fused.peer.check()
        E = -vals
        fused.close()
        if stats is not None:
            stats["n_sharded_solves"] = stats.get("n_sharded_solves", 0) + 1
            stats["n_fused_sharded_solves"] = stats.get("n_fused_sharded_solves", 0) + 1
    elif alg == "davidson" and sharded is not None:
        # identical start vectors on every rank: together with the all-reduced images and the deterministic vector
        # kernels this keeps the ranks' Krylov bases identical (a residual of norm 1e-9 that differed by 1e-13
        # between ranks would otherwise become a 1e-4 relative inconsistency of the next, normalised, direction)
        for gk in guesses:
            shard.consensus(gk)
        vals, vecs = davidson(lambda x, y: sharded.apply(x, y, -1.0), guesses, n, ctx.code, ctx.device, nroots=nS,
                              lindep=lindep, res_tol=res_tol, fixed_cycles=fixed_cycles, stats=stats, plan=None,
                              native=False, max_cycle=max_cycle, consensus=shard.consensus, max_space=max_space,
                              restart_keep=restart_keep, stagnation=stagnation, **pc)
        E = -vals
        if stats is not None:
            stats["n_sharded_solves"] = stats.get("n_sharded_solves", 0) + 1
    elif alg == "davidson":
        vals, vecs = davidson(lambda x, y: plan.apply(x, y, -1.0), guesses, n, ctx.code, ctx.device, nroots=nS,
                              lindep=lindep, res_tol=res_tol, fixed_cycles=fixed_cycles, stats=stats, plan=plan,
                              native=native, max_cycle=max_cycle, max_space=max_space, restart_keep=restart_keep,
                              stagnation=stagnation, **pc)
        E = -vals                                               # diag_tools.py:265,273
    elif alg == "exact":
        E, vecs = _eigs_exact(plan, n, nS, ctx)
    elif alg == "arnoldi":
        ncv = max(2 * nS + 1, 20)                               # scipy.sparse.linalg.eigs default
        vals, vecs = davidson(lambda x, y: plan.apply(x, y, -1.0), guesses[:1] if nS == 1 else guesses, n, ctx.code, ctx.device,
                              nroots=nS, lindep=lindep, res_tol=arnoldi_tol, tol_relative=True, stats=stats, plan=plan, native=native,
                              max_cycle=max_cycle, max_space=max(ncv - 3 * nS, nS + 2), restart_keep=min(2 * nS + 2, ncv - nS - 1))
        E = -vals
    elif alg == "arnoldi_host":
        E, vecs = _eigs_arnoldi(plan, n, nS, guesses[0], ctx)
    else:
        raise ValueError("alg must be 'davidson', 'exact', 'arnoldi' or 'arnoldi_host'")
    if shard is not None and shard.world > 1:
        # EVERY solve of a sharded run -- also the small bonds every rank solves on its own -- ends with rank 0's
        # eigenpairs on all ranks: a rounding difference that flips one convergence test would otherwise leave the
        # ranks with states that differ at the solver tolerance, and the next sharded mat-vec would mix blocks
        # built from different bases.  E steers checkConv, i.e. the number of sweeps (= collectives): same on all.
        vecs = [shard.consensus(vk.contiguous()) for vk in vecs]
        E = shard.consensus_host(np.asarray(E, dtype=np.complex128), ctx)
    if not isinstance(orthonormalize, bool):                    # diag_tools.py:275-283
        orthonormalize = (nS > 1) and (abs(E[0] - E[1]) < orthonormalize)
    if nS > 1 and orthonormalize:
        vecs = _orth(vecs, ctx)
    N = len(mpsL[0])
    if (site == 0 or site == N - 1) and edgePreserveState:
        preserveState = True
    if oneSite or alg != "exact":                               # diag_tools.py:208-211
        E, vecs, ovlp = check_overlap(guesses[0], list(vecs), E, preserveState, ctx)
    else:
        ovlp = None
    E = np.asarray(E)
    Eout = E[0] if (nStates == 1 and E.ndim == 1) else E
    V = torch.stack(list(vecs), dim=1)
    if own_plan:
        plan.close()
    if sharded is not None:
        sharded.plan.close()
    return Eout, (V if dev_in else V.cpu().numpy()), ovlp


def _eigs_exact(plan, n, k, ctx):
    """calc_eigs_exact: dense H (built column by column with the device mat-vec, which is
    what diag_tools.py:30-63 contracts explicitly) + LAPACK geev on the host (:197)."""
    if n > 4096:
        raise ValueError("alg='exact' is a validation path for small local problems (n <= 4096)")
    eye = torch.eye(n, dtype=ctx.tdtype, device=ctx.device)
    H = torch.empty((n, n), dtype=ctx.tdtype, device=ctx.device)
    for c in range(n):
        plan.apply(eye[c], H[c], 1.0)                           # row c of H^T
    Hh = H.cpu().numpy().T
    vals, vv = sla.eig(Hh)
    order = np.argsort(vals)[::-1][:k]                          # :198-200 (descending)
    E = vals[order]
    vecs = [ctx.dev(np.ascontiguousarray(vv[:, o]) if ctx.code == D.C128 else np.real(vv[:, o] * np.exp(-1j * np.angle(vv[np.argmax(np.abs(vv[:, o])), o]))))
            for o in order]
    return E, vecs


def _eigs_arnoldi(plan, n, k, guess, ctx):
    """calc_eigs_arnoldi (:214-241): ARPACK on the host driving the device mat-vec."""
    from scipy.sparse.linalg import LinearOperator, eigs as arnoldi
    xd = torch.empty(n, dtype=ctx.tdtype, device=ctx.device)
    yd = torch.empty_like(xd)

    def mv(x):
        xd.copy_(ctx.dev(np.asarray(x).reshape(-1)))
        plan.apply(xd, yd, -1.0)
        return yd.cpu().numpy()

    k = min(k, n - 2)
    op = LinearOperator((n, n), matvec=mv, dtype=np.complex128)
    try:
        vals, vv = arnoldi(op, k=k, which="SR", v0=guess.cpu().numpy().astype(np.complex128), tol=1e-5)
    except Exception as exc:                                    # :227-231
        vals, vv = exc.eigenvalues, exc.eigenvectors
    order = np.argsort(vals)[:k]
    E = -vals[order]
    if ctx.code == D.F64:
        vv = np.real(vv * np.exp(-1j * np.angle(vv[np.argmax(np.abs(vv), axis=0), np.arange(vv.shape[1])])))
    return E, [ctx.dev(np.ascontiguousarray(vv[:, o])) for o in order]


def _orth(vecs, ctx):
    """Orthonormal basis of span(vecs) (sla.orth, diag_tools.py:285) by device Gram-Schmidt."""
    from .davidson import _orthonormalise
    n = vecs[0].numel()
    ops = D.VecOps(ctx.code, n, ctx.device)
    tmp = torch.zeros(8, dtype=ctx.tdtype, device=ctx.device)
    nrm = torch.zeros(1, dtype=torch.float64, device=ctx.device)
    out = [v.clone() for v in vecs]
    return _orthonormalise(ops, out, out, nrm, tmp)
