"""Device-resident restarted non-symmetric Davidson.

Mirrors the algorithm the reference runs for every local eigenproblem
(``davidson_nosym1``, pydmrg/tools/la_tools.py:402-557, called from
pydmrg/tools/diag_tools.py:260 with the identity preconditioner, ``pick_eigs``,
``tol=1e-16, max_cycle=1000, follow_state=False``): same subspace growth, same restart rule
(``space + nroots > max_space``, la_tools.py:544, with max_space = 12 + 3*nroots, :412), same
linear-dependence pruning (``lindep``, :505-537) -- which is what actually terminates the
reference's loop, because its convergence flags are never set (:485-488).

What is different is where the work happens: the basis ``xs`` and its images ``ax`` are two
HBM-resident (max_space+nroots, n) slabs, the mat-vec is the DMMA H_eff plan, and every
vector operation is one fused pass of a CUDA kernel (csrc/krylov.cu).  The host only solves
the (<= 15+) x (<= 15+) projected eigenproblem with ``scipy.linalg.eig`` -- exactly the call the
reference makes (la_tools.py:461) -- and reads back O(m) scalars twice per cycle.

The in-loop orthogonalisation is block classical Gram-Schmidt (one multi-dot pass + one
multi-axpy pass) instead of the reference's sequential modified Gram-Schmidt (:524-527);
both project the same residual against the same orthonormal basis and agree to rounding.
"""
import ctypes
import os

import numpy as np
import scipy.linalg as sla
import torch

from . import device as D
from ._lib import check, lib


def pick_lowest_real(w, v):
    """pick_eigs, pydmrg/tools/diag_tools.py:175-179: ascending real part -- with a tie-break the reference lacks: among
    (numerically) equal real parts, i.e. the two members of a complex-conjugate pair, the larger imaginary part comes
    first, so the target cannot flip between them from cycle to cycle (csrc/davidson.cu)."""
    idx = list(np.argsort(np.real(w), kind="stable"))
    for _ in range(2):
        for i in range(len(idx) - 1):
            a, b = w[idx[i]], w[idx[i + 1]]
            if abs(a.real - b.real) <= 1e-9 * max(1.0, abs(a), abs(b)) and a.imag < b.imag:
                idx[i], idx[i + 1] = idx[i + 1], idx[i]
    idx = np.asarray(idx)
    return w[idx], v[:, idx]


class DavidsonWorkspace:
    """Basis slabs + scratch for vectors of length <= n (re-used across sites)."""

    def __init__(self, code, n, rows, nroots, device):
        self.code, self.n, self.rows, self.nroots = code, n, rows, nroots
        dt = D.torch_dtype(code)
        self.XS = torch.empty((rows, n), dtype=dt, device=device)
        self.AX = torch.empty((rows, n), dtype=dt, device=device)
        self.R = torch.empty((2 * nroots + 2, n), dtype=dt, device=device)   # residual / Ritz vectors
        self.small = torch.zeros(2 * rows * (nroots + 2) + 16, dtype=dt, device=device)
        self.norms = torch.zeros(2 * nroots + 16, dtype=torch.float64, device=device)
        # scratch of the C++ loop for the largest subspace these slabs can hold (rows = max_space + 4 * nroots)
        self.native_scratch = torch.zeros(int(lib.pdmrg_davidson_scratch_bytes(nroots, max(12, rows - 4 * nroots))) + 256,
                                          dtype=torch.uint8, device=device)
        self._ops = {}

    def ops(self, n):
        o = self._ops.get(n)
        if o is None:
            if len(self._ops) > 8:
                self._ops.clear()
            o = D.VecOps(self.code, n, self.XS.device)
            self._ops[n] = o
        return o


_WS_CACHE = {}


def get_workspace(code, n, rows, nroots, device):
    import threading
    key = (code, str(device), threading.get_ident())            # one set of slabs per host thread (scan workers)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.n < n or ws.rows < rows or ws.nroots < nroots:
        _WS_CACHE.pop(key, None)
        ws = None     # free the old slabs before allocating the new ones
        ws = DavidsonWorkspace(code, n, rows, nroots, device)
        _WS_CACHE[key] = ws
    return ws


def release_workspaces():
    _WS_CACHE.clear()


def _orthonormalise(ops, src, dst, norm_slot, small):
    """Gram-Schmidt of a short list of device vectors (la_tools.py:586-598) into ``dst``."""
    out = []
    for s, d in zip(src, dst):
        if d.data_ptr() != s.data_ptr():
            d.copy_(s)
        if not out:
            ops.project(d.unsqueeze(0), 0, small, d, scale_dev=None, norm2_out=norm_slot)   # m = 0: norm only
        for u in out:
            c = ops.dots(u.unsqueeze(0), 1, d, out=small)
            ops.project(u.unsqueeze(0), 1, c, d, scale_dev=None, norm2_out=norm_slot)
        ops.normalize(d, norm_slot, d)
        out.append(d)
    return out


def davidson_native(plan, x0, n, code, device, nroots=1, max_space=12, lindep=1e-14, max_cycle=1000,
                    res_tol=None, fixed_cycles=None, stats=None, sign=-1.0, diag=None, precond_floor=0.1,
                    restart_keep=0, tol_relative=False, stagnation=0, comm=None):
    """The same solver sequenced in C++ (csrc/davidson.cu, ``pdmrg_davidson``): one C-ABI call per
    local eigenproblem, operator = ``sign`` * H_eff of ``plan``.  ``comm`` (a pdmrg_comm handle): ``plan`` is this rank's
    ket-split shard and every mat-vec is the fused peer-memory exchange (``pdmrg_davidson_sharded``)."""
    ms = max_space + 3 * nroots
    rows = ms + nroots + int(restart_keep)
    W = get_workspace(code, n, rows, max(nroots, len(x0)), device)
    for k, g in enumerate(x0):
        W.R[nroots + k][:n].copy_(g)
    out = torch.empty((nroots, n), dtype=W.XS.dtype, device=device)
    e_host = (ctypes.c_double * (2 * nroots))()
    n_mv, cyc, last = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
    ws = plan.workspace.get(plan.ws_bytes)
    args = (plan._plan, plan._Lp, plan._Rp, float(sign), code, n, D._ptr(W.R[nroots]), W.R.stride(0),
            len(x0), nroots, max_space, float(lindep), -1.0 if res_tol is None else float(res_tol),
            int(max_cycle), int(fixed_cycles or 0), D._ptr(W.XS), D._ptr(W.AX), D._ptr(W.R),
            W.XS.stride(0), D._ptr(ws), ws.numel(), D._ptr(W.native_scratch), W.native_scratch.numel(),
            e_host, D._ptr(out), out.stride(0), ctypes.byref(n_mv), ctypes.byref(cyc), ctypes.byref(last), D._stream())
    n_rs = ctypes.c_int(0)
    if comm is not None:
        check(lib.pdmrg_davidson_sharded(args[0], comm, *args[1:], D._ptr(diag), float(precond_floor), int(restart_keep),
                                         int(bool(tol_relative)), ctypes.byref(n_rs), int(stagnation)), "pdmrg_davidson_sharded")
    elif restart_keep or tol_relative or stagnation:
        check(lib.pdmrg_davidson_ex(*args, D._ptr(diag), float(precond_floor), int(restart_keep), int(bool(tol_relative)),
                                    ctypes.byref(n_rs), int(stagnation)), "pdmrg_davidson_ex")
    elif diag is None:
        check(lib.pdmrg_davidson(*args), "pdmrg_davidson")
    else:
        check(lib.pdmrg_davidson_precond(*args, D._ptr(diag), float(precond_floor)), "pdmrg_davidson_precond")
    plan.n_apply += n_mv.value
    nr = min(nroots, max(1, min(nroots, n)))
    e = np.array([complex(e_host[2 * k], e_host[2 * k + 1]) for k in range(nr)])
    if stats is not None:
        stats["n_matvec"] = stats.get("n_matvec", 0) + n_mv.value
        stats["cycles"] = stats.get("cycles", 0) + cyc.value
        stats["last_residual"] = last.value
        if restart_keep:
            stats["n_thick_restarts"] = stats.get("n_thick_restarts", 0) + n_rs.value
    return e, [out[k] for k in range(nr)]


def davidson(matvec, x0, n, code, device, nroots=1, max_space=12, lindep=1e-14, max_cycle=1000,
             res_tol=None, fixed_cycles=None, stats=None, plan=None, native=None, consensus=None, diag=None,
             precond_floor=0.1, restart_keep=0, tol_relative=False, stagnation=0, comm=None):
    """Eigenpairs with the smallest real part of the operator behind ``matvec(x, y)``
    (y <- A x, both device vectors of length n).

    x0: list of device vectors (initial guesses).  Returns (e ndarray (k,), [k device vectors]).
    ``res_tol``: optional residual-norm stop; default is the reference's effective criterion,
    ||r||^2 <= lindep.  ``fixed_cycles``: stop after exactly that many cycles (benchmarking).
    When the H_eff ``plan`` is given the loop runs in C++ (``pdmrg_davidson``) unless
    ``native=False`` / PDMRG_PY_DAVIDSON=1; this Python loop is the specification it mirrors.
    ``diag``: diagonal of the operator (device vector) -- switches the correction step from the reference's identity
    preconditioner (diag_tools.py:137-139) to t = r / (diag - theta), |diag - theta| >= precond_floor (absolute):
    same eigenpairs, fewer cycles in cold sweeps; opt-in because the iteration counts then differ from the reference.
    ``restart_keep`` > 0: thick restart -- at a restart (la_tools.py:544) the basis is compressed to an orthonormal basis
    of that many lowest Ritz vectors (XS <- XS Q, AX <- AX Q, projected matrix Q^H H Q; no new mat-vec) and the iteration
    continues with the current residual direction, instead of starting again from the Ritz vectors alone.  Nearly
    degenerate partners of the target survive the restart: orders of magnitude fewer mat-vecs when the gap closes (scan
    points next to s = 0).  With the identity preconditioner this is thick-restarted Arnoldi (Krylov-Schur).
    ``tol_relative``: ``res_tol`` is relative to |theta| (ARPACK's stopping rule).
    ``stagnation`` > 0: stop when the residual has not been halved within that many cycles (the reference goes on to
    max_cycle; nearly defective local problems next to s = 0 stall at ||r|| ~ 1e-4 whatever the subspace).
    ``consensus``: optional callable applied to the small device arrays right before the host reads them
    (sharded sweeps: a broadcast from rank 0, so that every rank takes the same control decisions and the
    collectives inside ``matvec`` stay matched whatever the last bit of a reduction does).
    """
    if not isinstance(x0, (list, tuple)):
        x0 = [x0]
    x0 = list(x0)
    if native is None:
        native = os.environ.get("PDMRG_PY_DAVIDSON", "0") != "1"
    if native and plan is not None and nroots <= 4 and len(x0) <= nroots and n > nroots:
        return davidson_native(plan, x0, n, code, device, nroots, max_space, lindep, max_cycle, res_tol, fixed_cycles, stats,
                               diag=diag, precond_floor=precond_floor, restart_keep=restart_keep, tol_relative=tol_relative,
                               stagnation=stagnation, comm=comm)
    if comm is not None:
        raise ValueError("the fused sharded mat-vec is driven by the C++ loop only (nroots <= 4, native=True)")
    max_space = max_space + 3 * nroots
    rows = max_space + nroots + int(restart_keep)
    W = get_workspace(code, n, rows, max(nroots, len(x0)), device)
    ops = W.ops(n)
    XS, AX = W.XS[:, :n], W.AX[:, :n]
    heff = np.zeros((rows, rows), dtype=np.complex128)
    thresh2 = lindep if res_tol is None else res_tol * res_tol
    small_tail = W.small[-8:]
    gs_norm = W.norms[-2:-1]

    use_pc, pc_best, pc_stall = diag is not None, float("inf"), 0
    stag_best, stag_count = float("inf"), 0
    n_mv = 0
    space = 0
    fresh = True
    e = v = None
    trial = []
    dx2 = np.zeros(1)
    cyc = 0
    for cyc in range(max_cycle):
        if fresh:
            space = 0
            trial = _orthonormalise(ops, x0, [XS[k] for k in range(len(x0))], gs_norm, small_tail)
        elif len(trial) > 1:
            trial = _orthonormalise(ops, trial, trial, gs_norm, small_tail)
        rnow = len(trial)
        head = space
        for k, tv in enumerate(trial):
            if tv.data_ptr() != XS[head + k].data_ptr():
                XS[head + k].copy_(tv)
            matvec(XS[head + k], AX[head + k])
            n_mv += 1
        space = head + rnow
        # projected matrix (la_tools.py:449-457): new columns <xs_i|A xt_k>, i < space, and new
        # rows <xt_k|A xs_i> = conj <A xs_i|xt_k>, i < head -- two multi-dot passes per new vector
        for k in range(rnow):
            ops.dots(W.XS, space, AX[head + k], out=W.small[k * 2 * rows:])
            if head > 0:
                ops.dots(W.AX, head, XS[head + k], out=W.small[k * 2 * rows + rows:])
        if consensus is not None:
            consensus(W.small[:rnow * 2 * rows])
        host = W.small[:rnow * 2 * rows].cpu().numpy()            # D2H #1 of the cycle
        for k in range(rnow):
            heff[:space, head + k] = host[k * 2 * rows: k * 2 * rows + space]
            if head > 0:
                heff[head + k, :head] = np.conj(host[k * 2 * rows + rows: k * 2 * rows + rows + head])
        w, vv = sla.eig(heff[:space, :space])                      # la_tools.py:461
        w, vv = pick_lowest_real(w, vv)
        e = w[:nroots].copy()
        v = vv[:, :nroots].copy()
        if code == D.F64:
            e, v = _force_real(e, v)
        nr = e.shape[0]
        if tol_relative and res_tol is not None:
            thresh2 = (res_tol * max(abs(e[0]), 2.3e-16)) ** 2
        # residual r_k = sum_i v_ik (A xs_i - e_k xs_i) and its norm (la_tools.py:479-482), then --
        # with no host round trip in between -- projection against the basis and the new norm
        for k in range(nr):
            ops.residual(v[:, k], e[k], W.XS, W.AX, space, W.R[k][:n], norm2_out=W.norms[2 * k:2 * k + 1])
            scale = W.norms[2 * k:2 * k + 1]
            if use_pc:                                             # la_tools.py:510 with a diagonal preconditioner
                scale = W.norms[2 * nroots + 8 + k:2 * nroots + 9 + k]
                ops.precond(W.R[k][:n], diag, e[k], precond_floor, scale)
            c = ops.dots(W.XS, space, W.R[k][:n])
            ops.project(W.XS, space, c, W.R[k][:n], scale_dev=scale, norm2_out=W.norms[2 * k + 1:2 * k + 2])
        if consensus is not None:
            consensus(W.norms[:2 * nr])
        nrm = W.norms[:2 * nr].cpu().numpy()                       # D2H #2 of the cycle
        dx2, px2 = nrm[0::2], nrm[1::2]
        if stats is not None:
            stats.setdefault("res_history", []).append(float(np.sqrt(dx2.max())))
        if use_pc:
            # safeguard (as in csrc/davidson.cu): no halving of the residual within two restart periods -> identity
            res_now = float(np.sqrt(dx2.max()))
            if res_now < 0.5 * pc_best:
                pc_best, pc_stall = res_now, 0
            else:
                pc_stall += 1
                use_pc = pc_stall <= 2 * max_space
        if stagnation:
            res_now = float(np.sqrt(dx2.max()))
            if res_now < 0.5 * stag_best:
                stag_best, stag_count = res_now, 0
            else:
                stag_count += 1
        trial = []
        for k in range(nr):
            if dx2[k] > thresh2 and px2[k] > lindep:               # la_tools.py:505-537
                ops.normalize(W.R[k][:n], W.norms[2 * k + 1:2 * k + 2], W.R[k][:n])
                trial.append(W.R[k][:n])
        if not trial:                                              # la_tools.py:539-541
            break
        if stagnation and stag_count >= stagnation:
            break
        if fixed_cycles is not None and cyc + 1 >= fixed_cycles:
            break
        fresh = space + nroots > max_space                         # la_tools.py:544
        if fresh and restart_keep:
            # thick restart: keep an orthonormal basis Q of the lowest Ritz vectors (see csrc/davidson.cu)
            kk = min(max(int(restart_keep), nr), space - 1)
            Y = np.concatenate([v, vv[:, nr:]], axis=1)
            if code == D.F64:
                wall = w
                cols = [c for c in range(Y.shape[1]) if c < nr or abs(wall[c].imag) <= 1e-9 * max(1.0, abs(wall[c].real))]
                Y = Y[:, cols]
                piv = np.argmax(np.abs(Y), axis=0)
                ph = Y[piv, np.arange(Y.shape[1])]
                Y = np.real(Y * (np.conj(ph) / np.abs(ph))).astype(np.complex128)
            Q = np.linalg.qr(Y[:, :kk])[0]
            kk = Q.shape[1]
            if kk >= nr:
                for j in range(kk):
                    ops.combine(Q[:, j], W.XS, space, XS[space + j])
                    ops.combine(Q[:, j], W.AX, space, AX[space + j])
                for j in range(kk):
                    XS[j].copy_(XS[space + j]); AX[j].copy_(AX[space + j])
                Hn = Q.conj().T @ heff[:space, :space] @ Q
                v = Q.conj().T @ v
                heff[:] = 0
                heff[:kk, :kk] = Hn
                space = kk
                fresh = False
                if stats is not None:
                    stats["n_thick_restarts"] = stats.get("n_thick_restarts", 0) + 1
        if fresh:
            x0 = [W.R[nr + k][:n] for k in range(nr)]
            for k in range(nr):
                ops.combine(v[:, k], W.XS, space, x0[k])
    out = []
    for k in range(e.shape[0]):
        o = torch.empty(n, dtype=W.XS.dtype, device=W.XS.device)
        ops.combine(v[:, k], W.XS, space, o)                       # Ritz vector, la_tools.py:469
        out.append(o)
    if stats is not None:
        stats["n_matvec"] = stats.get("n_matvec", 0) + n_mv
        stats["cycles"] = stats.get("cycles", 0) + cyc + 1
        stats["last_residual"] = float(np.sqrt(dx2.max()))
    return e, out


def _force_real(e, v):
    """f64 mode: the targeted Ritz pairs must be real (symmetric / Perron-Frobenius problems)."""
    scale = max(1.0, float(np.max(np.abs(e.real))))
    if np.max(np.abs(e.imag)) > 1e-9 * scale:
        raise ArithmeticError("f64 Davidson met a complex Ritz value %r; run with dtype='c128'" % (e,))
    piv = np.argmax(np.abs(v), axis=0)
    ph = v[piv, np.arange(v.shape[1])]
    v = v * (np.conj(ph) / np.abs(ph))
    return e.real.astype(np.complex128), np.real(v).astype(np.complex128)
