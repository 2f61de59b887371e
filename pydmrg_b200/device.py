"""Device-side operator layer: torch tensors in, C-ABI calls out.

PyTorch owns device memory and the CUDA stream; every FLOP of the hot path is executed by the
hand-written sm_100a kernels in ``libpydmrg_b200.so`` (plus cuSOLVER for the SVD / eigh
back-ends).  Nothing here falls back to torch or NumPy arithmetic.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import C128, F64, check, lib


_DEV_INDEX = None


def _stream():
    """Raw cudaStream_t of torch's current stream (the fast private accessor: this is called once
    per kernel launch and torch.cuda.current_stream() costs ~15 us)."""
    global _DEV_INDEX
    if _DEV_INDEX is None:
        _DEV_INDEX = torch.cuda.current_device()
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(_DEV_INDEX))
    except AttributeError:          # private API moved: fall back to the public one
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_device_index(index):
    """Call after torch.cuda.set_device() if the device changes within a process."""
    global _DEV_INDEX
    _DEV_INDEX = int(index)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def dtype_code(t):
    if t.dtype == torch.complex128:
        return C128
    if t.dtype == torch.float64:
        return F64
    raise TypeError("pydmrg_b200 supports float64 / complex128 only, got %s" % t.dtype)


def torch_dtype(code):
    return torch.complex128 if code == C128 else torch.float64


def np_dtype(code):
    return np.complex128 if code == C128 else np.float64


def _req(t, name):
    if not (t.is_cuda and t.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA tensor" % name)


class Workspace:
    """Grow-only device scratch buffer (one per solver object; no per-call allocation)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self.buf


# ------------------------------------------------------------------------------------------
def gemm_nt(A, B, C=None, *, M=None, N=None, K=None, batch=1, lda=None, ldb=None, ldc=None,
            strideA=0, strideB=0, strideC=0, conjB=False, alpha=1.0, beta=0):
    """C[b] = alpha * A[b] @ op(B[b]).T + beta * C[b]; see pdmrg_gemm_nt in include/pydmrg_b200.h.
    With only A (M,K) and B (N,K) given, the extents are taken from the 2-D tensors."""
    code = dtype_code(A)
    if M is None:
        M, K = A.shape
        N = B.shape[0]
        lda, ldb = K, K
    if C is None:
        C = torch.empty((batch, M, N) if batch > 1 else (M, N), dtype=A.dtype, device=A.device)
        ldc = N
        strideC = M * N
    if ldc is None:
        ldc = N
    alpha = complex(alpha)
    check(lib.pdmrg_gemm_nt(code, M, N, K, batch, _ptr(A), lda, strideA, _ptr(B), ldb, strideB, int(conjB),
                            _ptr(C), ldc, strideC, alpha.real, alpha.imag, int(beta), _stream()), "pdmrg_gemm_nt")
    return C


def permute4(src, out, dims, strides, conj=False, scale=None, scale_axis=0):
    code = dtype_code(src)
    n = list(dims) + [1] * (4 - len(dims))
    s = list(strides) + [0] * (4 - len(strides))
    check(lib.pdmrg_permute4(code, _ptr(src), _ptr(out), n[0], n[1], n[2], n[3], s[0], s[1], s[2], s[3],
                             int(conj), _ptr(scale), scale_axis, _stream()), "pdmrg_permute4")
    return out


# ------------------------------------------------------------------------------------------
def combine_onesite(W, wl, wr, d, np_dt):
    """Wc[(n,o),(j,l)] = W[n,j,o,l]  (tools/diag_tools.py:118); None = identity."""
    if W is None:
        W = np.einsum("ab,ij->abij", np.eye(wl), np.eye(d))
    return np.ascontiguousarray(np.transpose(W, (0, 2, 1, 3)).reshape(wl * d, wr * d).astype(np_dt))


def combine_twosite(W1, W2, np_dt):
    """Wc[(j,(m,p)),(o,(n,q))] = sum_l W1[j,l,m,n] W2[l,o,p,q]  (tools/diag_tools.py:154-155)."""
    wl, _, d1, _ = W1.shape
    _, wr, d2, _ = W2.shape
    Wc = np.einsum("jlmn,lopq->jmponq", W1, W2).reshape(wl * d1 * d2, wr * d1 * d2)
    return np.ascontiguousarray(Wc.astype(np_dt))


class HeffPlan:
    """y = sign * H_eff x for one local problem (pdmrg_heff_* in the C ABI).

    ``ops`` is a list over MPO terms of (L, Wc, R): L (Dl,wL,Dl), R (Dr,wR,Dr) CUDA tensors
    and Wc the combined local operator (host ndarray, (wL*P, wR*P))."""

    def __init__(self, code, P, Dl, Dr, ops, workspace, block_masks=None, k_range=None):
        self.code, self.P, self.Dl, self.Dr = code, P, Dl, Dr
        self.n = P * Dl * Dr
        self.Ls = [o[0] for o in ops]
        self.Rs = [o[2] for o in ops]
        for t in self.Ls + self.Rs:
            _req(t, "environment block")
            if dtype_code(t) != code:
                raise TypeError("environment dtype does not match the plan")
        n_mpo = len(ops)
        self._wc = [np.ascontiguousarray(o[1], dtype=np_dtype(code)) for o in ops]
        wL = (ctypes.c_int * n_mpo)(*[int(L.shape[1]) for L in self.Ls])
        wR = (ctypes.c_int * n_mpo)(*[int(R.shape[1]) for R in self.Rs])
        for k, (L, wc, R) in enumerate(ops):
            if tuple(L.shape) != (Dl, wL[k], Dl) or tuple(R.shape) != (Dr, wR[k], Dr):
                raise ValueError("environment shape mismatch: L %s R %s for Dl=%d Dr=%d" % (tuple(L.shape), tuple(R.shape), Dl, Dr))
            if self._wc[k].shape != (wL[k] * P, wR[k] * P):
                raise ValueError("combined operator shape %s != %s" % (self._wc[k].shape, (wL[k] * P, wR[k] * P)))
        wc_ptrs = (ctypes.c_void_p * n_mpo)(*[w.ctypes.data for w in self._wc])
        self._plan = ctypes.c_void_p(0)
        mptrs = None
        if block_masks is not None:
            self._masks = [np.ascontiguousarray(m, dtype=np.uint8) for m in block_masks]
            mptrs = (ctypes.c_void_p * n_mpo)(*[m.ctypes.data for m in self._masks])
        k0, klen = (0, -1) if k_range is None else (int(k_range[0]), int(k_range[1]))
        # host tables now, device blob owned by torch, the (tiny) copy ordered on the current stream: no cudaMalloc /
        # cudaFree / legacy-stream synchronisation per bond step (they would stall concurrent sweeps on other streams)
        check(lib.pdmrg_heff_plan_create_host(ctypes.byref(self._plan), code, P, Dl, Dr, n_mpo, wL, wR, wc_ptrs, mptrs, k0, klen),
              "pdmrg_heff_plan_create_host")
        nb = int(lib.pdmrg_heff_plan_blob_bytes(self._plan))
        self._blob = torch.empty(nb, dtype=torch.uint8, device=self.Ls[0].device)
        check(lib.pdmrg_heff_plan_upload(self._plan, _ptr(self._blob), nb, _stream()), "pdmrg_heff_plan_upload")
        self._Lp = (ctypes.c_void_p * n_mpo)(*[t.data_ptr() for t in self.Ls])
        self._Rp = (ctypes.c_void_p * n_mpo)(*[t.data_ptr() for t in self.Rs])
        self.ws_bytes = int(lib.pdmrg_heff_workspace_bytes(self._plan))
        self.workspace = workspace
        self.flops_dense = float(lib.pdmrg_heff_flops_dense(self._plan))
        self.flops_sparse = float(lib.pdmrg_heff_flops_sparse(self._plan))
        self.flops_gemm = float(lib.pdmrg_heff_flops_gemm(self._plan))
        self.n_apply = 0

    def apply(self, x, y, sign=-1.0):
        """y <- sign * H x  (the reference's Hfun returns -H x: tools/diag_tools.py:125,162)."""
        ws = self.workspace.get(self.ws_bytes)
        check(lib.pdmrg_heff_apply(self._plan, self._Lp, self._Rp, _ptr(x), _ptr(y), float(sign), _ptr(ws),
                                   ws.numel(), _stream()), "pdmrg_heff_apply")
        self.n_apply += 1
        return y

    def diag(self, sign=-1.0, out=None):
        """Diagonal of  sign * H_eff  (pdmrg_heff_diag), for the diagonal Davidson preconditioner."""
        P, Dl, Dr = self.P, self.Dl, self.Dr
        if out is None:
            out = torch.empty(self.n, dtype=self.Ls[0].dtype, device=self.Ls[0].device)
        for m, (L, R, wc) in enumerate(zip(self.Ls, self.Rs, self._wc)):
            wL, wR = int(L.shape[1]), int(R.shape[1])
            W4 = wc.reshape(wL, P, wR, P)
            Wpp = np.ascontiguousarray(np.einsum("jpop->pjo", W4))
            Wd = torch.from_numpy(Wpp).to(L.device)
            check(lib.pdmrg_heff_diag(self.code, P, Dl, Dr, wL, wR, _ptr(L), _ptr(R), _ptr(Wd), float(sign), int(m > 0),
                                      _ptr(out), _stream()), "pdmrg_heff_diag")
            self._keep = Wd                                     # alive until the kernel has run (stream-ordered free)
        return out

    def apply_timed(self, x, y, sign=-1.0):
        """apply() with per-stage CUDA-event timing; returns [ms stage A, ms mix, ms stage C]."""
        ws = self.workspace.get(self.ws_bytes)
        ms = (ctypes.c_float * 3)()
        check(lib.pdmrg_heff_apply_timed(self._plan, self._Lp, self._Rp, _ptr(x), _ptr(y), float(sign), _ptr(ws),
                                         ws.numel(), _stream(), ms), "pdmrg_heff_apply_timed")
        return [float(v) for v in ms]

    def close(self):
        if self._plan:
            lib.pdmrg_heff_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
def _w_host(W, code):
    if W is None:
        return None, ctypes.c_void_p(0)
    Wh = np.ascontiguousarray(W, dtype=np_dtype(code))
    return Wh, ctypes.c_void_p(Wh.ctypes.data)


def env_update_right(A, W, L, Abra, workspace):
    """tools/env_tools.py:40-50 on the device; returns the new left block (Dr, wr, Dr)."""
    code = dtype_code(A)
    d, Dl, Dr = A.shape
    wl = L.shape[1]
    wr = wl if W is None else W.shape[1]
    Wh, Wp = _w_host(W, code)
    out = torch.empty((Dr, wr, Dr), dtype=A.dtype, device=A.device)
    nb = lib.pdmrg_env_workspace_bytes(code, d, Dl, Dr, wl, wr)
    ws = workspace.get(nb)
    check(lib.pdmrg_env_update_right(code, d, Dl, Dr, wl, wr, Wp, _ptr(L), _ptr(A), _ptr(Abra), _ptr(out),
                                     _ptr(ws), ws.numel(), _stream()), "pdmrg_env_update_right")
    return out


def env_update_left(B, W, R, Bbra, workspace):
    """tools/env_tools.py:28-38 on the device; returns the new right block (Dl, wl, Dl)."""
    code = dtype_code(B)
    d, Dl, Dr = B.shape
    wr = R.shape[1]
    wl = wr if W is None else W.shape[0]
    Wh, Wp = _w_host(W, code)
    out = torch.empty((Dl, wl, Dl), dtype=B.dtype, device=B.device)
    nb = lib.pdmrg_env_workspace_bytes(code, d, Dl, Dr, wl, wr)
    ws = workspace.get(nb)
    check(lib.pdmrg_env_update_left(code, d, Dl, Dr, wl, wr, Wp, _ptr(R), _ptr(B), _ptr(Bbra), _ptr(out),
                                    _ptr(ws), ws.numel(), _stream()), "pdmrg_env_update_left")
    return out


def env_update_cols(direction, T, W, blk, Tbra, k0, kn, workspace, out=None):
    """The slice of a block update owned by one rank of a sharded sweep (csrc/env_cols.cu): the entries of the new
    block whose ket index lies in [k0, k0+kn), written in place into the full-size ``out`` (zero-initialised when
    not given).  ``direction`` 'right': new left block from ``blk`` = L and T = A; 'left': new right block from
    ``blk`` = R and T = B."""
    code = dtype_code(T)
    d, Dl, Dr = T.shape
    Wh, Wp = _w_host(W, code)
    if direction == "right":
        wl = blk.shape[1]
        wr = wl if W is None else W.shape[1]
        shape, fn, name = (Dr, wr, Dr), lib.pdmrg_env_update_right_cols, "pdmrg_env_update_right_cols"
    else:
        wr = blk.shape[1]
        wl = wr if W is None else W.shape[0]
        shape, fn, name = (Dl, wl, Dl), lib.pdmrg_env_update_left_cols, "pdmrg_env_update_left_cols"
    if out is None:
        out = torch.zeros(shape, dtype=T.dtype, device=T.device)
    assert tuple(out.shape) == shape and out.is_contiguous()
    nb = lib.pdmrg_env_workspace_bytes(code, d, Dl, Dr, wl, wr)
    ws = workspace.get(nb)
    check(fn(code, d, Dl, Dr, wl, wr, Wp, _ptr(blk), _ptr(T), _ptr(Tbra), int(k0), int(kn), _ptr(out), _ptr(ws), ws.numel(),
             _stream()), name)
    return out


# ------------------------------------------------------------------------------------------
class VecOps:
    """Krylov vector kernels bound to one dtype / vector length / scratch buffer."""

    def __init__(self, code, n, device):
        self.code, self.n = code, int(n)
        self.dt = torch_dtype(code)
        self.npdt = np_dtype(code)
        nb = int(lib.pdmrg_vec_scratch_bytes(64))
        self.scratch = torch.zeros(nb, dtype=torch.uint8, device=device)      # holds the reduction ticket: must start at zero
        self.norm2 = torch.zeros(2, dtype=torch.float64, device=device)
        self.coefs = torch.zeros(128, dtype=self.dt, device=device)

    def _host(self, a):
        a = np.asarray(a)
        if self.code == F64 and np.iscomplexobj(a):
            a = a.real
        return np.ascontiguousarray(a, dtype=self.npdt)

    def dots(self, X, m, y, out=None):
        """out[i] = <X_i|y>, i < m;  X is a (rows, n) slab.  Result stays on the device."""
        out = self.coefs if out is None else out
        check(lib.pdmrg_vec_dots(self.code, self.n, m, _ptr(X), X.stride(0), _ptr(y), _ptr(out), _ptr(self.scratch),
                                 self.scratch.numel(), _stream()), "pdmrg_vec_dots")
        return out

    def combine(self, coef, X, m, out, accumulate=False):
        c = self._host(coef)
        check(lib.pdmrg_vec_combine(self.code, self.n, m, ctypes.c_void_p(c.ctypes.data), _ptr(X), X.stride(0),
                                    _ptr(out), int(accumulate), _stream()), "pdmrg_vec_combine")
        return out

    def residual(self, v, e, XS, AX, m, r, norm2_out=None):
        norm2_out = self.norm2 if norm2_out is None else norm2_out
        vv = self._host(v)
        ee = self._host([e])
        check(lib.pdmrg_davidson_residual(self.code, self.n, m, ctypes.c_void_p(vv.ctypes.data),
                                          ctypes.c_void_p(ee.ctypes.data), _ptr(XS), _ptr(AX), XS.stride(0), _ptr(r),
                                          _ptr(norm2_out), _ptr(self.scratch), self.scratch.numel(), _stream()),
              "pdmrg_davidson_residual")
        return norm2_out

    def project(self, X, m, c_dev, r, scale_dev=None, norm2_out=None):
        norm2_out = self.norm2[1:] if norm2_out is None else norm2_out
        check(lib.pdmrg_vec_project(self.code, self.n, m, _ptr(X), X.stride(0), _ptr(c_dev), _ptr(scale_dev), _ptr(r),
                                    _ptr(norm2_out), _ptr(self.scratch), self.scratch.numel(), _stream()),
              "pdmrg_vec_project")
        return norm2_out

    def precond(self, r, diag, theta, floor, norm2_out):
        """r <- r / (diag - theta) (floored); norm2_out[0] = ||r||^2 of the result."""
        th = complex(theta)
        check(lib.pdmrg_vec_precond(self.code, self.n, _ptr(r), _ptr(diag), th.real, th.imag, float(floor), _ptr(norm2_out),
                                    _ptr(self.scratch), self.scratch.numel(), _stream()), "pdmrg_vec_precond")
        return norm2_out

    def normalize(self, src, norm2_dev, out):
        check(lib.pdmrg_vec_normalize(self.code, self.n, _ptr(src), _ptr(norm2_dev), _ptr(out), _stream()),
              "pdmrg_vec_normalize")
        return out


# ------------------------------------------------------------------------------------------
SVD_GESVD = 0
SVD_GESVDJ = 1
SVD_GRAM = 2


def svd(a, workspace, method=SVD_GESVD, keep=0):
    """Thin SVD of a row-major (rows, cols) CUDA matrix (destroyed).  Returns U, S, Vh, info.
    ``keep`` > 0: only the leading ``keep`` triplets are needed (lets the Gram route skip work)."""
    code = dtype_code(a)
    rows, cols = a.shape
    k = min(rows, cols)
    U = torch.empty((rows, k), dtype=a.dtype, device=a.device)
    Vh = torch.empty((k, cols), dtype=a.dtype, device=a.device)
    S = torch.empty(k, dtype=torch.float64, device=a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    nb = int(lib.pdmrg_svd_workspace_bytes(code, rows, cols, method))
    if nb == 0:
        check(1, "pdmrg_svd_workspace_bytes")
    ws = workspace.get(nb)
    check(lib.pdmrg_svd_topk(code, rows, cols, int(keep), _ptr(a), _ptr(U), _ptr(S), _ptr(Vh), method, _ptr(ws),
                             ws.numel(), _ptr(info), _stream()), "pdmrg_svd")
    return U, S, Vh, info


def left_basis(X, workspace, keep=0):
    """Refined left singular basis of X (n, c): returns (Ut (n,n) rows = u_a descending, lam (n,) = S^2,
    info) or None when cuSOLVER's eigensolver did not converge."""
    code = dtype_code(X)
    n, c = X.shape
    Ut = torch.empty((n, n), dtype=X.dtype, device=X.device)
    lam = torch.empty(n, dtype=torch.float64, device=X.device)
    info = torch.zeros(1, dtype=torch.int32, device=X.device)
    nb = int(lib.pdmrg_left_basis_workspace_bytes(code, n, c))
    if nb == 0:
        check(1, "pdmrg_left_basis_workspace_bytes")
    ws = workspace.get(nb)
    rc = lib.pdmrg_left_basis(code, n, c, int(keep), _ptr(X), _ptr(Ut), _ptr(lam), _ptr(ws), ws.numel(), _ptr(info), _stream())
    if rc == 77:
        return None
    check(rc, "pdmrg_left_basis")
    return Ut, lam, info


RECYCLE_HOUSEHOLDER = 0
RECYCLE_CHOLQR2 = 1
RECYCLE_MAX_DEFECT2 = 1e-2       # ||P1 P1^H - I||_F^2 tolerated after the first Cholesky pass (the second squares it)


def recycle_basis(X1, X2, Wt, k, b, workspace, method=None):
    """One warm-started subspace step (csrc/solver.cu pdmrg_recycle_basis): X1 (a,c), X2 = X1^T (c,a),
    Wt (K,c) previous basis rows (k kept + K-k spare).  Returns Pt (K,a), Ct (K,c), S (K,) on the HOST,
    ok, method_used.  ``method`` None: Cholesky-QR2 first, Householder QR when its first pass was not
    well conditioned (or cuSOLVER reported a failure)."""
    code = dtype_code(X1)
    a, c = X1.shape
    K = Wt.shape[0]
    assert X2.shape == (c, a) and Wt.shape[1] == c and X1.is_contiguous() and X2.is_contiguous() and Wt.is_contiguous()
    Pt = torch.empty((K, a), dtype=X1.dtype, device=X1.device)
    Ct = torch.empty((K, c), dtype=X1.dtype, device=X1.device)
    S = torch.empty(K + 2, dtype=torch.float64, device=X1.device)
    info = torch.zeros(4, dtype=torch.int32, device=X1.device)
    nb = int(lib.pdmrg_recycle_basis_workspace_bytes(code, a, c, K, int(k), int(b)))
    if nb == 0:
        check(1, "pdmrg_recycle_basis_workspace_bytes")
    ws = workspace.get(nb)
    Sh = None
    for meth in ([RECYCLE_CHOLQR2, RECYCLE_HOUSEHOLDER] if method is None else [method]):
        check(lib.pdmrg_recycle_basis(code, a, c, K, int(k), int(b), int(meth), _ptr(X1), _ptr(X2), _ptr(Wt), _ptr(Pt),
                                      _ptr(Ct), _ptr(S), _ptr(ws), ws.numel(), _ptr(info), _stream()), "pdmrg_recycle_basis")
        Sh = S.cpu().numpy()                                    # sync
        ok = bool(np.all(np.isfinite(Sh))) and Sh[K + 1] == 0 and Sh[K] < RECYCLE_MAX_DEFECT2
        if ok:
            return Pt, Ct, Sh[:K], True, meth
    return Pt, Ct, Sh[:K], False, meth


def heevd(a, workspace):
    """In-place Hermitian eigen-decomposition of a row-major (n,n) matrix; returns (w, a)
    with row i of ``a`` the i-th eigenvector (ascending)."""
    code = dtype_code(a)
    n = a.shape[0]
    w = torch.empty(n, dtype=torch.float64, device=a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    nb = int(lib.pdmrg_heevd_workspace_bytes(code, n))
    if nb == 0:
        check(1, "pdmrg_heevd_workspace_bytes")
    ws = workspace.get(nb)
    check(lib.pdmrg_heevd(code, n, _ptr(a), _ptr(w), _ptr(ws), ws.numel(), _ptr(info), _stream()), "pdmrg_heevd")
    return w, a, info
