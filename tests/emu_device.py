"""Torch-CPU emulations of the device entry points (pydmrg_b200.device), following their documented semantics, so
that the HOST logic of the product (truncation bookkeeping, Davidson loop, sweeps, sharding) can execute without a
GPU.  Test infrastructure only; the CUDA kernels themselves are covered by the ``-m gpu`` tests."""
import numpy as np
import torch


class _Ctx:
    tdtype = torch.complex128
    device = torch.device("cpu")
    svd_method = -1
    svd_ws = None

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.tdtype)

    def svd_method_for(self, rows, cols):
        return 0


def _permute4(src, out, dims, strides, conj=False, scale=None, scale_axis=0):
    # pdmrg_permute4: the first prod(dims) elements of `out` receive src[sum_i idx_i * stride_i] (conjugated)
    if all(st >= 0 for st in strides):
        view = torch.as_strided(src.reshape(-1), tuple(dims), tuple(strides))
    else:
        # negative strides walk back from ``src`` (a view into a larger buffer, e.g. the last row of heevd's output)
        base = src.as_strided((src.untyped_storage().nbytes() // src.element_size(),), (1,), 0)
        idx = torch.full(tuple(dims), src.storage_offset(), dtype=torch.int64)
        for ax, (n, st) in enumerate(zip(dims, strides)):
            shp = [1] * len(dims)
            shp[ax] = n
            idx = idx + (torch.arange(n, dtype=torch.int64) * st).reshape(shp)
        assert int(idx.min()) >= 0
        view = base[idx]
    res = (view.conj() if conj else view).resolve_conj()
    if scale is not None:                      # out[..., i, ...] *= scale[i] along scale_axis
        shp = [1] * len(dims)
        shp[scale_axis] = dims[scale_axis]
        res = res * scale.reshape(shp).to(res.dtype)
    res = res.reshape(-1)
    out.reshape(-1)[: res.numel()].copy_(res)
    return out


def _gemm_nt(A, B, C=None, *, M=None, N=None, K=None, batch=1, lda=None, ldb=None, ldc=None, strideA=0, strideB=0,
             strideC=0, conjB=False, alpha=1.0, beta=0):
    # pdmrg_gemm_nt: C[b] = alpha * A[b] op(B[b])^T + beta * C[b], element strides as in the C ABI
    assert beta in (0, 1)
    if M is None:
        M, K = A.shape
        N = B.shape[0]
        lda = ldb = K
    if C is None:
        C = torch.empty((batch, M, N) if batch > 1 else (M, N), dtype=A.dtype)
        ldc, strideC = N, M * N
    ldc = N if ldc is None else ldc

    def flat(t):                                                 # the buffer from t's first element to the end of its storage
        return t.as_strided((t.untyped_storage().nbytes() // t.element_size() - t.storage_offset(),), (1,))

    Af, Bf, Cf = flat(A), flat(B), flat(C)
    for b in range(batch):
        a = torch.as_strided(Af[b * strideA:], (M, K), (lda, 1))
        bm = torch.as_strided(Bf[b * strideB:], (N, K), (ldb, 1))
        r = alpha * (a @ (bm.conj() if conjB else bm).T)
        out = torch.as_strided(Cf[b * strideC:], (M, N), (ldc, 1))
        out.copy_(r + out if beta else r)
    return C


def _heevd(a, ws):
    # pdmrg_heevd: in place, row i of ``a`` <- i-th eigenvector (ascending eigenvalues)
    w, v = torch.linalg.eigh(a)
    a.copy_(v.T)
    return w, a, torch.zeros(1, dtype=torch.int32)


def _left_basis(X, ws, keep=0):
    G = X @ X.conj().T
    lam, V = torch.linalg.eigh(G)
    lam, V = lam.flip(0), V.flip(1)
    return V.T.contiguous(), lam.clamp_min(0.0), torch.zeros(1, dtype=torch.int32)       # rows = eigenvectors u_a


def _svd(a, ws, method=0, keep=0):
    U, S, Vh = torch.linalg.svd(a, full_matrices=False)
    return U.contiguous(), S.to(torch.float64), Vh.contiguous(), torch.zeros(1, dtype=torch.int32)


def _recycle_basis(X1, X2, Wt, k, b, ws, method=None):
    from oracle import recycle_oracle as R
    Pt, Ct, S = R.recycle_basis(X1.numpy(), X2.numpy(), Wt.numpy(), int(k), int(b))
    return torch.from_numpy(np.ascontiguousarray(Pt)), torch.from_numpy(np.ascontiguousarray(Ct)), S, True, 1


class VecOps:
    """pydmrg_b200.device.VecOps on CPU tensors (csrc/krylov.cu semantics; X is a (rows, >= n) slab)."""

    def __init__(self, code, n, device):
        self.code, self.n = code, int(n)
        self.dt = torch.complex128 if code == 1 else torch.float64
        self.norm2 = torch.zeros(2, dtype=torch.float64)
        self.coefs = torch.zeros(128, dtype=self.dt)

    def _coef(self, a):
        a = np.asarray(a)
        if self.dt == torch.float64 and np.iscomplexobj(a):
            a = a.real
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dt)

    def dots(self, X, m, y, out=None):
        out = self.coefs if out is None else out
        if m > 0:
            out[:m] = X[:m, :self.n].conj() @ y[:self.n]
        return out

    def combine(self, coef, X, m, out, accumulate=False):
        r = self._coef(coef)[:m] @ X[:m, :self.n]
        out[:self.n] = out[:self.n] + r if accumulate else r
        return out

    def residual(self, v, e, XS, AX, m, r, norm2_out=None):
        norm2_out = self.norm2 if norm2_out is None else norm2_out
        c = self._coef(v)[:m]
        ee = self._coef([e])[0]
        r[:self.n] = c @ (AX[:m, :self.n] - ee * XS[:m, :self.n])
        norm2_out[0] = torch.linalg.vector_norm(r[:self.n]) ** 2
        return norm2_out

    def project(self, X, m, c_dev, r, scale_dev=None, norm2_out=None):
        norm2_out = self.norm2[1:] if norm2_out is None else norm2_out
        if m > 0:
            r[:self.n] = r[:self.n] - c_dev[:m].to(self.dt) @ X[:m, :self.n]
        if scale_dev is not None:                              # (r - sum c_i X_i) / sqrt(scale), csrc/krylov.cu
            r[:self.n] = r[:self.n] / torch.sqrt(scale_dev[0])
        norm2_out[0] = torch.linalg.vector_norm(r[:self.n]) ** 2
        return norm2_out

    def precond(self, r, diag, theta, floor, norm2_out):
        den = diag[:self.n].to(torch.complex128) - complex(theta)
        a = den.abs()
        unit = torch.where(a > 0, den / a.clamp_min(1e-300), torch.ones_like(den))
        den = torch.where(a < floor, unit * floor, den)
        t = r[:self.n] / (den if self.dt.is_complex else den.real)
        r[:self.n] = t
        norm2_out[0] = torch.linalg.vector_norm(t) ** 2
        return norm2_out

    def normalize(self, src, norm2_dev, out):
        out[:self.n] = src[:self.n] / torch.sqrt(norm2_dev[0])
        return out


class HeffPlan:
    """pydmrg_b200.device.HeffPlan: y[pout,i,r] = sign * sum L[i,j,k] Wc[(j,pout),(o,pin)] R[r,o,s] x[pin,k,s], summed over
    the MPO terms (csrc/heff.cu); ``block_masks`` keeps the (j,o) blocks this rank owns, ``k_range`` the range of the
    left ket index k it contracts (the two sharded modes)."""

    def __init__(self, code, P, Dl, Dr, ops, workspace, block_masks=None, k_range=None):
        self.code, self.P, self.Dl, self.Dr = code, P, Dl, Dr
        self.n = P * Dl * Dr
        self.ops, self.masks, self.k_range = ops, block_masks, k_range
        self.workspace, self.ws_bytes, self.n_apply = workspace, 0, 0
        self.flops_dense = self.flops_sparse = self.flops_gemm = 0.0
        self._plan = self

    def apply(self, x, y, sign=-1.0):
        P, Dl, Dr = self.P, self.Dl, self.Dr
        x3 = x.reshape(P, Dl, Dr)
        if self.k_range is not None:
            k0, kn = self.k_range
            xm = torch.zeros_like(x3)
            xm[:, k0:k0 + kn, :] = x3[:, k0:k0 + kn, :]
            x3 = xm
        acc = torch.zeros((P, Dl, Dr), dtype=x.dtype)
        for m, (L, Wc, R) in enumerate(self.ops):
            wL, wR = L.shape[1], R.shape[1]
            W4 = torch.from_numpy(np.ascontiguousarray(Wc)).to(x.dtype).reshape(wL, P, wR, P)
            if self.masks is not None:
                W4 = W4 * torch.from_numpy(np.asarray(self.masks[m], dtype=np.float64)).to(x.dtype)[:, None, :, None]
            T = torch.einsum("ros,qks->roqk", R, x3)
            U = torch.einsum("jpoq,roqk->prjk", W4, T)
            acc += torch.einsum("ijk,prjk->pir", L, U)
        y.reshape(P, Dl, Dr).copy_(sign * acc)
        self.n_apply += 1
        return y

    def diag(self, sign=-1.0, out=None):
        P, Dl, Dr = self.P, self.Dl, self.Dr
        acc = torch.zeros((P, Dl, Dr), dtype=self.ops[0][0].dtype)
        for (L, Wc, R) in self.ops:
            wL, wR = L.shape[1], R.shape[1]
            W4 = torch.from_numpy(np.ascontiguousarray(Wc)).to(acc.dtype).reshape(wL, P, wR, P)
            Wpp = torch.einsum("jpop->pjo", W4)
            acc += torch.einsum("kjk,pjo,sos->pks", L, Wpp, R)
        return (sign * acc).reshape(-1).contiguous()

    def close(self):
        pass


def _w_or_identity(W, w, d):
    return np.einsum("ab,ij->abij", np.eye(w), np.eye(d)) if W is None else np.asarray(W)


def _env_update_right(A, W, L, Abra, ws):
    # tools/env_tools.py:40-50: F'[k,m,q] = sum L[j,l,p] bra[i,j,k] W[l,m,i,n] A[n,p,q], bra = conj(A) by default
    Wt = torch.from_numpy(_w_or_identity(W, L.shape[1], A.shape[0])).to(A.dtype)
    bra = A.conj() if Abra is None else Abra
    return torch.einsum("jlp,ijk,lmin,npq->kmq", L, bra, Wt, A).contiguous()


def _env_update_left(B, W, R, Bbra, ws):
    # tools/env_tools.py:28-38: F'[x,y,a] = sum B[e,a,f] R[c,d,f] W[y,d,b,e] bra[b,x,c]
    Wt = torch.from_numpy(_w_or_identity(W, R.shape[1], B.shape[0])).to(B.dtype)
    bra = B.conj() if Bbra is None else Bbra
    return torch.einsum("eaf,cdf,ydbe,bxc->xya", B, R, Wt, bra).contiguous()


def _env_update_cols(direction, T, W, blk, Tbra, k0, kn, ws, out=None):
    full = (_env_update_right if direction == "right" else _env_update_left)(T, W, blk, Tbra, ws)
    if out is None:
        out = torch.zeros_like(full)
    out[:, :, k0:k0 + kn] = full[:, :, k0:k0 + kn]
    return out


class Context(_Ctx):
    """core.Context on the CPU (same attributes and conversions)."""

    def __init__(self, dtype="c128"):
        self.code = 1 if dtype in ("c128", 1) else 0
        self.tdtype = torch.complex128 if self.code == 1 else torch.float64
        self.npdtype = np.complex128 if self.code == 1 else np.float64
        self.device = torch.device("cpu")
        self.ws = self.svd_ws = None
        self.svd_method = -1

    def svd_method_for(self, rows, cols):
        return 2 if min(rows, cols) >= 128 else 0

    def dev(self, a):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(self.tdtype).contiguous()
        arr = np.asarray(a)
        if np.iscomplexobj(arr) and self.code == 0:
            assert np.abs(arr.imag).max(initial=0.0) == 0
            arr = arr.real
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=self.npdtype).copy())

    @staticmethod
    def host(t):
        return t.detach().numpy() if isinstance(t, torch.Tensor) else t


def install(setattr_):
    """Replace the device entry points of pydmrg_b200.device by the emulations; ``setattr_(obj, name, value)``
    is monkeypatch.setattr or the builtin setattr (worker processes)."""
    from pydmrg_b200 import device as D
    for name, fn in (("permute4", _permute4), ("gemm_nt", _gemm_nt), ("left_basis", _left_basis), ("recycle_basis", _recycle_basis),
                     ("svd", _svd), ("heevd", _heevd), ("VecOps", VecOps), ("HeffPlan", HeffPlan),
                     ("env_update_right", _env_update_right), ("env_update_left", _env_update_left),
                     ("env_update_cols", _env_update_cols)):
        setattr_(D, name, fn)
    return D
