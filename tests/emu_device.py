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
    assert batch == 1 and beta in (0, 1)
    if M is None:
        M, K = A.shape
        N = B.shape[0]
        lda = ldb = K
    a = torch.as_strided(A.reshape(-1), (M, K), (lda, 1))
    b = torch.as_strided(B.reshape(-1), (N, K), (ldb, 1))
    r = alpha * (a @ (b.conj() if conjB else b).T)
    if C is None:
        return r.contiguous()
    ldc = N if ldc is None else ldc
    out = torch.as_strided(C.reshape(-1), (M, N), (ldc, 1))
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

    def normalize(self, src, norm2_dev, out):
        out[:self.n] = src[:self.n] / torch.sqrt(norm2_dev[0])
        return out


def install(setattr_):
    """Replace the device entry points of pydmrg_b200.device by the emulations; ``setattr_(obj, name, value)``
    is monkeypatch.setattr or the builtin setattr (worker processes)."""
    from pydmrg_b200 import device as D
    for name, fn in (("permute4", _permute4), ("gemm_nt", _gemm_nt), ("left_basis", _left_basis), ("recycle_basis", _recycle_basis),
                     ("svd", _svd), ("heevd", _heevd), ("VecOps", VecOps)):
        setattr_(D, name, fn)
    return D
