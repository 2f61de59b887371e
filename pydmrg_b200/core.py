"""Execution context and host<->device conversion at the reference's list-of-ndarray boundary.

The reference keeps ``mps`` / ``mpo`` / ``env`` as Python lists of NumPy arrays
(pydmrg/tools/mps_tools.py:216, pydmrg/mpo/asep.py:61-66, pydmrg/tools/env_tools.py:6-26).
The public functions of this package accept exactly those objects; tensors are moved to the
device on entry and back on exit, and stay resident (as torch CUDA tensors) in between.
Every function also accepts torch CUDA tensors directly, in which case nothing is copied and
the results are device tensors ("like in, like out").
"""
import numpy as np
import torch

from . import device as D


def parse_dtype(dtype):
    if isinstance(dtype, str):
        name = dtype.lower()
    elif dtype in (np.complex128, torch.complex128, complex):
        name = "c128"
    elif dtype in (np.float64, torch.float64, float):
        name = "f64"
    else:
        name = str(dtype)
    if name in ("c128", "complex128", "complex"):
        return D.C128
    if name in ("f64", "float64", "float", "double"):
        return D.F64
    raise ValueError("dtype must be 'c128' or 'f64', got %r" % (dtype,))


class Context:
    """dtype + device + scratch for one calculation."""

    def __init__(self, dtype="c128", device=None, svd_method="auto", projective_min=128):
        self.code = parse_dtype(dtype)
        if not torch.cuda.is_available():
            raise RuntimeError("pydmrg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        D.set_device_index(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.tdtype = D.torch_dtype(self.code)
        self.npdtype = D.np_dtype(self.code)
        self.ws = D.Workspace(self.device)
        self.svd_ws = D.Workspace(self.device)
        # 'auto': cuSOLVER gesvd for small cuts, the Gram/eigh/QR route (5-8x faster) from 128 up
        self.svd_method = {"gesvd": D.SVD_GESVD, "gesvdj": D.SVD_GESVDJ, "gram": D.SVD_GRAM, "auto": -1}[svd_method]
        # finite two-site sweeps: cuts of at least this size take the projective / warm-started truncation
        # (truncate.truncate_twosite), smaller ones a full cuSOLVER SVD (cheaper there); tests lower it to put the
        # projective path in a heavily truncating regime that a dense oracle can afford
        self.projective_min = int(projective_min)

    def svd_method_for(self, rows, cols):
        if self.svd_method >= 0:
            return self.svd_method
        return D.SVD_GRAM if min(rows, cols) >= 128 else D.SVD_GESVD

    # -- conversions ---------------------------------------------------------------------
    def dev(self, a):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            if a.dtype != self.tdtype:
                if a.is_complex() and not self.tdtype.is_complex:
                    raise TypeError("complex tensor passed to an f64 context")
                a = a.to(self.tdtype)
            return a.to(self.device).contiguous()
        arr = np.asarray(a)
        if np.iscomplexobj(arr) and self.code == D.F64:
            if np.abs(arr.imag).max(initial=0.0) > 0:
                raise TypeError("complex array passed to an f64 context")
            arr = arr.real
        arr = np.ascontiguousarray(arr, dtype=self.npdtype)
        return torch.from_numpy(arr).to(self.device)

    @staticmethod
    def host(t):
        if isinstance(t, torch.Tensor):
            return t.detach().cpu().numpy()
        return t

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.tdtype, device=self.device)


_DEFAULT = {}


def default_context(dtype="c128"):
    key = (dtype, torch.cuda.current_device())
    ctx = _DEFAULT.get(key)
    if ctx is None:
        ctx = Context(dtype)
        _DEFAULT[key] = ctx
    return ctx


def is_dev(x):
    return isinstance(x, torch.Tensor)


def like(result, proto):
    """Return ``result`` (a device tensor) in the container type of ``proto``."""
    return result if is_dev(proto) else Context.host(result)
