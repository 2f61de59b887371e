"""The orchestration of the dominant path -- pydmrg_b200/csrc/heff.cu: mat-vec plans (stage-A / mix / stage-C batching,
MPO block-sparsity lists, the bond-sharded and ket-split variants) and the block updates -- compiled with g++ against
host stand-ins for the kernels it launches (tests/native_host/kernels_host_shim.cpp) and executed on HOST memory
through the same C ABI the product binds.  Checked against the reference's kernel fixtures (tests/golden/kernels.npz)
and the oracle.  The CUDA kernels themselves (DMMA GEMM, mix, permute, slab reduce) are what ``-m gpu`` covers; this
keeps the host-side index arithmetic under test in every CPU run."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import dmrg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp, i32, f64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
PV, PI = ctypes.POINTER(vp), ctypes.POINTER(i32)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("heff_host") / "libheff_host.so")
    csrc = os.path.join(ROOT, "pydmrg_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-w", "-include", "cmath",
                    "-I/usr/local/cuda/include", "-x", "c++", os.path.join(csrc, "heff.cu"), os.path.join(csrc, "env_cols.cu"),
                    os.path.join(ROOT, "tests", "native_host", "kernels_host_shim.cpp"), "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.shim_last_error.restype = ctypes.c_char_p
    L.pdmrg_heff_plan_create.argtypes = [PV, i32, i32, i32, i32, i32, PI, PI, PV]
    L.pdmrg_heff_plan_create_sharded.argtypes = [PV, i32, i32, i32, i32, i32, PI, PI, PV, PV]
    L.pdmrg_heff_plan_create_ketsplit.argtypes = [PV, i32, i32, i32, i32, i32, PI, PI, PV, PV, i32, i32]
    L.pdmrg_heff_plan_destroy.argtypes = [vp]
    L.pdmrg_heff_workspace_bytes.restype = sz
    L.pdmrg_heff_workspace_bytes.argtypes = [vp]
    L.pdmrg_heff_apply.argtypes = [vp, PV, PV, vp, vp, f64, vp, sz, vp]
    for name in ("pdmrg_heff_flops_dense", "pdmrg_heff_flops_sparse"):
        getattr(L, name).restype = f64
        getattr(L, name).argtypes = [vp]
    L.pdmrg_env_workspace_bytes.restype = sz
    L.pdmrg_env_workspace_bytes.argtypes = [i32] * 6
    L.pdmrg_env_update_right.argtypes = [i32] * 6 + [vp, vp, vp, vp, vp, vp, sz, vp]
    L.pdmrg_env_update_left.argtypes = [i32] * 6 + [vp, vp, vp, vp, vp, vp, sz, vp]
    return L


def ptr(a):
    return vp(a.ctypes.data) if a is not None else None


def apply_plan(L, code, P, Dl, Dr, ops, x, sign=-1.0, masks=None, k_range=None):
    """ops: [(L, Wc, R)] NumPy arrays; returns y.  Mirrors pydmrg_b200/device.py:HeffPlan."""
    dt = np.complex128 if code == 1 else np.float64
    n_mpo = len(ops)
    Ls = [np.ascontiguousarray(o[0], dtype=dt) for o in ops]
    Wcs = [np.ascontiguousarray(o[1], dtype=dt) for o in ops]
    Rs = [np.ascontiguousarray(o[2], dtype=dt) for o in ops]
    wL = (i32 * n_mpo)(*[a.shape[1] for a in Ls])
    wR = (i32 * n_mpo)(*[a.shape[1] for a in Rs])
    wc = (vp * n_mpo)(*[a.ctypes.data for a in Wcs])
    plan = vp(0)
    mk = None
    if masks is not None:
        mk_arr = [np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
        mk = (vp * n_mpo)(*[a.ctypes.data for a in mk_arr])
    if k_range is not None:
        rc = L.pdmrg_heff_plan_create_ketsplit(ctypes.byref(plan), code, P, Dl, Dr, n_mpo, wL, wR, wc, mk, k_range[0], k_range[1])
    elif masks is not None:
        rc = L.pdmrg_heff_plan_create_sharded(ctypes.byref(plan), code, P, Dl, Dr, n_mpo, wL, wR, wc, mk)
    else:
        rc = L.pdmrg_heff_plan_create(ctypes.byref(plan), code, P, Dl, Dr, n_mpo, wL, wR, wc)
    assert rc == 0, L.shim_last_error()
    ws = np.zeros(int(L.pdmrg_heff_workspace_bytes(plan)) + 64, np.uint8)
    Lp = (vp * n_mpo)(*[a.ctypes.data for a in Ls])
    Rp = (vp * n_mpo)(*[a.ctypes.data for a in Rs])
    xx = np.ascontiguousarray(x, dtype=dt).reshape(-1)
    y = np.full(xx.size, np.nan, dt)
    rc = L.pdmrg_heff_apply(plan, Lp, Rp, ptr(xx), ptr(y), sign, ptr(ws), ws.size, None)
    assert rc == 0, L.shim_last_error()
    dense, sparse = L.pdmrg_heff_flops_dense(plan), L.pdmrg_heff_flops_sparse(plan)
    L.pdmrg_heff_plan_destroy(plan)
    return y, dense, sparse


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


@pytest.mark.parametrize("c", ["a", "b", "c", "d"])
def test_matvec_plans_on_reference_fixtures(lib, golden, c):
    """One-site and two-site H_eff through pdmrg_heff_plan_create / pdmrg_heff_apply = the reference's Hfun."""
    from pydmrg_b200.device import combine_onesite, combine_twosite
    g = golden["kernels"]
    Lb, Rb, W1, W2, x1, x2 = g[c + "_L"], g[c + "_R"], g[c + "_W1"], g[c + "_W2"], g[c + "_x1"], g[c + "_x2"]
    d, Dl, Dr = x1.shape
    y, dense, sparse = apply_plan(lib, 1, d, Dl, Dr, [(Lb, combine_onesite(W1, Lb.shape[1], Rb.shape[1], d, np.complex128), Rb)], x1)
    assert rel(y, g[c + "_y1"]) < 1e-13
    y, dense, sparse = apply_plan(lib, 1, d * d, Dl, Dr, [(Lb, combine_twosite(W1, W2, np.complex128), Rb)], x2)
    assert rel(y, g[c + "_y2"]) < 1e-13
    assert 0 < sparse <= dense                                   # ASEP W is block sparse: skipped slabs cost no FLOPs
    # the sign argument and a second MPO term (sum over mpoInd, diag_tools.py:111)
    y2, _, _ = apply_plan(lib, 1, d * d, Dl, Dr, [(Lb, combine_twosite(W1, W2, np.complex128), Rb)] * 2, x2, sign=1.0)
    assert rel(y2, -2 * g[c + "_y2"]) < 1e-13


@pytest.mark.parametrize("c", ["a", "c"])
def test_sharded_plans_add_up(lib, golden, c):
    """Bond-sharded (block masks from the exact partitioner) and ket-split plans: the ranks' partial results sum to
    the unsharded mat-vec, for 2, 3 and 4 ranks."""
    from pydmrg_b200.device import combine_twosite
    from pydmrg_b200.parallel import block_pattern, ket_chunks, partition_blocks
    g = golden["kernels"]
    Lb, Rb, W1, W2, x2 = g[c + "_L"], g[c + "_R"], g[c + "_W1"], g[c + "_W2"], g[c + "_x2"]
    d, _, Dl, Dr = x2.shape
    Wc = combine_twosite(W1, W2, np.complex128)
    ref = g[c + "_y2"]
    for world in (2, 3, 4):
        parts = partition_blocks(block_pattern(Wc, Lb.shape[1], Rb.shape[1], d * d), world)
        acc = sum(apply_plan(lib, 1, d * d, Dl, Dr, [(Lb, Wc, Rb)], x2, masks=[parts[r]])[0] for r in range(world))
        assert rel(acc, ref) < 1e-13
        acc = 0
        for k0, kn in ket_chunks(Dl, world, align=2):
            if kn > 0:
                acc = acc + apply_plan(lib, 1, d * d, Dl, Dr, [(Lb, Wc, Rb)], x2, k_range=(k0, kn))[0]
        assert rel(acc, ref) < 1e-13


def test_real_dtype_heisenberg(lib):
    """f64 plans (Heisenberg, w = 5) against the oracle's two-site Hfun."""
    from pydmrg_b200.device import combine_twosite
    rng = np.random.default_rng(3)
    W = orc.heisenberg_mpo(6)[0]
    W1, W2 = W[2], W[3]
    Dl, Dr = 7, 5
    Lb, Rb, x = rng.standard_normal((Dl, 5, Dl)), rng.standard_normal((Dr, 5, Dr)), rng.standard_normal((2, 2, Dl, Dr))
    y, _, _ = apply_plan(lib, 0, 4, Dl, Dr, [(Lb, combine_twosite(W1, W2, np.float64), Rb)], x)
    ref = orc.heff_twosite(x.reshape(-1).astype(complex), [Lb], [W1], [W2], [Rb], x.shape, "blas")
    assert rel(y, ref.real) < 1e-13 and np.abs(ref.imag).max() == 0


@pytest.mark.parametrize("c", ["a", "b", "c", "d"])
def test_block_updates_on_reference_fixtures(lib, golden, c):
    """pdmrg_env_update_right / _left = the reference's update_envR / update_envL, with the default (conjugated) bra,
    an explicit bra and a None MPO entry."""
    g = golden["kernels"]
    A, W, Lb, Rb, bra = g[c + "_A"], g[c + "_W1"], g[c + "_L"], g[c + "_R"], g[c + "_Abra"]
    d, Dl, Dr = A.shape
    w = W.shape[0]
    A, W, Lb, Rb, bra = (np.ascontiguousarray(t, dtype=np.complex128) for t in (A, W, Lb, Rb, bra))
    ws = np.zeros(int(lib.pdmrg_env_workspace_bytes(1, d, Dl, Dr, w, w)) + 64, np.uint8)

    def run(fn, Wh, blk, T, Tb, shape):
        out = np.full(shape, np.nan, np.complex128)
        rc = fn(1, d, Dl, Dr, w, w, ptr(Wh), ptr(blk), ptr(T), ptr(Tb), ptr(out), ptr(ws), ws.size, None)
        assert rc == 0, lib.shim_last_error()
        return out

    assert rel(run(lib.pdmrg_env_update_right, W, Lb, A, None, (Dr, w, Dr)), g[c + "_envR_out"]) < 1e-13
    assert rel(run(lib.pdmrg_env_update_left, W, Rb, A, None, (Dl, w, Dl)), g[c + "_envL_out"]) < 1e-13
    assert rel(run(lib.pdmrg_env_update_right, W, Lb, A, bra, (Dr, w, Dr)), g[c + "_envR_bra_out"]) < 1e-13
    assert rel(run(lib.pdmrg_env_update_left, None, Rb, A, bra, (Dl, w, Dl)), g[c + "_envL_none_out"]) < 1e-13
