"""CPU model check of csrc/env_cols.cu (column-restricted block updates for sharded sweeps).

The entry points pdmrg_env_update_right_cols / _left_cols only ORCHESTRATE kernels that the GPU tests already
validate (pdmrg_permute4, pdmrg_gemm_nt, the sparse W mix): what is new is their index arithmetic -- pointer
offsets, extents, leading dimensions, batch strides.  This file re-issues exactly those calls, argument for
argument, against flat-buffer NumPy models of the three kernels' documented semantics and compares the result with
the oracle's block update restricted to the owned ket columns.  It cannot see the CUDA execution (first GPU run:
tests/mgpu_sharded_sweep_check.py), but it pins the arithmetic the .cu file passes to the kernels."""
import numpy as np
import pytest

from oracle import dmrg_oracle as orc


# ---- flat-buffer models of the kernels (element offsets, as the C ABI documents them) --------------------------
def permute4(src, src_off, out, out_off, n, s, conj):
    """pdmrg_permute4: out[contiguous (n0,n1,n2,n3)] = src[sum idx_i * s_i] (optionally conjugated)."""
    idx = sum(np.arange(n[a]).reshape([-1 if b == a else 1 for b in range(4)]) * s[a] for a in range(4))
    vals = src[src_off + idx.reshape(-1)]
    out[out_off:out_off + vals.size] = np.conj(vals) if conj else vals


def gemm_nt(M, N, K, batch, A, a_off, lda, sA, B, b_off, ldb, sB, C, c_off, ldc, sC):
    """pdmrg_gemm_nt (alpha = 1, beta = 0, no conj): C[b][m,n] = sum_k A[b][m,k] B[b][n,k]."""
    for b in range(batch):
        a = np.lib.stride_tricks.as_strided(A[a_off + b * sA:], (M, K), (lda * A.itemsize, A.itemsize))
        bb = np.lib.stride_tricks.as_strided(B[b_off + b * sB:], (N, K), (ldb * B.itemsize, B.itemsize))
        r = a @ bb.T
        for m in range(M):
            C[c_off + b * sC + m * ldc: c_off + b * sC + m * ldc + N] = r[m]


def mix(rows, inp, out, nx, ny, six, siy, sox, soy):
    """launch_mix: out[out_off[row] + x*sox + y*soy] = sum_nz val * in[in_off[nz] + x*six + y*siy]."""
    x, y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    for out_off, terms in rows:
        acc = np.zeros((nx, ny), dtype=inp.dtype)
        for in_off, val in terms:
            acc += val * inp[in_off + x * six + y * siy]
        out[out_off + x * sox + y * soy] = acc


# ---- transliteration of csrc/env_cols.cu ------------------------------------------------------------------------
def env_update_right_cols(d, Dl, Dr, wl, wr, W, L, A, Abra, k0, kn, out):
    At = np.zeros(d * Dl * Dr, complex); AbT = np.zeros(d * Dl * Dr, complex)
    wmax = max(wl, wr)
    T1 = np.zeros(d * Dl * Dr * wmax, complex); V = np.zeros(d * Dl * Dr * wmax, complex)
    Af, Lf = A.reshape(-1), L.reshape(-1)
    permute4(Af, k0, At, 0, (1, d, kn, Dl), (0, Dl * Dr, 1, Dr), False)
    bra = Af if Abra is None else Abra.reshape(-1)
    permute4(bra, 0, AbT, 0, (1, Dr, d, Dl), (0, 1, Dl * Dr, Dr), Abra is None)
    gemm_nt(d * kn, Dl * wl, Dl, 1, At, 0, Dl, 0, Lf, 0, Dl, 0, T1, 0, Dl * wl, 0)
    rows = []
    for m in range(wr):
        for i in range(d):
            terms = []
            for l in range(wl):
                for n in range(d):
                    w = W[l, m, i, n] if W is not None else (1.0 if (l == m and i == n) else 0.0)
                    if w != 0:
                        terms.append((l + n * kn * Dl * wl, w))
            rows.append((m * kn * d * Dl + i * Dl, terms))
    mix(rows, T1, V, kn, Dl, Dl * wl, wl, d * Dl, 1)
    gemm_nt(Dr, kn, d * Dl, wr, AbT, 0, d * Dl, 0, V, 0, d * Dl, kn * d * Dl, out, k0, wr * Dr, Dr)


def env_update_left_cols(d, Dl, Dr, wl, wr, W, R, B, Bbra, k0, kn, out):
    BbT = np.zeros(d * Dl * Dr, complex)
    wmax = max(wl, wr)
    T1 = np.zeros(d * Dl * Dr * wmax, complex); V = np.zeros(d * Dl * Dr * wmax, complex)
    Bf, Rf = B.reshape(-1), R.reshape(-1)
    bra = Bf if Bbra is None else Bbra.reshape(-1)
    permute4(bra, 0, BbT, 0, (1, Dl, d, Dr), (0, Dr, Dl * Dr, 1), Bbra is None)
    gemm_nt(kn, Dr * wr, Dr, d, Bf, k0 * Dr, Dr, Dl * Dr, Rf, 0, Dr, 0, T1, 0, Dr * wr, kn * Dr * wr)
    rows = []
    for y in range(wl):
        for b in range(d):
            terms = []
            for dd in range(wr):
                for e in range(d):
                    w = W[y, dd, b, e] if W is not None else (1.0 if (y == dd and b == e) else 0.0)
                    if w != 0:
                        terms.append((dd + e * kn * Dr * wr, w))
            rows.append((y * kn * d * Dr + b * Dr, terms))
    mix(rows, T1, V, kn, Dr, Dr * wr, wr, d * Dr, 1)
    gemm_nt(Dl, kn, d * Dr, wl, BbT, 0, d * Dr, 0, V, 0, d * Dr, kn * d * Dr, out, k0, wl * Dl, Dl)


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


@pytest.mark.parametrize("use_W,use_bra", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("dims", [(2, 6, 9), (2, 8, 8), (3, 5, 4)])
def test_column_restricted_updates_assemble_the_full_block(dims, use_W, use_bra):
    d, Dl, Dr = dims
    rng = np.random.default_rng(d * 100 + Dl * 10 + Dr)
    wl, wr = (4, 3) if use_W else (3, 3)
    A = crand(rng, d, Dl, Dr)
    bra = crand(rng, d, Dl, Dr) if use_bra else None
    # ---- right-growing: new left block (Dr, wr, Dr) from L (Dl, wl, Dl)
    W = crand(rng, wl, wr, d, d) * (rng.random((wl, wr, 1, 1)) > 0.4) if use_W else None
    L = crand(rng, Dl, wl, Dl)
    full = orc.env_update_right(A, W, L, bra, "blas")
    out = np.zeros(Dr * wr * Dr, complex)
    cuts = [0, Dr // 3, Dr // 3, Dr]                          # three ranks, one of them with an empty range
    for k0, k1 in zip(cuts[:-1], cuts[1:]):
        part = np.full(Dr * wr * Dr, np.nan + 0j)             # "everything else is left untouched"
        env_update_right_cols(d, Dl, Dr, wl, wr, W, L, A, bra, k0, k1 - k0, part)
        owned = np.zeros((Dr, wr, Dr), bool); owned[:, :, k0:k1] = True
        assert np.all(np.isnan(part.reshape(Dr, wr, Dr)[~owned])) and not np.any(np.isnan(part.reshape(Dr, wr, Dr)[owned]))
        out += np.where(np.isnan(part), 0, part)
    assert np.allclose(out.reshape(Dr, wr, Dr), full, rtol=0, atol=1e-12 * np.abs(full).max())
    # ---- left-growing: new right block (Dl, wl, Dl) from R (Dr, wr, Dr)
    W = crand(rng, wl, wr, d, d) * (rng.random((wl, wr, 1, 1)) > 0.4) if use_W else None
    R = crand(rng, Dr, wr, Dr)
    full = orc.env_update_left(A, W, R, bra, "blas")
    out = np.zeros(Dl * wl * Dl, complex)
    cuts = [0, Dl // 2, Dl]
    for k0, k1 in zip(cuts[:-1], cuts[1:]):
        part = np.full(Dl * wl * Dl, np.nan + 0j)
        env_update_left_cols(d, Dl, Dr, wl, wr, W, R, A, bra, k0, k1 - k0, part)
        out += np.where(np.isnan(part), 0, part)
    assert np.allclose(out.reshape(Dl, wl, Dl), full, rtol=0, atol=1e-12 * np.abs(full).max())


def test_shard_spec_env_update_sums_to_the_full_block(monkeypatch):
    """parallel.ShardSpec.env_update: the ranks' slices (ket_chunks ranges) add up to the whole block."""
    torch = pytest.importorskip("torch")
    from pydmrg_b200 import device as D, parallel

    def fake_cols(direction, T, W, blk, Tbra, k0, kn, ws, out=None):
        f = orc.env_update_right if direction == "right" else orc.env_update_left
        full = f(T.numpy(), W, blk.numpy(), None, "blas")
        res = np.zeros_like(full)
        res[:, :, k0:k0 + kn] = full[:, :, k0:k0 + kn]
        return torch.from_numpy(res)

    class Dist:
        class ReduceOp:
            SUM = 0

        def all_reduce(self, t, op=None):
            return t

    class Ctx:
        ws = None

    monkeypatch.setattr(D, "env_update_cols", fake_cols)
    rng = np.random.default_rng(1)
    A = torch.from_numpy(crand(rng, 2, 32, 48))
    W = crand(rng, 3, 3, 2, 2)
    for direction, blk in (("right", torch.from_numpy(crand(rng, 32, 3, 32))), ("left", torch.from_numpy(crand(rng, 48, 3, 48)))):
        acc = 0
        for rank in range(3):
            acc = acc + parallel.ShardSpec(rank, 3, Dist(), min_dim=16).env_update(direction, A, W, blk, Ctx()).numpy()
        f = orc.env_update_right if direction == "right" else orc.env_update_left
        assert np.allclose(acc, f(A.numpy(), W, blk.numpy(), None, "blas"), atol=1e-12)


# ---- the .cu file itself, compiled with g++ against host stand-ins for the kernels it launches ----------------------
@pytest.fixture(scope="module")
def env_cols_host(tmp_path_factory):
    import ctypes
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("env_cols_host") / "libenv_cols_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-w", "-include", "cmath",
                    "-I/usr/local/cuda/include", "-x", "c++", os.path.join(root, "pydmrg_b200", "csrc", "env_cols.cu"),
                    os.path.join(root, "tests", "native_host", "kernels_host_shim.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.shim_last_error.restype = ctypes.c_char_p
    return lib


@pytest.mark.parametrize("dtype", ["c128", "f64"])
@pytest.mark.parametrize("use_W,use_bra", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("dims", [(2, 6, 9), (2, 8, 8), (3, 5, 4)])
def test_env_cols_cu_on_host_memory(env_cols_host, dims, use_W, use_bra, dtype):
    """pydmrg_b200/csrc/env_cols.cu executed as it is (host stand-ins for permute / GEMM / mix): the column slices
    written in place assemble the oracle's full block update, and nothing outside the owned columns is touched."""
    import ctypes
    lib = env_cols_host
    d, Dl, Dr = dims
    rng = np.random.default_rng(7 * d + Dl + 3 * Dr)
    cplx = dtype == "c128"
    dt = np.complex128 if cplx else np.float64
    rnd = (lambda *s: crand(rng, *s)) if cplx else (lambda *s: rng.standard_normal(s))
    wl, wr = (4, 3) if use_W else (3, 3)
    A = np.ascontiguousarray(rnd(d, Dl, Dr))
    bra = np.ascontiguousarray(rnd(d, Dl, Dr)) if use_bra else None
    P = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
    ws = np.zeros(1 << 20, np.uint8)
    for direction in ("right", "left"):
        W = np.ascontiguousarray(rnd(wl, wr, d, d) * (rng.random((wl, wr, 1, 1)) > 0.4)).astype(dt) if use_W else None
        if direction == "right":
            blk = np.ascontiguousarray(rnd(Dl, wl, Dl)); n, w_out = Dr, wr
            full = orc.env_update_right(A, W, blk, bra, "blas")
            fn = lib.pdmrg_env_update_right_cols
        else:
            blk = np.ascontiguousarray(rnd(Dr, wr, Dr)); n, w_out = Dl, wl
            full = orc.env_update_left(A, W, blk, bra, "blas")
            fn = lib.pdmrg_env_update_left_cols
        out = np.full((n, w_out, n), np.nan, dt)
        cuts = [0, n // 3, n // 3, n]
        for k0, k1 in zip(cuts[:-1], cuts[1:]):
            before = out.copy()
            rc = fn(1 if cplx else 0, d, Dl, Dr, wl, wr, P(W), P(blk), P(A), P(bra), k0, k1 - k0, P(out), P(ws),
                    ctypes.c_size_t(ws.size), None)
            assert rc == 0, lib.shim_last_error()
            untouched = np.ones(out.shape, bool); untouched[:, :, k0:k1] = False
            assert np.array_equal(np.isnan(out[untouched]), np.isnan(before[untouched]))
            same = ~np.isnan(before) & untouched
            assert np.array_equal(out[same], before[same])
        assert not np.isnan(out).any()
        assert np.allclose(out, full, rtol=0, atol=1e-12 * np.abs(full).max())
