"""GPU parity tests of the CUDA kernels, called through the C ABI (pydmrg_b200._lib / .device),
against the NumPy oracle (oracle/dmrg_oracle.py) and the golden fixtures generated from the
unmodified reference (tests/golden/*.npz).

Tolerances: the contractions are pure sums of products -> 1e-13 relative (norm-wise)
(BASELINE.md section 4); singular values 1e-12 relative to the largest one."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-13


def rel(a, b):
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def D():
    from pydmrg_b200 import device
    return device


@pytest.fixture(scope="module")
def ws(D):
    return D.Workspace(torch.device("cuda"))


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 192, 1000), (130, 70, 33), (64, 64, 64), (1, 1, 1),
                                   (3, 5, 2), (200, 64, 8), (129, 257, 130), (512, 384, 96)])
def test_gemm_nt_shapes(D, cplx, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    mk = (lambda *s: crand(rng, *s)) if cplx else (lambda *s: rng.standard_normal(s))
    A, B = mk(M, K), mk(N, K)
    C = D.gemm_nt(dev(A), dev(B))
    torch.cuda.synchronize()
    assert rel(C.cpu().numpy(), A @ B.T) < TOL


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(256, 1024, 256, 5), (128, 128, 64, 1), (130, 70, 33, 3), (256, 192, 1000, 1), (512, 2048, 512, 2),
                                   (129, 257, 130, 1), (200, 64, 8, 7), (1024, 1024, 96, 1)])
def test_gemm_nt_stream_k(D, cplx, shape):
    """Stream-K scheduling (tiles split between CTAs along K, partials added by the finisher in a fixed order): same
    numbers as the data-parallel schedule to rounding, bit-identical from run to run, with beta = 1, a batch level,
    ragged extents and complex alpha."""
    M, N, K, nb = shape
    rng = np.random.default_rng(M + 3 * N + 7 * K + nb)
    mk = (lambda *s: crand(rng, *s)) if cplx else (lambda *s: rng.standard_normal(s))
    A, B, C0 = mk(nb, M, K), mk(nb, N, K), mk(nb, M, N)
    alpha = (0.7 - 0.3j) if cplx else -1.3
    ref = alpha * np.einsum("bmk,bnk->bmn", A, B) + C0
    outs = []
    for mode in (2, 2, 0):
        D.lib.pdmrg_set_streamk_mode(mode)
        try:
            Cd = dev(C0)
            D.gemm_nt(dev(A), dev(B), Cd, M=M, N=N, K=K, batch=nb, lda=K, ldb=K, ldc=N, strideA=M * K, strideB=N * K,
                      strideC=M * N, alpha=alpha, beta=1)
            used = D.lib.pdmrg_last_gemm_streamk()
        finally:
            D.lib.pdmrg_set_streamk_mode(1)
        torch.cuda.synchronize()
        eligible = ((M + 127) // 128) * ((N + (63 if cplx else 127)) // (64 if cplx else 128)) * nb * ((K * (2 if cplx else 1) + 15) // 16) >= 8 \
            and (cplx or K % 2 == 0) and K * (2 if cplx else 1) > 16   # odd real ld: generic kernel; one k-iteration: nothing to split
        assert used == (1 if (mode == 2 and eligible) else 0)
        outs.append(Cd.cpu().numpy())
        assert rel(outs[-1], ref) < TOL
    assert np.array_equal(outs[0], outs[1])                     # deterministic summation order


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("chi", [64, 128, 256, 512])
def test_heff_stream_k_matches_data_parallel(D, dtype, chi):
    """The two-site mat-vec at the bond dimensions where the cost model switches stream-K on (configs[1], [3]) against the
    same plan with stream-K disabled and against the oracle."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import core, mpo as M
    ctx = core.Context(dtype)
    cplx = dtype == "c128"
    W1, W2 = (M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)) if cplx else M.heisenberg_mpo(6))[0][2:4]
    w = W1.shape[0]
    rng = np.random.default_rng(chi)
    mk = (lambda *s: crand(rng, *s)) if cplx else (lambda *s: rng.standard_normal(s))
    L, R, x = mk(chi, w, chi), mk(chi, w, chi), mk(2, 2, chi, chi)
    plan = D.HeffPlan(ctx.code, 4, chi, chi, [(ctx.dev(L), D.combine_twosite(W1, W2, ctx.npdtype), ctx.dev(R))], ctx.ws)
    xd = ctx.dev(x.reshape(-1))
    ys = []
    for mode in (1, 0):
        D.lib.pdmrg_set_streamk_mode(mode)
        try:
            y = torch.empty_like(xd)
            plan.apply(xd, y, -1.0)
            torch.cuda.synchronize()
        finally:
            D.lib.pdmrg_set_streamk_mode(1)
        ys.append(y.cpu().numpy())
    ref = orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas")
    if not cplx:
        ref = ref.real
    assert rel(ys[0], ref) < TOL and rel(ys[1], ref) < TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_nt_paths_agree(D, cplx):
    """TMA/DMMA path and the generic kernel give the same numbers; path selection is visible."""
    rng = np.random.default_rng(5)
    mk = (lambda *s: crand(rng, *s)) if cplx else (lambda *s: rng.standard_normal(s))
    A, B = mk(192, 80), mk(160, 80)
    C1 = D.gemm_nt(dev(A), dev(B))
    assert D.lib.pdmrg_last_gemm_path() == 1
    D.lib.pdmrg_set_force_generic_gemm(1)
    try:
        C2 = D.gemm_nt(dev(A), dev(B))
        assert D.lib.pdmrg_last_gemm_path() == 2
    finally:
        D.lib.pdmrg_set_force_generic_gemm(0)
    torch.cuda.synchronize()
    assert rel(C1.cpu().numpy(), A @ B.T) < TOL and rel(C2.cpu().numpy(), A @ B.T) < TOL


@pytest.mark.parametrize("force_generic", [0, 1])
def test_gemm_nt_options(D, force_generic):
    """conjB, complex alpha, beta=1 accumulation, batch with broadcast A, leading dimensions."""
    rng = np.random.default_rng(11)
    D.lib.pdmrg_set_force_generic_gemm(force_generic)
    try:
        M, N, K, nb = 96, 72, 40, 3
        A = crand(rng, M, K + 8)          # lda = K+8
        B = crand(rng, nb, N, K)
        C0 = crand(rng, nb, M, N + 4)     # ldc = N+4
        Cd = dev(C0)
        alpha = 0.7 - 0.3j
        D.gemm_nt(dev(A), dev(B), Cd, M=M, N=N, K=K, batch=nb, lda=K + 8, ldb=K, ldc=N + 4, strideA=0,
                  strideB=N * K, strideC=M * (N + 4), conjB=True, alpha=alpha, beta=1)
        torch.cuda.synchronize()
        ref = C0.copy()
        for b in range(nb):
            ref[b, :, :N] += alpha * (A[:, :K] @ B[b].conj().T)
        assert rel(Cd.cpu().numpy(), ref) < TOL
        # real, odd leading dimension (not 16-byte aligned rows) must still be right
        Ar, Br = rng.standard_normal((70, 33)), rng.standard_normal((50, 33))
        Cr = D.gemm_nt(dev(Ar), dev(Br), alpha=-1.0)
        torch.cuda.synchronize()
        assert rel(Cr.cpu().numpy(), -(Ar @ Br.T)) < TOL
    finally:
        D.lib.pdmrg_set_force_generic_gemm(0)


def test_permute4(D):
    rng = np.random.default_rng(3)
    a = crand(rng, 3, 4, 5, 6)
    out = torch.empty((6, 4, 3, 5), dtype=torch.complex128, device="cuda")
    sc = dev(rng.standard_normal(4))
    # out[i3,i1,i0,i2] = conj(a[i0,i1,i2,i3]) * sc[i1]
    D.permute4(dev(a), out, (6, 4, 3, 5), (1, 30, 120, 6), conj=True, scale=sc, scale_axis=1)
    torch.cuda.synchronize()
    ref = np.conj(a).transpose(3, 1, 0, 2) * sc.cpu().numpy()[None, :, None, None]
    assert rel(out.cpu().numpy(), ref) < 1e-15


# ---------------------------------------------------------------------------------------
def _heff(D, ws, x, Ls, Wcs, Rs, P, Dl, Dr, code):
    plan = D.HeffPlan(code, P, Dl, Dr, [(dev(L), Wc, dev(R)) for L, Wc, R in zip(Ls, Wcs, Rs)], ws)
    xd = dev(x.reshape(-1))
    yd = torch.empty_like(xd)
    plan.apply(xd, yd, -1.0)
    torch.cuda.synchronize()
    return yd.cpu().numpy(), plan


@pytest.mark.parametrize("c", ["a", "b", "c", "d"])
def test_heff_golden(D, ws, golden, c):
    """H1 and H2 against reference outputs (tools/diag_tools.py:108-125, 146-162)."""
    g = golden["kernels"]
    L, R, W1, W2 = g[c + "_L"], g[c + "_R"], g[c + "_W1"], g[c + "_W2"]
    x1, x2 = g[c + "_x1"], g[c + "_x2"]
    d, Dl, Dr = x1.shape
    w = W1.shape[0]
    y1, _ = _heff(D, ws, x1, [L], [D.combine_onesite(W1, w, w, d, np.complex128)], [R], d, Dl, Dr, D.C128)
    assert rel(y1, g[c + "_y1"]) < TOL
    y2, plan = _heff(D, ws, x2, [L], [D.combine_twosite(W1, W2, np.complex128)], [R], d * d, Dl, Dr, D.C128)
    assert rel(y2, g[c + "_y2"]) < TOL
    assert plan.flops_sparse <= plan.flops_dense


@pytest.mark.parametrize("chi,real", [(64, False), (96, False), (128, True), (160, False)])
def test_heff_vs_oracle_large(D, ws, chi, real):
    """TMA/DMMA tile path at sizes the oracle finishes in seconds; ASEP and Heisenberg W."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    rng = np.random.default_rng(chi)
    d = 2
    Dl, Dr = chi, chi - 8
    if real:
        W1, W2 = M.heisenberg_mpo(6)[0][2:4]
        mk = lambda *s: rng.standard_normal(s)
        code, npdt = D.F64, np.float64
    else:
        W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
        mk = lambda *s: crand(rng, *s)
        code, npdt = D.C128, np.complex128
    w = W1.shape[0]
    L, R = mk(Dl, w, Dl), mk(Dr, w, Dr)
    x2 = mk(d, d, Dl, Dr)
    y2, plan = _heff(D, ws, x2, [L], [D.combine_twosite(W1, W2, npdt)], [R], d * d, Dl, Dr, code)
    ref = orc.heff_twosite(x2.reshape(-1), [L], [W1], [W2], [R], x2.shape)
    assert rel(y2, ref) < TOL
    x1 = mk(d, Dl, Dr)
    y1, _ = _heff(D, ws, x1, [L, L], [D.combine_onesite(W1, w, w, d, npdt)] * 2, [R, R], d, Dl, Dr, code)
    ref1 = 2 * orc.heff_onesite(x1.reshape(-1), [L], [W1], [R], x1.shape)
    assert rel(y1, ref1) < TOL


def test_heff_linearity_full_size(D, ws):
    """Size-independent property at a large bond (chi=512): H(a x + b y) = a Hx + b Hy, and
    MPO-bond-sharded partial sums add up to the full result."""
    from pydmrg_b200 import mpo as M
    chi = 512
    W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.complex128)
    torch.manual_seed(0)
    L = torch.randn(chi, 6, chi, dtype=torch.complex128, device="cuda")
    R = torch.randn(chi, 6, chi, dtype=torch.complex128, device="cuda")
    plan = D.HeffPlan(D.C128, 4, chi, chi, [(L, Wc, R)], ws)
    n = 4 * chi * chi
    x = torch.randn(n, dtype=torch.complex128, device="cuda")
    y = torch.randn(n, dtype=torch.complex128, device="cuda")
    a, b = 0.3 - 1.1j, -0.8 + 0.2j
    hx, hy, hz = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    plan.apply(x, hx); plan.apply(y, hy); plan.apply(a * x + b * y, hz)
    torch.cuda.synchronize()
    err = (hz - (a * hx + b * hy)).norm() / hz.norm()
    assert err.item() < 1e-13
    mask0 = np.zeros((6, 6), dtype=np.uint8); mask0[:, 0] = 1
    mask1 = 1 - mask0
    p0 = D.HeffPlan(D.C128, 4, chi, chi, [(L, Wc, R)], ws, block_masks=[mask0])
    p1 = D.HeffPlan(D.C128, 4, chi, chi, [(L, Wc, R)], ws, block_masks=[mask1])
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    p0.apply(x, y0); p1.apply(x, y1)
    torch.cuda.synchronize()
    assert ((y0 + y1) - hx).norm().item() / hx.norm().item() < 1e-13
    # ket-index sharding: three unequal ranges of the left ket index (one of them empty) sum to the full result
    acc = torch.zeros_like(x)
    for k_range in ((0, 192), (192, 320), (512, 0)):
        pk = D.HeffPlan(D.C128, 4, chi, chi, [(L, Wc, R)], ws, k_range=k_range)
        yk = torch.empty_like(x)
        pk.apply(x, yk)
        acc += yk
    torch.cuda.synchronize()
    assert (acc - hx).norm().item() / hx.norm().item() < 1e-13


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("c", ["a", "b", "c", "d"])
def test_env_updates_golden(D, ws, golden, c):
    """E1/E2 against reference update_envR / update_envL outputs (tools/env_tools.py:28-50)."""
    g = golden["kernels"]
    A, W, L, R, Abra = g[c + "_A"], g[c + "_W1"].astype(np.complex128), g[c + "_L"], g[c + "_R"], g[c + "_Abra"]
    out = D.env_update_right(dev(A), W, dev(L), None, ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), g[c + "_envR_out"]) < TOL
    out = D.env_update_left(dev(A), W, dev(R), None, ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), g[c + "_envL_out"]) < TOL
    out = D.env_update_right(dev(A), W, dev(L), dev(Abra), ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), g[c + "_envR_bra_out"]) < TOL
    out = D.env_update_left(dev(A), None, dev(R), dev(Abra), ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), g[c + "_envL_none_out"]) < TOL


@pytest.mark.parametrize("real", [False, True])
def test_env_updates_large(D, ws, real):
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    rng = np.random.default_rng(17)
    d, Dl, Dr = 2, 144, 120
    if real:
        W = M.heisenberg_mpo(6)[0][2]; mk = lambda *s: rng.standard_normal(s)
    else:
        W = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2].astype(np.complex128); mk = lambda *s: crand(rng, *s)
    w = W.shape[0]
    A, L, R = mk(d, Dl, Dr), mk(Dl, w, Dl), mk(Dr, w, Dr)
    out = D.env_update_right(dev(A), W, dev(L), None, ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), orc.env_update_right(A, W, L)) < TOL
    out = D.env_update_left(dev(A), W, dev(R), None, ws)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), orc.env_update_left(A, W, R)) < TOL


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("real", [False, True])
def test_vec_ops(D, real):
    rng = np.random.default_rng(23)
    n, m = 70001, 19
    mk = (lambda *s: rng.standard_normal(s)) if real else (lambda *s: crand(rng, *s))
    code = D.F64 if real else D.C128
    X, AX, y = mk(m, n), mk(m, n), mk(n)
    v, e = mk(m), mk(1)[0]
    ops = D.VecOps(code, n, torch.device("cuda"))
    Xd, AXd, yd = dev(X), dev(AX), dev(y)
    dots = ops.dots(Xd, m, yd)
    torch.cuda.synchronize()
    assert rel(dots[:m].cpu().numpy(), X.conj() @ y) < 1e-13
    out = torch.empty_like(yd)
    ops.combine(v, Xd, m, out)
    torch.cuda.synchronize()
    assert rel(out.cpu().numpy(), v @ X) < 1e-13
    r = torch.empty_like(yd)
    nrm = ops.residual(v, e, Xd, AXd, m, r)
    torch.cuda.synchronize()
    rr = v @ AX - e * (v @ X)
    assert rel(r.cpu().numpy(), rr) < 1e-13
    assert abs(nrm[0].item() - np.vdot(rr, rr).real) < 1e-12 * np.vdot(rr, rr).real
    c = ops.dots(Xd, m, r)
    n2 = ops.project(Xd, m, c, r, scale_dev=nrm)
    torch.cuda.synchronize()
    ref = (rr - (X.conj() @ rr) @ X) / np.sqrt(np.vdot(rr, rr).real)
    assert rel(r.cpu().numpy(), ref) < 1e-12
    assert abs(n2[0].item() - np.vdot(ref, ref).real) < 1e-12 * np.vdot(ref, ref).real
    ops.normalize(r, n2, r)
    torch.cuda.synchronize()
    assert abs(np.linalg.norm(r.cpu().numpy()) - 1) < 1e-13


@pytest.mark.parametrize("real", [False, True])
def test_svd_gram_decaying_spectrum(D, ws, real):
    """Method 2 (Gram + heevd + Householder QR + tail refinement) on a physical-looking spectrum
    spanning 14 decades: both factors exact isometries, singular values and reconstruction to
    ~1e-12 S_0 (the error model stated in csrc/solver.cu)."""
    rng = np.random.default_rng(5)
    n = 192
    mk = (lambda *s: rng.standard_normal(s)) if real else (lambda *s: crand(rng, *s))
    u, _ = np.linalg.qr(mk(n, n)); v, _ = np.linalg.qr(mk(n, n))
    s = np.exp(-np.arange(n) * (np.log(1e14) / n))
    a = (u * s) @ v.conj().T
    U, S, Vh, info = D.svd(dev(a.copy()), ws, D.SVD_GRAM)
    torch.cuda.synchronize()
    U, S, Vh = U.cpu().numpy(), S.cpu().numpy(), Vh.cpu().numpy()
    assert np.allclose(U.conj().T @ U, np.eye(n), atol=1e-12) and np.allclose(Vh @ Vh.conj().T, np.eye(n), atol=1e-12)
    assert np.max(np.abs(S ** 2 - s ** 2)) < 1e-13 * s[0] ** 2
    assert np.max(np.abs(S - s)) < 1e-11 * s[0]
    assert rel((U * S) @ Vh, a) < 1e-11
    # rectangular + rank deficient: exact zeros in the tail must not break the refinement
    b = a[:, :64] @ mk(64, 300)
    U, S, Vh, info = D.svd(dev(b.copy()), ws, D.SVD_GRAM)
    torch.cuda.synchronize()
    U, S, Vh = U.cpu().numpy(), S.cpu().numpy(), Vh.cpu().numpy()
    assert info.item() == 0 and np.all(np.isfinite(U)) and np.all(np.isfinite(Vh))
    # the refinement gate tolerates a level-0 error estimate of up to 1e-9 S_0 (csrc/solver.cu)
    assert rel((U * S) @ Vh, b) < 2e-9
    assert np.allclose(Vh @ Vh.conj().T, np.eye(n), atol=1e-11) and np.allclose(U.conj().T @ U, np.eye(n), atol=1e-11)


def test_svd_gram_zero_padded_regression(D, ws):
    """A two-site wave-function captured right after a chi ramp (zero-padded tensors: 384 exactly-zero
    columns, rows of norm ~1e-85): cuSOLVER's heevd does not converge on its raw Gram matrix; the
    Gram route must still return finite, isometric factors (diagonal shift / gesvd fallback)."""
    import os
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "svd_rank_deficient.npz"))
    a = np.zeros((512, 512), dtype=np.complex128)
    a[:, f["cols"]] = f["block"]
    for keep in (0, 256):
        U, S, Vh, info = D.svd(dev(a.copy()), ws, D.SVD_GRAM, keep=keep)
        torch.cuda.synchronize()
        k = keep or 512
        U, S, Vh = U.cpu().numpy()[:, :k], S.cpu().numpy()[:k], Vh.cpu().numpy()[:k]
        assert info.item() == 0 and np.all(np.isfinite(U)) and np.all(np.isfinite(S)) and np.all(np.isfinite(Vh))
        assert rel((U * S) @ Vh, a) < 1e-11
        assert np.allclose(Vh @ Vh.conj().T, np.eye(k), atol=1e-11) and np.allclose(U.conj().T @ U, np.eye(k), atol=1e-11)
        Sref = np.linalg.svd(a, compute_uv=False)
        assert np.max(np.abs(S - Sref[:k])) < 1e-12 * Sref[0]


@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("real", [False, True])
@pytest.mark.parametrize("shape", [(64, 64), (48, 96), (96, 48), (2, 4), (1, 2), (200, 200)])
def test_svd(D, ws, method, real, shape):
    rng = np.random.default_rng(shape[0] * 13 + shape[1])
    a = rng.standard_normal(shape) if real else crand(rng, *shape)
    U, S, Vh, info = D.svd(dev(a.copy()), ws, method)
    torch.cuda.synchronize()
    assert info.item() == 0
    U, S, Vh = U.cpu().numpy(), S.cpu().numpy(), Vh.cpu().numpy()
    Sref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(S - Sref)) < 1e-12 * Sref[0]
    assert rel((U * S) @ Vh, a) < 1e-12
    k = min(shape)
    assert np.allclose(U.conj().T @ U, np.eye(k), atol=1e-11) and np.allclose(Vh @ Vh.conj().T, np.eye(k), atol=1e-11)


@pytest.mark.parametrize("real", [False, True])
def test_heevd(D, ws, real):
    rng = np.random.default_rng(2)
    n = 96
    m = rng.standard_normal((n, 40)) if real else crand(rng, n, 40)
    rho = m @ m.conj().T
    w, vecs, info = D.heevd(dev(rho.copy()), ws)
    torch.cuda.synchronize()
    assert info.item() == 0
    w, vecs = w.cpu().numpy(), vecs.cpu().numpy()
    assert np.allclose(w, np.linalg.eigvalsh(rho), atol=1e-11 * abs(w).max())
    # row i of vecs is eigenvector i:  rho v = w v
    assert rel(rho @ vecs.T, vecs.T * w) < 1e-11


@pytest.mark.parametrize("mode", ["bond", "ket"])
def test_fused_sharded_path_world1(D, ws, mode):
    """Peer-memory path (scatter-epilogue GEMM, flags, reduce + all-gather kernel) with a single rank:
    must reproduce the plain mat-vec.  mode='ket' runs stage C as one long-K GEMM (MPO-bond slabs folded
    into the k-loop, pdmrg_gemm_nt_scatter(accumulate_b2=1))."""
    from pydmrg_b200 import mpo as M, parallel
    chi = 192
    W1, W2 = M.asep_mpo(6, (0.35, 0.0, 1.0, 0.0, 2 / 3, 0.0, -0.5))[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.complex128)
    torch.manual_seed(3)
    L = torch.randn(chi, 6, chi, dtype=torch.complex128, device="cuda")
    R = torch.randn(chi - 16, 6, chi - 16, dtype=torch.complex128, device="cuda")
    x = torch.randn(4 * chi * (chi - 16), dtype=torch.complex128, device="cuda")
    plan = D.HeffPlan(D.C128, 4, chi, chi - 16, [(L, Wc, R)], ws)
    y0 = torch.empty_like(x); plan.apply(x, y0)
    fused = parallel.FusedShardedHeff(D.C128, 4, chi, chi - 16, [(L, Wc, R)], ws, 0, 1, mode=mode)
    y1 = torch.empty_like(x)
    for _ in range(3):                      # epochs advance; buffers are re-used
        fused.apply(x, y1)
    torch.cuda.synchronize()
    assert ((y1 - y0).norm() / y0.norm()).item() < 1e-14
    fused.close()


def test_sharded_paths_two_gpus():
    """N = 2 ranks (torchrun): NCCL-all-reduce path and the fused peer-memory path both equal the
    unsharded mat-vec, and the fused result is bit-identical on both ranks."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "mgpu_sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout


@pytest.mark.parametrize("dtype", ["c128", "f64"])
@pytest.mark.parametrize("method", [0, 1])
def test_recycle_basis_methods(dtype, method):
    """pdmrg_recycle_basis, Householder (0) and Cholesky-QR2 (1): P has orthonormal rows, C = P X2^T, the rows
    keep the nested ordering (row i only mixes with rows <= i outside the window), the window rows of C are
    mutually orthogonal and sorted, and both methods span the same kept subspace."""
    import torch
    from pydmrg_b200 import core, device as D
    ctx = core.Context(dtype)
    rng = np.random.default_rng(11)
    a, c, k, p, b = 192, 160, 64, 8, 8
    K = k + p
    cplx = dtype == "c128"
    rnd = (lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if cplx else (lambda *s: rng.standard_normal(s))
    U, _ = np.linalg.qr(rnd(a, a)); V, _ = np.linalg.qr(rnd(c, c))
    q = min(a, c)
    s = np.exp(-np.arange(q) * 0.15)
    X1 = (U[:, :q] * s) @ V[:, :q].conj().T                       # (a, c)
    W = V[:, :K].conj().T + 1e-5 * rnd(K, c)                       # rows ~ conj(right vectors)^T, slightly off
    Pt, Ct, Sh, ok, used = D.recycle_basis(ctx.dev(X1), ctx.dev(np.ascontiguousarray(X1.T)), ctx.dev(np.ascontiguousarray(W)), k, b,
                                           ctx.svd_ws, method=method)
    assert ok and used == method
    P, C = ctx.host(Pt), ctx.host(Ct)
    assert np.max(np.abs(P @ P.conj().T - np.eye(K))) < 1e-12
    assert np.max(np.abs(C - P @ X1)) < 1e-12 * s[0] * a          # C[i,x] = sum_r P[i,r] X1[r,x]
    assert np.max(np.abs(Sh - np.linalg.norm(C, axis=1))) < 1e-12
    # rows of P are conj(left vectors): P[i] ~ conj(u_i) up to a phase for well separated weights
    ov = np.abs(P[: k - b] @ U[:, : k - b])                        # sum_r conj(u_i)[r] u_j[r]
    assert np.max(np.abs(np.diag(ov) - 1.0)) < 1e-6
    # window: C rows orthogonal, norms descending and equal to the singular values there
    w0 = k - b
    Cw = C[w0:]
    Gw = Cw @ Cw.conj().T
    assert np.max(np.abs(Gw - np.diag(np.diag(Gw)))) < 1e-12 * s[w0] ** 2 * 10
    assert np.all(np.diff(Sh[w0:]) <= 1e-15)
    assert np.max(np.abs(Sh[:K] - s[:K])) < 1e-8 * s[0]


def test_recycle_basis_rank_deficient_cholqr():
    """Rank-deficient wave-function (numerical rank 12 << K): the noise rows of T are replaced by pseudo-random
    vectors, so the Cholesky-QR2 route still delivers an orthonormal basis whose leading rows span the range."""
    import torch
    from pydmrg_b200 import core, device as D
    ctx = core.Context("c128")
    rng = np.random.default_rng(4)
    a, c, k, p, r0 = 256, 256, 128, 16, 12
    K = k + p
    rnd = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    U, _ = np.linalg.qr(rnd(a, a)); V, _ = np.linalg.qr(rnd(c, c))
    s = np.zeros(a); s[:r0] = np.exp(-np.arange(r0) * 0.8)
    X1 = (U * s) @ V.conj().T
    W = V[:, :K].conj().T.copy()
    Pt, Ct, Sh, ok, used = D.recycle_basis(ctx.dev(X1), ctx.dev(np.ascontiguousarray(X1.T)), ctx.dev(W), k, 16, ctx.svd_ws,
                                           method=D.RECYCLE_CHOLQR2)
    assert ok and used == D.RECYCLE_CHOLQR2
    P, C = ctx.host(Pt), ctx.host(Ct)
    assert np.max(np.abs(P @ P.conj().T - np.eye(K))) < 1e-12
    assert np.max(np.abs(Sh[:r0] - s[:r0])) < 1e-12 and np.max(Sh[r0:]) < 1e-13
    # projection on the kept rows loses nothing
    assert abs(np.linalg.norm(C[:k]) - np.linalg.norm(X1)) < 1e-13
