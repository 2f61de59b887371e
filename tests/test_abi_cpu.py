"""CPU: the C-ABI shared library loads without a GPU and exports every symbol declared in
include/pydmrg_b200.h; host-side logic of the package (MPS utilities, MPO combination)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "pydmrg_b200", "libpydmrg_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    from pydmrg_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "pydmrg_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pdmrg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib.lib, name), name
    assert set(lib.EXPORTS) == set(declared)
    assert lib.lib.pdmrg_version() >= 100
    assert lib.lib.pdmrg_launch_count() == 0        # nothing ran: no compute without a GPU


def test_argument_validation_without_gpu(lib):
    """Error behaviour: bad arguments are rejected on the host with a message (no launch)."""
    rc = lib.lib.pdmrg_gemm_nt(7, 4, 4, 4, 1, None, 4, 0, None, 4, 0, 0, None, 4, 0, 1.0, 0.0, 0, None)
    assert rc != 0 and b"dtype" in lib.lib.pdmrg_last_error()
    rc = lib.lib.pdmrg_gemm_nt(0, 4, 4, 4, 1, None, 4, 0, None, 4, 0, 0, None, 4, 0, 1.0, 0.0, 3, None)
    assert rc != 0 and b"beta" in lib.lib.pdmrg_last_error()
    with pytest.raises(lib.PdmrgError):
        lib.check(rc, "pdmrg_gemm_nt")


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "pydmrg_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|tools\.|dmrg\b|idmrg\b)", src, re.M), fn
            assert "ref_loader" not in src and "sys.path.insert(0, '/root/reference" not in src, fn


def test_mps_utils_match_oracle_draw_order():
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mps as mt
    for N in (5, 6, 9):
        np.random.seed(4)
        a = mt.create_rand_mps(N, 7)
        np.random.seed(4)
        b = orc.random_mps(N, 7)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        a = mt.make_mps_right(a); b = orc.make_mps_right(b)
        assert all(np.allclose(x, y) for x, y in zip(a, b))
        for t in a[1:]:
            assert np.allclose(np.einsum("ijk,ilk->jl", t, t.conj()), np.eye(t.shape[1]), atol=1e-12)


def test_increase_mbd_and_io(tmp_path):
    from pydmrg_b200 import mps as mt
    np.random.seed(1)
    mpsL = mt.create_all_mps(7, 4, 2)
    big = mt.increase_all_mbd([[t.copy() for t in M] for M in mpsL], 16)
    dims = [t.shape for t in big[0]]
    assert dims[0] == (2, 1, 2) and dims[3] == (2, 8, 8) and dims[-1] == (2, 2, 1)
    assert np.allclose(big[0][3][:, :4, :4], mpsL[0][3])
    fn = str(tmp_path / "state")
    mt.save_mps(big, fn, gaugeSite=3, fformat="npz")
    back, site = mt.load_mps(fn, fformat="npz")
    assert int(site) == 3 and len(back) == 2
    assert all(np.array_equal(x, y) for x, y in zip(back[1], big[1]))


def test_combined_operator_matches_oracle_dense():
    """Wc built on the host is exactly the operator the oracle contracts."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    # device.combine_* are pure NumPy helpers; import without touching CUDA
    from pydmrg_b200.device import combine_onesite, combine_twosite
    W1, W2 = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
    rng = np.random.default_rng(0)
    Dl, Dr, w, d = 3, 2, 6, 2
    L = rng.standard_normal((Dl, w, Dl)); R = rng.standard_normal((Dr, w, Dr))
    Wc = combine_twosite(W1, W2, np.float64).reshape(w, d * d, w, d * d)
    H = np.einsum("ijk,jaob,ros->airbks", L, Wc, R).reshape(d * d * Dl * Dr, -1)
    Hd = orc.heff_dense_twosite([L], [W1], [W2], [R], (d, d, Dl, Dr))
    assert np.allclose(H, Hd)
    Wc1 = combine_onesite(W1, w, w, d, np.float64).reshape(w, d, w, d)
    H1 = np.einsum("pnm,nojl,ijk->opilmk", L, Wc1, R).reshape(d * Dl * Dr, -1)
    assert np.allclose(H1, orc.heff_dense_onesite([L], [W1], [R], (d, Dl, Dr)))


def test_host_eig_matches_numpy(lib):
    """The small general eigen-solver inside pdmrg_davidson (host side; no GPU needed)."""
    import ctypes
    rng = np.random.default_rng(0)
    dp = ctypes.POINTER(ctypes.c_double)
    for n in (1, 2, 3, 7, 15, 18, 33):
        for real in (False, True):
            a = rng.standard_normal((n, n)) + (0 if real else 1j) * rng.standard_normal((n, n))
            a = a.astype(np.complex128)
            buf = np.ascontiguousarray(a.T).view(np.float64).ravel()          # column-major, interleaved
            w = np.zeros(2 * n); v = np.zeros(2 * n * n)
            assert lib.lib.pdmrg_host_eig(n, buf.ctypes.data_as(dp), w.ctypes.data_as(dp), v.ctypes.data_as(dp)) == 0
            wc = w.view(np.complex128); vc = v.view(np.complex128).reshape(n, n).T
            assert np.abs(a @ vc - vc * wc).max() < 1e-11 * max(1.0, np.abs(a).max())
            ref = np.linalg.eigvals(a)
            assert max(min(abs(x - ref)) for x in wc) < 1e-9 * max(1.0, abs(ref).max())
            assert np.allclose(np.linalg.norm(vc, axis=0), 1.0)


def test_recycle_window_bookkeeping():
    """Host-side bookkeeping of the warm-started truncation: number of spare vectors per cut."""
    from pydmrg_b200 import truncate
    assert truncate.spare_count(1024, 2048, 2048) == 128
    assert truncate.spare_count(256, 512, 512) == 32
    assert truncate.spare_count(8, 64, 64) == 4                 # at least four spares ...
    assert truncate.spare_count(60, 64, 200) == 4               # ... but never beyond the cut's rank
    assert truncate.spare_count(64, 64, 200) == 0               # nothing is discarded: no spares, no window
    # start policy: once exactly with nothing discarded, or twice exactly, at the CURRENT shape
    cache = {}
    assert not truncate.recycle_allowed(cache, 7, (2048, 2048, 1024))
    truncate.note_exact_visit(cache, 7, (2048, 2048, 1024), 1e-9)
    assert not truncate.recycle_allowed(cache, 7, (2048, 2048, 1024))
    truncate.note_exact_visit(cache, 7, (2048, 2048, 1024), 1e-9)
    assert truncate.recycle_allowed(cache, 7, (2048, 2048, 1024))
    assert not truncate.recycle_allowed(cache, 7, (2048, 2048, 512))      # the bond changed shape: start over
    truncate.note_exact_visit(cache, 8, (512, 512, 256), 1e-20)
    assert truncate.recycle_allowed(cache, 8, (512, 512, 256))


def test_new_entry_points_marshal_and_validate_without_gpu(lib):
    """The ctypes signatures of the entry points added for sharded sweeps and the preconditioner match the library
    (a wrong arity / type raises in ctypes), and their host-side argument checks answer before any launch."""
    import ctypes
    L = lib.lib
    e_host = (ctypes.c_double * 2)()
    n_mv, cyc, last = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
    args = (None, None, None, -1.0, 1, 8, None, 8, 1, 1, 12, 1e-14, -1.0, 10, 0, None, None, None, 8, None, 0, None, 0,
            e_host, None, 8, ctypes.byref(n_mv), ctypes.byref(cyc), ctypes.byref(last), None)
    assert L.pdmrg_davidson(*args) != 0 and b"null plan" in L.pdmrg_last_error()
    assert L.pdmrg_davidson_precond(*args, None, 0.1) != 0 and b"diagonal" in L.pdmrg_last_error()
    rc = L.pdmrg_env_update_right_cols(1, 2, 8, 8, 3, 3, None, None, None, None, 6, 4, None, None, 0, None)
    assert rc != 0 and b"ket range" in L.pdmrg_last_error()
    rc = L.pdmrg_env_update_left_cols(7, 2, 8, 8, 3, 3, None, None, None, None, 0, 4, None, None, 0, None)
    assert rc != 0 and b"dtype" in L.pdmrg_last_error()
    rc = L.pdmrg_env_update_right_cols(1, 2, 8, 8, 3, 3, None, None, None, None, 0, 8, None, None, 0, None)
    assert rc != 0 and b"workspace too small" in L.pdmrg_last_error()
    rc = L.pdmrg_vec_precond(1, 16, None, None, 0.0, 0.0, 0.0, None, None, 1 << 24, None)
    assert rc != 0 and b"floor" in L.pdmrg_last_error()
    rc = L.pdmrg_heff_diag(5, 4, 8, 8, 3, 3, None, None, None, -1.0, 0, None, None)
    assert rc != 0 and b"dtype" in L.pdmrg_last_error()
    assert L.pdmrg_launch_count() == 0
