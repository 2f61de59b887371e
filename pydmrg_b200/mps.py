"""MPS container utilities kept at the reference's boundary (host side; SURVEY.md section 2
marks these "out of scope for kernels": initial states, chi ramps, save / load).  Same names
and layouts as pydmrg/tools/mps_tools.py so that initial guesses and saved states plug in
unchanged:  ``mpsList[state][site]`` is an ndarray (d, Dl, Dr)."""
import os

import numpy as np


def create_rand_mps(N, mbd, d=2, const=1e-2):
    """tools/mps_tools.py:211-221 -- identical np.random draw order, so np.random.seed(k)
    reproduces the reference's initial state."""
    M = []
    for i in range(N // 2):
        M.append(const * np.random.rand(d, min(d ** i, mbd), min(d ** (i + 1), mbd)))
    if N % 2 == 1:
        i = N // 2 - 1
        M.append(const * np.random.rand(d, min(d ** (i + 1), mbd), min(d ** (i + 1), mbd)))
    for i in range(N // 2 - 1, -1, -1):
        M.append(const * np.random.rand(d, min(d ** (i + 1), mbd), min(d ** i, mbd)))
    return M


def create_const_mps(N, mbd, d=2, const=1e-2):
    """tools/mps_tools.py:223-233."""
    M = []
    for i in range(N // 2):
        M.append(const * np.ones((d, min(d ** i, mbd), min(d ** (i + 1), mbd))))
    if N % 2 == 1:
        i = N // 2 - 1
        M.append(const * np.ones((d, min(d ** (i + 1), mbd), min(d ** (i + 1), mbd))))
    for i in range(N // 2 - 1, -1, -1):
        M.append(const * np.ones((d, min(d ** (i + 1), mbd), min(d ** i, mbd))))
    return M


def create_all_mps(N, mbd, nStates, rand=True, const=1e-2, d=2):
    """tools/mps_tools.py:235-242."""
    return [(create_rand_mps if rand else create_const_mps)(N, mbd, d=d, const=const) for _ in range(nStates)]


def make_mps_right(M):
    """Right-canonical form by SVD from the right.  tools/mps_tools.py:36-46."""
    for i in range(len(M) - 1, 0, -1):
        m = np.swapaxes(M[i], 0, 1)
        n1, n2, n3 = m.shape
        U, s, V = np.linalg.svd(m.reshape(n1, n2 * n3), full_matrices=False)
        M[i] = np.swapaxes(V.reshape(n1, n2, n3), 0, 1)
        M[i - 1] = np.einsum("klj,ji,i->kli", M[i - 1], U, s)
    return M


def make_all_mps_right(mpsList):
    return [make_mps_right(M) for M in mpsList]


def conj_mps(mps):
    """tools/mps_tools.py:54-62."""
    return [[np.conj(t) for t in M] for M in mps]


def orthonormalize_states(mps, mpo=None, gSite=None, printEnergies=False):
    """tools/mps_tools.py:392-430: the centre tensors of all states (every other site tensor is shared, they are
    the state-averaged isometries) are replaced by an orthonormal basis of their span -- ``scipy.linalg.orth`` of the
    (d*Dl*Dr, nStates) matrix, exactly the reference's call.  Host-side post-processing of saved / returned states
    (the energy print-outs of the reference are omitted; ``mpo`` is accepted for signature compatibility)."""
    import scipy.linalg as sla
    if isinstance(mps, str):
        mps, gSite = load_mps(mps)
    gSite = int(gSite)
    nStates = len(mps)
    shape = np.shape(mps[0][gSite])
    vecs = np.zeros((int(np.prod(shape)), nStates), dtype=np.complex128)
    for state in range(nStates):
        vecs[:, state] = np.reshape(mps[state][gSite], -1)
    vecs = sla.orth(vecs, rcond=1e-100)                                  # :416
    for state in range(nStates):
        mps[state][gSite] = np.reshape(vecs[:, state], shape)
    return mps


def increase_mbd(M, mbd, d=2):
    """Zero-pad every site tensor to the bond profile of a larger chi (open chain).
    tools/mps_tools.py:244-279 (periodic=False, constant=False branch)."""
    N = len(M)
    for site in range(N):
        dl = min(d ** min(site, N - site), mbd)
        dr = min(d ** min(site + 1, N - site - 1), mbd)
        _, ny, nz = M[site].shape
        M[site] = np.pad(M[site].astype(np.complex128), ((0, 0), (0, max(dl - ny, 0)), (0, max(dr - nz, 0))))
    return M


def increase_all_mbd(mpsL, mbd, d=2):
    return [increase_mbd(M, mbd, d=d) for M in mpsL]


def nSites(mpsL):
    return len(mpsL[0])


def maxBondDim(mpsL):
    return max(max(t.shape) for t in mpsL[0])


# ---- checkpoint / resume: the reference's on-disk layouts -----------
def save_mps_npz(mpsL, fname, gaugeSite=0):
    """One file per state ``fname + 'state{s}.npz'`` with keys M0..M{N-1}, site
    (tools/mps_tools.py:355-362)."""
    for state, M in enumerate(mpsL):
        np.savez(fname + "state" + str(state) + ".npz", site=gaugeSite,
                 **{"M" + str(i): np.asarray(t) for i, t in enumerate(M)})


def load_mps_npz(fname):
    """tools/mps_tools.py:287-316."""
    mpsL, state, site = [], 0, 0
    while os.path.exists(fname + "state" + str(state) + ".npz"):
        f = np.load(fname + "state" + str(state) + ".npz")
        M, i = [], 0
        while "M" + str(i) in f:
            M.append(f["M" + str(i)])
            i += 1
        site = f["site"]
        mpsL.append(M)
        state += 1
    if not mpsL:
        raise FileNotFoundError(fname + "state0.npz")
    return mpsL, site


def save_mps_hdf5(mpsL, fname, gaugeSite=0, comp_opts=4):
    """groups state{s}/, dataset state{s}/site, state{s}/M{i}/real|imag float64 (d,Dl,Dr), gzip
    (tools/mps_tools.py:364-373).  Needs h5py."""
    import h5py
    with h5py.File(fname + ".hdf5", "w") as f:
        for state, M in enumerate(mpsL):
            g = f.create_group("state" + str(state))
            g.create_dataset("site", data=gaugeSite)
            for i, t in enumerate(M):
                t = np.asarray(t)
                g.create_dataset("M%d/real" % i, data=np.real(t), compression="gzip", compression_opts=comp_opts)
                g.create_dataset("M%d/imag" % i, data=np.imag(t), compression="gzip", compression_opts=comp_opts)


def load_mps_hdf5(fname):
    """tools/mps_tools.py:318-347."""
    import h5py
    mpsL = []
    with h5py.File(fname + ".hdf5", "r") as f:
        state = 0
        while "state%d" % state in f:
            M, i = [], 0
            while "state%d/M%d/real" % (state, i) in f:
                M.append(np.array(f["state%d/M%d/real" % (state, i)]) + 1j * np.array(f["state%d/M%d/imag" % (state, i)]))
                i += 1
            mpsL.append(M)
            state += 1
        site = np.array(f["state0/site"])
    return mpsL, site


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return hasattr(h5py, "File")
    except Exception:
        return False


def save_mps(mpsL, fname, gaugeSite=0, fformat=None, comp_opts=4):
    """tools/mps_tools.py:375-381.  The reference always writes hdf5; where h5py is not
    installed the NPZ layout (which the reference also reads, load_mps(fformat='npz')) is
    used instead."""
    if fname is None:
        return
    if fformat is None:
        fformat = "hdf5" if _have_h5py() else "npz"
    if fformat == "npz":
        save_mps_npz(mpsL, fname, gaugeSite)
    else:
        save_mps_hdf5(mpsL, fname, gaugeSite, comp_opts)


def load_mps(fname, fformat=None):
    """tools/mps_tools.py:349-353."""
    if fformat is None:
        fformat = "hdf5" if (os.path.exists(fname + ".hdf5") and _have_h5py()) else "npz"
    return load_mps_npz(fname) if fformat == "npz" else load_mps_hdf5(fname)
