"""Model MPO data builders in PyDMRG's layout -- list of operators, each a list of N
ndarrays ``(wl, wr, d_out, d_in)`` (first site ``(1,w,d,d)``, last ``(w,1,d,d)``).

The product consumes MPOs built by the reference's own ``mpo.<model>.return_mpo`` unchanged;
these builders exist because the reference is not present on the GPU box (tests / bench need
the TASEP / ASEP / Heisenberg operators named in BASELINE.json).  Element-for-element
equality with ``mpo/asep.py`` (open_mpo :36-68, open_curr :202-228) is asserted in
``tests/test_oracle_vs_reference.py``.  The reference has no Heisenberg model
and its ``mpo/tasep.py`` does not import: TASEP is ASEP with gamma=q=delta=0.
"""
import numpy as np

_OPS = {
    "I": np.eye(2),
    "Sp": np.array([[0., 1.], [0., 0.]]),      # mpo/ops.py:4-5
    "Sm": np.array([[0., 0.], [1., 0.]]),
    "n": np.array([[0., 0.], [0., 1.]]),
    "v": np.array([[1., 0.], [0., 0.]]),
    "Sz": np.array([[0.5, 0.], [0., -0.5]]),
}


def _assemble(N, w, site_entries):
    """site_entries(site) -> dict {(row, col): 2x2 array}; corner convention: the last row
    and the first column carry the running operators."""
    ops = []
    for site in range(N):
        W = np.zeros((w, w, 2, 2))
        for (r, c), m in site_entries(site).items():
            W[r, c] += m
        if site == 0:
            W = W[w - 1:w, :]
        elif site == N - 1:
            W = W[:, 0:1]
        ops.append(np.ascontiguousarray(W))
    return [ops]


def asep_mpo(N, params, periodic=False):
    """Open-boundary ASEP tilted generator, params = (alpha, gamma, p, q, beta, delta, s)
    (scalars; alpha/gamma act on site 0, beta/delta on site N-1).  w = 6.
    ``periodic=True`` appends the reference's wrap-around terms as extra w = 1 operators with
    ``None`` (identity) entries in between (mpo/asep.py:70-94)."""
    a, g, p, q, b, d, s = [float(t) for t in params]
    O = _OPS
    es, ems = np.exp(s), np.exp(-s)

    def entries(site):
        av = a if site == 0 else 0.0
        gv = g if site == 0 else 0.0
        bv = b if site == N - 1 else 0.0
        dv = d if site == N - 1 else 0.0
        corner = (av * es + dv * ems) * O["Sm"] - (av + dv) * O["v"] \
            + (bv * es + gv * ems) * O["Sp"] - (bv + gv) * O["n"]
        return {(0, 0): O["I"], (1, 0): p * es * O["Sm"], (2, 0): p * O["v"],
                (3, 0): q * ems * O["Sp"], (4, 0): q * O["n"], (5, 0): corner,
                (5, 1): O["Sp"], (5, 2): -O["n"], (5, 3): O["Sm"], (5, 4): -O["v"], (5, 5): O["I"]}

    mpoL = _assemble(N, 6, entries)
    if periodic:
        def string(first, last):
            op = [None] * N
            op[0] = first.reshape(1, 1, 2, 2).copy()
            op[-1] = last.reshape(1, 1, 2, 2).copy()
            return op
        if p != 0:
            mpoL.append(string(O["Sm"], p * es * O["Sp"]))
            mpoL.append(string(O["v"], -p * O["n"]))
        if q != 0:
            mpoL.append(string(O["Sp"], q * ems * O["Sm"]))
            mpoL.append(string(O["n"], -q * O["v"]))
    return mpoL


def tasep_mpo(N, alpha, beta, s):
    """TASEP (unit forward hopping) through the ASEP builder."""
    return asep_mpo(N, (alpha, 0.0, 1.0, 0.0, beta, 0.0, s))


def asep_curr_mpo(N, params):
    """Total-current observable MPO for the open ASEP, w = 4."""
    a, g, p, q, b, d, s = [float(t) for t in params]
    O = _OPS
    es, ems = np.exp(s), np.exp(-s)

    def entries(site):
        av = a if site == 0 else 0.0
        gv = g if site == 0 else 0.0
        bv = b if site == N - 1 else 0.0
        dv = d if site == N - 1 else 0.0
        corner = (av * es - dv * ems) * O["Sm"] + (bv * es - gv * ems) * O["Sp"]
        return {(0, 0): O["I"], (1, 0): p * es * O["Sm"], (2, 0): -q * ems * O["Sp"], (3, 0): corner,
                (3, 1): O["Sp"], (3, 2): O["Sm"], (3, 3): O["I"]}

    return _assemble(N, 4, entries)


def heisenberg_mpo(N, J=1.0, Jz=1.0, h=0.0, sign=-1.0):
    """XXZ chain  H = sum_i J/2 (S+_i S-_{i+1} + S-_i S+_{i+1}) + Jz Sz_i Sz_{i+1} - h Sz_i,
    w = 5.  PyDMRG's solver targets the eigenvalue with the LARGEST real part
    (tools/diag_tools.py:125,175-179,265), so the ground state needs the MPO of -H:
    ``sign=-1`` (default) multiplies H by -1; pass ``sign=+1`` for +H."""
    O = _OPS

    def entries(site):
        return {(0, 0): O["I"], (1, 0): O["Sm"], (2, 0): O["Sp"], (3, 0): O["Sz"],
                (4, 0): -h * sign * O["Sz"], (4, 1): sign * J / 2 * O["Sp"], (4, 2): sign * J / 2 * O["Sm"],
                (4, 3): sign * Jz * O["Sz"], (4, 4): O["I"]}

    return _assemble(N, 5, entries)


def mpo_conj_trans(mpo):
    """W -> W^dagger on the physical indices (None entries kept): the operator whose
    right eigenvector is the LEFT eigenvector of ``mpo`` (tools/mpo_tools.py:22-32)."""
    return [[None if W is None else np.ascontiguousarray(np.conj(np.transpose(W, (0, 1, 3, 2)))) for W in op]
            for op in mpo]
