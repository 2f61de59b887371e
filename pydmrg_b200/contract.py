"""<l|O|r> contraction of saved / returned states on the device: N successive right-block
updates, exactly how pydmrg/tools/contract.py:5-44 builds it out of update_envL.
Used for currents, densities and <l|H|r> (tests/asep/currentCheck.py:29-44)."""
import numpy as np

from . import device as D
from . import mps as mps_tools
from .core import default_context


def full_contract(mpo=None, mps=None, lmps=None, state=None, lstate=None, orth=False, gSite=None, glSite=None,
                  ctx=None):
    """Same arguments as the reference.  ``mps`` / ``lmps``: mpsList (or a file name);
    ``lmps`` holds the bra tensors as stored (default conj(mps), contract.py:12-13);
    ``mpo=None`` gives the overlap (contract.py:33-34)."""
    ctx = ctx or default_context()
    if isinstance(mps, str):
        mps, gSite = mps_tools.load_mps(mps)
    if isinstance(lmps, str):
        lmps, glSite = mps_tools.load_mps(lmps)
    assert not (lmps is None and mps is None)
    if orth:                                                     # contract.py:16-19
        host = lambda L: [[np.array(ctx.host(t)) for t in M] for M in L]
        if mps is not None:
            mps = mps_tools.orthonormalize_states(host(mps), gSite=gSite)
        if lmps is not None:
            lmps = mps_tools.orthonormalize_states(host(lmps), gSite=glSite)
    if state is None and lstate is None:
        state, lstate = 0, 0
    elif state is None:
        state = lstate
    elif lstate is None:
        lstate = state
    if mps is None:                                              # contract.py:14-15
        ket = [ctx.dev(np.conj(ctx.host(t))) for t in lmps[lstate]]
        bra = [ctx.dev(t) for t in lmps[lstate]]
    else:
        ket = [ctx.dev(t) for t in mps[state]]
        bra = None if lmps is None else [ctx.dev(t) for t in lmps[lstate]]
    N = len(ket)
    if mpo is None:
        mpo = [[None] * N]
    result = 0
    for op in mpo:
        R = ctx.dev(np.ones((1, 1, 1)))
        for site in range(N - 1, -1, -1):
            R = D.env_update_left(ket[site], op[site], R, None if bra is None else bra[site], ctx.ws)
        result = result + R.reshape(-1)[0].item()
    return result
