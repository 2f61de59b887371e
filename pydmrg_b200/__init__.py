"""pydmrg_b200 -- B200-native (sm_100a) implementation of PyDMRG's DMRG sweep hot path.

Public surface (names and argument meaning follow philliphelms/PyDMRG):
    run_dmrg, run_idmrg, rightStep, leftStep, single_iter, calc_eigs, make_ham_func,
    update_envR, update_envL, calc_env, renorm_inf, full_contract, load_mps, save_mps.
The CUDA extension ``libpydmrg_b200.so`` is mandatory: importing any compute module without it
raises (there is no CPU fallback).  ``pydmrg_b200.mpo`` / ``pydmrg_b200.mps`` are NumPy-only.
"""
__version__ = "0.1.0"

_LAZY = {
    "run_dmrg": "dmrg", "rightStep": "dmrg", "leftStep": "dmrg", "twoSiteStep": "dmrg", "run_sweeps": "dmrg",
    "run_idmrg": "idmrg", "single_iter": "idmrg",
    "calc_eigs": "eigs", "make_ham_func": "eigs",
    "update_envR": "env", "update_envL": "env", "calc_env": "env", "alloc_env": "env",
    "renorm_inf": "truncate", "full_contract": "contract",
    "load_mps": "mps", "save_mps": "mps", "Context": "core",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError("module 'pydmrg_b200' has no attribute %r" % name)
    import importlib
    return getattr(importlib.import_module("." + mod, __name__), name)
