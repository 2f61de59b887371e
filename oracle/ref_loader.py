"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (philliphelms/PyDMRG).

Imports the reference package from ``/root/reference/pydmrg`` with zero edits, by
installing four in-memory shims first:

1. NumPy / stdlib aliases the 2018-era code expects (``np.complex_``, ``np.float_``,
   ``collections.Sequence``) -- used at e.g. pydmrg/dmrg.py:341, pydmrg/mpo/asep.py:26.
2. ``cyclomps.tools.utils`` -- an un-vendored, un-pinned third-party module that
   pydmrg/tools/la_tools.py:12 star-imports.  It only re-exports NumPy primitives
   (call sites la_tools.py:20,28-30,451,459-464,591-601,638-639); the Davidson arithmetic
   itself (``davidson_nosym1``, la_tools.py:402-557) lives in the reference.
3. ``h5py`` -- imported at pydmrg/tools/mps_tools.py:6; when it is not installed a mini-shim
   (File / create_group / create_dataset / get over one .npz archive) stands in, enough for the
   file-based flows of the reference (``fname=...``, ``calcLeftState=True``).
4. ``pyscf.lib`` for pydmrg/idmrg.py:3-4 with ``eig = tools.la_tools.eig`` (la_tools is
   the reference's adapted copy of that routine) and ``einsum = np.einsum``.

``/root/reference`` exists only in the build container, so this loader is used by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that skip when
the path is absent.  Nothing in the product package imports it.
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PYDMRG_REFERENCE", "/root/reference/pydmrg")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "tools"))


def _install_shims():
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence
    if "cyclomps.tools.utils" not in sys.modules:
        pkg = types.ModuleType("cyclomps")
        tools = types.ModuleType("cyclomps.tools")
        utils = types.ModuleType("cyclomps.tools.utils")
        for name in ("sqrt", "zeros", "dot", "conj", "real", "take", "transpose", "abs",
                     "argsort", "ones", "eye", "array", "einsum"):
            setattr(utils, name, getattr(np, name))
        utils.qr = np.linalg.qr
        utils.summ = np.sum
        utils.USE_CTF = False
        utils.to_nparray = np.asarray
        utils.from_nparray = lambda a: a
        utils.__all__ = [k for k in vars(utils) if not k.startswith("__")]
        pkg.tools = tools
        tools.utils = utils
        sys.modules["cyclomps"] = pkg
        sys.modules["cyclomps.tools"] = tools
        sys.modules["cyclomps.tools.utils"] = utils
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules["h5py"] = _h5py_shim()


def _h5py_shim():
    """The few h5py calls of pydmrg/tools/mps_tools.py:320-373 (File as a context manager, create_group,
    create_dataset(name, data=...), get(path)) over one .npz archive written at the same path -- enough for the
    reference's file-based flows (run_dmrg(fname=...), calcLeftState, chi schedules) where h5py is not installed."""
    mod = types.ModuleType("h5py")

    class _Group:
        def __init__(self, store, prefix):
            self._store, self._prefix = store, prefix

        def create_group(self, name):
            return _Group(self._store, self._prefix + name + "/")

        def create_dataset(self, name, data=None, **kw):
            self._store[self._prefix + name] = np.asarray(data)

    class File(_Group):
        def __init__(self, path, mode="r"):
            _Group.__init__(self, {}, "")
            self._path, self._mode = path, mode
            if mode == "r":
                with np.load(path) as z:
                    self._store = {k.replace("|", "/"): z[k] for k in z.files}

        def get(self, key):
            return self._store.get(key)

        def __getitem__(self, key):
            return self._store[key]

        def __contains__(self, key):                            # a dataset, or a group (prefix of dataset paths)
            return key in self._store or any(k.startswith(key + "/") for k in self._store)

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            if self._mode != "r":
                with open(self._path, "wb") as fh:
                    np.savez(fh, **{k.replace("/", "|"): v for k, v in self._store.items()})
            return False

    mod.File = File
    return mod


_REF = None


def load(with_idmrg=True):
    """Return a namespace with the reference modules: dmrg, idmrg, diag_tools, env_tools,
    mps_tools, la_tools, mpo_tools, contract, ed, asep."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.la_tools = importlib.import_module("tools.la_tools")
    ns.mps_tools = importlib.import_module("tools.mps_tools")
    ns.diag_tools = importlib.import_module("tools.diag_tools")
    ns.env_tools = importlib.import_module("tools.env_tools")
    ns.mpo_tools = importlib.import_module("tools.mpo_tools")
    ns.contract = importlib.import_module("tools.contract")
    ns.ed = importlib.import_module("tools.ed")
    ns.asep = importlib.import_module("mpo.asep")
    ns.dmrg = importlib.import_module("dmrg")
    if with_idmrg:
        if "pyscf.lib" not in sys.modules:
            pyscf = types.ModuleType("pyscf")
            lib = types.ModuleType("pyscf.lib")
            lib.eig = ns.la_tools.eig
            lib.einsum = np.einsum
            pyscf.lib = lib
            sys.modules["pyscf"] = pyscf
            sys.modules["pyscf.lib"] = lib
        ns.idmrg = importlib.import_module("idmrg")
    # silence the reference's print-based logging
    ns.dmrg.VERBOSE = 0
    ns.diag_tools.VERBOSE = 0
    if with_idmrg:
        ns.idmrg.VERBOSE = 0
    _REF = ns
    return ns
