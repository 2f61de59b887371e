"""Dry-run of the Python code of tests/test_pending_gpu.py on the CPU emulation (tests/emu_device.py): catches typos and
unreachable tolerances before GPU time is spent on them.  Not a test: it says nothing about the device.

    python tests/experiments/dryrun_pending_on_emulation.py"""
import os, sys, types, traceback
os.environ["PDMRG_RUN_PENDING"] = "1"; os.environ["PDMRG_PY_DAVIDSON"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import emu_device
emu_device.install(setattr)
from pydmrg_b200 import core
class Ctx(emu_device.Context):
    def __init__(self, dtype="c128", **kw): super().__init__(dtype)
core.Context = Ctx
core.default_context = lambda dtype="c128": Ctx(dtype)
import pydmrg_b200.dmrg as dm, pydmrg_b200.contract as ct, pydmrg_b200.gauge as gg, pydmrg_b200.eigs as eg, pydmrg_b200.env as ev
for m in (dm, ct, gg, eg, ev):
    if hasattr(m, "default_context"): m.default_context = core.default_context
    if hasattr(m, "Context"): m.Context = Ctx
torch.cuda.synchronize = lambda *a, **k: None
import test_pending_gpu as T
import pydmrg_b200 as pkg
from pydmrg_b200 import core as _c, device, dmrg, mpo, parallel, truncate, contract, eigs
P = pkg
ok = True
def run(name, *args):
    global ok
    try:
        getattr(T, name)(P, *args); print("ok  ", name, args)
    except Exception as ex:
        ok = False; print("FAIL", name, args, type(ex).__name__, str(ex)[:300]); traceback.print_exc(limit=-4)
for tag in ("asep6", "asep8"): run("test_two_site_two_states_gap", tag)
for d in ("right", "left"):
    for dt in ("c128", "f64"): run("test_state_averaged_twosite_truncation", d, dt)
for tag in ("tasep8", "asep8"): run("test_left_right_run_matches_reference_fixture", tag)
for dt in ("c128", "f64"):
    for dims in ((2, 96, 80), (2, 64, 64), (2, 40, 200)): run("test_column_restricted_block_updates", dt, dims)
for dt in ("c128", "f64"):
    for m in (0, 5, 16, 17, 28): run("test_project_with_more_than_sixteen_vectors", dt, m)
for model in ("heisenberg_f64", "asep_c128"): run("test_gauge_shortcut_and_solver_options", model)
print("ALL OK" if ok else "SOME FAILED")
