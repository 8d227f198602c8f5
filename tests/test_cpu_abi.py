"""The host-core implementation of the C ABI (oracle/cpu_abi, SURVEY.md 2.2 / 8b: "the same ABI implemented by the CPU/OpenMP
lib"; baseline infrastructure, never loaded by the package): exports every symbol of include/frx.h and answers through the
ctypes binding exactly what the C oracle answers when called directly."""
import os
import re

import numpy as np
import pytest

from helpers import load_golden, configure_handler
from oracle import c_oracle
from oracle.build import build_cpu_abi
from frenetix_motion_planner_b200 import _capi, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpu_lib():
    return _capi.load_library(path=build_cpu_abi())


def test_cpu_build_exports_the_whole_abi(cpu_lib):
    hdr = open(os.path.join(ROOT, "include", "frx.h")).read()
    declared = set(re.findall(r"\b(frx_[a-z_0-9]+)\s*\(", hdr)) - {"frx_ctx"}
    for name in sorted(declared):
        assert hasattr(cpu_lib, name), name
    assert cpu_lib.frx_abi_version() == _capi.load_library.__globals__["C"].CDLL(_capi.LIB_PATH).frx_abi_version() \
        if os.path.exists(_capi.LIB_PATH) else True


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_lowvel_nodraw", "tjunction_nodraw"])
def test_cpu_abi_equals_the_c_oracle(cpu_lib, name):
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(g["sampling"])
    h = _capi.Handler(0, library=cpu_lib)
    configure_handler(h, ref, prm, preds, None, sampling=S)
    res = h.plan(S, row_index_base=1000)
    want = c_oracle.plan(S, ref, prm, preds, check_all_collisions=False)
    assert int(res.argmin) == want["argmin"] + 1000 and float(res.min_cost) == want["min_cost"]
    flags, tl = h.get_flags()
    costs, total = h.get_costs()
    assert np.array_equal(flags, want["flags"]) and np.array_equal(tl, want["traj_len"])
    assert np.array_equal(costs, want["costs"]) and np.array_equal(total, want["total"])
    assert np.array_equal(h.get_states_range(), want["states"])
    assert np.array_equal(h.winner_states(), want["states"][:, want["argmin"], :])
    fl, wtl, tot, wc = h.winner_record()
    assert (fl, wtl, tot) == (int(want["flags"][want["argmin"]]), int(want["traj_len"][want["argmin"]]), want["min_cost"])
    assert int(res.n_feasible) == want["n_feasible"] and int(res.collision_counter) == want["collision_counter"]
    with pytest.raises(_capi.FrxError, match="not available in the CPU build"):
        h.plan_device(0, 10)


def test_cpu_abi_grid_mode_equals_matrix_mode(cpu_lib):
    g, ref, prm, preds = load_golden("arc_hv_draw_pred")
    t1, v1, d1 = np.array([1.1, 2.0, 3.0]), np.linspace(2.0, 12.0, 5), np.linspace(-2.0, 2.0, 6)
    x_cl = (list(g["x_cl_lon"]), list(g["x_cl_lat"]))
    S = syn.grid_sampling_matrix(t1, v1, d1, x_cl)
    h = _capi.Handler(0, library=cpu_lib)
    configure_handler(h, ref, prm, preds, None, sampling=S)
    a = h.plan(S)
    fa, _ = h.get_flags()
    b = h.plan_grid(t1, v1, d1, x_cl)
    fb, _ = h.get_flags()
    assert (int(a.argmin), float(a.min_cost)) == (int(b.argmin), float(b.min_cost)) and np.array_equal(fa, fb)
    c = h.plan_grid(t1, v1, d1, x_cl, row_first=30, row_count=40)
    fc, _ = h.get_flags()
    assert np.array_equal(fc, fa[30:70]) and int(c.n_rows) == 40


def test_planner_class_end_to_end_on_the_cpu_backend(cpu_lib):
    """ReactivePlannerB200.plan() through the REAL ctypes binding (structs, result record, winner record, lazy read-backs)
    with the CPU implementation of the ABI behind it: same selection, statistics and trajectory pair as with the
    oracle-backed stand-in handler the drop-in tests use."""
    from test_gpu_planner import make_planner
    from oracle_handler import OracleHandler
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    g, ref, prm, preds = load_golden("tjunction_draw")
    outs = []
    for handler in (_capi.Handler(0, library=cpu_lib), OracleHandler()):
        orig = ReactivePlannerB200.__init__

        def init(self, *a, **k):
            k["handler"] = handler
            orig(self, *a, **k)
        ReactivePlannerB200.__init__ = init
        try:
            p = make_planner(g, prm, preds, 8.0)
        finally:
            ReactivePlannerB200.__init__ = orig
        p.collision_check_enabled = True
        p.obstacle_order = [100 + i for i in range(len(preds))]
        pair = p.plan()
        assert pair is not None and p.optimal_trajectory is not None
        outs.append((p.optimal_trajectory.uniqueId, p.optimal_trajectory.cost, p._total_count, list(p._infeasible_count_kinematics),
                     p.infeasible_kinematics_percentage, p.infeasible_count_collision,
                     np.array([s.position for s in pair[0].state_list]), np.array(pair[2]), np.array(pair[3]),
                     [t.uniqueId for t in p.all_traj[:40]]))
    a, b = outs
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[3] == b[3] and a[4] == b[4] and a[5] == b[5]
    assert np.array_equal(a[6], b[6]) and np.array_equal(a[7], b[7]) and np.array_equal(a[8], b[8]) and a[9] == b[9]
