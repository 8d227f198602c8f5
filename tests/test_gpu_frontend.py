"""the step in front of the hot path on the device -- reference tables from a polyline
(utils_coordinate_system.py:203-207) and the Frenet initial state (planner.py:567-635) -- against numpy / against the
reference's own function (tests/golden/ref_initial_states.npz)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import GOLDEN_DIR, load_golden, compare_with_oracle, rel_err
from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn
from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem

pytestmark = pytest.mark.gpu


def test_device_reference_tables_equal_numpy():
    from frenetix_motion_planner_b200 import _capi
    g = np.load(os.path.join(GOLDEN_DIR, "ref_initial_states.npz"))
    h = _capi.Handler(0)
    h.set_params(dt=0.1, N=30, low_vel_mode=False, draw_traj_set=True, kinematic_debug=True, cost_names=[], cost_weights=[],
                 x0_orientation=0.0, desired_velocity=5.0, **{k: syn.VEHICLE_2[k] for k in
                                                               ("a_max", "v_switch", "delta_max", "wheelbase", "wb_rear_axle", "length", "width")})
    for name in ("arc", "scurve", "straight", "tjunction"):
        poly = g[f"{name}_polyline"]
        cs = CoordinateSystem(poly)
        tab = h.set_reference_polyline(poly)
        assert np.array_equal(tab[4], poly[:, 0]) and np.array_equal(tab[5], poly[:, 1])
        assert np.array_equal(tab[0], cs.ref_pos)                                   # sqrt and the running sum are exact
        assert np.allclose(tab[1], cs.ref_theta, rtol=0, atol=1e-14)                # atan2 of another math library
        scale = np.abs(cs.ref_curv).max() + 1e-12
        assert np.abs(tab[2] - cs.ref_curv).max() <= 1e-12 * max(1.0, scale)
        assert np.abs(tab[3] - cs.ref_curv_d).max() <= 1e-11 * max(1.0, np.abs(cs.ref_curv_d).max())


def test_device_initial_state_equals_the_reference_code():
    from frenetix_motion_planner_b200 import _capi
    g = np.load(os.path.join(GOLDEN_DIR, "ref_initial_states.npz"))
    h = _capi.Handler(0)
    n = 0
    for name in ("arc", "scurve", "straight", "tjunction"):
        h.set_reference_polyline(g[f"{name}_polyline"])
        for row in g[f"{name}_cases"]:
            lon, lat = h.initial_state(row[0], row[1], row[2], row[3], row[4], row[5], bool(row[6]), syn.VEHICLE_2["wheelbase"])
            got, want = np.array(lon + lat), row[7:]
            assert np.all(np.abs(got - want) <= 1e-9 * np.maximum(1.0, np.abs(want))), (name, got, want)
            n += 1
    assert n == 48
    # driving against the reference direction: the reference raises (planner.py:607-609)
    h.set_reference_polyline(g["straight_polyline"])
    with pytest.raises(_capi.FrxError, match="Curvilinear velocity is negative"):
        h.initial_state(20.0, 0.5, np.pi, 5.0, 0.0, 0.0, False, syn.VEHICLE_2["wheelbase"])


def test_planner_with_device_frontend_matches_oracle_on_the_device_tables():
    """Polyline + Cartesian ego state in, plan out -- no host geometry; the oracle evaluated on the SAME (device-built)
    tables must agree on every mask and on the selected candidate."""
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    fx = np.load(os.path.join(GOLDEN_DIR, "tjunction.npz"))
    g, ref, prm, preds = load_golden("tjunction_draw")
    cfg_plan = SimpleNamespace(
        planning=SimpleNamespace(planning_horizon=3.0, dt=0.1, low_vel_mode_threshold=2.0, sampling_min=2, sampling_max=3,
                                 t_min=1.1, d_min=-3, d_max=3, d_ego_pos=False),
        debug=SimpleNamespace(multiproc=True, num_workers=6, draw_traj_set=True, kinematic_debug=True, save_all_traj=False,
                              log_risk=False, device_frontend=True),
        cost=SimpleNamespace(cost_weights=dict(prm.cost_weights)))
    p = ReactivePlannerB200(cfg_plan, SimpleNamespace(vehicle=SimpleNamespace(**syn.VEHICLE_2)), None, None, None, None, None)
    x_0 = SimpleNamespace(position=fx["ego_position_rear"], orientation=float(fx["ego_orientation"]), velocity=float(fx["ego_velocity"]),
                          acceleration=float(fx["ego_acceleration"]), yaw_rate=float(fx["ego_yaw_rate"]), steering_angle=0.0, time_step=0)
    p.update_externals(reference_path=fx["reference_path"], x_0=x_0, x_cl=None, desired_velocity=8.0,
                       predictions={100 + i: q for i, q in enumerate(preds)})
    assert np.allclose(p.x_cl[0], g["x_cl_lon"], rtol=1e-9) and np.allclose(p.x_cl[1], g["x_cl_lat"], rtol=1e-9, atol=1e-12)
    p.obstacle_order = [100 + i for i in range(5)]
    p.plan()
    cs = p.coordinate_system
    ref_dev = fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, np.ascontiguousarray(fx["reference_path"][:, 0]),
                         np.ascontiguousarray(fx["reference_path"][:, 1]))
    S = p._sampling_matrix(2)
    ora = fo.plan(S, ref_dev, prm, preds)
    b = p._bundle
    dev = dict(flags=b.flags, traj_len=b.traj_len, costs=b.costs, total=b.total, states=b.states(np.arange(S.shape[0])),
               argmin=p.last_plan_stats.argmin, min_cost=p.last_plan_stats.min_cost, reason_counts=p.last_plan_stats.reason_counts,
               n_in_list=p.last_plan_stats.n_in_list, n_feasible=p.last_plan_stats.n_feasible,
               collision_counter=p.last_plan_stats.collision_counter)
    compare_with_oracle(dev, ora, prm)
    assert p.optimal_trajectory.uniqueId == ora["argmin"] or ora["margins"][ora["argmin"]] < 1e-9
