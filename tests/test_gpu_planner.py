"""ReactivePlannerB200.plan() end to end on the GPU, against what the reference's own
``_create_trajectory_bundle`` + ``_get_optimal_trajectory`` produced (tests/golden)."""
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import load_golden, BAND, rel_err
from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def make_planner(g, prm, preds, v0, multiproc=False):
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    cfg_plan = SimpleNamespace(
        planning=SimpleNamespace(planning_horizon=3.0, dt=0.1, low_vel_mode_threshold=2.0, sampling_min=2, sampling_max=3,
                                 t_min=1.1, d_min=-3, d_max=3, d_ego_pos=False),
        debug=SimpleNamespace(multiproc=multiproc, num_workers=6, draw_traj_set=prm.draw_traj_set,
                              kinematic_debug=prm.kinematic_debug, save_all_traj=False, log_risk=False),
        cost=SimpleNamespace(cost_weights=dict(prm.cost_weights, acceleration=0.0, jerk=0.0)))
    cfg_sim = SimpleNamespace(vehicle=SimpleNamespace(**syn.VEHICLE_2))
    p = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None)
    x_0 = SimpleNamespace(position=None, orientation=float(g["x0_orientation"]), velocity=v0, acceleration=0.0,
                          yaw_rate=0.0, steering_angle=0.0, time_step=0)
    x_cl = (list(g["x_cl_lon"]), list(g["x_cl_lat"]))
    p.collision_check_enabled = False          # the goldens were recorded without pycrcc (make_golden.py)
    p.update_externals(reference_path=g["polyline"], x_0=x_0, x_cl=x_cl, desired_velocity=float(g["desired_velocity"]),
                       predictions={100 + i: q for i, q in enumerate(preds)} if preds else None)
    return p


@pytest.mark.parametrize("name,v0", [("straight_hv_draw", 8.0), ("arc_hv_draw_pred", 9.5), ("arc_hv_nodraw_nodebug", 9.5),
                                     ("scurve_lowvel_draw", 1.2), ("scurve_brake_hv_nodraw_debug", 3.0),
                                     ("short_hv_nodraw", 8.0)])
def test_plan_matches_reference_run(name, v0):
    g, ref, prm, preds = load_golden(name)
    p = make_planner(g, prm, preds, v0)
    pair = p.plan()
    margins = fo.plan(g["sampling"], ref, prm, preds, collision_check=False)["margins"]
    ok = margins >= BAND
    n_band = int((~ok).sum())
    # the reference's statistics (reactive_planner.py:233-235)
    assert p._total_count == g["sampling"].shape[0]
    assert abs(p._infeasible_count_kinematics[0] - g["infeasible_count_kinematics"][0]) <= n_band
    assert p._infeasible_count_kinematics[1:] == [0] * 10            # not multiproc: per-reason counts stay 0
    if n_band == 0:
        assert abs(p.infeasible_kinematics_percentage - float(g["percentage"])) < 1e-9
    # the selected trajectory
    opt = p.optimal_trajectory
    assert opt is not None and pair is not None
    if ok[opt.uniqueId] and ok[int(g["optimal_id"])]:
        assert opt.uniqueId == int(g["optimal_id"])
        r = opt.uniqueId
        assert abs(opt.cost - g["total"][r]) <= 1e-6 * max(1.0, abs(g["total"][r]))
        assert rel_err(opt.cartesian.x, g["states"][0, r]) < 1e-6 and rel_err(opt.cartesian.kappa, g["states"][5, r]) < 1e-6
        assert rel_err(opt.curvilinear.d_ddot, g["states"][13, r]) < 1e-6
        assert rel_err(opt.trajectory_long.coeffs, g["coeffs"][r, :6]) < 1e-12
        assert rel_err(opt.trajectory_lat.coeffs, g["coeffs"][r, 6:]) < 1e-12
        for k, nm in enumerate(g["cost_names"]):
            assert abs(opt.costMap[str(nm)][0] - g["costs"][r, k]) <= 1e-6 * max(1.0, abs(g["costs"][r, k]))
        cart, curv, lon, lat = pair
        assert len(cart.state_list) == 31 and cart.state_list[3].time_step == 3
        assert np.allclose(cart.state_list[5].position, [g["states"][0, r, 5], g["states"][1, r, 5]], rtol=1e-9)
        assert np.allclose(lon[7], g["states"][[7, 10, 11], r, 7], rtol=1e-9, atol=1e-12)
    # the cost-sorted trajectory set kept for visualisation (draw_traj_set)
    if prm.draw_traj_set:
        got = [t.uniqueId for t in p.all_traj]
        want = [int(i) for i in g["sorted_ids"]]
        assert [i for i in got if ok[i]] == [i for i in want if ok[i]]
        s = p.all_traj[0]
        assert s.feasible in (True, False) and s.valid in (True, False) and len(s.cartesian.x) == 31
    else:
        assert p.all_traj is None


def test_multiproc_debug_reports_reason_counters_and_resampling_loop():
    g, ref, prm, preds = load_golden("scurve_brake_hv_draw")
    p = make_planner(g, prm, preds, 3.0, multiproc=True)
    p.plan()
    c = p._infeasible_count_kinematics
    ora = fo.plan(g["sampling"], ref, prm, preds, collision_check=False)
    n_band = int((ora["margins"] < BAND).sum())
    assert all(abs(int(c[k]) - int(ora["reason_counts"][k])) <= n_band for k in range(11)) and c[10] > 0


def test_resampling_loop_and_last_level_fallback():
    """Nothing selectable at one level -> next level (reactive_planner.py:84-97); at the last level the
    reference falls back to the feasible trajectory of lowest risk (:262-269).  Without an attached risk function the
    planner ranks by: stays on the road, latest first collision, prediction cost, total cost."""
    g, ref, prm, preds = load_golden("scurve_brake_hv_draw")
    p = make_planner(g, prm, preds, 3.0)
    p._sampling_min = 1
    p.sampling_handler.update_static_params(0.9, 3.0, -3, 3)       # t_min = 0.9 keeps every level inside the horizon
    p.sampling_handler.set_v_sampling(*syn.velocity_interval(3.0, 11.5, 3.0, 50.8))
    p.set_static_obstacles([[30.0, 0.0, 0.0, 60.0, 30.0]])        # a wall over everything: nothing is selectable
    p.collision_check_enabled = True
    pair = p.plan()
    assert p._total_count == 8 * 9 * 10                             # level 2 was evaluated last (level 1: 5 x 5 x 6)
    st = p.last_plan_stats
    assert st.argmin == -1 and st.n_boundary == st.n_candidates > 0
    opt = p.optimal_trajectory
    assert pair is not None and opt.feasible and opt.valid
    b = p._bundle
    feas = np.flatnonzero((b.flags & 3) == 3)
    f = b.flags[feas]
    off = (f & (1 << 13)) != 0
    first_hit = np.where((f & (1 << 12)) != 0, (f >> 18) & 63, 64).astype(np.int64)
    pred = b.costs[feas, p.cost_names.index("prediction")]
    keys = sorted(zip(off.tolist(), (-first_hit).tolist(), pred.tolist(), b.total[feas].tolist(), feas.tolist()))
    assert off.all() and opt.uniqueId == keys[0][4]
    assert opt.boundary_harm > 0 and opt.boundary_harm == 1 / (1 + np.exp(4.591 - 0.185 * opt.cartesian.v[(int(b.flags[opt.uniqueId]) >> 24) & 63]))
    p.risk_function = lambda t: -t.cost                             # user-supplied risk: prefers the most expensive
    p.plan()
    assert p.optimal_trajectory.cost == p._bundle.total[(p._bundle.flags & 3) == 3].max()


def test_multi_agent_batched_launch_equals_individual_plans():
    """BASELINE.json configs[3] shape (several agents, own reference path / state / predictions each):
    one batched launch must give every agent exactly what its own plan() gives."""
    from frenetix_motion_planner_b200.reactive_planner_b200 import plan_batched
    cases = [("arc_hv_draw_pred", 9.5), ("straight_hv_draw", 8.0), ("scurve_lowvel_draw", 1.2), ("short_hv_draw", 8.0),
             ("scurve_brake_hv_draw", 3.0)]
    solo, batch = [], []
    for name, v0 in cases:
        g, ref, prm, preds = load_golden(name)
        for lst in (solo, batch):
            p = make_planner(g, prm, preds, v0)
            p.collision_check_enabled = True
            lst.append(p)
    for p in solo:
        p.plan()
    pairs = plan_batched(batch)
    assert len(pairs) == len(cases) and all(pr is not None for pr in pairs)
    for a, b in zip(solo, batch):
        assert a.optimal_trajectory.uniqueId == b.optimal_trajectory.uniqueId
        assert a.optimal_trajectory.cost == b.optimal_trajectory.cost
        assert a._infeasible_count_kinematics == b._infeasible_count_kinematics
        assert a.infeasible_count_collision == b.infeasible_count_collision
        assert np.array_equal(a._bundle.flags, b._bundle.flags) and np.array_equal(a._bundle.total, b._bundle.total)
        rows = np.arange(0, a._bundle.n_rows, 37)
        assert np.array_equal(a._bundle.states(rows), b._bundle.states(rows))
        assert [t.uniqueId for t in a.all_traj[:50]] == [t.uniqueId for t in b.all_traj[:50]]


def test_tjunction_scenario_from_cartesian_state():
    """BASELINE.json configs[0]: the shipped ZAM_Tjunction-1_42_T-1 scenario.  The planner gets the ego's
    CARTESIAN state (rear axle) and derives the Frenet state itself (planner.py:567-635), then plans."""
    import os
    from helpers import GOLDEN_DIR
    fx = np.load(os.path.join(GOLDEN_DIR, "tjunction.npz"))
    g, ref, prm, preds = load_golden("tjunction_draw")
    p = make_planner(g, prm, preds, float(fx["ego_velocity"]))
    x_0 = SimpleNamespace(position=fx["ego_position_rear"], orientation=float(fx["ego_orientation"]),
                          velocity=float(fx["ego_velocity"]), acceleration=float(fx["ego_acceleration"]),
                          yaw_rate=float(fx["ego_yaw_rate"]), steering_angle=0.0, time_step=0)
    p.x_cl = None
    p.update_externals(reference_path=fx["reference_path"], x_0=x_0, x_cl=None, desired_velocity=8.0)
    # golden x_cl: the reference's own _compute_initial_states on an independent projection (make_golden.py)
    assert np.allclose(p.x_cl[0], g["x_cl_lon"], rtol=1e-9) and np.allclose(p.x_cl[1], g["x_cl_lat"], rtol=1e-9, atol=1e-12)
    p.x_cl = (list(g["x_cl_lon"]), list(g["x_cl_lat"]))      # bit-identical sampling rows for the comparison below
    p.plan()
    ok = fo.plan(g["sampling"], ref, prm, preds, collision_check=False)["margins"] >= BAND
    opt = p.optimal_trajectory
    assert (opt.uniqueId == int(g["optimal_id"])) or not ok[opt.uniqueId] or not ok[int(g["optimal_id"])]
    # with the collision sweep switched on the five predicted cars are checked as well
    p.collision_check_enabled = True
    p.obstacle_order = [100 + i for i in range(5)]
    p.plan()
    assert p.optimal_trajectory is not None and p.last_plan_stats.n_candidates > 0
