"""GPU parity at scale: the CUDA path against the C oracle (tests/test_c_oracle.py pins that one to the
numpy oracle and the reference goldens), plus size-independent properties at BASELINE.json sizes."""
import numpy as np
import pytest

from oracle import frenet_oracle as fo
from oracle import c_oracle
from helpers import device_plan, compare_with_oracle, configure_handler, BAND
from frenetix_motion_planner_b200 import synthetic as syn
from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem

pytestmark = pytest.mark.gpu


def make_ref(poly):
    cs = CoordinateSystem(poly)
    return fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, poly[:, 0].copy(), poly[:, 1].copy())


def veh_params(**kw):
    base = {k: syn.VEHICLE_2[k] for k in ("a_max", "v_switch", "delta_max", "wheelbase", "wb_rear_axle", "length", "width")}
    base.update(kw)
    return fo.Params(**base)


def config2(n_t=20, n_v=50, n_d=50):
    poly = syn.straight_polyline(400)
    x_cl = ([10.0, 8.0, 0.0], [0.2, 0.0, 0.0])
    t1 = np.round(np.linspace(11, 30, n_t).round() * 0.1, 2)
    v_lo, v_hi = syn.velocity_interval(8.0, 11.5, 3.0, 50.8)
    v1, d1 = np.linspace(v_lo, v_hi, n_v), np.linspace(-3, 3, n_d)
    return poly, x_cl, np.unique(t1), v1, d1


def test_config2_full_size_matches_c_oracle():
    """BASELINE.json configs[1]: 50,000 candidates, straight path, no obstacles -- every candidate compared."""
    poly, x_cl, t1, v1, d1 = config2()
    S = syn.grid_sampling_matrix(t1, v1, d1, x_cl)
    assert S.shape[0] == 50_000
    ref, prm = make_ref(poly), veh_params(desired_velocity=8.0)
    ora = c_oracle.plan(S, ref, prm, [])
    dev = device_plan(S, ref, prm, [])
    errs = compare_with_oracle(dev, ora, prm, band=0.0)     # same closed-form coefficients: no band needed
    print({k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in errs.items()})


@pytest.mark.parametrize("draw,debug", [(True, True), (False, False), (False, True)])
def test_curved_with_obstacles_matches_c_oracle(draw, debug):
    """configs[2]-shaped (curved path, 20 predicted obstacles, full cost stack, collision sweep), 24,000 rows."""
    poly = syn.arc_polyline(R=80.0, M=400)
    x_cl = ([12.0, 9.5, 0.4], [-0.3, 0.2, -0.1])
    t1 = np.round(np.arange(11, 31) * 0.1, 2)
    v_lo, v_hi = syn.velocity_interval(9.5, 11.5, 3.0, 50.8)
    S = syn.grid_sampling_matrix(t1, np.linspace(v_lo, v_hi, 40), np.linspace(-3, 3, 30), x_cl)
    preds = syn.synthetic_predictions(poly, 20, 31, 0.1, seed=1234)
    ref = make_ref(poly)
    prm = veh_params(x0_orientation=0.2, desired_velocity=10.0, draw_traj_set=draw, kinematic_debug=debug)
    ora = c_oracle.plan(S, ref, prm, preds)
    dev = device_plan(S, ref, prm, preds)
    errs = compare_with_oracle(dev, ora, prm, band=1e-12)
    assert (ora["flags"] & fo.FLAG_COLLIDE).any() and ora["argmin"] >= 0
    print(draw, debug, {k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in errs.items()})


def test_long_horizon_two_chunks_lowvel_and_static_boxes():
    """N = 50 (51 samples -> two 32-lane chunks), low-velocity mode, inactive cost terms, static boxes."""
    poly = syn.scurve_polyline(M=300)
    x_cl = ([15.0, 1.2, 0.3], [0.1, 0.01, 0.0])
    t1 = np.round(np.array([1.1, 2.0, 3.3, 4.1, 5.0]), 2)
    S = syn.grid_sampling_matrix(t1, np.linspace(0.001, 6.0, 24), np.linspace(-2, 2, 21), x_cl)
    preds = syn.synthetic_predictions(poly, 7, 51, 0.1, seed=5)
    preds[2]["pos_list"] = preds[2]["pos_list"][:20]; preds[2]["cov_list"] = preds[2]["cov_list"][:20]   # ragged
    preds[2]["orientation_list"] = preds[2]["orientation_list"][:20]
    preds[4]["pos_list"] = preds[4]["pos_list"][:2]; preds[4]["cov_list"] = preds[4]["cov_list"][:2]     # too short to collide
    preds[4]["orientation_list"] = preds[4]["orientation_list"][:2]
    weights = dict(syn.DEFAULT_COST_WEIGHTS, acceleration=0.3, jerk=0.1, orientation_offset=0.7, path_length=0.2,
                   distance_to_obstacles=1.5)
    boxes = np.array([[45.0, 6.0, 0.4, 3.0, 1.0], [30.0, -4.5, 0.0, 5.0, 0.5]])
    ref = make_ref(poly)
    for low_v, x0 in ((True, x_cl), (False, ([15.0, 6.0, 0.3], [0.1, 0.2, 0.0]))):
        Sx = syn.grid_sampling_matrix(t1, np.linspace(0.001, 6.0 if low_v else 12.0, 24), np.linspace(-2, 2, 21), x0)
        prm = veh_params(N=50, low_vel_mode=low_v, x0_orientation=0.1, desired_velocity=3.0, cost_weights=weights,
                         obstacle_positions=np.array([[40.0, 3.0], [55.0, -2.0]]))
        ora = c_oracle.plan(Sx, ref, prm, preds, static_obbs=boxes)
        dev = device_plan(Sx, ref, prm, preds, static_obbs=boxes)
        errs = compare_with_oracle(dev, ora, prm, band=1e-12)
        print(low_v, {k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in errs.items()})


def test_row_order_does_not_matter_and_grid_mode_equals_matrix_mode():
    """Per-row results are a pure function of the row (the longitudinal memo must not leak between rows);
    rows generated on device equal the host matrix; arg-min is permutation invariant up to ties."""
    from frenetix_motion_planner_b200 import _capi
    poly, x_cl, t1, v1, d1 = config2(6, 12, 10)
    S = syn.grid_sampling_matrix(t1, v1, d1, x_cl)
    ref, prm = make_ref(poly), veh_params(desired_velocity=8.0)
    a = device_plan(S, ref, prm, [])
    perm = np.random.default_rng(3).permutation(S.shape[0])
    b = device_plan(S[perm], ref, prm, [])
    assert np.array_equal(a["flags"][perm], b["flags"])
    assert np.array_equal(a["total"][perm], b["total"])
    assert np.array_equal(a["states"][:, perm, :], b["states"])
    assert a["total"][a["argmin"]] == b["total"][b["argmin"]]
    h = _capi.Handler(0)
    configure_handler(h, ref, prm, [], sampling=S)
    res = h.plan_grid(t1, v1, d1, x_cl)
    assert res.argmin == a["argmin"] and res.min_cost == a["min_cost"]
    assert np.array_equal(h.get_flags()[0], a["flags"]) and np.array_equal(h.get_costs()[1], a["total"])
    assert np.array_equal(h.get_states_range(), a["states"])
    # shards of the same grid reduce to the single-device answer
    from frenetix_motion_planner_b200.dist import shard_rows, reduce_winners
    costs, rows = [], []
    for rk in range(3):
        first, count = shard_rows(S.shape[0], 3, rk)
        r = h.plan_grid(t1, v1, d1, x_cl, row_first=first, row_count=count)
        costs.append(r.min_cost); rows.append(r.argmin)
    c, r = reduce_winners(np.array(costs), np.array(rows))
    assert (c, r) == (a["min_cost"], a["argmin"])


def test_full_size_properties_config3_shape():
    """200,000 candidates, 20 obstacles (BASELINE.json configs[2] size): properties that need no oracle."""
    from frenetix_motion_planner_b200 import _capi
    poly = syn.arc_polyline(R=80.0, M=400)
    x_cl = ([12.0, 9.5, 0.4], [-0.3, 0.2, -0.1])
    t1 = np.round(np.arange(11, 31) * 0.1, 2)
    v_lo, v_hi = syn.velocity_interval(9.5, 11.5, 3.0, 50.8)
    v1, d1 = np.linspace(v_lo, v_hi, 100), np.linspace(-3, 3, 100)
    preds = syn.synthetic_predictions(poly, 20, 31, 0.1, seed=1234)
    ref = make_ref(poly)
    prm = veh_params(x0_orientation=0.2, desired_velocity=10.0)
    h = _capi.Handler(0)
    configure_handler(h, ref, prm, preds, T_values=t1)
    res = h.plan_grid(t1, v1, d1, x_cl)
    flags, tl = h.get_flags()
    costs, total = h.get_costs()
    assert res.n_rows == 200_000 and res.n_in_list == 200_000
    cand = (flags & fo.FLAG_CANDIDATE) != 0
    free = cand & ((flags & (fo.FLAG_COLLIDE | fo.FLAG_BOUNDARY)) == 0)
    assert res.argmin == int(np.flatnonzero(free)[np.argmin(total[free])])          # arg-min == host arg-min
    assert res.min_cost == total[res.argmin]
    assert res.n_feasible == int((((flags & fo.FLAG_VALID) != 0) & ((flags & fo.FLAG_FEASIBLE) != 0)).sum())
    assert res.n_collide == int((cand & ((flags & fo.FLAG_COLLIDE) != 0)).sum())
    w = np.array([prm.cost_weights[n] for n in prm.active_costs()])
    assert np.allclose(costs @ w, total, rtol=1e-12, atol=0)                           # weighted sum
    # end conditions of the polynomials on every candidate (reactive_planner.py:154-171)
    idx = np.arange(0, 200_000, 997)
    st = h.get_states(idx)
    S = syn.grid_sampling_matrix(t1, v1, d1, x_cl)[idx]
    last = tl[idx] - 1
    full = np.isclose(S[:, 1], 3.0)
    k = np.flatnonzero(full)
    assert np.allclose(st[fo.F_S_DOT][k, 30], S[k, 5], atol=1e-9)                       # s_dot(T) = ss1
    assert np.allclose(st[fo.F_S_DDOT][k, 30], 0.0, atol=1e-8)                          # s_ddot(T) = 0
    assert np.allclose(st[fo.F_D][k, 30], S[k, 10], atol=1e-9)                          # d(T) = d1
    # the sampled subset agrees with the C oracle row by row
    ora = c_oracle.plan(S, ref, prm, preds, T_values=t1)
    assert np.array_equal(ora["flags"] & 0x3ffff, flags[idx] & 0x3ffff)
    assert np.max(np.abs(ora["total"] - total[idx]) / np.maximum(1, np.abs(ora["total"]))) < 1e-9
