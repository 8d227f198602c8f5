"""Pin the C oracle: against the numpy oracle (itself pinned to the reference's own outputs) and
directly against the reference goldens."""
import numpy as np
import pytest

from oracle import frenet_oracle as fo
from oracle import c_oracle
from helpers import GOLDEN_CASES, BAND, load_golden, compare_with_oracle, rel_err
from frenetix_motion_planner_b200 import synthetic as syn


def _as_dev(c):
    return dict(flags=c["flags"], traj_len=c["traj_len"], costs=c["costs"], total=c["total"], states=c["states"],
                argmin=c["argmin"], min_cost=c["min_cost"], reason_counts=c["reason_counts"].astype(np.int64),
                n_in_list=c["n_in_list"], n_feasible=c["n_feasible"], collision_counter=c["collision_counter"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_oracle_matches_numpy_oracle(name):
    g, ref, prm, preds = load_golden(name)
    ora = fo.plan(g["sampling"], ref, prm, preds)
    c = c_oracle.plan(g["sampling"], ref, prm, preds)
    errs = compare_with_oracle(_as_dev(c), ora, prm, tol=1e-9)
    # the two oracles must also agree on which candidates are structurally marginal
    assert np.array_equal(c["margins"] < BAND, ora["margins"] < BAND)


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_lowvel_draw", "short_hv_nodraw"])
def test_c_oracle_lazy_collision_walk_equals_check_all(name):
    g, ref, prm, preds = load_golden(name)
    a = c_oracle.plan(g["sampling"], ref, prm, preds, check_all_collisions=True)
    b = c_oracle.plan(g["sampling"], ref, prm, preds, check_all_collisions=False, want_states=False)
    assert a["argmin"] == b["argmin"] and a["collision_counter"] == b["collision_counter"]


def test_np_sum_emulation_is_numpys_order():
    """np_sum in frx_oracle.c claims numpy's pairwise order: bit-identical to np.sum for every
    length the cost terms use; simps likewise against scipy's simpson."""
    import ctypes as C
    from scipy.integrate import simpson
    L = c_oracle.lib()
    L.orc_np_sum.restype = C.c_double
    L.orc_simps.restype = C.c_double
    rng = np.random.default_rng(5)
    for n in range(1, 65):
        for _ in range(20):
            a = rng.normal(0, 1, n) * 10.0 ** rng.integers(-8, 8, n)
            got = L.orc_np_sum(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(n))
            assert got == float(np.sum(a)), n
    for n in (3, 30, 31, 50, 51):
        y = rng.normal(0, 1, n) ** 2
        got = L.orc_simps(y.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(n), C.c_double(0.1))
        assert abs(got - float(simpson(y, dx=0.1))) <= 1e-14 * max(1.0, abs(got))
        assert abs(got - fo.simps(y, 0.1)) <= 1e-14 * max(1.0, abs(got))


def test_c_oracle_inactive_costs_and_static_boxes():
    poly = syn.scurve_polyline(M=220)
    from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem
    cs = CoordinateSystem(poly)
    ref = fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, poly[:, 0].copy(), poly[:, 1].copy())
    weights = {"acceleration": 0.3, "jerk": 0.1, "orientation_offset": 0.7, "path_length": 0.2,
               "distance_to_obstacles": 1.5, "velocity_offset": 1.0, "prediction": 0.4}
    prm = fo.Params(x0_orientation=0.3, desired_velocity=9.0, cost_weights=weights,
                    obstacle_positions=np.array([[40.0, 3.0], [55.0, -2.0]]))
    x_cl = ([12.0, 9.5, 0.4], [-0.3, 0.2, -0.1])
    S = syn.grid_sampling_matrix([1.1, 2.0, 3.0], np.linspace(2.0, 14.0, 6), np.linspace(-2.5, 2.5, 7), x_cl)
    preds = syn.synthetic_predictions(poly, 3, 31, 0.1, seed=3)
    boxes = np.array([[45.0, 6.0, 0.4, 3.0, 1.0], [30.0, -4.5, 0.0, 5.0, 0.5]])
    ora = fo.plan(S, ref, prm, preds, static_obbs=boxes)
    c = c_oracle.plan(S, ref, prm, preds, static_obbs=boxes)
    compare_with_oracle(_as_dev(c), ora, prm, tol=1e-9)
    assert (ora["flags"] & fo.FLAG_BOUNDARY).any() and (ora["flags"] & fo.FLAG_COLLIDE).any()
