"""Property tests (hypothesis): the CUDA path against the C oracle on RANDOM sampling rows, reference paths,
Frenet states and obstacle sets; permutation invariance of the selection; shard-and-reduce == one plan."""
import dataclasses

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from helpers import compare_with_oracle, configure_handler, rel_err
from oracle import c_oracle, frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn
from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem
from frenetix_motion_planner_b200.dist import shard_rows, reduce_winners

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])

_HANDLER = {}


def handler():
    from frenetix_motion_planner_b200 import _capi
    if "h" not in _HANDLER:
        _HANDLER["h"] = _capi.Handler(0)
    return _HANDLER["h"]


def make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    rng = np.random.default_rng(seed)
    poly = {0: syn.straight_polyline(260), 1: syn.arc_polyline(R=float(rng.uniform(40, 300)), M=260),
            2: syn.scurve_polyline(M=260, amp=float(rng.uniform(1, 6)))}[path_kind]
    cs = CoordinateSystem(poly)
    ref = fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, np.ascontiguousarray(poly[:, 0]), np.ascontiguousarray(poly[:, 1]))
    v0 = float(rng.uniform(0.3, 1.9)) if low_vel else float(rng.uniform(2.5, 14.0))
    x_cl = ([float(rng.uniform(5, 40)), v0, float(rng.uniform(-2, 2))],
            [float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.2, 0.2))])
    S = np.zeros((n_rows, 13))
    S[:, 1] = np.round(rng.integers(5, 31, n_rows) * 0.1, 2)                  # durations on the dt raster
    S[:, 2], S[:, 3], S[:, 4] = x_cl[0]
    v_lo, v_hi = syn.velocity_interval(v0, syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"])
    S[:, 5] = rng.uniform(v_lo, v_hi, n_rows)
    S[:, 7], S[:, 8], S[:, 9] = x_cl[1]
    S[:, 10] = rng.uniform(-3.5, 3.5, n_rows)
    order = np.lexsort((S[:, 10], S[:, 5], S[:, 1]))                            # cartesian-like order (any order is legal)
    if seed % 2:
        S = S[order]
    prm = fo.Params(low_vel_mode=low_vel, x0_orientation=float(cs.ref_theta[10] + rng.uniform(-0.2, 0.2)),
                    desired_velocity=float(rng.uniform(1, 14)), draw_traj_set=draw, kinematic_debug=debug,
                    **{k: syn.VEHICLE_2[k] for k in ("a_max", "v_switch", "delta_max", "wheelbase", "wb_rear_axle", "length", "width")})
    preds = syn.synthetic_predictions(poly, n_obs, int(rng.integers(12, 40)), 0.1, seed=seed) if n_obs else []
    static = None
    if walls:
        static = np.array([[float(rng.uniform(20, 80)), float(rng.uniform(-6, 6)), float(rng.uniform(-1, 1)),
                            float(rng.uniform(2, 15)), float(rng.uniform(0.1, 1.0))] for _ in range(walls)])
    return S, ref, prm, preds, static


def device(S, ref, prm, preds, static, row_base=0):
    h = handler()
    configure_handler(h, ref, prm, preds, static, T_values=np.round(np.arange(5, 31) * 0.1, 2))
    res = h.plan(np.ascontiguousarray(S), row_index_base=row_base)
    flags, traj_len = h.get_flags()
    costs, total = h.get_costs()
    return dict(res=res, flags=flags, traj_len=traj_len, costs=costs, total=total, states=h.get_states_range(),
                argmin=int(res.argmin), min_cost=float(res.min_cost), reason_counts=np.array(list(res.reason_counts), dtype=np.int64),
                n_in_list=int(res.n_in_list), n_feasible=int(res.n_feasible), collision_counter=int(res.collision_counter))


case_args = dict(seed=st.integers(0, 2 ** 31 - 1), n_rows=st.integers(1, 700), path_kind=st.integers(0, 2), low_vel=st.booleans(),
                 draw=st.booleans(), debug=st.booleans(), n_obs=st.integers(0, 9), walls=st.integers(0, 3))


@settings(**SETTINGS)
@given(**case_args)
def test_device_equals_c_oracle_on_random_inputs(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    S, ref, prm, preds, static = make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls)
    ora = c_oracle.plan(S, ref, prm, preds, static_obbs=static, T_values=np.round(np.arange(5, 31) * 0.1, 2))
    dev = device(S, ref, prm, preds, static)
    # same closed forms on both sides: a decision band of 1e-12 (ulp-level libm differences) is all that is left
    out = dict(ora)
    out["argmin"] = ora["argmin"]
    compare_with_oracle(dev, out, prm, band=1e-10)


@settings(**SETTINGS)
@given(**case_args)
def test_selection_is_invariant_under_row_permutation(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    S, ref, prm, preds, static = make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls)
    a = device(S, ref, prm, preds, static)
    perm = np.random.default_rng(seed ^ 0x5bd1e995).permutation(S.shape[0])
    b = device(S[perm], ref, prm, preds, static)
    assert np.array_equal(a["flags"][perm], b["flags"]) and np.array_equal(a["total"][perm], b["total"])
    assert np.array_equal(a["states"][:, perm, :], b["states"])
    assert a["min_cost"] == b["min_cost"] or (a["argmin"] < 0 and b["argmin"] < 0)
    if a["argmin"] >= 0:
        # ties break towards the lowest row of the matrix AS GIVEN: the winner of the permuted matrix is the first row, in
        # its order, among those that share the minimum cost
        free = ((a["flags"] & fo.FLAG_CANDIDATE) != 0) & ((a["flags"] & (fo.FLAG_COLLIDE | fo.FLAG_BOUNDARY)) == 0)
        best = np.flatnonzero(free & (a["total"] == a["min_cost"]))
        assert a["argmin"] == best.min()
        assert b["argmin"] == np.flatnonzero(np.isin(perm, best)).min()
    assert np.array_equal(a["reason_counts"], b["reason_counts"]) and a["n_feasible"] == b["n_feasible"]


@settings(**SETTINGS)
@given(shards=st.integers(2, 5), **case_args)
def test_shard_and_reduce_equals_one_plan(shards, seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    S, ref, prm, preds, static = make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls)
    whole = device(S, ref, prm, preds, static)
    costs, rows, n_feas = [], [], 0
    for r in range(shards):
        first, count = shard_rows(S.shape[0], shards, r)
        if count == 0:
            costs.append(np.inf); rows.append(-1)
            continue
        part = device(S[first:first + count], ref, prm, preds, static, row_base=first)
        costs.append(part["min_cost"]); rows.append(part["argmin"]); n_feas += part["n_feasible"]
        assert np.array_equal(part["flags"], whole["flags"][first:first + count])
    c, r = reduce_winners(np.array(costs), np.array(rows))
    assert (r, c if r >= 0 else None) == (whole["argmin"], whole["min_cost"] if whole["argmin"] >= 0 else None)
    assert n_feas == whole["n_feasible"]
