"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = ["straight_hv_draw", "arc_hv_draw_pred", "arc_hv_nodraw_nodebug", "arc_hv_nodraw_debug",
                "scurve_lowvel_draw", "scurve_lowvel_nodraw", "scurve_slow_hv_draw", "scurve_slow_hv_nodraw",
                "scurve_brake_hv_draw", "scurve_brake_hv_nodraw_debug", "scurve_brake_hv_nodraw_nodebug",
                "short_hv_draw", "short_hv_nodraw",
                # ZAM_Tjunction-1_42_T-1 (BASELINE.json configs[0]): reference path, ego state and predicted cars
                # of the shipped scenario (tests/golden/make_tjunction_fixture.py)
                "tjunction_draw", "tjunction_nodraw"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, f"ref_{name}.npz"))
    poly = g["polyline"]
    ref = fo.RefPath(ref_pos=g["ref_pos"], ref_theta=g["ref_theta"], ref_curv=g["ref_curv"],
                     ref_curv_d=g["ref_curv_d"], ref_x=np.ascontiguousarray(poly[:, 0]),
                     ref_y=np.ascontiguousarray(poly[:, 1]))
    names = [str(x) for x in g["cost_names"]]
    prm = fo.Params(low_vel_mode=bool(g["low_vel_mode"]), x0_orientation=float(g["x0_orientation"]),
                    desired_velocity=float(g["desired_velocity"]), draw_traj_set=bool(g["draw"]),
                    kinematic_debug=bool(g["debug"]),
                    cost_weights={n: float(w) for n, w in zip(names, g["cost_weights"])},
                    **{k: syn.VEHICLE_2[k] for k in ("a_max", "v_switch", "delta_max", "wheelbase",
                                                      "wb_rear_axle", "length", "width")})
    preds = []
    for i in range(int(g["n_obs"])):
        sh = g[f"pred{i}_shape"]
        preds.append({"pos_list": g[f"pred{i}_pos"], "cov_list": g[f"pred{i}_cov"],
                      "orientation_list": g[f"pred{i}_ori"],
                      "shape": {"length": float(sh[0]), "width": float(sh[1])}})
    return g, ref, prm, preds


SINGULAR_V = 1e4        # m/s: beyond this a row sits on the Frenet singularity (see compare_with_oracle)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1.0)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / scale
    e = np.where(np.isnan(a) & np.isnan(b), 0.0, e)
    e = np.where(np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)), 0.0, e)
    return float(np.max(e)) if e.size else 0.0


# ---------------------------------------------------------------------------------------------
# device side (through the C ABI)
# ---------------------------------------------------------------------------------------------
def configure_handler(h, ref, prm, preds, static_obbs=None, sampling=None, T_values=None,
                      store_states=True, check_collisions=True):
    from frenetix_motion_planner_b200 import hotpath
    names = prm.active_costs()
    h.set_params(dt=prm.dt, N=prm.N, low_vel_mode=prm.low_vel_mode, draw_traj_set=prm.draw_traj_set,
                 kinematic_debug=prm.kinematic_debug, a_max=prm.a_max, v_switch=prm.v_switch,
                 delta_max=prm.delta_max, wheelbase=prm.wheelbase, wb_rear_axle=prm.wb_rear_axle,
                 length=prm.length, width=prm.width, x0_orientation=prm.x0_orientation,
                 desired_velocity=prm.desired_velocity, cost_names=names,
                 cost_weights=[prm.cost_weights[n] for n in names], store_states=store_states,
                 check_collisions=check_collisions, curvature_rate_from_v_delta=prm.curvature_rate_from_v_delta,
                 v_delta_max=prm.v_delta_max, velocity_offset_norm=prm.velocity_offset_norm,
                 prediction_cost_mode=prm.prediction_cost_mode)
    h.set_reference(ref.ref_pos, ref.ref_theta, ref.ref_curv, ref.ref_curv_d, ref.ref_x, ref.ref_y)
    if T_values is None:
        T_values = hotpath.distinct_durations(sampling)
    h.set_time_tables(*hotpath.time_tables(T_values, prm.dt, prm.N + 1))
    packed = hotpath.pack_predictions(list(preds)) if preds else None
    if packed is None:
        h.set_predictions(None, None, None, [], [], [])
    else:
        h.set_predictions(*packed)
    h.set_obstacle_positions(prm.obstacle_positions)
    h.set_static_obbs(static_obbs)


def device_plan(sampling, ref, prm, preds, static_obbs=None, device=0, handler=None, **kw):
    """Run one plan through libfrx_b200 and read everything back (test use: small N)."""
    from frenetix_motion_planner_b200 import _capi
    h = handler or _capi.Handler(device)
    configure_handler(h, ref, prm, preds, static_obbs, sampling=sampling, **kw)
    res = h.plan(np.ascontiguousarray(sampling, dtype=np.float64))
    flags, traj_len = h.get_flags()
    costs, total = h.get_costs()
    states = h.get_states_range() if kw.get("store_states", True) else None
    out = dict(res=res, flags=flags, traj_len=traj_len, costs=costs, total=total, states=states,
               argmin=int(res.argmin), min_cost=float(res.min_cost),
               reason_counts=np.array(list(res.reason_counts), dtype=np.int64),
               n_in_list=int(res.n_in_list), n_feasible=int(res.n_feasible),
               collision_counter=int(res.collision_counter), handler=h)
    return out


BAND = 1e-9   # decision margin below which the reference's own outcome is rounding noise


def band_alternatives(sampling, ref, prm, preds, rows, static_obbs=None):
    """Both legal outcomes of the candidates `rows` that sit on the `s_velocity > 0.001` tie (oracle docstring): the
    oracle with the stand-still threshold moved just below and just above 0.001."""
    import dataclasses
    rows = np.asarray(rows, dtype=np.int64)
    outs = []
    for thr in (0.001 - 2e-9, 0.001 + 2e-9):
        outs.append(fo.plan(np.asarray(sampling)[rows], ref, dataclasses.replace(prm, standstill_threshold=thr), preds,
                            static_obbs=static_obbs))
    return rows, outs


def assert_band_rows_take_a_legal_branch(dev, alts, bits, tol):
    """Every in-band candidate must equal the oracle on ONE side of the tie: same mask bits, states / costs within `tol`."""
    rows, outs = alts
    for j, r in enumerate(rows):
        ok_any = False
        why = []
        for o in outs:
            if (int(dev["flags"][r]) ^ int(o["flags"][j])) & bits:
                why.append(f"flags {int(dev['flags'][r]) & bits:#x} vs {int(o['flags'][j]) & bits:#x}")
                continue
            e = 0.0
            if dev["states"] is not None and (int(o["flags"][j]) & fo.FLAG_STORED):
                e = max(e, rel_err(dev["states"][:, r, :], o["states"][:, j, :]))
            if int(o["flags"][j]) & fo.FLAG_COSTED:
                e = max(e, rel_err(dev["costs"][r], o["costs"][j]), rel_err(dev["total"][r], o["total"][j]))
            if e < tol and dev["traj_len"][r] == o["traj_len"][j]:
                ok_any = True
                break
            why.append(f"rel err {e:.2e}")
        assert ok_any, f"in-band row {r} matches neither side of the stand-still tie: {why}"


def compare_with_oracle(dev, ora, prm, tol=1e-6, check_collide=True, band=BAND, alts=None):
    """Assert the parity contract: masks / selected index bit-exact, states & costs within `tol`
    relative -- for every candidate whose decision margins exceed `band`.  The rest
    sit on a structural tie of the reference (see oracle docstring): with `alts` (band_alternatives) each of them
    must equal the oracle on one of the two sides of the tie; without, they are only counted."""
    from oracle import frenet_oracle as fo
    ok = ora["margins"] >= band if "margins" in ora else np.ones(len(ora["flags"]), bool)
    n_band = int((~ok).sum())
    fl_d, fl_o = dev["flags"].astype(np.uint64), ora["flags"].astype(np.uint64)
    bits = fo.FLAG_VALID | fo.FLAG_FEASIBLE | fo.FLAG_STORED | fo.FLAG_IN_LIST | fo.FLAG_COSTED | fo.FLAG_CANDIDATE
    for r in range(1, 11):
        bits |= fo.reason_bit(r)
    if check_collide:
        bits |= fo.FLAG_COLLIDE | fo.FLAG_BOUNDARY
    if alts is not None:
        assert set(alts[0].tolist()) == set(np.flatnonzero(~ok).tolist()), "alternatives must cover exactly the in-band rows"
        assert_band_rows_take_a_legal_branch(dev, alts, bits, tol)
    diff = ((fl_d ^ fl_o) & np.uint64(bits))
    diff[~ok] = 0
    assert not diff.any(), f"{int((diff != 0).sum())} flag mismatches, first rows {np.nonzero(diff)[0][:5]}, " \
                           f"bits {[hex(int(x)) for x in diff[np.nonzero(diff)[0][:5]]]}"
    if check_collide:
        # index of the first colliding ego hull (feeds boundary_harm, planner.py:370-372): exact wherever a hit is flagged
        for flag, shift in ((fo.FLAG_COLLIDE, fo.COLLIDE_STEP_SHIFT), (fo.FLAG_BOUNDARY, fo.BOUNDARY_STEP_SHIFT)):
            hit = ((fl_o & np.uint64(flag)) != 0) & ok
            kd, ko = (fl_d[hit] >> np.uint64(shift)) & np.uint64(63), (fl_o[hit] >> np.uint64(shift)) & np.uint64(63)
            assert np.array_equal(kd, ko), f"first-hit hull index differs for flag {flag:#x}: rows {np.flatnonzero(hit)[kd != ko][:5]}"
    stored = ((fl_o & np.uint64(fo.FLAG_STORED)) != 0) & ok
    costed = ((fl_o & np.uint64(fo.FLAG_COSTED)) != 0) & ok
    errs = {}
    if dev["states"] is not None:
        # A candidate that runs into the singularity of the Frenet transform (1 - kappa_r d -> 0, theta_cl -> pi/2: speeds of
        # 1e10 m/s) amplifies the last-ulp difference between the device's atan and libm's by 1 / cos(theta_cl); it is
        # infeasible on both sides (the flags above are compared exactly) and only has to be singular on both sides.
        with np.errstate(invalid="ignore"):
            singular = np.nanmax(np.abs(ora["states"][fo.F_V]), axis=1) > SINGULAR_V
            assert np.array_equal(singular[stored], (np.nanmax(np.abs(dev["states"][fo.F_V]), axis=1) > SINGULAR_V)[stored])
        cmp_rows = stored & ~singular
        for f, name in enumerate(fo.FIELDS):
            errs[name] = rel_err(dev["states"][f][cmp_rows], ora["states"][f][cmp_rows])
    errs["costs"] = rel_err(dev["costs"][costed], ora["costs"][costed])
    errs["total"] = rel_err(dev["total"][costed], ora["total"][costed])
    worst = max(errs.values()) if errs else 0.0
    assert worst < tol, f"relative error {worst:.3e} over tolerance: {errs}"
    assert np.array_equal(dev["traj_len"][stored], ora["traj_len"][stored])
    # selected index: exact, unless one of the two winners is itself an in-band candidate
    wd, wo = dev["argmin"], ora["argmin"]
    if wd != wo:
        assert (wd >= 0 and not ok[wd]) or (wo >= 0 and not ok[wo]), (wd, wo)
    elif wo >= 0:
        assert abs(dev["min_cost"] - ora["min_cost"]) <= tol * max(1.0, abs(ora["min_cost"]))
    assert abs(dev["n_in_list"] - ora["n_in_list"]) <= n_band
    assert abs(dev["n_feasible"] - ora["n_feasible"]) <= n_band
    assert np.all(np.abs(dev["reason_counts"].astype(float) - ora["reason_counts"]) <= n_band)
    if check_collide and n_band == 0:
        assert dev["collision_counter"] == ora["collision_counter"]
    errs["in_band"] = n_band
    return errs
