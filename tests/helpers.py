"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = ["straight_hv_draw", "arc_hv_draw_pred", "arc_hv_nodraw_nodebug", "arc_hv_nodraw_debug",
                "scurve_lowvel_draw", "scurve_lowvel_nodraw", "scurve_slow_hv_draw", "scurve_slow_hv_nodraw",
                "scurve_brake_hv_draw", "scurve_brake_hv_nodraw_debug", "scurve_brake_hv_nodraw_nodebug",
                "short_hv_draw", "short_hv_nodraw"]


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


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1.0)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / scale
    e = np.where(np.isnan(a) & np.isnan(b), 0.0, e)
    e = np.where(np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)), 0.0, e)
    return float(np.max(e)) if e.size else 0.0
