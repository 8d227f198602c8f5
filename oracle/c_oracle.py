"""ctypes wrapper of the C oracle (oracle/c/frx_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Same call shape and result dict as :func:`oracle.frenet_oracle.plan`; used for parity checks at
sizes the numpy oracle is too slow for and as the timed CPU baseline of bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import frenet_oracle as fo
from .build import build, build_native, OUT

_lib = None
_native = None


class OrcParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("N", C.c_int32), ("low", C.c_int32), ("draw", C.c_int32), ("debug", C.c_int32),
                ("a_max", C.c_double), ("v_switch", C.c_double), ("delta_max", C.c_double), ("wheelbase", C.c_double),
                ("wb_rear", C.c_double), ("length", C.c_double), ("width", C.c_double), ("x0_orientation", C.c_double),
                ("v_des", C.c_double), ("n_costs", C.c_int32), ("cost_ids", C.c_int32 * 10), ("w", C.c_double * 10),
                ("check_all_collisions", C.c_int32), ("collision_check", C.c_int32),
                ("kd_from_v_delta", C.c_int32), ("vo_norm", C.c_int32), ("v_delta_max", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [("argmin", C.c_int64), ("min_cost", C.c_double), ("n_in_list", C.c_int64), ("n_feasible", C.c_int64),
                ("n_candidates", C.c_int64), ("collision_counter", C.c_int64), ("reason_counts", C.c_int64 * 11)]


def _declare(L):
    L.orc_plan.restype = C.c_int
    L.orc_max_threads.restype = C.c_int
    L.orc_last_threads.restype = C.c_int
    return L


def lib():
    global _lib
    if _lib is None:
        build()                      # no-op when the library is newer than its source
        _lib = _declare(C.CDLL(OUT))
    return _lib


def native_lib():
    """The host-tuned build (-O3 -march=native), compiled on this machine on first use; bench.py's CPU arm."""
    global _native
    if _native is None:
        _native = _declare(C.CDLL(build_native()))
    return _native


def threads_used(library=None) -> int:
    return int((library or lib()).orc_last_threads())


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def time_tables(T_values, dt, Nt):
    T_values = np.asarray(T_values, dtype=np.float64)
    lens = np.zeros(T_values.size, dtype=np.int32)
    tpow = np.zeros((T_values.size, 5, Nt))
    for k, T in enumerate(T_values):
        t, t2, t3, t4, t5 = fo.time_grid(T, dt)
        n = len(t)
        lens[k] = n
        for q, arr in enumerate((t, t2, t3, t4, t5)):
            tpow[k, q, :n] = arr
    return T_values, lens, tpow


def pack_predictions(predictions):
    O = len(predictions)
    if O == 0:
        return 0, 0, None, None, None, None, None, None
    T = max(len(p["pos_list"]) for p in predictions)
    pos = np.zeros((O, T, 2)); cov = np.tile(np.eye(2), (O, T, 1, 1)); th = np.zeros((O, T))
    hl = np.zeros(O); hw = np.zeros(O); ln = np.zeros(O, dtype=np.int32)
    for o, p in enumerate(predictions):
        n = len(p["pos_list"]); ln[o] = n
        pos[o, :n] = p["pos_list"]; cov[o, :n] = p["cov_list"]; th[o, :n] = np.asarray(p["orientation_list"])[:n]
        hl[o] = p["shape"]["length"] / 2; hw[o] = p["shape"]["width"] / 2
    return O, T, pos, cov, th, hl, hw, ln


def plan(sampling, ref: fo.RefPath, prm: fo.Params, predictions=(), static_obbs=None,
         check_all_collisions=True, collision_check=True, want_states=True, want_margins=True, nthreads=0,
         T_values=None, buffers=None, library=None):
    """`buffers`: dict reused across calls (timing runs: keeps page faults of fresh arrays out of the loop).
    `nthreads` > 0 overrides OMP_NUM_THREADS (torchrun exports 1).  `library`: native_lib() for timing runs."""
    if getattr(prm, "prediction_cost_mode", 0) != 0:
        raise NotImplementedError("the C oracle only has the inverse-Mahalanobis prediction cost (use frenet_oracle.plan)")
    L = library or lib()
    S = np.ascontiguousarray(sampling, dtype=np.float64)
    n = S.shape[0]
    Nt = prm.N + 1
    names = prm.active_costs()
    K = len(names)
    P = OrcParams()
    P.dt, P.N, P.low, P.draw, P.debug = prm.dt, prm.N, int(prm.low_vel_mode), int(prm.draw_traj_set), int(prm.kinematic_debug)
    P.a_max, P.v_switch, P.delta_max, P.wheelbase = prm.a_max, prm.v_switch, prm.delta_max, prm.wheelbase
    P.wb_rear, P.length, P.width = prm.wb_rear_axle, prm.length, prm.width
    P.x0_orientation, P.v_des = prm.x0_orientation, prm.desired_velocity
    P.n_costs = K
    for k, nm in enumerate(names):
        P.cost_ids[k] = fo.COST_ID[nm]
        P.w[k] = prm.cost_weights[nm]
    P.check_all_collisions, P.collision_check = int(check_all_collisions), int(collision_check)
    P.kd_from_v_delta, P.vo_norm, P.v_delta_max = int(prm.curvature_rate_from_v_delta), int(prm.velocity_offset_norm), prm.v_delta_max
    if T_values is None:
        T_values = np.unique(S[:, 1])
    Tv, Tl, tp = time_tables(T_values, prm.dt, Nt)
    O, T, pos, cov, th, hl, hw, ln = pack_predictions(list(predictions))
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
            (ref.ref_pos, ref.ref_theta, ref.ref_curv, ref.ref_curv_d, ref.ref_x, ref.ref_y)]
    obs_pos = None if prm.obstacle_positions is None else np.ascontiguousarray(prm.obstacle_positions, dtype=np.float64)
    sobb = None if static_obbs is None or len(static_obbs) == 0 else np.ascontiguousarray(static_obbs, dtype=np.float64)
    if buffers is not None and buffers.get("n") == (n, Nt, K):
        states, costs, total, flags, tl, margins = (buffers[k] for k in ("states", "costs", "total", "flags", "tl", "margins"))
    else:
        states = np.zeros((14, n, Nt)) if want_states else None
        costs = np.zeros((n, max(K, 1))); total = np.zeros(n)
        flags = np.zeros(n, dtype=np.uint32); tl = np.zeros(n, dtype=np.int32)
        margins = np.zeros(n) if want_margins else None
        if buffers is not None:
            buffers.update(n=(n, Nt, K), states=states, costs=costs, total=total, flags=flags, tl=tl, margins=margins)
    res = OrcResult()
    rc = L.orc_plan(C.byref(P), C.c_int64(n), _p(S), C.c_int(arrs[0].size), *[_p(a) for a in arrs],
                    C.c_int(Tv.size), _p(Tv), _p(Tl, C.c_int32), _p(tp),
                    C.c_int(O), C.c_int(T), _p(pos), _p(cov), _p(th), _p(hl), _p(hw), _p(ln, C.c_int32),
                    C.c_int(0 if obs_pos is None else obs_pos.shape[0]), _p(obs_pos),
                    C.c_int(0 if sobb is None else sobb.shape[0]), _p(sobb),
                    _p(states), _p(costs), _p(total), _p(flags, C.c_uint32), _p(tl, C.c_int32), _p(margins),
                    C.byref(res), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"orc_plan failed: {rc}")
    n_list = int(res.n_in_list)
    return dict(states=states, flags=flags, traj_len=tl, costs=costs[:, :K], total=total, margins=margins,
                reason_counts=np.array(list(res.reason_counts), dtype=float), n_in_list=n_list,
                n_feasible=int(res.n_feasible), percentage=(100.0 * res.n_feasible / n_list if n_list else 0.0),
                argmin=int(res.argmin), min_cost=float(res.min_cost), collision_counter=int(res.collision_counter),
                cost_names=names)


def max_threads():
    return int(lib().orc_max_threads())
