"""TEST INFRASTRUCTURE: a stand-in for ``_capi.Handler`` that answers from the C oracle instead of the GPU.

The drop-in tests drive ``ReactivePlannerB200`` with the REFERENCE'S OWN callers (``FrenetPlannerInterface``,
``AgentBatch``; imported unmodified under tests/golden/ref_stubs.py).  Those run in the build container, which has the
reference tree but no GPU; the GPU box has a GPU but no reference tree.  So the host logic of the planner is exercised
here against this oracle-backed handler, the result is recorded (tests/golden/interface_trace.npz), and the GPU suite
replays the same call sequence on the device and compares.  The product never imports this file."""
import numpy as np

from oracle import c_oracle, frenet_oracle as fo
from frenetix_motion_planner_b200 import _capi


class OracleHandler:
    def __init__(self, device=0):
        self.device, self.n_rows, self.n_costs, self.Nt, self.generation = device, 0, 0, 0, 0
        self._prm = self._ref = None
        self._preds, self._static, self._out, self._row_base = [], None, None, 0

    # ---- set-up (same signatures as _capi.Handler)
    def set_params(self, *, dt, N, low_vel_mode, draw_traj_set, kinematic_debug, a_max, v_switch, delta_max, wheelbase,
                   wb_rear_axle, length, width, x0_orientation, desired_velocity, cost_names, cost_weights,
                   store_states=True, check_collisions=True, curvature_rate_from_v_delta=False, v_delta_max=0.4,
                   velocity_offset_norm=1, prediction_cost_mode=0):
        self._prm = fo.Params(dt=dt, N=N, a_max=a_max, v_switch=v_switch, delta_max=delta_max, wheelbase=wheelbase,
                              wb_rear_axle=wb_rear_axle, length=length, width=width, low_vel_mode=bool(low_vel_mode),
                              x0_orientation=x0_orientation, desired_velocity=desired_velocity,
                              draw_traj_set=bool(draw_traj_set), kinematic_debug=bool(kinematic_debug),
                              cost_weights=dict(zip(cost_names, cost_weights)),
                              curvature_rate_from_v_delta=bool(curvature_rate_from_v_delta), v_delta_max=v_delta_max,
                              velocity_offset_norm=int(velocity_offset_norm), prediction_cost_mode=int(prediction_cost_mode))
        self._check = bool(check_collisions)
        self.n_costs, self.Nt = len(cost_names), N + 1

    def set_reference(self, ref_pos, ref_theta, ref_curv, ref_curv_d, ref_x, ref_y):
        self._ref = fo.RefPath(*[np.ascontiguousarray(a, dtype=np.float64) for a in
                                 (ref_pos, ref_theta, ref_curv, ref_curv_d, ref_x, ref_y)])

    def set_time_tables(self, T_values, traj_len, tpow):
        pass

    def set_predictions(self, pos, cov, theta, half_len, half_wid, len_valid):
        self._preds = []
        if pos is None or len(half_len) == 0:
            return
        for o in range(len(half_len)):
            n = int(len_valid[o])
            self._preds.append({"pos_list": np.asarray(pos[o][:n]), "cov_list": np.asarray(cov[o][:n]),
                                "orientation_list": np.asarray(theta[o][:n]),
                                "shape": {"length": 2 * float(half_len[o]), "width": 2 * float(half_wid[o])}})

    def set_obstacle_positions(self, pos_xy):
        self._prm.obstacle_positions = None if pos_xy is None else np.asarray(pos_xy, dtype=np.float64)

    def set_static_obbs(self, obbs):
        self._static = None if obbs is None or len(obbs) == 0 else np.asarray(obbs, dtype=np.float64)

    def set_stream(self, h):
        pass

    # ---- plans
    def _run(self, S, row_base):
        self.generation += 1
        plan = fo.plan if self._prm.prediction_cost_mode == 1 else c_oracle.plan      # the C port has no collision-probability cost
        out = plan(S, self._ref, self._prm, self._preds if self._check else [], static_obbs=self._static if self._check else None,
                   check_all_collisions=True, collision_check=self._check)
        self._out, self._row_base, self.n_rows = out, row_base, S.shape[0]
        res = _capi.FrxResult()
        res.argmin = out["argmin"] + row_base if out["argmin"] >= 0 else -1
        res.min_cost = out["min_cost"]
        res.n_rows, res.n_in_list, res.n_feasible = S.shape[0], out["n_in_list"], out["n_feasible"]
        fl = out["flags"]
        cand = (fl & fo.FLAG_CANDIDATE) != 0
        res.n_candidates = int(cand.sum())
        res.n_collide = int((cand & ((fl & fo.FLAG_COLLIDE) != 0)).sum())
        res.n_boundary = int((cand & ((fl & fo.FLAG_BOUNDARY) != 0)).sum())
        res.collision_counter = out["collision_counter"]
        for k in range(11):
            res.reason_counts[k] = int(out["reason_counts"][k])
        return res

    def plan(self, sampling, row_index_base=0):
        return self._run(np.ascontiguousarray(sampling, dtype=np.float64), int(row_index_base))

    def plan_grid(self, t1, ss1, d1, x_cl, row_first=0, row_count=None):
        t1, ss1, d1 = (np.asarray(a, dtype=np.float64) for a in (t1, ss1, d1))
        (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
        n = t1.size * ss1.size * d1.size
        S = np.zeros((n, 13))
        S[:, 1] = np.repeat(t1, ss1.size * d1.size)
        S[:, 2], S[:, 3], S[:, 4] = s0, ss0, sss0
        S[:, 5] = np.tile(np.repeat(ss1, d1.size), t1.size)
        S[:, 7], S[:, 8], S[:, 9] = d0, dd0, ddd0
        S[:, 10] = np.tile(d1, t1.size * ss1.size)
        row_count = n - row_first if row_count is None else row_count
        return self._run(S[row_first:row_first + row_count], int(row_first))

    # ---- read-back
    def last_launches(self):
        return 0

    def get_flags(self, first=0, count=None):
        count = self.n_rows - first if count is None else count
        return self._out["flags"][first:first + count].copy(), self._out["traj_len"][first:first + count].copy()

    def get_costs(self, first=0, count=None):
        count = self.n_rows - first if count is None else count
        return self._out["costs"][first:first + count].copy(), self._out["total"][first:first + count].copy()

    def get_states(self, idx, fields=None):
        idx = np.asarray(idx, dtype=np.int64)
        st = self._out["states"][:, idx, :]
        if fields is not None:
            st = st[[(_capi.FIELD_ID[f] if isinstance(f, str) else int(f)) for f in fields]]
        return st.copy()

    def winner_states(self, fields=None):
        w = self._out["argmin"]
        return self.get_states(np.array([w]), fields)[:, 0, :]

    def winner_record(self):
        w = self._out["argmin"]
        return int(self._out["flags"][w]), int(self._out["traj_len"][w]), float(self._out["total"][w]), self._out["costs"][w].copy()

    def close(self):
        pass


def install_batched():
    """Route ``_capi.plan_batched`` through the oracle handlers (one oracle plan per agent)."""
    def plan_batched(handlers, samplings):
        return [h.plan(S) for h, S in zip(handlers, samplings)]
    _capi.plan_batched = plan_batched
