"""Drive ReactivePlannerB200 with the REFERENCE'S OWN caller and record what happens -> tests/golden/interface_trace.npz.

Run here (the reference tree is mounted, no GPU): the unmodified ``FrenetPlannerInterface.update_planner`` and
``FrenetPlannerInterface.step_interface`` (cr_scenario_handler/planner_interfaces/frenet_interface.py:178-287, imported
under ref_stubs) run NINE simulation steps = three replanning cycles (planning.replanning_frequency = 3) of the
ZAM_Tjunction-1_42_T-1 fixture with the B200 planner in the place ``frenet_interface.py:71-73`` gives it.  The interface
object is created with ``__new__`` and given the attributes its ``__init__`` sets (that constructor needs commonroad-io,
the route planner and the velocity planner, none of which are installed); the planner calls ``__init__`` makes are made
here in the same order (:84-136).  The device is replaced by the oracle-backed handler of tests/oracle_handler.py; the
GPU suite replays the recorded calls on the real device (tests/test_reference_dropin.py).
"""
import json
import logging
import os
import sys
import types
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, os.path.dirname(HERE), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402

N_STEPS, REPLAN = 9, 3
DESIRED_VELOCITY = 8.0


def configs(draw=True):
    cfg_plan = types.SimpleNamespace(
        planning=types.SimpleNamespace(planning_horizon=3.0, dt=0.1, low_vel_mode_threshold=2.0, sampling_min=2, sampling_max=3,
                                       t_min=1.1, d_min=-3, d_max=3, d_ego_pos=False, replanning_frequency=REPLAN),
        debug=types.SimpleNamespace(multiproc=True, num_workers=6, draw_traj_set=draw, kinematic_debug=True,
                                    save_all_traj=False, log_risk=False, use_cpp=False, msg_log_mode="WARNING"),
        cost=types.SimpleNamespace(cost_weights=dict(syn.DEFAULT_COST_WEIGHTS)))
    cfg_sim = types.SimpleNamespace(vehicle=types.SimpleNamespace(**syn.VEHICLE_2),
                                    behavior=types.SimpleNamespace(use_behavior_planner=False),
                                    simulation=types.SimpleNamespace(ego_agent_id=60000))
    return cfg_plan, cfg_sim


def fixture():
    fx = np.load(os.path.join(HERE, "tjunction.npz"))
    raw = json.load(open(os.path.join(HERE, "tjunction_lanelets.json")))
    lanelets = {int(k): dict(left=np.array(v["left"]), right=np.array(v["right"]), adj_left=v["adj_left"],
                             adj_right=v["adj_right"]) for k, v in raw.items()}
    return fx, lanelets


def predictions_at(fx, t):
    """Ground-truth predictions at simulation step t (prediction_helpers.py:207-257 + the :167-170 safety margins)."""
    preds = {}
    for o, oid in enumerate(fx["obstacle_ids"]):
        st = fx["obstacle_states"][o, t + 1:t + 32]
        preds[int(oid)] = {"pos_list": st[:, :2].copy(), "cov_list": np.tile(np.array([[0.1, 0.0], [0.0, 0.1]]), (len(st), 1, 1)),
                           "orientation_list": st[:, 2].copy(), "v_list": st[:, 3].copy(),
                           "shape": {"length": float(fx["obstacle_shapes"][o, 0]) + 0.5,
                                     "width": float(fx["obstacle_shapes"][o, 1]) + 0.2}}
    return preds


def initial_state(fx):
    from frenetix_motion_planner_b200.reactive_planner_b200 import PlannerState
    v, yr = float(fx["ego_velocity"]), float(fx["ego_yaw_rate"])
    return PlannerState(time_step=0, position=np.array(fx["ego_position_rear"]), orientation=float(fx["ego_orientation"]),
                        velocity=v, acceleration=float(fx["ego_acceleration"]), yaw_rate=yr,
                        steering_angle=float(np.arctan2(syn.VEHICLE_2["wheelbase"] * yr, v)))      # state.py:70-72


def make_interface(iface_cls, planner, fx, lanelets, cfg_plan, cfg_sim):
    """frenet_interface.py:35-141 without the third-party construction work."""
    it = iface_cls.__new__(iface_cls)
    it.config_plan, it.config_sim, it.scenario, it.id = cfg_plan, cfg_sim, None, 60000
    it.DT = cfg_plan.planning.dt
    it.replanning_counter, it.replanning_traj, it.behavior_module_state = 0, None, None
    it.planning_problem, it.log_path, it.mod_path = None, None, None
    it.msg_logger = logging.getLogger("Message_logger_60000")
    it.planner = planner
    x_0 = initial_state(fx)
    planner.set_ego_vehicle_state(current_ego_vehicle=types.SimpleNamespace(obstacle_id=60000, initial_state=deepcopy(x_0)))
    it.x_0 = x_0
    planner.record_state_and_input(it.x_0)
    it.x_cl = it.desired_velocity = it.occlusion_module = it.behavior_module = it.route_planner = None
    it.reference_path = fx["reference_path"]
    it.goal_area = None
    planner.set_road_boundary(lanelets)                       # Planner.set_scenario builds it once (planner.py:550-565)
    planner.update_externals(x_0=it.x_0, reference_path=it.reference_path, goal_area=it.goal_area,
                             occlusion_module=it.occlusion_module)
    it.x_cl = planner.x_cl
    it.velocity_planner = types.SimpleNamespace(calculate_desired_velocity=lambda x_0, s: DESIRED_VELOCITY)
    return it


def run(iface_cls, handler_factory, record=True):
    fx, lanelets = fixture()
    cfg_plan, cfg_sim = configs()
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    planner = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, logging.getLogger("Message_logger_60000"),
                                  handler=handler_factory())
    planner.obstacle_order = [int(i) for i in fx["obstacle_ids"]]
    it = make_interface(iface_cls, planner, fx, lanelets, cfg_plan, cfg_sim)
    scenario = types.SimpleNamespace(lanelet_network=None)
    out = {"x_cl0": np.array(it.x_cl[0] + it.x_cl[1])}
    for t in range(N_STEPS):
        it.update_planner(scenario, predictions_at(fx, t))
        x_in = it.x_0
        replanned = it.replanning_counter == 0 or int(it.replanning_counter / REPLAN) == 1
        traj, counter = it.step_interface(t)
        assert traj is not None, f"no trajectory at step {t}"
        out[f"s{t}_counter"] = counter
        out[f"s{t}_x0"] = np.array([x_in.position[0], x_in.position[1], x_in.orientation, x_in.velocity, x_in.acceleration,
                                    x_in.yaw_rate, x_in.steering_angle, x_in.time_step])
        out[f"s{t}_xcl_after"] = np.array(list(it.x_cl[0]) + list(it.x_cl[1]))
        out[f"s{t}_x0_after"] = np.array([it.x_0.position[0], it.x_0.position[1], it.x_0.orientation, it.x_0.velocity])
        if replanned:
            opt = planner.optimal_trajectory
            st = np.stack([getattr(opt.cartesian, f) for f in ("x", "y", "theta", "v", "a", "kappa", "kappa_dot")] +
                          [getattr(opt.curvilinear, f) for f in ("s", "d", "theta", "s_dot", "s_ddot", "d_dot", "d_ddot")])
            out[f"s{t}_plan_xcl_in"] = np.array(list(planner.x_cl[0]) + list(planner.x_cl[1]))
            out[f"s{t}_opt_id"], out[f"s{t}_opt_cost"], out[f"s{t}_opt_states"] = opt.uniqueId, opt.cost, st
            out[f"s{t}_counts"] = np.array(planner._infeasible_count_kinematics, dtype=np.int64)
            out[f"s{t}_percentage"] = planner.infeasible_kinematics_percentage
            out[f"s{t}_collisions"] = planner.infeasible_count_collision
            out[f"s{t}_n_all_traj"] = len(planner.all_traj)
            out[f"s{t}_stats"] = np.array([planner.last_plan_stats.n_candidates, planner.last_plan_stats.n_collide,
                                           planner.last_plan_stats.n_boundary])
    out["n_history"] = len(planner.ego_vehicle_history)
    out["n_states"] = len(planner.record_state_list)
    return out, planner, it


def main():
    import ref_stubs
    ref_stubs.install()
    from cr_scenario_handler.planner_interfaces.frenet_interface import FrenetPlannerInterface
    from oracle_handler import OracleHandler
    out, planner, it = run(FrenetPlannerInterface, OracleHandler)
    np.savez_compressed(os.path.join(HERE, "interface_trace.npz"), **out)
    plans = [t for t in range(N_STEPS) if f"s{t}_opt_id" in out]
    print("plans at steps", plans, "selected", [int(out[f"s{t}_opt_id"]) for t in plans],
          "collisions", [int(out[f"s{t}_collisions"]) for t in plans], "stats", [out[f"s{t}_stats"].tolist() for t in plans])
    print("x_0 after 9 steps", out["s8_x0_after"], "history", out["n_history"], "states", out["n_states"])


if __name__ == "__main__":
    main()
