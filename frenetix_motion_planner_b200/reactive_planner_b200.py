"""``ReactivePlannerB200`` -- drop-in for ``ReactivePlannerPython`` / ``ReactivePlannerCpp`` whose inner loop
(sample -> back-project -> kinematic gates -> costs -> collision -> arg-min) runs on a B200 through
libfrx_b200.so.

Plug-in point in the reference: ``FrenetPlannerInterface.__init__`` picks the planner class
(cr_scenario_handler/planner_interfaces/frenet_interface.py:71-73); INTEGRATION.md shows the two-line
patch.  Same 7-argument constructor and the same public surface as
frenetix_motion_planner/reactive_planner.py:36-130 and the ``Planner`` base (planner.py:45-710):
``plan()``, ``update_externals()``, ``set_reference_and_coordinate_system()``, ``set_cost_function()``,
``set_predictions()``, ``set_x_0/set_x_cl/set_desired_velocity``, attributes ``all_traj``,
``optimal_trajectory``, ``trajectory_pair``, ``infeasible_count_collision``,
``_infeasible_count_kinematics``, ``infeasible_kinematics_percentage``, ``x_0``, ``x_cl``,
``coordinate_system``, ``record_state_list``, ``record_input_list``, ``ego_vehicle_history``.

Host-side work per plan() is what the reference also does on the host once per step (level sets,
sampling matrix, bookkeeping); nothing per-candidate happens in Python.  The CommonRoad object
conversions of the reference (``Trajectory`` / ``DynamicObstacle``) need commonroad-io; when it is not
importable, light-weight containers with the same attribute names are returned instead.
"""
from __future__ import annotations

import logging
import math
import time
from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np

from . import _capi, hotpath
from .coordinate_system import CoordinateSystem
from .sampling_matrix import SamplingHandler, python_path_rows, sampling_axes
from .trajectories import TrajectoryBundle, TrajectorySample, CartesianSample, CurviLinearSample

_EPS = 1e-5


def _get(obj, name, default=None):
    """Attribute-or-key access (OmegaConf objects, plain namespaces and dicts all work)."""
    if obj is None:
        return default
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


class PlannerState(SimpleNamespace):
    """Minimal stand-in for ReactivePlannerState (state.py:14-39): rear-axle position, orientation,
    velocity, acceleration, yaw_rate, steering_angle, time_step."""

    def shift_positions_to_center(self, wb_rear_axle: float):
        c = PlannerState(**self.__dict__)
        c.position = np.asarray(self.position, dtype=float) + wb_rear_axle * np.array(
            [np.cos(self.orientation), np.sin(self.orientation)])
        return c


class _Trajectory(SimpleNamespace):
    """(initial_time_step, state_list) like commonroad's Trajectory."""


class ReactivePlannerB200:
    def __init__(self, config_plan, config_sim, scenario=None, planning_problem=None, log_path=None, work_dir=None,
                 msg_logger=None, device: int = 0, handler=None):
        self.config_plan, self.config_sim = config_plan, config_sim
        planning, debug = _get(config_plan, "planning"), _get(config_plan, "debug")
        self.horizon = _get(planning, "planning_horizon")
        self.dT = _get(planning, "dt")
        self.N = int(self.horizon / self.dT)
        self._check_valid_settings()
        self.vehicle_params = _get(config_sim, "vehicle")
        self._low_vel_mode_threshold = _get(planning, "low_vel_mode_threshold", 2.0)
        self.msg_logger = msg_logger or logging.getLogger("Message_logger")
        self._multiproc = bool(_get(debug, "multiproc", False))
        self._num_workers = _get(debug, "num_workers", 1)
        self.x_0 = None
        self.x_cl: Optional[Tuple[List, List]] = None
        self.reference_path = None
        self.record_state_list, self.record_input_list, self.ego_vehicle_history = [], [], []
        self._LOW_VEL_MODE = False
        self.coordinate_system: Optional[CoordinateSystem] = None
        self.scenario, self.road_boundary = scenario, None
        self.planning_problem = planning_problem
        self.predictions = None
        self.reach_set = self.behavior = self.set_new_ref_path = None
        self.goal_status, self.full_goal_status, self.goal_area = False, None, None
        self.occlusion_module, self.use_occ_model = None, False
        self.goal_message = "Planner is in time step 0!"
        self.desired_velocity = None
        self._desired_d = 0.
        self.use_prediction = False
        self._collision_counter = 0
        self._total_count = 0
        self._infeasible_count_collision = 0
        self._optimal_cost = 0
        self.max_seen_costs = 1
        self._infeasible_count_kinematics = None
        self.infeasible_kinematics_percentage = None
        self._sampling_min = _get(planning, "sampling_min", 2)
        self._sampling_max = _get(planning, "sampling_max", 3)
        self.sampling_handler = SamplingHandler(dt=self.dT, max_sampling_number=self._sampling_max,
                                                t_min=_get(planning, "t_min", 1.1), horizon=self.horizon,
                                                delta_d_max=_get(planning, "d_max", 3), delta_d_min=_get(planning, "d_min", -3),
                                                d_ego_pos=_get(planning, "d_ego_pos", False))
        self.stopping_s = None
        self.log_risk = bool(_get(debug, "log_risk", False))
        self.save_all_traj = bool(_get(debug, "save_all_traj", False))
        self.all_traj = None
        self.optimal_trajectory = None
        self.trajectory_pair = None
        self.logger = None                      # DataLoggingCosts of the reference plugs in here when available
        self._draw_traj_set = bool(_get(debug, "draw_traj_set", True))
        self._kinematic_debug = bool(_get(debug, "kinematic_debug", True))
        # "cpp" sampling adds {N*dT}, {ss0} to the level sets like reactive_planner_cpp.py:235-237
        self.sampling_style = _get(debug, "sampling_style", "python")
        # the other things ReactivePlannerCpp does differently on the hot path (reactive_planner_cpp.py:109-112, 151-155,
        # 170-178): curvature-rate limit from vehicle.v_delta_max, prediction cost = collision probability
        # (CalculateCollisionProbabilityFast), velocity-offset cost with norm_order = 2
        self.cpp_flavour = bool(_get(debug, "cpp_flavour", False))
        self.emergency_mode = _get(planning, "emergency_mode", "risk")      # planning.yaml: "stopping" | "risk" (cpp path)
        if self.cpp_flavour:
            self.sampling_style = "cpp"
        # True: the reference tables and the Frenet initial state are computed on the device (frx_set_reference_polyline,
        # frx_initial_state) -- no numpy / CCosy work on the host between a polyline + Cartesian state and the plan
        self.device_frontend = bool(_get(debug, "device_frontend", False))
        self.static_obbs = None                 # road-boundary stand-in: [[cx, cy, theta, half_len, half_wid], ...]
        self.obstacle_order = None              # ids in scenario.obstacles order (collision_check.py:127-131)
        self.collision_check_enabled = True     # False: selection = first of the cost-sorted list (tests)

        self.cost_weights = dict(_get(_get(config_plan, "cost"), "cost_weights", {}) or {})
        self.set_cost_function(self.cost_weights)
        # risk / harm parameters (planner.py:156-159 loads configurations/{harm_parameters,risk}.json): only the LR1S
        # "ignore_angle" coefficients reach the hot path (boundary_harm, planner.py:370-381)
        from .trajectories import DEFAULT_HARM_COEFF
        self.params_harm = {"log_reg": {"ignore_angle": dict(DEFAULT_HARM_COEFF)}}
        self.params_risk = None

        # ---- device side: fails loudly if libfrx_b200.so or the GPU is missing (no CPU fallback).  `handler` lets a caller
        # share an existing context (and the drop-in tests substitute their recorder); nothing in the package passes it.
        self.handler = handler if handler is not None else _capi.Handler(device)
        self._bundle: Optional[TrajectoryBundle] = None
        self._device_tables = None
        self._static_dirty = True
        self._tables_key = None                 # durations the device's time tables were last built for
        self._prefetched = None                 # (input signature, optimal trajectory) left by prefetch_plans()
        self._ref_dirty = True
        self._pred_dirty = True
        self.last_plan_stats = None

    # ------------------------------------------------------------------------------------------
    # setters of the Planner base class (planner.py:172-310, 671-710)
    # ------------------------------------------------------------------------------------------
    @property
    def infeasible_count_collision(self):
        return self._collision_counter

    def set_cost_function(self, cost_weights):
        """Accepts a weight dict (what the cpp planner takes, reactive_planner_cpp.py:114-141) or an object with
        ``cost_weights`` (an AdaptableCostFunction of the reference)."""
        w = _get(cost_weights, "cost_weights", cost_weights)
        self.cost_weights = dict(w)
        self.cost_names, self.cost_weight_list = hotpath.active_costs(self.cost_weights)
        unsupported = [n for n in self.cost_names if n not in _capi.COST_ID]
        if unsupported:
            raise NotImplementedError(f"cost terms {unsupported} are host-only / unimplemented in the reference "
                                      f"(partial_cost_functions.py:67-117,133-138,199-293,359-387)")
        self.cost_function = SimpleNamespace(cost_weights=self.cost_weights, cost_weights_names=self.cost_names)

    def set_predictions(self, predictions: dict):
        self.use_prediction = True
        self.predictions = predictions
        self._pred_dirty = True

    def set_reference_and_coordinate_system(self, reference_path: np.ndarray = None, coordinate_system=None):
        """Reference tables for the device; an existing CoordinateSystem (e.g. the reference's CCosy wrapper)
        can be handed in instead of a polyline."""
        if coordinate_system is None and self.device_frontend:
            ref = np.ascontiguousarray(reference_path, dtype=np.float64)
            tab = self.handler.set_reference_polyline(ref)            # built on the device, read back once for the host view
            coordinate_system = CoordinateSystem.from_tables(ref, tab[0], tab[1], tab[2], tab[3])
            self._device_tables = coordinate_system
        elif coordinate_system is None:
            coordinate_system = CoordinateSystem(reference=reference_path)
        self.coordinate_system = coordinate_system
        self.reference_path = np.asarray(coordinate_system.reference)
        self.set_new_ref_path = True
        self._ref_dirty = True

    def set_scenario(self, scenario):
        """planner.py:550-565: keeps the scenario and, once, builds the road boundary from its lanelet network (here as
        thin static boxes, road_boundary.py -- the drivability checker's triangulated boundary is not available)."""
        self.scenario = scenario
        network = getattr(scenario, "lanelet_network", None)
        if self.static_obbs is None and network is not None:
            self.set_road_boundary(network)

    def set_road_boundary(self, lanelets, **kw):
        """`lanelets`: a commonroad LaneletNetwork, or the plain dict of road_boundary.lanelets_from_commonroad_xml."""
        from . import road_boundary as rb
        if not isinstance(lanelets, dict):
            lanelets = rb.lanelets_from_network(lanelets)
        self.set_static_obstacles(rb.road_boundary_obbs(lanelets, **kw))

    def set_static_obstacles(self, obbs):
        self.static_obbs = None if obbs is None else np.asarray(obbs, dtype=np.float64).reshape(-1, 5)
        self._static_dirty = True

    def set_x_0(self, x_0):
        self.x_0 = x_0
        self._LOW_VEL_MODE = bool(x_0.velocity < self._low_vel_mode_threshold)     # planner.py:222-229

    def set_x_cl(self, x_cl):
        # planner.py:231-237: a given x_cl is only trusted once one exists and the reference path is unchanged;
        # (extension) a caller without a Cartesian position can seed the Frenet state directly
        if x_cl is not None and ((self.x_cl is not None and not self.set_new_ref_path)
                                 or getattr(self.x_0, "position", None) is None):
            self.x_cl = x_cl
        else:
            self.x_cl = self._compute_initial_states(self.x_0)
        self.set_new_ref_path = False

    def set_desired_velocity(self, desired_velocity: float, current_speed: float = None, stopping: bool = False,
                             v_limit: float = 36):
        self.desired_velocity = desired_velocity
        a_max, v_max = _get(self.vehicle_params, "a_max"), _get(self.vehicle_params, "v_max")
        min_v = max(0.001, current_speed - a_max * self.horizon)                    # planner.py:304-306
        max_v = min(min(current_speed + (a_max / 6.0) * self.horizon, v_limit), v_max)
        self.sampling_handler.set_v_sampling(min_v, max_v)

    def set_goal_area(self, goal_area):
        self.goal_area = goal_area

    def set_planning_problem(self, planning_problem):
        self.planning_problem = planning_problem

    def set_occlusion_module(self, occ_module):
        raise NotImplementedError("the occlusion module re-ranks candidates on the host (external package); "
                                  "out of scope of the device hot path")

    def set_reach_set(self, reach_set):
        self.reach_set = reach_set

    def set_behavior(self, behavior):
        self.behavior = behavior

    def set_ego_vehicle_state(self, current_ego_vehicle):
        self.ego_vehicle_history.append(current_ego_vehicle)

    def set_sampling_parameters(self, t_min: float, horizon: float, delta_d_min: float, delta_d_max: float):
        self.sampling_handler.update_static_params(t_min, horizon, delta_d_min, delta_d_max)

    def record_state_and_input(self, state):
        self.record_state_list.append(state)
        if len(self.record_state_list) > 1:
            rate = (state.steering_angle - self.record_state_list[-2].steering_angle) / self.dT
        else:
            rate = 0.0
        self.record_input_list.append(SimpleNamespace(time_step=state.time_step, acceleration=state.acceleration,
                                                      steering_angle_speed=rate))

    def update_externals(self, scenario=None, reference_path=None, planning_problem=None, goal_area=None, x_0=None,
                         x_cl=None, cost_weights=None, occlusion_module=None, desired_velocity=None, predictions=None,
                         reach_set=None, behavior=None):
        """planner.py:172-217, same order of effects."""
        if scenario is not None:
            self.set_scenario(scenario)
        if reference_path is not None:
            self.set_reference_and_coordinate_system(reference_path)
        if planning_problem is not None:
            self.set_planning_problem(planning_problem)
        if goal_area is not None:
            self.set_goal_area(goal_area)
        if x_0 is not None:
            self.set_x_0(x_0)
            self.set_x_cl(x_cl)
        if cost_weights is not None:
            self.set_cost_function(cost_weights)
        if occlusion_module is not None:
            self.set_occlusion_module(occlusion_module)
        if desired_velocity is not None:
            self.set_desired_velocity(desired_velocity, x_0.velocity)
        if predictions is not None:
            self.set_predictions(predictions)
        if reach_set is not None:
            self.set_reach_set(reach_set)
        if behavior is not None:
            self.set_behavior(behavior)
        if self.sampling_handler.d_ego_pos:
            self.sampling_handler.set_d_sampling(self.x_cl[1][0])

    # ------------------------------------------------------------------------------------------
    # Frenet initial state (planner.py:567-635)
    # ------------------------------------------------------------------------------------------
    def _compute_initial_states(self, x_0):
        if getattr(self, "device_frontend", False) and getattr(self, "_device_tables", None) is self.coordinate_system:
            return self.handler.initial_state(x_0.position[0], x_0.position[1], x_0.orientation, x_0.velocity,
                                              getattr(x_0, "acceleration", 0.0), getattr(x_0, "steering_angle", 0.0),
                                              self._LOW_VEL_MODE, _get(self.vehicle_params, "wheelbase"))
        cs = self.coordinate_system
        s, d = cs.convert_to_curvilinear_coords(x_0.position[0], x_0.position[1])
        s_idx = int(np.argmax(cs.ref_pos > s)) - 1
        s_lambda = (s - cs.ref_pos[s_idx]) / (cs.ref_pos[s_idx + 1] - cs.ref_pos[s_idx])
        ref_theta = np.unwrap(cs.ref_theta)
        theta_ref = (ref_theta[s_idx + 1] - ref_theta[s_idx]) * (s - cs.ref_pos[s_idx]) / \
                    (cs.ref_pos[s_idx + 1] - cs.ref_pos[s_idx]) + ref_theta[s_idx]
        theta_ref = hotpath_make_valid_orientation(theta_ref)
        theta_cl = x_0.orientation - theta_ref
        kr = (cs.ref_curv[s_idx + 1] - cs.ref_curv[s_idx]) * s_lambda + cs.ref_curv[s_idx]
        kr_d = (cs.ref_curv_d[s_idx + 1] - cs.ref_curv_d[s_idx]) * s_lambda + cs.ref_curv_d[s_idx]
        wheelbase = _get(self.vehicle_params, "wheelbase")
        kappa_0 = np.tan(getattr(x_0, "steering_angle", 0.0)) / wheelbase
        d_p = (1 - kr * d) * np.tan(theta_cl)
        d_pp = -(kr_d * d + kr * d_p) * np.tan(theta_cl) + ((1 - kr * d) / (math.cos(theta_cl) ** 2)) * (
                kappa_0 * (1 - kr * d) / math.cos(theta_cl) - kr)
        s_velocity = x_0.velocity * math.cos(theta_cl) / (1 - kr * d)
        if s_velocity < 0:
            raise Exception("Initial state or reference incorrect! Curvilinear velocity is negative which indicates "
                            "that the ego vehicle is not driving in the same direction as specified by the reference")
        s_acceleration = getattr(x_0, "acceleration", 0.0)
        s_acceleration -= (s_velocity ** 2 / math.cos(theta_cl)) * (
                (1 - kr * d) * np.tan(theta_cl) * (kappa_0 * (1 - kr * d) / (math.cos(theta_cl)) - kr) -
                (kr_d * d + kr * d_p))
        s_acceleration /= ((1 - kr * d) / (math.cos(theta_cl)))
        if self._LOW_VEL_MODE:
            d_velocity, d_acceleration = d_p, d_pp
        else:
            d_velocity = x_0.velocity * math.sin(theta_cl)
            d_acceleration = s_acceleration * d_p + s_velocity ** 2 * d_pp
        return [s, s_velocity, s_acceleration], [d, d_velocity, d_acceleration]

    # ------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------
    def _push_static_inputs(self):
        h, cs, vp = self.handler, self.coordinate_system, self.vehicle_params
        h.set_params(dt=self.dT, N=self.N, low_vel_mode=self._LOW_VEL_MODE, draw_traj_set=self._draw_traj_set,
                     kinematic_debug=self._kinematic_debug, a_max=_get(vp, "a_max"), v_switch=_get(vp, "v_switch"),
                     delta_max=_get(vp, "delta_max"), wheelbase=_get(vp, "wheelbase"),
                     wb_rear_axle=_get(vp, "wb_rear_axle"), length=_get(vp, "length"), width=_get(vp, "width"),
                     x0_orientation=self.x_0.orientation, desired_velocity=self.desired_velocity,
                     cost_names=self.cost_names, cost_weights=self.cost_weight_list, store_states=True,
                     check_collisions=self.collision_check_enabled and (self.use_prediction or self.static_obbs is not None),
                     curvature_rate_from_v_delta=self.cpp_flavour, v_delta_max=_get(vp, "v_delta_max", 0.4),
                     velocity_offset_norm=2 if self.cpp_flavour else 1, prediction_cost_mode=1 if self.cpp_flavour else 0)
        if self._ref_dirty:
            if getattr(self, "_device_tables", None) is not cs:       # device-built tables are already where they belong
                ref = np.asarray(cs.reference)
                h.set_reference(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, ref[:, 0], ref[:, 1])
            self._ref_dirty = False
        if self._pred_dirty:
            packed = hotpath.pack_predictions(self.predictions, self.obstacle_order) if self.use_prediction else None
            if packed is None:
                h.set_predictions(None, None, None, [], [], [])
            else:
                h.set_predictions(*packed)
            self._pred_dirty = False
        if self._static_dirty:                   # uploads synchronise the stream: only when something changed
            h.set_static_obbs(self.static_obbs)
            self._static_dirty = False
        if "distance_to_obstacles" in self.cost_names:
            h.set_obstacle_positions(getattr(self, "obstacle_positions", None))

    def _level_axes(self, samp_level: int):
        """The (t1, ss1, d1) axes of a sampling level as arrays in the reference's set-iteration order: the candidates are
        their cartesian product with t1 slowest and d1 fastest, which is the loop nest of reactive_planner.py:149-158 (row
        index == the reference's uniqueId) and the row order of generate_sampling_matrix (reactive_planner_cpp.py:228-253)."""
        t_set, v_set, d_set = sampling_axes(self.sampling_handler, samp_level, self.x_cl,
                                            cpp_style=(self.sampling_style == "cpp"))
        return (np.fromiter(t_set, dtype=np.float64), np.fromiter(v_set, dtype=np.float64),
                np.fromiter(d_set, dtype=np.float64))

    def _sampling_matrix(self, samp_level: int) -> np.ndarray:
        return python_path_rows(*self._level_axes(samp_level), self.x_cl)

    def _prepare_level(self, samp_level: int, as_matrix: bool = False):
        """Host work of one sampling level: the three axes (or, for the multi-agent batch, the expanded matrix) and the
        time tables of its durations.  A level is ALWAYS a cartesian product, so the single-planner path hands the device
        the axes (frx_plan_grid, ~1 kB over PCIe) and never builds the [N, 13] matrix on the host."""
        axes = self._level_axes(samp_level)
        self._total_count = axes[0].size * axes[1].size * axes[2].size
        self._bundle = None                      # device buffers are recycled by the next plan
        key = (tuple(np.sort(axes[0]).tolist()), self.N)
        if key != self._tables_key:              # the level's durations only change with the sampling parameters
            self.handler.set_time_tables(*hotpath.time_tables(np.unique(axes[0]), self.dT, self.N + 1))
            self._tables_key = key
        return python_path_rows(*axes, self.x_cl) if as_matrix else axes

    def _make_bundle(self, sampling=None, axes=None) -> TrajectoryBundle:
        n = sampling.shape[0] if sampling is not None else axes[0].size * axes[1].size * axes[2].size
        self._bundle = TrajectoryBundle(self.handler, n, self.cost_names, self.cost_weight_list, self.dT,
                                        self.horizon, self.N + 1, self._LOW_VEL_MODE, sampling=sampling,
                                        grid=None if axes is None else (*axes, self.x_cl))
        return self._bundle

    def _create_trajectory_bundle(self, x_0_lon=None, x_0_lat=None, cost_function=None, samp_level: int = None) -> TrajectoryBundle:
        """reactive_planner.py:132-182.  The reference builds one TrajectorySample object per (t, v, d) here and evaluates
        them later; on the device sampling, feasibility, costs, collision sweep and arg-min are ONE launch, so the bundle
        that comes back is already evaluated (``bundle.result`` holds the plan's summary record)."""
        if x_0_lon is not None and x_0_lat is not None:
            self.x_cl = (list(x_0_lon), list(x_0_lat))
        if cost_function is not None and cost_function is not self.cost_function:
            self.set_cost_function(cost_function)
        axes = self._prepare_level(samp_level)
        res = hotpath.PlanOutput.from_result(self.handler.plan_grid(*axes, self.x_cl))
        bundle = self._make_bundle(axes=axes)
        bundle.result = res
        return bundle

    def _input_signature(self):
        """Everything a plan depends on that changes between simulation steps: a prefetched plan is only used by the
        plan() call that sees the very same inputs (and no other plan on this handler in between)."""
        x = self.x_0
        pos = getattr(x, "position", None)
        return (None if pos is None else (float(pos[0]), float(pos[1])), float(x.orientation), float(x.velocity),
                getattr(x, "time_step", None), tuple(float(v) for v in self.x_cl[0]), tuple(float(v) for v in self.x_cl[1]),
                None if self.desired_velocity is None else float(self.desired_velocity), id(self.predictions),
                id(self.coordinate_system), self.handler.generation)

    def plan(self) -> tuple:
        """reactive_planner.py:67-130 with the per-candidate work on the device."""
        optimal_trajectory = None
        t0 = time.time()
        pre, self._prefetched = self._prefetched, None
        if pre is not None and pre[0] == self._input_signature():
            # the multi-agent batch already evaluated this very plan in its one launch (prefetch_plans)
            return self._finish_plan(pre[1], time.time() - t0)
        self._push_static_inputs()
        samp_level = self._sampling_min
        while optimal_trajectory is None and samp_level < self._sampling_max:
            bundle = self._create_trajectory_bundle(self.x_cl[0], self.x_cl[1], self.cost_function, samp_level=samp_level)
            optimal_trajectory = self._get_optimal_trajectory(bundle, samp_level)
            samp_level += 1
        return self._finish_plan(optimal_trajectory, time.time() - t0)

    def _finish_plan(self, optimal_trajectory, planning_time: float):
        """reactive_planner.py:106-130: output conversion, stand-still fallback, post-processing."""
        self.trajectory_pair = self._compute_trajectory_pair(optimal_trajectory) if optimal_trajectory is not None else None
        if self.trajectory_pair is not None:
            self.set_ego_vehicle_state(self.convert_state_list_to_commonroad_object(self.trajectory_pair[0].state_list))
        if optimal_trajectory is None and self.x_0.velocity <= 0.1:
            self.msg_logger.warning('Planning standstill for the current scenario')
            optimal_trajectory = self._compute_standstill_trajectory()
        self.optimal_trajectory = optimal_trajectory
        self.plan_postprocessing(optimal_trajectory=optimal_trajectory, planning_time=planning_time)
        return self.trajectory_pair

    def _get_optimal_trajectory(self, bundle: TrajectoryBundle, samp_lvl, res=None):
        """reactive_planner.py:184-272: statistics, all_traj, selection (device arg-min == first collision-free
        entry of the cost-sorted feasible list)."""
        res = res if res is not None else bundle.result
        self._collision_counter = res.collision_counter
        if res.argmin is not None and int(res.argmin) >= 0:
            bundle.winner_row = int(res.argmin) - bundle.row_base
        counts = [0] * 11
        if self._multiproc and self._kinematic_debug:          # only then do the per-reason counts travel back
            counts = [int(v) for v in res.reason_counts]       # (reactive_planner.py:218-220)
        counts[0] = int(res.reason_counts[0])
        self._infeasible_count_kinematics = counts
        self.infeasible_kinematics_percentage = float(res.n_feasible / res.n_in_list) * 100 if res.n_in_list else 0.0
        self.last_plan_stats = res
        if self._draw_traj_set or self.save_all_traj:
            bundle.sort()                        # reads flags + costs of every row to the host: they outlive the next plan
            self.all_traj = bundle.trajectories
        if res.argmin >= 0:
            return bundle.sample(res.argmin - bundle.row_base).detach()
        if samp_lvl >= self._sampling_max - 1 and res.n_feasible > 0:
            if self.cpp_flavour and self.emergency_mode == "stopping":
                # reactive_planner_cpp.py:403-407: lowest end velocity, then shortest duration, then the lateral target
                # closest to the current offset, among the feasible candidates
                fl = bundle.flags
                feas = np.flatnonzero(((fl & _capi.FLAG_VALID) != 0) & ((fl & _capi.FLAG_FEASIBLE) != 0) &
                                      ((fl & _capi.FLAG_IN_LIST) != 0))
                self.msg_logger.warning("No optimal trajectory available. Select stopping trajectory!")
                params = np.array([bundle.sampling_row(int(r)) for r in feas]) if feas.size < 4096 else bundle.sampling_rows(feas)
                pick = self._select_stopping_row(params, self.x_cl[1][0])
                return bundle.sample(int(feas[pick])).detach()
            return self._select_min_risk(bundle)
        return None

    @staticmethod
    def _select_stopping_row(params: np.ndarray, d_pos: float) -> int:
        """Index into `params` ([n, 13] sampling rows of the feasible candidates) of the trajectory
        ReactivePlannerCpp._select_stopping_trajectory (reactive_planner_cpp.py:446-469) returns: the first hit when
        iterating end velocities ascending, durations ascending, lateral targets by distance to `d_pos`."""
        v, t, d = params[:, 5], params[:, 1], params[:, 10]
        d_vals = np.unique(d)
        d_rank = {val: k for k, val in enumerate(d_vals[np.argsort(np.abs(d_vals - d_pos))])}
        rank = np.array([d_rank[val] for val in d])
        return int(np.lexsort((rank, t, v))[0])

    @staticmethod
    def _select_stopping_trajectory(trajectories, sampling_matrix, d_pos):
        """Same call as the reference's static method: picks from a list of samples."""
        if not trajectories:
            return None
        params = np.array([t.sampling_parameters for t in trajectories])
        return trajectories[ReactivePlannerB200._select_stopping_row(params, d_pos)]

    def _select_min_risk(self, bundle: TrajectoryBundle):
        """reactive_planner.py:262-269: no collision-free candidate at the last level -> the feasible trajectory with the
        lowest ``ego_risk + obst_risk``.  The risk model (risk_assessment.calc_risk: harm x collision probability) lives
        on the host in the reference; a ``risk_function(sample) -> float`` attribute plugs it in unchanged.  Without one
        the device results give a proxy in the same spirit -- every remaining candidate collides or leaves the road, so
        prefer (1) staying on the road, (2) the LATEST first collision, (3) the lowest prediction cost (inverse
        Mahalanobis proximity to the predicted obstacles), (4) the total cost -- and say loudly that this is not the
        reference's risk."""
        fl = bundle.flags
        feas = np.flatnonzero(((fl & _capi.FLAG_VALID) != 0) & ((fl & _capi.FLAG_FEASIBLE) != 0) &
                              ((fl & _capi.FLAG_IN_LIST) != 0))
        if feas.size == 0:
            return None
        self.msg_logger.warning("No optimal trajectory available. Select lowest risk trajectory!")
        risk_fn = getattr(self, "risk_function", None)
        if risk_fn is not None:
            risks = np.array([risk_fn(bundle.sample(int(r))) for r in feas])
            return bundle.sample(int(feas[np.argmin(risks)])).detach()
        self.msg_logger.warning("risk_assessment is not attached (planner.risk_function): ranking by road departure, time "
                                "of first collision, prediction cost and total cost instead of ego_risk + obst_risk")
        f = fl[feas]
        off_road = (f & _capi.FLAG_BOUNDARY) != 0
        hit = (f & _capi.FLAG_COLLIDE) != 0
        first_hit = np.where(hit, (f >> _capi.FLAG_COLLIDE_STEP_SHIFT) & 63, 64).astype(np.int64)
        pred = bundle.costs[feas, self.cost_names.index("prediction")] if "prediction" in self.cost_names \
            else np.zeros(feas.size)
        order = np.lexsort((feas, bundle.total[feas], pred, -first_hit, off_road))     # last key is the primary one
        return bundle.sample(int(feas[order[0]])).detach()

    # ------------------------------------------------------------------------------------------
    # output conversion (planner.py:394-515) without commonroad-io
    # ------------------------------------------------------------------------------------------
    def _compute_trajectory_pair(self, trajectory: TrajectorySample) -> tuple:
        """planner.py:394-447: (Cartesian trajectory, curvilinear trajectory, lon samples, lat samples) of the selected
        candidate.  Same values as the reference's per-state loop; the arithmetic is done on whole arrays first and the
        state objects are filled from plain Python floats (the loop itself was most of a small plan's host time)."""
        c, cl = trajectory.cartesian, trajectory.curvilinear
        t0 = getattr(self.x_0, "time_step", 0)
        theta = np.asarray(c.theta, dtype=np.float64)
        n = theta.size
        yaw = np.empty(n)
        yaw[0] = getattr(self.x_0, "yaw_rate", 0.0)
        yaw[1:] = (theta[1:] - theta[:-1]) / self.dT
        steer = np.arctan2(_get(self.vehicle_params, "wheelbase") * np.asarray(c.kappa, dtype=np.float64), 1.0)
        # shift_orientation (planner.py:536-542) into [x_0.orientation - pi, x_0.orientation + pi]; almost always a no-op
        lo, hi = self.x_0.orientation - np.pi, self.x_0.orientation + np.pi
        th = theta.tolist()
        if theta.min() < lo or theta.max() > hi:
            for i, o in enumerate(th):
                while o < lo:
                    o += 2 * np.pi
                while o > hi:
                    o -= 2 * np.pi
                th[i] = o
        pos = np.stack([np.asarray(c.x, dtype=np.float64), np.asarray(c.y, dtype=np.float64)], axis=1)
        pos_cl = np.stack([np.asarray(cl.s, dtype=np.float64), np.asarray(cl.d, dtype=np.float64)], axis=1)
        v, a, kap = np.asarray(c.v).tolist(), np.asarray(c.a).tolist(), np.asarray(c.kappa).tolist()
        yaw_l, steer_l, th_raw = yaw.tolist(), steer.tolist(), theta.tolist()
        cart_list = [PlannerState(time_step=t0 + i, position=pos[i], orientation=th[i], velocity=v[i], acceleration=a[i],
                                  yaw_rate=yaw_l[i], steering_angle=steer_l[i]) for i in range(n)]
        cl_list = [SimpleNamespace(time_step=t0 + i, position=pos_cl[i], velocity=v[i], acceleration=a[i], orientation=th_raw[i],
                                   yaw_rate=kap[i]) for i in range(n)]
        lon_list = np.stack([cl.s, cl.s_dot, cl.s_ddot], axis=1).tolist()
        lat_list = np.stack([cl.d, cl.d_dot, cl.d_ddot], axis=1).tolist()
        return (_Trajectory(initial_time_step=t0, state_list=cart_list), _Trajectory(initial_time_step=t0, state_list=cl_list),
                lon_list, lat_list)

    def convert_state_list_to_commonroad_object(self, state_list, obstacle_id: int = 42):
        """planner.py:488-515: the ego as a dynamic obstacle whose positions are shifted from the rear axle to the centre."""
        wb_rear = _get(self.vehicle_params, "wb_rear_axle")
        ori = np.array([s.orientation for s in state_list], dtype=np.float64)
        pos = np.array([s.position for s in state_list], dtype=np.float64) + wb_rear * np.stack([np.cos(ori), np.sin(ori)], axis=1)
        shifted = []
        for k, s in enumerate(state_list):
            c = PlannerState(**s.__dict__)
            c.position = pos[k]
            shifted.append(c)
        return SimpleNamespace(obstacle_id=obstacle_id, initial_state=shifted[0],
                               obstacle_shape=SimpleNamespace(length=_get(self.vehicle_params, "length"),
                                                              width=_get(self.vehicle_params, "width")),
                               prediction=SimpleNamespace(trajectory=_Trajectory(
                                   initial_time_step=shifted[0].time_step, state_list=shifted)))

    def _compute_standstill_trajectory(self):
        """reactive_planner.py:579-626 (host only; used when nothing is selectable and v <= 0.1)."""
        x_0, (x_0_lon, x_0_lat) = self.x_0, self.x_cl
        N = self.N
        kappa_0 = np.tan(getattr(x_0, "steering_angle", 0.0)) / _get(self.vehicle_params, "wheelbase")
        a = np.repeat(0.0, N)
        a[1] = -x_0.velocity / self.dT
        cart = CartesianSample(np.repeat(x_0.position[0], N), np.repeat(x_0.position[1], N),
                               np.repeat(x_0.orientation, N), np.repeat(0.0, N), a, np.repeat(kappa_0, N),
                               np.repeat(0.0, N), current_time_step=N)
        cs = self.coordinate_system
        s_idx = int(np.argmax(cs.ref_pos > x_0_lon[0])) - 1
        ref_theta = np.unwrap(cs.ref_theta)
        th = (ref_theta[s_idx + 1] - ref_theta[s_idx]) * (x_0_lon[0] - cs.ref_pos[s_idx]) / \
             (cs.ref_pos[s_idx + 1] - cs.ref_pos[s_idx]) + ref_theta[s_idx]
        theta_cl = x_0.orientation - hotpath_make_valid_orientation(th)
        curv = CurviLinearSample(np.repeat(x_0_lon[0], N), np.repeat(x_0_lat[0], N), np.repeat(theta_cl, N), N,
                                 dd=np.repeat(x_0_lat[1], N), ddd=np.repeat(x_0_lat[2], N),
                                 ss=np.repeat(x_0_lon[1], N), sss=np.repeat(x_0_lon[2], N))
        return SimpleNamespace(cartesian=cart, curvilinear=curv, uniqueId=0, cost=0.0, feasible=True, valid=True,
                               horizon=self.horizon, dt=self.dT, costMap={n: (0, 0) for n in self.cost_names},
                               _ego_risk=None, _obst_risk=None, boundary_harm=None, _coll_detected=None,
                               actual_traj_length=N)

    # ------------------------------------------------------------------------------------------
    # the remaining members of the reference classes (planner.py / reactive_planner.py), device-backed
    # ------------------------------------------------------------------------------------------
    def _check_valid_settings(self):
        """planner.py:544-548"""
        assert self.dT > 0, 'provided dt is not correct! dt = {}'.format(self.dT)
        assert self.N > 0 and isinstance(self.N, int), 'N is not correct!'
        assert self.horizon > 0, 'provided t_h is not correct! dt = {}'.format(self.horizon)

    def set_stopping_point(self, stop_s_coordinate):
        self.stopping_s = stop_s_coordinate                   # planner.py:664-669

    def shift_orientation(self, trajectory, interval_start=-np.pi, interval_end=np.pi):
        """planner.py:536-542"""
        for state in trajectory.state_list:
            while state.orientation < interval_start:
                state.orientation += 2 * np.pi
            while state.orientation > interval_end:
                state.orientation -= 2 * np.pi
        return trajectory

    def _compute_cart_traj(self, trajectory):
        """planner.py:449-486: the Cartesian state list of a sample (yaw rate by np.gradient, no orientation shift)."""
        c = trajectory.cartesian
        t0 = getattr(self.x_0, "time_step", 0)
        yaw = np.gradient(np.asarray(c.theta)) / self.dT
        yaw[0] = getattr(self.x_0, "yaw_rate", 0.0)
        steer = np.arctan2(_get(self.vehicle_params, "wheelbase") * np.asarray(c.kappa), 1.0)
        return _Trajectory(initial_time_step=t0, state_list=[
            PlannerState(time_step=int(t0 + i), position=np.array([c.x[i], c.y[i]]), orientation=c.theta[i], velocity=c.v[i],
                         acceleration=c.a[i], yaw_rate=yaw[i], steering_angle=steer[i]) for i in range(len(c.x))])

    def create_coll_object(self, trajectory, vehicle_params=None, ego_state=None):
        """planner.py:517-534 builds a pycrcc time-variant obstacle of obb-sum hulls.  pycrcc is not a dependency of this
        package; what it would hold is returned as an array: one row per hull k = (cx, cy, theta, half_len, half_wid), the
        box in the frame of box k that contains the ego boxes k and k + 1 -- the same hulls the
        kernels test.  `trajectory`: the object convert_state_list_to_commonroad_object returns."""
        vp = vehicle_params or self.vehicle_params
        hl, hw = _get(vp, "length") / 2, _get(vp, "width") / 2
        st = trajectory.prediction.trajectory.state_list
        rows = []
        for a, b in zip(st[:-1], st[1:]):
            ux, uy = math.cos(a.orientation), math.sin(a.orientation)
            dx, dy = b.position[0] - a.position[0], b.position[1] - a.position[1]
            du, dv = dx * ux + dy * uy, dy * ux - dx * uy
            c = abs(ux * math.cos(b.orientation) + uy * math.sin(b.orientation))
            sn = abs(ux * math.sin(b.orientation) - uy * math.cos(b.orientation))
            eu, ev = hl * c + hw * sn, hl * sn + hw * c
            lo_u, hi_u, lo_v, hi_v = min(-hl, du - eu), max(hl, du + eu), min(-hw, dv - ev), max(hw, dv + ev)
            mu, mv = 0.5 * (lo_u + hi_u), 0.5 * (lo_v + hi_v)
            rows.append([a.position[0] + mu * ux - mv * uy, a.position[1] + mu * uy + mv * ux, a.orientation,
                         0.5 * (hi_u - lo_u), 0.5 * (hi_v - lo_v)])
        return np.array(rows)

    def check_feasibility(self, trajectories, queue_1=None, queue_2=None):
        """reactive_planner.py:274-577 evaluates the kinematics of a list of samples.  Samples of a device bundle arrive
        evaluated (``feasible`` / ``valid`` / the state arrays are read from the plan's result), so this only hands the
        list back -- through the queues when the caller passed them, like the multiprocessing variant (:207-224)."""
        trajectory_list = list(trajectories)
        if queue_1 is not None:
            queue_1.put(trajectory_list)
            if self._kinematic_debug and queue_2 is not None:
                queue_2.put(list(self._infeasible_count_kinematics or [0] * 11))
            return None
        return trajectory_list

    def trajectory_collision_check(self, feasible_trajectories):
        """planner.py:329-392 on the device's per-candidate verdicts: walk the (cost-sorted) list, count the candidates
        that meet a predicted obstacle, return the first one that neither collides nor leaves the road.  plan() does not
        call this -- the arg-min kernel already returns that candidate -- it serves callers that walk a list themselves."""
        for trajectory in feasible_trajectories:
            if self.use_occ_model and trajectory.valid is False:
                continue
            collision_detected = bool(trajectory._coll_detected) if self.use_prediction else False
            if collision_detected:
                self._collision_counter += 1
            boundary_harm = trajectory.boundary_harm
            trajectory.boundary_harm = boundary_harm
            trajectory._coll_detected = collision_detected
            if not collision_detected and boundary_harm == 0:
                return trajectory
        return None

    def set_risk_costs(self, trajectory):
        """planner.py:312-327 (risk_assessment.calc_risk, host side in the reference, out of scope here): with a
        ``risk_function(sample) -> float`` attached its value is booked as the ego risk; otherwise the device-side proxies
        are -- boundary harm for the ego, the inverse-Mahalanobis prediction cost for the obstacles."""
        risk_fn = getattr(self, "risk_function", None)
        if risk_fn is not None:
            trajectory._ego_risk, trajectory._obst_risk = float(risk_fn(trajectory)), 0.0
        else:
            pred = trajectory.costMap.get("prediction", (0.0, 0.0))[0] if hasattr(trajectory, "costMap") else 0.0
            trajectory._ego_risk, trajectory._obst_risk = float(trajectory.boundary_harm or 0.0), float(pred)
        return trajectory

    def plan_postprocessing(self, optimal_trajectory, planning_time, replanning_counter=0):
        """planner.py:637-649: hand the result to the reference's logger when one is attached."""
        if optimal_trajectory is not None and self.logger:
            self.logger.log(optimal_trajectory, time_step=self.x_0.time_step,
                            infeasible_kinematics=self._infeasible_count_kinematics,
                            percentage_kinematics=self.infeasible_kinematics_percentage, planning_time=planning_time,
                            ego_vehicle=self.ego_vehicle_history[-1], desired_velocity=self.desired_velocity,
                            replanning_counter=replanning_counter)
            self.logger.log_predicition(self.predictions)
        if self.save_all_traj and self.logger:
            self.logger.log_all_trajectories(self.all_traj, self.x_0.time_step)


def hotpath_make_valid_orientation(angle: float) -> float:
    """commonroad.common.util.make_valid_orientation (restated)."""
    two_pi = 2.0 * np.pi
    angle = angle % two_pi
    if np.pi <= angle <= two_pi:
        angle = angle - two_pi
    return angle


def prefetch_plans(planners: List["ReactivePlannerB200"]) -> None:
    """All agents of a multi-agent step in ONE eval-kernel launch per sampling level.

    The reference steps its agents one after the other (cr_scenario_handler/simulation/agent_batch.py:186-189 ->
    agent.py:230-236 -> planner_interface.update_planner / step_interface -> planner.plan()).  Call this between the
    agents' ``update_planner`` and their ``step_interface``: every planner does its host-side preparation, the sampling
    levels of all of them go through ``frx_plan_batched`` together (planners that found nothing re-enter the next round
    with their next level, exactly like the while-loop of reactive_planner.py:84-97), and each planner keeps the
    outcome.  Its next ``plan()`` -- called by the unmodified ``step_interface`` -- recognises the inputs, skips the
    device and only does the per-agent tail (output conversion, history, logging).  INTEGRATION.md shows the
    ``AgentBatch._step_agents`` patch."""
    level, optimal = {}, {}
    for p in planners:
        p._prefetched = None
        p._push_static_inputs()
        level[id(p)] = p._sampling_min
        optimal[id(p)] = None
    pending = [p for p in planners if level[id(p)] < p._sampling_max]
    while pending:
        mats = [p._prepare_level(level[id(p)], as_matrix=True) for p in pending]
        results = _capi.plan_batched([p.handler for p in pending], mats)
        nxt = []
        for p, S, r in zip(pending, mats, results):
            res = hotpath.PlanOutput.from_result(r)
            opt = p._get_optimal_trajectory(p._make_bundle(sampling=S), level[id(p)], res)
            level[id(p)] += 1
            optimal[id(p)] = opt
            if opt is None and level[id(p)] < p._sampling_max:
                nxt.append(p)
        pending = nxt
    for p in planners:
        p._prefetched = (p._input_signature(), optimal[id(p)])


def plan_batched(planners: List["ReactivePlannerB200"]) -> list:
    """prefetch_plans + every planner's own plan(): the trajectory pairs in the order of `planners`."""
    prefetch_plans(planners)
    return [p.plan() for p in planners]
