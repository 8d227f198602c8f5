"""Generate golden vectors by running the REFERENCE'S OWN code (unmodified files under
/root/reference, third-party imports stubbed by ``ref_stubs``).

Run once in the build container:  ``python tests/golden/make_golden.py``
Writes ``tests/golden/ref_*.npz``.  What is executed from the reference, per case:

* ``SamplingHandler`` / ``generate_sampling_matrix``       (sampling_matrix.py)
* ``ReactivePlannerPython._create_trajectory_bundle``      (reactive_planner.py:132-182)
  -> ``QuarticTrajectory`` / ``QuinticTrajectory``         (polynomial_trajectory.py)
* ``ReactivePlannerPython._get_optimal_trajectory``        (reactive_planner.py:184-272)
  -> ``check_feasibility``                                 (reactive_planner.py:274-577)
  -> ``TrajectoryBundle.sort`` -> ``AdaptableCostFunction.calc_cost`` (trajectories.py:524-561,
     cost_function.py:78-91) -> partial cost functions incl. ``get_inv_mahalanobis_dist``
* ``Planner.trajectory_collision_check`` is replaced by "first of the sorted list" because pycrcc
  is not installed -- collision parity is therefore NOT pinned by these files.

Third-party stand-ins used while generating (documented, parity-unpinned): the CCosy point
conversion (our definition, frenetix_motion_planner_b200/coordinate_system.py) and
``make_valid_orientation``.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_stubs  # noqa: E402

ref_stubs.install()

from frenetix_motion_planner.reactive_planner import ReactivePlannerPython  # noqa: E402
from frenetix_motion_planner.sampling_matrix import SamplingHandler, generate_sampling_matrix  # noqa: E402
from frenetix_motion_planner.cost_functions.cost_function import AdaptableCostFunction  # noqa: E402
import frenetix_motion_planner.cost_functions.partial_cost_functions as pcf  # noqa: E402

from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem  # noqa: E402
from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402


class _Log:
    def debug(self, *a, **k): pass
    info = warning = critical = error = debug


def build_planner(polyline, x_cl, x0_velocity, x0_orientation, desired_velocity, draw, debug,
                  predictions, cost_weights, dt=0.1, horizon=3.0, t_min=1.1, samp_max=3):
    veh = types.SimpleNamespace(**syn.VEHICLE_2)
    cfg = types.SimpleNamespace(
        cost=types.SimpleNamespace(cost_weights=dict(cost_weights)),
        debug=types.SimpleNamespace(save_unweighted_costs=False))
    p = object.__new__(ReactivePlannerPython)
    p.config_plan = cfg
    p.horizon, p.dT, p.N = horizon, dt, int(horizon / dt)
    p.vehicle_params = veh
    p.msg_logger = _Log()
    p._multiproc, p._num_workers = False, 1
    p.x_0 = types.SimpleNamespace(orientation=x0_orientation, velocity=x0_velocity, time_step=0,
                                  position=np.zeros(2))
    p.x_cl = x_cl
    p._LOW_VEL_MODE = x0_velocity < 2.0                       # planner.py:222-229
    p.coordinate_system = CoordinateSystem(polyline)
    p.scenario = None
    p.predictions = predictions
    p.reach_set = None
    p.occlusion_module = None
    p.use_occ_model = False
    p.desired_velocity = desired_velocity
    p._sampling_min, p._sampling_max = 2, samp_max
    p.sampling_handler = SamplingHandler(dt=dt, max_sampling_number=samp_max, t_min=t_min, horizon=horizon,
                                         delta_d_max=3, delta_d_min=-3, d_ego_pos=False)
    min_v, max_v = syn.velocity_interval(x0_velocity, veh.a_max, horizon, veh.v_max)
    p.sampling_handler.set_v_sampling(min_v, max_v)
    p.save_all_traj = False
    p.all_traj = None
    p._draw_traj_set, p._kinematic_debug = draw, debug
    p._collision_counter = 0
    p._infeasible_count_kinematics = None
    p.infeasible_kinematics_percentage = None
    p.log_risk = False
    p.logger = None
    p.cost_function = AdaptableCostFunction(rp=p, configuration=cfg)
    p.cost_function.update_state(scenario=None, rp=p, predictions=predictions, reachset=None)
    # pycrcc is unavailable: selection = first of the cost-sorted candidate list
    p.trajectory_collision_check = lambda feasible_trajectories: (feasible_trajectories[0]
                                                                   if feasible_trajectories else None)
    return p


FIELDS_CART = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot")
FIELDS_CL = ("s", "d", "theta", "s_dot", "s_ddot", "d_dot", "d_ddot")


def run_case(name, polyline, x_cl, v0, th0, v_des, draw, debug, n_obs, cost_weights=None, seed=7, preds_list=None,
             samp_level=2, samp_max=3):
    cost_weights = cost_weights or syn.DEFAULT_COST_WEIGHTS
    Nt = 31
    if preds_list is None:
        preds_list = syn.synthetic_predictions(polyline, n_obs, 31, 0.1, seed) if n_obs else []
    n_obs = len(preds_list)
    predictions = {100 + i: p for i, p in enumerate(preds_list)}
    pl = build_planner(polyline, x_cl, v0, th0, v_des, draw, debug, predictions, cost_weights, samp_max=samp_max)

    bundle = pl._create_trajectory_bundle(x_cl[0], x_cl[1], pl.cost_function, samp_level=samp_level)
    trajs = list(bundle.trajectories)
    n = len(trajs)
    # sampling matrix in *generation order* (python set iteration order of the reference)
    S = np.zeros((n, 13))
    coeffs = np.zeros((n, 12))
    for r, t in enumerate(trajs):
        assert t.uniqueId == r
        S[r] = [0.0, t.trajectory_long.delta_tau, x_cl[0][0], x_cl[0][1], x_cl[0][2],
                t.trajectory_long.x_d[0], t.trajectory_long.x_d[1],
                x_cl[1][0], x_cl[1][1], x_cl[1][2],
                t.trajectory_lat.x_d[0], t.trajectory_lat.x_d[1], t.trajectory_lat.x_d[2]]
        coeffs[r, :6] = t.trajectory_long.coeffs
        coeffs[r, 6:] = t.trajectory_lat.coeffs
    lat_delta_tau = np.array([t.trajectory_lat.delta_tau for t in trajs])

    optimal = pl._get_optimal_trajectory(bundle, samp_level)

    in_list = np.zeros(n, bool); stored = np.zeros(n, bool)
    feasible = np.zeros(n, bool); valid = np.zeros(n, bool)
    costed = np.zeros(n, bool)
    states = np.zeros((14, n, Nt)); traj_len = np.zeros(n, np.int32)
    total = np.zeros(n); names = list(pl.cost_function.cost_weights_names)
    costs = np.zeros((n, len(names))); costs_w = np.zeros((n, len(names)))
    # every sample object that survived is reachable from the original list (same objects)
    if draw:
        listed = pl.all_traj
    else:
        listed = None
    for r, t in enumerate(trajs):
        feasible[r] = bool(t.feasible) if t.feasible is not None else False
        valid[r] = bool(t.valid) if t.valid is not None else False
        if hasattr(t, "_cartesian"):
            stored[r] = True
            c, cl = t.cartesian, t.curvilinear
            for k, f in enumerate(FIELDS_CART):
                states[k, r] = getattr(c, f)
            for k, f in enumerate(FIELDS_CL):
                states[7 + k, r] = getattr(cl, f)
            traj_len[r] = t.actual_traj_length
        if t.cost != 0 or any(v != (0, 0) for v in t.costMap.values()):
            costed[r] = True
            total[r] = t.cost
            for k, nm in enumerate(names):
                costs[r, k], costs_w[r, k] = t.costMap[nm]
    if listed is not None:
        for t in listed:
            in_list[t.uniqueId] = True
        sorted_ids = np.array([t.uniqueId for t in listed], dtype=np.int64)
    else:
        sorted_ids = np.array([t.uniqueId for t in bundle.trajectories], dtype=np.int64)
        # trajectories_all is not kept by the reference in this mode; membership follows
        # reactive_planner.py:353-354,379,385,567 and is re-derived by the oracle
    out = dict(
        polyline=polyline, x_cl_lon=np.array(x_cl[0]), x_cl_lat=np.array(x_cl[1]),
        x0_velocity=v0, x0_orientation=th0, desired_velocity=v_des, draw=draw, debug=debug,
        low_vel_mode=pl._LOW_VEL_MODE, sampling=S, coeffs=coeffs, lat_delta_tau=lat_delta_tau,
        feasible=feasible, valid=valid, stored=stored, in_list=in_list, costed=costed,
        states=states, traj_len=traj_len, total=total, costs=costs, costs_weighted=costs_w,
        cost_names=np.array(names), cost_weights=np.array([cost_weights[k] for k in names]),
        sorted_ids=sorted_ids,
        infeasible_count_kinematics=np.array(pl._infeasible_count_kinematics, dtype=float),
        percentage=pl.infeasible_kinematics_percentage,
        optimal_id=(optimal.uniqueId if optimal is not None else -1),
        n_obs=n_obs,
        ref_pos=pl.coordinate_system.ref_pos, ref_theta=pl.coordinate_system.ref_theta,
        ref_curv=pl.coordinate_system.ref_curv, ref_curv_d=pl.coordinate_system.ref_curv_d,
    )
    for i, p in enumerate(preds_list):
        out[f"pred{i}_pos"] = p["pos_list"]; out[f"pred{i}_cov"] = p["cov_list"]
        out[f"pred{i}_ori"] = p["orientation_list"]
        out[f"pred{i}_shape"] = np.array([p["shape"]["length"], p["shape"]["width"]])
    path = os.path.join(HERE, f"ref_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} stored={stored.sum()} feasible={feasible.sum()} valid={valid.sum()} "
          f"costed={costed.sum()} optimal={out['optimal_id']} counts={out['infeasible_count_kinematics']}")


def sampling_matrix_case():
    """generate_sampling_matrix + level sets (sampling_matrix.py) known answers."""
    sh = SamplingHandler(dt=0.1, max_sampling_number=3, t_min=1.1, horizon=3.0, delta_d_max=3, delta_d_min=-3,
                         d_ego_pos=False)
    sh.set_v_sampling(0.001, 13.75)
    out = {}
    for lvl in range(3):
        out[f"t_{lvl}"] = np.array(sorted(sh.t_sampling.to_range(lvl)))
        out[f"v_{lvl}"] = np.array(sorted(sh.v_sampling.to_range(lvl)))
        out[f"d_{lvl}"] = np.array(sorted(sh.d_sampling.to_range(lvl)))
    t1 = np.array([1.1, 2.0, 3.0]); v1 = np.array([0.001, 5.0]); d1 = np.array([-1.0, 0.0, 1.0, 0.2])
    out["matrix"] = generate_sampling_matrix(t0_range=0.0, t1_range=t1, s0_range=10.0, ss0_range=8.0, sss0_range=0.5,
                                             ss1_range=v1, sss1_range=0, d0_range=0.2, dd0_range=0.1,
                                             ddd0_range=-0.1, d1_range=d1, dd1_range=0.0, ddd1_range=0.0)
    out["m_t1"], out["m_v1"], out["m_d1"] = t1, v1, d1
    np.savez_compressed(os.path.join(HERE, "ref_sampling.npz"), **out)
    print("sampling: matrix", out["matrix"].shape)


def inactive_cost_case():
    """acceleration / jerk / orientation_offset / path_length / distance_to_obstacles on a fixed
    sample (partial_cost_functions.py:24-46,141-151,172-196)."""
    rng = np.random.default_rng(11)
    out = {}
    for Nt in (31, 51, 30):
        a = rng.normal(0, 2, Nt); v = np.abs(rng.normal(8, 2, Nt)); th = rng.normal(0, 0.2, Nt)
        x = np.cumsum(v) * 0.1; y = rng.normal(0, 1, Nt)
        traj = types.SimpleNamespace(dt=0.1, cartesian=types.SimpleNamespace(a=a, v=v, x=x, y=y),
                                     curvilinear=types.SimpleNamespace(theta=th))
        obs_pos = np.array([[5.0, 2.0], [12.0, -3.0]])
        scen = types.SimpleNamespace(obstacles=[
            types.SimpleNamespace(state_at_time=lambda t, p=p: types.SimpleNamespace(position=p)) for p in obs_pos])
        planner = types.SimpleNamespace(x_0=types.SimpleNamespace(time_step=0))
        out[f"a_{Nt}"], out[f"v_{Nt}"], out[f"th_{Nt}"], out[f"x_{Nt}"], out[f"y_{Nt}"] = a, v, th, x, y
        out[f"obs_{Nt}"] = obs_pos
        out[f"acceleration_{Nt}"] = pcf.acceleration_costs(traj)
        out[f"jerk_{Nt}"] = pcf.jerk_costs(traj)
        out[f"orientation_offset_{Nt}"] = pcf.orientation_offset_costs(traj)
        out[f"path_length_{Nt}"] = pcf.path_length_costs(traj)
        out[f"distance_to_obstacles_{Nt}"] = pcf.distance_to_obstacles_costs(traj, planner=planner, scenario=scen)
    np.savez_compressed(os.path.join(HERE, "ref_inactive_costs.npz"), **out)
    print("inactive costs ok")


def refpath_cases():
    """extend_ref_path_both_ends / smooth_ref_path of the reference (utils_coordinate_system.py:20-58,110-134) on three
    dense centre lines -> ref_refpath.npz.  commonroad_dc's resample_polyline is stood in for by the package's resampler."""
    import commonroad_dc.geometry.util as cdc_util                 # stub module
    from frenetix_motion_planner_b200 import reference_path as rp
    cdc_util.resample_polyline = rp.resample_polyline
    import importlib
    ucs = importlib.import_module("cr_scenario_handler.utils.utils_coordinate_system")
    ucs.resample_polyline = rp.resample_polyline
    tj = np.load(os.path.join(HERE, "tjunction.npz"))["centre_line"]
    routes = {"tjunction": rp.resample_polyline(tj, 0.125), "scurve": rp.resample_polyline(syn.scurve_polyline(M=200), 0.125),
              "arc": rp.resample_polyline(syn.arc_polyline(R=45.0, M=90), 0.125)}
    out = {}
    for name, route in routes.items():
        ext = ucs.extend_ref_path_both_ends(route)
        out[f"{name}_route"], out[f"{name}_extended"], out[f"{name}_smooth"] = route, ext, ucs.smooth_ref_path(ext)
        out[f"{name}_extended_80"] = ucs.extend_ref_path_both_ends(route, 80)
    np.savez_compressed(os.path.join(HERE, "ref_refpath.npz"), **out)
    print("ref paths:", {k: v.shape for k, v in out.items() if k.endswith("_smooth")})


def collision_probability_module():
    """The reference's risk_assessment.collision_probability with its two absent third-party calls stood in for:
    scipy.stats.mvn.mvnun (removed from scipy) by multivariate_normal.cdf(upper, lower_limit=lower), which evaluates the same
    rectangle probability, and pycrcc.RectOBB (centre, half length, x axis) by a three-line class."""

    import importlib
    from scipy.stats import multivariate_normal
    cp = importlib.import_module("risk_assessment.collision_probability")

    class _Mvn:
        @staticmethod
        def mvnun(lower, upper, mu, cov):
            return multivariate_normal.cdf(np.asarray(upper), mean=np.asarray(mu), cov=np.asarray(cov), lower_limit=np.asarray(lower)), 0

    class _RectOBB:
        def __init__(self, r_x, r_y, orientation, cx, cy):
            self._r, self._o, self._c = r_x, orientation, np.array([cx, cy])

        def center(self):
            return self._c

        def r_x(self):
            return self._r

        def local_x_axis(self):
            return np.array([np.cos(self._o), np.sin(self._o)])
    cp.mvn = _Mvn
    cp.pycrcc = types.SimpleNamespace(RectOBB=_RectOBB)
    return cp


def collision_probability_cases():
    """get_collision_probability_fast of the reference (risk_assessment/collision_probability.py:141-261) on stored
    trajectories -> ref_collision_probability.npz.  Two absent third-party calls are stood in for: scipy.stats.mvn.mvnun
    (removed from scipy) by multivariate_normal.cdf(upper, lower_limit=lower), which evaluates the same rectangle
    probability, and pycrcc.RectOBB (centre, half length, x axis) by a three-line class."""
    cp = collision_probability_module()
    veh = types.SimpleNamespace(**syn.VEHICLE_2)
    out = {}
    n_case = 0
    for name in ("arc_hv_draw_pred", "tjunction_draw", "scurve_lowvel_draw"):
        g = np.load(os.path.join(HERE, f"ref_{name}.npz"))
        rows = np.flatnonzero(g["stored"])[::37][:6]
        for variant in range(2):
            preds = {}
            for o in range(int(g["n_obs"])):
                cov = g[f"pred{o}_cov"].copy()
                pos = g[f"pred{o}_pos"].copy()
                if variant == 1:                      # correlated covariances, a zero matrix, obstacles pulled next to the ego
                    for k in range(cov.shape[0]):
                        r = [0.5, -0.8, 0.95, -0.97, 0.2][(k + o) % 5]
                        sx, sy = 0.3 + 0.02 * k, 0.5 + 0.01 * k
                        cov[k] = [[sx * sx, r * sx * sy], [r * sx * sy, sy * sy]]
                    cov[3] = 0.0
                    pos = pos + (np.array([g["states"][0, rows[0], 5], g["states"][1, rows[0], 5]]) - pos[4]) * 0.9
                preds[100 + o] = {"pos_list": pos, "cov_list": cov, "orientation_list": g[f"pred{o}_ori"],
                                  "shape": {"length": float(g[f"pred{o}_shape"][0]), "width": float(g[f"pred{o}_shape"][1])}}
            for r in rows:
                traj = types.SimpleNamespace(cartesian=types.SimpleNamespace(x=g["states"][0, r], y=g["states"][1, r], theta=g["states"][2, r]))
                probs = cp.get_collision_probability_fast(traj, preds, veh)
                key = f"c{n_case}"
                out[key + "_xyt"] = np.stack([g["states"][0, r], g["states"][1, r], g["states"][2, r]])
                out[key + "_n_obs"] = len(preds)
                for o, oid in enumerate(preds):
                    out[f"{key}_o{o}_pos"], out[f"{key}_o{o}_cov"] = preds[oid]["pos_list"], preds[oid]["cov_list"]
                    out[f"{key}_o{o}_ori"], out[f"{key}_o{o}_shape"] = preds[oid]["orientation_list"], g[f"pred{o}_shape"]
                    out[f"{key}_o{o}_probs"] = probs[oid]
                n_case += 1
    out["n_cases"] = n_case
    np.savez_compressed(os.path.join(HERE, "ref_collision_probability.npz"), **out)
    nz = sum(int((out[k] > 0).sum()) for k in out if k.endswith("_probs"))
    print("collision probability cases:", n_case, "non-zero step probabilities:", nz)


def sampling_order_cases():
    """Iteration order of the reference's level SETS (python hash order: it fixes uniqueId = row index and every
    equal-cost tie) for random v / d intervals, all levels, plus the cpp path's unions -> ref_sampling_order.npz."""
    rng = np.random.default_rng(99)
    out = {}
    cases = []
    for k in range(24):
        v_lo = float(np.round(rng.uniform(0.001, 8.0), int(rng.integers(1, 6))))
        v_hi = v_lo + float(np.round(rng.uniform(0.5, 20.0), int(rng.integers(1, 6))))
        d_half = float(rng.choice([3.0, 2.5, 1.75, 4.0]))
        t_min = float(rng.choice([1.1, 0.9, 0.5]))
        horizon = float(rng.choice([3.0, 5.0]))
        sh = SamplingHandler(dt=0.1, max_sampling_number=4, t_min=t_min, horizon=horizon, delta_d_max=d_half,
                             delta_d_min=-d_half, d_ego_pos=False)
        sh.set_v_sampling(v_lo, v_hi)
        d0, ss0 = float(rng.uniform(-1, 1)), float(rng.uniform(v_lo, v_hi))
        cases.append([v_lo, v_hi, d_half, t_min, horizon, d0, ss0])
        for lvl in range(4):
            # reactive_planner.py:149-158: the python path iterates these very objects
            out[f"c{k}_l{lvl}_t"] = np.array(list(sh.t_sampling.to_range(lvl)))
            out[f"c{k}_l{lvl}_v"] = np.array(list(sh.v_sampling.to_range(lvl)))
            out[f"c{k}_l{lvl}_d"] = np.array(list(sh.d_sampling.to_range(lvl).union({d0})))
            # reactive_planner_cpp.py:235-237
            N = int(horizon / 0.1)
            out[f"c{k}_l{lvl}_t_cpp"] = np.array(list(sh.t_sampling.to_range(lvl).union({N * 0.1})))
            out[f"c{k}_l{lvl}_v_cpp"] = np.array(list(sh.v_sampling.to_range(lvl).union({ss0})))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "ref_sampling_order.npz"), **out)
    print("sampling order cases:", len(cases))


def independent_projection(cs, X):
    """(x, y) -> (s, d) WITHOUT the product's inverse: scan the forward map P(s) + d n(s) on a 1 mm raster for the sign
    change of the tangential miss (X - P(s)) . t(s) next to X, then bisect it.  Stands in for pycrccosy's
    convert_to_curvilinear_coords (absent), which is the inverse of ITS forward map in the same sense."""
    poly, pos, theta = np.asarray(cs.reference), cs.ref_pos, cs.ref_theta

    def frame(s):
        i = int(np.argmax(pos > s)) - 1
        lam = (s - pos[i]) / (pos[i + 1] - pos[i])
        P = (1 - lam) * poly[i] + lam * poly[i + 1]
        th = theta[i] + lam * (theta[i + 1] - theta[i])
        r = np.asarray(X) - P
        return r[0] * np.cos(th) + r[1] * np.sin(th), r[1] * np.cos(th) - r[0] * np.sin(th), float(np.hypot(*r))
    ss = np.arange(pos[0] + 1e-6, pos[-1] - 1e-3, 1e-3)
    vals = np.array([frame(s) for s in ss])
    near = vals[:, 2] < vals[:, 2].min() + 0.5
    k = np.flatnonzero(near[:-1] & (np.sign(vals[:-1, 0]) != np.sign(vals[1:, 0])))
    assert k.size >= 1, "no foot point found"
    k = k[np.argmin(vals[k, 2])]
    lo, hi = ss[k], ss[k + 1]
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if (frame(mid)[0] > 0) == (frame(lo)[0] > 0):
            lo = mid
        else:
            hi = mid
    s = 0.5 * (lo + hi)
    return np.array([s, frame(s)[1]])


class _ProjectingCS:
    """What Planner._compute_initial_states reads from its coordinate system (planner.py:567-635)."""

    def __init__(self, cs):
        self._cs = cs
        self.ref_pos, self.ref_theta, self.ref_curv, self.ref_curv_d = cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d

    def convert_to_curvilinear_coords(self, x, y):
        return independent_projection(self._cs, (x, y))


def reference_initial_state(cs, x_0, low_vel_mode):
    """The reference's OWN Planner._compute_initial_states (unmodified, planner.py:567-635) on a fake self."""
    from frenetix_motion_planner.planner import Planner
    me = types.SimpleNamespace(coordinate_system=_ProjectingCS(cs), vehicle_params=types.SimpleNamespace(**syn.VEHICLE_2),
                               _LOW_VEL_MODE=bool(low_vel_mode), msg_logger=_Log())
    lon, lat = Planner._compute_initial_states(me, x_0)
    return [float(v) for v in lon], [float(v) for v in lat]


def initial_state_cases():
    """Frenet initial states of random ego poses next to four reference paths, both velocity modes -> ref_initial_states.npz
    (pins ReactivePlannerB200._compute_initial_states and the device-side frx_initial_state)."""
    rng = np.random.default_rng(2026)
    tj = np.load(os.path.join(HERE, "tjunction.npz"))["reference_path"]
    paths = {"arc": syn.arc_polyline(R=60.0, M=220), "scurve": syn.scurve_polyline(M=220),
             "straight": syn.straight_polyline(120), "tjunction": tj}
    out = {}
    for name, poly in paths.items():
        cs = CoordinateSystem(poly)
        rows = []
        for k in range(12):
            s, d = rng.uniform(5.0, cs.ref_pos[-1] - 40.0), rng.uniform(-2.5, 2.5)
            X = cs.convert_to_cartesian_coords(s, d)
            i = int(np.argmax(cs.ref_pos > s)) - 1
            th = cs.ref_theta[i] + rng.uniform(-0.3, 0.3)
            low = k % 3 == 0
            v = rng.uniform(0.2, 1.9) if low else rng.uniform(2.1, 15.0)
            x_0 = types.SimpleNamespace(position=np.array(X), orientation=float(th), velocity=float(v),
                                        acceleration=float(rng.uniform(-2, 2)), yaw_rate=0.0,
                                        steering_angle=float(rng.uniform(-0.2, 0.2)), time_step=0)
            lon, lat = reference_initial_state(cs, x_0, low)
            rows.append([X[0], X[1], x_0.orientation, x_0.velocity, x_0.acceleration, x_0.steering_angle, float(low)] + lon + lat)
        out[f"{name}_polyline"] = poly
        out[f"{name}_cases"] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "ref_initial_states.npz"), **out)
    print("initial states:", {k: v.shape for k, v in out.items() if k.endswith("_cases")})


def tjunction_inputs():
    """Inputs of the ZAM_Tjunction-1_42_T-1 fixture (tests/golden/tjunction.npz): smoothed reference path, the
    ego's Frenet state at time step 0 and the ground-truth predictions of the five cars."""
    g = np.load(os.path.join(HERE, "tjunction.npz"))
    poly = g["reference_path"]
    cs = CoordinateSystem(poly)
    x_0 = types.SimpleNamespace(position=g["ego_position_rear"], orientation=float(g["ego_orientation"]),
                                velocity=float(g["ego_velocity"]), acceleration=float(g["ego_acceleration"]),
                                yaw_rate=float(g["ego_yaw_rate"]), steering_angle=0.0, time_step=0)
    # the reference's own Frenet-state code on an independent projection -- nothing of the product is involved
    x_cl = reference_initial_state(cs, x_0, x_0.velocity < 2.0)
    preds = []
    for o in range(g["obstacle_states"].shape[0]):
        st = g["obstacle_states"][o, 1:32]                       # time steps 1..31 (prediction_helpers.py:239-247)
        preds.append({"pos_list": st[:, :2].copy(), "cov_list": np.tile(np.array([[0.1, 0.0], [0.0, 0.1]]), (31, 1, 1)),
                      "orientation_list": st[:, 2].copy(), "v_list": st[:, 3].copy(),
                      "shape": {"length": float(g["obstacle_shapes"][o, 0]) + 0.5,
                                "width": float(g["obstacle_shapes"][o, 1]) + 0.2}})
    return poly, x_cl, x_0, preds


if __name__ == "__main__":
    if "--initial-states" in sys.argv:
        initial_state_cases()
        sampling_order_cases()
        refpath_cases()
        collision_probability_cases()
        sys.exit(0)
    if "--tjunction" in sys.argv:
        poly, x_cl, x_0, preds = tjunction_inputs()
        print("x_cl", x_cl)
        run_case("tjunction_draw", poly, x_cl, x_0.velocity, x_0.orientation, 8.0, True, True, 0, preds_list=preds)
        run_case("tjunction_nodraw", poly, x_cl, x_0.velocity, x_0.orientation, 8.0, False, False, 0, preds_list=preds)
        sys.exit(0)
    straight = syn.straight_polyline(200)
    arc = syn.arc_polyline(R=60.0, M=220)
    scurve = syn.scurve_polyline(M=220)
    x_cl_a = ([10.0, 8.0, 0.0], [0.2, 0.0, 0.0])
    x_cl_b = ([12.0, 9.5, 0.4], [-0.3, 0.2, -0.1])
    x_cl_low = ([15.0, 1.2, 0.3], [0.1, 0.01, 0.0])
    x_cl_slow = ([20.0, 2.5, -0.5], [0.4, -0.05, 0.02])
    run_case("straight_hv_draw", straight, x_cl_a, 8.0, 0.0, 8.0, True, True, 0)
    run_case("arc_hv_draw_pred", arc, x_cl_b, 9.5, 0.2, 10.0, True, True, 5)
    run_case("arc_hv_nodraw_nodebug", arc, x_cl_b, 9.5, 0.2, 10.0, False, False, 5)
    run_case("arc_hv_nodraw_debug", arc, x_cl_b, 9.5, 0.2, 10.0, False, True, 3)
    run_case("scurve_lowvel_draw", scurve, x_cl_low, 1.2, 0.1, 3.0, True, True, 4)
    run_case("scurve_lowvel_nodraw", scurve, x_cl_low, 1.2, 0.1, 3.0, False, False, 4)
    run_case("scurve_slow_hv_draw", scurve, x_cl_slow, 2.5, 0.3, 1.0, True, True, 2)
    run_case("scurve_slow_hv_nodraw", scurve, x_cl_slow, 2.5, 0.3, 1.0, False, False, 2)
    # invalid candidates: s_dot < -EPS (hard braking overshoot) and out-of-projection-domain (short path)
    x_cl_brake = ([20.0, 3.0, -6.0], [0.4, -0.05, 0.02])
    run_case("scurve_brake_hv_draw", scurve, x_cl_brake, 3.0, 0.3, 1.0, True, True, 2)
    run_case("scurve_brake_hv_nodraw_debug", scurve, x_cl_brake, 3.0, 0.3, 1.0, False, True, 2)
    run_case("scurve_brake_hv_nodraw_nodebug", scurve, x_cl_brake, 3.0, 0.3, 1.0, False, False, 2)
    short = syn.straight_polyline(40)
    run_case("short_hv_draw", short, x_cl_a, 8.0, 0.0, 8.0, True, True, 2)
    run_case("short_hv_nodraw", short, x_cl_a, 8.0, 0.0, 8.0, False, False, 2)
    sampling_matrix_case()
    inactive_cost_case()
    initial_state_cases()
    sampling_order_cases()
    refpath_cases()
    collision_probability_cases()
