"""Live differential sweep (build container only: needs /root/reference): the REFERENCE'S OWN code (the same entry points
make_golden.py records, third-party imports stubbed) against the numpy oracle on RANDOM cases -- reference path kind and
shape, Frenet state, speed regime, desired velocity, debug flags, obstacle sets -- with the assertions of
tests/test_oracle_golden.py.  The committed goldens are 15 hand-picked cases; this widens the pin.

    python tests/golden/sweep_reference_vs_oracle.py [n_cases] [first_seed]
    python tests/golden/sweep_reference_vs_oracle.py --initial-states [n_poses] [first_seed]     (Frenet front end)
    python tests/golden/sweep_reference_vs_oracle.py --collision-probability [n] [first_seed]   (cpp prediction cost)
    python tests/golden/sweep_reference_vs_oracle.py --sampling-order [n] [first_seed]          (level sets and their order, a1)
    python tests/golden/sweep_reference_vs_oracle.py --refpath [n] [first_seed]                 (reference-path preparation, 8f-2)
    python tests/golden/sweep_reference_vs_oracle.py --trajectory-pair [n] [first_seed]         (output conversion of plan(), a13)
    python tests/golden/sweep_reference_vs_oracle.py --inactive-costs [n] [first_seed]          (the five optional cost terms, a9)

Prints one line per case and a summary; exit code 1 on any mismatch.  Nothing is written into the repository."""
import os
import sys
import tempfile
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402  (installs the stubs, imports the reference)
import helpers  # noqa: E402
import test_oracle_golden as tog  # noqa: E402
from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402


def random_case(seed):
    rng = np.random.default_rng(seed)
    kind = int(rng.integers(0, 4))
    poly = [syn.straight_polyline(260), syn.arc_polyline(R=float(rng.uniform(35, 300)), M=260),
            syn.scurve_polyline(M=260, amp=float(rng.uniform(1, 6))),
            syn.arc_polyline(R=float(rng.uniform(50, 200)), M=200, start_heading=float(rng.uniform(-1, 1)))][kind]
    low = bool(rng.integers(0, 3) == 0)
    v0 = float(rng.uniform(0.3, 1.9)) if low else float(rng.uniform(2.1, 14.0))
    x_cl = ([float(rng.uniform(5, 40)), v0, float(rng.uniform(-2, 2))],
            [float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.2, 0.2))])
    th0 = float(rng.uniform(-0.3, 0.3))
    v_des = float(rng.uniform(0.5, 14.0))
    draw, debug = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    n_obs = int(rng.integers(0, 7))
    level = int(rng.choice([1, 2, 2, 2, 3], p=[0.2, 0.25, 0.25, 0.25, 0.05]))      # 3 x 3 x 4 ... 17 x 17 x 18 end states per duration
    weights = None
    if rng.integers(0, 3) == 0:        # other weights, and the optional terms switched on (cost_function.py:55-91 picks up w > 0)
        weights = {k: float(np.round(w * rng.uniform(0.2, 3.0), 3)) for k, w in syn.DEFAULT_COST_WEIGHTS.items()}
        for k in ("acceleration", "jerk", "orientation_offset", "path_length"):
            if rng.integers(0, 2):
                weights[k] = float(np.round(rng.uniform(0.01, 2.0), 3))
    return dict(polyline=poly, x_cl=x_cl, v0=v0, th0=th0, v_des=v_des, draw=draw, debug=debug, n_obs=n_obs, seed=int(seed), level=level,
                weights=weights)


def initial_state_sweep(n_poses, first):
    """Planner._compute_initial_states of the reference (planner.py:567-635, unmodified, on make_golden's independent
    projection) against ReactivePlannerB200._compute_initial_states on random paths and ego poses, both velocity modes."""
    import types
    from frenetix_motion_planner_b200.reactive_planner_b200 import ReactivePlannerB200
    from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem
    bad = 0
    worst = 0.0
    t_all = time.time()
    for k in range(n_poses):
        rng = np.random.default_rng(first + k)
        kind = int(rng.integers(0, 3))
        poly = [syn.straight_polyline(120), syn.arc_polyline(R=float(rng.uniform(30, 300)), M=160),
                syn.scurve_polyline(M=160, amp=float(rng.uniform(1, 6)))][kind]
        cs = CoordinateSystem(poly)
        s, d = rng.uniform(5.0, cs.ref_pos[-1] - 40.0), rng.uniform(-3.0, 3.0)
        X = cs.convert_to_cartesian_coords(s, d)
        i = int(np.argmax(cs.ref_pos > s)) - 1
        low = bool(rng.integers(0, 3) == 0)
        x_0 = types.SimpleNamespace(position=np.array(X), orientation=float(cs.ref_theta[i] + rng.uniform(-0.4, 0.4)),
                                    velocity=float(rng.uniform(0.2, 1.9) if low else rng.uniform(2.1, 15.0)),
                                    acceleration=float(rng.uniform(-3, 3)), yaw_rate=0.0,
                                    steering_angle=float(rng.uniform(-0.3, 0.3)), time_step=0)
        want = np.array(sum(mg.reference_initial_state(cs, x_0, low), []))
        me = types.SimpleNamespace(coordinate_system=cs, vehicle_params=types.SimpleNamespace(**syn.VEHICLE_2), _LOW_VEL_MODE=low)
        lon, lat = ReactivePlannerB200._compute_initial_states(me, x_0)
        got = np.array(list(lon) + list(lat))
        err = float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))
        worst = max(worst, err)
        if not err <= 1e-9:
            bad += 1
            print(f"pose seed {first + k}: MISMATCH err {err:.3e} got {got} want {want}", flush=True)
    print(f"{n_poses - bad} of {n_poses} random ego poses: Frenet initial state == reference (worst {worst:.2e}, {time.time() - t_all:.0f} s)")
    return bad


def collision_probability_sweep(n_cases, first):
    """get_collision_probability_fast of the reference (risk_assessment/collision_probability.py:141-261; scipy's rectangle
    probability behind its mvn call) against the oracle's restatement with its own Genz bivariate-normal routine: random
    ego trajectories, obstacles within a few metres of them, random correlations (all three |rho| branches of the algorithm),
    now and then a zero covariance (the ground-truth case, :214-216)."""
    import types
    from oracle import frenet_oracle as fo
    cp = mg.collision_probability_module()
    veh = types.SimpleNamespace(**syn.VEHICLE_2)
    bad, worst, nonzero = 0, 0.0, 0
    t_all = time.time()
    for k in range(n_cases):
        rng = np.random.default_rng(first + k)
        n = int(rng.integers(8, 32))
        t = np.arange(n) * 0.1
        v, yaw_rate, th0 = rng.uniform(0.5, 14.0), rng.uniform(-0.3, 0.3), rng.uniform(-3.1, 3.1)
        th = th0 + yaw_rate * t
        x = rng.uniform(-50, 50) + np.cumsum(v * 0.1 * np.cos(th))
        y = rng.uniform(-50, 50) + np.cumsum(v * 0.1 * np.sin(th))
        preds = {}
        for o in range(int(rng.integers(1, 5))):
            m = int(rng.integers(4, n + 3))
            off = rng.uniform(-6, 6, 2)
            j = np.minimum(np.arange(m), n - 1)
            pos = np.stack([x[j], y[j]], axis=1) + off + rng.normal(0, 0.4, (m, 2))
            cov = np.zeros((m, 2, 2))
            for q in range(m):
                rho = float(rng.choice([0.0, rng.uniform(-0.29, 0.29), rng.uniform(0.3, 0.74), -rng.uniform(0.3, 0.74),
                                        rng.uniform(0.75, 0.92), rng.uniform(0.93, 0.995), -rng.uniform(0.93, 0.995)]))
                sx, sy = rng.uniform(0.1, 1.5), rng.uniform(0.1, 1.5)
                cov[q] = [[sx * sx, rho * sx * sy], [rho * sx * sy, sy * sy]]
                if rng.integers(0, 12) == 0:
                    cov[q] = 0.0
            preds[100 + o] = {"pos_list": pos, "cov_list": cov, "orientation_list": rng.uniform(-3.1, 3.1, m),
                              "shape": {"length": float(rng.uniform(3, 6)), "width": float(rng.uniform(1.5, 2.5))}}
        traj = types.SimpleNamespace(cartesian=types.SimpleNamespace(x=x, y=y, theta=th))
        want = cp.get_collision_probability_fast(traj, preds, veh)
        got = fo.collision_probability_fast(x, y, th, list(preds.values()), veh.length, veh.width)
        for o, oid in enumerate(preds):
            w = np.asarray(want[oid])
            err = float(np.abs(got[o] - w).max()) if got[o].shape == w.shape else float("inf")
            worst = max(worst, err)
            nonzero += int((w > 0).sum())
            if not err < 1e-12:
                bad += 1
                print(f"collision-probability seed {first + k} obstacle {o}: MISMATCH err {err:.3e}", flush=True)
    print(f"{n_cases} random trajectories, {nonzero} non-zero step probabilities: {'all equal' if not bad else str(bad) + ' MISMATCHES'} "
          f"(worst {worst:.2e}, {time.time() - t_all:.0f} s)")
    return bad


def sampling_order_sweep(n_cases, first):
    """Level sets of the reference's SamplingHandler (sampling_matrix.py:17-195) and their ITERATION ORDER (python set order:
    it fixes uniqueId = row index and how equal-cost ties break) against the package's handler + sampling_axes, python and
    cpp style, for random v / d / t configurations and every level."""
    from frenetix_motion_planner_b200.sampling_matrix import SamplingHandler as OurHandler, sampling_axes
    bad = 0
    for k in range(n_cases):
        rng = np.random.default_rng(first + k)
        v_lo = float(np.round(rng.uniform(0.001, 8.0), int(rng.integers(1, 6))))
        v_hi = v_lo + float(np.round(rng.uniform(0.5, 20.0), int(rng.integers(1, 6))))
        d_half = float(rng.choice([3.0, 2.5, 1.75, 4.0, float(np.round(rng.uniform(1, 5), 2))]))
        t_min = float(rng.choice([1.1, 0.9, 0.5, 1.5]))
        horizon = float(rng.choice([3.0, 5.0, 4.0, 2.0]))
        levels = int(rng.integers(3, 6))
        kw = dict(dt=0.1, max_sampling_number=levels, t_min=t_min, horizon=horizon, delta_d_max=d_half, delta_d_min=-d_half, d_ego_pos=False)
        ref, our = mg.SamplingHandler(**kw), OurHandler(**kw)
        ref.set_v_sampling(v_lo, v_hi); our.set_v_sampling(v_lo, v_hi)
        d0, ss0 = float(rng.uniform(-1, 1)), float(rng.uniform(v_lo, v_hi))
        x_cl = ([3.0, ss0, 0.0], [d0, 0.0, 0.0])
        N = int(horizon / 0.1)
        for lvl in range(levels):
            t, v, d = sampling_axes(our, lvl, x_cl)
            tc, vc, dc = sampling_axes(our, lvl, x_cl, cpp_style=True)
            ok = (list(t) == list(ref.t_sampling.to_range(lvl)) and list(v) == list(ref.v_sampling.to_range(lvl))
                  and list(d) == list(ref.d_sampling.to_range(lvl).union({d0}))
                  and list(tc) == list(ref.t_sampling.to_range(lvl).union({N * 0.1}))
                  and list(vc) == list(ref.v_sampling.to_range(lvl).union({ss0})))
            if not ok:
                bad += 1
                print(f"sampling seed {first + k} level {lvl}: MISMATCH", flush=True)
    print(f"{n_cases} random sampling configurations x every level: {'all equal' if not bad else str(bad) + ' MISMATCHES'}")
    return bad


def refpath_sweep(n_cases, first):
    """extend_ref_path_both_ends / smooth_ref_path of the reference (utils_coordinate_system.py:20-58,110-134) against
    frenetix_motion_planner_b200.reference_path on random dense centre lines.  (commonroad_dc's resample_polyline is
    absent: both sides use the package's resampler, as in make_golden.refpath_cases -- the extension, the chaikin
    smoothing and the spline pipeline around it are the reference's.)"""
    import importlib
    import commonroad_dc.geometry.util as cdc_util                 # stub module
    from frenetix_motion_planner_b200 import reference_path as rp
    cdc_util.resample_polyline = rp.resample_polyline
    ucs = importlib.import_module("cr_scenario_handler.utils.utils_coordinate_system")
    ucs.resample_polyline = rp.resample_polyline
    bad, worst = 0, 0.0
    for k in range(n_cases):
        rng = np.random.default_rng(first + k)
        kind = int(rng.integers(0, 3))
        raw = [syn.arc_polyline(R=float(rng.uniform(25, 200)), M=int(rng.integers(60, 160)), start_heading=float(rng.uniform(-3, 3))),
               syn.scurve_polyline(M=int(rng.integers(120, 260)), amp=float(rng.uniform(1, 8))),
               syn.straight_polyline(int(rng.integers(60, 200)))][kind]
        if kind == 2:                                              # a straight line with a gentle random wobble
            raw = raw + np.stack([np.zeros(len(raw)), np.cumsum(rng.normal(0, 0.02, len(raw)))], axis=1)
        route = rp.resample_polyline(raw, float(rng.choice([0.125, 0.25, 0.5])))
        ext_r, ext_o = ucs.extend_ref_path_both_ends(route), rp.extend_ref_path_both_ends(route)
        ok = np.array_equal(ext_r, ext_o)
        far = float(rng.choice([30, 80, 120]))
        ok = ok and np.array_equal(ucs.extend_ref_path_both_ends(route, far), rp.extend_ref_path_both_ends(route, far))
        sm_r, sm_o = ucs.smooth_ref_path(ext_r), rp.smooth_ref_path(ext_o)
        err = float(np.abs(sm_r - sm_o).max()) if sm_r.shape == sm_o.shape else float("inf")
        worst = max(worst, err)
        if not (ok and err <= 1e-10):
            bad += 1
            print(f"refpath seed {first + k}: MISMATCH extended_equal {ok} smooth err {err:.3e}", flush=True)
    print(f"{n_cases} random centre lines: {'all equal' if not bad else str(bad) + ' MISMATCHES'} (smoothed path: worst {worst:.2e} m)")
    return bad


def trajectory_pair_sweep(n_cases, first):
    """Planner._compute_trajectory_pair + shift_orientation of the reference (planner.py:394-447,536-542) -- what every caller
    of plan() consumes -- against ReactivePlannerB200._compute_trajectory_pair on random selected trajectories (oracle states of
    random cases), random x_0 time step / yaw rate / orientation (incl. orientations that force the 2 pi shift).  commonroad's
    state / trajectory containers are absent: the reference function fills plain attribute bags instead."""
    import types
    import importlib
    from oracle import frenet_oracle as fo
    from frenetix_motion_planner_b200.reactive_planner_b200 import ReactivePlannerB200
    from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem
    pl = importlib.import_module("frenetix_motion_planner.planner")

    class Bag(types.SimpleNamespace):
        pass
    pl.ReactivePlannerState = pl.CustomState = Bag
    pl.Trajectory = lambda t0, states: types.SimpleNamespace(initial_time_step=t0, state_list=states)
    bad, worst = 0, 0.0
    for k in range(n_cases):
        c = random_case(first + k)
        rng = np.random.default_rng(first + k + 7)
        cs = CoordinateSystem(c["polyline"])
        ref = fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, np.ascontiguousarray(c["polyline"][:, 0]),
                         np.ascontiguousarray(c["polyline"][:, 1]))
        prm = fo.Params(low_vel_mode=c["v0"] < 2.0, x0_orientation=c["th0"], desired_velocity=c["v_des"], draw_traj_set=True,
                        **{q: syn.VEHICLE_2[q] for q in ("a_max", "v_switch", "delta_max", "wheelbase", "wb_rear_axle", "length", "width")})
        v_lo, v_hi = syn.velocity_interval(c["v0"], syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"])
        S = syn.grid_sampling_matrix(np.array([1.1, 2.0, 3.0]), np.linspace(v_lo, v_hi, 3), np.linspace(-3, 3, 4), c["x_cl"])
        out = fo.plan(S, ref, prm, [], collision_check=False)
        stored = np.flatnonzero((out["flags"] & fo.FLAG_STORED) != 0)
        for r in stored[:: max(1, len(stored) // 4)]:
            st = out["states"][:, r, :]
            traj = types.SimpleNamespace(
                cartesian=types.SimpleNamespace(x=st[0], y=st[1], theta=st[2], v=st[3], a=st[4], kappa=st[5], kappa_dot=st[6]),
                curvilinear=types.SimpleNamespace(s=st[7], d=st[8], theta=st[9], s_dot=st[10], s_ddot=st[11], d_dot=st[12], d_ddot=st[13]))
            x_0 = types.SimpleNamespace(time_step=int(rng.integers(0, 200)), yaw_rate=float(rng.uniform(-0.5, 0.5)),
                                        orientation=float(st[2][0] + rng.choice([0.0, 0.1, 2 * np.pi, -2 * np.pi, 3.0, -3.0])))
            veh = types.SimpleNamespace(**syn.VEHICLE_2)
            me = types.SimpleNamespace(x_0=x_0, dT=0.1, vehicle_params=veh)
            me.shift_orientation = lambda *a, **kw: pl.Planner.shift_orientation(me, *a, **kw)
            want = pl.Planner._compute_trajectory_pair(me, traj)
            me2 = types.SimpleNamespace(x_0=x_0, dT=0.1, vehicle_params=veh)
            got = ReactivePlannerB200._compute_trajectory_pair(me2, traj)
            ok = len(got[0].state_list) == len(want[0].state_list) and got[0].initial_time_step == want[0].initial_time_step
            for a, b in zip(got[0].state_list, want[0].state_list):
                for f in ("time_step", "orientation", "velocity", "acceleration", "yaw_rate", "steering_angle"):
                    ok = ok and (getattr(a, f) == getattr(b, f))
                    if isinstance(getattr(b, f), float) and getattr(a, f) != getattr(b, f):
                        worst = max(worst, abs(getattr(a, f) - getattr(b, f)))
                ok = ok and np.array_equal(a.position, b.position)
            for a, b in zip(got[1].state_list, want[1].state_list):
                for f in ("time_step", "velocity", "acceleration", "orientation", "yaw_rate"):
                    ok = ok and (getattr(a, f) == getattr(b, f))
                ok = ok and np.array_equal(a.position, b.position)
            ok = ok and np.array_equal(np.array(got[2]), np.array(want[2])) and np.array_equal(np.array(got[3]), np.array(want[3]))
            if not ok:
                bad += 1
                print(f"trajectory pair seed {first + k} row {r}: MISMATCH (worst scalar difference so far {worst:.3e})", flush=True)
    print(f"{n_cases} random cases x 4 trajectories: {'all equal' if not bad else str(bad) + ' MISMATCHES'}")
    return bad


def inactive_cost_sweep(n_cases, first):
    """The five cost terms that are inactive by default but implemented on the device (acceleration, jerk, orientation_offset,
    path_length, distance_to_obstacles: partial_cost_functions.py:24-46,141-151,172-196) -- the reference's functions on
    random samples of random length against the oracle's restatement."""
    import types
    from oracle import frenet_oracle as fo
    bad, worst = 0, 0.0
    for k in range(n_cases):
        rng = np.random.default_rng(first + k)
        Nt = int(rng.integers(6, 64))
        a = rng.normal(0, 2, Nt); v = np.abs(rng.normal(8, 3, Nt)); th = rng.normal(0, 0.3, Nt)
        x = np.cumsum(v) * 0.1 + rng.uniform(-100, 100); y = rng.normal(0, 2, Nt) + rng.uniform(-100, 100)
        traj = types.SimpleNamespace(dt=0.1, cartesian=types.SimpleNamespace(a=a, v=v, x=x, y=y), curvilinear=types.SimpleNamespace(theta=th))
        obs_pos = np.stack([x[rng.integers(0, Nt, 3)] + rng.normal(0, 6, 3), y[rng.integers(0, Nt, 3)] + rng.normal(0, 6, 3)], axis=1)
        scen = types.SimpleNamespace(obstacles=[types.SimpleNamespace(state_at_time=lambda t, p=p: types.SimpleNamespace(position=p)) for p in obs_pos])
        planner = types.SimpleNamespace(x_0=types.SimpleNamespace(time_step=0))
        want = {"acceleration": mg.pcf.acceleration_costs(traj), "jerk": mg.pcf.jerk_costs(traj),
                "orientation_offset": mg.pcf.orientation_offset_costs(traj), "path_length": mg.pcf.path_length_costs(traj),
                "distance_to_obstacles": mg.pcf.distance_to_obstacles_costs(traj, planner=planner, scenario=scen)}
        st = np.zeros((14, Nt))
        st[fo.F_A], st[fo.F_V], st[fo.F_THETA_CL], st[fo.F_X], st[fo.F_Y] = a, v, th, x, y
        prm = fo.Params(N=Nt - 1, obstacle_positions=obs_pos)
        for name, w in want.items():
            got = fo._costs_for(name, st, None, None, prm, [], [], Nt)
            err = abs(float(got) - float(w)) / max(1.0, abs(float(w)))
            worst = max(worst, err)
            if not err < 1e-12:
                bad += 1
                print(f"cost seed {first + k} {name}: MISMATCH {got} vs {w}", flush=True)
    print(f"{n_cases} random samples x 5 cost terms: {'all equal' if not bad else str(bad) + ' MISMATCHES'} (worst {worst:.2e})")
    return bad


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--inactive-costs":
        return 1 if inactive_cost_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 200, int(sys.argv[3]) if len(sys.argv) > 3 else 3000) else 0
    if len(sys.argv) > 1 and sys.argv[1] == "--trajectory-pair":
        return 1 if trajectory_pair_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 50, int(sys.argv[3]) if len(sys.argv) > 3 else 1500) else 0
    if len(sys.argv) > 1 and sys.argv[1] == "--refpath":
        return 1 if refpath_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 50, int(sys.argv[3]) if len(sys.argv) > 3 else 1200) else 0
    if len(sys.argv) > 1 and sys.argv[1] == "--sampling-order":
        return 1 if sampling_order_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 200, int(sys.argv[3]) if len(sys.argv) > 3 else 900) else 0
    if len(sys.argv) > 1 and sys.argv[1] == "--collision-probability":
        return 1 if collision_probability_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 700) else 0
    if len(sys.argv) > 1 and sys.argv[1] == "--initial-states":
        return 1 if initial_state_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 300) else 0
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    tmp = tempfile.mkdtemp(prefix="frx_sweep_")
    mg.HERE = tmp                      # run_case writes ref_<name>.npz there
    helpers.GOLDEN_DIR = tmp           # load_golden reads it back
    bad = 0
    t_all = time.time()
    for k in range(n_cases):
        c = random_case(first + k)
        name = f"sweep{first + k}"
        t0 = time.time()
        try:
            devnull = open(os.devnull, "w")
            old = sys.stdout
            sys.stdout = devnull
            try:
                mg.run_case(name, c["polyline"], c["x_cl"], c["v0"], c["th0"], c["v_des"], c["draw"], c["debug"], c["n_obs"],
                            seed=c["seed"], samp_level=c["level"], samp_max=max(3, c["level"] + 1), cost_weights=c["weights"])
            finally:
                sys.stdout = old
            tog.test_oracle_matches_reference_golden(name)
            g = np.load(os.path.join(tmp, f"ref_{name}.npz"))
            print(f"seed {first + k}: ok   rows {g['sampling'].shape[0]:4d} stored {int(g['stored'].sum()):4d} feasible {int(g['feasible'].sum()):4d} "
                  f"optimal {int(g['optimal_id']):4d} low_vel {bool(g['low_vel_mode'])} draw {c['draw']} debug {c['debug']} obs {c['n_obs']} terms {len(g['cost_names'])} "
                  f"({time.time() - t0:.1f} s)", flush=True)
        except Exception:
            bad += 1
            print(f"seed {first + k}: MISMATCH {c}", flush=True)
            traceback.print_exc()
        finally:
            try:
                os.remove(os.path.join(tmp, f"ref_{name}.npz"))
            except OSError:
                pass
    print(f"{n_cases - bad} of {n_cases} random cases: oracle == reference ({time.time() - t_all:.0f} s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
