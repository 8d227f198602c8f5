"""The drop-in claim, tested against the reference's OWN callers.

CPU part (needs the reference tree at /root/reference, i.e. the build container; skipped elsewhere): the unmodified
``Planner`` / ``ReactivePlannerPython`` / ``FrenetPlannerInterface`` are imported under tests/golden/ref_stubs.py and
  (i)   their public surface is diffed against ``ReactivePlannerB200`` -- any missing name fails;
  (ii)  ``FrenetPlannerInterface.update_planner`` / ``step_interface`` (frenet_interface.py:178-287) drive the B200 planner
        through three replanning cycles of the ZAM_Tjunction fixture; the run must reproduce the committed trace
        (tests/golden/interface_trace.npz, made by tests/golden/make_interface_trace.py);
  (iii) the multi-agent protocol of INTEGRATION.md (update_planner for all agents -> prefetch_plans -> step_interface for
        all agents) gives every agent exactly what stepping them one after the other gives.
The device is stood in for by the oracle-backed handler (tests/oracle_handler.py) there.

GPU part (no reference tree on the GPU box): the planner calls recorded in the trace are replayed on the real device and
must give the same selected trajectories, costs and counters; the batched launch is checked against the ORACLE per agent.
"""
import inspect
import json
import os
import re
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
sys.path.insert(0, GOLDEN)

from helpers import rel_err  # noqa: E402
from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402

HAVE_REFERENCE = os.path.isdir("/root/reference/frenetix_motion_planner")
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is only mounted in the build container")


def _trace():
    return np.load(os.path.join(GOLDEN, "interface_trace.npz"))


def _reference_classes():
    import ref_stubs
    ref_stubs.install()
    from frenetix_motion_planner.planner import Planner
    from frenetix_motion_planner.reactive_planner import ReactivePlannerPython
    from cr_scenario_handler.planner_interfaces.frenet_interface import FrenetPlannerInterface
    return Planner, ReactivePlannerPython, FrenetPlannerInterface


# ---------------------------------------------------------------------------------------------------------------
# (i) surface
# ---------------------------------------------------------------------------------------------------------------
@needs_reference
def test_public_surface_of_the_reference_planner_classes_is_present():
    Planner, ReactivePlannerPython, _ = _reference_classes()
    from frenetix_motion_planner_b200 import ReactivePlannerB200

    def members(cls):
        return {n for n, v in inspect.getmembers(cls) if not n.startswith("__") and (inspect.isfunction(v) or isinstance(v, property))}
    ours = members(ReactivePlannerB200)
    missing = (members(Planner) | members(ReactivePlannerPython)) - ours
    # the end-point sampler is dead code in the reference (never called, reactive_planner.py:628-673)
    assert missing <= {"_create_end_point_trajectory_bundle"}, f"methods / properties missing: {sorted(missing)}"
    # same signatures for what callers invoke with keywords (frenet_interface.py:129,209,263-267; agent.py)
    for name in ("update_externals", "set_desired_velocity", "plan_postprocessing", "convert_state_list_to_commonroad_object",
                 "set_ego_vehicle_state", "record_state_and_input", "set_sampling_parameters", "set_reference_and_coordinate_system"):
        want = [p for p in inspect.signature(getattr(ReactivePlannerPython, name)).parameters]
        got = [p for p in inspect.signature(getattr(ReactivePlannerB200, name)).parameters]
        assert got[:len(want)] == want, (name, want, got)
    # the 7-argument constructor of frenet_interface.py:71-73
    want = list(inspect.signature(ReactivePlannerPython.__init__).parameters)
    assert list(inspect.signature(ReactivePlannerB200.__init__).parameters)[:len(want)] == want
    # instance attributes the reference constructors create
    src = inspect.getsource(Planner.__init__) + inspect.getsource(ReactivePlannerPython.__init__)
    attrs = set(re.findall(r"self\\.([A-Za-z_][A-Za-z_0-9]*)\\s*(?::[^=\\n]+)?=[^=]", src))
    from oracle_handler import OracleHandler
    sys.path.insert(0, GOLDEN)
    import make_interface_trace as mit
    cfg_plan, cfg_sim = mit.configs()
    p = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None, handler=OracleHandler())
    absent = {a for a in attrs if not hasattr(p, a)}
    assert not absent, f"instance attributes missing: {sorted(absent)}"


@needs_reference
def test_stopping_trajectory_selection_equals_the_reference_static_method():
    """reactive_planner_cpp.py:446-469 (emergency_mode == "stopping"), on random feasible subsets of a cpp-style grid."""
    import ref_stubs
    ref_stubs.install()
    from frenetix_motion_planner.reactive_planner_cpp import ReactivePlannerCpp
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    rng = np.random.default_rng(4)
    t1, v1, d1 = np.round(np.arange(11, 31, 3) * 0.1, 2), np.linspace(0.001, 9.0, 10), np.append(np.linspace(-3, 3, 9), 0.37)
    S = syn.grid_sampling_matrix(t1, v1, d1, ([5.0, 4.0, 0.0], [0.37, 0.0, 0.0]))
    for _ in range(200):
        keep = np.flatnonzero(rng.random(S.shape[0]) < rng.uniform(0.01, 0.5))
        if keep.size == 0:
            continue
        rng.shuffle(keep)
        trajs = [types.SimpleNamespace(sampling_parameters=S[r], row=int(r)) for r in keep]
        d_pos = float(rng.uniform(-3, 3))
        want = ReactivePlannerCpp._select_stopping_trajectory(trajs, S, d_pos)
        got = ReactivePlannerB200._select_stopping_trajectory(trajs, S, d_pos)
        assert got.row == want.row


# ---------------------------------------------------------------------------------------------------------------
# (ii) FrenetPlannerInterface drives the planner
# ---------------------------------------------------------------------------------------------------------------
@needs_reference
def test_reference_interface_drives_the_planner_for_three_replanning_cycles():
    _, _, FrenetPlannerInterface = _reference_classes()
    from oracle_handler import OracleHandler
    import make_interface_trace as mit
    out, planner, it = mit.run(FrenetPlannerInterface, OracleHandler)
    g = _trace()
    assert set(out.keys()) == set(g.files)
    for k in g.files:
        assert np.allclose(np.asarray(out[k], dtype=float), np.asarray(g[k], dtype=float), rtol=1e-9, atol=1e-12), k
    # what the interface and the simulation read off the planner afterwards (frenet_interface.py:149-176)
    assert len(it.record_state_list) == 10 and len(it.record_input_list) == 10 and len(it.vehicle_history) == 10
    assert it.optimal_trajectory is planner.optimal_trajectory and it.trajectory_pair is planner.trajectory_pair
    assert len(it.all_trajectories) == int(g["s6_n_all_traj"]) and it.coordinate_system is planner.coordinate_system
    plans = [0, 3, 6]
    assert [int(g[f"s{t}_counter"]) for t in range(9)] == [0, 1, 2] * 3 and all(f"s{t}_opt_id" in g.files for t in plans)
    # a selected trajectory that was kept from an earlier cycle still shows ITS numbers (stale views)
    first = planner.ego_vehicle_history[1]
    assert first.initial_state.time_step == 0
    kept = out["s0_opt_states"]
    assert np.array_equal(kept, g["s0_opt_states"])


@needs_reference
def test_reference_sql_logger_writes_the_full_trajectory_set(tmp_path):
    """SURVEY 8f-4, logging sinks: the reference's own SqlLogger.log_all_trajectories (logging_helpers.py:275-294) walks
    planner.all_traj and reads every attribute of the sample surface (cartesian / curvilinear arrays, costMap,
    feasabilityMap, sampling_parameters, _ego_risk, _coll_detected, boundary_harm ...).  It runs unmodified on the B200
    planner's lazy samples, with one device gather per 256 trajectories instead of one per trajectory."""
    from pathlib import Path
    _, _, FrenetPlannerInterface = _reference_classes()
    from frenetix_motion_planner.utility.logging_helpers import SqlLogger
    from oracle_handler import OracleHandler
    import make_interface_trace as mit
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    fx, lanelets = mit.fixture()
    cfg_plan, cfg_sim = mit.configs()

    class Counting(OracleHandler):
        gathers = 0

        def get_states(self, idx, fields=None):
            Counting.gathers += 1
            return super().get_states(idx, fields)
    planner = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None, handler=Counting())
    it = mit.make_interface(FrenetPlannerInterface, planner, fx, lanelets, cfg_plan, cfg_sim)
    it.update_planner(types.SimpleNamespace(lanelet_network=None), mit.predictions_at(fx, 0))
    it.step_interface(0)
    sql = SqlLogger(Path(tmp_path), cfg_plan, cfg_sim, None, None)
    sql.set_cost_names(list(planner.cost_names))
    n = len(planner.all_traj)
    before = Counting.gathers
    sql.log_all_trajectories(planner.all_traj, 0)
    assert Counting.gathers - before <= -(-n // 256) + 1
    con = sql.con                                   # (the logger holds the database in EXCLUSIVE locking mode)
    assert con.execute("SELECT COUNT(*) FROM trajectories").fetchone()[0] == n == 630
    b = planner._bundle
    row = planner.all_traj[7].uniqueId
    x_txt, = con.execute("SELECT x FROM trajectories WHERE id = ?", (str(row),)).fetchone()
    assert np.allclose(np.array(json.loads(x_txt)), b._h.get_states(np.array([row]))[0, 0], rtol=1e-4)
    cost, = con.execute("SELECT costs_cumulative_weighted FROM costs WHERE id = ?", (row,)).fetchone()
    assert cost == b.total[row]
    feas = con.execute("SELECT feasible, inf_curvature FROM infeasability WHERE id = ?", (row,)).fetchone()
    assert bool(feas[0]) == bool(b.flags[row] & 2)
    t1, d1 = con.execute("SELECT t1, d1 FROM sampling_params WHERE id = ?", (row,)).fetchone()
    assert (t1, d1) == (b.sampling_row(row)[1], b.sampling_row(row)[10])
    harm = [r[0] for r in con.execute("SELECT boundary_harm FROM trajectories_meta")]
    assert sum(h > 0 for h in harm) == int(((b.flags & (1 << 13)) != 0).sum()) > 0


@needs_reference
def test_stale_bundle_views_raise_instead_of_showing_the_next_plan():
    _, _, FrenetPlannerInterface = _reference_classes()
    from oracle_handler import OracleHandler
    from frenetix_motion_planner_b200.trajectories import StaleBundleError
    import make_interface_trace as mit
    fx, lanelets = mit.fixture()
    cfg_plan, cfg_sim = mit.configs()
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    planner = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None, handler=OracleHandler())
    it = mit.make_interface(FrenetPlannerInterface, planner, fx, lanelets, cfg_plan, cfg_sim)
    it.update_planner(types.SimpleNamespace(lanelet_network=None), mit.predictions_at(fx, 0))
    it.step_interface(0)
    opt0, all0 = planner.optimal_trajectory, planner.all_traj
    x0 = np.array(opt0.cartesian.x)
    cost0, other = opt0.cost, all0[5]
    other_cost = other.cost                               # flags / costs of all_traj were read by sort(): they survive
    for t in (1, 2, 3):
        it.update_planner(types.SimpleNamespace(lanelet_network=None), mit.predictions_at(fx, t))
        it.step_interface(t)
    assert planner.optimal_trajectory is not opt0
    assert np.array_equal(opt0.cartesian.x, x0) and opt0.cost == cost0          # detached: still the first plan's numbers
    assert other.cost == other_cost
    with pytest.raises(StaleBundleError):
        other.cartesian                                                          # never fetched -> recycled by the next plan


# ---------------------------------------------------------------------------------------------------------------
# (iii) multi-agent protocol
# ---------------------------------------------------------------------------------------------------------------
def _agents(iface_cls, handler_factory, n=3):
    import make_interface_trace as mit
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    fx, lanelets = mit.fixture()
    agents = []
    for a in range(n):
        cfg_plan, cfg_sim = mit.configs()
        planner = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None, handler=handler_factory())
        planner.obstacle_order = [int(i) for i in fx["obstacle_ids"]]
        fxa = dict(fx)
        fxa["ego_velocity"] = np.array(float(fx["ego_velocity"]) + 0.7 * a)       # agents differ in their initial speed
        it = mit.make_interface(iface_cls, planner, fxa, lanelets, cfg_plan, cfg_sim)
        agents.append(it)
    return agents, fx


@needs_reference
def test_agent_batch_protocol_prepare_one_launch_finish_equals_sequential_stepping():
    _, _, FrenetPlannerInterface = _reference_classes()
    import oracle_handler
    import make_interface_trace as mit
    from frenetix_motion_planner_b200 import _capi
    from frenetix_motion_planner_b200.reactive_planner_b200 import prefetch_plans
    saved = _capi.plan_batched
    oracle_handler.install_batched()
    try:
        seq, fx = _agents(FrenetPlannerInterface, oracle_handler.OracleHandler)
        bat, _ = _agents(FrenetPlannerInterface, oracle_handler.OracleHandler)
        scenario = types.SimpleNamespace(lanelet_network=None)
        launches = []
        for t in range(6):
            preds = mit.predictions_at(fx, t)
            for it in seq:                                   # agent_batch.py:186-189 as shipped
                it.update_planner(scenario, preds)
                it.step_interface(t)
            # INTEGRATION.md patch: prepare -> one launch -> finish
            replanning = [it for it in bat if it.replanning_counter == 0 or int(it.replanning_counter / mit.REPLAN) == 1]
            for it in replanning:
                it.update_planner(scenario, preds)
            gens = [it.planner.handler.generation for it in bat]
            if replanning:
                prefetch_plans([it.planner for it in replanning])
            for it in bat:
                it.update_planner(scenario, preds)
                it.step_interface(t)
            launches.append([it.planner.handler.generation - g0 for it, g0 in zip(bat, gens)])
            for a, b in zip(seq, bat):
                assert a.planner.optimal_trajectory.uniqueId == b.planner.optimal_trajectory.uniqueId
                assert a.planner.optimal_trajectory.cost == b.planner.optimal_trajectory.cost
                assert np.array_equal(a.x_0.position, b.x_0.position) and a.x_cl == b.x_cl
                assert a.planner._infeasible_count_kinematics == b.planner._infeasible_count_kinematics
        # replanning steps cost exactly ONE plan per agent (the batched one), the steps in between none
        assert launches == [[1, 1, 1], [0, 0, 0], [0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0]]
    finally:
        _capi.plan_batched = saved


# ---------------------------------------------------------------------------------------------------------------
# GPU: replay of the recorded planner calls on the device
# ---------------------------------------------------------------------------------------------------------------
def _device_planner():
    import make_interface_trace as mit
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    fx, lanelets = mit.fixture()
    cfg_plan, cfg_sim = mit.configs()
    p = ReactivePlannerB200(cfg_plan, cfg_sim, None, None, None, None, None)
    p.obstacle_order = [int(i) for i in fx["obstacle_ids"]]
    p.set_road_boundary(lanelets)
    return p, fx, mit


def _state_from(row):
    from frenetix_motion_planner_b200.reactive_planner_b200 import PlannerState
    return PlannerState(position=np.array(row[:2]), orientation=float(row[2]), velocity=float(row[3]), acceleration=float(row[4]),
                        yaw_rate=float(row[5]), steering_angle=float(row[6]), time_step=int(row[7]))


@pytest.mark.gpu
def test_recorded_interface_run_replayed_on_the_device():
    g = _trace()
    p, fx, mit = _device_planner()
    p.update_externals(x_0=_state_from(g["s0_x0"]), reference_path=fx["reference_path"])
    assert np.allclose(np.array(p.x_cl[0] + p.x_cl[1]), g["x_cl0"], rtol=1e-12)
    for t in (0, 3, 6):
        xcl = g[f"s{t}_plan_xcl_in"]
        p.update_externals(scenario=types.SimpleNamespace(lanelet_network=None), x_0=_state_from(g[f"s{t}_x0"]),
                           x_cl=(list(xcl[:3]), list(xcl[3:])), desired_velocity=mit.DESIRED_VELOCITY,
                           predictions=mit.predictions_at(fx, t))
        pair = p.plan()
        opt = p.optimal_trajectory
        assert pair is not None and opt.uniqueId == int(g[f"s{t}_opt_id"])
        assert abs(opt.cost - float(g[f"s{t}_opt_cost"])) <= 1e-6 * max(1.0, abs(float(g[f"s{t}_opt_cost"])))
        st = np.stack([getattr(opt.cartesian, f) for f in ("x", "y", "theta", "v", "a", "kappa", "kappa_dot")] +
                      [getattr(opt.curvilinear, f) for f in ("s", "d", "theta", "s_dot", "s_ddot", "d_dot", "d_ddot")])
        assert rel_err(st, g[f"s{t}_opt_states"]) < 1e-6
        assert list(p._infeasible_count_kinematics) == [int(v) for v in g[f"s{t}_counts"]]
        assert abs(p.infeasible_kinematics_percentage - float(g[f"s{t}_percentage"])) < 1e-9
        assert p.infeasible_count_collision == int(g[f"s{t}_collisions"])
        assert [p.last_plan_stats.n_candidates, p.last_plan_stats.n_collide, p.last_plan_stats.n_boundary] == g[f"s{t}_stats"].tolist()
        assert len(p.all_traj) == int(g[f"s{t}_n_all_traj"])
        # the next cycle starts from the state the interface would take over (frenet_interface.py:252-253)
        nxt = pair[0].state_list[1]
        assert np.allclose([nxt.position[0], nxt.position[1], nxt.orientation, nxt.velocity], g[f"s{t}_x0_after"], rtol=1e-9)
        assert np.allclose(pair[2][1] + pair[3][1], g[f"s{t}_xcl_after"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_batched_launch_against_the_oracle_per_agent():
    """every agent of ONE batched launch against the oracle on its own inputs (not against solo CUDA)."""
    from oracle import frenet_oracle as fo
    from helpers import load_golden, compare_with_oracle, band_alternatives, BAND, configure_handler
    from frenetix_motion_planner_b200 import _capi
    cases = ["arc_hv_draw_pred", "straight_hv_draw", "scurve_lowvel_draw", "short_hv_draw", "scurve_brake_hv_draw", "tjunction_nodraw"]
    handlers, mats, inputs = [], [], []
    for name in cases:
        g, ref, prm, preds = load_golden(name)
        h = _capi.Handler(0)
        configure_handler(h, ref, prm, preds, None, sampling=g["sampling"])
        handlers.append(h); mats.append(np.ascontiguousarray(g["sampling"])); inputs.append((g, ref, prm, preds))
    results = _capi.plan_batched(handlers, mats)
    for h, S, res, (g, ref, prm, preds) in zip(handlers, mats, results, inputs):
        flags, traj_len = h.get_flags()
        costs, total = h.get_costs()
        dev = dict(res=res, flags=flags, traj_len=traj_len, costs=costs, total=total, states=h.get_states_range(),
                   argmin=int(res.argmin), min_cost=float(res.min_cost),
                   reason_counts=np.array(list(res.reason_counts), dtype=np.int64), n_in_list=int(res.n_in_list),
                   n_feasible=int(res.n_feasible), collision_counter=int(res.collision_counter))
        ora = fo.plan(S, ref, prm, preds)
        alts = band_alternatives(S, ref, prm, preds, np.flatnonzero(ora["margins"] < BAND))
        compare_with_oracle(dev, ora, prm, alts=alts)
