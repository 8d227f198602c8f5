"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs."""
import numpy as np
import pytest

from oracle import frenet_oracle as fo
from helpers import GOLDEN_CASES, BAND, load_golden, device_plan, compare_with_oracle, rel_err, band_alternatives

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_device_matches_oracle_on_golden_inputs(name):
    g, ref, prm, preds = load_golden(name)
    ora = fo.plan(g["sampling"], ref, prm, preds)
    dev = device_plan(g["sampling"], ref, prm, preds)
    # no row is exempt: the candidates on the stand-still tie must equal the oracle on one of its two sides
    alts = band_alternatives(g["sampling"], ref, prm, preds, np.flatnonzero(ora["margins"] < BAND))
    errs = compare_with_oracle(dev, ora, prm, alts=alts)
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_lowvel_draw", "short_hv_draw", "scurve_brake_hv_nodraw_debug"])
def test_device_matches_reference_golden_directly(name):
    """Skip the oracle: compare with what the reference's own code produced (tests/golden)."""
    g, ref, prm, preds = load_golden(name)
    dev = device_plan(g["sampling"], ref, prm, preds, check_collisions=False)
    # candidates on a structural tie of the reference (decision margin ~ 1 ulp) are excluded
    ok = fo.plan(g["sampling"], ref, prm, preds, collision_check=False)["margins"] >= BAND
    stored = g["stored"] & ok
    assert np.array_equal(((dev["flags"] & fo.FLAG_STORED) != 0)[ok], g["stored"][ok])
    assert rel_err(dev["states"][:, stored, :], g["states"][:, stored, :]) < 1e-6
    assert np.array_equal(((dev["flags"] & fo.FLAG_FEASIBLE) != 0)[stored], g["feasible"][stored])
    costed = g["costed"] & ok
    assert rel_err(dev["total"][costed], g["total"][costed]) < 1e-6
    assert rel_err(dev["costs"][costed], g["costs"][costed]) < 1e-6
    assert dev["argmin"] == int(g["optimal_id"]) or not ok[dev["argmin"]] or not ok[int(g["optimal_id"])]


def test_fdiv_is_ieee_division():
    """The kernels' slow-path-free division (ddivg = range-checked divisor + ddivf) must equal IEEE division
    bit for bit on its documented domain: ANY divisor, dividend 0 or within 2^+-500 (sign of a zero
    quotient excepted)."""
    from frenetix_motion_planner_b200 import _capi
    h = _capi.Handler(0)
    rng = np.random.default_rng(99)
    n = 2_000_000
    a = rng.normal(0, 1, n) * 10.0 ** rng.integers(-12, 12, n)
    b = rng.normal(0, 1, n) * 10.0 ** rng.integers(-12, 12, n)
    a[:1000] = 0.0; a[1000:2000] = -0.0
    b[2100:2200] = np.array([1e-310, 1e300, np.inf, 0.0, np.nan] * 20)     # degenerate divisors -> fallback
    b[2200:2300] = 0.1; a[2200:2300] = np.linspace(-3, 3, 100)
    b[2300:2400] = 100000.0; a[2300:2400] = np.rint(rng.normal(0, 1e5, 100))
    q1, q2 = h.selftest_fdiv(a, b)
    same = (q1 == q2) | (np.isnan(q1) & np.isnan(q2))
    assert same.all(), f"{(~same).sum()} mismatches, e.g. {a[~same][:3]} / {b[~same][:3]}"
    with np.errstate(all="ignore"):
        ref = a / b
    same_host = (q1 == ref) | (np.isnan(q1) & np.isnan(ref))
    assert same_host.all()


@pytest.mark.gpu
def test_division_by_plan_constants_is_ieee_exact():
    """ddivc(a, b, 1/b) (kernels: yaw rate / dt, rint(.)/100000, sum / Nt) equals IEEE a / b bit for bit."""
    from frenetix_motion_planner_b200 import _capi
    h = _capi.Handler(0)
    rng = np.random.default_rng(7)
    n = 1_000_000
    a = rng.normal(0, 1, n) * 10.0 ** rng.integers(-12, 12, n)
    a[:1000] = np.rint(rng.normal(0, 1e5, 1000))
    a[1000:1100] = 0.0
    for b in (0.1, 0.2, 0.05, 100000.0, 31.0, 51.0, 21.0, 61.0):
        q1, q2 = h.selftest_divc(a, b)
        assert (q1 == q2).all(), f"b={b}: {(q1 != q2).sum()} mismatches"
        assert (q1 == a / b).all()


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_brake_hv_nodraw_nodebug"])
def test_winner_states_record_equals_gathered_row(name):
    """The selected trajectory published with the arg-min (mapped result record) is the gathered state row."""
    g, ref, prm, preds = load_golden(name)
    dev = device_plan(g["sampling"], ref, prm, preds)
    h = dev["handler"]
    assert dev["argmin"] >= 0
    w = h.winner_states()
    ref_rows = h.get_states(np.array([dev["argmin"]], dtype=np.int64))[:, 0, :]
    assert w.shape == ref_rows.shape
    assert np.array_equal(w, ref_rows)
    assert np.array_equal(w, dev["states"][:, dev["argmin"], :])
    fl, tl, tot, costs = h.winner_record()                      # the winner's scalars travel in the same record
    a = dev["argmin"]
    assert (fl, tl, tot) == (int(dev["flags"][a]), int(dev["traj_len"][a]), float(dev["total"][a]))
    assert np.array_equal(costs, dev["costs"][a])


def test_pinned_host_matrix_is_read_in_place():
    """A pinned sampling matrix takes the zero-copy path (kernel prefetches rows over PCIe); same result bit for bit."""
    import torch
    from frenetix_motion_planner_b200 import _capi
    from helpers import configure_handler
    g, ref, prm, preds = load_golden("arc_hv_draw_pred")
    S = np.ascontiguousarray(np.tile(g["sampling"], (7, 1))[:5003])      # ragged last tile, several tiles per warp
    dev = device_plan(S, ref, prm, preds)
    h = _capi.Handler(0)
    configure_handler(h, ref, prm, preds, None, sampling=S)
    Sp = torch.from_numpy(S.copy()).pin_memory()
    res = h.plan(Sp.numpy())
    assert int(res.argmin) == dev["argmin"] and float(res.min_cost) == dev["min_cost"]
    flags, traj_len = h.get_flags()
    costs, total = h.get_costs()
    assert np.array_equal(flags, dev["flags"]) and np.array_equal(traj_len, dev["traj_len"])
    assert np.array_equal(total, dev["total"]) and np.array_equal(costs, dev["costs"])
    assert np.array_equal(h.get_states_range(), dev["states"])
    # an unaligned view into the pinned buffer (row offset, 8-byte alignment only) works as well
    res2 = h.plan(Sp.numpy()[3:])
    d2 = device_plan(S[3:], ref, prm, preds)
    assert int(res2.argmin) == d2["argmin"] and float(res2.min_cost) == d2["min_cost"]


@pytest.mark.parametrize("seg", [1, 2, 4])
@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_lowvel_draw", "scurve_brake_hv_nodraw_nodebug", "tjunction_draw"])
def test_every_lanes_per_candidate_instance_matches_oracle(name, seg, monkeypatch):
    """SEG = 1, 2, 4 lanes per candidate (time steps split into segments) are separate kernel instances; the library picks
    one from the row count, FRX_SEG forces it.  All three must satisfy the parity contract, ragged last tile included."""
    monkeypatch.setenv("FRX_SEG", str(seg))
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(np.tile(g["sampling"], (3, 1))[:g["sampling"].shape[0] * 2 + 7])
    ora = fo.plan(S, ref, prm, preds)
    dev = device_plan(S, ref, prm, preds)
    compare_with_oracle(dev, ora, prm)       # asserts the contract (exact masks / index, 1e-6 on states and costs)


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "arc_hv_nodraw_nodebug", "tjunction_draw", "tjunction_nodraw"])
def test_split_obstacle_kernel_equals_fused_pass(name, monkeypatch):
    """Large plans run the obstacle pass + arg-min as a second kernel (FRX_SPLIT_OBS forces it here): every output is
    bit-identical to the fused pass, and the oracle contract holds."""
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(np.tile(g["sampling"], (4, 1))[:g["sampling"].shape[0] * 3 + 11])
    monkeypatch.setenv("FRX_SEG", "1")
    monkeypatch.setenv("FRX_OBS_CHUNKS", "1")          # one step chunk: same summation order as the fused pass
    monkeypatch.setenv("FRX_SPLIT_OBS", "0")
    fused = device_plan(S, ref, prm, preds)
    monkeypatch.setenv("FRX_SPLIT_OBS", "1")
    split = device_plan(S, ref, prm, preds)
    for k in ("flags", "traj_len", "costs", "total", "states", "reason_counts"):
        assert np.array_equal(fused[k], split[k]), k
    for k in ("argmin", "min_cost", "n_in_list", "n_feasible", "collision_counter"):
        assert fused[k] == split[k], k
    assert fused["res"].n_collide == split["res"].n_collide and fused["res"].n_boundary == split["res"].n_boundary
    assert np.array_equal(fused["handler"].winner_states(), split["handler"].winner_states())
    compare_with_oracle(split, fo.plan(S, ref, prm, preds), prm)


@pytest.mark.gpu
def test_counters_of_a_plan_do_not_leak_into_the_next_one():
    """One handler, three plans: with colliding obstacles, without any obstacle, with the obstacles again.  The device
    keeps Planner._collision_counter in a global counter that the NEXT plan's last block snapshots and clears; a plan that
    runs no collision pass must report 0, not what its predecessor left there (found by the hypothesis suite)."""
    from frenetix_motion_planner_b200 import _capi
    g, ref, prm, preds = load_golden("tjunction_draw")
    S = g["sampling"]
    h = _capi.Handler(0)
    free = fo.plan(S, ref, prm, [])
    w = free["argmin"]
    crowded = [dict(p) for p in preds]
    n = len(crowded[0]["pos_list"])                      # one predicted car drives onto the unobstructed winner from step 18 on
    pos = np.stack([free["states"][fo.F_X][w, :n], free["states"][fo.F_Y][w, :n]], axis=1)
    pos[:18] = pos[18] + np.array([0.0, 300.0])
    crowded[0]["pos_list"] = pos
    ora = fo.plan(S, ref, prm, crowded)
    assert ora["collision_counter"] > 0 and ora["argmin"] >= 0 and ora["argmin"] != w
    n_band = int((ora["margins"] < BAND).sum())         # rows on the stand-still tie may legally fall either way (helpers.py)
    first = device_plan(S, ref, prm, crowded, handler=h)
    assert first["collision_counter"] > 0 and abs(first["collision_counter"] - ora["collision_counter"]) <= n_band
    empty = device_plan(S, ref, prm, [], handler=h)
    assert empty["collision_counter"] == 0 and empty["res"].n_collide == 0 and empty["res"].n_boundary == 0
    again = device_plan(S, ref, prm, crowded, handler=h)
    for k in ("argmin", "min_cost", "n_in_list", "n_feasible", "collision_counter"):
        assert again[k] == first[k], k
    assert np.array_equal(again["flags"], first["flags"]) and np.array_equal(again["reason_counts"], first["reason_counts"])


@pytest.mark.gpu
@pytest.mark.parametrize("n_obs", [7, 37, 70])
def test_obstacle_block_shapes_and_record_staging_are_bit_identical(n_obs, monkeypatch):
    """The split obstacle pass keeps the prediction records of the leading steps in shared memory and runs in one of two
    block shapes (frx_obstacle.cuh), its units dealt to blocks round-robin or to warps by ticket; which shape, which
    dealing, and how many steps are staged (all, some, none -- the rest is read through the L1), must not change a
    single bit.  7 / 37 / 70 obstacles: record lists with every tail length of the
    four-record groups and more obstacles than the 32 lanes of the cooperative cull; the oracle contract holds."""
    from frenetix_motion_planner_b200 import synthetic as syn
    g, ref, prm, _ = load_golden("arc_hv_draw_pred")
    S = np.ascontiguousarray(np.tile(g["sampling"], (4, 1))[:g["sampling"].shape[0] * 3 + 11])
    preds = syn.synthetic_predictions(g["polyline"], n_obs, 31, 0.1, seed=77 + n_obs, lateral_spread=20.0)   # some collide, most do not
    for o, p in enumerate(preds):                      # ragged prediction horizons: per-step lists of different lengths
        keep = 31 - (o % 5) * 4
        for key in ("pos_list", "cov_list", "orientation_list", "v_list"):
            if key in p:
                p[key] = np.asarray(p[key])[:keep]
    walls = np.array([[float(np.mean(ref.ref_x[:40])), float(np.mean(ref.ref_y[:40])) + 6.0, 0.3, 30.0, 0.1]])
    monkeypatch.setenv("FRX_SEG", "1")
    monkeypatch.setenv("FRX_SPLIT_OBS", "1")
    monkeypatch.setenv("FRX_OBS_CHUNKS", "1")
    base = None
    for wide, stage_kb, ticket in (("0", None, "1"), ("0", "0", "0"), ("0", "4", "1"), ("0", None, "0"), ("1", None, "0"),
                                   ("1", "0", "1"), ("1", "9", "0"), ("1", None, "1")):
        monkeypatch.setenv("FRX_OBS_WIDE", wide)
        monkeypatch.setenv("FRX_OBS_TICKET", ticket)      # units dealt to warps by ticket / to blocks round-robin
        if stage_kb is None:
            monkeypatch.delenv("FRX_OBS_STAGE_KB", raising=False)
        else:
            monkeypatch.setenv("FRX_OBS_STAGE_KB", stage_kb)
        dev = device_plan(S, ref, prm, preds, static_obbs=walls)
        assert dev["res"].obstacle_kernel_ms > 0
        if base is None:
            base = dev
            ora = fo.plan(S, ref, prm, preds, static_obbs=walls)
            compare_with_oracle(dev, ora, prm, alts=band_alternatives(S, ref, prm, preds, np.flatnonzero(ora["margins"] < BAND), static_obbs=walls))
            continue
        for k in ("flags", "traj_len", "costs", "total", "states", "reason_counts"):
            assert np.array_equal(base[k], dev[k]), (k, wide, stage_kb, ticket)
        for k in ("argmin", "min_cost", "n_in_list", "n_feasible", "collision_counter"):
            assert base[k] == dev[k], (k, wide, stage_kb, ticket)
    monkeypatch.setenv("FRX_SPLIT_OBS", "0")          # and the fused pass of the eval kernel: same operations, same order
    monkeypatch.delenv("FRX_OBS_WIDE", raising=False)
    monkeypatch.delenv("FRX_OBS_TICKET", raising=False)
    monkeypatch.delenv("FRX_OBS_STAGE_KB", raising=False)
    fused = device_plan(S, ref, prm, preds, static_obbs=walls)
    for k in ("flags", "traj_len", "costs", "total", "states", "reason_counts"):
        assert np.array_equal(base[k], fused[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [2, 4, 8])
@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "tjunction_draw", "tjunction_nodraw"])
def test_step_chunked_obstacle_pass(name, chunks, monkeypatch):
    """Mid-size plans cut the obstacle pass into chunks of time steps (frx_obstacle.cuh): masks, first-hit indices, counters
    and the selected row equal the one-chunk pass exactly; the prediction cost is the same sum in a different association
    (<= 1e-12 relative); the oracle contract holds."""
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(np.tile(g["sampling"], (4, 1))[:g["sampling"].shape[0] * 3 + 11])
    walls = np.array([[float(np.mean(ref.ref_x[:40])), float(np.mean(ref.ref_y[:40])) + 6.0, 0.3, 30.0, 0.1]])
    monkeypatch.setenv("FRX_SEG", "1")
    monkeypatch.setenv("FRX_SPLIT_OBS", "1")
    monkeypatch.setenv("FRX_OBS_CHUNKS", "1")
    one = device_plan(S, ref, prm, preds, static_obbs=walls)
    monkeypatch.setenv("FRX_OBS_CHUNKS", str(chunks))
    cut = device_plan(S, ref, prm, preds, static_obbs=walls)
    assert cut["res"].obstacle_kernel_ms > 0
    for k in ("flags", "traj_len", "states", "reason_counts"):
        assert np.array_equal(one[k], cut[k]), k
    assert rel_err(cut["costs"], one["costs"]) < 1e-12 and rel_err(cut["total"], one["total"]) < 1e-12
    for k in ("argmin", "n_in_list", "n_feasible", "collision_counter"):
        assert one[k] == cut[k], k
    assert one["res"].n_collide == cut["res"].n_collide and one["res"].n_boundary == cut["res"].n_boundary
    compare_with_oracle(cut, fo.plan(S, ref, prm, preds, static_obbs=walls), prm)


@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "scurve_lowvel_draw", "tjunction_nodraw", "scurve_brake_hv_nodraw_debug"])
def test_cpp_flavour_switches_match_the_oracle(name):
    """the use_cpp = True variants on the hot path -- curvature-rate limit from v_delta_max
    (reactive_planner_cpp.py:109-112) and velocity_offset with norm_order = 2 (:170-178) -- device vs oracle, and they
    really change the outcome."""
    import dataclasses
    g, ref, prm, preds = load_golden(name)
    p2 = dataclasses.replace(prm, curvature_rate_from_v_delta=True, v_delta_max=0.4, velocity_offset_norm=2)
    ora = fo.plan(g["sampling"], ref, p2, preds)
    dev = device_plan(g["sampling"], ref, p2, preds)
    compare_with_oracle(dev, ora, p2)
    base = fo.plan(g["sampling"], ref, prm, preds)
    assert not np.array_equal(base["flags"] & fo.FLAG_FEASIBLE, ora["flags"] & fo.FLAG_FEASIBLE)
    k = list(ora["cost_names"]).index("velocity_offset")
    assert not np.allclose(base["costs"][:, k], ora["costs"][:, k])


@pytest.mark.parametrize("seg", [1, 2, 4])
@pytest.mark.parametrize("name", ["tjunction_draw", "arc_hv_draw_pred"])
def test_static_boxes_fused_pass_tile_cull_equals_oracle(name, seg, monkeypatch):
    """Road-boundary boxes in the FUSED pass (small plans, every SEG instance): the per-tile cull list must select what
    testing every box selects -- against the oracle, which tests every box, with the T-junction network's 82 walls, with more
    boxes than one 32-lane round, with a box list longer than the cull list (falls back to the linear scan) and with walls
    far away (empty list)."""
    import json
    import os
    from helpers import GOLDEN_DIR
    from frenetix_motion_planner_b200.road_boundary import road_boundary_obbs
    monkeypatch.setenv("FRX_SEG", str(seg))
    monkeypatch.setenv("FRX_SPLIT_OBS", "0")
    g, ref, prm, preds = load_golden(name)
    S = g["sampling"]
    raw = json.load(open(os.path.join(GOLDEN_DIR, "tjunction_lanelets.json")))
    net = {int(k): dict(left=np.array(v["left"]), right=np.array(v["right"]), adj_left=v["adj_left"], adj_right=v["adj_right"])
           for k, v in raw.items()}
    walls = road_boundary_obbs(net)
    if not name.startswith("tjunction"):          # move the network onto this case's reference path
        walls = walls.copy()
        walls[:, 0] += ref.ref_x[30] - walls[:, 0].mean(); walls[:, 1] += ref.ref_y[30] - walls[:, 1].mean()
    rng = np.random.default_rng(3)
    far = np.column_stack([rng.uniform(4000, 5000, 40), rng.uniform(4000, 5000, 40), rng.uniform(-3, 3, 40), rng.uniform(1, 9, 40),
                           rng.uniform(0.05, 1, 40)])
    many = np.vstack([walls] * 7)[:600]                                          # > FRX_WALL_LIST boxes
    for boxes in (walls, np.vstack([far, walls[:45]]), far, many):
        ora = fo.plan(S, ref, prm, preds, static_obbs=boxes)
        dev = device_plan(S, ref, prm, preds, static_obbs=boxes)
        compare_with_oracle(dev, ora, prm)
    assert (fo.plan(S, ref, prm, preds, static_obbs=walls)["flags"] & fo.FLAG_BOUNDARY).any()


@pytest.mark.parametrize("correlated", [False, True])
@pytest.mark.parametrize("name", ["arc_hv_draw_pred", "tjunction_draw", "scurve_lowvel_nodraw"])
def test_collision_probability_prediction_cost_matches_the_oracle(name, correlated, monkeypatch):
    """prediction term = CalculateCollisionProbabilityFast (reactive_planner_cpp.py:151-155 = the in-tree
    get_collision_probability_fast): bivariate-normal rectangle probabilities on the device (Genz's BVND, all three |rho|
    branches) against the oracle restatement that is pinned to the reference's own function; small plans (forced into the
    split obstacle kernel) and step-chunked ones."""
    import dataclasses
    g, ref, prm, preds = load_golden(name)
    preds = [dict(p) for p in preds]
    S = g["sampling"]
    if correlated:
        ego = np.array([ref.ref_x[int(np.argmax(ref.ref_pos > S[0, 2]))], ref.ref_y[int(np.argmax(ref.ref_pos > S[0, 2]))]])
        for o, p in enumerate(preds):
            cov = np.array(p["cov_list"], dtype=float).copy()
            for k in range(cov.shape[0]):
                r = [0.5, -0.8, 0.95, -0.97, 0.2, 0.0][(k + o) % 6]
                sx, sy = 0.4 + 0.03 * k, 0.6 + 0.02 * k
                cov[k] = [[sx * sx, r * sx * sy], [r * sx * sy, sy * sy]]
            cov[2] = 0.0                                                      # ground-truth style zero covariance
            p["cov_list"] = cov
            p["pos_list"] = np.array(p["pos_list"]) + (ego - np.array(p["pos_list"])[3]) * 0.8     # pull them next to the ego
    p1 = dataclasses.replace(prm, prediction_cost_mode=1)
    ora = fo.plan(S, ref, p1, preds)
    k = list(ora["cost_names"]).index("prediction")
    costed = (ora["flags"] & fo.FLAG_COSTED) != 0
    if correlated or name != "scurve_lowvel_nodraw":                          # (that case's obstacles start > 5 m away: all-zero term)
        assert (ora["costs"][costed, k] > 0).sum() > 10                       # the term is alive in this case
    for chunks in ("1", "4"):
        monkeypatch.setenv("FRX_OBS_CHUNKS", chunks)
        dev = device_plan(S, ref, p1, preds)
        assert dev["res"].obstacle_kernel_ms > 0
        alts = band_alternatives(S, ref, p1, preds, np.flatnonzero(ora["margins"] < BAND))
        compare_with_oracle(dev, ora, p1, alts=alts)
    if not correlated:          # (the correlated variant carries a zero covariance, which only the probability cost accepts)
        base = fo.plan(S, ref, prm, preds)
        assert not np.allclose(base["costs"][costed, k], ora["costs"][costed, k])


def test_prediction_cost_exact_fallback(monkeypatch):
    """The fast prediction-cost path factors the inverse covariance (Cholesky) and shares reciprocals; covariances without a
    factor (indefinite) and an ego that sits exactly ON a predicted mean (the reference returns inf) must take the exact
    term-by-term fallback -- fused pass, split pass and chunked pass."""
    g, ref, prm, preds = load_golden("arc_hv_draw_pred")
    S = g["sampling"]
    ora0 = fo.plan(S, ref, prm, preds)
    preds = [dict(p) for p in preds]
    cov = np.array(preds[0]["cov_list"], dtype=float).copy()
    cov[4] = [[1.0, 2.0], [2.0, 1.0]]                      # indefinite
    cov[9] = [[0.3, 0.1], [0.25, 0.4]]                     # not symmetric (the quadratic form only sees the symmetric part)
    preds[0]["cov_list"] = cov
    pos = np.array(preds[1]["pos_list"], dtype=float).copy()
    # a costed row whose position at step 7 is the same double on both sides (states agree to an ulp or two, most of them
    # exactly): "the ego ON the mean" must mean the same thing to the device and to the oracle
    dev0 = device_plan(S, ref, prm, preds)
    same = (dev0["states"][fo.F_X][:, 7] == ora0["states"][fo.F_X][:, 7]) & (dev0["states"][fo.F_Y][:, 7] == ora0["states"][fo.F_Y][:, 7])
    rows = np.flatnonzero(((ora0["flags"] & fo.FLAG_COSTED) != 0) & same)
    assert rows.size > 20
    r = int(rows[17])
    pos[6] = [ora0["states"][fo.F_X][r, 7], ora0["states"][fo.F_Y][r, 7]]       # obstacle mean == ego position of row r at step 7
    preds[1]["pos_list"] = pos
    ora = fo.plan(S, ref, prm, preds)
    k = list(ora["cost_names"]).index("prediction")
    assert np.isinf(ora["costs"][r, k])
    for split, chunks in (("0", "1"), ("1", "1"), ("1", "4")):
        monkeypatch.setenv("FRX_SPLIT_OBS", split)
        monkeypatch.setenv("FRX_OBS_CHUNKS", chunks)
        dev = device_plan(S, ref, prm, preds)
        assert np.isinf(dev["costs"][r, k]) and np.isinf(dev["total"][r])
        alts = band_alternatives(S, ref, prm, preds, np.flatnonzero(ora["margins"] < BAND))
        compare_with_oracle(dev, ora, prm, alts=alts)
