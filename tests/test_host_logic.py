"""CPU-only tests of the host layer: level sets / sampling matrix against the reference's own outputs,
time tables, packing, sharding, the C-ABI surface, and loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN_DIR, load_golden
from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import hotpath, synthetic as syn
from frenetix_motion_planner_b200.sampling_matrix import (SamplingHandler, generate_sampling_matrix, python_path_rows,
                                                          sampling_axes)
from frenetix_motion_planner_b200.dist import shard_rows, reduce_winners

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level_sets_and_matrix_match_reference_outputs():
    g = np.load(os.path.join(GOLDEN_DIR, "ref_sampling.npz"))
    sh = SamplingHandler(dt=0.1, max_sampling_number=3, t_min=1.1, horizon=3.0, delta_d_max=3, delta_d_min=-3,
                         d_ego_pos=False)
    sh.set_v_sampling(0.001, 13.75)
    for lvl in range(3):
        assert np.array_equal(np.array(sorted(sh.t_sampling.to_range(lvl))), g[f"t_{lvl}"])
        assert np.array_equal(np.array(sorted(sh.v_sampling.to_range(lvl))), g[f"v_{lvl}"])
        assert np.array_equal(np.array(sorted(sh.d_sampling.to_range(lvl))), g[f"d_{lvl}"])
    M = generate_sampling_matrix(t0_range=0.0, t1_range=g["m_t1"], s0_range=10.0, ss0_range=8.0, sss0_range=0.5,
                                 ss1_range=g["m_v1"], sss1_range=0, d0_range=0.2, dd0_range=0.1, ddd0_range=-0.1,
                                 d1_range=g["m_d1"], dd1_range=0.0, ddd1_range=0.0)
    assert np.array_equal(M, g["matrix"])
    x_cl = ([10.0, 8.0, 0.5], [0.2, 0.1, -0.1])
    assert np.array_equal(syn.grid_sampling_matrix(g["m_t1"], g["m_v1"], g["m_d1"], x_cl), g["matrix"])


@pytest.mark.parametrize("name,v0", [("straight_hv_draw", 8.0), ("arc_hv_draw_pred", 9.5), ("scurve_lowvel_draw", 1.2)])
def test_python_path_row_order_equals_reference_generation_order(name, v0):
    """Row index == the reference's uniqueId: same sets, same iteration order (reactive_planner.py:149-175)."""
    g, ref, prm, preds = load_golden(name)
    x_cl = (list(g["x_cl_lon"]), list(g["x_cl_lat"]))
    sh = SamplingHandler(dt=0.1, max_sampling_number=3, t_min=1.1, horizon=3.0, delta_d_max=3, delta_d_min=-3,
                         d_ego_pos=False)
    sh.set_v_sampling(*syn.velocity_interval(v0, syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"]))
    rows = python_path_rows(*sampling_axes(sh, 2, x_cl), x_cl)
    assert np.array_equal(rows, g["sampling"])
    t, v, d = sampling_axes(sh, 2, x_cl, cpp_style=True)
    assert len(t) * len(v) * len(d) == 800        # the C++ path's 8 x 10 x 10


def test_time_tables_follow_numpy_semantics():
    for T in (1.1, 1.4, 1.7, 2.0, 2.3, 2.6, 2.9, 3.0):
        n, tp = hotpath.time_table(T, 0.1, 31)
        t, t2, t3, t4, t5 = fo.time_grid(T, 0.1)
        assert n == len(t)
        for k, a in enumerate((t, t2, t3, t4, t5)):
            assert np.array_equal(tp[k, :n], a) and not tp[k, n:].any()
    assert hotpath.time_table(1.1, 0.1, 31)[0] == 13        # np.arange length quirk
    with pytest.raises(ValueError):
        hotpath.time_table(3.2, 0.1, 31)


def test_distinct_durations_and_packing():
    S = syn.grid_sampling_matrix([1.1, 2.0, 3.0], [1.0, 2.0], [0.0, 0.5, 1.0], ([0, 1, 0], [0, 0, 0]))
    assert np.array_equal(hotpath.distinct_durations(S), [1.1, 2.0, 3.0])
    assert np.array_equal(hotpath.distinct_durations(S[np.random.default_rng(0).permutation(len(S))]), [1.1, 2.0, 3.0])
    preds = syn.synthetic_predictions(syn.straight_polyline(200), 3, 31, 0.1, seed=1)
    preds[1]["pos_list"] = preds[1]["pos_list"][:10]; preds[1]["cov_list"] = preds[1]["cov_list"][:10]
    pos, cov, th, hl, hw, ln = hotpath.pack_predictions({7: preds[0], 3: preds[1], 9: preds[2]}, obstacle_order=[3, 9, 7])
    assert list(ln) == [10, 31, 31] and pos.shape == (3, 31, 2) and np.array_equal(pos[2], preds[0]["pos_list"])
    assert np.array_equal(cov[0, 10:], np.tile(np.eye(2), (21, 1, 1))) and hl[0] == 2.75
    assert hotpath.pack_predictions({}) is None
    names, w = hotpath.active_costs({"b": 1.0, "a": 0.5, "zero": 0.0})
    assert names == ["a", "b"] and w == [0.5, 1.0]


def test_shards_cover_rows_once_and_reduce_is_deterministic():
    for n, ws in ((10, 3), (50_000, 8), (7, 8), (10_010_624, 8)):
        spans = [shard_rows(n, ws, r) for r in range(ws)]
        assert sum(c for _, c in spans) == n
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] or spans[i + 1][1] == 0 for i in range(ws - 1))
    assert reduce_winners(np.array([3.0, 1.0, 1.0, np.inf]), np.array([5, 40, 12, -1])) == (1.0, 12)
    assert reduce_winners(np.array([np.inf, np.inf]), np.array([-1, -1])) == (float("inf"), -1)


def test_c_abi_library_exports_every_declared_symbol():
    from frenetix_motion_planner_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "frx.h")).read()
    declared = set(re.findall(r"\b(frx_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("frx_ctx")
    assert len(declared) >= 20
    lib = _capi.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/frx.h but not exported by libfrx_b200.so"
    assert set(_capi.EXPORTS) == declared
    assert lib.frx_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from frenetix_motion_planner_b200 import _capi
    with pytest.raises(_capi.FrxError):
        _capi.Handler(0)
    with pytest.raises(_capi.FrxError):
        _capi.load_library("/nonexistent/libfrx_b200.so")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "frenetix_motion_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").replace("the oracle", "").lower() or f == "synthetic.py", f


def test_bundle_reads_the_selected_row_from_the_result_record():
    """TrajectoryBundle.states_of(argmin) uses the winner rows published with the arg-min (no device gather); any other
    row goes through the gather."""
    from frenetix_motion_planner_b200.trajectories import TrajectoryBundle

    class FakeHandler:
        Nt = 4
        calls = []

        def winner_states(self, fields=None):
            self.calls.append("winner")
            return np.full((14, 4), 7.0)

        def get_states(self, idx, fields=None):
            self.calls.append(("gather", int(idx[0])))
            return np.full((14, len(idx), 4), 3.0)

    h = FakeHandler()
    b = TrajectoryBundle(h, 10, ["lateral_jerk"], [1.0], 0.1, 0.3, 4, False, sampling=np.zeros((10, 13)))
    assert b.winner_row is None
    assert b.states_of(5)[0, 0] == 3.0 and h.calls[-1] == ("gather", 5)
    b.winner_row = 5
    assert b.states_of(5)[0, 0] == 7.0 and h.calls[-1] == "winner"
    assert b.states_of(6)[0, 0] == 3.0 and h.calls[-1] == ("gather", 6)


def test_state_tensor_index_is_a_bijection_onto_blocks_of_32_candidates():
    """frx_state_index (csrc/frx_device.cuh): [block of 32 candidates][step][field][32] -- restated here, checked to be a
    bijection onto [0, 14 * Nt * Np) with the 14 fields of one (block, step) contiguous."""
    def index(row, Nt, nf, f, i):
        return (((row >> 5) * Nt + i) * nf + f) * 32 + (row & 31)
    Nt, nf, N = 5, 14, 70
    Np = (N + 31) // 32 * 32
    seen = set()
    for row in range(Np):
        for i in range(Nt):
            for f in range(nf):
                seen.add(index(row, Nt, nf, f, i))
    assert seen == set(range(nf * Nt * Np))
    base = index(37, Nt, nf, 0, 2)
    assert [index(37, Nt, nf, f, 2) - base for f in range(nf)] == [32 * f for f in range(nf)]
    src = open(os.path.join(ROOT, "frenetix_motion_planner_b200", "csrc", "frx_device.cuh")).read()
    assert "(((size_t)(row >> 5) * (size_t)Nt + (size_t)i) * (size_t)nf + (size_t)f) * 32 + (size_t)(row & 31)" in src


def test_reference_arm_prints_one_json_line_on_cpu():
    """bench.py --impl reference runs on the host cores only (oracle port) and emits the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_set_iteration_order_equals_the_reference_for_every_level():
    """the row order (= uniqueId, = how equal-cost ties break) is the iteration order of the reference's SET
    objects; a copy of a set may iterate differently.  24 random v / d / t configurations x 4 levels, python and cpp
    style, against lists recorded from the reference's own SamplingHandler (tests/golden/ref_sampling_order.npz)."""
    g = np.load(os.path.join(GOLDEN_DIR, "ref_sampling_order.npz"))
    n_unsorted = 0
    for k, (v_lo, v_hi, d_half, t_min, horizon, d0, ss0) in enumerate(g["cases"]):
        sh = SamplingHandler(dt=0.1, max_sampling_number=4, t_min=t_min, horizon=horizon, delta_d_max=d_half,
                             delta_d_min=-d_half, d_ego_pos=False)
        sh.set_v_sampling(v_lo, v_hi)
        x_cl = ([3.0, ss0, 0.0], [d0, 0.0, 0.0])
        for lvl in range(4):
            t, v, d = sampling_axes(sh, lvl, x_cl)
            assert list(t) == list(g[f"c{k}_l{lvl}_t"]) and list(v) == list(g[f"c{k}_l{lvl}_v"])
            assert list(d) == list(g[f"c{k}_l{lvl}_d"])
            n_unsorted += list(v) != sorted(v)
            rows = python_path_rows(t, v, d, x_cl)
            want = [(a, b, c) for a in g[f"c{k}_l{lvl}_t"] for b in g[f"c{k}_l{lvl}_v"] for c in g[f"c{k}_l{lvl}_d"]]
            assert np.array_equal(rows[:, [1, 5, 10]], np.array(want))
            tc, vc, dc = sampling_axes(sh, lvl, x_cl, cpp_style=True)
            # the cpp union value is N * dT of the reference; round(horizon / dt) * dt here: same double
            assert list(tc) == list(g[f"c{k}_l{lvl}_t_cpp"]) and list(vc) == list(g[f"c{k}_l{lvl}_v_cpp"])
    assert n_unsorted > 10          # the cases really exercise hash order, not sorted order


def test_initial_frenet_state_equals_the_reference_code():
    """planner.py:567-635: ReactivePlannerB200._compute_initial_states against the reference's OWN function run on an
    independent projection (tests/golden/make_golden.py: initial_state_cases) -- 48 poses, both velocity modes."""
    from types import SimpleNamespace
    from frenetix_motion_planner_b200.reactive_planner_b200 import ReactivePlannerB200
    from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem
    g = np.load(os.path.join(GOLDEN_DIR, "ref_initial_states.npz"))
    for name in ("arc", "scurve", "straight", "tjunction"):
        cs = CoordinateSystem(g[f"{name}_polyline"])
        for row in g[f"{name}_cases"]:
            x_0 = SimpleNamespace(position=row[:2], orientation=row[2], velocity=row[3], acceleration=row[4],
                                  steering_angle=row[5], yaw_rate=0.0, time_step=0)
            me = SimpleNamespace(coordinate_system=cs, vehicle_params=SimpleNamespace(**syn.VEHICLE_2), _LOW_VEL_MODE=bool(row[6]))
            lon, lat = ReactivePlannerB200._compute_initial_states(me, x_0)
            got, want = np.array(list(lon) + list(lat)), row[7:]
            assert np.all(np.abs(got - want) <= 1e-9 * np.maximum(1.0, np.abs(want))), (name, got, want)


def test_reference_path_preparation_equals_the_reference_code():
    """extend_ref_path_both_ends / smooth_ref_path (utils_coordinate_system.py:20-58,110-134) against outputs of the
    reference's own functions (tests/golden/make_golden.py: refpath_cases)."""
    from frenetix_motion_planner_b200 import reference_path as rp
    g = np.load(os.path.join(GOLDEN_DIR, "ref_refpath.npz"))
    for name in ("tjunction", "scurve", "arc"):
        ext = rp.extend_ref_path_both_ends(g[f"{name}_route"])
        assert np.array_equal(ext, g[f"{name}_extended"])
        assert np.array_equal(rp.extend_ref_path_both_ends(g[f"{name}_route"], 80), g[f"{name}_extended_80"])
        sm = rp.smooth_ref_path(ext)
        assert sm.shape == g[f"{name}_smooth"].shape and np.allclose(sm, g[f"{name}_smooth"], rtol=0, atol=1e-10)
        seg = np.sqrt(np.sum(np.diff(sm, axis=0) ** 2, axis=1))
        assert np.all(np.abs(seg[:-1] - 1.0) < 2e-2)                        # 1 m of arc length between vertices
    # the T-junction fixture's reference path IS this pipeline's output
    assert np.array_equal(np.load(os.path.join(GOLDEN_DIR, "tjunction.npz"))["reference_path"], g["tjunction_smooth"])
