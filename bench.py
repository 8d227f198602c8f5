#!/usr/bin/env python
"""bench.py -- candidate trajectories evaluated per second per planning step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config5|config3|config2|config4]
                    [--impl reference] [--no-also] [--no-cpu-baseline]

One "step" = one pass of the hot path (sample -> back-project -> gates -> costs -> collision ->
arg-min) over one batch of synthetic candidates.

Default workload at EVERY N = BASELINE.json configs[4] ("config5"): ~10^7 candidates, 51 samples, 5 cost
terms, 50 predicted obstacles, rows generated on the device from the three host axes -- the largest
configuration that fits one B200 (57 GB of states) and the one the multi-GPU target is quoted on;
with N > 1 (torchrun) the SAME grid is sharded over the ranks (strong scaling) and the ranks exchange
the 16-byte (min_cost, row) record per step.  The line also carries, under "also", complete lines
(own roofline / e2e / cpu_baseline) for configs[2] ("config3": 200,000 candidates, 20 obstacles -- the
1-GPU target configuration of the north star) and configs[1] ("config2": 50,000 candidates, no
obstacles); those two scale weakly (every rank its own 200k / 50k shard of an N-times denser v axis).

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each number is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402
from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem  # noqa: E402

_SAVED_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print on fd 1 meanwhile (NCCL banner ...) goes to stderr."""
    global _SAVED_STDOUT
    if _SAVED_STDOUT is None:
        sys.stdout.flush()
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    if _SAVED_STDOUT is not None:
        os.dup2(_SAVED_STDOUT, 1)
    print(json.dumps(obj), flush=True)
    if _SAVED_STDOUT is not None:
        os.dup2(2, 1)


METRIC = "candidate trajectories evaluated/sec per planning step"
UNIT = "candidates/s"
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def host_threads() -> int:
    """Hardware threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arms ignore that)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def build_workload(name: str, world: int):
    """-> dict(polyline, x_cl, t1, v1, d1, N, dt, x0_orientation, v_des, low_vel, preds, scaling, label, grid_mode)"""
    if name == "config2":
        poly = syn.straight_polyline(400)
        x_cl = ([10.0, 8.0, 0.0], [0.2, 0.0, 0.0])
        t1 = np.round(np.arange(11, 31) * 0.1, 2)
        v_lo, v_hi = syn.velocity_interval(8.0, syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"])
        scale = int(os.environ.get("FRX_BENCH_SCALE", "1"))   # tuning aid: N-times more rows per GPU (not the headline)
        v1 = np.linspace(v_lo, v_hi, 50 * world * scale)     # weak scaling: denser v axis, 50k rows per rank
        d1 = np.linspace(-3.0, 3.0, 50)
        return dict(name=name, polyline=poly, x_cl=x_cl, t1=t1, v1=v1, d1=d1, N=30, dt=0.1, x0_orientation=0.0, v_des=8.0,
                    v0=8.0, preds=[], scaling="weak", grid_mode=False,
                    label="configs[1]: straight ref path (M=400), 20t x 50v x 50d = 50,000 candidates per GPU, "
                          "31 samples, 5 cost terms, 0 obstacles, fp64")
    if name == "config3":
        poly = syn.arc_polyline(R=80.0, M=400)
        x_cl = ([12.0, 9.5, 0.4], [-0.3, 0.2, -0.1])
        t1 = np.round(np.arange(11, 31) * 0.1, 2)
        v_lo, v_hi = syn.velocity_interval(9.5, syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"])
        v1 = np.linspace(v_lo, v_hi, 100 * world)
        d1 = np.linspace(-3.0, 3.0, 100)
        preds = syn.synthetic_predictions(poly, 20, 31, 0.1, seed=1234)
        return dict(name=name, polyline=poly, x_cl=x_cl, t1=t1, v1=v1, d1=d1, N=30, dt=0.1, x0_orientation=0.2, v_des=10.0,
                    v0=9.5, preds=preds, scaling="weak", grid_mode=False,
                    label="configs[2]-shaped: curved ref path, 20t x 100v x 100d = 200,000 candidates per GPU, "
                          "31 samples, 5 cost terms, 20 predicted obstacles (synthetic), fp64")
    if name == "config5":
        poly = syn.arc_polyline(R=200.0, M=600)
        x_cl = ([15.0, 12.0, 0.2], [0.1, 0.05, 0.0])
        t1 = np.unique(np.round(np.linspace(1.1, 5.0, 50), 2))
        v_lo, v_hi = syn.velocity_interval(12.0, syn.VEHICLE_2["a_max"], 5.0, syn.VEHICLE_2["v_max"])
        nv5 = int(os.environ.get("FRX_BENCH_C5_V", "448"))      # tuning aid (profiler captures): fewer rows, not the headline
        v1 = np.linspace(v_lo, v_hi, nv5)
        d1 = np.linspace(-3.0, 3.0, 447)
        preds = syn.synthetic_predictions(poly, 50, 51, 0.1, seed=2025, s_hi=200.0)
        total = t1.size * v1.size * d1.size
        return dict(name=name, polyline=poly, x_cl=x_cl, t1=t1, v1=v1, d1=d1, N=50, dt=0.1, x0_orientation=0.1, v_des=13.0,
                    v0=12.0, preds=preds, scaling="strong", grid_mode=True,
                    label=f"configs[4]: R=200 m arc, {t1.size}t x {nv5}v x 447d = {total:,} candidates sharded over the "
                          f"GPUs, 51 samples, 5 cost terms, 50 predicted obstacles, rows generated on device, fp64")
    if name == "config4":
        # multi-agent: 6 agents (the ego + the 5 cars of the T-junction fixture are agents in main_multiagent.py),
        # 50,000 candidates each, own reference path / Frenet state / predictions (the others' motion) per agent
        agents = []
        for a in range(6):
            poly = [syn.straight_polyline(400), syn.arc_polyline(R=80.0, M=400), syn.scurve_polyline(M=400),
                    syn.arc_polyline(R=150.0, M=400, start_heading=0.5), syn.scurve_polyline(M=400, amp=3.0),
                    syn.arc_polyline(R=60.0, M=400, start_heading=-0.3)][a]
            v0 = 6.0 + a
            x_cl = ([10.0 + 2 * a, v0, 0.1 * a], [0.1 * a - 0.2, 0.02 * a, 0.0])
            v_lo, v_hi = syn.velocity_interval(v0, syn.VEHICLE_2["a_max"], 3.0, syn.VEHICLE_2["v_max"])
            agents.append(dict(polyline=poly, x_cl=x_cl, v0=v0, x0_orientation=0.05 * a, v_des=v0 + 1.0,
                               t1=np.round(np.arange(11, 31) * 0.1, 2), v1=np.linspace(v_lo, v_hi, 50),
                               d1=np.linspace(-3.0, 3.0, 50), preds=syn.synthetic_predictions(poly, 5, 31, 0.1, seed=100 + a)))
        return dict(name=name, agents=agents, N=30, dt=0.1, scaling="weak", preds=agents[0]["preds"], grid_mode=False,
                    polyline=agents[0]["polyline"], x_cl=agents[0]["x_cl"], t1=agents[0]["t1"], v1=agents[0]["v1"],
                    d1=agents[0]["d1"], x0_orientation=0.0, v_des=7.0, v0=6.0,
                    label="configs[3]: 6 agents x 50,000 candidates (own reference path, state and 5 predicted obstacles "
                          "each) batched into ONE eval-kernel launch, 31 samples, 5 cost terms, fp64")
    raise SystemExit(f"unknown workload {name}")


def grid_rows(w, idx) -> np.ndarray:
    """Sampling rows [len(idx), 13] of the global row numbers `idx` of the (t1, v1, d1) grid (generate_sampling_matrix
    order: t1 slowest, then ss1, then d1)."""
    idx = np.asarray(idx, dtype=np.int64)
    nv, nd = w["v1"].size, w["d1"].size
    it, rem = np.divmod(idx, nv * nd)
    iv, idd = np.divmod(rem, nd)
    (s0, ss0, sss0), (d0, dd0, ddd0) = w["x_cl"]
    S = np.zeros((idx.size, 13))
    S[:, 1] = w["t1"][it]
    S[:, 2], S[:, 3], S[:, 4] = s0, ss0, sss0
    S[:, 5] = w["v1"][iv]
    S[:, 7], S[:, 8], S[:, 9] = d0, dd0, ddd0
    S[:, 10] = w["d1"][idd]
    return S


def cpu_sample_indices(w, n_total: int, first: int, count: int, target: int) -> np.ndarray:
    """A bounded, representative sample of this rank's rows for the CPU arms: every k-th row of the shard."""
    stride = max(1, count // target)
    return first + np.arange(0, count, stride, dtype=np.int64)


def oracle_inputs(w):
    from oracle import frenet_oracle as fo
    cs = CoordinateSystem(w["polyline"])
    ref = fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d,
                     np.ascontiguousarray(w["polyline"][:, 0]), np.ascontiguousarray(w["polyline"][:, 1]))
    prm = fo.Params(dt=w["dt"], N=w["N"], low_vel_mode=w["v0"] < 2.0, x0_orientation=w["x0_orientation"],
                    desired_velocity=w["v_des"], draw_traj_set=True, kinematic_debug=True,
                    cost_weights=dict(syn.DEFAULT_COST_WEIGHTS),
                    **{k: syn.VEHICLE_2[k] for k in ("a_max", "v_switch", "delta_max", "wheelbase", "wb_rear_axle",
                                                      "length", "width")})
    return ref, prm


def algorithmic_bytes_per_candidate(Nt: int, K: int, matrix_input: bool) -> int:
    """104 B sampling row (0 when rows are generated on device) + 14 fields x Nt x 8
    + 8 (total) + 8K (unweighted costs) + 4 (flags) + 4 (traj_len)."""
    return (104 if matrix_input else 0) + 112 * Nt + 8 * K + 16


def obstacle_flops_per_candidate(Nt: int, preds, check_collisions: bool = True) -> int:
    """Algorithmic fp64 work of the obstacle pass per candidate: per step i >= 1 and obstacle predicted
    at that step the inverse-Mahalanobis term = 12 flop (2 subtractions, the quadratic form 7, square, reciprocal and
    accumulate counted 1 each); per step the ego obb-sum hull = 60 flop (box centre, hull in the frame of box k, bounding
    radius).  Pair tests of the collision sweep are NOT counted: the warp-level cull removes almost all of them."""
    pairs = 0
    for p in preds:
        pairs += max(0, min(Nt, len(p["pos_list"])) - 1)
    return 12 * pairs + (60 * (Nt - 1) if (check_collisions and len(preds)) else 0)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.samples if t0 <= t <= t1] or [l for (_, l) in self.samples[-3:]]
        sm, mx, reasons = [], [], set()
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------------------------
def cpu_abi_handler(w, S):
    """The reference's CPU path behind the SAME C ABI the GPU arm is called through: oracle/cpu_abi (include/frx.h on the
    host cores = the C/OpenMP port of the reference's Python path with its lazy collision walk), built for THIS host
    (-O3 -march=native, contraction off) and configured exactly like the device handler."""
    from oracle.build import build_cpu_abi
    from frenetix_motion_planner_b200 import _capi, hotpath
    threads = host_threads()
    os.environ["FRX_CPU_THREADS"] = str(threads)               # torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores that
    lib = _capi.load_library(path=build_cpu_abi(native=True))
    lib.orc_last_threads.restype = __import__("ctypes").c_int
    h = _capi.Handler(0, library=lib)
    cs = CoordinateSystem(w["polyline"])
    names, weights = hotpath.active_costs(syn.DEFAULT_COST_WEIGHTS)
    veh = syn.VEHICLE_2
    h.set_params(dt=w["dt"], N=w["N"], low_vel_mode=w["v0"] < 2.0, draw_traj_set=True, kinematic_debug=True,
                 a_max=veh["a_max"], v_switch=veh["v_switch"], delta_max=veh["delta_max"], wheelbase=veh["wheelbase"],
                 wb_rear_axle=veh["wb_rear_axle"], length=veh["length"], width=veh["width"],
                 x0_orientation=w["x0_orientation"], desired_velocity=w["v_des"], cost_names=names,
                 cost_weights=weights, store_states=True, check_collisions=True)
    h.set_reference(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, w["polyline"][:, 0], w["polyline"][:, 1])
    h.set_time_tables(*hotpath.time_tables(np.unique(S[:, 1]), w["dt"], w["N"] + 1))
    packed = hotpath.pack_predictions(w["preds"])
    if packed is not None:
        h.set_predictions(*packed)
    return h, lib, threads


def cpu_port_throughput(w, S, budget_s=12.0, min_reps=2, max_reps=200):
    """Time the CPU implementation of the ABI on `S` with all host threads; returns (cand/s, threads, reps, seconds)."""
    h, lib, threads = cpu_abi_handler(w, S)
    h.plan(S)                                                  # warm-up (threads, pages)
    reps, t_acc = 0, 0.0
    while reps < min_reps or (t_acc < budget_s and reps < max_reps):
        t0 = time.perf_counter()
        h.plan(S)
        t_acc += time.perf_counter() - t0
        reps += 1
    assert int(lib.orc_last_threads()) == threads, "CPU arm did not get every host thread"
    return S.shape[0] * reps / t_acc, threads, reps, t_acc


def python_path_throughput(w, S, n=64):
    """The restated Python path itself (numpy oracle, 1 core) on a small sample."""
    from oracle import frenet_oracle as fo
    ref, prm = oracle_inputs(w)
    idx = np.linspace(0, S.shape[0] - 1, n).astype(int)
    t0 = time.perf_counter()
    fo.plan(S[idx], ref, prm, w["preds"], check_all_collisions=False)
    return n / (time.perf_counter() - t0)


def oracle_confirms_winner(w, n_total: int, winner_row: int, winner_cost: float, n_sample: int = 60_000):
    """Selected-trajectory check at sizes the oracle cannot cover exhaustively: the C oracle (every collision checked)
    evaluates a strided sample of the WHOLE grid plus the winner's (t1, ss1) slab and its ss1 neighbours.  The device
    winner is the global arg-min, so it must also be the arg-min of any subset that contains it, with the same cost."""
    from oracle import c_oracle
    if winner_row < 0:
        return None
    nd, nv = w["d1"].size, w["v1"].size
    idx = [np.arange(0, n_total, max(1, n_total // n_sample), dtype=np.int64)]
    slab = (winner_row // nd) * nd
    for dv in (-1, 0, 1):
        s0 = slab + dv * nd
        if 0 <= s0 and s0 + nd <= n_total:
            idx.append(np.arange(s0, s0 + nd, dtype=np.int64))
    idx = np.unique(np.concatenate(idx))
    ref, prm = oracle_inputs(w)
    out = c_oracle.plan(grid_rows(w, idx), ref, prm, w["preds"], check_all_collisions=True, want_states=False,
                        want_margins=False, nthreads=host_threads())
    o_row = int(idx[out["argmin"]]) if out["argmin"] >= 0 else -1
    same_cost = abs(out["min_cost"] - winner_cost) <= 1e-9 * max(1.0, abs(winner_cost))
    return {"rows_checked": int(idx.size), "oracle_row": o_row, "match": bool(o_row == winner_row and same_cost),
            "oracle_cost": float(out["min_cost"])}


def run_reference_arm(args, w, world):
    """--impl reference: the reference's CPU path on the host cores.  The reference itself (pure Python + un-vendored
    frenetix/commonroad wheels) cannot be installed offline, so this times the C/OpenMP port of its Python path behind
    the same C ABI (oracle/cpu_abi), with every hardware thread of the host."""
    n_total = w["t1"].size * w["v1"].size * w["d1"].size
    rows = n_total if w["scaling"] == "strong" else n_total // world     # the workload of ONE GPU-arm rank 0 step
    target = 100_000 if w["preds"] else 200_000
    idx = cpu_sample_indices(w, n_total, 0, rows, target)
    S = grid_rows(w, idx)
    h, lib, threads = cpu_abi_handler(w, S)
    vals = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = h.plan(S)
        if r.argmin >= 0:
            h.winner_states()
        if i >= args.warmup:
            vals.append(time.perf_counter() - t0)
    assert int(lib.orc_last_threads()) == threads == host_threads(), "CPU arm did not get every host thread"
    sec_per_step = float(np.mean(vals))
    value = S.shape[0] / sec_per_step
    sample = (f"each step = every {max(1, rows // target)}-th row of the workload ({S.shape[0]:,} of {rows:,} rows) through "
              f"frx_plan of the CPU implementation of the C ABI (oracle/cpu_abi: C/OpenMP port of the reference's Python path, "
              f"gcc -O3 -march=native -ffp-contract=off) on {threads} threads; frenetix 0.4.0 / the Python reference are not "
              f"installable offline")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "rows_per_step": int(S.shape[0]), "note": f"CPU, {threads} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_config4(args, w, local_rank):
    """Multi-agent batch: one frx_plan_batched call per step (host matrices in, H2D inside)."""
    import torch
    from frenetix_motion_planner_b200 import _capi, hotpath
    veh = syn.VEHICLE_2
    names, weights = hotpath.active_costs(syn.DEFAULT_COST_WEIGHTS)
    handlers, mats = [], []
    for ag in w["agents"]:
        h = _capi.Handler(local_rank)
        cs = CoordinateSystem(ag["polyline"])
        h.set_params(dt=w["dt"], N=w["N"], low_vel_mode=ag["v0"] < 2.0, draw_traj_set=True, kinematic_debug=True,
                     a_max=veh["a_max"], v_switch=veh["v_switch"], delta_max=veh["delta_max"], wheelbase=veh["wheelbase"],
                     wb_rear_axle=veh["wb_rear_axle"], length=veh["length"], width=veh["width"],
                     x0_orientation=ag["x0_orientation"], desired_velocity=ag["v_des"], cost_names=names,
                     cost_weights=weights)
        h.set_reference(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, ag["polyline"][:, 0], ag["polyline"][:, 1])
        h.set_time_tables(*hotpath.time_tables(np.unique(ag["t1"]), w["dt"], w["N"] + 1))
        h.set_predictions(*hotpath.pack_predictions(ag["preds"]))
        handlers.append(h)
        mats.append(torch.from_numpy(syn.grid_sampling_matrix(ag["t1"], ag["v1"], ag["d1"], ag["x_cl"])).pin_memory().numpy())
    rows = sum(m.shape[0] for m in mats)
    for _ in range(args.warmup):
        res = _capi.plan_batched(handlers, mats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kms = []
    for _ in range(args.steps):
        res = _capi.plan_batched(handlers, mats)
        kms.append(res[0].eval_kernel_ms)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same agents one after the other (what AgentBatch._step_agents does)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solo = [h.plan(m) for h, m in zip(handlers, mats)]
    dt_solo = time.perf_counter() - t0
    same = all(a.argmin == b.argmin and a.min_cost == b.min_cost for a, b in zip(res, solo))
    B_cand = algorithmic_bytes_per_candidate(w["N"] + 1, len(names), True)
    peak, _src = hbm_peak()
    kmean = float(np.mean(kms))
    ach = rows * B_cand / (kmean * 1e-3) / 1e9
    return {
        "metric": METRIC, "value": rows * args.steps / dt, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "agents": len(handlers), "rows_total": rows,
                   "timing": "wall clock around frx_plan_batched (pinned host matrices in, H2D inside)",
                   "sequential_plans_ms_per_step": dt_solo / args.steps * 1e3, "batched_equals_sequential": bool(same)},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "frx_eval_batched_kernel", "kernel_ms": kmean, "algorithmic_bytes_per_candidate": B_cand},
        "e2e": {"value": rows * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": rows * 104,
                "d2h_bytes_per_step": 152 * len(handlers)},
        "gpu_launches": args.steps * (1 + len(handlers))}


def planner_e2e(local_rank: int):
    """The public Python API a user of the reference calls: ReactivePlannerB200.plan() end to end (host work, H2D of the
    sampling axes, kernels, D2H of the result record and the selected trajectory, output conversion to the trajectory
    pair), wall clock per call.  (a) configs[0]: the ZAM_Tjunction-1_42_T-1 fixture at the reference's default sampling
    (630 candidates, 5 predicted cars, road boundary), with the reference's default debug flags (draw_traj_set=True keeps
    the cost-sorted candidate set for visualisation) and without; (b) the configs[2] grid (200,000 candidates,
    20 obstacles) through the same call."""
    import types
    from frenetix_motion_planner_b200 import ReactivePlannerB200
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    fx = np.load(os.path.join(gold, "tjunction.npz"))
    raw = json.load(open(os.path.join(gold, "tjunction_lanelets.json")))
    lanelets = {int(k): dict(left=np.array(v["left"]), right=np.array(v["right"]), adj_left=v["adj_left"], adj_right=v["adj_right"])
                for k, v in raw.items()}

    def make(draw, horizon=3.0, d_levels=None):
        cfg_plan = types.SimpleNamespace(
            planning=types.SimpleNamespace(planning_horizon=horizon, dt=0.1, low_vel_mode_threshold=2.0, sampling_min=2, sampling_max=3,
                                           t_min=1.1, d_min=-3, d_max=3, d_ego_pos=False, replanning_frequency=3),
            debug=types.SimpleNamespace(multiproc=True, num_workers=6, draw_traj_set=draw, kinematic_debug=draw, save_all_traj=False,
                                        log_risk=False),
            cost=types.SimpleNamespace(cost_weights=dict(syn.DEFAULT_COST_WEIGHTS)))
        return ReactivePlannerB200(cfg_plan, types.SimpleNamespace(vehicle=types.SimpleNamespace(**syn.VEHICLE_2)), None, None, None, None,
                                   None, device=local_rank)

    def timeit(p, n=200, warm=20):
        for _ in range(warm):
            pair = p.plan()
        t0 = time.perf_counter()
        for _ in range(n):
            pair = p.plan()
        dt = (time.perf_counter() - t0) / n
        return dt, pair

    preds = {}
    for o, oid in enumerate(fx["obstacle_ids"]):
        st = fx["obstacle_states"][o, 1:32]
        preds[int(oid)] = {"pos_list": st[:, :2].copy(), "cov_list": np.tile(np.array([[0.1, 0.0], [0.0, 0.1]]), (31, 1, 1)),
                           "orientation_list": st[:, 2].copy(), "v_list": st[:, 3].copy(),
                           "shape": {"length": float(fx["obstacle_shapes"][o, 0]) + 0.5, "width": float(fx["obstacle_shapes"][o, 1]) + 0.2}}
    x_0 = types.SimpleNamespace(position=fx["ego_position_rear"], orientation=float(fx["ego_orientation"]), velocity=float(fx["ego_velocity"]),
                                acceleration=float(fx["ego_acceleration"]), yaw_rate=float(fx["ego_yaw_rate"]), steering_angle=0.0, time_step=0)
    for draw in (True, False):
        p = make(draw)
        p.obstacle_order = [int(i) for i in fx["obstacle_ids"]]
        p.set_road_boundary(lanelets)
        p.update_externals(reference_path=fx["reference_path"], x_0=x_0, x_cl=None, desired_velocity=8.0, predictions=preds)
        dt, pair = timeit(p)
        out["config1_draw_traj_set" if draw else "config1"] = {
            "us_per_plan": dt * 1e6, "candidates": int(p._total_count), "candidates_per_s": p._total_count / dt,
            "selected": int(p.optimal_trajectory.uniqueId), "eval_kernel_us": 1e3 * float(p.last_plan_stats.eval_kernel_ms),
            "obstacles": len(preds), "road_boundary_boxes": int(len(p.static_obbs))}
        p.handler.close()
    # (b) the configs[2] grid through plan(): the planner's level sets replaced by the dense axes
    w = build_workload("config3", 1)
    p = make(False)
    cs_poly = w["polyline"]
    x_0b = types.SimpleNamespace(position=None, orientation=w["x0_orientation"], velocity=w["v0"], acceleration=0.0, yaw_rate=0.0,
                                 steering_angle=0.0, time_step=0)
    p.update_externals(reference_path=cs_poly, x_0=x_0b, x_cl=w["x_cl"], desired_velocity=w["v_des"],
                       predictions={100 + i: q for i, q in enumerate(w["preds"])})
    dense = (w["t1"], w["v1"], w["d1"])
    p._level_axes = lambda samp_level: dense
    dt, pair = timeit(p, n=50, warm=5)
    out["config3_grid"] = {"us_per_plan": dt * 1e6, "candidates": int(p._total_count), "candidates_per_s": p._total_count / dt,
                           "selected": int(p.optimal_trajectory.uniqueId), "obstacles": len(w["preds"]),
                           "eval_kernel_us": 1e3 * float(p.last_plan_stats.eval_kernel_ms)}
    p.handler.close()
    return out


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# one workload on the GPU(s) -> one bench line
# ----------------------------------------------------------------------------------------------
class Env:
    """Process-wide state shared by the workloads of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.l2_drain = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.int64, device=self.dev)
        self.fp64_peak = None

    def flush_l2(self):
        """Cold L2 for the timed step: write 256 MiB (> 126 MB L2: evicts inputs and outputs of the previous step),
        then read another 256 MiB so that the dirty lines of the write pass are drained to HBM before the timed
        kernel starts (otherwise it pays for writing the flush buffer back while it streams its own output)."""
        self.l2_flush.zero_()
        self.l2_drain.sum()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


def run_gpu_workload(env: Env, name: str, steps: int, warmup: int, cpu_baseline: bool, headline: bool):
    torch, dist = env.torch, env.dist
    from frenetix_motion_planner_b200 import _capi, hotpath
    from frenetix_motion_planner_b200.dist import ArgminExchange, SharedPageExchange, shard_rows, single_node
    rank, world, dev, stream = env.rank, env.world, env.dev, env.stream
    w = build_workload(name, world)
    grid_mode = w["grid_mode"]
    n_total = w["t1"].size * w["v1"].size * w["d1"].size
    first, count = shard_rows(n_total, world, rank)

    # ---- set-up (not timed: once per scenario / planning cycle in the reference, too)
    cs = CoordinateSystem(w["polyline"])
    names, weights = hotpath.active_costs(syn.DEFAULT_COST_WEIGHTS)
    K, Nt = len(names), w["N"] + 1
    h = _capi.Handler(env.local_rank)
    h.set_stream(stream.cuda_stream)
    veh = syn.VEHICLE_2
    h.set_params(dt=w["dt"], N=w["N"], low_vel_mode=w["v0"] < 2.0, draw_traj_set=True, kinematic_debug=True,
                 a_max=veh["a_max"], v_switch=veh["v_switch"], delta_max=veh["delta_max"], wheelbase=veh["wheelbase"],
                 wb_rear_axle=veh["wb_rear_axle"], length=veh["length"], width=veh["width"],
                 x0_orientation=w["x0_orientation"], desired_velocity=w["v_des"], cost_names=names,
                 cost_weights=weights, store_states=True, check_collisions=True)
    h.set_reference(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, w["polyline"][:, 0], w["polyline"][:, 1])
    h.set_time_tables(*hotpath.time_tables(np.unique(w["t1"]), w["dt"], Nt))
    packed = hotpath.pack_predictions(w["preds"])
    if packed is not None:
        h.set_predictions(*packed)
    # the per-plan exchange of the 16-byte winner records: a shared pinned page the kernels' last CTAs write into (one node),
    # else one NCCL all-gather (FRX_BENCH_EXCHANGE=nccl forces it)
    use_page = world > 1 and single_node() and os.environ.get("FRX_BENCH_EXCHANGE", "page") != "nccl"
    page = SharedPageExchange(h) if use_page else None
    ex = ArgminExchange() if (world > 1 and not use_page) else None
    if env.fp64_peak is None:
        env.fp64_peak = h.fp64_peak_tflops()

    if grid_mode:
        S_host = None
    else:
        S_host_t = torch.from_numpy(grid_rows(w, np.arange(first, first + count))).pin_memory()   # pinned host rows of this rank
        S_host = S_host_t.numpy()
        S_dev = S_host_t.to(dev)
    launches = {"n": 0}

    def step_resident():
        if page is not None:
            r = (h.plan_grid(w["t1"], w["v1"], w["d1"], w["x_cl"], row_first=first, row_count=count) if grid_mode
                 else h.plan_device(S_dev.data_ptr(), count, row_index_base=first))
            page.finish()
        elif grid_mode:
            r = h.plan_grid(w["t1"], w["v1"], w["d1"], w["x_cl"], row_first=first, row_count=count)
            if ex is not None:
                ex.exchange(r.min_cost, r.argmin, handler=h)
        elif ex is not None:
            # plan, all-gather of the 16-byte winner records and their read-back are queued back to back;
            # the host blocks once
            h.plan_device_async(S_dev.data_ptr(), count, row_index_base=first)
            ex.enqueue(h)
            r = h.plan_wait()
            ex.finish()
        else:
            r = h.plan_device(S_dev.data_ptr(), count, row_index_base=first)
        launches["n"] += h.last_launches()
        return r

    def step_e2e():
        if grid_mode:
            r = h.plan_grid(w["t1"], w["v1"], w["d1"], w["x_cl"], row_first=first, row_count=count)
        else:
            r = h.plan(S_host, row_index_base=first)
        if page is not None:
            cost, row, owner = page.finish()
        else:
            cost, row, owner = (r.min_cost, r.argmin, 0) if ex is None else ex.exchange(r.min_cost, r.argmin, handler=h)
        win = None
        if row >= 0 and owner == rank:
            # the selected trajectory: published by the eval kernel with the arg-min (mapped result record)
            win = h.winner_states() if row == r.argmin else h.get_states(np.array([row - first], dtype=np.int64))
        return r, cost, row, win

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            env.flush_l2()
            step_resident()
        stream.synchronize()
        env.barrier()

        # ---- timed region 1: device-resident inputs, CUDA events per step on the kernel's stream
        sampler = ClockSampler(env.local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        env.barrier()                             # all ranks enter the timed loop together
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kern_ms, obs_ms = [], []
        launches["n"] = 0
        t_wall0 = time.perf_counter()
        for i in range(steps):
            env.flush_l2()                        # cold, clean L2 between timed iterations
            ev[i][0].record(stream)
            r = step_resident()
            ev[i][1].record(stream)
            kern_ms.append(r.eval_kernel_ms); obs_ms.append(r.obstacle_kernel_ms)
        stream.synchronize()
        env.barrier()
        t_wall1 = time.perf_counter()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(np.sum(step_ms))
        gpu_launches = launches["n"]
        last_launches = h.last_launches()

        # ---- timed region 2: end to end through the C ABI with HOST buffers (wall clock)
        for _ in range(3):
            step_e2e()
        stream.synchronize()
        env.barrier()
        t0 = time.perf_counter()
        e2e_kern_ms, e2e_dev_ms = [], []
        for i in range(steps):
            r_e2e, cost_e2e, row_e2e, win = step_e2e()
            e2e_kern_ms.append(r_e2e.eval_kernel_ms); e2e_dev_ms.append(r_e2e.total_device_ms)
        stream.synchronize()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        # ---- for information (N = 1, cartesian workloads): the same plan through frx_plan_grid, i.e. handing the
        # library the three host axes + x_cl the planner owns instead of the expanded [N, 13] matrix
        e2e_grid_s, grid_same = None, None
        if world == 1 and not grid_mode:
            def step_grid():
                rg = h.plan_grid(w["t1"], w["v1"], w["d1"], w["x_cl"])
                return rg, (h.winner_states() if rg.argmin >= 0 else None)
            for _ in range(3):
                rg, _w = step_grid()
            t0 = time.perf_counter()
            for i in range(steps):
                rg, _w = step_grid()
            e2e_grid_s = time.perf_counter() - t0
            grid_same = bool(rg.argmin == r_e2e.argmin and rg.min_cost == r_e2e.min_cost)
        env.barrier()
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

        # ---- multi-GPU selection checks (untimed): every rank must hold the same global record, and it must equal what
        # ONE GPU selects on the whole grid (rank 0 re-plans all rows once)
        ranks_agree, single_gpu = None, None
        if world > 1:
            rec = torch.tensor([float(cost_e2e), float(row_e2e)], dtype=torch.float64, device=dev)
            allrec = [torch.empty_like(rec) for _ in range(world)]
            dist.all_gather(allrec, rec)
            ranks_agree = bool(all(torch.equal(a, allrec[0]) for a in allrec))
            if rank == 0:
                if grid_mode:
                    r1 = h.plan_grid(w["t1"], w["v1"], w["d1"], w["x_cl"])
                else:
                    r1 = h.plan(grid_rows(w, np.arange(n_total)))
                single_gpu = {"row": int(r1.argmin), "cost": float(r1.min_cost),
                              "equal": bool(int(r1.argmin) == int(row_e2e) and float(r1.min_cost) == float(cost_e2e))}
            env.barrier()

    # ---- max over ranks
    eval_only = [a - b for a, b in zip(kern_ms, obs_ms)]
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, float(np.mean(kern_ms)), float(np.mean(obs_ms)), float(np.mean(eval_only))],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kern_mean_ms, obs_mean_ms, eval_mean_ms = t.tolist()
        cnt = torch.tensor([count], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        rows_all = int(cnt.item())
    else:
        kern_mean_ms, obs_mean_ms, eval_mean_ms = float(np.mean(kern_ms)), float(np.mean(obs_ms)), float(np.mean(eval_only))
        rows_all = count

    line = None
    if rank == 0:
        ms_per_step = total_ms / steps
        value = rows_all / (ms_per_step * 1e-3)
        e2e_value = rows_all * steps / e2e_s
        B_cand = algorithmic_bytes_per_candidate(Nt, K, matrix_input=not grid_mode)
        peak, peak_src = hbm_peak()
        achieved = count * B_cand / (eval_mean_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        roof_eval = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                     "kernel": "frx_eval_kernel", "kernel_ms": eval_mean_ms, "algorithmic_bytes_per_candidate": B_cand,
                     "peak_source": peak_src}
        roof_obs = None
        if obs_mean_ms > 0:
            F_cand = obstacle_flops_per_candidate(Nt, w["preds"])
            ach_tf = count * F_cand / (obs_mean_ms * 1e-3) / 1e12
            roof_obs = {"bound": "fp64", "achieved": ach_tf, "peak": env.fp64_peak, "unit": "TFLOP/s",
                        "frac": ach_tf / env.fp64_peak,
                        "traffic": None if traffic is None else traffic.get("obstacle_dram_bytes_per_launch"),
                        "kernel": "frx_obstacle_kernel", "kernel_ms": obs_mean_ms,
                        "algorithmic_flops_per_candidate": F_cand,
                        "peak_source": "measured live (frx_selftest_fp64_peak: independent DFMA streams on every SM, 2 flop per FMA)",
                        "note": "fp64-bound, not HBM-bound: 12 flop per (candidate, step, obstacle) + 60 per (candidate, step); the "
                                "kernel spends 9.25 fp64 issue slots on those 12 flop and the peak counts 2 flop per slot, so "
                                "12 / 18.5 = 0.65 is the ceiling of this fraction (DESIGN.md section 5)"}
        # the dominant kernel of the step carries the `roofline` key, the other one rides along
        dominant_obs = roof_obs is not None and obs_mean_ms > eval_mean_ms
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["label"], "rows_per_gpu": int(count), "rows_total": int(rows_all),
                       "Nt": Nt, "cost_terms": names, "obstacles": len(w["preds"]),
                       "input": "rows generated on device" if grid_mode else "sampling matrix [N,13] resident in HBM",
                       "l2": "L2 flushed between timed iterations (256 MiB write pass, then a 256 MiB read pass that drains the dirty flush lines); per-step state output "
                             f"{count * 112 * Nt / 1e6:.0f} MB > 126 MB L2",
                       "timing": "sum of per-step CUDA-event intervals on the launch stream, max over ranks",
                       "parallelism": (f"{world} x B200, contiguous row shards, one 16-B record exchanged per rank and step: " +
                                       ("each rank's last CTA stores it into a shared pinned page (64 B posted PCIe write per rank), every host "
                                        "reads all slots -- no collective kernel" if use_page else "one NCCL all-gather of 16 B per rank + read-back"))
                                      if world > 1 else "1 x B200"},
            "roofline": roof_obs if dominant_obs else roof_eval,
            ("roofline_eval_kernel" if dominant_obs else "roofline_obstacle_kernel"): roof_eval if dominant_obs else roof_obs,
            "kernels_ms": {"frx_eval_kernel": eval_mean_ms, "frx_obstacle_kernel": obs_mean_ms, "both": kern_mean_ms},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(8 * (len(w["t1"]) + len(w["v1"]) + len(w["d1"]) + 6) if grid_mode else count * 13 * 8),
                    "d2h_bytes_per_step": int(16 + 8 * 17 + 14 * Nt * 8),
                    "ms_per_step": 1e3 * e2e_s / steps, "eval_kernel_ms": float(np.mean(e2e_kern_ms)),
                    "device_ms": float(np.mean(e2e_dev_ms)),
                    "grid_api": (None if e2e_grid_s is None else
                                 {"value": count * steps / e2e_grid_s, "unit": UNIT, "h2d_bytes_per_step":
                                  int(8 * (len(w["t1"]) + len(w["v1"]) + len(w["d1"]) + 6)),
                                  "note": "same plan through frx_plan_grid (host axes t1/ss1/d1 + x_cl, rows expanded on "
                                          "the device)", "same_argmin_and_cost_as_matrix_api": grid_same}),
                    "note": ("frx_plan_grid on HOST axes t1/ss1/d1 + x_cl (rows expanded on the device); " if grid_mode else
                             "frx_plan on a PINNED HOST sampling matrix: the eval kernel reads the rows in place over PCIe "
                             "(cp.async prefetch one tile ahead, no staging copy; FRX_ZEROCOPY=0 restores cudaMemcpyAsync); ") +
                            "the result record and the selected trajectory's 14 state rows come back through mapped host "
                            "memory written by the kernel's last CTA; wall clock around the C-ABI call"},
            "gpu_launches": gpu_launches, "gpu_launches_per_step": last_launches, "clocks": clocks,
            "selected": {"row": int(row_e2e), "cost": float(cost_e2e), "n_feasible": int(r_e2e.n_feasible),
                         "n_collide": int(r_e2e.n_collide)},
        }
        if world > 1:
            line["selected"]["all_ranks_hold_the_same_record"] = ranks_agree
            line["selected"]["global_row_equals_single_gpu"] = single_gpu
        if world == 1:
            # selected-trajectory match against the oracle (every size): the device winner must be the oracle's arg-min
            # of a subset that contains it (for <= 200k rows the subset is the whole matrix)
            line["selected"]["selected_row_matches"] = oracle_confirms_winner(
                w, n_total, int(row_e2e), float(cost_e2e), n_sample=(n_total if n_total <= 200_000 else 60_000))
        if world == 1 and cpu_baseline:
            idx = cpu_sample_indices(w, n_total, first, count, 100_000 if w["preds"] else 200_000)
            S_cpu = S_host[idx - first] if S_host is not None else grid_rows(w, idx)
            thr, threads, reps, secs = cpu_port_throughput(w, S_cpu, budget_s=(12.0 if headline else 6.0))
            line["cpu_baseline"] = {
                "value": thr, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{reps} x every {max(1, count // S_cpu.shape[0])}-th row of the workload "
                          f"({S_cpu.shape[0]:,} rows, {secs:.1f} s) through the CPU implementation of the C ABI (oracle/cpu_abi: "
                          f"C/OpenMP port of the reference's Python path, gcc -O3 -march=native -ffp-contract=off)",
                "python_path_1core": python_path_throughput(w, S_cpu) if headline else None}
    if page is not None:
        page.close()
    h.close()
    del h
    torch.cuda.empty_cache()
    return line


# ----------------------------------------------------------------------------------------------
# main
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the config3 / config2 lines under 'also'")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, build_workload(args.workload, world), world)
        return

    quiet_stdout()
    env = Env(args)
    if args.workload == "config4":
        line = run_config4(args, build_workload("config4", 1), env.local_rank)
        if rank == 0:
            emit(line)
        return
    line = run_gpu_workload(env, args.workload, args.steps, args.warmup, not args.no_cpu_baseline, headline=True)
    if not args.no_also:
        also = {}
        for name in ("config3", "config2"):
            if name == args.workload:
                continue
            l2 = run_gpu_workload(env, name, args.steps, args.warmup, not args.no_cpu_baseline, headline=False)
            if rank == 0:
                also[name] = l2
        if rank == 0:
            line["also"] = also
    if rank == 0 and world == 1 and not args.no_also:
        line["e2e"]["planner"] = planner_e2e(env.local_rank)
    if rank == 0:
        emit(line)
    if world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
