"""Host-side plumbing between the reference-shaped planner objects and the C ABI.

Everything here is small per-plan host work (the reference does the same things once per
planning step, see the cited lines); the per-candidate work lives in csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np

from . import _capi


# --------------------------------------------------------------------------------------------
# time grid (reactive_planner.py:296-303): numpy semantics matter (np.arange length, np.round of
# the powers), so the tables are produced with numpy exactly like the reference and shipped.
# --------------------------------------------------------------------------------------------
_TABLE_CACHE: Dict[tuple, tuple] = {}


def time_table(T: float, dt: float, Nt: int):
    """-> (traj_len, tpow[5, Nt]) for one trajectory duration."""
    key = (float(T), float(dt), int(Nt))
    hit = _TABLE_CACHE.get(key)
    if hit is not None:
        return hit
    t = np.round(np.arange(0, T + dt, dt), 5)
    n = len(t)
    if n > Nt:
        raise ValueError(f"duration {T} s needs {n} samples but the planning horizon only has {Nt} "
                         f"(the reference would fail at reactive_planner.py:314)")
    tp = np.zeros((5, Nt))
    tp[0, :n] = t
    for k in range(2, 6):
        tp[k - 1, :n] = np.round(np.power(t, k), 10)
    _TABLE_CACHE[key] = (n, tp)
    return n, tp


def time_tables(T_values: Sequence[float], dt: float, Nt: int):
    T_values = np.asarray(T_values, dtype=np.float64)
    lens = np.zeros(T_values.size, dtype=np.int32)
    tpow = np.zeros((T_values.size, 5, Nt))
    for k, T in enumerate(T_values):
        lens[k], tpow[k] = time_table(T, dt, Nt)
    return T_values, lens, tpow


def distinct_durations(sampling: np.ndarray) -> np.ndarray:
    """Distinct values of column 1 (t1) of a sampling matrix.  Sampling matrices are cartesian
    products with t1 the slowest axis (sampling_matrix.py:85-121), so the column is piecewise
    constant and a change-point scan beats a sort; falls back to np.unique otherwise."""
    col = sampling[:, 1]
    change = np.flatnonzero(col[1:] != col[:-1])
    vals = np.concatenate((col[:1], col[change + 1]))
    if vals.size > 64:
        vals = np.unique(col)
    else:
        vals = np.unique(vals)
    return vals


# --------------------------------------------------------------------------------------------
# predictions (prediction_helpers.py:164-170,256-257 dict format) -> dense arrays
# --------------------------------------------------------------------------------------------
def pack_predictions(predictions, obstacle_order: Optional[Sequence] = None):
    """`predictions`: dict {id: {...}} or list of dicts with 'pos_list', 'cov_list',
    'orientation_list', 'shape'.  `obstacle_order`: ids in ``scenario.obstacles`` order
    (collision_check.py:127-131 iterates the scenario, not the dict)."""
    if predictions is None:
        return None
    if isinstance(predictions, dict):
        ids = list(predictions.keys()) if obstacle_order is None else [i for i in obstacle_order if i in predictions]
        # ids that are predicted but not in the scenario order still feed the prediction cost
        ids += [i for i in predictions.keys() if i not in ids]
        plist = [predictions[i] for i in ids]
    else:
        plist = list(predictions)
    O = len(plist)
    if O == 0:
        return None
    T = max(len(p["pos_list"]) for p in plist)
    pos = np.zeros((O, T, 2)); cov = np.tile(np.eye(2), (O, T, 1, 1)); theta = np.zeros((O, T))
    hl = np.zeros(O); hw = np.zeros(O); lens = np.zeros(O, dtype=np.int32)
    for o, p in enumerate(plist):
        n = len(p["pos_list"])
        lens[o] = n
        pos[o, :n] = np.asarray(p["pos_list"], dtype=np.float64)
        cov[o, :n] = np.asarray(p["cov_list"], dtype=np.float64)
        theta[o, :n] = np.asarray(p["orientation_list"], dtype=np.float64)[:n]
        hl[o] = p["shape"]["length"] / 2
        hw[o] = p["shape"]["width"] / 2
    return pos, cov, theta, hl, hw, lens


@dataclass
class PlanOutput:
    """What one device plan returns to the host layer (everything else stays in HBM)."""
    argmin: int
    min_cost: float
    n_rows: int
    n_in_list: int
    n_feasible: int
    n_candidates: int
    n_collide: int
    n_boundary: int
    collision_counter: int
    reason_counts: np.ndarray
    eval_kernel_ms: float
    total_device_ms: float

    @classmethod
    def from_result(cls, r: "_capi.FrxResult"):
        return cls(int(r.argmin), float(r.min_cost), int(r.n_rows), int(r.n_in_list), int(r.n_feasible),
                   int(r.n_candidates), int(r.n_collide), int(r.n_boundary), int(r.collision_counter),
                   np.array(list(r.reason_counts), dtype=np.int64), float(r.eval_kernel_ms), float(r.total_device_ms))


def active_costs(cost_weights: dict):
    """cost_function.py:55-60: zero weights dropped, names sorted."""
    names = [k for k, w in cost_weights.items() if w != 0]
    names.sort()
    return names, [float(cost_weights[n]) for n in names]
