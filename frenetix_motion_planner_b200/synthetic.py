"""Deterministic synthetic workloads (BASELINE.json configs 2, 3, 5 shapes).

Pure numpy, no device code: reference polylines, dense (t, v, d) grids, predicted-obstacle sets.
Used by ``bench.py`` and the tests so that every leg (CUDA, oracle, CPU baseline) sees the
same inputs.
"""
from __future__ import annotations

import numpy as np

# BMW 320i (commonroad-vehicle-models id 2); explicit inputs everywhere, see SURVEY.md 8c
VEHICLE_2 = dict(length=4.508, width=1.610, wheelbase=2.5789, wb_rear_axle=1.4227,
                 a_max=11.5, v_max=50.8, v_switch=7.319, delta_max=1.066, v_delta_max=0.4)

DEFAULT_COST_WEIGHTS = {"lateral_jerk": 0.2, "longitudinal_jerk": 0.2, "velocity_offset": 1.0,
                        "distance_to_reference_path": 5.0, "prediction": 0.2}


def straight_polyline(M: int = 400, spacing: float = 1.0) -> np.ndarray:
    x = np.arange(M, dtype=np.float64) * spacing
    return np.stack([x, np.zeros(M)], axis=1)


def arc_polyline(R: float = 200.0, M: int = 600, spacing: float = 1.0, start_heading: float = 0.0) -> np.ndarray:
    """Constant-curvature (left turn) polyline with ~`spacing` m between vertices."""
    phi = np.arange(M, dtype=np.float64) * (spacing / R)
    x = R * np.sin(phi)
    y = R * (1.0 - np.cos(phi))
    c, s = np.cos(start_heading), np.sin(start_heading)
    return np.stack([c * x - s * y, s * x + c * y], axis=1)


def scurve_polyline(M: int = 300, spacing: float = 1.0, amp: float = 6.0, wavelength: float = 120.0) -> np.ndarray:
    """Gentle S-curve with varying curvature (exercises curv / curv_d interpolation)."""
    xs = np.linspace(0.0, (M - 1) * spacing * 0.98, 8 * M)
    ys = amp * np.sin(2 * np.pi * xs / wavelength)
    pts = np.stack([xs, ys], axis=1)
    seg = np.sqrt(np.sum(np.diff(pts, axis=0) ** 2, axis=1))
    L = np.concatenate(([0.0], np.cumsum(seg)))
    target = np.arange(M, dtype=np.float64) * spacing
    target = target[target <= L[-1]]
    return np.stack([np.interp(target, L, pts[:, 0]), np.interp(target, L, pts[:, 1])], axis=1)


def velocity_interval(v: float, a_max: float, horizon: float, v_max: float, v_limit: float = 36.0):
    """frenetix_motion_planner/planner.py:304-306."""
    min_v = max(0.001, v - a_max * horizon)
    max_v = min(min(v + (a_max / 6.0) * horizon, v_limit), v_max)
    return min_v, max_v


def grid_sampling_matrix(t1_range, ss1_range, d1_range, x_cl) -> np.ndarray:
    """Cartesian product in the row order of ``generate_sampling_matrix``
    (frenetix_motion_planner/sampling_matrix.py:85-121): t1 slowest, then ss1, then d1."""
    t1 = np.asarray(t1_range, dtype=np.float64)
    v1 = np.asarray(ss1_range, dtype=np.float64)
    d1 = np.asarray(d1_range, dtype=np.float64)
    (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
    n = t1.size * v1.size * d1.size
    S = np.zeros((n, 13), dtype=np.float64)
    S[:, 1] = np.repeat(t1, v1.size * d1.size)
    S[:, 2], S[:, 3], S[:, 4] = s0, ss0, sss0
    S[:, 5] = np.tile(np.repeat(v1, d1.size), t1.size)
    S[:, 7], S[:, 8], S[:, 9] = d0, dd0, ddd0
    S[:, 10] = np.tile(d1, t1.size * v1.size)
    return S


def time_range(t_min: float, horizon: float, dt: float, n: int) -> np.ndarray:
    """`n` two-decimal durations between t_min and horizon on the dt raster (like TimeSampling)."""
    k0, k1 = int(round(t_min / dt)), int(round(horizon / dt))
    ks = np.unique(np.round(np.linspace(k0, k1, n)).astype(int))
    return np.round(ks * dt, 2)


def synthetic_predictions(polyline: np.ndarray, n_obstacles: int, T: int, dt: float, seed: int,
                          lateral_spread: float = 6.0, s_lo: float = 15.0, s_hi: float = 80.0,
                          length: float = 5.5, width: float = 2.2):
    """`n_obstacles` predicted cars moving roughly along the path with a lateral offset.
    Covariance grows per step and is slightly rotated so that the inverse
    is a full 2x2 matrix.  Returns the reference's prediction dict format
    (cr_scenario_handler/utils/prediction_helpers.py:164-170,256-257) as a list."""
    rng = np.random.default_rng(seed)
    seg = np.sqrt(np.sum(np.diff(polyline, axis=0) ** 2, axis=1))
    L = np.concatenate(([0.0], np.cumsum(seg)))
    head = np.arctan2(np.diff(polyline[:, 1]), np.diff(polyline[:, 0]))
    head = np.concatenate((head, head[-1:]))
    preds = []
    for o in range(n_obstacles):
        s_start = rng.uniform(s_lo, max(s_lo + 1.0, min(s_hi, L[-1] - 60.0)))
        v = rng.uniform(0.0, 12.0)
        lat = rng.normal(0.0, lateral_spread)
        if abs(lat) < 1.2:      # keep a corridor so that not every candidate collides
            lat = np.sign(lat + 1e-9) * (1.2 + abs(lat))
        s = s_start + v * dt * np.arange(1, T + 1)
        px = np.interp(s, L, polyline[:, 0])
        py = np.interp(s, L, polyline[:, 1])
        th = np.interp(s, L, np.unwrap(head))
        pos = np.stack([px - lat * np.sin(th), py + lat * np.cos(th)], axis=1)
        ori = th + rng.normal(0.0, 0.05)
        cov = np.zeros((T, 2, 2))
        for k in range(T):
            sc = 0.1 * (1.0 + 0.05 * k)
            a = 0.3 * o + 0.01 * k
            R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            cov[k] = R @ np.diag([sc * 1.5, sc]) @ R.T
        preds.append({"pos_list": pos, "cov_list": cov, "orientation_list": ori,
                      "v_list": np.full(T, v), "shape": {"length": length, "width": width}})
    return preds
