"""Reference-path preparation -- the step in front of the hot path.

``FrenetPlannerInterface.__init__`` takes the route planner's centre line, extends it by 30 m at both ends and smooths it
with a cubic spline before it reaches ``Planner.set_reference_and_coordinate_system``
(cr_scenario_handler/planner_interfaces/frenet_interface.py:100-114,
cr_scenario_handler/utils/utils_coordinate_system.py:20-58,110-134).  Same functions, same names, same results
(tests/test_host_logic.py replays outputs of the reference's own code, tests/golden/ref_refpath.npz) so that a caller who
only has a raw polyline gets the tables the reference would plan on.

``resample_polyline`` belongs to the un-vendored commonroad-drivability-checker (``commonroad_dc.geometry.util``); the
arc-length resampler below is the library's stand-in for it (fixed step along the polyline, the end point kept).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import splev, splprep


def resample_polyline(polyline: np.ndarray, step: float = 2.0) -> np.ndarray:
    polyline = np.asarray(polyline, dtype=np.float64)
    seg = np.sqrt(np.sum(np.diff(polyline, axis=0) ** 2, axis=1))
    arc = np.concatenate(([0.0], np.cumsum(seg)))
    n = int(np.floor(arc[-1] / step))
    target = np.arange(n + 1) * step
    if arc[-1] - n * step > 1e-9:
        target = np.concatenate((target, [arc[-1]]))
    return np.stack([np.interp(target, arc, polyline[:, 0]), np.interp(target, arc, polyline[:, 1])], axis=1)


def extend_path_linearly(points: np.ndarray, extension_length: float = 50, at_start: bool = True) -> np.ndarray:
    """Continue the first / last segment in a straight line, one new vertex per segment length."""
    points = np.asarray(points, dtype=np.float64)
    a, b = (points[0], points[1]) if at_start else (points[-2], points[-1])
    delta = b - a
    dist = float(np.sqrt(delta[0] ** 2 + delta[1] ** 2))
    if dist == 0:
        return points
    unit = delta / dist
    k = np.arange(1, int(extension_length / dist) + 1, dtype=np.float64)[:, None]
    if k.size == 0:
        return points
    if at_start:
        return np.vstack(((a - k * unit * dist)[::-1], points))
    return np.vstack((points, b + k * unit * dist))


def extend_ref_path_both_ends(ref_path: np.ndarray, extension_length: float = 30) -> np.ndarray:
    return extend_path_linearly(extend_path_linearly(ref_path, extension_length, at_start=True), extension_length,
                                at_start=False)


def _drop_duplicate_vertices(p: np.ndarray) -> np.ndarray:
    _, first = np.unique(p, axis=0, return_index=True)
    return p[np.sort(first)]


def smooth_ref_path(reference: np.ndarray, smoothing_interval: float = 4) -> np.ndarray:
    """Cubic interpolating spline through every (smoothing_interval / 0.125)-th vertex of the (densely sampled) path,
    evaluated at 6 points per metre and resampled to 1 m (utils_coordinate_system.py:110-134)."""
    reference = _drop_duplicate_vertices(np.asarray(reference, dtype=np.float64))
    # path length as the reference measures it: every other segment of the dense polyline, rounded to mm
    half = np.sqrt(np.sum((reference[0:-2:2] - reference[1:-1:2]) ** 2, axis=1))
    length = np.round(np.sum(half), 3)
    stride = int(smoothing_interval / 0.125)
    knots = reference[::stride]
    tck, u = splprep(knots.T, u=None, k=3, s=0.0)
    xs, ys = splev(np.linspace(u.min(), u.max(), int(6 * length)), tck, der=0)
    return _drop_duplicate_vertices(resample_polyline(np.array([xs, ys]).transpose(), 1))
