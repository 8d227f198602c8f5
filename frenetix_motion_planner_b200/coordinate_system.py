"""Reference-path tables consumed by the device hot path.

Mirrors the attribute surface of the reference's ``CoordinateSystem``
(cr_scenario_handler/utils/utils_coordinate_system.py:187-274): ``reference``, ``ref_pos``,
``ref_theta``, ``ref_curv``, ``ref_curv_d``, ``convert_to_cartesian_coords``.

The three ``compute_*_from_polyline`` helpers the reference calls live in the un-vendored
commonroad-drivability-checker (``commonroad_dc.geometry.util``); the definitions below are the
ones SURVEY.md A.0 fixes (cumulative chord length; heading of the outgoing segment, last vertex
repeats; index-based ``np.gradient`` curvature).  ``ref_curv_d = np.gradient(curv, pos)`` is the
reference's own line (utils_coordinate_system.py:206).  Callers that own a real CCosy can pass its
arrays straight to :class:`CoordinateSystem.from_tables` instead.
"""
from __future__ import annotations

import numpy as np


def compute_pathlength_from_polyline(polyline: np.ndarray) -> np.ndarray:
    seg = np.sqrt(np.sum(np.diff(polyline, axis=0) ** 2, axis=1))
    return np.concatenate(([0.0], np.cumsum(seg)))


def compute_orientation_from_polyline(polyline: np.ndarray) -> np.ndarray:
    d = np.diff(polyline, axis=0)
    th = np.arctan2(d[:, 1], d[:, 0])
    return np.concatenate((th, th[-1:]))


def compute_curvature_from_polyline(polyline: np.ndarray) -> np.ndarray:
    x_d = np.gradient(polyline[:, 0])
    x_dd = np.gradient(x_d)
    y_d = np.gradient(polyline[:, 1])
    y_dd = np.gradient(y_d)
    return (x_d * y_dd - x_dd * y_d) / ((x_d ** 2 + y_d ** 2) ** 1.5)


class CoordinateSystem:
    """Host-side holder of the six reference tables (pos, theta, curv, curv_d, x, y)."""

    def __init__(self, reference: np.ndarray):
        reference = np.ascontiguousarray(reference, dtype=np.float64)
        if reference.ndim != 2 or reference.shape[1] != 2 or reference.shape[0] < 3:
            raise ValueError("reference path must be an [M>=3, 2] polyline")
        self._reference = reference
        self._ref_pos = compute_pathlength_from_polyline(reference)
        self._ref_curv = compute_curvature_from_polyline(reference)
        self._ref_theta = np.unwrap(compute_orientation_from_polyline(reference))
        self._ref_curv_d = np.gradient(self._ref_curv, self._ref_pos)
        self._ref_curv_dd = np.gradient(self._ref_curv_d, self._ref_pos)

    @classmethod
    def from_tables(cls, reference, ref_pos, ref_theta, ref_curv, ref_curv_d):
        self = cls.__new__(cls)
        self._reference = np.ascontiguousarray(reference, dtype=np.float64)
        self._ref_pos = np.ascontiguousarray(ref_pos, dtype=np.float64)
        self._ref_theta = np.ascontiguousarray(ref_theta, dtype=np.float64)
        self._ref_curv = np.ascontiguousarray(ref_curv, dtype=np.float64)
        self._ref_curv_d = np.ascontiguousarray(ref_curv_d, dtype=np.float64)
        self._ref_curv_dd = np.gradient(self._ref_curv_d, self._ref_pos)
        return self

    reference = property(lambda self: self._reference)
    ref_pos = property(lambda self: self._ref_pos)
    ref_theta = property(lambda self: self._ref_theta)
    ref_curv = property(lambda self: self._ref_curv)
    ref_curv_d = property(lambda self: self._ref_curv_d)
    ref_cruv_dd = property(lambda self: self._ref_curv_dd)  # (sic) name of the reference property

    def convert_to_cartesian_coords(self, s: float, d: float):
        """(s, d) -> (x, y) with the library's CCosy definition; ``None``
        outside the projection domain, like the reference wrapper (:262-270)."""
        p = self._ref_pos
        if not (s >= p[0]) or not (s < p[-1]):
            return None
        i = int(np.searchsorted(p, s, side="right")) - 1
        lam = (s - p[i]) / (p[i + 1] - p[i])
        px = (1.0 - lam) * self._reference[i, 0] + lam * self._reference[i + 1, 0]
        py = (1.0 - lam) * self._reference[i, 1] + lam * self._reference[i + 1, 1]
        th = self._ref_theta[i] + lam * (self._ref_theta[i + 1] - self._ref_theta[i])
        return np.array([px - d * np.sin(th), py + d * np.cos(th)])

    def convert_to_curvilinear_coords(self, x: float, y: float):
        """(x, y) -> (s, d): the EXACT inverse of :meth:`convert_to_cartesian_coords` (what pycrccosy's
        ``convert_to_curvilinear_coords`` is to its own forward map; planner.py:567-571 seeds ``x_cl`` with it).

        On segment i the forward map is P(lam) + d * n(theta(lam)) with a heading that is interpolated along the segment,
        so the foot point is where the offset vector is perpendicular to the INTERPOLATED tangent:
        f(lam) = (X - P(lam)) . t(theta(lam)) = 0.  Start from the orthogonal projection onto the closest segment and
        run Newton on f, stepping into the neighbour segment when lam leaves [0, 1)."""
        ref, pos, theta = self._reference, self._ref_pos, self._ref_theta
        X = np.array([x, y], dtype=np.float64)
        a, b = ref[:-1], ref[1:]
        ab = b - a
        lam0 = np.clip(np.sum((X - a) * ab, axis=1) / np.sum(ab * ab, axis=1), 0.0, 1.0)
        q = a + lam0[:, None] * ab
        i = int(np.argmin(np.sum((X - q) ** 2, axis=1)))
        lam = float(lam0[i])
        n_seg = ref.shape[0] - 1
        for _ in range(50):
            th = theta[i] + lam * (theta[i + 1] - theta[i])
            t = np.array([np.cos(th), np.sin(th)])
            n = np.array([-np.sin(th), np.cos(th)])
            r = X - (a[i] + lam * ab[i])
            f = float(np.dot(r, t))
            df = -float(np.dot(ab[i], t)) + float(np.dot(r, n)) * (theta[i + 1] - theta[i])
            step = f / df if df != 0.0 else 0.0
            lam_new = lam - step
            if lam_new < 0.0 and i > 0:
                i, lam = i - 1, 1.0 - 1e-12     # continue at the end of the previous segment
                continue
            if lam_new >= 1.0 and i < n_seg - 1:
                i, lam = i + 1, 0.0
                continue
            lam_new = min(max(lam_new, 0.0), 1.0)
            if abs(lam_new - lam) <= 1e-15:
                lam = lam_new
                break
            lam = lam_new
        th = theta[i] + lam * (theta[i + 1] - theta[i])
        n = np.array([-np.sin(th), np.cos(th)])
        s = pos[i] + lam * (pos[i + 1] - pos[i])
        d = float(np.dot(X - (a[i] + lam * ab[i]), n))
        return np.array([s, d])
