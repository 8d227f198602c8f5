"""Sampling level sets and the [N x 13] sampling matrix -- host side, once per planning step.

API-compatible with the reference's ``frenetix_motion_planner/sampling_matrix.py``
(``SamplingHandler``, ``TimeSampling`` / ``VelocitySampling`` / ``LateralPositionSampling`` /
``LongitudinalPositionSampling`` with ``to_range(level) -> set``, ``generate_sampling_matrix``), so
planner code written against the reference keeps working; tests/test_host_logic.py replays the
reference's own outputs (tests/golden/ref_sampling.npz).

Additions for the device path: :func:`sampling_axes` returns the three 1-D axes (t1, ss1, d1) of a
level so that large grids can be expanded on the GPU (``frx_plan_grid``) instead of on the host.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

# column order of a sampling-matrix row (sampling_matrix.py:85-121 of the reference)
COLUMNS = ("t0", "t1", "s0", "ss0", "sss0", "ss1", "sss1", "d0", "dd0", "ddd0", "d1", "dd1", "ddd1")


class _LevelSets:
    """Base of the per-quantity samplers: level i holds a set of values, denser with i."""

    def __init__(self, minimum: float, maximum: float, max_density: int):
        if not maximum >= minimum:
            raise AssertionError("sampling interval is empty")
        if not (isinstance(max_density, (int, np.integer)) and max_density > 0):
            raise AssertionError("max_density must be a positive integer")
        self.minimum, self.maximum, self.max_density = minimum, maximum, int(max_density)
        self._levels = [self._level(i) for i in range(self.max_density)]

    def _level(self, i: int) -> set:
        # 3, 5, 9, 17, ... equidistant values: n_{i+1} = 2 n_i - 1
        n = 2 ** (i + 1) + 1
        return set(np.linspace(self.minimum, self.maximum, n))

    def to_range(self, sampling_stage: int = 0) -> set:
        if not 0 <= sampling_stage < self.max_density:
            raise AssertionError(f"sampling stage {sampling_stage} outside [0, {self.max_density})")
        return self._levels[sampling_stage]


class VelocitySampling(_LevelSets):
    pass


class LateralPositionSampling(_LevelSets):
    pass


class LongitudinalPositionSampling(_LevelSets):
    def __init__(self, maximum: float, minimum: float, density: int):   # (sic) argument order of the reference
        super().__init__(maximum, minimum, density)


class TimeSampling(_LevelSets):
    """Durations on the dt raster: level i steps by int((1 / (i + 1)) / dt) * dt, rounded to 2 decimals."""

    def __init__(self, minimum: float, maximum: float, density: int, dT: float):
        self.dT = dT
        super().__init__(minimum, maximum, density)

    def _level(self, i: int) -> set:
        step = int((1 / (i + 1)) / self.dT)
        return set(np.round(np.arange(self.minimum, self.maximum + self.dT, step * self.dT), 2))


class SamplingHandler:
    def __init__(self, dt: float, max_sampling_number: int, t_min: float, horizon: float, delta_d_min: float,
                 delta_d_max: float, d_ego_pos: bool):
        self.dt = dt
        self.max_sampling_number = max_sampling_number
        self.s_sampling_mode = False
        self.d_ego_pos = d_ego_pos
        self.t_min, self.horizon = t_min, horizon
        self.delta_d_min, self.delta_d_max = delta_d_min, delta_d_max
        self.t_sampling = self.d_sampling = self.v_sampling = self.s_sampling = None
        self.set_t_sampling()
        if not self.d_ego_pos:
            self.set_d_sampling()

    def update_static_params(self, t_min: float, horizon: float, delta_d_min: float, delta_d_max: float):
        assert t_min > 0, "t_min cant be <= 0"
        self.t_min, self.horizon, self.delta_d_min, self.delta_d_max = t_min, horizon, delta_d_min, delta_d_max
        self.set_t_sampling()
        self.set_d_sampling()

    def change_max_sampling_level(self, max_samp_lvl):
        self.max_sampling_number = max_samp_lvl

    def set_t_sampling(self):
        self.t_sampling = TimeSampling(self.t_min, self.horizon, self.max_sampling_number, self.dt)

    def set_d_sampling(self, lat_pos=None):
        lo, hi = self.delta_d_min, self.delta_d_max
        if self.d_ego_pos:
            lo, hi = lat_pos + lo, lat_pos + hi
        self.d_sampling = LateralPositionSampling(lo, hi, self.max_sampling_number)

    def set_v_sampling(self, v_min, v_max):
        self.v_sampling = VelocitySampling(v_min, v_max, self.max_sampling_number)

    def set_s_sampling(self, delta_s_min, delta_s_max):
        self.s_sampling = LongitudinalPositionSampling(delta_s_min, delta_s_max, self.max_sampling_number)


def generate_sampling_matrix(*, t0_range, t1_range, s0_range, ss0_range, sss0_range, ss1_range, sss1_range, d0_range,
                             dd0_range, ddd0_range, d1_range, dd1_range, ddd1_range) -> np.ndarray:
    """Cartesian product of the 13 ranges, one row per combination, first argument slowest.
    Vectorised (np.meshgrid) instead of itertools.product over Python tuples; same rows, same order."""
    axes = [np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in
            (t0_range, t1_range, s0_range, ss0_range, sss0_range, ss1_range, sss1_range, d0_range, dd0_range,
             ddd0_range, d1_range, dd1_range, ddd1_range)]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def python_path_rows(t_set: Iterable[float], v_set: Iterable[float], d_set: Iterable[float], x_cl) -> np.ndarray:
    """Rows in the generation order of the reference's Python path (reactive_planner.py:149-175):
    ``for t: for v: for d`` over the *sets* in their native iteration order, so that the row index
    equals the reference's ``uniqueId``.  Vectorised: one repeat / tile per axis, no Python loop over rows."""
    (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
    t = np.fromiter(t_set, dtype=np.float64)
    v = np.fromiter(v_set, dtype=np.float64)
    d = np.fromiter(d_set, dtype=np.float64)
    n = t.size * v.size * d.size
    S = np.zeros((n, 13), dtype=np.float64)
    S[:, 1] = np.repeat(t, v.size * d.size)
    S[:, 2], S[:, 3], S[:, 4] = s0, ss0, sss0
    S[:, 5] = np.tile(np.repeat(v, d.size), t.size)
    S[:, 7], S[:, 8], S[:, 9] = d0, dd0, ddd0
    S[:, 10] = np.tile(d, t.size * v.size)
    return S


def sampling_axes(handler: SamplingHandler, level: int, x_cl, cpp_style: bool = False):
    """The three axes of sampling level `level`, as the very set objects the reference iterates.

    Iteration order of a Python set depends on its hash-table size, and a *copy* of a set can land in a different table
    size than the original -- so the sets are handed out exactly as the reference builds them: the Python path iterates
    ``t_sampling.to_range(level)`` and ``v_sampling.to_range(level)`` themselves and one ``.union({d0})`` copy of the d
    set (reactive_planner.py:149-158); the C++ path takes one ``.union`` of each (reactive_planner_cpp.py:235-237:
    t with {N*dT}, v with {ss0}, d with {d0}).  Row index == the reference's ``uniqueId`` depends on it."""
    (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
    if cpp_style:
        t_set = handler.t_sampling.to_range(level).union({round(handler.horizon / handler.dt) * handler.dt})
        v_set = handler.v_sampling.to_range(level).union({ss0})
    else:
        t_set = handler.t_sampling.to_range(level)
        v_set = handler.v_sampling.to_range(level)
    d_set = handler.d_sampling.to_range(level).union({d0})
    return t_set, v_set, d_set
