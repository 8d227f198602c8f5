"""Trajectory containers of the B200 planner: the reference's object surface over device tensors.

The reference materialises one ``TrajectorySample`` (+ two polynomial objects, a ``CartesianSample``
and a ``CurviLinearSample``) per candidate (frenetix_motion_planner/trajectories.py:56-478).  At
2e5 .. 1e7 candidates that would dwarf the kernel, so here the *bundle* owns the result of one device
plan (everything stays in HBM) and samples are index proxies created on demand; array attributes are
fetched lazily (``frx_get_states`` gather) and cached.  Duck-type compatible with what the
reference's logging / visualisation / interfaces read:

``.cartesian.{x,y,theta,v,a,kappa,kappa_dot}``, ``.curvilinear.{s,d,theta,s_dot,s_ddot,d_dot,d_ddot}``,
``.cost``, ``.costMap{name: (unweighted, weighted)}``, ``.feasible``, ``.valid``, ``.uniqueId``, ``.dt``,
``.horizon``, ``.sampling_parameters[13]``, ``.feasabilityMap``, ``.actual_traj_length``,
``.trajectory_long/.trajectory_lat`` (``.coeffs``, ``.delta_tau``, ``.squared_jerk_integral(t)``), and the
writable ``._ego_risk``, ``._obst_risk``, ``._coll_detected``, ``.boundary_harm``, ``.harm_occ_module``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi

_UNSET = object()
_PREFETCH = 256          # rows gathered at once when a list of samples is being iterated
# risk_assessment harm model LR1S "ignore_angle" (configurations/harm_parameters.json: log_reg.ignore_angle)
DEFAULT_HARM_COEFF = {"const": -4.591, "speed": 0.185}
_CART = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot")
_CURV = ("s", "d", "theta", "s_dot", "s_ddot", "d_dot", "d_ddot")   # state fields 7..13


class CartesianSample:
    """(x, y, theta, v, a, kappa, kappa_dot) arrays of one candidate (trajectories.py:56-198)."""

    def __init__(self, x, y, theta, v, a, kappa, kappa_dot, current_time_step: int):
        self.x, self.y, self.theta, self.v, self.a = x, y, theta, v, a
        self.kappa, self.kappa_dot = kappa, kappa_dot
        self.current_time_step = current_time_step

    def length(self) -> int:
        return len(self.x)


class CurviLinearSample:
    """(s, d, theta, s_dot, s_ddot, d_dot, d_ddot) arrays of one candidate (trajectories.py:200-335)."""

    def __init__(self, s, d, theta, current_time_step: int, dd=None, ddd=None, ss=None, sss=None):
        self.s, self.d, self.theta = s, d, theta
        self.d_dot, self.d_ddot, self.s_dot, self.s_ddot = dd, ddd, ss, sss
        self.current_time_step = current_time_step

    def length(self) -> int:
        return len(self.s)


class PolynomialView:
    """Stand-in for Quartic/QuinticTrajectory: coefficients via the reference's own linear solve
    (polynomial_trajectory.py:293-343,452-488), computed on first access (one sample, host)."""

    def __init__(self, kind: str, x_0: Sequence[float], x_d: Sequence[float], delta_tau: float):
        self.kind, self.x_0, self.x_d = kind, np.asarray(x_0, dtype=np.float64), np.asarray(x_d, dtype=np.float64)
        self.tau_0, self.delta_tau = 0, float(delta_tau)
        self._coeffs = None

    @property
    def coeffs(self) -> np.ndarray:
        if self._coeffs is None:
            T = self.delta_tau
            xs, vxs, axs = self.x_0
            if self.kind == "quartic":
                vxe = self.x_d[0]
                A = np.array([[3 * T ** 2, 4 * T ** 3], [6 * T, 12 * T ** 2]])
                b = np.array([vxe - vxs - axs * T, -axs])
                x = np.linalg.solve(A, b)
                self._coeffs = np.array([xs, vxs, axs / 2.0, x[0], x[1], 0.0])
            else:
                xe, vxe, axe = self.x_d
                A = np.array([[T ** 3., T ** 4, T ** 5], [3. * T ** 2, 4. * T ** 3, 5. * T ** 4],
                              [6. * T, 12. * T ** 2, 20. * T ** 3]])
                b = np.array([xe - xs - vxs * T - .5 * axs * T ** 2, vxe - vxs - axs * T, axe - axs])
                x = np.linalg.solve(A, b)
                self._coeffs = np.array([xs, vxs, .5 * axs, x[0], x[1], x[2]])
        return self._coeffs

    def squared_jerk_integral(self, t: float) -> float:
        c = self.coeffs
        t2 = t * t; t3 = t2 * t; t4 = t3 * t; t5 = t4 * t
        return (36 * c[3] * c[3] * t + 144 * c[3] * c[4] * t2 + 240 * c[3] * c[5] * t3 + 192 * c[4] * c[4] * t3 +
                720 * c[4] * c[5] * t4 + 720 * c[5] * c[5] * t5)


class StaleBundleError(RuntimeError):
    """A lazy view was asked for data of a plan whose device buffers a later plan has recycled."""


class TrajectorySample:
    """Index proxy of row `row` of a :class:`TrajectoryBundle` (attribute surface: module docstring).

    Ownership: in the reference every sample owns its arrays.  Here the arrays stay in HBM until somebody asks, and the
    next plan on the same handler recycles them.  :meth:`detach` pulls everything the sample can still be asked for to
    the host (the planner does that for the selected trajectory before it returns it); a sample that was not detached
    raises :class:`StaleBundleError` instead of silently showing the newer plan's numbers."""

    def __init__(self, bundle: "TrajectoryBundle", row: int):
        self._b, self._row = bundle, int(row)
        self.uniqueId = int(row) + bundle.row_base
        self.horizon, self.dt = bundle.horizon, bundle.dt
        self._ego_risk = self._obst_risk = None
        self._harm_override = _UNSET
        self.harm_occ_module = None
        self._states = None
        self._coll_override = None
        self._own = None           # detached copy: (flags, cost, cost row, traj_len, sampling row)

    def detach(self) -> "TrajectorySample":
        """Make the sample self-contained (host copies of its states and scalars): safe to keep across later plans."""
        if self._own is None:
            self._fetch()
            b, r = self._b, self._row
            if b.winner_row is not None and r == b.winner_row and b._flags is None and hasattr(b._h, "winner_record"):
                # the selected candidate's scalars came back with the arg-min: no read-back of whole arrays for one row
                b._live()
                fl, tl, tot, costs = b._h.winner_record()
                self._own = (fl, tot, np.array(costs, dtype=np.float64), tl, b.sampling_row(r))
            else:
                self._own = (int(b.flags[r]), float(b.total[r]), np.array(b.costs[r], dtype=np.float64), int(b.traj_len[r]),
                             b.sampling_row(r))
        return self

    # ---- scalars --------------------------------------------------------------------------
    @property
    def _flags(self) -> int:
        return self._own[0] if self._own is not None else int(self._b.flags[self._row])

    @property
    def feasible(self) -> bool:
        return bool(self._flags & _capi.FLAG_FEASIBLE)

    @property
    def valid(self) -> bool:
        return bool(self._flags & _capi.FLAG_VALID)

    @property
    def cost(self) -> float:
        return self._own[1] if self._own is not None else float(self._b.total[self._row])

    @property
    def _cost_row(self) -> np.ndarray:
        return self._own[2] if self._own is not None else self._b.costs[self._row]

    @property
    def costMap(self) -> Dict[str, tuple]:
        c = self._cost_row
        return {n: (float(c[k]), float(self._b.weights[k] * c[k])) for k, n in enumerate(self._b.cost_names)}

    @property
    def cost_list(self) -> list:
        return [float(v) for v in self._cost_row]

    @property
    def _coll_detected(self):
        return bool(self._flags & _capi.FLAG_COLLIDE) if self._coll_override is None else self._coll_override

    @_coll_detected.setter
    def _coll_detected(self, v):
        self._coll_override = v

    @property
    def feasabilityMap(self) -> Dict[str, float]:
        f = self._flags
        return {"Yaw rate Constraint": float(bool(f & _capi.flag_reason(6))),
                "Acceleration Constraint": float(bool(f & (_capi.flag_reason(8) | _capi.flag_reason(1)))),
                "Curvature Constraint": float(bool(f & _capi.flag_reason(5))),
                "Curvature Rate Constraint": float(bool(f & _capi.flag_reason(7)))}

    @property
    def actual_traj_length(self) -> int:
        return self._own[3] if self._own is not None else int(self._b.traj_len[self._row])

    @property
    def sampling_parameters(self) -> np.ndarray:
        return self._own[4] if self._own is not None else self._b.sampling_row(self._row)

    @property
    def boundary_harm(self):
        """planner.py:362-381: MAIS3+ probability (logistic regression "ignore_angle", harm_parameters.json) at the velocity
        of the step where the ego first overlaps the road boundary, 0 when it never does.  A value assigned by a caller
        (the reference writes the attribute) takes precedence."""
        if self._harm_override is not _UNSET:
            return self._harm_override
        f = self._flags
        if not f & _capi.FLAG_BOUNDARY:
            return 0
        k = (f >> _capi.FLAG_BOUNDARY_STEP_SHIFT) & 63
        v = float(self._fetch()[3][k])
        c = self._b.harm_coeff
        return float(1.0 / (1.0 + np.exp(-c["const"] - c["speed"] * v)))

    @boundary_harm.setter
    def boundary_harm(self, value):
        self._harm_override = value

    @property
    def first_collision_step(self) -> int:
        """Index of the first ego hull that meets a predicted obstacle (time index t0 + k), -1 when there is none."""
        f = self._flags
        return int((f >> _capi.FLAG_COLLIDE_STEP_SHIFT) & 63) if f & _capi.FLAG_COLLIDE else -1

    # ---- polynomials ----------------------------------------------------------------------
    @property
    def trajectory_long(self) -> PolynomialView:
        p = self.sampling_parameters
        return PolynomialView("quartic", p[2:5], [p[5], 0.0], p[1])

    @property
    def trajectory_lat(self) -> PolynomialView:
        p = self.sampling_parameters
        tau = p[1]
        if self._b.low_vel_mode:       # reactive_planner.py:161-166
            lon = self.trajectory_long
            c = lon.coeffs
            t = p[1]
            goal = (c[0] + c[1] * t + c[2] * t ** 2 + c[3] * t ** 3 + c[4] * t ** 4) - p[2]
            tau = t if goal <= 0 else goal
        return PolynomialView("quintic", p[7:10], p[10:13], tau)

    # ---- arrays (lazy gather from HBM) ------------------------------------------------------
    def _fetch(self) -> np.ndarray:
        if self._states is None:
            self._states = self._b.states_of(self._row)
        return self._states

    @property
    def cartesian(self) -> CartesianSample:
        st = self._fetch()
        return CartesianSample(*[st[k] for k in range(7)], current_time_step=self.actual_traj_length)

    @property
    def curvilinear(self) -> CurviLinearSample:
        st = self._fetch()
        return CurviLinearSample(st[7], st[8], st[9], self.actual_traj_length, dd=st[12], ddd=st[13], ss=st[10], sss=st[11])

    def length(self) -> int:
        return self._b.Nt

    def __repr__(self):
        return f"TrajectorySample(uniqueId={self.uniqueId}, cost={self.cost:.6g}, feasible={self.feasible}, valid={self.valid})"


class _SampleList:
    """Lazy, list-like view (len / index / slice / iterate) over a set of rows."""

    def __init__(self, bundle: "TrajectoryBundle", rows: np.ndarray):
        self._b, self._rows = bundle, np.asarray(rows, dtype=np.int64)

    def __len__(self):
        return int(self._rows.size)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return _SampleList(self._b, self._rows[i])
        return self._b.sample(int(self._rows[i]))

    def __iter__(self):
        # sinks that walk the whole list and read the state arrays of every sample (logging_helpers.py:275-294,580-647,
        # visualisation) would otherwise issue one device gather per sample: while iterating, a state miss fetches the
        # block of rows around it in one gather
        rows = self._rows
        for k0 in range(0, rows.size, _PREFETCH):
            block = rows[k0:k0 + _PREFETCH]
            for r in block:
                self._b._prefetch_hint = block
                yield self._b.sample(int(r))
        self._b._prefetch_hint = None

    def __bool__(self):
        return self._rows.size > 0

    @property
    def rows(self) -> np.ndarray:
        return self._rows


class TrajectoryBundle:
    """Result of one device plan.  ``trajectories`` / ``get_sorted_list()`` mirror
    trajectories.py:480-602 (stable cost order, ties keep generation order)."""

    def __init__(self, handler: "_capi.Handler", n_rows: int, cost_names: List[str], weights: List[float], dt: float,
                 horizon: float, Nt: int, low_vel_mode: bool, sampling: Optional[np.ndarray] = None, grid=None,
                 row_first: int = 0, row_base: int = 0, keep_all: bool = True):
        self._h, self.n_rows = handler, int(n_rows)
        self.cost_names, self.weights = list(cost_names), np.asarray(weights, dtype=np.float64)
        self.dt, self.horizon, self.Nt, self.low_vel_mode = dt, horizon, Nt, bool(low_vel_mode)
        self._sampling, self._grid, self._row_first, self.row_base = sampling, grid, int(row_first), int(row_base)
        self._flags = self._traj_len = self._costs = self._total = None
        self._order = None
        self._members: Optional[np.ndarray] = None
        self._cache: Dict[int, TrajectorySample] = {}
        self._is_sorted = False
        self.winner_row: Optional[int] = None     # local row of the last plan's arg-min (set by the planner)
        self._gen = getattr(handler, "generation", 0)     # the plan these views belong to
        self._prefetch_hint = None                        # rows of the block a list iteration is currently in
        self._state_cache: Dict[int, np.ndarray] = {}
        self.harm_coeff = dict(DEFAULT_HARM_COEFF)

    def _live(self):
        if getattr(self._h, "generation", self._gen) != self._gen:
            raise StaleBundleError("this TrajectoryBundle belongs to an earlier plan: the handler has planned again and "
                                   "recycled the device buffers (read what you need, or detach() samples, before the next plan())")

    # ---- bulk, lazily read back ------------------------------------------------------------
    def _read_flags(self):
        if self._flags is None:
            self._live()
            self._flags, self._traj_len = self._h.get_flags(0, self.n_rows)

    def _read_costs(self):
        if self._total is None:
            self._live()
            self._costs, self._total = self._h.get_costs(0, self.n_rows)

    @property
    def flags(self) -> np.ndarray:
        self._read_flags()
        return self._flags

    @property
    def traj_len(self) -> np.ndarray:
        self._read_flags()
        return self._traj_len

    @property
    def costs(self) -> np.ndarray:
        self._read_costs()
        return self._costs

    @property
    def total(self) -> np.ndarray:
        self._read_costs()
        return self._total

    def states_of(self, row: int) -> np.ndarray:
        hit = self._state_cache.pop(int(row), None)
        if hit is not None:
            return hit
        self._live()
        if self.winner_row is not None and int(row) == self.winner_row:
            # the selected candidate's rows came back with the arg-min (mapped result record): no device round trip
            return self._h.winner_states()
        hint = self._prefetch_hint
        if hint is not None and hint.size > 1 and int(row) in hint:
            st = self._h.get_states(np.asarray(hint, dtype=np.int64))
            for k, r in enumerate(hint.tolist()):
                if r != int(row):
                    self._state_cache[r] = st[:, k, :]
            self._prefetch_hint = None            # one gather per block
            return st[:, hint.tolist().index(int(row)), :]
        return self._h.get_states(np.array([row], dtype=np.int64))[:, 0, :]

    def states(self, rows, fields=None) -> np.ndarray:
        """[n_fields, len(rows), Nt] gather for many rows at once (logging / visualisation)."""
        self._live()
        return self._h.get_states(np.asarray(rows, dtype=np.int64), fields)

    def sampling_row(self, row: int) -> np.ndarray:
        if self._sampling is not None:
            return np.array(self._sampling[row], dtype=np.float64)
        t1, v1, d1, x_cl = self._grid
        g = self._row_first + row
        it, rem = divmod(g, len(v1) * len(d1))
        iv, idd = divmod(rem, len(d1))
        (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
        return np.array([0.0, t1[it], s0, ss0, sss0, v1[iv], 0.0, d0, dd0, ddd0, d1[idd], 0.0, 0.0])

    def sampling_rows(self, rows) -> np.ndarray:
        """[len(rows), 13] sampling rows, vectorised (grid mode computes them from the axes)."""
        rows = np.asarray(rows, dtype=np.int64)
        if self._sampling is not None:
            return np.array(self._sampling[rows], dtype=np.float64)
        t1, v1, d1, x_cl = self._grid
        g = self._row_first + rows
        it, rem = np.divmod(g, len(v1) * len(d1))
        iv, idd = np.divmod(rem, len(d1))
        (s0, ss0, sss0), (d0, dd0, ddd0) = x_cl
        out = np.zeros((rows.size, 13))
        out[:, 1], out[:, 5], out[:, 10] = np.asarray(t1)[it], np.asarray(v1)[iv], np.asarray(d1)[idd]
        out[:, 2], out[:, 3], out[:, 4], out[:, 7], out[:, 8], out[:, 9] = s0, ss0, sss0, d0, dd0, ddd0
        return out

    # ---- list semantics of the reference ---------------------------------------------------
    def sample(self, row: int) -> TrajectorySample:
        s = self._cache.get(row)
        if s is None:
            s = self._cache[row] = TrajectorySample(self, row)
        return s

    def restrict(self, mask_or_rows):
        """``bundle.trajectories = [...]`` of the reference: choose which rows are members."""
        a = np.asarray(mask_or_rows)
        self._members = np.flatnonzero(a) if a.dtype == bool else a.astype(np.int64)
        self._order = None
        self._is_sorted = False

    @property
    def member_rows(self) -> np.ndarray:
        if self._members is None:
            self._members = np.flatnonzero((self.flags & _capi.FLAG_IN_LIST) != 0)
        return self._members

    @property
    def trajectories(self) -> _SampleList:
        return _SampleList(self, self.member_rows if self._order is None else self._order)

    @trajectories.setter
    def trajectories(self, rows):
        self.restrict(rows.rows if isinstance(rows, _SampleList) else rows)

    def sort(self, occlusion_module=None):
        """Stable sort of the members by total cost (trajectories.py:524-561); costs are already on the
        device result, nothing is recomputed."""
        if not self._is_sorted:
            m = self.member_rows
            self._order = m[np.argsort(self.total[m], kind="stable")]
            self._is_sorted = True

    def get_sorted_list(self, occlusion_module=None) -> _SampleList:
        self.sort(occlusion_module)
        return _SampleList(self, self._order)

    def min_costs(self):
        return self.sample(int(self._order[0])) if self._is_sorted and len(self._order) else None

    def max_costs(self):
        return self.sample(int(self._order[-1])) if self._is_sorted and len(self._order) else None

    @property
    def empty(self) -> bool:
        return self.member_rows.size == 0
