"""ctypes binding of libfrx_b200.so (include/frx.h) -- the only door to the device hot path.

There is deliberately NO fallback: if the shared library is missing, or the machine has no CUDA
device, creating a :class:`Handler` raises.  ``python __graft_entry__.py`` (``build()``) compiles it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrx_b200.so")

FRX_MAX_COSTS = 10
EXCHANGE_PAGE_BYTES = 8192
NUM_FIELDS = 14
FIELDS = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot",
          "s", "d", "theta_cl", "s_dot", "s_ddot", "d_dot", "d_ddot")
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}

FLAG_VALID = 1 << 0
FLAG_FEASIBLE = 1 << 1
FLAG_COLLIDE = 1 << 12
FLAG_BOUNDARY = 1 << 13
FLAG_STORED = 1 << 14
FLAG_IN_LIST = 1 << 15
FLAG_COSTED = 1 << 16
FLAG_CANDIDATE = 1 << 17
FLAG_COLLIDE_STEP_SHIFT, FLAG_BOUNDARY_STEP_SHIFT = 18, 24     # 6-bit index of the first colliding ego hull (frx.h)


def flag_reason(r: int) -> int:
    return 1 << (1 + r)


COST_NAMES = ("acceleration", "distance_to_obstacles", "distance_to_reference_path", "jerk",
              "lateral_jerk", "longitudinal_jerk", "orientation_offset", "path_length",
              "prediction", "velocity_offset")
COST_ID = {n: i for i, n in enumerate(COST_NAMES)}
# terms the reference itself cannot evaluate in batch (lanelet lookups, reach sets, broken helpers,
# NotImplementedError stubs): partial_cost_functions.py:67-117,133-138,199-293,359-387
HOST_ONLY_COSTS = ("lane_center_offset", "velocity", "responsibility", "steering_angle", "steering_rate", "yaw",
                   "longitudinal_velocity_offset", "time", "inverse_duration")


class FrxParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("N", C.c_int32), ("low_vel_mode", C.c_int32), ("draw_traj_set", C.c_int32),
                ("kinematic_debug", C.c_int32),
                ("a_max", C.c_double), ("v_switch", C.c_double), ("delta_max", C.c_double), ("wheelbase", C.c_double),
                ("wb_rear_axle", C.c_double), ("length", C.c_double), ("width", C.c_double),
                ("x0_orientation", C.c_double), ("desired_velocity", C.c_double),
                ("n_costs", C.c_int32), ("cost_ids", C.c_int32 * FRX_MAX_COSTS),
                ("cost_weights", C.c_double * FRX_MAX_COSTS),
                ("store_states", C.c_int32), ("check_collisions", C.c_int32),
                ("curvature_rate_from_v_delta", C.c_int32), ("velocity_offset_norm", C.c_int32), ("v_delta_max", C.c_double),
                ("prediction_cost_mode", C.c_int32), ("reserved_", C.c_int32)]


class FrxResult(C.Structure):
    _fields_ = [("argmin", C.c_int64), ("min_cost", C.c_double), ("n_rows", C.c_int64), ("n_in_list", C.c_int64),
                ("n_feasible", C.c_int64), ("n_candidates", C.c_int64), ("n_collide", C.c_int64),
                ("n_boundary", C.c_int64), ("collision_counter", C.c_int64), ("reason_counts", C.c_int64 * 11),
                ("eval_kernel_ms", C.c_float), ("total_device_ms", C.c_float),
                ("obstacle_kernel_ms", C.c_float), ("reserved_", C.c_float)]


EXPORTS = ("frx_abi_version", "frx_create", "frx_destroy", "frx_last_error", "frx_set_reference", "frx_set_params",
           "frx_set_reference_polyline", "frx_get_reference", "frx_initial_state",
           "frx_set_time_tables", "frx_set_predictions", "frx_set_obstacle_positions", "frx_set_static_obbs",
           "frx_plan", "frx_plan_device", "frx_plan_device_async", "frx_plan_wait", "frx_plan_grid", "frx_plan_batched", "frx_state_pitch", "frx_last_launches", "frx_get_states",
           "frx_get_states_range", "frx_winner_states", "frx_winner_record", "frx_get_costs", "frx_get_flags", "frx_device_pointers", "frx_winner_device_pointer",
           "frx_selftest_fdiv", "frx_selftest_divc", "frx_selftest_fp64_peak", "frx_set_exchange", "frx_exchange_wait", "frx_set_stream",
           "frx_synchronize")

_lib = None


class FrxError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the library and declare the prototypes.  Raises FrxError if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FRX_LIB") or LIB_PATH     # FRX_LIB: tuning builds of the same ABI
    if not os.path.exists(p):
        raise FrxError(f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                       f"g.build()'); there is no CPU fallback")
    lib = C.CDLL(p)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    vp = C.c_void_p
    lib.frx_abi_version.restype = C.c_int
    lib.frx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.frx_destroy.argtypes = [vp]
    lib.frx_last_error.argtypes = [vp]; lib.frx_last_error.restype = C.c_char_p
    lib.frx_set_reference.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp, dp]
    lib.frx_set_params.argtypes = [vp, C.POINTER(FrxParams)]
    lib.frx_set_reference_polyline.argtypes = [vp, C.c_int32, dp]
    lib.frx_get_reference.argtypes = [vp, C.c_int32, dp]
    lib.frx_initial_state.argtypes = [vp, dp, C.c_int32, C.c_double, dp]
    lib.frx_set_time_tables.argtypes = [vp, C.c_int32, dp, ip, dp]
    lib.frx_set_predictions.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, dp, dp, ip]
    lib.frx_set_obstacle_positions.argtypes = [vp, C.c_int32, dp]
    lib.frx_set_static_obbs.argtypes = [vp, C.c_int32, dp]
    lib.frx_plan.argtypes = [vp, C.c_int64, dp, C.c_int64, C.POINTER(FrxResult)]
    lib.frx_plan_device.argtypes = [vp, C.c_int64, vp, C.c_int64, C.POINTER(FrxResult)]
    lib.frx_plan_device_async.argtypes = [vp, C.c_int64, vp, C.c_int64]
    lib.frx_plan_wait.argtypes = [vp, C.POINTER(FrxResult)]
    lib.frx_plan_grid.argtypes = [vp, C.c_int32, dp, C.c_int32, dp, C.c_int32, dp, dp, C.c_int64, C.c_int64,
                                  C.POINTER(FrxResult)]
    lib.frx_plan_batched.argtypes = [C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(dp), C.POINTER(FrxResult)]
    lib.frx_state_pitch.argtypes = [vp]; lib.frx_state_pitch.restype = C.c_int32
    lib.frx_last_launches.argtypes = [vp]; lib.frx_last_launches.restype = C.c_int32
    lib.frx_get_states.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64), C.c_uint32, dp]
    lib.frx_get_states_range.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint32, dp]
    lib.frx_winner_states.argtypes = [vp, C.c_uint32, dp]
    lib.frx_winner_record.argtypes = [vp, C.POINTER(C.c_uint32), ip, dp, dp]
    lib.frx_get_costs.argtypes = [vp, C.c_int64, C.c_int64, dp, dp]
    lib.frx_get_flags.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint32), ip]
    lib.frx_device_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.frx_winner_device_pointer.argtypes = [vp, C.POINTER(vp)]
    lib.frx_selftest_fdiv.argtypes = [vp, C.c_int64, dp, dp, dp, dp]
    lib.frx_selftest_divc.argtypes = [vp, C.c_int64, dp, C.c_double, dp, dp]
    lib.frx_selftest_fp64_peak.argtypes = [vp, dp]
    lib.frx_set_exchange.argtypes = [vp, vp, C.c_int32, C.c_int32]
    lib.frx_exchange_wait.argtypes = [vp, C.c_int64, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.frx_set_stream.argtypes = [vp, vp]
    lib.frx_synchronize.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("frx_last_error", "frx_state_pitch", "frx_last_launches", "frx_abi_version"):
            getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Handler:
    """Thin object wrapper of one ``frx_ctx`` (the counterpart of ``frenetix.TrajectoryHandler``,
    reactive_planner_cpp.py:49): owns the device buffers of one planner instance."""

    def __init__(self, device: int = 0, library=None):
        # `library`: another build of the same ABI loaded with load_library(path) (tuning builds; the benchmark's CPU arm
        # binds the host-core implementation of the ABI this way).  The package itself always uses libfrx_b200.so.
        self._lib = library if library is not None else load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.frx_create(int(device), C.byref(self._ctx))
        if rc != 0 or not self._ctx:
            raise FrxError(f"frx_create(device={device}) failed with code {rc}: no usable CUDA device "
                           f"(this library has no CPU fallback)")
        self.device = device
        self.n_rows = 0
        self.n_costs = 0
        self.Nt = 0
        self.generation = 0       # bumped by every plan*: lazy views of an older plan (TrajectoryBundle) detect that the
                                  # device buffers they point at have been recycled

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.frx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.frx_last_error(self._ctx)
            raise FrxError(f"libfrx_b200 error {rc}: {msg.decode() if msg else '?'}")

    # ---- set-up -----------------------------------------------------------------------------
    def set_reference(self, ref_pos, ref_theta, ref_curv, ref_curv_d, ref_x, ref_y):
        arrs = [_f64(a) for a in (ref_pos, ref_theta, ref_curv, ref_curv_d, ref_x, ref_y)]
        M = arrs[0].size
        if any(a.size != M for a in arrs):
            raise ValueError("reference tables must have equal length")
        self._check(self._lib.frx_set_reference(self._ctx, M, *[_dptr(a) for a in arrs]))

    def set_reference_polyline(self, polyline) -> np.ndarray:
        """Build the six reference tables on the device from an [M, 2] polyline; returns them ([6, M]: pos, theta, curv,
        curv_d, x, y) so that the host keeps a view of what the device plans on."""
        xy = _f64(polyline)
        if xy.ndim != 2 or xy.shape[1] != 2:
            raise ValueError("polyline must be [M, 2]")
        self._check(self._lib.frx_set_reference_polyline(self._ctx, xy.shape[0], _dptr(xy)))
        out = np.empty((6, xy.shape[0]), dtype=np.float64)
        self._check(self._lib.frx_get_reference(self._ctx, xy.shape[0], _dptr(out)))
        return out

    def initial_state(self, x, y, orientation, velocity, acceleration, steering_angle, low_vel_mode, wheelbase):
        """Planner._compute_initial_states on the device -> ([s, s', s''], [d, d', d''])."""
        x0 = _f64([x, y, orientation, velocity, acceleration, steering_angle])
        out = np.empty(6, dtype=np.float64)
        self._check(self._lib.frx_initial_state(self._ctx, _dptr(x0), int(bool(low_vel_mode)), float(wheelbase), _dptr(out)))
        return [float(v) for v in out[:3]], [float(v) for v in out[3:]]

    def set_params(self, *, dt, N, low_vel_mode, draw_traj_set, kinematic_debug, a_max, v_switch, delta_max,
                   wheelbase, wb_rear_axle, length, width, x0_orientation, desired_velocity,
                   cost_names: Sequence[str], cost_weights: Sequence[float], store_states=True,
                   check_collisions=True, curvature_rate_from_v_delta=False, v_delta_max=0.4, velocity_offset_norm=1,
                   prediction_cost_mode=0):
        p = FrxParams()
        p.dt, p.N = float(dt), int(N)
        p.low_vel_mode, p.draw_traj_set, p.kinematic_debug = int(bool(low_vel_mode)), int(bool(draw_traj_set)), int(bool(kinematic_debug))
        p.a_max, p.v_switch, p.delta_max, p.wheelbase = float(a_max), float(v_switch), float(delta_max), float(wheelbase)
        p.wb_rear_axle, p.length, p.width = float(wb_rear_axle), float(length), float(width)
        p.x0_orientation, p.desired_velocity = float(x0_orientation), float(desired_velocity)
        if len(cost_names) > FRX_MAX_COSTS:
            raise ValueError("too many cost terms")
        p.n_costs = len(cost_names)
        for k, (n, w) in enumerate(zip(cost_names, cost_weights)):
            if n not in COST_ID:
                raise NotImplementedError(f"cost term '{n}' cannot be evaluated on the device "
                                          f"({'host-only in the reference' if n in HOST_ONLY_COSTS else 'unknown'})")
            p.cost_ids[k] = COST_ID[n]
            p.cost_weights[k] = float(w)
        p.store_states, p.check_collisions = int(bool(store_states)), int(bool(check_collisions))
        p.curvature_rate_from_v_delta, p.v_delta_max = int(bool(curvature_rate_from_v_delta)), float(v_delta_max)
        p.velocity_offset_norm = int(velocity_offset_norm)
        p.prediction_cost_mode = int(prediction_cost_mode)
        self._check(self._lib.frx_set_params(self._ctx, C.byref(p)))
        self.n_costs = p.n_costs
        self.Nt = int(N) + 1

    def set_time_tables(self, T_values, traj_len, tpow):
        T_values = _f64(T_values)
        traj_len = np.ascontiguousarray(traj_len, dtype=np.int32)
        tpow = _f64(tpow)
        assert tpow.shape == (T_values.size, 5, self.Nt)
        self._check(self._lib.frx_set_time_tables(self._ctx, T_values.size, _dptr(T_values),
                                                  traj_len.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(tpow)))

    def set_predictions(self, pos, cov, theta, half_len, half_wid, len_valid):
        if pos is None or len(half_len) == 0:
            self._check(self._lib.frx_set_predictions(self._ctx, 0, 0, None, None, None, None, None, None))
            return
        pos, cov, theta = _f64(pos), _f64(cov), _f64(theta)
        half_len, half_wid = _f64(half_len), _f64(half_wid)
        len_valid = np.ascontiguousarray(len_valid, dtype=np.int32)
        O, T = theta.shape
        assert pos.shape == (O, T, 2) and cov.shape == (O, T, 2, 2)
        self._check(self._lib.frx_set_predictions(self._ctx, O, T, _dptr(pos), _dptr(cov), _dptr(theta),
                                                  _dptr(half_len), _dptr(half_wid),
                                                  len_valid.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_obstacle_positions(self, pos_xy):
        if pos_xy is None or len(pos_xy) == 0:
            self._check(self._lib.frx_set_obstacle_positions(self._ctx, 0, None))
            return
        pos_xy = _f64(pos_xy)
        self._check(self._lib.frx_set_obstacle_positions(self._ctx, pos_xy.shape[0], _dptr(pos_xy)))

    def set_static_obbs(self, obbs):
        if obbs is None or len(obbs) == 0:
            self._check(self._lib.frx_set_static_obbs(self._ctx, 0, None))
            return
        obbs = _f64(obbs)
        assert obbs.ndim == 2 and obbs.shape[1] == 5
        self._check(self._lib.frx_set_static_obbs(self._ctx, obbs.shape[0], _dptr(obbs)))

    def set_stream(self, cuda_stream_handle: int):
        self._check(self._lib.frx_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)))

    # ---- the hot path -----------------------------------------------------------------------
    def plan(self, sampling: np.ndarray, row_index_base: int = 0) -> FrxResult:
        """Host sampling matrix [N, 13] (H2D copy inside)."""
        if sampling.dtype != np.float64 or not sampling.flags.c_contiguous:
            sampling = _f64(sampling)
        if sampling.ndim != 2 or sampling.shape[1] != 13:
            raise ValueError("sampling matrix must be [N, 13]")
        res = FrxResult()
        self.generation += 1
        self._check(self._lib.frx_plan(self._ctx, sampling.shape[0], _dptr(sampling), int(row_index_base), C.byref(res)))
        self.n_rows = sampling.shape[0]
        return res

    def plan_device(self, device_ptr: int, n_rows: int, row_index_base: int = 0) -> FrxResult:
        res = FrxResult()
        self.generation += 1
        self._check(self._lib.frx_plan_device(self._ctx, int(n_rows), C.c_void_p(device_ptr), int(row_index_base), C.byref(res)))
        self.n_rows = int(n_rows)
        return res

    def plan_device_async(self, device_ptr: int, n_rows: int, row_index_base: int = 0) -> None:
        """Enqueue only; pair with :meth:`plan_wait`."""
        self.generation += 1
        self._check(self._lib.frx_plan_device_async(self._ctx, int(n_rows), C.c_void_p(device_ptr), int(row_index_base)))
        self.n_rows = int(n_rows)

    def plan_wait(self) -> FrxResult:
        res = FrxResult()
        self._check(self._lib.frx_plan_wait(self._ctx, C.byref(res)))
        return res

    def plan_grid(self, t1, ss1, d1, x_cl, row_first: int = 0, row_count: Optional[int] = None) -> FrxResult:
        t1, ss1, d1 = _f64(t1), _f64(ss1), _f64(d1)
        xcl = _f64(np.concatenate([np.asarray(x_cl[0], dtype=np.float64), np.asarray(x_cl[1], dtype=np.float64)]))
        total = t1.size * ss1.size * d1.size
        if row_count is None:
            row_count = total - row_first
        res = FrxResult()
        self.generation += 1
        self._check(self._lib.frx_plan_grid(self._ctx, t1.size, _dptr(t1), ss1.size, _dptr(ss1), d1.size, _dptr(d1),
                                            _dptr(xcl), int(row_first), int(row_count), C.byref(res)))
        self.n_rows = int(row_count)
        return res

    # ---- read-back --------------------------------------------------------------------------
    def state_pitch(self) -> int:
        return int(self._lib.frx_state_pitch(self._ctx))

    def last_launches(self) -> int:
        """Kernels of the library the last plan launched."""
        return int(self._lib.frx_last_launches(self._ctx))

    @staticmethod
    def _mask(fields) -> int:
        if fields is None:
            return (1 << NUM_FIELDS) - 1
        m = 0
        for f in fields:
            m |= 1 << (FIELD_ID[f] if isinstance(f, str) else int(f))
        return m

    def get_states(self, idx, fields=None) -> np.ndarray:
        """Gather rows `idx` -> array [n_fields, len(idx), Nt] (ascending field id)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        mask = self._mask(fields)
        nf = bin(mask).count("1")
        pitch = self.state_pitch()
        out = np.empty((nf, idx.size, pitch), dtype=np.float64)
        self._check(self._lib.frx_get_states(self._ctx, idx.size, idx.ctypes.data_as(C.POINTER(C.c_int64)), mask, _dptr(out)))
        return out[:, :, :self.Nt]

    def winner_states(self, fields=None) -> np.ndarray:
        """State rows of the selected candidate of the last plan -> array [n_fields, Nt]; host-side copy out of the
        mapped result record (no device round trip)."""
        mask = self._mask(fields)
        nf = bin(mask).count("1")
        pitch = self.state_pitch()
        out = np.empty((nf, pitch), dtype=np.float64)
        self._check(self._lib.frx_winner_states(self._ctx, mask, _dptr(out)))
        return out[:, :self.Nt]

    def winner_record(self):
        """(flags, traj_len, total cost, unweighted cost terms) of the selected candidate, from the mapped result record."""
        fl, tl, tot = C.c_uint32(), C.c_int32(), C.c_double()
        costs = np.empty(max(self.n_costs, 1), dtype=np.float64)
        self._check(self._lib.frx_winner_record(self._ctx, C.byref(fl), C.byref(tl), C.byref(tot), _dptr(costs)))
        return int(fl.value), int(tl.value), float(tot.value), costs[:self.n_costs]

    def get_states_range(self, first=0, count=None, fields=None) -> np.ndarray:
        count = self.n_rows - first if count is None else count
        mask = self._mask(fields)
        nf = bin(mask).count("1")
        pitch = self.state_pitch()
        out = np.empty((nf, count, pitch), dtype=np.float64)
        self._check(self._lib.frx_get_states_range(self._ctx, first, count, mask, _dptr(out)))
        return out[:, :, :self.Nt]

    def get_costs(self, first=0, count=None):
        count = self.n_rows - first if count is None else count
        costs = np.empty((count, self.n_costs), dtype=np.float64)
        total = np.empty(count, dtype=np.float64)
        self._check(self._lib.frx_get_costs(self._ctx, first, count, _dptr(costs) if self.n_costs else None, _dptr(total)))
        return costs, total

    def get_flags(self, first=0, count=None):
        count = self.n_rows - first if count is None else count
        flags = np.empty(count, dtype=np.uint32)
        tl = np.empty(count, dtype=np.int32)
        self._check(self._lib.frx_get_flags(self._ctx, first, count, flags.ctypes.data_as(C.POINTER(C.c_uint32)),
                                            tl.ctypes.data_as(C.POINTER(C.c_int32))))
        return flags, tl

    def device_pointers(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        self._check(self._lib.frx_device_pointers(self._ctx, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def selftest_fdiv(self, a, b):
        a, b = _f64(a), _f64(b)
        q1, q2 = np.empty_like(a), np.empty_like(a)
        self._check(self._lib.frx_selftest_fdiv(self._ctx, a.size, _dptr(a), _dptr(b), _dptr(q1), _dptr(q2)))
        return q1, q2

    def selftest_divc(self, a, b: float):
        a = _f64(a)
        q1, q2 = np.empty_like(a), np.empty_like(a)
        self._check(self._lib.frx_selftest_divc(self._ctx, a.size, _dptr(a), float(b), _dptr(q1), _dptr(q2)))
        return q1, q2

    def fp64_peak_tflops(self) -> float:
        """Measured DFMA throughput of this GPU (2 flop per FMA), the roofline denominator of the obstacle kernel."""
        v = C.c_double()
        self._check(self._lib.frx_selftest_fp64_peak(self._ctx, C.byref(v)))
        return float(v.value)

    def set_exchange(self, page_address: Optional[int], rank: int = 0, world: int = 1):
        """Attach the node's shared exchange page (EXCHANGE_PAGE_BYTES of zeroed POSIX shm); None detaches."""
        self._check(self._lib.frx_set_exchange(self._ctx, C.c_void_p(page_address) if page_address else None, int(rank), int(world)))

    def exchange_wait(self, timeout_us: int = 20_000_000):
        """-> (global min cost, global row, owner rank) once every rank's record of the last plan has arrived."""
        c, r, o = C.c_double(), C.c_int64(), C.c_int32()
        self._check(self._lib.frx_exchange_wait(self._ctx, int(timeout_us), C.byref(c), C.byref(r), C.byref(o)))
        return float(c.value), int(r.value), int(o.value)

    def winner_device_pointer(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.frx_winner_device_pointer(self._ctx, C.byref(p)))
        return p.value

    def synchronize(self):
        self._check(self._lib.frx_synchronize(self._ctx))


def plan_batched(handlers: Sequence["Handler"], samplings: Sequence[np.ndarray]):
    """One eval-kernel launch for several planners (agents): -> list of FrxResult, one per handler."""
    n = len(handlers)
    if n == 0 or n != len(samplings):
        raise ValueError("need one sampling matrix per handler")
    lib = load_library()
    mats = [np.ascontiguousarray(S, dtype=np.float64) for S in samplings]
    for S in mats:
        if S.ndim != 2 or S.shape[1] != 13:
            raise ValueError("sampling matrix must be [N, 13]")
    ctxs = (C.c_void_p * n)(*[h._ctx for h in handlers])
    rows = (C.c_int64 * n)(*[S.shape[0] for S in mats])
    ptrs = (C.POINTER(C.c_double) * n)(*[_dptr(S) for S in mats])
    res = (FrxResult * n)()
    for h in handlers:
        h.generation += 1
    rc = lib.frx_plan_batched(n, ctxs, rows, ptrs, res)
    if rc != 0:
        msg = lib.frx_last_error(handlers[0]._ctx)
        raise FrxError(f"libfrx_b200 error {rc}: {msg.decode() if msg else '?'}")
    for h, S in zip(handlers, mats):
        h.n_rows = S.shape[0]
    return list(res)
