"""CPU ORACLE (numpy) for the reactive planner's inner loop -- TEST INFRASTRUCTURE ONLY.

This file is a checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``frenetix_motion_planner_b200``) never does.

It restates, op-for-op and in the same operation order, the Python path (``use_cpp=False``)
of TUM-AVS/Frenetix-Motion-Planner:

* trajectory generation ............ frenetix_motion_planner/reactive_planner.py:132-182
* quartic / quintic coefficients ... frenetix_motion_planner/polynomial_trajectory.py:293-343, 452-488
* polynomial evaluation ............ frenetix_motion_planner/polynomial_trajectory.py:172-272
* feasibility / back-projection .... frenetix_motion_planner/reactive_planner.py:274-577
* angle interpolation .............. cr_scenario_handler/utils/utils_coordinate_system.py:137-155
* cost terms ....................... frenetix_motion_planner/cost_functions/partial_cost_functions.py:24-64,120-196,341-356
* weighted sum ..................... frenetix_motion_planner/cost_functions/cost_function.py:55-91
* inverse Mahalanobis .............. risk_assessment/collision_probability.py:264-299
* sort / selection ................. frenetix_motion_planner/trajectories.py:524-561, reactive_planner.py:229-272
* collision vs. predictions ........ frenetix_motion_planner/planner.py:329-392,488-534,
                                     cr_scenario_handler/utils/collision_check.py:110-200,
                                     frenetix_motion_planner/state.py:30-39

PINNING STATUS.  The reference ships no tests or golden vectors.  The parts of this oracle
that restate code that IS in the reference tree are pinned against that code itself:
``tests/golden/make_golden.py`` imports the real ``reactive_planner.py`` /
``polynomial_trajectory.py`` / ``partial_cost_functions.py`` / ``cost_function.py`` /
``collision_probability.py`` / ``sampling_matrix.py`` (third-party imports stubbed) and
records their outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
The parts that stand in for third-party code that is NOT in the tree are *defined* here and are
PARITY-UNPINNED: (1) CCosy ``convert_to_cartesian_coords`` (commonroad-drivability-checker
2024.1) -> :func:`ccosy_to_cartesian`; (2) ``make_valid_orientation`` (commonroad-io 2024.2)
-> :func:`make_valid_orientation`; (3) pycrcc ``trajectory_preprocess_obb_sum`` + OBB/OBB
overlap -> :func:`obb_sum_hull`, :func:`obb_overlap`.  See DESIGN.md section 3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_EPS = 1e-5  # reactive_planner.py:26

# field order of the state tensor (shared with the CUDA library, include/frx.h)
FIELDS = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot",
          "s", "d", "theta_cl", "s_dot", "s_ddot", "d_dot", "d_ddot")
F_X, F_Y, F_THETA, F_V, F_A, F_KAPPA, F_KAPPA_DOT, F_S, F_D, F_THETA_CL, F_S_DOT, F_S_DDOT, F_D_DOT, F_D_DDOT = range(14)

# flag bits (shared with include/frx.h)
FLAG_VALID = 1 << 0
FLAG_FEASIBLE = 1 << 1
# reason r (1..10, slots of _infeasible_count_kinematics, logging_helpers.py:365-375) -> bit (1 + r)
def reason_bit(r: int) -> int:
    return 1 << (1 + r)
FLAG_COLLIDE = 1 << 12
FLAG_BOUNDARY = 1 << 13
COLLIDE_STEP_SHIFT, BOUNDARY_STEP_SHIFT = 18, 24      # 6-bit index of the first colliding ego hull (include/frx.h)
FLAG_STORED = 1 << 14      # reached reactive_planner.py:551-567 (has .cartesian/.curvilinear)
FLAG_IN_LIST = 1 << 15     # member of `trajectories_all`
FLAG_COSTED = 1 << 16      # cost function was evaluated for it
FLAG_CANDIDATE = 1 << 17   # handed to trajectory_collision_check

# cost term ids (alphabetical = the reference's evaluation order, cost_function.py:58-60)
COST_NAMES = ("acceleration", "distance_to_obstacles", "distance_to_reference_path", "jerk",
              "lateral_jerk", "longitudinal_jerk", "orientation_offset", "path_length",
              "prediction", "velocity_offset")
COST_ID = {n: i for i, n in enumerate(COST_NAMES)}


@dataclass
class RefPath:
    """Reference-path tables, utils_coordinate_system.py:203-207 (+ polyline vertices)."""
    ref_pos: np.ndarray
    ref_theta: np.ndarray
    ref_curv: np.ndarray
    ref_curv_d: np.ndarray
    ref_x: np.ndarray
    ref_y: np.ndarray


@dataclass
class Params:
    dt: float = 0.1
    N: int = 30
    a_max: float = 11.5
    v_switch: float = 7.319
    delta_max: float = 1.066
    wheelbase: float = 2.5789
    wb_rear_axle: float = 1.4227
    length: float = 4.508
    width: float = 1.610
    low_vel_mode: bool = False
    x0_orientation: float = 0.0
    desired_velocity: float = 8.0
    draw_traj_set: bool = True
    kinematic_debug: bool = True
    # cost.yaml:3-17 defaults (zero weights are dropped, cost_function.py:55-57)
    cost_weights: dict = field(default_factory=lambda: {
        "lateral_jerk": 0.2, "longitudinal_jerk": 0.2, "velocity_offset": 1.0,
        "distance_to_reference_path": 5.0, "prediction": 0.2})
    # distance_to_obstacles needs the obstacles' current positions [[x, y], ...]
    obstacle_positions: Optional[np.ndarray] = None
    # the `s_velocity > 0.001` stand-still threshold of reactive_planner.py:393-447.  Tests move it by +-1e-9 to get BOTH
    # outcomes of a candidate whose velocity profile is constructed to end at exactly 0.001 m/s (a structural tie that
    # LAPACK rounding noise decides in the reference itself): the device must reproduce one of the two.
    standstill_threshold: float = 0.001
    # ---- the use_cpp = True flavour (reactive_planner_cpp.py:96-178).  frenetix 0.4.0's sources are not in the tree, so
    # these follow what the reference's Python files say about them (parity unpinned):
    # CheckCurvatureRateConstraint(wheelbase, velocityDeltaMax): the limit the Python path carries as a comment,
    # reactive_planner.py:513-515 -- kappa_dot_max = v_delta_max / (wheelbase * cos(steering_angle) ** 2)
    curvature_rate_from_v_delta: bool = False
    v_delta_max: float = 0.4
    # CalculateVelocityOffsetCost(..., norm_order=2): squared instead of absolute offsets
    velocity_offset_norm: int = 1
    # prediction cost: 0 = inverse Mahalanobis (python path), 1 = CalculateCollisionProbabilityFast, i.e. the in-tree
    # get_collision_probability_fast (risk_assessment/collision_probability.py:141-261; prediction_costs :344-348)
    prediction_cost_mode: int = 0

    def active_costs(self):
        names = [k for k, w in self.cost_weights.items() if w != 0]
        names.sort()
        for n in names:
            if n not in COST_ID:
                raise NotImplementedError(f"cost term '{n}' is host-only / unimplemented in the reference")
        return names


# ----------------------------------------------------------------------------------------------
# third-party stand-ins (PARITY-UNPINNED, see module docstring)
# ----------------------------------------------------------------------------------------------
def make_valid_orientation(angle):
    """commonroad.common.util.make_valid_orientation (commonroad-io 2024.2), restated."""
    two_pi = 2.0 * np.pi
    angle = angle % two_pi
    if np.pi <= angle <= two_pi:
        angle = angle - two_pi
    return angle


def ccosy_to_cartesian(ref: RefPath, s, d):
    """Definition used for CCosy.convert_to_cartesian_coords: linear point on the
    segment plus d times the normal of the linearly interpolated (unwrapped) heading.
    Returns None outside [ref_pos[0], ref_pos[-1])."""
    p = ref.ref_pos
    if not (s >= p[0]) or not (s < p[-1]):
        return None
    i = int(np.argmax(p > s)) - 1
    lam = (s - p[i]) / (p[i + 1] - p[i])
    px = (1.0 - lam) * ref.ref_x[i] + lam * ref.ref_x[i + 1]
    py = (1.0 - lam) * ref.ref_y[i] + lam * ref.ref_y[i + 1]
    th = ref.ref_theta[i] + lam * (ref.ref_theta[i + 1] - ref.ref_theta[i])
    return np.array([px - d * math.sin(th), py + d * math.cos(th)])


def obb_sum_hull(c0x, c0y, th0, c1x, c1y, th1, hl, hw):
    """Definition used for pycrcc trajectory_preprocess_obb_sum: the smallest box, in the frame of
    box k, that contains boxes k and k+1.  Returns (cx, cy, ux, uy, ha, hb)."""
    ux, uy = math.cos(th0), math.sin(th0)
    u1x, u1y = math.cos(th1), math.sin(th1)
    dx = c1x - c0x
    dy = c1y - c0y
    du = dx * ux + dy * uy
    dv = dy * ux - dx * uy
    c = abs(ux * u1x + uy * u1y)
    sn = abs(ux * u1y - uy * u1x)
    eu = hl * c + hw * sn
    ev = hl * sn + hw * c
    lo_u = min(-hl, du - eu)
    hi_u = max(hl, du + eu)
    lo_v = min(-hw, dv - ev)
    hi_v = max(hw, dv + ev)
    mu = 0.5 * (lo_u + hi_u)
    mv = 0.5 * (lo_v + hi_v)
    ha = 0.5 * (hi_u - lo_u)
    hb = 0.5 * (hi_v - lo_v)
    cx = c0x + (mu * ux - mv * uy)
    cy = c0y + (mu * uy + mv * ux)
    return (cx, cy, ux, uy, ha, hb)


def obb_overlap(e, o):
    """Exact OBB-vs-OBB separating-axis test; touching counts as overlap."""
    ecx, ecy, eux, euy, eha, ehb = e
    ocx, ocy, oux, ouy, oha, ohb = o
    dx = ocx - ecx
    dy = ocy - ecy
    c = abs(eux * oux + euy * ouy)
    sn = abs(eux * ouy - euy * oux)
    if abs(dx * eux + dy * euy) > eha + (oha * c + ohb * sn):
        return False
    if abs(dy * eux - dx * euy) > ehb + (oha * sn + ohb * c):
        return False
    if abs(dx * oux + dy * ouy) > oha + (eha * c + ehb * sn):
        return False
    if abs(dy * oux - dx * ouy) > ohb + (eha * sn + ehb * c):
        return False
    return True


# ----------------------------------------------------------------------------------------------
# restated reference code
# ----------------------------------------------------------------------------------------------
def time_grid(T: float, dt: float):
    """reactive_planner.py:296-300 -- time samples and their *rounded* powers."""
    t = np.round(np.arange(0, T + dt, dt), 5)
    t2 = np.round(np.power(t, 2), 10)
    t3 = np.round(np.power(t, 3), 10)
    t4 = np.round(np.power(t, 4), 10)
    t5 = np.round(np.power(t, 5), 10)
    return t, t2, t3, t4, t5


def quartic_coeffs(xs, vxs, axs, vxe, axe_unused, T):
    """polynomial_trajectory.py:452-488.  NB: the RHS hard-wires the end acceleration to 0
    (``-axs``), the sampled value is ignored."""
    A = np.array([[3 * T ** 2, 4 * T ** 3],
                  [6 * T, 12 * T ** 2]])
    b = np.array([vxe - vxs - axs * T,
                  - axs])
    x = np.linalg.solve(A, b)
    return np.array([xs, vxs, axs / 2.0, x[0], x[1], 0.0])


def quintic_coeffs(xs, vxs, axs, xe, vxe, axe, T):
    """polynomial_trajectory.py:293-343."""
    A = np.array([[T ** 3., T ** 4, T ** 5],
                  [3. * T ** 2, 4. * T ** 3, 5. * T ** 4],
                  [6. * T, 12. * T ** 2, 20. * T ** 3]])
    b = np.array([xe - xs - vxs * T - .5 * axs * T ** 2,
                  vxe - vxs - axs * T,
                  axe - axs])
    x = np.linalg.solve(A, b)
    return np.array([xs, vxs, .5 * axs, x[0], x[1], x[2]])


def calc_position(c, tau, tau2, tau3, tau4, tau5):
    return c[0] + c[1] * tau + c[2] * tau2 + c[3] * tau3 + c[4] * tau4 + c[5] * tau5


def calc_velocity(c, tau, tau2, tau3, tau4):
    return c[1] + 2. * c[2] * tau + 3. * c[3] * tau2 + 4. * c[4] * tau3 + 5. * c[5] * tau4


def calc_acceleration(c, tau, tau2, tau3):
    return 2 * c[2] + 6 * c[3] * tau + 12 * c[4] * tau2 + 20 * c[5] * tau3


def squared_jerk_integral(c, t):
    """polynomial_trajectory.py:172-191."""
    t2 = t * t
    t3 = t2 * t
    t4 = t3 * t
    t5 = t4 * t
    return (36 * c[3] * c[3] * t + 144 * c[3] * c[4] * t2 +
            240 * c[3] * c[5] * t3 + 192 * c[4] * c[4] * t3 +
            720 * c[4] * c[5] * t4 + 720 * c[5] * c[5] * t5)


def evaluate_position_at_tau(c, tau, delta_tau):
    """polynomial_trajectory.py:193-228 (position only; tau_0 = 0)."""
    if tau < 0:
        tau = 0.0
    elif tau > delta_tau:
        tau = delta_tau
    tau2 = np.power(tau, 2)
    tau3 = tau2 * tau
    tau4 = tau2 * tau2
    tau5 = tau3 * tau2
    return calc_position(c, tau, tau2, tau3, tau4, tau5)


def interpolate_angle(x, x1, x2, y1, y2):
    """utils_coordinate_system.py:137-155."""
    delta = y2 - y1
    return make_valid_orientation(delta * (x - x1) / (x2 - x1) + y1)


def simps(y, dx):
    """scipy 1.13.1 ``integrate.simps(y, dx=dx)`` (default even='simpson'), restated.
    Odd sample count: composite Simpson.  Even sample count: Simpson on the first n-1 samples plus
    the Cartwright end correction on the last interval."""
    y = np.asarray(y, dtype=float)
    n = len(y)
    if n == 1:
        return 0.0
    if n == 2:
        return 0.5 * dx * (y[0] + y[1])
    if n % 2 == 1:
        return dx / 3.0 * np.sum(y[0:-2:2] + 4.0 * y[1:-1:2] + y[2::2])
    result = dx / 3.0 * np.sum(y[0:n - 3:2] + 4.0 * y[1:n - 2:2] + y[2:n - 1:2])
    # Cartwright correction of the last interval, written as scipy computes it (h0 = h1 = dx)
    alpha = (2 * dx ** 2 + 3 * dx * dx) / (6 * (dx + dx))
    beta = (dx ** 2 + 3.0 * dx * dx) / (6 * dx)
    eta = (1 * dx ** 3) / (6 * dx * (dx + dx))
    result += alpha * y[-1] + beta * y[-2] - eta * y[-3]
    return result


# ----------------------------------------------------------------------------------------------
# bivariate normal rectangle probability: what scipy.stats.mvn.mvnun(lower, upper, mean, cov) returns for
# two dimensions (Genz's MVNDST calls the deterministic bivariate routine there).  scipy removed the
# module; this is Genz's BVND algorithm (A. Genz, "Numerical computation of rectangular bivariate and
# trivariate normal and t probabilities", Statistics and Computing 14, 2004): Gauss-Legendre quadrature of
# the Plackett integral with 6 / 12 / 20 points depending on |rho|, the |rho| > 0.925 branch with its
# asymptotic corrections.  Checked against scipy.stats.multivariate_normal.cdf(..., lower_limit=...)
# to 4e-16 (tests/test_oracle_golden.py).
# ----------------------------------------------------------------------------------------------
_GL_W = {3: [0.1713244923791705, 0.3607615730481384, 0.4679139345726904],
         6: [0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547,
             0.2491470458134029],
         10: [0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
              0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259]}
_GL_X = {3: [0.9324695142031522, 0.6612093864662647, 0.2386191860831970],
         6: [0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802,
             0.1252334085114692],
         10: [0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
              0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733]}


def _phid(z):
    return 0.5 * math.erfc(-z / math.sqrt(2.0))


def bvnu(dh, dk, r):
    """P(X > dh, Y > dk), standard bivariate normal with correlation r."""
    if r == 0:
        return _phid(-dh) * _phid(-dk)
    tp = 2 * math.pi
    h, k = dh, dk
    hk = h * k
    bvn = 0.0
    lg = 3 if abs(r) < 0.3 else (6 if abs(r) < 0.75 else 10)
    w, x = _GL_W[lg], _GL_X[lg]
    if abs(r) < 0.925:
        hs = (h * h + k * k) / 2
        asr = math.asin(r) / 2
        for wi, xi in zip(w, x):
            for xs in (1 - xi, 1 + xi):
                sn = math.sin(asr * xs)
                bvn += wi * math.exp((sn * hk - hs) / (1 - sn * sn))
        bvn = bvn * asr / tp + _phid(-h) * _phid(-k)
    else:
        if r < 0:
            k = -k
            hk = -hk
        if abs(r) < 1:
            as_ = 1 - r * r
            a = math.sqrt(as_)
            bs = (h - k) ** 2
            asr = -(bs / as_ + hk) / 2
            c = (4 - hk) / 8
            d = (12 - hk) / 80
            if asr > -100:
                bvn = a * math.exp(asr) * (1 - c * (bs - as_) * (1 - d * bs) / 3 + c * d * as_ * as_)
            if hk > -100:
                b = math.sqrt(bs)
                sp = math.sqrt(tp) * _phid(-b / a)
                bvn = bvn - math.exp(-hk / 2) * sp * b * (1 - c * bs * (1 - d * bs) / 3)
            a = a / 2
            acc = 0.0
            for wi, xi in zip(w, x):
                for xx in (1 - xi, 1 + xi):
                    xs = (a * xx) ** 2
                    asr = -(bs / xs + hk) / 2
                    if asr > -100:
                        sp = 1 + c * xs * (1 + 5 * d * xs)
                        rs = math.sqrt(1 - xs)
                        ep = math.exp(-(hk / 2) * xs / (1 + rs) ** 2) / rs
                        acc += wi * math.exp(asr) * (sp - ep)
            bvn = (a * acc - bvn) / tp
        if r > 0:
            bvn = bvn + _phid(-max(h, k))
        elif h >= k:
            bvn = -bvn
        else:
            L = _phid(k) - _phid(h) if h < 0 else _phid(-h) - _phid(-k)
            bvn = L - bvn
    return max(0.0, min(1.0, bvn))


def mvnun2(lower, upper, mu, cov):
    """Rectangle probability of N(mu, cov) in two dimensions (the value part of scipy.stats.mvn.mvnun)."""
    sx, sy = math.sqrt(cov[0][0]), math.sqrt(cov[1][1])
    r = cov[0][1] / (sx * sy)
    a1, a2 = (lower[0] - mu[0]) / sx, (lower[1] - mu[1]) / sy
    b1, b2 = (upper[0] - mu[0]) / sx, (upper[1] - mu[1]) / sy
    return bvnu(a1, a2, r) - bvnu(b1, a2, r) - bvnu(a1, b2, r) + bvnu(b1, b2, r)


def collision_probability_fast(x, y, theta, predictions, veh_length, veh_width):
    """risk_assessment/collision_probability.py:141-261, restated: per obstacle the per-step probability that the ego
    (three axis-aligned rectangles of length / 3 x width around the REAR-AXLE position and +- length / 3 along its heading,
    :336-371) overlaps the obstacle (three normal distributions: predicted mean and the mean +- length / 2 along the NEXT
    step's yaw, :173-178), zero when all three means are more than 5 m away."""
    ego_pos = np.stack((x, y), axis=-1)
    offset = np.array([veh_length / 6, veh_width / 2])
    out = {}
    for oid, pred in enumerate(predictions):
        mean_list, cov_list, yaw_list = pred["pos_list"], pred["cov_list"], pred["orientation_list"]
        length = pred["shape"]["length"]
        probs = []
        min_len = min(len(x), len(mean_list))
        ego_pos_array = ego_pos[1:min_len]
        dev = np.stack((np.cos(yaw_list[1:min_len]), np.sin(yaw_list[1:min_len])), axis=-1) * length / 2
        mean_array = np.array(mean_list[:min_len - 1])
        total_mean_array = np.array([mean_array, mean_array + dev, mean_array - dev])
        dist = total_mean_array - ego_pos_array
        dist = np.sqrt(dist[:, :, 0] ** 2 + dist[:, :, 1] ** 2)
        far = dist.min(axis=0) > 5.0
        for i in range(1, len(x)):
            if i < len(mean_list):
                if far[i - 1]:
                    prob = 0.0
                else:
                    cov = cov_list[i - 1]
                    if all(c == 0 for c in (cov[0][0], cov[0][1], cov[1][0], cov[1][1])):
                        cov = [[0.1, 0.0], [0.0, 0.1]]
                    prob = 0.0
                    c0 = ego_pos_array[i - 1]
                    ax = np.array([math.cos(theta[i]), math.sin(theta[i])])
                    r_x = veh_length / 2
                    centers = np.array([c0, c0 + r_x * (2 / 3) * ax, c0 - r_x * (2 / 3) * ax])
                    upper, lower = centers + offset, centers - offset
                    for mu in total_mean_array[:, i - 1]:
                        for q in range(3):
                            prob += mvnun2(lower[q], upper[q], mu, cov)
            else:
                prob = 0.0
            probs.append(prob / 3)
        out[oid] = np.array(probs)
    return out


def _costs_for(name, st, c_lon, c_lat, prm: Params, predictions, inv_covs, Nt):
    """partial_cost_functions.py -- one unweighted cost term for one candidate."""
    x, y = st[F_X], st[F_Y]
    if name == "lateral_jerk":                       # :49-55  (evaluated at t = trajectory.dt !)
        return squared_jerk_integral(c_lat, prm.dt)
    if name == "longitudinal_jerk":                  # :58-64
        return squared_jerk_integral(c_lon, prm.dt)
    if name == "velocity_offset":                    # :120-130
        vel = st[F_V]
        half_idx = int(len(vel) / 2)
        if prm.velocity_offset_norm == 2:
            cost = np.sum(np.square(vel[half_idx:-1] - prm.desired_velocity))
        else:
            cost = np.sum(np.abs(vel[half_idx:-1] - prm.desired_velocity))
        cost += np.abs(((vel[-1] - prm.desired_velocity) ** 2))
        return float(cost)
    if name == "distance_to_reference_path":         # :154-169 (len(d + 4) == len(d))
        d = st[F_D]
        return float((np.sum(np.abs(d)) + np.abs(d[-1]) * 5) / len(d + 4))
    if name == "prediction" and prm.prediction_cost_mode == 1:
        probs = collision_probability_fast(x, y, st[F_THETA], predictions, prm.length, prm.width)
        pred_costs = 0
        for key in probs:
            pred_costs += np.sum(probs[key])
        return pred_costs
    if name == "prediction":                         # :341-356 + collision_probability.py:264-299
        pred_costs = 0
        for o, pred in enumerate(predictions):
            mean_list = pred["pos_list"]
            inv_cov_list = inv_covs[o]
            inv_dist = []
            for i in range(1, len(x)):
                if i < len(mean_list):
                    u = np.array([x[i], y[i]])
                    v = np.array(mean_list[i - 1])
                    iv = np.array(inv_cov_list[i - 1])
                    delta = u - v
                    mahalanobis_squared = delta.T @ iv @ delta
                    with np.errstate(divide="ignore", invalid="ignore"):
                        inv_dist.append(1.0 / (mahalanobis_squared ** 2))
                else:
                    inv_dist.append(0.0)
            pred_costs += np.sum(np.array(inv_dist))
        return pred_costs
    if name == "acceleration":                       # :24-33
        return simps(np.square(st[F_A]), dx=prm.dt)
    if name == "jerk":                               # :36-46
        jerk = np.diff(st[F_A]) / prm.dt
        return simps(np.square(jerk), dx=prm.dt)
    if name == "orientation_offset":                 # :141-151
        th = np.diff(st[F_THETA_CL]) / prm.dt
        return simps(np.square(th), dx=prm.dt)
    if name == "path_length":                        # :189-196
        return simps(st[F_V], dx=prm.dt)
    if name == "distance_to_obstacles":              # :172-186 (cdist euclidean, 1/dist**2)
        cost = 0.0
        if prm.obstacle_positions is not None:
            for p in np.asarray(prm.obstacle_positions, dtype=float):
                dists = np.sqrt((x - p[0]) ** 2 + (y - p[1]) ** 2)
                with np.errstate(divide="ignore"):
                    cost += np.sum(np.reciprocal(dists ** 2))
        return float(cost)
    raise NotImplementedError(name)


def check_feasibility_one(row, ref: RefPath, prm: Params):
    """reactive_planner.py:154-171 (polynomials for one sampling row) + :290-569 (one candidate).

    Returns dict(state[14,Nt], flags, traj_len, reasons[11], c_lon, c_lat, in_list, stored)."""
    N, dT = prm.N, prm.dt
    Nt = N + 1
    t1 = row[1]
    s0, ss0, sss0, ss1 = row[2], row[3], row[4], row[5]
    d0, dd0, ddd0, d1, dd1, ddd1 = row[7], row[8], row[9], row[10], row[11], row[12]

    # ---- reactive_planner.py:154-171
    c_lon = quartic_coeffs(s0, ss0, sss0, ss1, 0, t1)
    if prm.low_vel_mode:
        s_lon_goal = evaluate_position_at_tau(c_lon, t1, t1) - s0
        if s_lon_goal <= 0:
            s_lon_goal = t1
        c_lat = quintic_coeffs(d0, dd0, ddd0, d1, dd1, ddd1, s_lon_goal)
    else:
        c_lat = quintic_coeffs(d0, dd0, ddd0, d1, dd1, ddd1, t1)

    reasons = np.zeros(11)
    feasible = True
    valid = True
    out = dict(c_lon=c_lon, c_lat=c_lat)
    # distance of every data-dependent DECISION to its threshold (SURVEY.md 4.5 "margin protocol"):
    # candidates whose smallest margin is ~1 ulp sit on a structural tie (e.g. a velocity profile
    # constructed to end at exactly 0.001 m/s, the `> 0.001` stand-still threshold) whose outcome
    # depends on LAPACK rounding noise even in the reference itself.
    margin = [np.inf]

    def upd(m):
        m = np.min(np.abs(np.asarray(m, dtype=float))) if np.size(m) else np.inf
        if np.isnan(m):
            m = 0.0
        if m < margin[0]:
            margin[0] = float(m)
    if prm.low_vel_mode:
        upd(evaluate_position_at_tau(c_lon, t1, t1) - s0)

    # ---- :296-303
    t, t2, t3, t4, t5 = time_grid(t1, dT)
    traj_len = len(t)
    out["traj_len"] = traj_len
    if traj_len > Nt:
        raise ValueError("time grid longer than planning horizon (reference would raise too)")

    s = np.zeros(Nt); s_velocity = np.zeros(Nt); s_acceleration = np.zeros(Nt)
    d = np.zeros(Nt); d_velocity = np.zeros(Nt); d_acceleration = np.zeros(Nt)

    # ---- :314-322
    s[:traj_len] = calc_position(c_lon, t, t2, t3, t4, t5)
    s_velocity[:traj_len] = calc_velocity(c_lon, t, t2, t3, t4)
    s_acceleration[:traj_len] = calc_acceleration(c_lon, t, t2, t3)
    for ext in range(traj_len, Nt):
        s[ext] = s[ext - 1] + dT * s_velocity[traj_len - 1]
    s_velocity[traj_len:] = s_velocity[traj_len - 1]
    s_acceleration[traj_len:] = 0.0

    # ---- :325-346
    if not prm.low_vel_mode:
        d[:traj_len] = calc_position(c_lat, t, t2, t3, t4, t5)
        d_velocity[:traj_len] = calc_velocity(c_lat, t, t2, t3, t4)
        d_acceleration[:traj_len] = calc_acceleration(c_lat, t, t2, t3)
    else:
        s1 = s[:traj_len] - s[0]
        s2 = np.square(s1)
        s3 = s2 * s1
        s4 = np.square(s2)
        s5 = s4 * s1
        d[:traj_len] = calc_position(c_lat, s1, s2, s3, s4, s5)
        d_velocity[:traj_len] = calc_velocity(c_lat, s1, s2, s3, s4)
        d_acceleration[:traj_len] = calc_acceleration(c_lat, s1, s2, s3)
    d[traj_len:] = d[traj_len - 1]
    d_velocity[traj_len:] = 0.0
    d_acceleration[traj_len:] = 0.0

    state = np.zeros((14, Nt))

    def finish(in_list, stored):
        fl = (FLAG_VALID if valid else 0) | (FLAG_FEASIBLE if feasible else 0)
        for r in range(1, 11):
            if reasons[r]:
                fl |= reason_bit(r)
        if in_list:
            fl |= FLAG_IN_LIST
        if stored:
            fl |= FLAG_STORED
        out.update(state=state, flags=fl, reasons=reasons, feasible=feasible, valid=valid, margin=margin[0])
        return out

    upd(s_velocity + _EPS)
    upd(np.abs(s_velocity) - _EPS)
    if not prm.draw_traj_set:
        upd(np.abs(s_acceleration) - prm.a_max)
    # ---- :350-355
    if np.any(s_velocity < -_EPS):
        valid = False
        reasons[10] += 1
        if not prm.draw_traj_set and not prm.kinematic_debug:
            return finish(False, False)
    s_velocity[np.abs(s_velocity) < _EPS] = 0.0

    x = np.zeros(Nt); y = np.zeros(Nt); v = np.zeros(Nt); a = np.zeros(Nt)
    theta_gl = np.zeros(Nt); theta_cl = np.zeros(Nt); kappa_gl = np.zeros(Nt); kappa_cl = np.zeros(Nt)

    # ---- :373-386
    if not prm.draw_traj_set:
        if np.any(np.abs(s_acceleration) > prm.a_max):
            feasible = False
            reasons[1] += 1
            return finish(True, False)
        if np.any(s_velocity < -_EPS):
            feasible = False
            reasons[2] += 1
            return finish(True, False)

    ref_pos, ref_theta, ref_curv, ref_curv_d = ref.ref_pos, ref.ref_theta, ref.ref_curv, ref.ref_curv_d
    brk = (not prm.draw_traj_set) and (not prm.kinematic_debug)

    # ---- :389-533
    with np.errstate(all="ignore"):
        for i in range(0, Nt):
            if not prm.low_vel_mode:
                upd(s_velocity[i] - 0.001)
                if s_velocity[i] > prm.standstill_threshold:
                    dp = d_velocity[i] / s_velocity[i]
                else:
                    dp = 0.
                ddot = d_acceleration[i] - dp * s_acceleration[i]
                if s_velocity[i] > prm.standstill_threshold:
                    dpp = ddot / (s_velocity[i] ** 2)
                else:
                    dpp = 0.
            else:
                dp = d_velocity[i]
                dpp = d_acceleration[i]

            s_idx = np.argmax(ref_pos > s[i]) - 1
            if s_idx + 1 >= len(ref_pos):
                feasible = False
                reasons[3] = 1
                break
            s_lambda = (s[i] - ref_pos[s_idx]) / (ref_pos[s_idx + 1] - ref_pos[s_idx])

            if s_velocity[i] > prm.standstill_threshold or prm.low_vel_mode:
                theta_cl[i] = np.arctan2(dp, 1.0)
                theta_gl[i] = theta_cl[i] + interpolate_angle(
                    s[i], ref_pos[s_idx], ref_pos[s_idx + 1], ref_theta[s_idx], ref_theta[s_idx + 1])
            else:
                theta_gl[i] = prm.x0_orientation if i == 0 else theta_gl[i - 1]
                theta_cl[i] = theta_gl[i] - interpolate_angle(
                    s[i], ref_pos[s_idx], ref_pos[s_idx + 1], ref_theta[s_idx], ref_theta[s_idx + 1])

            k_r = (ref_curv[s_idx + 1] - ref_curv[s_idx]) * s_lambda + ref_curv[s_idx]
            k_r_d = (ref_curv_d[s_idx + 1] - ref_curv_d[s_idx]) * s_lambda + ref_curv_d[s_idx]

            oneKrD = (1 - k_r * d[i])
            cosTheta = math.cos(theta_cl[i])
            tanTheta = np.tan(theta_cl[i])

            kappa_gl[i] = (dpp + (k_r * dp + k_r_d * d[i]) * tanTheta) * cosTheta * ((cosTheta / oneKrD) ** 2) + \
                          (cosTheta / oneKrD) * k_r
            kappa_cl[i] = kappa_gl[i] - k_r

            v[i] = s_velocity[i] * (oneKrD / (math.cos(theta_cl[i])))

            a[i] = s_acceleration[i] * (oneKrD / cosTheta) + ((s_velocity[i] ** 2) / cosTheta) * (
                    oneKrD * tanTheta * (kappa_gl[i] * (oneKrD / cosTheta) - k_r) - (
                    k_r_d * d[i] + k_r * dp))

            upd(v[i] + _EPS)
            if v[i] < -_EPS:
                feasible = False
                reasons[4] = 1
                if brk:
                    break

            kappa_max = np.tan(prm.delta_max) / prm.wheelbase
            upd(abs(kappa_gl[i]) - kappa_max)
            if abs(kappa_gl[i]) > kappa_max:
                feasible = False
                reasons[5] = 1
                if brk:
                    break

            yaw_rate = (theta_gl[i] - theta_gl[i - 1]) / dT if i > 0 else 0.
            theta_dot_max = kappa_max * v[i]
            if not (abs(round(yaw_rate, 5)) == 0 and theta_dot_max == 0):   # exact 0 > 0 is reproducible
                upd(abs(round(yaw_rate, 5)) - theta_dot_max)
            if abs(round(yaw_rate, 5)) > theta_dot_max:
                feasible = False
                reasons[6] = 1
                if brk:
                    break

            kappa_dot = (kappa_gl[i] - kappa_gl[i - 1]) / dT if i > 0 else 0.
            kappa_dot_max = 0.4
            if prm.curvature_rate_from_v_delta:
                steering_angle = np.arctan2(prm.wheelbase * kappa_gl[i], 1.0)
                kappa_dot_max = prm.v_delta_max / (prm.wheelbase * math.cos(steering_angle) ** 2)
            upd(abs(kappa_dot) - kappa_dot_max)
            if abs(kappa_dot) > kappa_dot_max:
                feasible = False
                reasons[7] = 1
                if brk:
                    break

            v_switch = prm.v_switch
            a_max = prm.a_max * v_switch / v[i] if v[i] > v_switch else prm.a_max
            a_min = -prm.a_max
            upd(a[i] - a_max)
            upd(a[i] - a_min)
            if not a_min <= a[i] <= a_max:
                feasible = False
                reasons[8] = 1
                if brk:
                    break

    # ---- :536-567
    stored = False
    upd(s - ref_pos[0])
    upd(s - ref_pos[-1])
    if feasible or prm.draw_traj_set:
        for i in range(0, Nt):
            pos = ccosy_to_cartesian(ref, s[i], d[i])
            if pos is not None:
                x[i] = pos[0]
                y[i] = pos[1]
            else:
                valid = False
                reasons[9] = 1
                break
        state[F_X], state[F_Y], state[F_THETA], state[F_V], state[F_A] = x, y, theta_gl, v, a
        state[F_KAPPA] = kappa_gl
        state[F_KAPPA_DOT] = np.append([0], np.diff(kappa_gl))
        state[F_S], state[F_D], state[F_THETA_CL] = s, d, theta_cl
        state[F_S_DOT], state[F_S_DDOT], state[F_D_DOT], state[F_D_DDOT] = \
            s_velocity, s_acceleration, d_velocity, d_acceleration
        stored = True
    return finish(stored, stored)


def collides_with_predictions(st, prm: Params, predictions, Nt):
    """planner.py:329-360 + collision_check.py:110-200 for ONE candidate (see module docstring for
    the obb-sum / overlap definitions)."""
    x, y, th = st[F_X], st[F_Y], st[F_THETA]
    hl, hw = prm.length / 2, prm.width / 2
    cx = x + prm.wb_rear_axle * np.cos(th)          # state.py:30-39
    cy = y + prm.wb_rear_axle * np.sin(th)
    ego = [obb_sum_hull(cx[k], cy[k], th[k], cx[k + 1], cy[k + 1], th[k + 1], hl, hw)
           for k in range(Nt - 1)]
    for pred in predictions:
        pos = np.asarray(pred["pos_list"], dtype=float)
        L = min(Nt, len(pos))
        if L <= 2:
            continue
        ori = np.asarray(pred["orientation_list"], dtype=float)
        ohl, ohw = pred["shape"]["length"] / 2, pred["shape"]["width"] / 2
        for k in range(1, min(Nt - 2, L - 1) + 1):
            j = k - 1
            oh = obb_sum_hull(pos[j, 0], pos[j, 1], ori[j], pos[j + 1, 0], pos[j + 1, 1], ori[j + 1], ohl, ohw)
            if obb_overlap(ego[k], oh):
                return True
    return False


def first_collision_step(st, prm: Params, predictions, Nt):
    """Index k of the first ego hull (time index t0 + k) that meets a predicted obstacle, -1 if none: what a pycrcc
    time-variant collision query reports; same pairs as :func:`collides_with_predictions`."""
    x, y, th = st[F_X], st[F_Y], st[F_THETA]
    hl, hw = prm.length / 2, prm.width / 2
    cx = x + prm.wb_rear_axle * np.cos(th)
    cy = y + prm.wb_rear_axle * np.sin(th)
    first = -1
    for pred in predictions:
        pos = np.asarray(pred["pos_list"], dtype=float)
        L = min(Nt, len(pos))
        if L <= 2:
            continue
        ori = np.asarray(pred["orientation_list"], dtype=float)
        ohl, ohw = pred["shape"]["length"] / 2, pred["shape"]["width"] / 2
        kmax = min(Nt - 2, L - 1)
        if first >= 0:
            kmax = min(kmax, first - 1)
        for k in range(1, kmax + 1):
            j = k - 1
            e = obb_sum_hull(cx[k], cy[k], th[k], cx[k + 1], cy[k + 1], th[k + 1], hl, hw)
            oh = obb_sum_hull(pos[j, 0], pos[j, 1], ori[j], pos[j + 1, 0], pos[j + 1, 1], ori[j + 1], ohl, ohw)
            if obb_overlap(e, oh):
                first = k
                break
    return first


def collides_with_static(st, prm: Params, static_obbs, Nt):
    """Road-boundary stand-in (planner.py:362-378): ego hulls vs caller-provided static boxes
    [[cx, cy, theta, half_len, half_wid], ...].  Returns first colliding step or -1."""
    if static_obbs is None or len(static_obbs) == 0:
        return -1
    x, y, th = st[F_X], st[F_Y], st[F_THETA]
    hl, hw = prm.length / 2, prm.width / 2
    cx = x + prm.wb_rear_axle * np.cos(th)
    cy = y + prm.wb_rear_axle * np.sin(th)
    for k in range(Nt - 1):
        e = obb_sum_hull(cx[k], cy[k], th[k], cx[k + 1], cy[k + 1], th[k + 1], hl, hw)
        for b in static_obbs:
            o = (b[0], b[1], math.cos(b[2]), math.sin(b[2]), b[3], b[4])
            if obb_overlap(e, o):
                return k
    return -1


def plan(sampling: np.ndarray, ref: RefPath, prm: Params, predictions: Sequence[dict] = (),
         static_obbs=None, check_all_collisions: bool = True, collision_check: bool = True):
    """One sampling level of ReactivePlannerPython.plan(): reactive_planner.py:89-94 on an explicit
    sampling matrix (rows = sampling_matrix.py:85-121 format).

    Batch-equivalent of the lazy selection: argmin of the total cost over the
    candidates handed to the collision check that neither collide nor leave the road, ties ->
    lowest row index.  With ``check_all_collisions`` every candidate of that set is checked (what
    the device does); otherwise only the ones the reference would visit."""
    sampling = np.asarray(sampling, dtype=np.float64)
    n = sampling.shape[0]
    Nt = prm.N + 1
    names = prm.active_costs()
    K = len(names)
    weights = [prm.cost_weights[k] for k in names]
    # (the collision-probability flavour never inverts a covariance: zero matrices mark ground-truth predictions there)
    inv_covs = [np.linalg.inv(np.asarray(p["cov_list"], dtype=float)) for p in predictions] \
        if prm.prediction_cost_mode == 0 else [None] * len(predictions)

    states = np.zeros((14, n, Nt))
    flags = np.zeros(n, dtype=np.uint32)
    traj_len = np.zeros(n, dtype=np.int32)
    costs = np.zeros((n, K))
    total = np.zeros(n)
    coeffs = np.zeros((n, 12))
    margins = np.zeros(n)
    reason_counts = np.zeros(11)

    for r in range(n):
        o = check_feasibility_one(sampling[r], ref, prm)
        states[:, r, :] = o["state"]
        flags[r] = o["flags"]
        traj_len[r] = o["traj_len"]
        coeffs[r, :6] = o["c_lon"]
        coeffs[r, 6:] = o["c_lat"]
        margins[r] = o["margin"]
        reason_counts += o["reasons"]

    valid = (flags & FLAG_VALID) != 0
    feas = (flags & FLAG_FEASIBLE) != 0
    in_list = (flags & FLAG_IN_LIST) != 0
    n_list = int(in_list.sum())
    n_feasible = int((in_list & valid & feas).sum())
    reason_counts[0] = int((in_list & ~(valid & feas)).sum())          # reactive_planner.py:229-234
    percentage = float(n_feasible / n_list) * 100 if n_list else 0.0   # :235

    # ---- :244-253 which candidates get costs, which go to the collision check
    if prm.draw_traj_set:
        costed = in_list.copy()
        cand = in_list & feas          # filter(lambda x: x.feasible is True) -- `valid` NOT consulted
    else:
        costed = in_list & valid & feas
        cand = costed.copy()
    with np.errstate(all="ignore"):
        for r in np.nonzero(costed)[0]:
            st = states[:, r, :]
            costlist = np.zeros(K)
            costlist_weighted = np.zeros(K)
            for num, name in enumerate(names):
                costlist[num] = _costs_for(name, st, coeffs[r, :6], coeffs[r, 6:], prm, predictions, inv_covs, Nt)
                costlist_weighted[num] = weights[num] * costlist[num]
            costs[r] = costlist
            total[r] = np.sum(costlist_weighted)
    flags[costed] |= FLAG_COSTED
    flags[cand] |= FLAG_CANDIDATE

    # ---- planner.py:329-392 in cost order (stable sort => ties by row index)
    order = [r for r in np.argsort(total, kind="stable") if cand[r]]
    winner = -1
    collision_counter = 0
    for r in order:
        st = states[:, r, :]
        hit_k = first_collision_step(st, prm, predictions, Nt) if (len(predictions) and collision_check) else -1
        off_k = collides_with_static(st, prm, static_obbs, Nt)
        hit, off = hit_k >= 0, off_k >= 0
        if hit:
            flags[r] |= FLAG_COLLIDE | (hit_k << COLLIDE_STEP_SHIFT)
        if off:
            flags[r] |= FLAG_BOUNDARY | (off_k << BOUNDARY_STEP_SHIFT)
        if winner < 0:
            if hit:
                collision_counter += 1
            if not hit and not off:
                winner = int(r)
                if not check_all_collisions:
                    break
    return dict(states=states, flags=flags, traj_len=traj_len, costs=costs, total=total,
                coeffs=coeffs, margins=margins, reason_counts=reason_counts, n_in_list=n_list, n_feasible=n_feasible,
                percentage=percentage, argmin=winner,
                min_cost=(float(total[winner]) if winner >= 0 else float("inf")),
                collision_counter=collision_counter, cost_names=names)
