"""Known-answer tests that do NOT go through the oracle's own definitions: the third-party pieces the
reference calls into (pycrcc OBB overlap / obb-sum hulls, the CCosy point conversion) are restated in oracle/ and in the
CUDA kernels, so here they are pinned against independent mathematics instead:

  * OBB-vs-OBB overlap against a brute-force convex-polygon intersection test (corner containment + edge crossing,
    no separating axes) on 10^5 random box pairs;
  * the obb-sum hull against its defining properties (contains both boxes, is tight on all four sides in the frame of
    the first box);
  * the Frenet -> Cartesian map against closed forms on a straight line and on a circular arc, and against its own
    inverse ((x, y) -> (s, d) -> (x, y) round trip);
  * the per-step back-projection of check_feasibility (reactive_planner.py:389-478) against the closed forms that hold
    on a straight reference: x = s, y = d, theta = atan(d'/s'), v = s'/cos(theta), kappa = d'' cos^3(theta).
"""
import math

import numpy as np
import pytest

from oracle import frenet_oracle as fo
from frenetix_motion_planner_b200 import synthetic as syn
from frenetix_motion_planner_b200.coordinate_system import CoordinateSystem


# ---------------------------------------------------------------------------------------------------------------
# brute-force rectangle intersection: closed sets, no separating-axis reasoning
# ---------------------------------------------------------------------------------------------------------------
def corners(cx, cy, ux, uy, ha, hb):
    vx, vy = -uy, ux
    return np.array([[cx + sa * ha * ux + sb * hb * vx, cy + sa * ha * uy + sb * hb * vy]
                     for sa, sb in ((1, 1), (-1, 1), (-1, -1), (1, -1))])


def _inside(pt, box):
    cx, cy, ux, uy, ha, hb = box
    dx, dy = pt[0] - cx, pt[1] - cy
    return abs(dx * ux + dy * uy) <= ha and abs(dy * ux - dx * uy) <= hb


def _orient(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _segments_cross(p1, p2, p3, p4):
    d1, d2, d3, d4 = _orient(p3, p4, p1), _orient(p3, p4, p2), _orient(p1, p2, p3), _orient(p1, p2, p4)
    return ((d1 > 0) != (d2 > 0)) and ((d3 > 0) != (d4 > 0))


def brute_force_overlap(e, o):
    ce, co = corners(*e), corners(*o)
    if any(_inside(p, o) for p in ce) or any(_inside(p, e) for p in co):
        return True
    for i in range(4):
        for j in range(4):
            if _segments_cross(ce[i], ce[(i + 1) % 4], co[j], co[(j + 1) % 4]):
                return True
    return False


def _sat_margin(e, o):
    """Smallest |gap| over the four axes: pairs this close to touching are decided by rounding, not geometry."""
    ecx, ecy, eux, euy, eha, ehb = e
    ocx, ocy, oux, ouy, oha, ohb = o
    dx, dy = ocx - ecx, ocy - ecy
    c, sn = abs(eux * oux + euy * ouy), abs(eux * ouy - euy * oux)
    gaps = [abs(dx * eux + dy * euy) - (eha + oha * c + ohb * sn), abs(dy * eux - dx * euy) - (ehb + oha * sn + ohb * c),
            abs(dx * oux + dy * ouy) - (oha + eha * c + ehb * sn), abs(dy * oux - dx * ouy) - (ohb + eha * sn + ehb * c)]
    return min(abs(g) for g in gaps)


def _random_box(rng, spread):
    th = rng.uniform(-math.pi, math.pi)
    return (rng.uniform(-spread, spread), rng.uniform(-spread, spread), math.cos(th), math.sin(th),
            rng.uniform(0.3, 4.0), rng.uniform(0.3, 2.0))


def test_obb_overlap_equals_brute_force_polygon_intersection():
    rng = np.random.default_rng(20260117)
    n, n_hit, skipped = 100_000, 0, 0
    for k in range(n):
        e, o = _random_box(rng, 6.0), _random_box(rng, 6.0)
        if _sat_margin(e, o) < 1e-9:
            skipped += 1
            continue
        got = fo.obb_overlap(e, o)
        assert got == brute_force_overlap(e, o), (k, e, o)
        n_hit += got
    assert skipped == 0 and 0.15 * n < n_hit < 0.85 * n         # the sample exercises both outcomes


def test_obb_overlap_special_configurations():
    unit = (0.0, 0.0, 1.0, 0.0, 2.0, 1.0)
    assert fo.obb_overlap(unit, (4.0, 0.0, 1.0, 0.0, 2.0, 1.0))               # touching edge to edge = overlap
    assert not fo.obb_overlap(unit, (4.0 + 1e-9, 0.0, 1.0, 0.0, 2.0, 1.0))
    assert fo.obb_overlap(unit, (0.2, 0.1, math.cos(0.7), math.sin(0.7), 0.3, 0.2))     # fully inside
    assert fo.obb_overlap((0.2, 0.1, math.cos(0.7), math.sin(0.7), 0.3, 0.2), unit)
    # a cross: no corner of either box lies in the other, only the edges intersect
    assert fo.obb_overlap((0.0, 0.0, 1.0, 0.0, 5.0, 0.2), (0.0, 0.0, 0.0, 1.0, 5.0, 0.2))
    # corners close, separated only by a diagonal axis of the rotated box
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    assert not fo.obb_overlap((0.0, 0.0, 1.0, 0.0, 1.0, 1.0), (2.5, 2.5, c, s, 1.0, 1.0))
    assert not brute_force_overlap((0.0, 0.0, 1.0, 0.0, 1.0, 1.0), (2.5, 2.5, c, s, 1.0, 1.0))


def test_obb_sum_hull_contains_both_boxes_and_is_tight():
    rng = np.random.default_rng(7)
    for _ in range(20_000):
        hl, hw = rng.uniform(0.5, 3.0), rng.uniform(0.3, 1.5)
        c0 = rng.uniform(-5, 5, 2)
        th0, th1 = rng.uniform(-math.pi, math.pi), rng.uniform(-math.pi, math.pi)
        c1 = c0 + rng.uniform(-3, 3, 2)
        hx, hy, ux, uy, ha, hb = fo.obb_sum_hull(c0[0], c0[1], th0, c1[0], c1[1], th1, hl, hw)
        assert (ux, uy) == (math.cos(th0), math.sin(th0))                                   # frame of box k
        pts = np.vstack([corners(c0[0], c0[1], math.cos(th0), math.sin(th0), hl, hw),
                         corners(c1[0], c1[1], math.cos(th1), math.sin(th1), hl, hw)])
        pu = (pts[:, 0] - hx) * ux + (pts[:, 1] - hy) * uy
        pv = (pts[:, 1] - hy) * ux - (pts[:, 0] - hx) * uy
        tol = 1e-12 * (1 + np.abs(pts).max())
        assert pu.max() <= ha + tol and pu.min() >= -ha - tol and pv.max() <= hb + tol and pv.min() >= -hb - tol
        # minimal: every side of the hull is touched by a corner
        assert abs(pu.max() - ha) <= tol and abs(pu.min() + ha) <= tol
        assert abs(pv.max() - hb) <= tol and abs(pv.min() + hb) <= tol


# ---------------------------------------------------------------------------------------------------------------
# Frenet -> Cartesian map
# ---------------------------------------------------------------------------------------------------------------
def _refpath(poly):
    cs = CoordinateSystem(poly)
    return cs, fo.RefPath(cs.ref_pos, cs.ref_theta, cs.ref_curv, cs.ref_curv_d, np.ascontiguousarray(poly[:, 0]),
                          np.ascontiguousarray(poly[:, 1]))


def test_frenet_to_cartesian_on_a_straight_line_is_the_identity():
    cs, ref = _refpath(syn.straight_polyline(200))
    rng = np.random.default_rng(1)
    for _ in range(2000):
        s, d = rng.uniform(0, 198.9), rng.uniform(-5, 5)
        assert np.allclose(fo.ccosy_to_cartesian(ref, s, d), [s, d], rtol=0, atol=1e-12)
        assert np.allclose(cs.convert_to_cartesian_coords(s, d), [s, d], rtol=0, atol=1e-12)
    assert fo.ccosy_to_cartesian(ref, -0.1, 0.0) is None and fo.ccosy_to_cartesian(ref, 199.0, 0.0) is None
    assert np.all(cs.ref_theta == 0) and np.all(cs.ref_curv == 0) and np.all(cs.ref_curv_d == 0)


def test_frenet_to_cartesian_on_an_arc_matches_the_circle():
    R = 200.0
    poly = syn.arc_polyline(R=R, M=600)                 # vertices ON the circle, 1 m apart along it
    cs, ref = _refpath(poly)
    chord = 2 * R * math.sin(0.5 / R)                   # polyline arclength per vertex step
    sagitta = R * (1 - math.cos(0.5 / R))               # how far a chord runs inside the circle
    rng = np.random.default_rng(2)
    for _ in range(2000):
        s, d = rng.uniform(1.0, 590.0), rng.uniform(-4, 4)
        phi = (s / chord) / R                           # the same fraction of the way round
        exact = np.array([(R - d) * math.sin(phi), R - (R - d) * math.cos(phi)])
        got = fo.ccosy_to_cartesian(ref, s, d)
        # a polyline vertex lies exactly on the circle; between vertices the chord is at most `sagitta` inside it; the
        # heading table holds the heading of the OUTGOING chord (half a step ahead of the tangent at the vertex), so the
        # normal is turned by at most one step angle 1 / R
        assert np.linalg.norm(got - exact) <= sagitta + abs(d) * (1.0 / R) + 1e-9
    # table values: constant curvature 1/R (interior), heading = chord heading
    assert np.allclose(cs.ref_curv[2:-2], 1.0 / R, rtol=1e-4)
    assert np.allclose(np.diff(cs.ref_theta)[:-1], 1.0 / R, rtol=1e-9)


@pytest.mark.parametrize("poly", [syn.arc_polyline(R=30.0, M=120), syn.scurve_polyline(M=300),
                                  syn.arc_polyline(R=200.0, M=600, start_heading=2.9)],
                         ids=["tight_arc", "s_curve", "arc_across_pi"])
def test_cartesian_to_frenet_round_trip(poly):
    cs, ref = _refpath(poly)
    rng = np.random.default_rng(3)
    L = cs.ref_pos[-1]
    for _ in range(3000):
        s, d = rng.uniform(0.5, L - 0.5), rng.uniform(-4, 4)
        X = fo.ccosy_to_cartesian(ref, s, d)
        s2, d2 = cs.convert_to_curvilinear_coords(X[0], X[1])
        assert abs(s2 - s) < 1e-8 and abs(d2 - d) < 1e-8
        assert np.allclose(fo.ccosy_to_cartesian(ref, s2, d2), X, rtol=0, atol=1e-9)


def test_initial_frenet_position_against_a_brute_force_projection():
    """x_cl[0][0], x_cl[1][0] of the ZAM_Tjunction fixture: a dense scan of the forward map, independent of the Newton
    inverse the product uses (planner.py:567-571 asks CCosy for this projection)."""
    t = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "tjunction.npz"))
    cs, ref = _refpath(t["reference_path"])
    X = t["ego_position_rear"]
    s_prod, d_prod = cs.convert_to_curvilinear_coords(X[0], X[1])
    # scan: for every s on a 1 mm raster the offset that would be needed and the tangential miss; refine the best by bisection
    ss = np.arange(cs.ref_pos[0] + 1e-6, cs.ref_pos[-1] - 1e-3, 1e-3)
    i = np.searchsorted(cs.ref_pos, ss, side="right") - 1
    lam = (ss - cs.ref_pos[i]) / (cs.ref_pos[i + 1] - cs.ref_pos[i])
    P = (1 - lam)[:, None] * t["reference_path"][i] + lam[:, None] * t["reference_path"][i + 1]
    th = cs.ref_theta[i] + lam * (cs.ref_theta[i + 1] - cs.ref_theta[i])
    miss = (X[0] - P[:, 0]) * np.cos(th) + (X[1] - P[:, 1]) * np.sin(th)
    near = np.hypot(X[0] - P[:, 0], X[1] - P[:, 1]) < 10.0
    k = np.flatnonzero(near[:-1] & (np.sign(miss[:-1]) != np.sign(miss[1:])))
    assert k.size == 1
    lo, hi = ss[k[0]], ss[k[0] + 1]

    def tangential(s):
        p = fo.ccosy_to_cartesian(ref, s, 0.0)
        j = int(np.argmax(cs.ref_pos > s)) - 1
        l = (s - cs.ref_pos[j]) / (cs.ref_pos[j + 1] - cs.ref_pos[j])
        a = cs.ref_theta[j] + l * (cs.ref_theta[j + 1] - cs.ref_theta[j])
        return (X[0] - p[0]) * math.cos(a) + (X[1] - p[1]) * math.sin(a), (X[1] - p[1]) * math.cos(a) - (X[0] - p[0]) * math.sin(a)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if (tangential(mid)[0] > 0) == (tangential(lo)[0] > 0):
            lo = mid
        else:
            hi = mid
    s_bf = 0.5 * (lo + hi)
    d_bf = tangential(s_bf)[1]
    assert abs(s_prod - s_bf) < 1e-9 and abs(d_prod - d_bf) < 1e-9


# ---------------------------------------------------------------------------------------------------------------
# back-projection closed forms on a straight reference (no curvature terms survive)
# ---------------------------------------------------------------------------------------------------------------
def _straight_case():
    cs, ref = _refpath(syn.straight_polyline(400))
    prm = fo.Params(x0_orientation=0.05, desired_velocity=9.0, draw_traj_set=True, kinematic_debug=True)
    x_cl = ([12.0, 7.0, 0.3], [0.4, 0.2, -0.1])
    S = syn.grid_sampling_matrix(np.array([1.1, 2.0, 3.0]), np.linspace(2.0, 12.0, 6), np.linspace(-2.5, 2.5, 7), x_cl)
    return ref, prm, S


def _assert_straight_closed_forms(st, stored):
    x, y, th, v, a, kap = (st[k][stored] for k in (fo.F_X, fo.F_Y, fo.F_THETA, fo.F_V, fo.F_A, fo.F_KAPPA))
    s, d, thc, sd, sdd, dd, ddd = (st[k][stored] for k in (fo.F_S, fo.F_D, fo.F_THETA_CL, fo.F_S_DOT, fo.F_S_DDOT,
                                                            fo.F_D_DOT, fo.F_D_DDOT))
    assert np.abs(x - s).max() < 1e-9 and np.abs(y - d).max() < 1e-12
    assert np.abs(th - np.arctan2(dd, sd)).max() < 1e-12 and np.abs(th - thc).max() < 1e-15
    assert np.abs(v - np.hypot(sd, dd)).max() < 1e-10                       # s' / cos(theta) = sqrt(s'^2 + d'^2)
    dpp = (ddd - (dd / sd) * sdd) / sd ** 2                                  # d'' w.r.t. arclength
    assert np.abs(kap - dpp * np.cos(th) ** 3).max() < 1e-10
    # a = s'' / cos + s'^2 / cos * tan * kappa / cos ... on a straight line: dv/dt of v = sqrt(s'^2 + d'^2)
    assert np.abs(a - (sd * sdd + dd * ddd) / np.hypot(sd, dd)).max() < 1e-9


def test_back_projection_closed_forms_on_a_straight_reference_oracle():
    ref, prm, S = _straight_case()
    out = fo.plan(S, ref, prm, [])
    stored = (out["flags"] & fo.FLAG_STORED) != 0
    assert stored.sum() > 50
    _assert_straight_closed_forms(out["states"], stored)
    # end conditions of the polynomials (polynomial_trajectory.py:293-343,452-488): reached at t = T (sample T / dt)
    for r in np.flatnonzero(stored)[:40]:
        T, k = S[r, 1], int(round(S[r, 1] / prm.dt))
        if k < prm.N + 1:
            assert abs(out["states"][fo.F_S_DOT][r, k] - S[r, 5]) < 1e-9 and abs(out["states"][fo.F_S_DDOT][r, k]) < 1e-9
            assert abs(out["states"][fo.F_D][r, k] - S[r, 10]) < 1e-9
            assert abs(out["states"][fo.F_D_DOT][r, k]) < 1e-9 and abs(out["states"][fo.F_D_DDOT][r, k]) < 1e-8


@pytest.mark.gpu
def test_back_projection_closed_forms_on_a_straight_reference_device():
    from helpers import device_plan
    ref, prm, S = _straight_case()
    dev = device_plan(S, ref, prm, [])
    stored = (dev["flags"] & fo.FLAG_STORED) != 0
    assert stored.sum() > 50
    _assert_straight_closed_forms(dev["states"], stored)


@pytest.mark.gpu
def test_device_collision_flags_equal_brute_force_polygon_intersection():
    """The CUDA sweep against the brute-force rectangle test (no oracle in between): one static box per case and the ego
    hulls rebuilt on the host from the device's own x, y, theta."""
    from helpers import device_plan
    ref, prm, S = _straight_case()
    rng = np.random.default_rng(5)
    n_checked = 0
    for case in range(6):
        wall = np.array([[rng.uniform(20, 45), rng.uniform(-4, 4), rng.uniform(-1.5, 1.5), rng.uniform(1, 6), rng.uniform(0.3, 1.5)]])
        dev = device_plan(S, ref, prm, [], static_obbs=wall)
        cand = (dev["flags"] & fo.FLAG_CANDIDATE) != 0
        o = (wall[0, 0], wall[0, 1], math.cos(wall[0, 2]), math.sin(wall[0, 2]), wall[0, 3], wall[0, 4])
        for r in np.flatnonzero(cand):
            x, y, th = dev["states"][fo.F_X][r], dev["states"][fo.F_Y][r], dev["states"][fo.F_THETA][r]
            cx, cy = x + prm.wb_rear_axle * np.cos(th), y + prm.wb_rear_axle * np.sin(th)
            first, tie = -1, False
            for k in range(prm.N):
                e = fo.obb_sum_hull(cx[k], cy[k], th[k], cx[k + 1], cy[k + 1], th[k + 1], prm.length / 2, prm.width / 2)
                if _sat_margin(e, o) < 1e-9:
                    tie = True
                    break
                if brute_force_overlap(e, o):
                    first = k
                    break
            if tie:
                continue
            hit = bool(dev["flags"][r] & fo.FLAG_BOUNDARY)
            assert hit == (first >= 0), (case, r)
            if hit:
                assert int((dev["flags"][r] >> fo.BOUNDARY_STEP_SHIFT) & 63) == first
            n_checked += 1
    assert n_checked > 300
