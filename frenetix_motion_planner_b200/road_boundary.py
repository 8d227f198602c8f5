"""Road boundary as static boxes for the collision sweep.

The reference builds the boundary once per scenario with the un-vendored drivability checker
(``commonroad_dc.boundary.create_road_boundary_obstacle(scenario, method='aligned_triangulation', axis=2)``,
frenetix_motion_planner/planner.py:550-565) and tests every candidate against it
(``trajectories_collision_static_obstacles``, planner.py:362-368).  That library is neither in the reference tree nor
installable offline, so -- like the other third-party pieces -- the construction is DEFINED here
and flagged parity-unpinned:

    the boundary is a chain of thin oriented boxes ("walls"), one per segment of every lanelet bound that is on the
    OUTSIDE of the road: a bound segment is outside iff a probe point just beyond it (to the right of a right bound,
    to the left of a left bound) lies in no lanelet polygon of the network.

A candidate leaves the road iff the swept hull of its footprint (obb-sum of two consecutive boxes, collision_check.py
:147-181) meets a wall -- the same exact SAT test the kernel runs for static boxes (`frx_set_static_obbs`).  Walls are
`[cx, cy, theta, half_len, half_wid]` rows, the format of `ReactivePlannerB200.set_static_obstacles`.

Inputs are plain arrays (no commonroad import needed): `lanelets_from_commonroad_xml` reads a CommonRoad 2020a file with
xml.etree, `lanelets_from_network` duck-types a `commonroad.scenario.lanelet.LaneletNetwork`.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import Dict, Optional

import numpy as np

Lanelets = Dict[int, dict]   # id -> {"left": [n,2], "right": [n,2], "adj_left": id|None, "adj_right": id|None}


def lanelets_from_commonroad_xml(path: str) -> Lanelets:
    def pts(node):
        return np.array([[float(p.find("x").text), float(p.find("y").text)] for p in node.findall("point")], dtype=np.float64)

    out: Lanelets = {}
    for l in ET.parse(path).getroot().findall("lanelet"):
        al, ar = l.find("adjacentLeft"), l.find("adjacentRight")
        out[int(l.attrib["id"])] = dict(left=pts(l.find("leftBound")), right=pts(l.find("rightBound")),
                                        adj_left=None if al is None else int(al.attrib["ref"]),
                                        adj_right=None if ar is None else int(ar.attrib["ref"]))
    return out


def lanelets_from_network(lanelet_network) -> Lanelets:
    """From a commonroad LaneletNetwork (lanelet.left_vertices / right_vertices / adj_left / adj_right)."""
    return {int(l.lanelet_id): dict(left=np.asarray(l.left_vertices, dtype=np.float64), right=np.asarray(l.right_vertices, dtype=np.float64),
                                    adj_left=getattr(l, "adj_left", None), adj_right=getattr(l, "adj_right", None))
            for l in lanelet_network.lanelets}


def _inside_any(points: np.ndarray, polygons) -> np.ndarray:
    """Even-odd point-in-polygon of `points` [m,2] against every polygon; True if inside at least one."""
    inside = np.zeros(points.shape[0], dtype=bool)
    px, py = points[:, 0][:, None], points[:, 1][:, None]
    for poly in polygons:
        x0, y0 = poly[:, 0][None, :], poly[:, 1][None, :]
        x1, y1 = np.roll(poly[:, 0], -1)[None, :], np.roll(poly[:, 1], -1)[None, :]
        crosses = (y0 > py) != (y1 > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = x0 + (py - y0) * (x1 - x0) / (y1 - y0)
        inside |= (np.sum(crosses & (px < xi), axis=1) % 2) == 1
    return inside


def _merge_run(pts: np.ndarray, tol: float):
    """Greedy polyline simplification of one run of outside segments: chords p[a] -> p[b] such that every skipped
    vertex stays within `tol` of the chord.  tol = 0 keeps every segment."""
    out, a, n = [], 0, len(pts)
    while a < n - 1:
        b = a + 1
        while tol > 0 and b + 1 < n:
            d = pts[b + 1] - pts[a]
            L = np.hypot(d[0], d[1])
            if L <= 0:
                break
            rel = pts[a + 1:b + 1] - pts[a]
            dev = np.abs(rel[:, 0] * d[1] - rel[:, 1] * d[0]) / L
            if np.max(dev) > tol:
                break
            b += 1
        out.append((pts[a], pts[b]))
        a = b
    return out


def road_boundary_obbs(lanelets: Lanelets, wall_half_width: float = 0.05, probe: float = 0.25,
                       min_length: float = 1e-6, merge_tol: float = 0.02) -> np.ndarray:
    """Walls `[cx, cy, theta, half_len, half_wid]` along the outer bounds of the lanelet network.

    wall_half_width: half thickness of a wall (the wall is centred ON the bound)
    probe:           how far beyond the bound the outside test looks (m); smaller than any lane width
    merge_tol:       consecutive outside segments of a bound are merged into one wall while the skipped vertices stay
                     within this distance of it (m); 0 = one wall per segment.  Every wall costs one broad-phase test
                     per candidate and step in the collision sweep, and straight arms are many collinear segments.
    """
    polygons = [np.vstack([l["left"], l["right"][::-1]]) for l in lanelets.values()]
    rows = []
    for l in lanelets.values():
        for side, sign in (("right", -1.0), ("left", +1.0)):     # outward normal: right of a right bound, left of a left bound
            v = np.asarray(l[side], dtype=np.float64)
            if len(v) < 2:
                continue
            a, b = v[:-1], v[1:]
            d = b - a
            length = np.hypot(d[:, 0], d[:, 1])
            ok = length > min_length
            t = d / np.where(ok, length, 1.0)[:, None]
            n = np.stack([-t[:, 1], t[:, 0]], axis=1) * sign      # left normal * sign
            outside = ok & ~_inside_any(0.5 * (a + b) + probe * n, polygons)
            k = 0
            while k < len(outside):                               # runs of consecutive outside segments
                if not outside[k]:
                    k += 1
                    continue
                e = k
                while e + 1 < len(outside) and outside[e + 1]:
                    e += 1
                for p0, p1 in _merge_run(v[k:e + 2], merge_tol):
                    dd = p1 - p0
                    rows.append([0.5 * (p0[0] + p1[0]), 0.5 * (p0[1] + p1[1]), np.arctan2(dd[1], dd[0]), 0.5 * np.hypot(dd[0], dd[1]),
                                 wall_half_width])
                k = e + 1
    return np.asarray(rows, dtype=np.float64).reshape(-1, 5)


def footprint_hits_walls(x: float, y: float, theta: float, half_len: float, half_wid: float, walls: np.ndarray) -> bool:
    """Host-side exact SAT of one oriented box against the walls (the kernel's static-box test, for tests and tools)."""
    if walls.size == 0:
        return False
    ux, uy = np.cos(theta), np.sin(theta)
    wux, wuy = np.cos(walls[:, 2]), np.sin(walls[:, 2])
    dx, dy = walls[:, 0] - x, walls[:, 1] - y
    c = np.abs(ux * wux + uy * wuy)
    s = np.abs(ux * wuy - uy * wux)
    ha, hb = walls[:, 3], walls[:, 4]
    sep = (np.abs(dx * ux + dy * uy) > half_len + (ha * c + hb * s)) | (np.abs(dy * ux - dx * uy) > half_wid + (ha * s + hb * c)) | \
          (np.abs(dx * wux + dy * wuy) > ha + (half_len * c + half_wid * s)) | (np.abs(dy * wux - dx * wuy) > hb + (half_len * s + half_wid * c))
    return bool((~sep).any())
