"""CPU tests of the road-boundary construction: outer bounds of the lanelet network as thin walls,
on the shipped ZAM_Tjunction-1_42_T-1 network (tests/golden/tjunction_lanelets.json, made by
tests/golden/make_tjunction_lanelets.py)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from frenetix_motion_planner_b200.road_boundary import road_boundary_obbs, footprint_hits_walls, lanelets_from_network

HALF_LEN, HALF_WID = 4.508 / 2, 1.610 / 2     # vehicle 2 (BMW 320i)


@pytest.fixture(scope="module")
def network():
    raw = json.load(open(os.path.join(GOLDEN_DIR, "tjunction_lanelets.json")))
    return {int(k): dict(left=np.array(v["left"]), right=np.array(v["right"]), adj_left=v["adj_left"], adj_right=v["adj_right"])
            for k, v in raw.items()}


@pytest.fixture(scope="module")
def walls(network):
    return road_boundary_obbs(network, merge_tol=0.0)          # one wall per outside segment


@pytest.fixture(scope="module")
def merged(network):
    return road_boundary_obbs(network)                          # default: collinear runs merged


def centre(l):
    return 0.5 * (l["left"] + l["right"])


def headings(c):
    d = np.diff(c, axis=0)
    th = np.arctan2(d[:, 1], d[:, 0])
    return np.concatenate([th, th[-1:]])


def test_walls_lie_on_lanelet_bounds(network, walls):
    assert walls.shape[1] == 5 and walls.shape[0] > 50
    segs = []
    for l in network.values():
        for side in ("left", "right"):
            v = l[side]
            segs += [(0.5 * (v[k] + v[k + 1]), 0.5 * np.linalg.norm(v[k + 1] - v[k])) for k in range(len(v) - 1)]
    mids = np.array([m for m, _ in segs]); halfs = np.array([h for _, h in segs])
    for w in walls:
        k = np.argmin(np.linalg.norm(mids - w[:2], axis=1))
        assert np.linalg.norm(mids[k] - w[:2]) < 1e-9 and abs(halfs[k] - w[3]) < 1e-9


def test_no_wall_inside_the_road(network, walls):
    """Shared bounds (left bounds between opposite lanes, connector bounds that cross the junction) carry no wall: driving
    along any lanelet's centre line, the junction connectors included, never touches one."""
    for l in network.values():
        c = centre(l)
        for (x, y), th in zip(c, headings(c)):
            assert not footprint_hits_walls(x, y, th, HALF_LEN, HALF_WID, walls)


def test_leaving_the_road_hits_a_wall_and_changing_lane_does_not(network, walls):
    l = network[50195]                      # the ego's start lanelet: straight, opposite lane 50197 on its left
    c, th = centre(l), headings(centre(l))
    k = len(c) // 2
    left = np.array([-np.sin(th[k]), np.cos(th[k])])
    width = np.linalg.norm(l["left"][k] - l["right"][k])
    off_right = c[k] - left * (0.5 * width + 0.2)            # footprint centre just beyond the right bound
    off_left = c[k] + left * width                            # centre of the opposite lane
    assert footprint_hits_walls(off_right[0], off_right[1], th[k], HALF_LEN, HALF_WID, walls)
    assert not footprint_hits_walls(off_left[0], off_left[1], th[k], HALF_LEN, HALF_WID, walls)
    far_left = c[k] + left * (1.5 * width + 0.2)              # beyond the opposite lane's outer bound
    assert footprint_hits_walls(far_left[0], far_left[1], th[k], HALF_LEN, HALF_WID, walls)


def test_outer_boundary_is_closed_along_the_arms(network, walls):
    """Every right-bound segment of the three straight arms is an outer segment (no right neighbours in this network)."""
    for lid in (50195, 50197, 50199, 50201, 50203, 50205):
        v = network[lid]["right"]
        mids = 0.5 * (v[:-1] + v[1:])
        for m in mids:
            assert np.min(np.linalg.norm(walls[:, :2] - m, axis=1)) < 1e-9


def test_duck_typed_lanelet_network_gives_the_same_walls(network, walls):
    class L:
        def __init__(self, lid, d):
            self.lanelet_id, self.left_vertices, self.right_vertices = lid, d["left"], d["right"]
            self.adj_left, self.adj_right = d["adj_left"], d["adj_right"]

    class Net:
        lanelets = [L(k, v) for k, v in network.items()]

    assert np.array_equal(road_boundary_obbs(lanelets_from_network(Net()), merge_tol=0.0), walls)


def test_planner_builds_the_boundary_from_a_scenario_once(network, walls):
    """ReactivePlannerB200.set_scenario mirrors planner.py:550-565 (boundary built once per scenario)."""
    from frenetix_motion_planner_b200.reactive_planner_b200 import ReactivePlannerB200

    class L:
        def __init__(self, lid, d):
            self.lanelet_id, self.left_vertices, self.right_vertices = lid, d["left"], d["right"]
            self.adj_left, self.adj_right = d["adj_left"], d["adj_right"]

    class Net:
        lanelets = [L(k, v) for k, v in network.items()]

    class Scenario:
        lanelet_network = Net()

    p = ReactivePlannerB200.__new__(ReactivePlannerB200)      # host-side logic only: no device handle needed
    p.static_obbs = None
    p.set_scenario(Scenario())
    assert np.array_equal(p.static_obbs, road_boundary_obbs(network))
    p.static_obbs = walls[:3]
    p.set_scenario(Scenario())                                 # already set: untouched
    assert p.static_obbs.shape[0] == 3
    p.set_road_boundary(network, wall_half_width=0.1, merge_tol=0.0)
    assert p.static_obbs.shape == walls.shape and np.all(p.static_obbs[:, 4] == 0.1)


def test_merged_walls_are_fewer_and_decide_the_same(network, walls, merged):
    """Collinear runs collapse (the straight arms), the total length stays, and on a dense probe of footprints around the
    network the merged walls give the same hit / no-hit answer as the per-segment walls."""
    assert merged.shape[0] <= 0.75 * walls.shape[0]
    assert abs(merged[:, 3].sum() - walls[:, 3].sum()) < 0.05
    rng = np.random.default_rng(3)
    n_hit = 0
    for l in network.values():
        c = centre(l)
        th = headings(c)
        for k in range(0, len(c), 2):
            left = np.array([-np.sin(th[k]), np.cos(th[k])])
            for off in rng.uniform(-6.0, 6.0, 6):
                x, y = c[k] + left * off
                a = th[k] + rng.normal(0, 0.2)
                h0 = footprint_hits_walls(x, y, a, HALF_LEN, HALF_WID, walls)
                h1 = footprint_hits_walls(x, y, a, HALF_LEN, HALF_WID, merged)
                assert h0 == h1
                n_hit += h0
    assert n_hit > 50
