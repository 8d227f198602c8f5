"""Build tests/golden/tjunction.npz from the reference's example scenario
/root/reference/example_scenarios/ZAM_Tjunction-1_42_T-1.xml (BASELINE.json configs[0]/[2]).

Run once in the build container (the XML is not available on the GPU box).  Only *inputs* are
extracted: the planning problem's initial state, the centre line of the lanelet chain
50195 -> 50209 -> 50203, and the five dynamic obstacles' state lists.  The reference
path goes through the reference's own ``extend_ref_path_both_ends`` and ``smooth_ref_path``
(cr_scenario_handler/utils/utils_coordinate_system.py:54-58,110-134, imported unmodified through
ref_stubs; ``resample_polyline`` of the un-vendored commonroad_dc is stood in for by a plain
arc-length resampler).  Predictions are the ground-truth form of
cr_scenario_handler/utils/prediction_helpers.py:207-257 (+ the 0.5 / 0.2 m safety margins of :167-170).
"""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_stubs  # noqa: E402

ref_stubs.install()
import commonroad_dc.geometry.util as cdc_util  # noqa: E402  (stub module)


def resample_polyline(polyline, step=2.0):
    """Arc-length resampling (stand-in for commonroad_dc.geometry.util.resample_polyline)."""
    polyline = np.asarray(polyline, dtype=float)
    seg = np.sqrt(np.sum(np.diff(polyline, axis=0) ** 2, axis=1))
    L = np.concatenate(([0.0], np.cumsum(seg)))
    n = int(np.floor(L[-1] / step))
    target = np.concatenate((np.arange(n + 1) * step, [L[-1]])) if L[-1] - n * step > 1e-9 else np.arange(n + 1) * step
    return np.stack([np.interp(target, L, polyline[:, 0]), np.interp(target, L, polyline[:, 1])], axis=1)


cdc_util.resample_polyline = resample_polyline
from cr_scenario_handler.utils.utils_coordinate_system import smooth_ref_path, extend_ref_path_both_ends  # noqa: E402
from frenetix_motion_planner_b200 import synthetic as syn  # noqa: E402

XML = "/root/reference/example_scenarios/ZAM_Tjunction-1_42_T-1.xml"


def pts(node):
    return np.array([[float(p.find("x").text), float(p.find("y").text)] for p in node.findall("point")])


def main():
    root = ET.parse(XML).getroot()
    lanelets = {l.attrib["id"]: l for l in root.findall("lanelet")}
    chain = ["50195", "50209", "50203"]
    centre = []
    for lid in chain:
        l = lanelets[lid]
        c = 0.5 * (pts(l.find("leftBound")) + pts(l.find("rightBound")))
        centre.append(c if not centre else c[1:])
    centre = np.vstack(centre)
    # commonroad-route-planner hands over a densely resampled centre line (smooth_ref_path assumes ~0.125 m
    # spacing, utils_coordinate_system.py:117-121); stand-in: arc-length resampling of the lanelet centre line
    route = resample_polyline(centre, 0.125)
    ref = smooth_ref_path(extend_ref_path_both_ends(route))

    ini = root.find("planningProblem").find("initialState")
    f = lambda path: float(ini.find(path).text)
    pos = np.array([f("position/point/x"), f("position/point/y")])
    th, v, a, yr = f("orientation/exact"), f("velocity/exact"), f("acceleration/exact"), f("yawRate/exact")
    wb_rear = syn.VEHICLE_2["wb_rear_axle"]
    rear = pos - wb_rear * np.array([np.cos(th), np.sin(th)])          # state.py:43-66 centre -> rear axle

    obs = []
    for d in root.findall("dynamicObstacle"):
        states = [d.find("initialState")] + d.find("trajectory").findall("state")
        arr = np.array([[float(s.find("position/point/x").text), float(s.find("position/point/y").text),
                         float(s.find("orientation/exact").text), float(s.find("velocity/exact").text)] for s in states])
        shape = d.find("shape/rectangle")
        obs.append((int(d.attrib["id"]), arr, float(shape.find("length").text), float(shape.find("width").text)))

    out = dict(reference_path=ref, centre_line=centre, ego_position_center=pos, ego_position_rear=rear,
               ego_orientation=th, ego_velocity=v, ego_acceleration=a, ego_yaw_rate=yr,
               obstacle_ids=np.array([o[0] for o in obs]),
               obstacle_states=np.stack([o[1] for o in obs]),                  # [5, 148, 4] x, y, theta, v
               obstacle_shapes=np.array([[o[2], o[3]] for o in obs]))
    np.savez_compressed(os.path.join(HERE, "tjunction.npz"), **out)
    print("reference path", ref.shape, "length", np.sum(np.sqrt(np.sum(np.diff(ref, axis=0) ** 2, axis=1))),
          "ego", rear, th, v, "obstacles", out["obstacle_states"].shape)


if __name__ == "__main__":
    main()
