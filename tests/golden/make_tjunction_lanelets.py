"""Build tests/golden/tjunction_lanelets.json (bounds and adjacency of the 12 lanelets of
/root/reference/example_scenarios/ZAM_Tjunction-1_42_T-1.xml) for the road-boundary tests.  Run once in the build
container; inputs only, read with xml.etree (no commonroad-io needed)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from frenetix_motion_planner_b200.road_boundary import lanelets_from_commonroad_xml  # noqa: E402

XML = "/root/reference/example_scenarios/ZAM_Tjunction-1_42_T-1.xml"

if __name__ == "__main__":
    ll = lanelets_from_commonroad_xml(XML)
    out = {str(k): dict(left=v["left"].tolist(), right=v["right"].tolist(), adj_left=v["adj_left"], adj_right=v["adj_right"])
           for k, v in ll.items()}
    with open(os.path.join(HERE, "tjunction_lanelets.json"), "w") as f:
        json.dump(out, f)
    print(len(out), "lanelets")
