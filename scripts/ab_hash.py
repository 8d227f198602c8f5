"""Print a digest of every output of a few plans (A/B builds must print identical lines): FRX_LIB=... python scripts/ab_hash.py"""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, device_plan, GOLDEN_CASES
from frenetix_motion_planner_b200 import synthetic as syn

def digest(dev):
    h = hashlib.sha256()
    for k in ("flags", "traj_len", "costs", "total", "states"):
        h.update(np.ascontiguousarray(dev[k]).tobytes())
    return h.hexdigest()[:16], dev["argmin"]

for name in GOLDEN_CASES:
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(np.tile(g["sampling"], (5, 1)))
    print(name, *digest(device_plan(S, ref, prm, preds)))
