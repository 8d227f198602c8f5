"""Plan latency at the reference's default sampling density (config 1: a few hundred candidates) through the C ABI.
usage: [FRX_SEG=1|2|4] python scripts/latency_small.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, configure_handler
from frenetix_motion_planner_b200 import _capi

for name in ("tjunction_draw", "tjunction_nodraw"):
    g, ref, prm, preds = load_golden(name)
    S = np.ascontiguousarray(g["sampling"], dtype=np.float64)
    h = _capi.Handler(0)
    configure_handler(h, ref, prm, preds, None, sampling=S)
    for _ in range(20):
        r = h.plan(S)
    t0 = time.perf_counter()
    n = 200
    kms = []
    for _ in range(n):
        r = h.plan(S)
        kms.append(r.eval_kernel_ms)
        w = h.winner_states() if r.argmin >= 0 else None
    dt = (time.perf_counter() - t0) / n
    print(f"{name}: {S.shape[0]} candidates, {len(preds)} obstacles, SEG={os.environ.get('FRX_SEG', 'auto')}: "
          f"{dt * 1e6:.1f} us per plan() incl. winner read-back, eval kernel {np.median(kms) * 1e3:.1f} us, argmin {r.argmin}")
