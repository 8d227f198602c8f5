"""Summarise a FRX_TRACE dump: per-warp phase durations (us) of the first tile.  usage: trace_report.py trace.bin"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 16).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["entry", "ref staged", "rows ready", "memo ready", "pass done", "tile done", "all tiles done", "exit"]
print("warps", len(t), " kernel span (first entry -> last exit): %.1f us" % ((t[:, 7].max() - t0) / 1e3))
for k in range(8):
    v = (t[:, k] - t0) / 1e3
    v = v[t[:, k] > 0]
    if len(v): print(f"  {names[k]:>15}: min {v.min():7.1f}  median {np.median(v):7.1f}  p90 {np.percentile(v, 90):7.1f}  max {v.max():7.1f} us   ({len(v)} warps)")
d = lambda a, b: np.median((t[:, b] - t[:, a])[(t[:, a] > 0) & (t[:, b] > 0)]) / 1e3
print("median durations: prologue %.1f | rows %.1f | memo fill(s) %.1f | candidates (last pass) %.1f | tile total %.1f us" %
      (d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(2, 5)))

if (t[:, 8:15] > 0).any():
    m = t[(t[:, 8] > 0) & (t[:, 14] > 0)]
    lab = ["table lookup + coefficients + first time-table loads", "per-lane time-table loads + samples + extension",
           "votes + segment search", "interpolation", "sincos", "stores + header"]
    print("first memo fill, median us per phase:", ", ".join(f"{lab[k]} {np.median(m[:, 9 + k] - m[:, 8 + k]) / 1e3:.2f}" for k in range(6)),
          f"| total {np.median(m[:, 14] - m[:, 8]) / 1e3:.2f}")
