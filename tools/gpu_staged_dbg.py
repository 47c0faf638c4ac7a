"""Temporary: staged kernel pipeline timing (build with -DRSM_STAGED_DEBUG)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
lib = ctypes.CDLL(matcher.LIB_PATH)
out = (ctypes.c_ulonglong * 32)()
for name, sc in (("cfg5/4", synth.config5(scale=0.25)), ("cfg2", synth.config2()), ("cfg5", synth.config5())):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    lib.rsm_debug_staged(out, 1)
    m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    lib.rsm_debug_staged(out, 1)
    v = list(out)
    r = max(v[0], 1)
    print(name, "rounds", v[0], "per round: empty-wait %.0f plan %.0f issue %.0f beams %.1f tall %.2f | w0 full-wait %.0f compute %.0f | w15 full-wait %.0f compute %.0f" % (
        v[1]/r, v[2]/r, v[3]/r, v[5]/r, v[6]/r, v[8]/r, v[9]/r, v[10]/r, v[11]/r), flush=True)
    for o, what in ((12, "unsplit"), (24, "cluster")):
        c = max(v[o], 1)
        print("   %s CTAs %d: prologue %.0f loop %.0f combine %.0f epilogue %.0f tail %.0f" % (what, v[o], v[o+1]/c, v[o+2]/c, v[o+3]/c, v[o+4]/c, v[o+5]/c), flush=True)
    grid.close(); scan.close()
