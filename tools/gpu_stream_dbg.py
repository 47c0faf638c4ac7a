"""Stream-plan staged kernel: where a CTA's cycles go (build the library with -DRSM_STAGED_DEBUG: tools/build_debug.sh)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
lib = ctypes.CDLL(matcher.LIB_PATH)
out = (ctypes.c_ulonglong * 32)()
for name, sc in (("cfg2", synth.config2()), ("cfg5/4", synth.config5(scale=0.25)), ("cfg5", synth.config5())):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    lib.rsm_debug_staged(out, 1)
    m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    lib.rsm_debug_staged(out, 1)
    v = list(out)
    it, fin, r = max(v[0], 1), max(v[24], 1), max(v[7], 1)
    print(name, "item visits %d (finished %d): setup %.0f loop %.0f exact+exchange %.0f | epilogue %.0f per finished item" % (
        v[0], v[24], v[1] / it, v[2] / it, v[3] / it, v[4] / fin), flush=True)
    print("   producer rounds %d: plan %.0f empty-wait %.0f publish+issue %.0f" % (v[7], v[5] / r, v[6] / r, v[13] / r))
    for o, w in ((8, 0), (16, 5)):
        n = max(v[o + 3], 1)
        print("   warp %d rounds %d: full-wait %.0f gather %.0f (%.1f beams, %.1f per beam) flush+arrive %.0f" % (
            w, v[o + 3], v[o] / n, v[o + 1] / n, v[o + 4] / n, v[o + 1] / max(v[o + 4], 1), v[o + 2] / n), flush=True)
    grid.close(); scan.close()
