"""Temporary: select kernel phase timing (build with -DRSM_SELECT_DEBUG)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
lib = ctypes.CDLL(matcher.LIB_PATH)
out = (ctypes.c_ulonglong * 16)()
for name, sc in (("cfg1", synth.config1()), ("cfg2", synth.config2()), ("cfg5/4", synth.config5(scale=0.25))):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    for rep in range(3):
        m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
        lib.rsm_debug_select(out)
        v = list(out)
        t0 = v[0]
        print(name, "us since first CTA start: last start %.1f | stage1 done first %.1f last %.1f | merge begin %.1f merge end %.1f gather end %.1f" % tuple((x - t0) / 1e3 for x in (v[1], v[2], v[3], v[4], v[5], v[6])), "| max over top_k calls: pass1 %.1f pass2 %.1f rank %.1f us, count %d" % (v[8] / 1e3, v[9] / 1e3, v[10] / 1e3, v[11]), flush=True)
    grid.close(); scan.close()
