"""Score-kernel time of the wide window (config 5) and config 2, L1 path vs staged path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
ctx.set_profiling(True)
m = matcher.BasedCorrelationScanMatch(ctx)
for name, sc in (("cfg5", synth.config5()), ("cfg5/4", synth.config5(scale=0.25)), ("cfg2", synth.config2())):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    for _ in range(2):
        m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    ctx.reset_stats()
    n = 3
    t0 = time.perf_counter()
    for _ in range(n):
        r = m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    wall = (time.perf_counter() - t0) / n
    st = ctx.stats()
    print("%s staged=%s: wall %.3f ms, score_k %.3f ms, sel_k %.3f ms, evals %.3g -> kernel %.3g evals/s, resp %.4f" % (
        name, os.environ.get("RSM_NO_STAGED") is None, wall * 1e3, st["score_kernel_ms"] / n, st["select_kernel_ms"] / n,
        st["evals"] / n, st["evals"] / n / (st["score_kernel_ms"] / n * 1e-3), r), flush=True)
    grid.close(); scan.close()
