"""Quick: config-2 / config-5-quarter score kernel time (profiling on) and parity of the pass result vs the golden-free oracle check (scores sha)."""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
for name, sc in (("cfg2", synth.config2()), ("cfg5/4", synth.config5(scale=0.25))):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    ctx.set_profiling(True)
    for _ in range(5):
        m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    ctx.reset_stats()
    n = 30
    for _ in range(n):
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        r = m.ScanMatch(grid, scan, sc.passes[0], pose, cov)
    st = ctx.stats()
    sd = m.scores(grid, sc.scan_pts, sc.passes[0], sc.seed_pose)
    print(name, "score_k %.1f us select_k %.1f us  evals/s kernel %.3g  resp %.6f pose %s scores sha %s" % (
        st["score_kernel_ms"] * 1e3 / n, st["select_kernel_ms"] * 1e3 / n, st["evals"] / (st["score_kernel_ms"] * 1e-3),
        r, pose, hashlib.sha256(sd.tobytes()).hexdigest()[:16]), flush=True)
    ctx.set_profiling(False)
    grid.close(); scan.close()
