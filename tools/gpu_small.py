"""Where does the time of a SMALL match go (BASELINE configs[0] and configs[2])?  Phase timers + kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
names = ["prep+launch", "wait score+select", "host1", "gather rt", "host2", "exact", "raster", "-"]
for name, sc in (("cfg1", synth.config1()), ("cfg3", synth.config3()), ("cfg3 shipped", synth.config3(True))):
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    for pi, p in enumerate(sc.passes):
        for prof in (False, True):
            ctx.set_profiling(prof)
            for _ in range(5):
                m.ScanMatch(grid, sc.scan_pts, p, sc.seed_pose.copy(), np.eye(3))
            ctx.reset_stats()
            n = 50
            t0 = time.perf_counter()
            for _ in range(n):
                m.ScanMatch(grid, sc.scan_pts, p, sc.seed_pose.copy(), np.eye(3))
            wall = (time.perf_counter() - t0) / n
            st = ctx.stats()
            print(name, "pass", pi, "prof", prof, "wall %.1f us" % (wall * 1e6), {k: round(v * 1e3 / n, 1) for k, v in zip(names, st["phase_ms"]) if v},
                  "score_k %.1f sel_k %.1f us, launches %.1f exact %d n_xy %d n_ang %d" % (st["score_kernel_ms"] * 1e3 / n, st["select_kernel_ms"] * 1e3 / n,
                  st["kernel_launches"] / n, st["exact_sort_passes"], m.last_detail.n_xy, m.last_detail.n_ang), flush=True)
    grid.close()
