"""Where a config-2 step goes: host wall clock per call (no L2 flush), the library's phase timers, kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
m = matcher.BasedCorrelationScanMatch(ctx)
sc = synth.config2()
g = sc.grid
grid = matcher.ScanMatchMap.from_spec(ctx, g)
grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
def run(n, flush=False, timer=False):
    ctx.reset_stats()
    t_ev = 0.0
    t0 = time.perf_counter()
    for _ in range(n):
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        if flush: ctx.flush_l2()
        if timer: ctx.timer_start()
        m.ScanMatch(grid, scan, sc.passes[0], pose, cov)
        if timer: t_ev += ctx.timer_stop()
    w = (time.perf_counter() - t0) / n * 1e3
    st = ctx.stats()
    return w, t_ev / n, st
for label, prof in (("graph", False), ("profiling (no graph)", True)):
    ctx.set_profiling(prof)
    run(10)
    w, _, st = run(200)
    ph = [round(v / 200 * 1e3, 1) for v in st["phase_ms"][:4]]
    print("%s: wall %.1f us per call; phases us [prepare+launch, wait, finalise, ..] %s; score_k %.1f select_k %.1f" % (
        label, w * 1e3, ph, st["score_kernel_ms"] / 200 * 1e3, st["select_kernel_ms"] / 200 * 1e3), flush=True)
    w, ev, st = run(100, flush=True, timer=True)
    print("   with flush + event timer: events %.1f us per step, wall %.1f us" % (ev * 1e3, w * 1e3), flush=True)
    w, ev, st = run(100, flush=False, timer=True)
    print("   event timer, no flush: events %.1f us per step, wall %.1f us" % (ev * 1e3, w * 1e3), flush=True)
