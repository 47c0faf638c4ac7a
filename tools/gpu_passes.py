"""Per-pass scoring-kernel time of a loop-closure batch (coarse / fine / super-fine)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
ctx.set_profiling(True)
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = synth.config4(npairs)
grids = []
for sc in pairs:
    dg = matcher.ScanMatchMap.from_spec(ctx, sc.grid)
    dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses)
    grids.append(dg)
m = matcher.BasedCorrelationScanMatch(ctx)
poses = [sc.seed_pose.copy() for sc in pairs]
import ctypes
for pi, name in enumerate(("coarse", "fine", "super")):
    # batch of single passes through rsm_match_batch is chain-only; emulate with per-pass params tripled? use chain w/ identical params
    sm = matcher.ScanMatchers(ctx, [pairs[0].passes[pi]] * 3)
    sm.ScanMatchBatch(grids, [sc.scan_pts for sc in pairs], poses, use_fine_scan_match=False)
    ctx.reset_stats()
    t0 = time.perf_counter()
    scores, newposes, covs, resp = sm.ScanMatchBatch(grids, [sc.scan_pts for sc in pairs], poses, use_fine_scan_match=False)
    wall = time.perf_counter() - t0
    st = ctx.stats()
    print("%s: wall %.3f ms score_k %.3f ms sel_k %.3f ms evals %.3g -> %.3g evals/s kernel; phases %s" % (
        name, wall * 1e3, st["score_kernel_ms"], st["select_kernel_ms"], st["evals"], st["evals"] / (st["score_kernel_ms"] * 1e-3),
        [round(v, 3) for v in st["phase_ms"][:6]]), flush=True)
    poses = list(newposes)
