"""Where does a step's wall time go?  Prints rsm_stats phase timers for config 2 and a loop-closure batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher

ctx = matcher.Context(0)
print("host cores:", os.cpu_count(), flush=True)
m = matcher.BasedCorrelationScanMatch(ctx)
names = ["prep+launch", "wait score+select", "host1", "gather rt", "host2", "exact", "raster", "-"]

def show(tag, st, n):
    print(tag, "per call [us]:", {k: round(v * 1e3 / n, 1) for k, v in zip(names, st["phase_ms"])},
          "score_k %.1f sel_k %.1f" % (st["score_kernel_ms"] * 1e3 / n, st["select_kernel_ms"] * 1e3 / n),
          "launches/call %.1f exact %d" % (st["kernel_launches"] / n, st["exact_sort_passes"]), flush=True)

sc = synth.config2()
g = sc.grid
grid = matcher.ScanMatchMap.from_spec(ctx, g)
grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
scan = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
for prof in ((True,) if "--quick" in sys.argv else (False, True)):
    ctx.set_profiling(prof)
    for _ in range(3):
        m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    ctx.reset_stats()
    n = 3 if "--quick" in sys.argv else 20
    t0 = time.perf_counter()
    for _ in range(n):
        m.ScanMatch(grid, scan, sc.passes[0], sc.seed_pose.copy(), np.eye(3))
    wall = (time.perf_counter() - t0) / n
    show("cfg2 profiling=%s wall %.1f us" % (prof, wall * 1e6), ctx.stats(), n)

QUICK = "--quick" in sys.argv
for npairs in ((64,) if QUICK else (64, 256)):
    pairs = synth.config4(npairs)
    packed = matcher.pack_loop_closure(pairs)
    import torch
    keep = []
    for key in ("base_pts", "pts", "base_poses", "centres", "poses"):
        t = torch.from_numpy(packed[key]).pin_memory(); keep.append(t)
        packed[key] = t.numpy()
    for prof in (False, True):
        ctx.set_profiling(prof)
        matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
        ctx.reset_stats()
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
        wall = (time.perf_counter() - t0) / n
        st = ctx.stats()
        show("loop_closure %d pairs profiling=%s wall %.2f ms (%.0f matches/s, %.3g evals/s)" % (npairs, prof, wall * 1e3, npairs / wall, st["evals"] / n / wall), st, n)
    # the same pairs from a device-resident scan store (ids only travel)
    st_ = matcher.ScanStore(ctx)
    chains, mids = [], []
    for sc_ in pairs:
        chains.append([st_.AddRangeData(p, q) for p, q in zip(sc_.base_pts, sc_.base_poses)])
        mids.append(st_.AddRangeData(sc_.scan_pts, sc_.seed_pose))
    centres = [sc_.grid_centre for sc_ in pairs]
    seeds = [sc_.seed_pose for sc_ in pairs]
    for prof in (False, True):
        ctx.set_profiling(prof)
        matcher.scan_match_interface_batch(ctx, st_, pairs[0].grid, centres, chains, mids, seeds, pairs[0].passes)
        ctx.reset_stats()
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            matcher.scan_match_interface_batch(ctx, st_, pairs[0].grid, centres, chains, mids, seeds, pairs[0].passes)
        wall = (time.perf_counter() - t0) / n
        st = ctx.stats()
        show("store loop_closure %d pairs profiling=%s wall %.2f ms (%.0f matches/s, %.3g evals/s) raster_k %.1f us" % (npairs, prof, wall * 1e3, npairs / wall, st["evals"] / n / wall, st["raster_kernel_ms"] * 1e3 / n), st, n)
    st_.close()
