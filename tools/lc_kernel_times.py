import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
pairs = synth.config4(512)
st = matcher.ScanStore(ctx)
chains, mids = [], []
for sc in pairs:
    chains.append([st.AddRangeData(p, q) for p, q in zip(sc.base_pts, sc.base_poses)])
    mids.append(st.AddRangeData(sc.scan_pts, sc.seed_pose))
centres = np.array([sc.grid_centre for sc in pairs]); seeds = np.array([sc.seed_pose for sc in pairs])
chains = matcher.pack_chains(chains); mids = np.array(mids, dtype=np.int32)
ctx.set_profiling(True)
for r in range(3):
    ctx.reset_stats()
    matcher.scan_match_interface_batch(ctx, st, pairs[0].grid, centres, chains, mids, seeds, pairs[0].passes)
    s = ctx.stats()
print("raster %.3f ms score %.3f select %.3f" % (s["raster_kernel_ms"], s["score_kernel_ms"], s["select_kernel_ms"]))
