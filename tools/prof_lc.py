"""One batched back-end step (rsm_scan_match_interface_batch) on N config-4 pairs, for ncu / timing runs.

    python tools/prof_lc.py [pairs] [reps]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = matcher.Context(0)
pairs = synth.config4(npairs)
st = matcher.ScanStore(ctx)
chains, mids = [], []
for sc in pairs:
    chains.append([st.AddRangeData(p, q) for p, q in zip(sc.base_pts, sc.base_poses)])
    mids.append(st.AddRangeData(sc.scan_pts, sc.seed_pose))
centres = np.array([sc.grid_centre for sc in pairs])
seeds = np.array([sc.seed_pose for sc in pairs])
chains = matcher.pack_chains(chains)
mids = np.array(mids, dtype=np.int32)
for r in range(reps + 1):
    ctx.reset_stats()
    t0 = time.perf_counter()
    res = matcher.scan_match_interface_batch(ctx, st, pairs[0].grid, centres, chains, mids, seeds, pairs[0].passes)
    dt = time.perf_counter() - t0
    s = ctx.stats()
    print("run %d: %.2f ms, %.0f matches/s, %.3g evals/s, launches %d, d2h %d B, h2d %d B, exact %d, phases %s" % (
        r, dt * 1e3, npairs / dt, s["evals"] / dt, s["kernel_launches"], s["d2h_bytes"], s["h2d_bytes"], s["exact_sort_passes"],
        [round(v, 2) for v in s["phase_ms"][:8]]), flush=True)
print("accepted", int((res[0] > 0.6).sum()))
