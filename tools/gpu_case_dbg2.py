import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
from oracle.oracle_py import Oracle
O = Oracle()
ctx = matcher.Context(0)
sc = synth.config1(); g = sc.grid
grid = O.build_grid(g, sc.base_pts, sc.base_poses)
dg = matcher.ScanMatchMap.from_spec(ctx, g); dg.upload(grid)
m = matcher.BasedCorrelationScanMatch(ctx)
p = synth.chain_yaml()[0]
seed = sc.truth_pose + np.array([3.0, 3.0, 1.0])
cm = O.world_to_map(g, seed)
so = O.scores(grid, g, sc.scan_pts, p, cm)
sd = m.scores(dg, sc.scan_pts, p, seed)
geo = O.geometry(g, p, len(sc.scan_pts), cm)
print(geo, repr(cm))
bad = np.nonzero(so != sd)[0]
print("mismatches", len(bad), "of", len(so))
n = geo["n_xy"]
for k in bad[:40]:
    ia, ix, iy = k // (n * n), (k // n) % n, k % n
    print(k, ia, ix, iy, repr(so[k]), repr(sd[k]), (sd[k] - so[k]) * len(sc.scan_pts))
np.save("gpurun_out/bad_idx.npy", bad)
np.save("gpurun_out/bad_sd.npy", sd)
