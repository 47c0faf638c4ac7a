"""Debug aid: one pass far from the truth (flat score landscape) -- scores, best, covariance vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
from oracle.oracle_py import Oracle
O = Oracle()
ctx = matcher.Context(0)
sc = synth.config1(); g = sc.grid
grid = O.build_grid(g, sc.base_pts, sc.base_poses)
dg = matcher.ScanMatchMap.from_spec(ctx, g); dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, True)
assert np.array_equal(dg.download(), grid)
m = matcher.BasedCorrelationScanMatch(ctx)
for d in ([3.0, 3.0, 1.0], [2.0, -1.0, 0.5], [0.12, -0.07, 0.1]):
    for p in synth.chain_yaml():
        seed = sc.truth_pose + np.array(d)
        so = O.scores(grid, g, sc.scan_pts, p, O.world_to_map(g, seed))
        sd = m.scores(dg, sc.scan_pts, p, seed)
        w = O.match(grid, g, sc.scan_pts, p, seed)
        pose, cov = seed.copy(), np.eye(3)
        r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
        D = m.last_detail
        srt = np.sort(so)[::-1]
        print(d, int(p[7]), "scores equal", np.array_equal(so, sd), "resp", r == w["response"], "n_avg", D.n_avg, w["n_avg"], "exact", D.exact_sort_used,
              "best", [D.best_pose_map[i] for i in range(3)] == list(w["best_map"][:3]), "cov", np.allclose(cov, w["cov"], rtol=1e-6, atol=0),
              "ties in top22:", int((np.diff(srt[:22]) == 0).sum()), "top20 vs 21st", srt[19], srt[20], flush=True)
        if not np.allclose(cov, w["cov"], rtol=1e-6, atol=0):
            print("   got", cov.ravel(), "\n   want", w["cov"].ravel())
            print("   best got", [D.best_pose_map[i] for i in range(3)], "want", w["best_map"])
