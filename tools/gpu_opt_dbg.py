"""Debug aid: chain-with-optimiser cases against the fixture, composed calls vs rsm_match_chain_opt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import optimize_cases, cov_close
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
z, cases = optimize_cases()
opt = matcher.BasedOptimizeScanMatch(ctx)
m = matcher.BasedCorrelationScanMatch(ctx)
for tag, sc, gc, base_c, scan_c in cases:
    gf = sc.grid
    dev_f = matcher.ScanMatchMap.from_spec(ctx, gf); dev_f.InitMapWithRangeVec(sc.base_pts, sc.base_poses, gf.default_prob, gf.sigma, gf.occu_offset, True)
    dev_c = matcher.ScanMatchMap.from_spec(ctx, gc); dev_c.InitMapWithRangeVec(base_c, sc.base_poses, gc.default_prob, gc.sigma, gc.occu_offset, True)
    seeds = sc.truth_pose + z["seed_deltas"]
    sm = matcher.ScanMatchers(ctx, synth.chain_yaml())
    P = synth.chain_yaml()
    for fi, fc in enumerate(z["failed_costs"]):
        for ui, use_fine in enumerate((True, False)):
            for si in range(5):
                pose, cov = seeds[si].copy(), np.eye(3)
                s = sm.ScanMatchWithOptimize(scan_c, sc.scan_pts, dev_c, dev_f, pose, cov, z["opt_sets"][0], fc, use_fine)
                ok = (s == z[tag + "_chain_score"][fi, ui, si], np.array_equal(pose, z[tag + "_chain_pose"][fi, ui, si]),
                      cov_close(cov, z[tag + "_chain_cov"][fi, ui, si]), np.array_equal(sm.last_responses, z[tag + "_chain_resp"][fi, ui, si, 1:]))
                # composed
                p2 = seeds[si].copy(); c2 = np.eye(3)
                cost = opt.ScanMatch(dev_c, scan_c, z["opt_sets"][0], p2)
                rr = [0, 0, 0]
                if (not use_fine) or cost > fc:
                    p2 = seeds[si].copy()
                    rr[0] = m.ScanMatch(dev_f, sc.scan_pts, P[0], p2, c2)
                if use_fine:
                    rr[1] = m.ScanMatch(dev_f, sc.scan_pts, P[1], p2, c2)
                    rr[2] = m.ScanMatch(dev_f, sc.scan_pts, P[2], p2, c2)
                ok2 = (np.array_equal(p2, z[tag + "_chain_pose"][fi, ui, si]), cov_close(c2, z[tag + "_chain_cov"][fi, ui, si]))
                if not all(ok) or not all(ok2):
                    print(tag, fi, ui, si, "chain_opt", ok, "composed", ok2, "resp", sm.last_responses, z[tag + "_chain_resp"][fi, ui, si], flush=True)
                    print("  cov got", cov.ravel(), "\n  cov composed", c2.ravel(), "\n  want", z[tag + "_chain_cov"][fi, ui, si].ravel())
print("done")
