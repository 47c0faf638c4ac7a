"""Golden vectors for the Gauss-Newton matcher (BasedOptimizeScanMatch) and the chain that starts with it
(ScanMatchers::ScanMatch, use_optimize_scan_match on), generated from the REFERENCE'S OWN HEADERS
(oracle/_ref/libref.so).  Run in the build container only:

    python tests/golden/make_optimize.py

The one step of that code that is not the reference's own arithmetic here is H.ldlt().solve(b): real Eigen
is not in this image, so it resolves to oracle/standin/Eigen/Core's restatement of Eigen 3.3's LDLT.
Inputs are re-synthesised by roborts_edu_slam_b200.synth (deterministic; guarded by a checksum), so the
fixture holds only seeds, parameters and the reference's outputs.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Ref  # noqa: E402
from roborts_edu_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (iterate_max_times, cost_decrease_threshold, cost_min_threshold, max_update_distance, max_update_angle)
OPT_SETS = np.array([
    [10, 0.1, 0.5, 0.5, 0.5],      # config/real_robot_param.yaml:43-47
    [10, 1.0, 2.0, 0.5, 0.2],      # in-code defaults, scan_matchers.h:127-131
    [3, 1e-9, 0.0, 0.02, 0.01],    # tight clamps, stops on the iteration count
    [1, 0.1, 0.5, 0.5, 0.5],       # a single evaluation
])
SEED_DELTAS = np.array([
    [0.12, -0.07, 0.1], [0.02, 0.01, 0.02], [-0.2, 0.15, -0.12], [0.5, 0.4, -0.3], [3.0, 3.0, 1.0],
    [0.0, 0.0, 0.0], [-0.04, 0.3, 0.25], [9.0, -9.5, 2.0],      # the last one: most points leave the map
])
FAILED_COSTS = np.array([2.0, 20.0, 200.0])   # yaml, in-code default, and one the optimiser always passes
# chain cases: only seeds for which MapSizeCheck (scan_matchers.h:365-390) leaves the map alone -- beyond that the
# reference's matcher reads outside the grid, which the CUDA path reports as RSM_ERR_WINDOW instead
N_CHAIN_SEEDS = 4


def scenarios():
    return [synth.config1(), synth.config4(1)[0]]


def coarse_of(sc):
    """Coarse map of the same scene: twice the cell length, the reference's default coarse blur ratio."""
    g = sc.grid
    gc = synth.backend_grid(g.res * 2, g.sigma * 2, 10.0, sc.grid_centre)
    return gc, [p * 0.5 for p in sc.base_pts], sc.scan_pts * 0.5


def checksum(sc):
    h = hashlib.sha256()
    for a in [sc.scan_pts, sc.base_poses] + list(sc.base_pts):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    R = Ref()
    out = {"opt_sets": OPT_SETS, "seed_deltas": SEED_DELTAS, "failed_costs": FAILED_COSTS}
    for sc in scenarios():
        g = sc.grid
        gc, base_c, scan_c = coarse_of(sc)
        mf = R.create_map(g)
        R.build_map(mf, g, sc.base_pts, sc.base_poses)
        mc = R.create_map(gc)
        R.build_map(mc, gc, base_c, sc.base_poses)
        n_o, n_s = len(OPT_SETS), len(SEED_DELTAS)
        cost = np.zeros((2, n_o, n_s))
        pose = np.zeros((2, n_o, n_s, 3))
        for mi, (m, pts) in enumerate(((mf, sc.scan_pts), (mc, scan_c))):
            for oi, op in enumerate(OPT_SETS):
                for si, d in enumerate(SEED_DELTAS):
                    r = R.optimize(m, pts, op, sc.truth_pose + d)
                    cost[mi, oi, si], pose[mi, oi, si] = r["cost"], r["pose"]
        n_f = len(FAILED_COSTS)
        ch_score = np.zeros((n_f, 2, n_s))
        ch_pose = np.zeros((n_f, 2, n_s, 3))
        ch_cov = np.zeros((n_f, 2, n_s, 3, 3))
        ch_resp = np.zeros((n_f, 2, n_s, 4))
        for fi, fc in enumerate(FAILED_COSTS):
            for ui, use_fine in enumerate((True, False)):
                for si, d in enumerate(SEED_DELTAS[:N_CHAIN_SEEDS]):
                    r = R.match_chain_opt(mc, scan_c, mf, sc.scan_pts, synth.chain_yaml(), OPT_SETS[0], fc, sc.truth_pose + d,
                                          use_fine=use_fine)
                    ch_score[fi, ui, si], ch_pose[fi, ui, si], ch_cov[fi, ui, si] = r["score"], r["pose"], r["cov"]
                    ch_resp[fi, ui, si, 0], ch_resp[fi, ui, si, 1:] = r["optimize_cost"], r["responses"]
        R.destroy_map(mf)
        R.destroy_map(mc)
        tag = sc.name.split("_")[-1]
        out.update({tag + "_checksum": checksum(sc), tag + "_cost": cost, tag + "_pose": pose, tag + "_chain_score": ch_score,
                    tag + "_chain_pose": ch_pose, tag + "_chain_cov": ch_cov, tag + "_chain_resp": ch_resp})
        print(tag, "costs", np.round(cost[0, 0], 3), "chain scores", np.round(ch_score[:, 0, 0], 4))
    dst = os.path.join(HERE, "optimize_cases.npz")
    np.savez_compressed(dst, **out)
    print("->", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
