"""Golden vectors for the map check (MapFeedbackResponsePenalty / MapCheckPenalize), generated from
the REFERENCE'S OWN CODE (oracle/_ref/libref.so).  Run in the build container only:

    python tests/golden/make_mapcheck.py

tests/golden/mapcheck_<map>.npz holds a publishing map built by the reference's own
OccuGridMap<CountCell>::UpdateMapByRange from the scenario's base scans (stored as the packed
occupancy the check reads), the scan, a set of candidate poses and parameter sets, and the
coefficient the reference returned for every (parameter set, pose).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Ref  # noqa: E402
from roborts_edu_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (check_point_num, bound_tolerance, penalty_gain, use_logistic, origin_x, origin_y)
PARAM_SETS = np.array([
    [100, 2.5, 0.015, 0, 0.0, 0.0],      # config/real_robot_param.yaml:77-80, front end
    [100, 2.5, 0.015, 1, 0.0, 0.0],      # loop closure (logistic)
    [20, 0.0, 0.1, 0, 0.0, 0.0],
    [7, 1.0, 0.3, 1, 1.5, -2.25],        # sensor origin off the pose
    [100, -1.0, 0.015, 0, 0.0, 0.0],     # knobs out of range: check switched off
    [100000, 2.5, 0.015, 0, 0.0, 0.0],   # every point
])


def poses_for(sc, rng, n):
    out = [sc.truth_pose.copy(), sc.seed_pose.copy()]
    while len(out) < n - 2:
        scale = rng.choice([0.05, 0.3, 1.0, 3.0])
        out.append(sc.truth_pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)]) * scale)
    g = sc.grid
    out.append(np.array([-g.off_x - 1.0, 1.0, 0.2]))                 # outside the map -> 0.0
    # near the map corner: most rays leave the map (skipped); the sensor cell stays inside for every
    # origin in PARAM_SETS (the reference reads out of bounds when the sensor cell itself is outside)
    out.append(np.array([-g.off_x + 3.5 * g.res, -g.off_y + 3.5 * g.res, 0.0]))
    return np.array(out)


def main():
    R = Ref()
    rng = np.random.default_rng(20261018)
    for sc in (synth.config1(), synth.config4(1)[0]):
        g = sc.grid
        m = R.pubmap_create(g)
        for pts, pose in zip(sc.base_pts, sc.base_poses):
            assert R.pubmap_update(m, pts, pose) == 0
        value, count, occ = R.pubmap_read(m, g)
        poses = poses_for(sc, rng, 48)
        coeff = np.zeros((len(PARAM_SETS), len(poses)))
        for k, ps in enumerate(PARAM_SETS):
            for i, pose in enumerate(poses):
                coeff[k, i] = R.pubmap_penalty(m, sc.scan_pts, pose, int(ps[0]), ps[1], ps[2], False, bool(ps[3]), ps[4:6])
        R.pubmap_destroy(m)
        name = "mapcheck_" + sc.name.split("_")[-1]
        dst = os.path.join(HERE, name + ".npz")
        np.savez_compressed(dst, name=name, grid_spec=np.array([g.res, g.size_x, g.size_y, g.off_x, g.off_y]),
                            occ_packed=np.packbits(occ.ravel()), scan_pts=sc.scan_pts, poses=poses,
                            param_sets=PARAM_SETS, coeff=coeff)
        print(name, "->", dst, os.path.getsize(dst), "bytes;", int(occ.sum()), "occupied cells;",
              len(np.unique(coeff)), "distinct coefficients, e.g.", np.unique(coeff)[:6])


if __name__ == "__main__":
    main()
