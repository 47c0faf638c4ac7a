"""Generate golden input/output vectors from the REFERENCE'S OWN CODE (oracle/_ref/libref.so,
i.e. /root/reference/src headers compiled in place).  Run in the build container only:

    python tests/golden/make_golden.py [name-prefix ...]

Each tests/golden/<case>.npz holds the complete inputs of one match (grid geometry, base scans,
scan, seed pose, pass parameters) and what the reference returned for them: the lookup grid
(as a sha256 + the sparse list of non-default cells), every pass response, the final pose, the
3x3 covariance, and for the first pass the sha256 of the sorted candidate scores plus its head.
The GPU box has no /root/reference; tests there check the oracle restatement and the CUDA path
against these files.
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


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tie_heavy_case():
    """Binary grid, penalty off, coarse steps: thousands of exactly tied candidates, so the result
    depends on libstdc++'s unstable sort order (SURVEY.md hard part H1)."""
    sc = synth.config1()
    sc.name = "ties_icra"
    g = sc.grid
    g.sigma = 0.03  # 1x1 kernel at 0.05 m -> cells are 0.3 or 1.0 only
    sc.passes = [synth.pass_param(0.6, 0.05, 0.349, 0.0349, 0.2, 40, False, synth.COARSE),
                 synth.pass_param(0.2, 0.05, 0.175, 0.0349, 0.2, 40, False, synth.FINE),
                 synth.pass_param(0.1, 0.05, 0.0349, 0.00349, 0.2, 40, False, synth.SUPER)]
    return sc


def non_blur_cases():
    """SET_CELL_OCCUPIED construction (map/occu_grid_map.h:317-321, 499-516): use_blur off on the config-1 map, and
    blur parameters the reference's GaussianBlur rejects (sigma >= 10 * resolution) on a config-4 pair, which makes
    UpdateMapByRange drop use_blur by itself (:265-268).  Cells end up 0.3, 0.8 (hit by one scan) or 1.0."""
    a = synth.config1()
    a.name = "nonblur_icra"
    a.grid.use_blur = False
    a.passes = synth.chain_yaml()
    b = synth.config4(1, seed=99)[0]
    b.name = "nonblur_badsigma_pair"
    b.grid.sigma = 5.0
    return [a, b]


def cases():
    c1 = synth.config1()
    c3 = synth.config3(shipped_points=True)
    c3.name = "cfg3_willow_shipped"
    # keep the fixture small: the 2400^2 grid is rebuilt from the stored scans
    c4 = synth.config4(1)[0]
    return [c1, c3, c4, tie_heavy_case()] + non_blur_cases()


def main():
    R = Ref()
    only = [a for a in sys.argv[1:] if not a.startswith("-")]      # optional: name prefixes to (re)generate
    for sc in cases():
        if only and not sc.name.startswith(tuple(only)):
            continue
        g = sc.grid
        m = R.create_map(g)
        R.build_map(m, g, sc.base_pts, sc.base_poses)
        grid = R.read_map(m, g)
        nz = np.flatnonzero(grid.ravel() != np.float32(g.default_prob)).astype(np.int32)
        out = dict(
            name=sc.name,
            grid_spec=np.array([g.res, g.sigma, g.size_x, g.size_y, g.off_x, g.off_y, g.default_prob, g.occu_offset, float(g.use_blur)]),
            base_n=np.array([len(p) for p in sc.base_pts], dtype=np.int32),
            base_pts=np.concatenate(sc.base_pts, axis=0),
            base_poses=sc.base_poses,
            scan_pts=sc.scan_pts,
            seed_pose=sc.seed_pose,
            passes=np.array(sc.passes),
            grid_sha=sha(grid), grid_nz_index=nz, grid_nz_value=grid.ravel()[nz],
        )
        centre = R.world_to_map(m, sc.seed_pose)
        cand = R.candidates(m, sc.scan_pts, sc.passes[0], centre)
        out["centre_map"] = centre
        out["pass0_sorted_scores_sha"] = sha(cand["score"])
        out["pass0_sorted_head"] = cand["score"][:64].copy()
        out["pass0_best"] = cand["best"]
        r1 = R.match(m, sc.scan_pts, sc.passes[0], sc.seed_pose)
        out["pass0_response"] = r1["response"]
        out["pass0_pose"] = r1["pose"]
        out["pass0_cov"] = r1["cov"]
        if len(sc.passes) == 3:
            rc = R.match_chain(m, sc.scan_pts, sc.passes, sc.seed_pose)
            out["chain_score"] = rc["score"]
            out["chain_pose"] = rc["pose"]
            out["chain_cov"] = rc["cov"]
            out["chain_responses"] = rc["responses"]
        R.destroy_map(m)
        dst = os.path.join(HERE, sc.name + ".npz")
        np.savez_compressed(dst, **out)
        print(sc.name, "->", dst, os.path.getsize(dst), "bytes; pass0 response", r1["response"],
              "chain" if len(sc.passes) == 3 else "", out.get("chain_score", ""))


if __name__ == "__main__":
    main()
