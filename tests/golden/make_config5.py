"""Golden vector for BASELINE configs[4] at FULL size (+-8 m / 360 deg at 0.05 m on the whole willow map:
321 x 321 x 721 = 74 293 761 candidates x ~940 beams = 7.0e10 evaluations).  Build container only:

    python tests/golden/make_config5.py [--skip-ref]

Two sources, which must agree before anything is written:
  * the REFERENCE'S OWN CODE (oracle/_ref/libref.so): BasedCorrelationScanMatch::ScanMatch on the full window
    (single thread, ~5 min, ~6 GB for its 40-byte candidate array and the by-value copy) -> response, pose,
    covariance; and its sorted candidate list -> sha256 of the sorted scores, the head of the list;
  * the oracle restatement, scored per angle slice on all host threads -> the same quantities plus the
    checksums a GPU test can localise a difference with (sha256 of the unsorted score array, one 64-bit sum of
    the score bit patterns per search angle).
The fixture stores no inputs: tests re-synthesise config 5 (roborts_edu_slam_b200/synth.py) and check the
input checksum stored here first.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Oracle, Ref, ref_available  # noqa: E402
from roborts_edu_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def input_checksum(sc):
    h = hashlib.sha256()
    for a in [sc.scan_pts, sc.base_poses, sc.seed_pose, sc.passes[0]] + list(sc.base_pts):
        h.update(np.ascontiguousarray(a).tobytes())
    g = sc.grid
    h.update(np.array([g.res, g.sigma, g.size_x, g.size_y, g.off_x, g.off_y, g.default_prob, g.occu_offset]).tobytes())
    return h.hexdigest()


def angle_sums(scores, n_ang):
    """One wrap-around uint64 sum of the scores' bit patterns per search angle (order-free within an angle)."""
    return scores.view(np.uint64).reshape(n_ang, -1).sum(axis=1, dtype=np.uint64)


def main():
    sc = synth.config5()
    g, p = sc.grid, sc.passes[0]
    O = Oracle()
    t0 = time.time()
    grid = O.build_grid(g, sc.base_pts, sc.base_poses)
    centre = O.world_to_map(g, sc.seed_pose)
    geo = O.geometry(g, p, len(sc.scan_pts), centre)
    scores = O.scores_threaded(grid, g, sc.scan_pts, p, centre)
    print("oracle: %d candidates x %d beams scored in %.1f s" % (scores.size, geo["visited"], time.time() - t0), flush=True)
    fin = O.finish_scores(scores, g, len(sc.scan_pts), p, sc.seed_pose)
    order = np.sort(scores)[::-1]
    out = dict(
        input_checksum=input_checksum(sc), grid_sha=sha(grid),
        n_ang=geo["n_ang"], n_xy=geo["n_xy"], visited=geo["visited"], centre_map=centre,
        scores_sha=sha(scores), sorted_scores_sha=sha(order), sorted_head=order[:64].copy(),
        angle_sums=angle_sums(scores, geo["n_ang"]),
        response=fin["response"], pose=fin["pose"], cov=fin["cov"], n_avg=fin["n_avg"], best_map=fin["best_map"],
        source="oracle",
    )
    print("oracle: response %.17g pose %s n_avg %d" % (fin["response"], fin["pose"], fin["n_avg"]), flush=True)
    del order
    if "--skip-ref" not in sys.argv and ref_available():
        R = Ref()
        m = R.create_map(g)
        R.build_map(m, g, sc.base_pts, sc.base_poses)
        assert np.array_equal(R.read_map(m, g), grid), "lookup grid: oracle != reference"
        t0 = time.time()
        r = R.match(m, sc.scan_pts, p, sc.seed_pose)
        print("reference: ScanMatch on the full window in %.1f s: response %.17g pose %s" % (time.time() - t0, r["response"], r["pose"]), flush=True)
        assert r["response"] == fin["response"] and np.array_equal(r["pose"], fin["pose"]) and np.array_equal(r["cov"], fin["cov"]), \
            "full-size config 5: oracle != reference"
        t0 = time.time()
        cand = R.candidates(m, sc.scan_pts, p, R.world_to_map(m, sc.seed_pose))
        print("reference: sorted candidate list in %.1f s" % (time.time() - t0), flush=True)
        assert sha(cand["score"]) == out["sorted_scores_sha"], "sorted candidate scores: oracle != reference"
        assert np.array_equal(cand["best"], fin["best_map"])
        R.destroy_map(m)
        out["source"] = "reference (oracle/_ref/libref.so), bit-equal to the oracle restatement"
    dst = os.path.join(HERE, "config5_full.npz")
    np.savez_compressed(dst, **out)
    print("->", dst, os.path.getsize(dst), "bytes;", out["source"])


if __name__ == "__main__":
    main()
