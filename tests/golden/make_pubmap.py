"""Golden vectors for the publishing map kept on the device: PubMap = OccuGridMap<CountCell> constructed like
CreateAllMap does (slam_processor.cpp:477-483) and updated scan by scan with the REFERENCE'S OWN UpdateMapByRange
(ray-traced free space, auto-resize) with the CountCellFunctions knobs SlamProcessor::UpdateMap sets
(:538-551), along the trajectory of make_frontend.py.  Run in the build container only:

    python tests/golden/make_pubmap.py

Per step: updated or extended, geometry afterwards, checksum over (value, pass count, hit count) of every cell.
At the end: the occupancy the map check reads and the reference's MapCheckPenalize coefficient for 48 poses.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.oracle_py import Ref  # noqa: E402
import make_frontend as mf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# (update_free_factor, update_occu_factor): first frame = (map_min_passthrough, 2 * map_min_passthrough), then the
# in-code defaults (param_config.h:48-49) and the shipped yaml values (config/real_robot_param.yaml:19-20) alternate
FIRST, CODE, YAML = (4.0, 8.0), (0.3, 0.7), (0.0, 0.0)
OCCU_THRESHOLD, MIN_PASS = 0.2, 3.0


def factors(k):
    return FIRST if k == 0 else (CODE if k % 2 else YAML)


def main():
    R = Ref()
    g = mf.spec()
    m = R.pubmap_create_frontend(g, mf.EXTEND)
    poses, pts = mf.trajectory(), mf.scans()
    stamped, geom, shas = [], [], []
    for k, (p, s) in enumerate(zip(poses, pts)):
        f = factors(k)
        R.pubmap_set_factors(m, f[0], f[1], OCCU_THRESHOLD, MIN_PASS)
        ok, ge = R.pubmap_update_geom(m, s, p)
        val, cnt, hit, occ = R.pubmap_read_all(m, ge[0], ge[1])
        h = hashlib.sha256()
        for a in (val, cnt, hit):
            h.update(a.tobytes())
        stamped.append(ok); geom.append(ge); shas.append(h.hexdigest())
    rng = np.random.default_rng(4242)
    check_poses = poses[rng.integers(0, len(poses), 48)] + rng.uniform(-1, 1, (48, 3)) * np.array([0.8, 0.8, 0.3])
    coeff = np.array([R.pubmap_penalty(m, pts[-1], q, 100, 2.5, 0.015, False, True) for q in check_poses])
    dst = os.path.join(HERE, "pubmap_willow.npz")
    np.savez_compressed(dst, stamped=np.array(stamped), geom=np.array(geom), shas=np.array(shas), occ_packed=np.packbits(occ.ravel()),
                        check_poses=check_poses, coeff=coeff, knobs=np.array([OCCU_THRESHOLD, MIN_PASS]))
    R.pubmap_destroy(m)
    print("steps", len(poses), "extensions", int((~np.array(stamped)).sum()), "final", geom[-1], "occupied", int(occ.sum()),
          "distinct coefficients", len(np.unique(coeff)), "->", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
