"""Golden vectors for front-end map maintenance: a scan-match map constructed like CreateAllMap does
(slam_processor.cpp:466-512: auto-resize on, just_update_occu on, never Reset) and then updated scan by scan
with the REFERENCE'S OWN OccuGridMap::UpdateMapByRange (oracle/_ref/libref.so) along a trajectory that leaves
the initial extent on every side.  Run in the build container only:

    python tests/golden/make_frontend.py

Per step the fixture holds whether the scan was stamped or the map extended instead (the reference drops that
scan), the map geometry afterwards and a checksum of every cell; plus the final map (sparse).  Scans are
re-synthesised by roborts_edu_slam_b200.synth (deterministic), guarded by a checksum.
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
RES, SIGMA, SIZE, BEAMS, FOV, RMAX, EXTEND = 0.05, 0.15, 480, 541, np.deg2rad(270.25), 10.0, 0.2
START = np.array([14.375, 28.625, 0.3])


def trajectory():
    """Right and up, then far back to the left and down: extends +x, +y, -x, -y."""
    out, p = [], START.copy()
    for k in range(14):
        p = p + np.array([0.9, 0.15, 0.07]); out.append(p.copy())
    for k in range(30):
        p = p + np.array([-0.9, -0.35, -0.05]); out.append(p.copy())
    return np.array(out)


def spec():
    half = 0.5 * SIZE * RES
    return synth.GridSpec(RES, SIGMA, SIZE, SIZE, -(START[0] - half), -(START[1] - half), 0.3, 0.88, True)


def scans():
    occ = synth.load_map("willow")
    return [synth.raycast(occ, p[0], p[1], p[2], BEAMS, FOV, RMAX) / RES for p in trajectory()]


def main():
    R = Ref()
    g = spec()
    m = R.frontend_map_create(g, EXTEND)
    poses, pts = trajectory(), scans()
    h = hashlib.sha256()
    for a in pts:
        h.update(np.ascontiguousarray(a).tobytes())
    stamped, geom, shas = [], [], []
    for p, s in zip(poses, pts):
        ok, ge = R.frontend_map_update(m, s, p, True)
        cells = R.read_map_sized(m, ge[0], ge[1])
        stamped.append(ok); geom.append(ge); shas.append(hashlib.sha256(cells.tobytes()).hexdigest())
    nz = np.flatnonzero(cells.ravel() != np.float32(0.5))
    dst = os.path.join(HERE, "frontend_willow.npz")
    np.savez_compressed(dst, inputs_sha=h.hexdigest(), stamped=np.array(stamped), geom=np.array(geom), shas=np.array(shas),
                        final_nz_index=nz.astype(np.int32), final_nz_value=cells.ravel()[nz])
    R.destroy_map(m)
    print("steps", len(poses), "extensions", int((~np.array(stamped)).sum()), "final", geom[-1], "->", dst, os.path.getsize(dst), "bytes")
    print("distinct geometries:", sorted(set(geom)))


if __name__ == "__main__":
    main()
