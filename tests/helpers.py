"""Shared helpers of the parity tests."""
import glob
import hashlib
import os

import numpy as np

from roborts_edu_slam_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("maps_", "mapcheck_", "optimize_", "frontend_", "pubmap_")))


def mapcheck_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "mapcheck_*.npz")))


def load_mapcheck(name):
    """-> (GridSpec of the publishing map, occupancy [size_y, size_x] uint8, npz with scan, poses, parameter sets and
    the reference's coefficients)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    gs = z["grid_spec"]
    g = synth.GridSpec(float(gs[0]), 0.0, int(gs[1]), int(gs[2]), float(gs[3]), float(gs[4]), 0.5, 0.88, False)
    occ = np.unpackbits(z["occ_packed"])[: g.size_x * g.size_y].reshape(g.size_y, g.size_x)
    return g, occ, z


def load_golden(name):
    """-> (synth.Scenario rebuilt from the stored inputs, dict of the reference's outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    gs = z["grid_spec"]
    g = synth.GridSpec(float(gs[0]), float(gs[1]), int(gs[2]), int(gs[3]), float(gs[4]), float(gs[5]),
                       float(gs[6]), float(gs[7]), bool(gs[8]))
    base_n = z["base_n"]
    offs = np.concatenate([[0], np.cumsum(base_n)])
    base_pts = [z["base_pts"][offs[i]:offs[i + 1]].copy() for i in range(len(base_n))]
    sc = synth.Scenario(str(z["name"]), g, base_pts, z["base_poses"].copy(), z["scan_pts"].copy(),
                        z["seed_pose"].copy(), z["seed_pose"].copy(), [p.copy() for p in z["passes"]])
    return sc, z


def golden_grid(z, g):
    grid = np.full(g.size_x * g.size_y, np.float32(g.default_prob), dtype=np.float32)
    grid[z["grid_nz_index"]] = z["grid_nz_value"]
    return grid.reshape(g.size_y, g.size_x)


def cov_close(a, b, rtol=1e-6):
    """Covariance tolerance of BASELINE.json's north_star: 1e-6 relative, element-wise."""
    return np.allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=0.0)


def random_scenario(rng, n_points=200, size=160, res=0.05, n_base=3, spread=60.0):
    """A small random match problem on a back-end style grid: points on a noisy ring."""
    centre = rng.uniform(5.0, 20.0, size=2)
    cell_len = 1 / (1.0 / res)
    g = synth.GridSpec(res, 0.15, size, size, -(centre[0] - 0.5 * size * cell_len), -(centre[1] - 0.5 * size * cell_len))
    ang = np.sort(rng.uniform(-np.pi, np.pi, n_points))
    rad = spread * (0.55 + 0.4 * rng.random(n_points)) * 0.5
    ring = np.stack([np.cos(ang) * rad, np.sin(ang) * rad], axis=1)
    truth = np.array([centre[0], centre[1], rng.uniform(-np.pi, np.pi)])

    def seen_from(pose):
        # express the ring (world cells around the centre) in the sensor frame of `pose`
        d = ring - (pose[:2] - centre) / cell_len
        c, s = np.cos(-pose[2]), np.sin(-pose[2])
        return np.stack([c * d[:, 0] - s * d[:, 1], s * d[:, 0] + c * d[:, 1]], axis=1)

    base_pts, base_poses = [], []
    for k in range(n_base):
        bp = truth + np.array([0.05 * (k - 1), 0.03 * (k - 1), 0.05 * k])
        base_pts.append(seen_from(bp))
        base_poses.append(bp)
    seed = truth + np.array([0.08, -0.05, 0.06])
    return synth.Scenario("random", g, base_pts, np.array(base_poses), seen_from(truth), seed, truth, [],
                          truth[:2].copy())


def optimize_cases():
    """-> (npz of the reference's Gauss-Newton outputs, [(tag, fine scenario, coarse GridSpec, coarse base scans,
    coarse scan)]) -- inputs re-synthesised, checked against the fixture's checksum."""
    import hashlib
    z = np.load(os.path.join(GOLDEN_DIR, "optimize_cases.npz"), allow_pickle=False)
    out = []
    for sc in (synth.config1(), synth.config4(1)[0]):
        tag = sc.name.split("_")[-1]
        h = hashlib.sha256()
        for a in [sc.scan_pts, sc.base_poses] + list(sc.base_pts):
            h.update(np.ascontiguousarray(a).tobytes())
        assert h.hexdigest() == str(z[tag + "_checksum"]), "synthetic inputs changed since the fixture was made"
        g = sc.grid
        gc = synth.backend_grid(g.res * 2, g.sigma * 2, 10.0, sc.grid_centre)
        out.append((tag, sc, gc, [p * 0.5 for p in sc.base_pts], sc.scan_pts * 0.5))
    return z, out
