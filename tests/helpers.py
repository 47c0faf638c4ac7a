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
                  if not os.path.basename(p).startswith(("maps_", "mapcheck_", "optimize_", "frontend_", "pubmap_", "config5_")))


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


def load_config5_golden():
    """-> (config 5 re-synthesised, the fixture of tests/golden/make_config5.py); the fixture holds no inputs, only their
    checksum, which is verified here."""
    z = np.load(os.path.join(GOLDEN_DIR, "config5_full.npz"), allow_pickle=False)
    sc = synth.config5()
    g = sc.grid
    h = hashlib.sha256()
    for a in [sc.scan_pts, sc.base_poses, sc.seed_pose, sc.passes[0]] + list(sc.base_pts):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.array([g.res, g.sigma, g.size_x, g.size_y, g.off_x, g.off_y, g.default_prob, g.occu_offset]).tobytes())
    assert h.hexdigest() == str(z["input_checksum"]), "synthetic config-5 inputs changed since the fixture was made"
    return sc, z


_SHIM = None


def host_shim():
    """The product's host-side arithmetic (csrc/rsm_host.h) compiled with g++ through tests/host_shim.cpp: needs neither
    nvcc nor librsm.so, so CPU tests can use the host logic on a fresh checkout."""
    global _SHIM
    if _SHIM is None:
        import ctypes
        import subprocess
        import tempfile
        here = os.path.dirname(os.path.abspath(__file__))
        root = os.path.dirname(here)
        out = os.path.join(tempfile.mkdtemp(prefix="rsm_host_shim_"), "libhostshim.so")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-I" + os.path.join(root, "roborts_edu_slam_b200", "csrc"), "-o", out, os.path.join(here, "host_shim.cpp")])
        L = ctypes.CDLL(out)
        c_d, c_i, c_p = ctypes.c_double, ctypes.c_int, ctypes.c_void_p
        L.hs_bounds_create.restype = c_p
        L.hs_bounds_create.argtypes = [c_i, c_i, c_d, c_d, c_d, c_d]
        L.hs_bounds_destroy.argtypes = [c_p]
        L.hs_bounds_update_scan.argtypes = [c_p, c_p, c_i, c_p, c_i, c_i, c_p]
        L.hs_bounds_size_check.argtypes = [c_p, c_p, c_d, c_d, c_p]
        _SHIM = L
    return _SHIM


class ShimBounds:
    """MapBounds of csrc/rsm_host.h through the g++ shim; the interface of matcher.MapBounds (which binds the same code
    inside librsm.so)."""

    def __init__(self, size_x, size_y, resolution, offset_x, offset_y, extend_factor=1.0):
        self.L = host_shim()
        self.h = self.L.hs_bounds_create(int(size_x), int(size_y), 1.0 / float(resolution), float(offset_x), float(offset_y), float(extend_factor))

    def _ret(self, fits, geom):
        return bool(fits), (int(geom[0]), int(geom[1]), float(geom[2]), float(geom[3])), (int(geom[4]), int(geom[5]))

    def UpdateMapByRange(self, pts_cells, sensor_pose, half_kernel=0, use_blur=False):
        pts = np.ascontiguousarray(np.asarray(pts_cells, dtype=np.float64).reshape(-1, 2))
        pose = np.ascontiguousarray(sensor_pose, dtype=np.float64)
        geom = np.zeros(6)
        fits = self.L.hs_bounds_update_scan(self.h, pts.ctypes.data, len(pts), pose.ctypes.data, int(half_kernel), int(use_blur), geom.ctypes.data)
        return self._ret(fits, geom)

    def MapSizeCheck(self, pose_world, range_max, offset):
        pose = np.ascontiguousarray(pose_world, dtype=np.float64)
        geom = np.zeros(6)
        fits = self.L.hs_bounds_size_check(self.h, pose.ctypes.data, float(range_max), float(offset), geom.ctypes.data)
        return self._ret(fits, geom)

    def close(self):
        if self.h:
            self.L.hs_bounds_destroy(self.h)
            self.h = None


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
