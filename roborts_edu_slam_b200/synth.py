"""Synthetic workloads for the correlative scan matcher.

Not the hot path and not the oracle: this module only *produces inputs* (scans ray-cast from
the Stage floor plans named in BASELINE.json, base-scan chains, grid geometry, pass
parameters).  The same arrays are then fed to the reference build, the oracle restatement
and the CUDA path, so parity never depends on how they were made.

Geometry conventions follow the reference:
  * scan points reach the matcher in CELL units, sensor frame
    (RangeDataContainer::CreateFrom(scan, 1/res), sensor_data_manager.h:99-115);
  * a back-end grid is a square of int((r_max+2)*2/res) cells centred on the current sensor
    pose: offset = -(pose - 0.5*size*cell_len)  (slam_processor.cpp:433-439, 451-455).
"""
import ctypes
import os
from dataclasses import dataclass, field
from typing import List

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

COARSE, FINE, SUPER, FAST = 0, 1, 2, 3
MAP_RES = 0.05  # metres per pixel of every Stage map (maps/*.yaml:2)
ALL_BEAMS = 100000  # use_point_size large enough that every beam is used

_lib = None


def _synth_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsynth.so")
        if not os.path.exists(path):
            raise RuntimeError("libsynth.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        lib.synth_raycast.restype = ctypes.c_int
        lib.synth_raycast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        lib.synth_is_clear.restype = ctypes.c_int
        lib.synth_is_clear.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


_maps = {}


def load_map(name):
    """Occupancy mask (uint8, row 0 = bottom) of 'icra' | 'rm' | 'willow'."""
    if name not in _maps:
        z = np.load(os.path.join(_ROOT, "tests", "golden", "maps_occ.npz"))
        h, w = (int(v) for v in z[name + "_shape"])
        occ = np.unpackbits(z[name + "_bits"])[: h * w].reshape(h, w).astype(np.uint8)
        _maps[name] = np.ascontiguousarray(occ)
    return _maps[name]


def raycast(occ, x, y, heading, beams, fov, r_max, r0=0.05, dr=0.0125, keep_frac=0.95):
    """Scan points in METRES, sensor frame, shape (P, 2) float64."""
    out = np.empty((beams, 2), dtype=np.float64)
    n = _synth_lib().synth_raycast(occ.ctypes.data, occ.shape[1], occ.shape[0], MAP_RES,
                                   float(x), float(y), float(heading), int(beams), float(fov),
                                   float(r_max), float(r0), float(dr), float(keep_frac),
                                   out.ctypes.data)
    return out[:n].copy()


def free_poses(occ, n, seed, clear_px=6):
    """n seeded poses (x, y in metres at pixel centres, heading) in free space."""
    rng = np.random.default_rng(seed)
    lib = _synth_lib()
    h, w = occ.shape
    poses = []
    while len(poses) < n:
        px = int(rng.integers(0, w))
        py = int(rng.integers(0, h))
        th = float(rng.uniform(-np.pi, np.pi))
        if lib.synth_is_clear(occ.ctypes.data, w, h, px, py, clear_px):
            poses.append(((px + 0.5) * MAP_RES, (py + 0.5) * MAP_RES, th))
    return np.array(poses, dtype=np.float64)


def pass_param(size, sres, aoff, ares, threshold, use_point_size, use_center_penalty, kind):
    """The 8-double parameter block shared by ref_driver, the oracle and the C ABI."""
    return np.array([size, sres, aoff, ares, threshold, float(use_point_size),
                     1.0 if use_center_penalty else 0.0, float(kind)], dtype=np.float64)


@dataclass
class GridSpec:
    res: float
    sigma: float
    size_x: int
    size_y: int
    off_x: float
    off_y: float
    default_prob: float = 0.3
    occu_offset: float = 0.88
    use_blur: bool = True


@dataclass
class Scenario:
    name: str
    grid: GridSpec
    base_pts: List[np.ndarray]      # per base scan: (n,2) CELL units, sensor frame
    base_poses: np.ndarray          # (n_scans, 3) world metres / rad
    scan_pts: np.ndarray            # (P,2) CELL units, sensor frame
    seed_pose: np.ndarray           # (3,) world
    truth_pose: np.ndarray          # (3,) world
    passes: List[np.ndarray] = field(default_factory=list)
    grid_centre: np.ndarray = None  # world (x,y) the back-end grid is centred on (None: explicit grid)


def backend_grid(res, sigma, r_max, centre_xy, default_prob=0.3, occu_offset=0.88):
    scale = 1.0 / res
    cell_len = 1.0 / scale
    init_map_size = (r_max + 2.0) * 2
    n = int(init_map_size / res)
    off_x = -(centre_xy[0] - 0.5 * n * cell_len)
    off_y = -(centre_xy[1] - 0.5 * n * cell_len)
    return GridSpec(res, sigma, n, n, off_x, off_y, default_prob, occu_offset, True)


def make_scenario(name, map_name, truth, beams, fov, r_max, res, sigma, passes,
                  seed_delta=(0.12, -0.07, 0.1), chain=8, grid=None):
    occ = load_map(map_name)
    truth = np.asarray(truth, dtype=np.float64)
    inv = 1.0 / res
    scan_m = raycast(occ, truth[0], truth[1], truth[2], beams, fov, r_max)
    base_pts, base_poses = [], []
    for k in range(chain):
        bp = np.array([truth[0] + 0.1 * (k - 4), truth[1] + 0.05 * (k - 4), 0.1 * k])
        pts = raycast(occ, bp[0], bp[1], bp[2], beams, fov, r_max)
        base_pts.append(pts * inv)
        base_poses.append(bp)
    g = grid if grid is not None else backend_grid(res, sigma, r_max, truth[:2])
    return Scenario(name, g, base_pts, np.array(base_poses), scan_m * inv,
                    truth + np.asarray(seed_delta), truth, list(passes),
                    truth[:2].copy() if grid is None else None)


DEG = 0.01745  # the reference writes angles as 0.01745 * degrees (scan_matchers.h:138-153)


def chain_defaults(use_point_size=(ALL_BEAMS, ALL_BEAMS, ALL_BEAMS)):
    """ScanMatchParam in-class defaults (scan_matchers.h:136-157)."""
    return [pass_param(0.8, 0.1, DEG * 80, DEG * 2, 0.6, use_point_size[0], True, COARSE),
            pass_param(0.2, 0.02, DEG * 20, DEG * 2, 0.7, use_point_size[1], True, FINE),
            pass_param(0.02, 0.01, DEG * 2, DEG * 0.2, 0.7, use_point_size[2], True, SUPER)]


def chain_yaml(use_point_size=(ALL_BEAMS, ALL_BEAMS, ALL_BEAMS)):
    """config/real_robot_param.yaml:52-71."""
    return [pass_param(0.6, 0.05, 0.523, 0.0349, 0.6, use_point_size[0], True, COARSE),
            pass_param(0.2, 0.02, 0.175, 0.0349, 0.6, use_point_size[1], True, FINE),
            pass_param(0.02, 0.01, 0.0349, 0.00349, 0.6, use_point_size[2], True, SUPER)]


def config1():
    """360-beam scan vs icra, +-0.3 m / +-20 deg at 0.05 m (BASELINE configs[0])."""
    return make_scenario("cfg1_icra", "icra", (1.0, 1.0, 0.3), 360, np.deg2rad(359.0), 10.0,
                         0.05, 0.15, [pass_param(0.6, 0.05, 0.349, 0.0349, 0.6, ALL_BEAMS, True, COARSE)])


def config2(truth=(8.4, 14.35, 0.3)):
    """720-beam scan, +-1 m / +-45 deg at 0.025 m over rm (BASELINE configs[1])."""
    return make_scenario("cfg2_rm", "rm", truth, 720, np.deg2rad(359.5), 12.0,
                         0.025, 0.03, [pass_param(2.0, 0.025, 0.7854, 0.0087266, 0.6, ALL_BEAMS, True, COARSE)])


def config3(shipped_points=False):
    """Hokuyo 1081-beam, coarse 0.1 + fine 0.01 chain with covariance (BASELINE configs[2])."""
    ups = (100, 100, 200) if shipped_points else (ALL_BEAMS,) * 3
    return make_scenario("cfg3_willow", "willow", (14.375, 28.625, 0.3), 1081, np.deg2rad(270.25), 10.0,
                         0.01, 0.03, chain_defaults(ups))


def config4(n_pairs, seed=1234, first=0):
    """Batched loop closure: seeded scan-vs-submap pairs on willow at 0.05 m (BASELINE configs[3]).

    Returns pairs [first, first+n_pairs) of the seeded sequence so ranks can shard it."""
    occ = load_map("willow")
    poses = free_poses(occ, first + n_pairs, seed)[first:]
    out = []
    for i, p in enumerate(poses):
        out.append(make_scenario("cfg4_pair%d" % (first + i), "willow", p, 1081, np.deg2rad(270.25), 10.0,
                                 0.05, 0.15, chain_yaml()))
    return out


def config5(scale=1.0):
    """Wide relocalisation: +-8 m / 360 deg at 0.05 m on the full willow map (BASELINE configs[4]).

    The grid covers the whole map plus a margin wide enough for the search window and is built
    from a seeded trajectory of base scans.  scale < 1 shrinks the window (the 1/64 variant of
    SURVEY 8d is scale=0.25 with aoff=pi/4)."""
    occ = load_map("willow")
    h, w = occ.shape
    res, sigma, r_max = 0.05, 0.15, 10.0
    truth = np.array([14.375, 28.625, 0.3])
    half = 8.0 * scale
    aoff = np.pi if scale == 1.0 else np.pi / 4
    margin = r_max + half + 1.0
    n_x = int((w * MAP_RES + 2 * margin) / res)
    n_y = int((h * MAP_RES + 2 * margin) / res)
    grid = GridSpec(res, sigma, n_x, n_y, margin, margin, 0.3, 0.88, True)
    traj = free_poses(occ, 64, 77)
    inv = 1.0 / res
    base_pts = [raycast(occ, p[0], p[1], p[2], 1081, np.deg2rad(270.25), r_max) * inv for p in traj]
    scan = raycast(occ, truth[0], truth[1], truth[2], 1081, np.deg2rad(270.25), r_max) * inv
    passes = [pass_param(2 * half, 0.05, aoff, 0.0087266, 0.6, ALL_BEAMS, True, COARSE)]
    return Scenario("cfg5_willow_wide", grid, base_pts, traj, scan,
                    truth + np.array([0.12, -0.07, 0.1]), truth, passes)


def random_grid(rng, size_x, size_y, levels=(0.3, 0.5642, 0.6666, 0.7046, 0.7875, 0.8324, 1.0), p_occ=0.08):
    """A random float32 lookup grid drawn from a small value set (upload-path tests)."""
    vals = np.array(levels, dtype=np.float32)
    g = np.full((size_y, size_x), vals[0], dtype=np.float32)
    mask = rng.random((size_y, size_x)) < p_occ
    g[mask] = vals[rng.integers(1, len(vals), size=int(mask.sum()))]
    return g
