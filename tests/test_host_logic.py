"""CPU tests of the product's host-side arithmetic (roborts_edu_slam_b200/csrc/rsm_host.h), compiled into a small
shim with g++ (tests/host_shim.cpp) and compared with the oracle: map transforms, pass geometry, the finalisation
(best pose, covariances) on real score arrays, and the 3x3 LDLT solve of the Gauss-Newton matcher."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import random_scenario
from roborts_edu_slam_b200 import matcher, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
c_d, c_i, c_l, c_p = ctypes.c_double, ctypes.c_int, ctypes.c_long, ctypes.c_void_p


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_shim") / "libhostshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           "-I" + os.path.join(ROOT, "roborts_edu_slam_b200", "csrc"), "-o", out, os.path.join(HERE, "host_shim.cpp")])
    L = ctypes.CDLL(out)
    L.hs_normalize_angle.restype = c_d
    L.hs_normalize_angle.argtypes = [c_d]
    L.hs_max_abs_limit.restype = c_d
    L.hs_max_abs_limit.argtypes = [c_d, c_d]
    L.hs_world_to_map.argtypes = [c_d, c_d, c_d, c_p, c_p]
    L.hs_map_to_world.argtypes = [c_d, c_d, c_d, c_p, c_p]
    L.hs_ldlt3.argtypes = [c_p, c_p, c_p]
    L.hs_geometry.argtypes = [ctypes.POINTER(matcher.PassParamStruct), c_i, c_d, c_p, c_p, c_p]
    L.hs_finalize.restype = c_d
    L.hs_finalize.argtypes = [ctypes.POINTER(matcher.PassParamStruct), c_i, c_d, c_p, c_p, c_l, c_p, c_p, ctypes.POINTER(c_i)]
    return L


def _param(p):
    return matcher.CorrelationScanMatchParam.from_array(p).struct()


def test_transforms_and_geometry(shim, oracle, rng):
    for _ in range(200):
        res = float(rng.choice([0.05, 0.025, 0.01, 0.08, 0.1]))
        g = synth.GridSpec(res, 0.15, 480, 480, float(rng.uniform(-30, 30)), float(rng.uniform(-30, 30)))
        w = rng.uniform(-40, 40, 3)
        m = np.zeros(3)
        shim.hs_world_to_map(1.0 / res, g.off_x, g.off_y, w.ctypes.data, m.ctypes.data)
        assert np.array_equal(m, oracle.world_to_map(g, w))
        back = np.zeros(3)
        shim.hs_map_to_world(1.0 / res, g.off_x, g.off_y, m.ctypes.data, back.ctypes.data)
        assert np.array_equal(back, oracle.map_to_world(g, m))
        p = synth.pass_param(float(rng.choice([0.6, 0.8, 0.2, 0.02, 2.0])), float(rng.choice([0.05, 0.1, 0.02, 0.01, 0.025])),
                             float(rng.choice([0.349, 0.523, 1.396, 0.0349, 0.175])), float(rng.choice([0.0349, 0.00349, 0.0087266])),
                             0.3, int(rng.choice([100000, 100, 200, 7])), True, int(rng.integers(0, 3)))
        P = int(rng.integers(1, 1200))
        oi, od = np.zeros(5, dtype=np.int64), np.zeros(4)
        st = _param(p)
        shim.hs_geometry(ctypes.byref(st), P, 1 / (1.0 / res), m.ctypes.data, oi.ctypes.data, od.ctypes.data)
        geo = oracle.geometry(g, p, P, m)
        assert list(oi) == [geo["n_ang"], geo["n_xy"], geo["step"], geo["divisor"], geo["visited"]]
        assert list(od) == [geo["start_x"], geo["start_y"], geo["factor"], geo["start_angle"]]


def test_finalisation_on_oracle_scores(shim, oracle, rng):
    """find_best / positional_cov / angular_cov of the product header on the oracle's score arrays reproduce the
    oracle's (= the reference's) response, best pose and covariance for every pass type -- bit for bit."""
    cases = [(synth.config1(), synth.config1().passes[0])]
    sc4 = synth.config4(1)[0]
    cases += [(sc4, p) for p in sc4.passes]
    for _ in range(6):
        sc = random_scenario(rng, n_points=150, size=160)
        cases.append((sc, synth.pass_param(0.5, 0.05, 0.2, 0.05, 0.3, 100000, bool(rng.integers(0, 2)), int(rng.integers(0, 3)))))
    for sc, p in cases:
        g = sc.grid
        grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
        centre = oracle.world_to_map(g, sc.seed_pose)
        scores = oracle.scores(grid, g, sc.scan_pts, p, centre)
        cov0 = np.diag([3.0, 2.0, 1.0]) + 0.25
        want = oracle.match(grid, g, sc.scan_pts, p, sc.seed_pose, cov=cov0)
        best, cov, navg = np.zeros(4), cov0.copy(), c_i(0)
        st = _param(p)
        resp = shim.hs_finalize(ctypes.byref(st), len(sc.scan_pts), 1 / (1.0 / g.res), centre.ctypes.data, scores.ctypes.data,
                                len(scores), best.ctypes.data, cov.ctypes.data, ctypes.byref(navg))
        assert resp == want["response"] and navg.value == want["n_avg"]
        assert np.array_equal(best, want["best_map"])
        assert np.array_equal(cov, want["cov"])


def test_gauss_newton_host_pieces(shim, oracle, rng):
    for k in range(300):
        A = rng.normal(size=(3, 3))
        H = A @ A.T if k % 2 else (A + A.T)
        H = np.ascontiguousarray(H[np.ix_(*[rng.permutation(3)] * 2)] * 10.0 ** rng.integers(-3, 4))
        if k % 7 == 0:
            H[1, :] = 0.0; H[:, 1] = 0.0          # a zero pivot: pseudo-inverse path
        b = rng.normal(size=3)
        x = np.zeros(3)
        shim.hs_ldlt3(H.ctypes.data, b.ctypes.data, x.ctypes.data)
        assert np.array_equal(x, oracle.ldlt3_solve(H, b)), k
    for a in list(rng.uniform(-20, 20, 200)) + [0.0, np.pi, -np.pi, 2 * np.pi, 3 * np.pi]:
        want = np.fmod(np.fmod(a, 2.0 * np.pi) + 2.0 * np.pi, 2.0 * np.pi)
        want = want - 2.0 * np.pi if want > np.pi else want
        assert shim.hs_normalize_angle(float(a)) == want
    assert shim.hs_max_abs_limit(3.0, 0.5) == 0.5 and shim.hs_max_abs_limit(-3.0, -0.5) == -0.5 and shim.hs_max_abs_limit(0.2, 0.5) == 0.2
