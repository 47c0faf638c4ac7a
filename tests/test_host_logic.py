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


def _bounds_api(shim):
    shim.hs_bounds_create.restype = c_p
    shim.hs_bounds_create.argtypes = [c_i, c_i, c_d, c_d, c_d, c_d]
    shim.hs_bounds_destroy.argtypes = [c_p]
    shim.hs_bounds_update_scan.restype = c_i
    shim.hs_bounds_update_scan.argtypes = [c_p, c_p, c_i, c_p, c_i, c_i, c_p]
    shim.hs_bounds_size_check.restype = c_i
    shim.hs_bounds_size_check.argtypes = [c_p, c_p, c_d, c_d, c_p]


def test_resize_policy_matches_fixtures(shim):
    """The host restatement of UpdateBound / ExtendSize (MapBounds) reproduces, step by step, the decisions and the
    geometry the reference took on the 44-scan trajectory (fixtures of make_frontend.py / make_pubmap.py): stamped
    or extended, size, map offset -- for the blurred scan-match map and for the publishing map."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_frontend", os.path.join(HERE, "golden", "make_frontend.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    _bounds_api(shim)
    g = mf.spec()
    poses, pts = mf.trajectory(), mf.scans()
    for fixture, half, blur in (("frontend_willow.npz", 2, 1), ("pubmap_willow.npz", 0, 0)):
        z = np.load(os.path.join(HERE, "golden", fixture), allow_pickle=False)
        b = shim.hs_bounds_create(g.size_x, g.size_y, 1.0 / g.res, g.off_x, g.off_y, mf.EXTEND)
        prev = (g.size_x, g.size_y, g.off_x, g.off_y)
        for k, (p, s) in enumerate(zip(poses, pts)):
            geom = np.zeros(6)
            s = np.ascontiguousarray(s)
            ok = shim.hs_bounds_update_scan(b, s.ctypes.data, len(s), np.ascontiguousarray(p).ctypes.data, half, blur, geom.ctypes.data)
            want = z["geom"][k]
            assert bool(ok) == bool(z["stamped"][k]), (fixture, k)
            assert (int(geom[0]), int(geom[1]), geom[2], geom[3]) == (int(want[0]), int(want[1]), float(want[2]), float(want[3])), (fixture, k)
            if not ok:       # where the old cell (0, 0) went
                assert (int(geom[4]), int(geom[5])) == (int(round((geom[2] - prev[2]) / g.res)), int(round((geom[3] - prev[3]) / g.res)))
            prev = (int(geom[0]), int(geom[1]), geom[2], geom[3])
        shim.hs_bounds_destroy(b)


def test_resize_policy_matches_reference(shim, ref, rng):
    """Random walks (with MapSizeCheck calls in between, as ScanMatchers::ScanMatch makes them) against the live
    reference map object: same decisions, same sizes, same offsets, bit for bit."""
    _bounds_api(shim)
    occ = synth.load_map("willow")
    total_ext = 0
    for trial in range(8):
        res = float(rng.choice([0.05, 0.1, 0.025]))
        sigma = res * 3
        half = int((sigma / res) * np.sqrt(np.log(2)))
        n = int(rng.choice([200, 480]))
        start = np.array([14.375, 28.625, 0.3]) + rng.uniform(-1, 1, 3)
        g = synth.GridSpec(res, sigma, n, n, -(start[0] - 0.5 * n * res), -(start[1] - 0.5 * n * res), 0.3, 0.88, True)
        extend = float(rng.choice([0.2, 1.0, 0.05]))
        m = ref.frontend_map_create(g, extend)
        b = shim.hs_bounds_create(g.size_x, g.size_y, 1.0 / res, g.off_x, g.off_y, extend)
        p = start.copy()
        n_ext = 0
        drift = rng.uniform(-0.5, 0.5, 2)
        for k in range(40):
            p = p + np.array([drift[0] + rng.uniform(-0.5, 0.5), drift[1] + rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3)])
            s = np.ascontiguousarray(synth.raycast(occ, p[0], p[1], p[2], 181, np.deg2rad(270.25), 6.0) / res)
            if len(s) == 0:
                continue
            geom = np.zeros(6)
            if k % 5 == 4:
                ok_r, ge_r = ref.frontend_map_size_check(m, p, 6.0, 0.6)
                ok = shim.hs_bounds_size_check(b, np.ascontiguousarray(p).ctypes.data, 6.0, 0.6, geom.ctypes.data)
            else:
                ok_r, ge_r = ref.frontend_map_update(m, s, p, True)
                ok = shim.hs_bounds_update_scan(b, s.ctypes.data, len(s), np.ascontiguousarray(p).ctypes.data, half, 1, geom.ctypes.data)
            assert bool(ok) == ok_r, (trial, k)
            assert (int(geom[0]), int(geom[1]), geom[2], geom[3]) == ge_r, (trial, k, geom, ge_r)
            n_ext += 0 if ok_r else 1
        total_ext += n_ext
        ref.destroy_map(m)
        shim.hs_bounds_destroy(b)
    assert total_ext >= 6


def test_resize_policy_through_the_c_abi(ref, rng):
    """The same policy through librsm.so's host-only entry points (rsm_map_bounds_*; callable without a GPU): the
    fixture trajectories and a random walk against the live reference, plus rsm_blur_half_size."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_frontend", os.path.join(HERE, "golden", "make_frontend.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    lib = matcher.load_library()
    assert lib.rsm_blur_half_size(0.15, 0.05) == 2 and lib.rsm_blur_half_size(0.4, 0.1) == 3
    assert lib.rsm_blur_half_size(0.03, 0.025) == 0 and lib.rsm_blur_half_size(0.0, 0.05) == -1
    g = mf.spec()
    poses, pts = mf.trajectory(), mf.scans()
    for fixture, half, blur in (("frontend_willow.npz", 2, True), ("pubmap_willow.npz", 0, False)):
        z = np.load(os.path.join(HERE, "golden", fixture), allow_pickle=False)
        mb = matcher.MapBounds(g.size_x, g.size_y, g.res, g.off_x, g.off_y, mf.EXTEND)
        for k, (p, s) in enumerate(zip(poses, pts)):
            fits, geom, pre = mb.UpdateMapByRange(s, p, half, blur)
            want = z["geom"][k]
            assert fits == bool(z["stamped"][k]) and geom == (int(want[0]), int(want[1]), float(want[2]), float(want[3])), (fixture, k)
        mb.close()
    occ = synth.load_map("willow")
    res, n = 0.05, 300
    start = np.array([14.375, 28.625, 0.3])
    gs = synth.GridSpec(res, 0.15, n, n, -(start[0] - 0.5 * n * res), -(start[1] - 0.5 * n * res), 0.3, 0.88, True)
    m = ref.frontend_map_create(gs, 0.3)
    mb = matcher.MapBounds(n, n, res, gs.off_x, gs.off_y, 0.3)
    p, n_ext = start.copy(), 0
    for k in range(60):
        p = p + np.array([0.35 + rng.uniform(-0.4, 0.4), -0.25 + rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3)])
        s = synth.raycast(occ, p[0], p[1], p[2], 181, np.deg2rad(270.25), 6.0) / res
        if len(s) == 0:
            continue
        if k % 7 == 6:
            ok_r, ge_r = ref.frontend_map_size_check(m, p, 6.0, 0.6)
            fits, geom, pre = mb.MapSizeCheck(p, 6.0, 0.6)
        else:
            ok_r, ge_r = ref.frontend_map_update(m, s, p, True)
            fits, geom, pre = mb.UpdateMapByRange(s, p, 2, True)
        assert fits == ok_r and geom == ge_r, (k, geom, ge_r)
        n_ext += 0 if ok_r else 1
    assert n_ext >= 2
    ref.destroy_map(m)
    mb.close()


def test_angle_tables_are_libm_cos_and_sin(shim, rng):
    """The product fills its angle tables with glibc's sincos; the reference calls std::cos and std::sin
    (correlate_scan_matcher.h:171-172).  They must be the same bits: 2 M search angles, small and large."""
    shim.hs_angle_trig.argtypes = [c_p, c_l, c_p]
    ang = np.concatenate([rng.uniform(-7.0, 7.0, 1_500_000), rng.uniform(-1e4, 1e4, 500_000), np.array([0.0, -0.0, np.pi, -np.pi / 2])])
    out = np.zeros((len(ang), 2))
    shim.hs_angle_trig(ang.ctypes.data, len(ang), out.ctypes.data)
    # numpy's cos / sin on float64 are its own SIMD kernels, not libm: compare with the C library through ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cos.restype = libm.sin.restype = c_d
    libm.cos.argtypes = libm.sin.argtypes = [c_d]
    idx = rng.integers(0, len(ang), 200_000)
    for i in idx:
        a = float(ang[i])
        assert out[i, 0] == libm.cos(a) and out[i, 1] == libm.sin(a), a


@pytest.mark.parametrize("runs,variant,max_ctas", [
    ([(707, 81, 181)], 0, 148),                 # BASELINE configs[1]: 181 items on 148 SMs
    ([(707, 81, 181)], 2, 148),
    ([(938, 321, 721)], 0, 148),                # configs[4]: 16 tiles per angle, partial tiles on two sides
    ([(360, 49, 3)], 1, 148),                   # three items over many CTAs: every item has many parts
    ([(360, 49, 3)], 1, 7),
    ([(100, 81, 5), (707, 65, 9), (64, 97, 2), (2048, 81, 1)], 0, 33),   # a batch of different windows
    ([(30, 81, 4)], 0, 148),                    # fewer beams than the cut guard
])
def test_stream_plan_covers_every_beam_once(runs, variant, max_ctas):
    """Host side of the staged kernel's stream plan (csrc/rsm_api.cu plan_stream): the shares tile the (item, beam)
    sequence exactly, shared items have consistent tickets, parts and disjoint partial slots."""
    shares, n_items, n_tickets, n_slots = matcher.stream_plan(runs, variant, max_ctas)
    tile = {0: (81, 96), 1: (64, 64), 2: (96, 96)}[variant]
    beams = []
    for v, n_xy, n_ang in runs:
        beams += [v] * (n_ang * (-(-n_xy // tile[0])) * (-(-n_xy // tile[1])))
    assert n_items == len(beams) and 1 <= len(shares) <= max_ctas
    seen = [np.zeros(v, dtype=np.int32) for v in beams]
    users = {}
    prev_end = (0, 0)
    for s in shares:
        i0, b0, i1, b1, t0, s0, p0, n0, t1, s1, p1, n1 = [int(v) for v in s]
        assert (i0, b0) == prev_end, "shares are contiguous"
        assert i0 <= i1 and 0 <= b0 < beams[i0] and 0 < b1 <= beams[i1] and (i0 < i1 or b0 < b1)
        for it in range(i0, i1 + 1):
            lo, hi = (b0 if it == i0 else 0), (b1 if it == i1 else beams[it])
            assert hi > lo, "no empty visit"
            seen[it][lo:hi] += 1
            whole = lo == 0 and hi == beams[it]
            t, sl, pa, n = (t0, s0, p0, n0) if it == i0 else (t1, s1, p1, n1) if it == i1 else (-1, 0, 0, 0)
            assert (t < 0) == whole, "an item is shared exactly when a share sees a part of it"
            if t >= 0:
                users.setdefault(it, []).append((t, sl, pa, n))
        prev_end = (i1 + 1, 0) if b1 == beams[i1] else (i1, b1)
    assert prev_end == (n_items, 0)
    assert all((v == 1).all() for v in seen), "every beam of every item exactly once"
    assert len(users) == n_tickets
    slots = []
    for it, us in users.items():
        t, sl, _, n = us[0]
        assert len(us) == n >= 2 and all(u[0] == t and u[1] == sl and u[3] == n for u in us)
        assert sorted(u[2] for u in us) == list(range(n))
        slots += list(range(sl, sl + n))
    assert sorted(u[0][0] for u in users.values()) == list(range(n_tickets))
    assert sorted(slots) == list(range(n_slots))
    if len(shares) == max_ctas and len(runs) == 1 and variant != 1:
        # equal shares: beams x tile weight per share within 15 % of the mean (single job, full-size launch)
        v, n_xy, _ = runs[0]
        if n_xy <= tile[0]:
            w = [sum((b1 if it == i1 else v) - (b0 if it == i0 else 0) for it in range(i0, i1 + 1)) for i0, b0, i1, b1 in shares[:, :4]]
            assert max(w) <= 1.15 * np.mean(w) and min(w) >= 0.85 * np.mean(w)


def _check_plan(runs, variant, max_ctas):
    shares, n_items, n_tickets, n_slots = matcher.stream_plan(runs, variant, max_ctas)
    tile = {0: (81, 96), 1: (64, 64), 2: (96, 96)}[variant]
    beams = []
    for v, n_xy, n_ang in runs:
        beams += [v] * (n_ang * (-(-n_xy // tile[0])) * (-(-n_xy // tile[1])))
    assert n_items == len(beams) and 1 <= len(shares) <= max_ctas
    covered = np.zeros(len(beams), dtype=np.int64)
    parts = {}
    prev_end = (0, 0)
    for s in shares:
        i0, b0, i1, b1, t0, s0, p0, n0, t1, s1, p1, n1 = [int(v) for v in s]
        assert (i0, b0) == prev_end and i0 <= i1 and (i0 < i1 or b0 < b1)
        for it in (range(i0, i1 + 1) if i1 - i0 < 3 else [i0, i1]):
            lo, hi = (b0 if it == i0 else 0), (b1 if it == i1 else beams[it])
            assert 0 <= lo < hi <= beams[it]
            covered[it] += hi - lo
            whole = lo == 0 and hi == beams[it]
            t, sl, pa, n = (t0, s0, p0, n0) if it == i0 else (t1, s1, p1, n1) if it == i1 else (-1, 0, 0, 0)
            assert (t < 0) == whole
            if t >= 0:
                parts.setdefault(t, []).append((sl, pa, n))
        if i1 - i0 >= 3:
            covered[i0 + 1:i1] += np.array(beams[i0 + 1:i1])
        prev_end = (i1 + 1, 0) if b1 == beams[i1] else (i1, b1)
    assert prev_end == (n_items, 0) and np.array_equal(covered, np.array(beams))
    assert sorted(parts) == list(range(n_tickets))
    slots = []
    for t, us in parts.items():
        assert len({u[0] for u in us}) == 1 and len(us) == us[0][2] >= 2 and sorted(u[1] for u in us) == list(range(len(us)))
        slots += list(range(us[0][0], us[0][0] + len(us)))
    assert sorted(slots) == list(range(n_slots))


def test_stream_plan_random_launch_shapes():
    """Seeded sweep over launch shapes (jobs of different beam counts, windows, angle counts; few and many CTAs): the
    invariants of test_stream_plan_covers_every_beam_once on 300 random plans."""
    rng = np.random.default_rng(20261019)
    for _ in range(300):
        variant = int(rng.integers(0, 3))
        n_jobs = int(rng.integers(1, 6))
        n_xy0 = int(rng.integers(48, 200))
        runs = [(int(rng.integers(2, 2049)), n_xy0, int(rng.integers(1, 40))) for _ in range(n_jobs)]
        _check_plan(runs, variant, int(rng.integers(1, 300)))
