"""CPU tests (-m "not gpu"): the oracle restatement against the reference's golden vectors and,
where the reference was compiled here (oracle/_ref), against the reference itself."""
import numpy as np
import pytest

from helpers import cov_close, golden_grid, golden_names, load_golden, random_scenario, sha
from roborts_edu_slam_b200 import synth


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(oracle, name):
    sc, z = load_golden(name)
    g = sc.grid
    grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
    assert sha(grid) == str(z["grid_sha"])
    assert np.array_equal(grid, golden_grid(z, g))
    centre = oracle.world_to_map(g, sc.seed_pose)
    assert np.array_equal(centre, z["centre_map"])
    scores = oracle.scores(grid, g, sc.scan_pts, sc.passes[0], centre)
    srt = np.sort(scores)[::-1]
    assert sha(srt) == str(z["pass0_sorted_scores_sha"])
    assert np.array_equal(srt[:64], z["pass0_sorted_head"])
    r = oracle.match(grid, g, sc.scan_pts, sc.passes[0], sc.seed_pose)
    assert r["response"] == float(z["pass0_response"])
    assert np.array_equal(r["pose"], z["pass0_pose"])
    assert np.array_equal(r["cov"], z["pass0_cov"])
    assert np.array_equal(r["best_map"], z["pass0_best"])
    if "chain_score" in z.files:
        rc = oracle.match_chain(grid, g, sc.scan_pts, sc.passes, sc.seed_pose)
        assert rc["score"] == float(z["chain_score"])
        assert np.array_equal(rc["pose"], z["chain_pose"])
        assert np.array_equal(rc["cov"], z["chain_cov"])
        assert np.array_equal(rc["responses"], z["chain_responses"])


def _compare_with_ref(oracle, ref, sc, chain):
    g = sc.grid
    grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
    m = ref.create_map(g)
    try:
        ref.build_map(m, g, sc.base_pts, sc.base_poses)
        assert np.array_equal(grid, ref.read_map(m, g))
        assert np.array_equal(oracle.world_to_map(g, sc.seed_pose), ref.world_to_map(m, sc.seed_pose))
        assert np.array_equal(oracle.map_to_world(g, [100.25, 77.5, 0.3]), ref.map_to_world(m, [100.25, 77.5, 0.3]))
        for p in sc.passes:
            centre = ref.world_to_map(m, sc.seed_pose)
            geo = oracle.geometry(g, p, len(sc.scan_pts), centre)
            so = oracle.scores(grid, g, sc.scan_pts, p, centre)
            cr = ref.candidates(m, sc.scan_pts, p, centre)
            ix = np.rint((cr["x"] - geo["start_x"]) / geo["factor"]).astype(np.int64)
            iy = np.rint((cr["y"] - geo["start_y"]) / geo["factor"]).astype(np.int64)
            k = (cr["angle_index"].astype(np.int64) * geo["n_xy"] + ix) * geo["n_xy"] + iy
            assert np.array_equal(np.sort(k), np.arange(len(k)))
            assert np.array_equal(so[k], cr["score"])          # every candidate score, bit for bit
            ro = oracle.match(grid, g, sc.scan_pts, p, sc.seed_pose)
            rr = ref.match(m, sc.scan_pts, p, sc.seed_pose)
            assert ro["response"] == rr["response"]
            assert np.array_equal(ro["pose"], rr["pose"])
            assert np.array_equal(ro["cov"], rr["cov"])       # same sort order even under ties
        if chain:
            ro = oracle.match_chain(grid, g, sc.scan_pts, sc.passes, sc.seed_pose)
            rr = ref.match_chain(m, sc.scan_pts, sc.passes, sc.seed_pose)
            assert ro["score"] == rr["score"]
            assert np.array_equal(ro["pose"], rr["pose"])
            assert np.array_equal(ro["cov"], rr["cov"])
            assert np.array_equal(ro["responses"], rr["responses"])
    finally:
        ref.destroy_map(m)


def test_oracle_vs_reference_named_configs(oracle, ref):
    _compare_with_ref(oracle, ref, synth.config1(), False)
    _compare_with_ref(oracle, ref, synth.config3(True), True)
    for sc in synth.config4(3, seed=99):
        _compare_with_ref(oracle, ref, sc, True)


def test_oracle_vs_reference_random(oracle, ref, rng):
    for trial in range(12):
        sc = random_scenario(rng, n_points=int(rng.integers(5, 300)))
        kind = [synth.COARSE, synth.FINE, synth.SUPER][trial % 3]
        ups = int(rng.choice([20, 50, 100000]))
        pen = bool(trial % 2)
        sres = float(rng.choice([0.05, 0.02, 0.1, 0.025]))
        sc.passes = [synth.pass_param(sres * int(rng.integers(2, 9)), sres, float(rng.uniform(0.02, 0.4)),
                                      float(rng.uniform(0.005, 0.05)), 0.3, ups, pen, kind)]
        _compare_with_ref(oracle, ref, sc, False)


def test_oracle_vs_reference_ties(oracle, ref):
    # binary grid + no penalty: the winner depends on the unstable sort's tie order
    sc, _ = load_golden("ties_icra")
    _compare_with_ref(oracle, ref, sc, True)


def test_blur_kernel(oracle, ref):
    for sigma, res in [(0.15, 0.05), (0.03, 0.025), (0.03, 0.01), (0.24, 0.08), (0.4, 0.1), (0.01, 0.05), (1.0, 0.05)]:
        ho, ko = oracle.blur_kernel(sigma, res)
        hr, kr = ref.blur_kernel(sigma, res)
        assert ho == hr
        if ho >= 0:
            assert np.array_equal(ko, kr)


def test_oracle_degenerate_inputs(oracle):
    sc, z = load_golden("cfg1_icra")
    g = sc.grid
    grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
    # empty scan: response 0, pose and covariance untouched (correlate_scan_matcher.h:792-795)
    cov = np.arange(9.0).reshape(3, 3)
    r = oracle.match(grid, g, np.zeros((0, 2)), sc.passes[0], sc.seed_pose, cov)
    assert r["response"] == 0.0 and np.array_equal(r["pose"], sc.seed_pose) and np.array_equal(r["cov"], cov)
    # scan in unknown space: every candidate ties; response below threshold -> pose untouched
    flat = np.full_like(grid, np.float32(0.3))
    r = oracle.match(flat, g, sc.scan_pts, sc.passes[0], sc.seed_pose)
    assert r["response"] < 0.6 and np.array_equal(r["pose"], sc.seed_pose)
    assert r["n_avg"] > 1
    nopen = sc.passes[0].copy()
    nopen[6] = 0.0   # no centre penalty: all 3549 candidates tie exactly
    r = oracle.match(flat, g, sc.scan_pts, nopen, sc.seed_pose)
    assert r["n_avg"] == 3549 and np.array_equal(r["pose"], sc.seed_pose)
    assert cov_close(r["cov"], r["cov"])


# ---- map check: MapFeedbackResponsePenalty / MapCheckPenalize ---------------------
def test_mapcheck_oracle_matches_golden(oracle):
    """The restatement against the coefficients the reference returned (committed fixtures)."""
    from helpers import load_mapcheck, mapcheck_names
    names = mapcheck_names()
    assert names
    for name in names:
        g, occ, z = load_mapcheck(name)
        for k, ps in enumerate(z["param_sets"]):
            for i, pose in enumerate(z["poses"]):
                got = oracle.map_check_penalize(occ, g, z["scan_pts"], pose, int(ps[0]), ps[1], ps[2], bool(ps[3]), ps[4:6])
                assert got == z["coeff"][k, i], (name, k, i, got, z["coeff"][k, i])


def test_mapcheck_oracle_matches_reference(oracle, ref, rng):
    """Live reference (its own UpdateMapByRange builds the publishing map) vs the restatement on random
    poses, scan subsets and knobs, with and without the loop-closure logistic."""
    sc = synth.config1()
    g = sc.grid
    m = ref.pubmap_create(g)
    for pts, pose in zip(sc.base_pts, sc.base_poses):
        assert ref.pubmap_update(m, pts, pose) == 0
    _, _, occ = ref.pubmap_read(m, g)
    assert occ.sum() > 100
    seen = set()
    for k in range(400):
        pose = sc.truth_pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)]) * rng.choice([0.05, 0.3, 1.0])
        n = int(rng.choice([len(sc.scan_pts), 150, 250, 30, 0]))
        cp = int(rng.choice([100, 20, 7, 2]))       # 1 makes the reference divide by zero (occu_grid_map.h:367)
        tol = float(rng.choice([2.5, 0.0, 1.0]))
        gain = float(rng.choice([0.015, 0.1]))
        logistic = bool(k & 1)
        org = (0.0, 0.0) if k % 3 else (float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        a = ref.pubmap_penalty(m, sc.scan_pts[:n], pose, cp, tol, gain, False, logistic, org)
        b = oracle.map_check_penalize(occ, g, sc.scan_pts[:n], pose, cp, tol, gain, logistic, org)
        assert a == b, (k, a, b)
        seen.add(a)
    assert len(seen) > 20
    ref.pubmap_destroy(m)


# ---- Gauss-Newton matcher: BasedOptimizeScanMatch ---------------------------------
def test_ldlt3_solve_properties(oracle, rng):
    """The restated 3x3 LDLT (the one step that cannot be pinned against real Eigen here): solves SPD and
    indefinite symmetric systems to rounding accuracy whatever the pivot order, returns the pseudo-inverse
    solution for a singular diagonal block and zeros for H = 0."""
    for k in range(300):
        A = rng.normal(size=(3, 3))
        H = A @ A.T if k % 2 else (A + A.T)
        H = H[np.ix_(*[rng.permutation(3)] * 2)] * 10.0 ** rng.integers(-3, 4)
        b = rng.normal(size=3)
        x = oracle.ldlt3_solve(H, b)
        assert np.allclose(H @ x, b, rtol=0, atol=1e-9 * np.linalg.cond(H) * max(1.0, np.abs(b).max()))
    assert np.array_equal(oracle.ldlt3_solve(np.zeros((3, 3)), [1.0, 2.0, 3.0]), np.zeros(3))
    x = oracle.ldlt3_solve(np.diag([4.0, 0.0, 2.0]), [8.0, 5.0, 1.0])
    assert np.array_equal(x, [2.0, 0.0, 0.5])


def test_optimize_oracle_matches_golden(oracle):
    """Restatement vs the outputs of the reference's own optimize_scan_matcher.h (committed fixture):
    cost and pose bit for bit, on the fine and the coarse map, plus the chain that starts with the optimiser."""
    from helpers import optimize_cases
    z, cases = optimize_cases()
    for tag, sc, gc, base_c, scan_c in cases:
        gf = sc.grid
        grid_f = oracle.build_grid(gf, sc.base_pts, sc.base_poses)
        grid_c = oracle.build_grid(gc, base_c, sc.base_poses)
        for mi, (grid, g, pts) in enumerate(((grid_f, gf, sc.scan_pts), (grid_c, gc, scan_c))):
            for oi, op in enumerate(z["opt_sets"]):
                for si, d in enumerate(z["seed_deltas"]):
                    r = oracle.optimize(grid, g, pts, op, sc.truth_pose + d)
                    assert r["cost"] == z[tag + "_cost"][mi, oi, si], (tag, mi, oi, si)
                    assert np.array_equal(r["pose"], z[tag + "_pose"][mi, oi, si])
        for fi, fc in enumerate(z["failed_costs"]):
            for ui, use_fine in enumerate((True, False)):
                for si, d in enumerate(z["seed_deltas"][:4]):
                    r = oracle.match_chain_opt(grid_c, gc, scan_c, grid_f, gf, sc.scan_pts, synth.chain_yaml(), z["opt_sets"][0],
                                               fc, sc.truth_pose + d, use_fine=use_fine)
                    assert r["score"] == z[tag + "_chain_score"][fi, ui, si], (tag, fi, ui, si)
                    assert np.array_equal(r["pose"], z[tag + "_chain_pose"][fi, ui, si])
                    assert np.array_equal(r["cov"], z[tag + "_chain_cov"][fi, ui, si])
                    assert r["optimize_cost"] == z[tag + "_chain_resp"][fi, ui, si, 0]
                    assert np.array_equal(r["responses"], z[tag + "_chain_resp"][fi, ui, si, 1:])
    # both branches of scan_matchers.h:224-226 are exercised by the fixture
    resp = z["pair0_chain_resp"]
    assert (resp[:, 0, :, 1] == 0).any() and (resp[:, 0, :, 1] != 0).any()


def test_optimize_oracle_matches_reference(oracle, ref, rng):
    """Live reference header vs the restatement on random problems (random grids, float-valued cells,
    scans that partly leave the map, random knobs)."""
    compared = 0
    for k in range(60):
        sc = random_scenario(rng, n_points=int(rng.integers(5, 300)), size=int(rng.choice([96, 160])))
        g = sc.grid
        grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
        if k % 4 == 3:   # arbitrary float cells instead of the few blur levels
            grid = rng.random(grid.shape).astype(np.float32)
        m = ref.create_map(g)
        ref.write_map(m, grid)
        op = (int(rng.integers(1, 12)), float(rng.choice([0.1, 1.0, 1e-6])), float(rng.choice([0.5, 2.0, 0.0])),
              float(rng.choice([0.5, 0.05])), float(rng.choice([0.5, 0.2, 0.02])))
        scale = rng.choice([0.02, 0.2, 1.0, 4.0])
        seed = sc.truth_pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)]) * scale
        a = oracle.optimize(grid, g, sc.scan_pts, op, seed)
        b = ref.optimize(m, sc.scan_pts, op, seed)
        ref.destroy_map(m)
        if a["oob_reads"]:     # a point in the strip size_y-1 < y < size_y: the reference reads past its cell array
            continue
        compared += 1
        assert a["cost"] == b["cost"] and np.array_equal(a["pose"], b["pose"]), (k, a, b)
    assert compared >= 40


# ---- front-end maps: incremental updates, extension, publishing map ---------------------------
def test_frontend_maps_oracle_matches_golden(oracle):
    """The restatement of the incremental scan-match map update, ExtendSize's copy and the publishing map's
    ray-traced CountCell update replays the 44-scan trajectory and reproduces the checksums the reference's own
    maps had after every step (fixtures), with the resize decisions coming from the product's host policy
    (csrc/rsm_host.h compiled with g++ through tests/host_shim.cpp: no nvcc, no librsm.so)."""
    import hashlib
    import importlib.util
    import os
    from helpers import ShimBounds
    here = os.path.dirname(os.path.abspath(__file__))
    mods = {}
    import sys
    sys.path.insert(0, os.path.join(here, "golden"))
    try:
        for name in ("make_frontend", "make_pubmap"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(here, "golden", name + ".py"))
            mods[name] = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mods[name])
    finally:
        sys.path.pop(0)
    mf, mp = mods["make_frontend"], mods["make_pubmap"]
    g = mf.spec()
    poses, pts = mf.trajectory(), mf.scans()
    zf = np.load(os.path.join(here, "golden", "frontend_willow.npz"), allow_pickle=False)
    zp = np.load(os.path.join(here, "golden", "pubmap_willow.npz"), allow_pickle=False)
    # scan-match map
    grid = np.full((g.size_y, g.size_x), np.float32(0.5), dtype=np.float32)
    grid[0, 0] = np.float32(g.default_prob)
    cur = synth.GridSpec(g.res, g.sigma, g.size_x, g.size_y, g.off_x, g.off_y, g.default_prob, g.occu_offset, True)
    bounds = ShimBounds(g.size_x, g.size_y, g.res, g.off_x, g.off_y, mf.EXTEND)
    for k, (p, s) in enumerate(zip(poses, pts)):
        fits, geom, pre = bounds.UpdateMapByRange(s, p, 2, True)
        assert fits == bool(zf["stamped"][k])
        if fits:
            oracle.grid_stamp(grid, cur, s, p)
        else:
            grid = oracle.grid_extend(grid, geom[0], geom[1], pre, 0.5, g.default_prob)
            cur = synth.GridSpec(g.res, g.sigma, geom[0], geom[1], geom[2], geom[3], g.default_prob, g.occu_offset, True)
        assert hashlib.sha256(grid.tobytes()).hexdigest() == str(zf["shas"][k]), "scan-match map, step %d" % k
    bounds.close()
    # publishing map
    pm = oracle.pubmap_new(g.size_x, g.size_y)
    off = (g.off_x, g.off_y)
    bounds = ShimBounds(g.size_x, g.size_y, g.res, g.off_x, g.off_y, mf.EXTEND)
    for k, (p, s) in enumerate(zip(poses, pts)):
        f = mp.factors(k)
        fits, geom, pre = bounds.UpdateMapByRange(s, p, 0, False)
        assert fits == bool(zp["stamped"][k])
        if fits:
            oracle.pubmap_update(pm, g.res, off[0], off[1], s, p, f[0], f[1])
        else:
            pm = oracle.pubmap_extend(pm, geom[0], geom[1], pre)
            off = (geom[2], geom[3])
        h = hashlib.sha256()
        for a in (pm["value"], pm["passc"], pm["hit"]):
            h.update(a.tobytes())
        assert h.hexdigest() == str(zp["shas"][k]), "publishing map, step %d" % k
    bounds.close()
    occ = ((pm["passc"] >= np.float32(zp["knobs"][1])) & ~(pm["value"] < np.float32(zp["knobs"][0]))).astype(np.uint8)
    assert np.array_equal(occ.ravel(), np.unpackbits(zp["occ_packed"])[: occ.size])
    gfin = synth.GridSpec(g.res, 0.0, occ.shape[1], occ.shape[0], off[0], off[1], 0.5, 0.88, False)
    got = np.array([oracle.map_check_penalize(occ, gfin, pts[-1], q, 100, 2.5, 0.015, True) for q in zp["check_poses"]])
    assert np.array_equal(got, zp["coeff"])


def test_config5_full_size_golden(oracle):
    """BASELINE configs[4] at full size (74.3 M candidates, 7.0e10 evaluations): the restatement, scored per angle slice on
    all host threads, against the fixture the reference's own code produced (tests/golden/make_config5.py)."""
    import hashlib
    from helpers import load_config5_golden
    sc, z = load_config5_golden()
    g, p = sc.grid, sc.passes[0]
    grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
    assert sha(grid) == str(z["grid_sha"])
    centre = oracle.world_to_map(g, sc.seed_pose)
    assert np.array_equal(centre, z["centre_map"])
    scores = oracle.scores_threaded(grid, g, sc.scan_pts, p, centre)
    assert hashlib.sha256(scores.tobytes()).hexdigest() == str(z["scores_sha"])
    fin = oracle.finish_scores(scores, g, len(sc.scan_pts), p, sc.seed_pose)
    assert fin["response"] == float(z["response"]) and np.array_equal(fin["pose"], z["pose"]) and np.array_equal(fin["cov"], z["cov"])
    assert fin["n_avg"] == int(z["n_avg"]) and np.array_equal(fin["best_map"], z["best_map"])
