"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the
same inputs and against the reference's golden vectors.

Bars (BASELINE.json north_star): bit-exact lookup grid, candidate scores, response and pose;
covariance within 1e-6 relative (cov_close).  Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from helpers import cov_close, golden_grid, golden_names, load_golden, random_scenario
from roborts_edu_slam_b200 import matcher, synth

pytestmark = pytest.mark.gpu


def device_grid(ctx, sc, rasterize=True, oracle=None):
    g = sc.grid
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    if rasterize:
        dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    else:
        dg.upload(oracle.build_grid(g, sc.base_pts, sc.base_poses))
    return dg


def assert_pass_equal(got_resp, got_pose, got_cov, want):
    assert got_resp == want["response"]
    assert np.array_equal(got_pose, want["pose"])
    assert cov_close(got_cov, want["cov"])


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(ctx, name):
    """Everything the reference produced for the committed fixtures, through the CUDA path."""
    sc, z = load_golden(name)
    g = sc.grid
    dg = device_grid(ctx, sc)
    assert np.array_equal(dg.download(), golden_grid(z, g))                 # rasteriser, every cell
    assert np.array_equal(dg.GetMapCoordsPose(sc.seed_pose), z["centre_map"])
    m = matcher.BasedCorrelationScanMatch(ctx)
    scores = m.scores(dg, sc.scan_pts, sc.passes[0], sc.seed_pose)
    assert np.array_equal(np.sort(scores)[::-1][:64], z["pass0_sorted_head"])
    import hashlib
    assert hashlib.sha256(np.sort(scores)[::-1].tobytes()).hexdigest() == str(z["pass0_sorted_scores_sha"])
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r = m.ScanMatch(dg, sc.scan_pts, sc.passes[0], pose, cov)
    assert r == float(z["pass0_response"])
    assert np.array_equal(pose, z["pass0_pose"])
    assert cov_close(cov, z["pass0_cov"])
    assert np.array_equal([m.last_detail.best_pose_map[i] for i in range(3)] + [m.last_detail.best_score], z["pass0_best"])
    if "chain_score" in z.files:
        sm = matcher.ScanMatchers(ctx, sc.passes)
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        s = sm.ScanMatch(sc.scan_pts, dg, pose, cov)
        assert s == float(z["chain_score"])
        assert np.array_equal(pose, z["chain_pose"])
        assert cov_close(cov, z["chain_cov"])
        assert np.array_equal(sm.last_responses, z["chain_responses"])
    dg.close()


def test_scores_and_indices_small(ctx, oracle, rng):
    """Every candidate score bit-equal on random small problems: all tile shapes (n_xy 1..40),
    beam subsampling, penalty on/off, all pass types, f < 1, f = 1, f > 1."""
    m = matcher.BasedCorrelationScanMatch(ctx)
    for trial in range(40):
        sc = random_scenario(rng, n_points=int(rng.integers(3, 400)), size=200)
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        dg = matcher.ScanMatchMap.from_spec(ctx, sc.grid)
        dg.upload(grid)
        sres = float(rng.choice([0.05, 0.02, 0.1, 0.025, 0.01, 0.2]))
        n_steps = int(rng.integers(0, 40))
        if sres * n_steps > 1.6:
            n_steps = int(1.6 / sres)
        p = synth.pass_param(sres * n_steps, sres, float(rng.uniform(0.0, 0.5)), float(rng.uniform(0.004, 0.06)),
                             0.3, int(rng.choice([10, 37, 100, 100000])), bool(trial % 2), trial % 3)
        so = oracle.scores(grid, sc.grid, sc.scan_pts, p, oracle.world_to_map(sc.grid, sc.seed_pose))
        sd = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
        assert so.shape == sd.shape
        assert np.array_equal(so, sd), "trial %d: %d of %d scores differ" % (trial, int((so != sd).sum()), len(so))
        want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
        assert_pass_equal(r, pose, cov, want)
        assert m.last_detail.n_avg == want["n_avg"]
        dg.close()


def test_rasterizer_matches_oracle(ctx, oracle, rng):
    for sigma, res in [(0.15, 0.05), (0.03, 0.025), (0.24, 0.08), (0.4, 0.1), (0.03, 0.01)]:
        sc = random_scenario(rng, n_points=300, size=260, res=res, n_base=5)
        sc.grid.sigma = sigma
        want = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        dg = device_grid(ctx, sc)
        assert dg.is_fixed_point()
        assert np.array_equal(dg.download(), want)
        dg.close()
    # points outside the grid / on the border band are skipped like the reference does
    sc = random_scenario(rng, n_points=200, size=60, spread=80.0)
    want = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    dg = device_grid(ctx, sc)
    assert np.array_equal(dg.download(), want)
    dg.close()
    # a kernel value that is not a multiple of 2^-25 switches the grid to float32 cells
    sc = random_scenario(rng, n_points=200, size=160)
    sc.grid.occu_offset = 0.2
    sc.grid.default_prob = 0.1
    want = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    dg = device_grid(ctx, sc)
    assert not dg.is_fixed_point()
    assert np.array_equal(dg.download(), want)
    dg.close()


def test_float32_grid_path(ctx, oracle, rng):
    """Arbitrary float cells (not 2^-25 multiples): float gather + FP64 accumulate in beam order."""
    m = matcher.BasedCorrelationScanMatch(ctx)
    for trial in range(4):
        sc = random_scenario(rng, n_points=257, size=200)
        grid = rng.random((sc.grid.size_y, sc.grid.size_x)).astype(np.float32) * np.float32(0.97)
        dg = matcher.ScanMatchMap.from_spec(ctx, sc.grid)
        dg.upload(grid)
        assert not dg.is_fixed_point()
        assert np.array_equal(dg.download(), grid)
        p = synth.pass_param(0.6, 0.05, 0.3, 0.02, 0.3, [100000, 50][trial % 2], True, trial % 3)
        so = oracle.scores(grid, sc.grid, sc.scan_pts, p, oracle.world_to_map(sc.grid, sc.seed_pose))
        sd = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
        assert np.array_equal(so, sd)
        want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, sc.scan_pts, p, pose, cov), pose, cov, want)
        dg.close()


def test_exact_ties_follow_the_reference_sort(ctx, oracle):
    """Binary grid without the centre penalty: thousands of exact ties, so the averaged pose and the
    covariance depend on libstdc++'s unstable sort order; the CUDA path must detect that and
    reproduce it (exact_sort_used) -- still bit-exact."""
    sc, z = load_golden("ties_icra")
    dg = device_grid(ctx, sc)
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    m = matcher.BasedCorrelationScanMatch(ctx)
    used = 0
    pose_in = sc.seed_pose.copy()
    for p in sc.passes:
        want = oracle.match(grid, sc.grid, sc.scan_pts, p, pose_in)
        pose, cov = pose_in.copy(), np.eye(3)
        r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
        assert_pass_equal(r, pose, cov, want)
        assert np.array_equal(cov, want["cov"])
        used += m.last_detail.exact_sort_used
        pose_in = want["pose"]
    assert used >= 1
    # fully degenerate: scan in unknown space, no penalty -> all candidates tie
    flat = np.full((sc.grid.size_y, sc.grid.size_x), np.float32(0.3))
    dg.upload(flat)
    p = sc.passes[0]
    want = oracle.match(flat, sc.grid, sc.scan_pts, p, sc.seed_pose)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
    assert_pass_equal(r, pose, cov, want)
    assert m.last_detail.n_avg == want["n_avg"] == 3549 and m.last_detail.exact_sort_used == 1
    dg.close()


def test_reference_error_behaviour(ctx, oracle):
    sc, _ = load_golden("cfg1_icra")
    m = matcher.BasedCorrelationScanMatch(ctx)
    dg = matcher.ScanMatchMap.from_spec(ctx, sc.grid)
    # map not initialised: response 0, outputs untouched (correlate_scan_matcher.h:792-795)
    pose, cov = sc.seed_pose.copy(), np.arange(9.0).reshape(3, 3)
    assert m.ScanMatch(dg, sc.scan_pts, sc.passes[0], pose, cov) == 0.0
    assert np.array_equal(pose, sc.seed_pose) and np.array_equal(cov, np.arange(9.0).reshape(3, 3))
    dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses)
    # empty scan: same
    assert m.ScanMatch(dg, np.zeros((0, 2)), sc.passes[0], pose, cov) == 0.0
    assert np.array_equal(pose, sc.seed_pose) and np.array_equal(cov, np.arange(9.0).reshape(3, 3))
    # response below threshold: covariance written, pose untouched
    p = sc.passes[0].copy()
    p[4] = 0.99
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
    assert_pass_equal(r, pose, cov, want)
    assert np.array_equal(pose, sc.seed_pose) and m.last_detail.pose_updated == 0
    # window leaves the grid: the reference would read out of bounds; the device path reports it
    far = sc.seed_pose + np.array([11.5, 0.0, 0.0])
    with pytest.raises(matcher.RsmError) as e:
        m.ScanMatch(dg, sc.scan_pts, sc.passes[0], far.copy(), np.eye(3))
    assert e.value.status == 4
    # FAST (branch and bound) is not provided
    p = sc.passes[0].copy()
    p[7] = 3
    with pytest.raises(matcher.RsmError) as e:
        m.ScanMatch(dg, sc.scan_pts, p, sc.seed_pose.copy(), np.eye(3))
    assert e.value.status == 5
    dg.close()


def test_chain_batch_and_loop_closure(ctx, oracle):
    pairs = synth.config4(12, seed=4321)
    want = []
    for sc in pairs:
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        want.append(oracle.match_chain(grid, sc.grid, sc.scan_pts, sc.passes, sc.seed_pose))
    # (1) one by one through ScanMatchers
    sm = matcher.ScanMatchers(ctx, pairs[0].passes)
    grids = [device_grid(ctx, sc) for sc in pairs]
    for sc, dg, w in zip(pairs, grids, want):
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        s = sm.ScanMatch(sc.scan_pts, dg, pose, cov)
        assert s == w["score"] and np.array_equal(pose, w["pose"]) and cov_close(cov, w["cov"])
        assert np.array_equal(sm.last_responses, w["responses"])
    # (2) batched over existing grids
    scores, poses, covs, resp = sm.ScanMatchBatch(grids, [sc.scan_pts for sc in pairs], [sc.seed_pose for sc in pairs])
    for i, w in enumerate(want):
        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
        assert np.array_equal(resp[i], w["responses"])
    # (3) batched with device-side grid construction
    packed = matcher.pack_loop_closure(pairs)
    scores, poses, covs, resp = matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
    for i, w in enumerate(want):
        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
        assert np.array_equal(resp[i], w["responses"])
    # coarse only (use_fine_scan_match = false)
    sc, dg = pairs[0], grids[0]
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    w = oracle.match_chain(grid, sc.grid, sc.scan_pts, sc.passes, sc.seed_pose, use_fine=False)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    s = sm.ScanMatch(sc.scan_pts, dg, pose, cov, use_fine_scan_match=False)
    assert s == w["score"] and np.array_equal(pose, w["pose"]) and cov_close(cov, w["cov"])
    for dg in grids:
        dg.close()


def test_config2_full_size(ctx, oracle):
    """BASELINE configs[1] at full size (8.4e8 evaluations): every score, pose, covariance."""
    sc = synth.config2()
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    dg = device_grid(ctx, sc)
    assert np.array_equal(dg.download(), grid)
    m = matcher.BasedCorrelationScanMatch(ctx)
    p = sc.passes[0]
    so = oracle.scores(grid, sc.grid, sc.scan_pts, p, oracle.world_to_map(sc.grid, sc.seed_pose))
    sd = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    assert np.array_equal(so, sd)
    want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    assert_pass_equal(m.ScanMatch(dg, sc.scan_pts, p, pose, cov), pose, cov, want)
    dg.close()


def test_config5_quarter_scale_vs_oracle(ctx, oracle):
    """BASELINE configs[4] at 1/4 scale (+-2 m / +-45 deg on the full willow map: 81 x 81 x 181 candidates x 938
    beams = 1.1e9 evaluations, multi-beam-round staged launches): every score, the pose and the covariance
    against the oracle (0.5 s on 8 host threads, 3.5 s on one), unsliced and angle-sliced."""
    sc = synth.config5(scale=0.25)
    g, p = sc.grid, sc.passes[0]
    grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
    dg = device_grid(ctx, sc)
    assert np.array_equal(dg.download(), grid)
    m = matcher.BasedCorrelationScanMatch(ctx)
    want_scores = oracle.scores_threaded(grid, g, sc.scan_pts, p, oracle.world_to_map(g, sc.seed_pose))
    got = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    assert np.array_equal(got, want_scores)
    want = oracle.finish_scores(want_scores, g, len(sc.scan_pts), p, sc.seed_pose)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    assert_pass_equal(m.ScanMatch(dg, sc.scan_pts, p, pose, cov), pose, cov, want)
    assert m.last_detail.n_avg == want["n_avg"]
    # the same window through the sliced calls, 3 emulated ranks on this context's GPU, one after the other
    gathered = sliced_on_one_gpu(sc, p, sc.seed_pose, 3)
    for r, pose_s, cov_s, navg in gathered:
        assert r == want["response"] and np.array_equal(pose_s, want["pose"]) and cov_close(cov_s, want["cov"]) and navg == want["n_avg"]
    dg.close()


def sliced_on_one_gpu(sc, p, pose_in, world):
    """SlicedScanMatch with `world` emulated ranks (one context each, lock-step threads, list all-gather)."""
    import threading
    ctxs = [matcher.Context(0) for _ in range(world)]
    grids = []
    for c in ctxs:
        dg = matcher.ScanMatchMap.from_spec(c, sc.grid)
        dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, sc.grid.default_prob, sc.grid.sigma, sc.grid.occu_offset, sc.grid.use_blur)
        grids.append(dg)
    slots, barrier, out = [None] * world, threading.Barrier(world), [None] * world

    def make_gather(rank):
        def gather(buf):
            slots[rank] = buf.copy()
            barrier.wait()
            res = [b.copy() for b in slots]
            barrier.wait()
            return res
        return gather

    def worker(rank):
        sm = matcher.SlicedScanMatch(ctxs[rank], rank, world, make_gather(rank))
        pose, cov = pose_in.copy(), np.eye(3)
        try:
            r = sm.ScanMatch(grids[rank], sc.scan_pts, p, pose, cov)
            out[rank] = (r, pose, cov, sm.last_detail.n_avg)
        except Exception as e:
            out[rank] = e
            barrier.abort()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for dg in grids:
        dg.close()
    for c in ctxs:
        c.close()
    for o in out:
        if isinstance(o, Exception):
            raise o
    return out


def test_config5_full_size_golden(ctx):
    """BASELINE configs[4] at FULL size (321 x 321 x 721 = 74.3 M candidates, 7.0e10 evaluations, 594 MB of scores,
    16 partial tiles per angle) against the fixture written by tests/golden/make_config5.py from the reference's own
    code: every score (sha256 of the array + one checksum per search angle), the head of the sorted list, response,
    pose, covariance, averaging-set size -- unsliced, and angle-sliced over 4 emulated ranks."""
    import hashlib
    from helpers import load_config5_golden
    sc, z = load_config5_golden()
    p = sc.passes[0]
    dg = device_grid(ctx, sc)
    assert hashlib.sha256(np.ascontiguousarray(dg.download()).tobytes()).hexdigest() == str(z["grid_sha"])
    assert np.array_equal(dg.GetMapCoordsPose(sc.seed_pose), z["centre_map"])
    m = matcher.BasedCorrelationScanMatch(ctx)
    scores = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    n_ang = int(z["n_ang"])
    assert scores.size == n_ang * int(z["n_xy"]) ** 2
    sums = scores.view(np.uint64).reshape(n_ang, -1).sum(axis=1, dtype=np.uint64)
    bad = np.flatnonzero(sums != z["angle_sums"])
    assert bad.size == 0, "scores differ from the reference at search angles %s" % bad[:8]
    assert hashlib.sha256(scores.tobytes()).hexdigest() == str(z["scores_sha"])
    head = np.partition(scores, scores.size - 64)[-64:]
    assert np.array_equal(np.sort(head)[::-1], z["sorted_head"])
    del scores, head
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
    assert r == float(z["response"]) and np.array_equal(pose, z["pose"]) and cov_close(cov, z["cov"])
    assert m.last_detail.n_avg == int(z["n_avg"])
    assert np.array_equal([m.last_detail.best_pose_map[i] for i in range(3)] + [m.last_detail.best_score], z["best_map"])
    dg.close()
    for r, pose_s, cov_s, navg in sliced_on_one_gpu(sc, p, sc.seed_pose, 4):
        assert r == float(z["response"]) and np.array_equal(pose_s, z["pose"]) and cov_close(cov_s, z["cov"]) and navg == int(z["n_avg"])


def test_size_independent_properties_wide_window(ctx):
    """1/4-scale config 5 (1.1e9 evaluations): properties that hold at any size, beside the oracle comparison of
    test_config5_quarter_scale_vs_oracle."""
    sc = synth.config5(scale=0.25)
    dg = device_grid(ctx, sc)
    m = matcher.BasedCorrelationScanMatch(ctx)
    p = sc.passes[0]
    full = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    # (1) angle slices reproduce the full pass bit for bit (what multi-GPU angle slicing relies on)
    n_ang = int(np.floor(p[2] * 2 / p[3]) + 1)
    n_xy = int(np.floor(p[0] / p[1] + 0.5) + 1)
    assert full.size == n_ang * n_xy * n_xy
    cuts = [0, 1, n_ang // 3, n_ang // 2 + 1, n_ang]
    parts = [m.scores(dg, sc.scan_pts, p, sc.seed_pose, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), full)
    # (2) the float32 kernel and the fixed-point kernel are two independent code paths that must
    #     agree exactly (float sums of 2^-25 multiples are exact in FP64)
    cells = dg.download()
    bumped = cells.copy()
    bumped[0, 0] = np.float32(0.1)   # one unreachable, non-representable cell forces float32 cells
    df = matcher.ScanMatchMap.from_spec(ctx, sc.grid)
    df.upload(bumped)
    assert dg.is_fixed_point() and not df.is_fixed_point()
    assert np.array_equal(m.scores(df, sc.scan_pts, p, sc.seed_pose), full)
    # (3) the matcher's winner is the maximum of the dump, and the unpenalised score is a mean of
    #     grid values, hence bounded by the grid's range
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
    assert m.last_detail.best_score == full.max() and r == min(full.max(), 1.0)
    p2 = p.copy()
    p2[6] = 0.0
    raw = m.scores(dg, sc.scan_pts, p2, sc.seed_pose, 0, 8)
    assert raw.min() >= cells.min() - 1e-12 and raw.max() <= cells.max() + 1e-12
    # (4) idempotence: same call, same bits
    assert np.array_equal(m.scores(dg, sc.scan_pts, p, sc.seed_pose, 5, 9), full[5 * n_xy * n_xy: 9 * n_xy * n_xy])
    dg.close()
    df.close()


def test_angle_sliced_match_equals_unsliced(oracle):
    """one window cut along the angle index over N ranks, merged through the partial /
    merge / finish calls, must equal the unsliced match (and hence the reference) -- also when exact
    ties send every rank through rsm_match_slice_scores / rsm_match_finish_exact.  The ranks are
    emulated by N contexts on this GPU, run in lock step; the all-gather is a list."""
    import threading
    from roborts_edu_slam_b200.sharding import contiguous_range

    def run_sliced(sc, p, pose_in, world):
        ctxs = [matcher.Context(0) for _ in range(world)]
        grids = []
        for c in ctxs:
            dg = matcher.ScanMatchMap.from_spec(c, sc.grid)
            dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, sc.grid.default_prob, sc.grid.sigma, sc.grid.occu_offset)
            grids.append(dg)
        # a barrier-style all-gather between threads
        slots, barrier, out = [None] * world, threading.Barrier(world), [None] * world

        def make_gather(rank):
            def gather(buf):
                slots[rank] = buf.copy()
                barrier.wait()
                res = [b.copy() for b in slots]
                barrier.wait()
                return res
            return gather

        def worker(rank):
            sm = matcher.SlicedScanMatch(ctxs[rank], rank, world, make_gather(rank))
            pose, cov = pose_in.copy(), np.eye(3)
            try:
                r = sm.ScanMatch(grids[rank], sc.scan_pts, p, pose, cov)
                out[rank] = (r, pose, cov, sm.last_detail.n_avg, sm.exact_fallback, sm.last_detail.exact_sort_used)
            except Exception as e:      # keep the other ranks' barrier from hanging the test
                out[rank] = e
                barrier.abort()

        ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for dg in grids:
            dg.close()
        for c in ctxs:
            c.close()
        for o in out:
            if isinstance(o, Exception):
                raise o
        return out

    # ties_icra: thousands of exact ties -> the sliced path gathers the slices and runs the reference's sort itself
    cases = [(synth.config1(), 3), (synth.config3(True), 2), (synth.config4(1, seed=5)[0], 4), (load_golden("ties_icra")[0], 3)]
    fallbacks = 0
    for sc, world in cases:
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        pose_in = sc.seed_pose.copy()
        for p in sc.passes:
            want = oracle.match(grid, sc.grid, sc.scan_pts, p, pose_in)
            res = run_sliced(sc, p, pose_in, world)
            assert all(r is not None for r in res)
            assert len({r[4] for r in res}) == 1, "the ranks must agree on the exact-tie fallback"
            for r, pose, cov, navg, fell_back, exact_used in res:     # every rank holds the same, reference-identical result
                assert r == want["response"] and np.array_equal(pose, want["pose"]) and cov_close(cov, want["cov"])
                assert navg == want["n_avg"]
                assert bool(exact_used) == bool(fell_back)
                if fell_back:
                    assert np.array_equal(cov, want["cov"])       # the same std::sort: bit-equal
            fallbacks += int(res[0][4])
            pose_in = want["pose"]
    assert fallbacks >= 1, "ties_icra must exercise the gathered exact path"
    # more ranks than angles: empty slices are fine
    sc = synth.config1()
    p = synth.pass_param(0.2, 0.05, 0.03, 0.0349, 0.3, 100000, True, 0)   # n_ang = 2
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
    for r, pose, cov, navg, _, _ in run_sliced(sc, p, sc.seed_pose, 4):
        assert r == want["response"] and np.array_equal(pose, want["pose"]) and cov_close(cov, want["cov"])


def test_staged_shared_memory_variant(ctx, monkeypatch):
    """The TMA-staged variant (default for wide unit-step windows) and the L1 variant
    (RSM_NO_STAGED=1) must agree bit for bit."""
    sc = synth.config5(scale=0.25)
    dg = device_grid(ctx, sc)
    m = matcher.BasedCorrelationScanMatch(ctx)
    p = sc.passes[0]
    got = m.scores(dg, sc.scan_pts, p, sc.seed_pose, 3, 40)
    pose, cov = sc.seed_pose.copy(), np.eye(3)
    r1 = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
    monkeypatch.setenv("RSM_NO_STAGED", "1")
    ref = m.scores(dg, sc.scan_pts, p, sc.seed_pose, 3, 40)
    assert np.array_equal(got, ref)
    pose2, cov2 = sc.seed_pose.copy(), np.eye(3)
    r2 = m.ScanMatch(dg, sc.scan_pts, p, pose2, cov2)
    assert r1 == r2 and np.array_equal(pose, pose2) and np.array_equal(cov, cov2)
    # a window that touches the grid border: beams fall back to exact per-thread indices
    edge = sc.seed_pose.copy()
    g = sc.grid
    edge[:2] = [-(g.off_x) + 12.5, -(g.off_y) + 12.5]       # 12.5 m from the grid corner
    ref = m.scores(dg, sc.scan_pts[:50] * 0.2, p, edge, 0, 6)
    monkeypatch.delenv("RSM_NO_STAGED")
    got = m.scores(dg, sc.scan_pts[:50] * 0.2, p, edge, 0, 6)
    assert np.array_equal(got, ref)
    dg.close()


@pytest.mark.parametrize("n_xy,beams", [(49, 90), (61, 700), (64, 300), (65, 130), (81, 1000), (96, 64), (97, 520), (131, 260)])
def test_staged_tiles_and_clusters(ctx, oracle, rng, n_xy, beams):
    """Staged variant against the oracle over window sizes that hit both tile shapes, partial
    tiles / rows, and beam counts that give cluster sizes from 1 to 8."""
    g = synth.GridSpec(0.05, 0.15, 700, 600, 3.0, 2.0)
    grid = synth.random_grid(rng, g.size_x, g.size_y)
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    dg.upload(grid)
    m = matcher.BasedCorrelationScanMatch(ctx)
    ang = np.sort(rng.uniform(-np.pi, np.pi, beams))
    rad = rng.uniform(1.0, 7.0, beams)
    pts = np.stack([np.cos(ang) * rad, np.sin(ang) * rad], axis=1)
    pose = np.array([700 * 0.05 / 2 - 3.0 + 0.013, 600 * 0.05 / 2 - 2.0 - 0.021, 0.3])
    p = synth.pass_param((n_xy - 1) * 0.05, 0.05, 0.06, 0.02, 0.3, 100000, True, 0)      # 6 or 7 angles
    so = oracle.scores(grid, g, pts, p, oracle.world_to_map(g, pose))
    assert so.size % (n_xy * n_xy) == 0 and so.size // (n_xy * n_xy) in (6, 7)
    sd = m.scores(dg, pts, p, pose)
    assert np.array_equal(so, sd)
    want = oracle.match(grid, g, pts, p, pose)
    pz, cz = pose.copy(), np.eye(3)
    assert_pass_equal(m.ScanMatch(dg, pts, p, pz, cz), pz, cz, want)
    dg.close()


@pytest.mark.parametrize("n_xy,beams,variant,ctas", [
    (81, 1000, 0, 148), (81, 1000, 0, 5), (81, 707, 2, 148), (65, 130, 0, 37), (97, 520, 0, 148), (131, 260, 0, 148),
    (131, 260, 2, 23), (49, 90, 1, 148), (61, 700, 1, 148), (64, 300, 0, 148), (80, 333, 0, 148), (78, 200, 0, 11),
    (33, 150, 0, 148), (163, 100, 0, 148), (163, 100, 1, 64)])
def test_stream_plan_of_the_staged_kernel(ctx, oracle, rng, monkeypatch, n_xy, beams, variant, ctas):
    """Stream plan (persistent CTAs over one beam sequence, items finished by the last CTA to arrive) with each of its
    three tile mappings, forced on small windows: every score against the oracle.  Partial tiles on both sides, shares
    that cut an item into many parts, more CTAs than items and fewer."""
    monkeypatch.setenv("RSM_STAGED_PLAN", "stream")
    monkeypatch.setenv("RSM_STREAM_VARIANT", str(variant))
    monkeypatch.setenv("RSM_STREAM_CTAS", str(ctas))
    g = synth.GridSpec(0.05, 0.15, 700, 600, 3.0, 2.0)
    grid = synth.random_grid(rng, g.size_x, g.size_y)
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    dg.upload(grid)
    m = matcher.BasedCorrelationScanMatch(ctx)
    ang = np.sort(rng.uniform(-np.pi, np.pi, beams))
    rad = rng.uniform(1.0, 7.0, beams)
    pts = np.stack([np.cos(ang) * rad, np.sin(ang) * rad], axis=1)
    pose = np.array([700 * 0.05 / 2 - 3.0 + 0.013, 600 * 0.05 / 2 - 2.0 - 0.021, 0.3])
    p = synth.pass_param((n_xy - 1) * 0.05, 0.05, 0.06, 0.02, 0.3, 100000, True, 0)      # 6 or 7 angles
    if n_xy < 48:
        pytest.skip("staged variants start at 48 translations per axis")
    so = oracle.scores(grid, g, pts, p, oracle.world_to_map(g, pose))
    before = ctx.stats()["kernel_launches"]
    sd = m.scores(dg, pts, p, pose)
    assert ctx.stats()["kernel_launches"] - before == 1, "one persistent launch"
    assert np.array_equal(so, sd)
    want = oracle.match(grid, g, pts, p, pose)
    for _ in range(3):          # the third call replays the pass as a CUDA graph: tickets must be zeroed again
        pz, cz = pose.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, pts, p, pz, cz), pz, cz, want)
    # a window at the grid border: beams whose footprint leaves the grid take exact indices from global memory
    half = min(n_xy - 1, 80) // 2
    edge = np.array([-g.off_x + (half + 15) * 0.05 + 0.02, -g.off_y + (half + 15) * 0.05 + 0.03, -0.2])
    q = synth.pass_param(2 * half * 0.05, 0.05, 0.04, 0.02, 0.3, 100000, True, 0)
    short = pts[:60] * 0.1        # at most 14 cells long: the lowest cell touched is 0 or 1
    so = oracle.scores(grid, g, short, q, oracle.world_to_map(g, edge))
    assert np.array_equal(so, m.scores(dg, short, q, edge))
    dg.close()


@pytest.mark.parametrize("variant,ctas", [(0, 148), (0, 9), (2, 31), (1, 148)])
def test_stream_plan_over_a_batch_of_jobs(ctx, oracle, rng, monkeypatch, variant, ctas):
    """One persistent launch over SEVERAL jobs (rsm_match_batch): different grids, scans of different sizes, shares that
    run across job boundaries (the CTA reloads the job, its tensor maps and its beam table mid-range)."""
    monkeypatch.setenv("RSM_STAGED_PLAN", "stream")
    monkeypatch.setenv("RSM_STREAM_VARIANT", str(variant))
    monkeypatch.setenv("RSM_STREAM_CTAS", str(ctas))
    n_xy = 65 if variant != 1 else 57
    coarse = synth.pass_param((n_xy - 1) * 0.05, 0.05, 0.08, 0.02, 0.3, 100000, True, 0)       # 9 angles
    sm = matcher.ScanMatchers(ctx, [coarse, coarse, coarse])
    grids, dgs, scans, poses = [], [], [], []
    for k, beams in enumerate((70, 333, 64, 1000, 129)):
        g = synth.GridSpec(0.05, 0.15, 420 + 40 * k, 400, 3.0 + k, 2.0)
        grid = synth.random_grid(rng, g.size_x, g.size_y)
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.upload(grid)
        ang = np.sort(rng.uniform(-np.pi, np.pi, beams))
        rad = rng.uniform(1.0, 5.5, beams)
        grids.append((g, grid)); dgs.append(dg)
        scans.append(np.stack([np.cos(ang) * rad, np.sin(ang) * rad], axis=1))
        poses.append(np.array([g.size_x * 0.05 / 2 - g.off_x + 0.011 * k, 400 * 0.05 / 2 - 2.0 - 0.007 * k, 0.2 + 0.1 * k]))
    before = ctx.stats()["score_launches"]
    scores, out_poses, covs, resp = sm.ScanMatchBatch(dgs, scans, np.array(poses), use_fine_scan_match=False)
    assert ctx.stats()["score_launches"] - before == 1, "one pass, one launch for the whole batch"
    for k, (g, grid) in enumerate(grids):
        want = oracle.match(grid, g, scans[k], coarse, poses[k])
        assert resp[k, 0] == want["response"] and np.array_equal(out_poses[k], want["pose"]) and cov_close(covs[k], want["cov"])
    for dg in dgs:
        dg.close()


def test_stream_plan_is_the_default_for_full_waves(ctx, oracle, monkeypatch):
    """BASELINE configs[1] (181 items) takes the stream plan by default and equals the cluster plan bit for bit; a
    batch of different windows shares one persistent launch."""
    sc = synth.config2()
    dg = device_grid(ctx, sc)
    m = matcher.BasedCorrelationScanMatch(ctx)
    p = sc.passes[0]
    before = ctx.stats()["kernel_launches"]
    got = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    assert ctx.stats()["kernel_launches"] - before == 1
    monkeypatch.setenv("RSM_STAGED_PLAN", "cluster")
    before = ctx.stats()["kernel_launches"]
    ref = m.scores(dg, sc.scan_pts, p, sc.seed_pose)
    assert ctx.stats()["kernel_launches"] - before == 2
    assert np.array_equal(got, ref)
    dg.close()


def test_wide_grid_runtime_pitch(ctx, oracle, rng):
    """Grids wider than the largest padded pitch (4128 cells) use the run-time stride variants."""
    g = synth.GridSpec(0.05, 0.15, 4300, 300, 1.0, 1.0)
    grid = synth.random_grid(rng, g.size_x, g.size_y)
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    dg.upload(grid)
    assert np.array_equal(dg.download(), grid)
    m = matcher.BasedCorrelationScanMatch(ctx)
    ang = np.sort(rng.uniform(-np.pi, np.pi, 300))
    pts = np.stack([np.cos(ang) * rng.uniform(20, 100, 300), np.sin(ang) * rng.uniform(20, 100, 300)], axis=1)
    pose = np.array([4300 * 0.05 - 12.0, 6.5, 0.4])       # near the right edge, past column 4128
    for p in (synth.pass_param(1.0, 0.05, 0.2, 0.02, 0.3, 100000, True, 0),
              synth.pass_param(0.4, 0.02, 0.1, 0.02, 0.3, 100000, True, 1)):
        so = oracle.scores(grid, g, pts, p, oracle.world_to_map(g, pose))
        sd = m.scores(dg, pts, p, pose)
        assert np.array_equal(so, sd)
        want = oracle.match(grid, g, pts, p, pose)
        pz, cz = pose.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, pts, p, pz, cz), pz, cz, want)
    dg.close()


def test_second_device_from_a_fresh_thread(oracle):
    """A context on device 1 used from a new host thread (whose current device is 0): every
    entry point has to select the context's device itself.  Needs two GPUs."""
    import threading
    try:
        c1 = matcher.Context(1)
    except matcher.RsmError:
        pytest.skip("one GPU only")
    sc = synth.config1()
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    want = oracle.match(grid, sc.grid, sc.scan_pts, sc.passes[0], sc.seed_pose)
    out = {}

    def work():
        dg = matcher.ScanMatchMap.from_spec(c1, sc.grid)
        dg.upload(grid)
        m = matcher.BasedCorrelationScanMatch(c1)
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        out["r"] = m.ScanMatch(dg, sc.scan_pts, sc.passes[0], pose, cov)
        out["pose"] = pose
        dg.close()

    t = threading.Thread(target=work)
    t.start()
    t.join()
    c1.close()
    assert out["r"] == want["response"] and np.array_equal(out["pose"], want["pose"])


# ---- map check ---------------------------------------------------------------
def test_mapcheck_golden_and_oracle(ctx, oracle, rng):
    """rsm_map_check_penalize against the reference's coefficients (fixtures) and, on fresh random
    poses / knobs / per-pose scans, against the oracle: bit-equal."""
    from helpers import load_mapcheck, mapcheck_names
    for name in mapcheck_names():
        g, occ, z = load_mapcheck(name)
        pm = matcher.ScanMatchMap.from_spec(ctx, g)
        pm.upload_occupancy(occ)
        for k, ps in enumerate(z["param_sets"]):
            got = pm.MapCheckPenalize(z["scan_pts"], z["poses"], int(ps[0]), ps[1], ps[2], bool(ps[3]), ps[4:6])
            assert np.array_equal(got, z["coeff"][k]), (name, k, got, z["coeff"][k])
        # one scan per pose (different lengths, one empty), random knobs
        poses = z["poses"][:24] + rng.uniform(-0.2, 0.2, (24, 3))
        scans = [z["scan_pts"][: int(rng.integers(0, len(z["scan_pts"]) + 1))] for _ in range(24)]
        scans[5] = z["scan_pts"][:0]
        for cp, tol, gain, logistic in ((100, 2.5, 0.015, True), (13, 0.5, 0.2, False)):
            got = pm.MapCheckPenalize(scans, poses, cp, tol, gain, logistic)
            want = np.array([oracle.map_check_penalize(occ, g, s, p, cp, tol, gain, logistic) for s, p in zip(scans, poses)])
            assert np.array_equal(got, want)
        pm.close()


def test_mapcheck_errors(ctx):
    g = synth.GridSpec(0.05, 0.0, 64, 48, 1.0, 1.0, 0.5, 0.88, False)
    pm = matcher.ScanMatchMap.from_spec(ctx, g)
    with pytest.raises(matcher.RsmError) as e:
        pm.MapCheckPenalize(np.zeros((4, 2)), np.zeros((1, 3)))
    assert "RSM_ERR_NOT_INIT" in str(e.value)
    pm.upload_occupancy(np.zeros((48, 64), dtype=np.uint8))
    assert pm.MapCheckPenalize(np.ones((4, 2)), np.zeros((0, 3))).shape == (0,)
    with pytest.raises(matcher.RsmError) as e:     # the reference divides by zero for check_point_num = 1
        pm.MapCheckPenalize(np.ones((4, 2)), np.zeros((1, 3)), 1, 2.5, 0.015)
    assert "RSM_ERR_INVALID" in str(e.value)
    # an empty map blocks nothing: 1 + 2 * gain
    assert np.array_equal(pm.MapCheckPenalize(np.array([[5.0, 1.0], [0.0, 7.0]]), np.array([[0.5, 0.4, 0.1]]), 100, 2.5, 0.015),
                          np.array([1.0 + 2 * 0.015]))
    pm.close()


# ---- scan store + batched ScanMatchInterface ---------------------------------
def test_scan_store_interface_batch(ctx, oracle):
    """rsm_scan_match_interface_batch: chains and match scans named by id in a device-resident store,
    back-end grid reset + chain + map check per candidate, against the oracle run candidate by candidate
    (slam_processor.cpp:250-326); pose updates of stored scans (UpdateRangeData) are followed."""
    from helpers import load_mapcheck
    pairs = synth.config4(4)                       # pair 0 = the scenario the mapcheck_pair0 fixture was built on
    store, pub_store = matcher.ScanStore(ctx), matcher.ScanStore(ctx)
    chains, match_ids = [], []
    for sc in pairs:
        ids = [store.AddRangeData(p, pose) for p, pose in zip(sc.base_pts, sc.base_poses)]
        mid = store.AddRangeData(sc.scan_pts, sc.seed_pose)
        for p, pose in zip(sc.base_pts, sc.base_poses):
            pub_store.AddRangeData(p, pose)
        assert pub_store.AddRangeData(sc.scan_pts, sc.seed_pose) == mid
        chains.append(ids)
        match_ids.append(mid)
    assert len(store) == len(pub_store) == 4 * 9 and np.array_equal(store.sensor_pose(chains[1][2]), pairs[1].base_poses[2])
    # candidates: every pair against its own chain, pair 0's scan against a sub-chain and against a chain
    # mixing scans of two pairs, one candidate with an empty chain (grid not initialised -> 0, pose untouched)
    cands = [(i, chains[i], pairs[i].grid_centre, pairs[i].seed_pose) for i in range(4)]
    cands.append((0, chains[0][2:6], pairs[0].grid_centre, pairs[0].seed_pose + np.array([0.03, 0.02, -0.04])))
    cands.append((0, chains[0][:4] + chains[1][:2], pairs[0].grid_centre, pairs[0].seed_pose))
    cands.append((1, [], pairs[1].grid_centre, pairs[1].seed_pose))
    g0 = pairs[0].grid
    gpub, occ, _ = load_mapcheck("mapcheck_pair0")
    pm = matcher.ScanMatchMap.from_spec(ctx, gpub)
    pm.upload_occupancy(occ)
    check = (100, 2.5, 0.015, True)

    def flat(ids):
        out_pts, out_poses = [], []
        for i in ids:
            p, k = divmod(i, 9)
            out_pts.append(pairs[p].base_pts[k])
            out_poses.append(store.sensor_pose(i))
        return out_pts, np.array(out_poses).reshape(-1, 3)

    def expected(with_check):
        want = []
        for who, ids, centre, seed in cands:
            sc = pairs[who]
            g = synth.backend_grid(g0.res, g0.sigma, 10.0, centre)
            assert (g.size_x, g.off_x) == (g0.size_x, sc.grid.off_x)
            if not ids:
                want.append(dict(score=0.0, pose=np.asarray(seed, dtype=np.float64), responses=np.zeros(3), cov=np.eye(3)))
                continue
            bpts, bposes = flat(ids)
            grid = oracle.build_grid(g, bpts, bposes)
            w = oracle.match_chain(grid, g, sc.scan_pts, sc.passes, seed)
            if with_check:
                c = oracle.map_check_penalize(occ, gpub, sc.scan_pts, w["pose"], *check)
                s = w["score"] * c
                w = dict(w, score=1.0 if s > 1.0 else s)
            want.append(w)
        return want

    def run(with_check):
        return matcher.scan_match_interface_batch(
            ctx, store, g0, [c[2] for c in cands], [c[1] for c in cands], [match_ids[c[0]] for c in cands],
            [c[3] for c in cands], pairs[0].passes, True,
            pm if with_check else None, pub_store if with_check else None, check if with_check else None)

    for with_check in (False, True):
        scores, poses, covs, resp = run(with_check)
        for i, w in enumerate(expected(with_check)):
            assert scores[i] == w["score"], (with_check, i, scores[i], w["score"])
            assert np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
            assert np.array_equal(resp[i], w["responses"])
    assert expected(True)[1]["score"] == 0.0 or expected(True)[1]["score"] < expected(False)[1]["score"]
    # the pose graph moves two stored scans: the next batch rasterises them at their new poses
    store.UpdateRangeData([chains[0][1], chains[0][5]],
                          [pairs[0].base_poses[1] + np.array([0.07, -0.03, 0.02]), pairs[0].base_poses[5] + np.array([-0.05, 0.04, -0.03])])
    scores, poses, covs, resp = run(False)
    for i, w in enumerate(expected(False)):
        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
    # errors: unknown ids
    with pytest.raises(matcher.RsmError) as e:
        matcher.scan_match_interface_batch(ctx, store, g0, [pairs[0].grid_centre], [[0, 999]], [match_ids[0]],
                                           [pairs[0].seed_pose], pairs[0].passes)
    assert "RSM_ERR_INVALID" in str(e.value)
    pm.close()
    store.close()
    pub_store.close()


# ---- Gauss-Newton matcher ----------------------------------------------------
def test_optimize_golden(ctx):
    """rsm_optimize / rsm_optimize_batch / rsm_match_chain_opt against the outputs of the reference's own
    optimize_scan_matcher.h (committed fixture): cost, pose, chain score and responses bit for bit."""
    from helpers import optimize_cases
    z, cases = optimize_cases()
    opt = matcher.BasedOptimizeScanMatch(ctx)
    for tag, sc, gc, base_c, scan_c in cases:
        gf = sc.grid
        dev_f = device_grid(ctx, sc)
        dev_c = matcher.ScanMatchMap.from_spec(ctx, gc)
        dev_c.InitMapWithRangeVec(base_c, sc.base_poses, gc.default_prob, gc.sigma, gc.occu_offset, gc.use_blur)
        seeds = sc.truth_pose + z["seed_deltas"]
        for mi, (dev, pts) in enumerate(((dev_f, sc.scan_pts), (dev_c, scan_c))):
            for oi, op in enumerate(z["opt_sets"]):
                costs, poses, iters = opt.ScanMatchBatch([dev] * len(seeds), [pts] * len(seeds), op, seeds)
                assert np.array_equal(costs, z[tag + "_cost"][mi, oi]), (tag, mi, oi, costs, z[tag + "_cost"][mi, oi])
                assert np.array_equal(poses, z[tag + "_pose"][mi, oi])
                assert iters.min() >= 1 and iters.max() <= int(op[0])
                pose = seeds[0].copy()                                   # one problem through rsm_optimize
                assert opt.ScanMatch(dev, pts, op, pose) == z[tag + "_cost"][mi, oi, 0]
                assert np.array_equal(pose, z[tag + "_pose"][mi, oi, 0]) and opt.last_iterations == iters[0]
        sm = matcher.ScanMatchers(ctx, synth.chain_yaml())
        for fi, fc in enumerate(z["failed_costs"]):
            for ui, use_fine in enumerate((True, False)):
                for si in range(4):
                    pose, cov = seeds[si].copy(), np.eye(3)
                    s = sm.ScanMatchWithOptimize(scan_c, sc.scan_pts, dev_c, dev_f, pose, cov, z["opt_sets"][0], fc, use_fine)
                    assert s == z[tag + "_chain_score"][fi, ui, si], (tag, fi, ui, si)
                    assert np.array_equal(pose, z[tag + "_chain_pose"][fi, ui, si])
                    assert cov_close(cov, z[tag + "_chain_cov"][fi, ui, si])
                    assert sm.last_optimize_cost == z[tag + "_chain_resp"][fi, ui, si, 0]
                    assert np.array_equal(sm.last_responses, z[tag + "_chain_resp"][fi, ui, si, 1:])
        dev_f.close()
        dev_c.close()


def test_optimize_random_and_errors(ctx, oracle, rng):
    """Random problems against the oracle (blur-level and arbitrary float cells, scans that partly leave the
    map, random knobs), a mixed batch, and the reference's error behaviour."""
    opt = matcher.BasedOptimizeScanMatch(ctx)
    devs, scans, seeds, want = [], [], [], []
    op = (10, 0.1, 0.5, 0.5, 0.5)
    for k in range(24):
        sc = random_scenario(rng, n_points=int(rng.integers(1, 700)), size=int(rng.choice([96, 160, 333])))
        g = sc.grid
        grid = oracle.build_grid(g, sc.base_pts, sc.base_poses)
        if k % 3 == 2:
            grid = rng.random(grid.shape).astype(np.float32)
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.upload(grid)
        assert dg.is_fixed_point() == (k % 3 != 2)
        scale = rng.choice([0.02, 0.2, 1.0, 4.0])
        seed = sc.truth_pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)]) * scale
        knobs = (int(rng.integers(1, 12)), float(rng.choice([0.1, 1.0, 1e-6])), float(rng.choice([0.5, 2.0, 0.0])),
                 float(rng.choice([0.5, 0.05])), float(rng.choice([0.5, 0.2, 0.02])))
        w = oracle.optimize(grid, g, sc.scan_pts, knobs, seed)
        pose = seed.copy()
        assert opt.ScanMatch(dg, sc.scan_pts, knobs, pose) == w["cost"], k
        assert np.array_equal(pose, w["pose"]) and opt.last_iterations == min(knobs[0], w["iterations"] + 1)
        devs.append(dg); scans.append(sc.scan_pts); seeds.append(seed)
        want.append(oracle.optimize(grid, g, sc.scan_pts, op, seed))
    costs, poses, iters = opt.ScanMatchBatch(devs, scans, op, np.array(seeds))
    assert np.array_equal(costs, [w["cost"] for w in want]) and np.array_equal(poses, [w["pose"] for w in want])
    # errors: empty scan / uninitialised grid -> kMaxCost, pose untouched (optimize_scan_matcher.h:73-76)
    pose = seeds[0].copy()
    assert opt.ScanMatch(devs[0], np.zeros((0, 2)), op, pose) == 1000.0 and np.array_equal(pose, seeds[0])
    fresh = matcher.ScanMatchMap.from_spec(ctx, synth.GridSpec(0.05, 0.15, 64, 64, 0.0, 0.0))
    assert opt.ScanMatch(fresh, scans[0], op, pose) == 1000.0 and np.array_equal(pose, seeds[0])
    with pytest.raises(matcher.RsmError) as e:
        opt.ScanMatch(devs[0], scans[0], (0, 0.1, 0.5, 0.5, 0.5), pose)
    assert "RSM_ERR_INVALID" in str(e.value)
    fresh.close()
    for dg in devs:
        dg.close()


@pytest.mark.parametrize("plan", ["default", "stream0", "stream1", "stream2"])
def test_cell_boundary_fallback_paths(ctx, oracle, rng, monkeypatch, plan):
    """Beams whose rotated end point sits exactly on a cell boundary for a whole tile fail the kernels'
    provable index test and take the exact per-thread path.  Regression: the tiled kernel used to read
    those beams' end points from a buffer that the prefetch of a later chunk had already overwritten.
    Axis-aligned search angle (index 0 is exactly 0 rad), an integer-aligned centre and end points on a
    half-cell lattice put every beam of angle 0 on that path; other angles run the fast path.
    plan = stream<v>: the staged kernel's stream plan with tile mapping v forced on the wide windows (few CTAs, so
    that rounds mix safe and unsafe beams and items are cut between CTAs)."""
    m = matcher.BasedCorrelationScanMatch(ctx)
    if plan != "default":
        monkeypatch.setenv("RSM_STAGED_PLAN", "stream")
        monkeypatch.setenv("RSM_STREAM_VARIANT", plan[-1])
        monkeypatch.setenv("RSM_STREAM_CTAS", "13")
    for n_pts, half_lattice, window, size in ((200, 1.0, 0.6, 200), (333, 0.2, 0.6, 200), (97, 1.0, 0.25, 200),
                                              (150, 1.0, 2.4, 320), (700, 0.05, 2.4, 320), (260, 1.0, 3.2, 400)):
        g = synth.GridSpec(0.05, 0.15, size, size, 0.0, 0.0)
        levels = np.array([0.3, 0.5642, 0.6666, 0.7046, 0.7875, 0.8324, 1.0], dtype=np.float32)
        grid = levels[rng.integers(0, len(levels), (size, size))]
        pts = rng.uniform(-60, 60, (n_pts, 2))
        snap = rng.random(n_pts) < half_lattice
        pts[snap] = np.round(pts[snap] * 2) / 2            # multiples of half a cell
        seed = np.array([size // 2 * 0.05, size // 2 * 0.05, 0.3])
        p = synth.pass_param(window, 0.05, 0.3, 0.1, 0.3, 100000, True, 0)     # angle index 0: 0.3 - 0.3 = 0 rad exactly
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.upload(grid)
        assert dg.is_fixed_point()
        centre = oracle.world_to_map(g, seed)
        assert centre[0] == size // 2 and centre[2] - 0.3 == 0.0
        so = oracle.scores(grid, g, pts, p, centre)
        sd = m.scores(dg, pts, p, seed)
        assert np.array_equal(so, sd), (n_pts, window, int((so != sd).sum()))
        want = oracle.match(grid, g, pts, p, seed)
        pose, cov = seed.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, pts, p, pose, cov), pose, cov, want)
        dg.close()


def test_fuzz_every_scoring_variant(ctx, oracle):
    """Seeded fuzz over what selects a scoring kernel and its slow paths: window width (flat / tiled /
    staged, one or several tiles, cluster splits), search step (unit, integer, fractional), grid width
    (both padded pitches and the run-time pitch), fixed-point vs float cells, beam subsampling, end points
    snapped to cell boundaries, integer-aligned centres.  Every candidate score bit-equal, then the pass."""
    rng = np.random.default_rng(987654321)
    m = matcher.BasedCorrelationScanMatch(ctx)
    levels = np.array([0.3, 0.5642, 0.6666, 0.7046, 0.7875, 0.8324, 1.0], dtype=np.float32)
    n_done = 0
    for trial in range(90):
        size = int(rng.choice([160, 300, 544, 700, 1200]))
        res = 0.05
        n_xy = int(rng.choice([3, 5, 7, 11, 13, 25, 33, 49, 65, 81, 97]))
        step_cells = float(rng.choice([1.0, 1.0, 1.0, 2.0, 0.4, 0.2, 0.5]))
        if n_xy >= 49 and rng.random() < 0.7:
            step_cells = 1.0                      # mostly the staged kernel for wide windows
        sres = res * step_cells
        window = sres * (n_xy - 1)
        n_ang = int(rng.integers(1, 8))
        ares = float(rng.choice([0.0349, 0.1, 0.01]))
        aoff = ares * (n_ang - 1) / 2 + 1e-9
        P = int(rng.choice([1, 7, 33, 100, 257, 640, 900]))
        if n_xy * n_xy * n_ang * P > 6e7:
            P = max(1, int(6e7 / (n_xy * n_xy * n_ang)))
        half_window_cells = window / res / 2
        reach = size / 2 - half_window_cells - 6
        if reach < 8:
            continue
        pts = rng.uniform(-1, 1, (P, 2))
        pts = pts / np.maximum(1.0, np.linalg.norm(pts, axis=1, keepdims=True)) * reach
        snap = rng.random(P) < rng.choice([0.0, 0.1, 1.0])
        pts[snap] = np.round(pts[snap] * 2) / 2
        g = synth.GridSpec(res, 0.15, size, size, float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        fixed = rng.random() < 0.75
        grid = levels[rng.integers(0, len(levels), (size, size))] if fixed else rng.random((size, size)).astype(np.float32)
        centre_cells = np.array([size / 2, size / 2]) + (rng.integers(-3, 4, 2) if rng.random() < 0.5 else rng.uniform(-3, 3, 2))
        theta = float(rng.choice([0.0, aoff, rng.uniform(-3, 3)]))
        seed = np.array([centre_cells[0] * res - g.off_x, centre_cells[1] * res - g.off_y, theta])
        use_pts = int(rng.choice([100000, 100000, 50, 7]))
        p = synth.pass_param(window, sres, aoff, ares, 0.3, use_pts, bool(rng.integers(0, 2)), int(rng.integers(0, 3)))
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.upload(grid)
        centre = oracle.world_to_map(g, seed)
        geo = oracle.geometry(g, p, P, centre)
        so = oracle.scores(grid, g, pts, p, centre)
        sd = m.scores(dg, pts, p, seed)
        assert np.array_equal(so, sd), (trial, size, n_xy, step_cells, n_ang, P, fixed, use_pts, int((so != sd).sum()))
        want = oracle.match(grid, g, pts, p, seed)
        pose, cov = seed.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, pts, p, pose, cov), pose, cov, want)
        assert (geo["n_xy"], geo["n_ang"]) == (m.last_detail.n_xy, m.last_detail.n_ang)
        dg.close()
        n_done += 1
    assert n_done >= 70


# ---- front-end map maintenance -----------------------------------------------
def test_frontend_map_stays_on_device(ctx):
    """A scan-match map kept on the device across 44 scans of a trajectory that leaves the initial extent on
    every side: rsm_grid_fill / rsm_grid_update_by_range / rsm_grid_extend against the cells the reference's own
    UpdateMapByRange + ExtendSize produced after every step (fixture), then a match on the grown map."""
    import hashlib
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_frontend", os.path.join(os.path.dirname(__file__), "golden", "make_frontend.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)              # only its pure helpers are used (trajectory, scans, spec); Ref is not touched
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend_willow.npz"), allow_pickle=False)
    g = mf.spec()
    poses, pts = mf.trajectory(), mf.scans()
    h = hashlib.sha256()
    for a in pts:
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == str(z["inputs_sha"])
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    dg.fill(0.5, g.default_prob)
    fresh = dg.download()
    assert fresh[0, 0] == np.float32(0.3) and (fresh.ravel()[1:] == np.float32(0.5)).all()
    # the resize decisions come from the library's own restatement of the policy (no reference map object involved)
    bounds = matcher.MapBounds(g.size_x, g.size_y, g.res, g.off_x, g.off_y, mf.EXTEND)
    half = ctx.lib.rsm_blur_half_size(g.sigma, g.res)
    n_ext = 0
    for k, (p, s) in enumerate(zip(poses, pts)):
        sx, sy, ox, oy = (int(z["geom"][k][0]), int(z["geom"][k][1]), float(z["geom"][k][2]), float(z["geom"][k][3]))
        fits, geom, pre = bounds.UpdateMapByRange(s, p, half, True)
        assert fits == bool(z["stamped"][k]) and geom == (sx, sy, ox, oy)
        if fits:
            dg.UpdateMapByRange(s, p, g.sigma, g.occu_offset, True)
        else:
            dg.ExtendSize(sx, sy, pre, (ox, oy), 0.5, g.default_prob)
            n_ext += 1
        assert dg.geometry() == (sx, sy, ox, oy)
        cells = dg.download()
        assert hashlib.sha256(cells.tobytes()).hexdigest() == str(z["shas"][k]), "step %d" % k
    assert n_ext == 5 and dg.is_fixed_point()
    final = np.full(cells.size, np.float32(0.5), dtype=np.float32)
    final[z["final_nz_index"]] = z["final_nz_value"]
    assert np.array_equal(cells.ravel(), final)
    # the grown map is a normal lookup grid: match the last scan against it
    m = matcher.BasedCorrelationScanMatch(ctx)
    pose, cov = poses[-1] + np.array([0.08, -0.05, 0.04]), np.eye(3)
    r = m.ScanMatch(dg, pts[-1], synth.chain_yaml()[0], pose, cov)
    assert r > 0.6 and np.hypot(*(pose[:2] - poses[-1][:2])) < 0.06
    dg.close()


def test_publishing_map_on_device(ctx):
    """PubMap (CountCell: hit / pass / value per cell) updated on the device with ray-traced free space across
    the 44-scan trajectory, with the knobs SlamProcessor::UpdateMap sets, against the reference's own cells after
    every step (fixture); then the occupancy it yields feeds the map check without any upload and reproduces
    the reference's MapCheckPenalize coefficients."""
    import hashlib
    import importlib.util
    import os
    here = os.path.dirname(__file__)
    mods = {}
    import sys
    sys.path.insert(0, os.path.join(here, "golden"))
    try:
        for name in ("make_frontend", "make_pubmap"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(here, "golden", name + ".py"))
            mods[name] = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mods[name])          # pure helpers only (trajectory, scans, spec, factors)
    finally:
        sys.path.pop(0)
    mf, mp = mods["make_frontend"], mods["make_pubmap"]
    z = np.load(os.path.join(here, "golden", "pubmap_willow.npz"), allow_pickle=False)
    g = mf.spec()
    poses, pts = mf.trajectory(), mf.scans()
    pm = matcher.PubMap(ctx, g.res, g.size_x, g.size_y, g.off_x, g.off_y)
    bounds = matcher.MapBounds(g.size_x, g.size_y, g.res, g.off_x, g.off_y, mf.EXTEND)
    for k, (p, s) in enumerate(zip(poses, pts)):
        sx, sy, ox, oy = (int(z["geom"][k][0]), int(z["geom"][k][1]), float(z["geom"][k][2]), float(z["geom"][k][3]))
        f = mp.factors(k)
        fits, geom, pre = bounds.UpdateMapByRange(s, p, 0, False)
        assert fits == bool(z["stamped"][k]) and geom == (sx, sy, ox, oy)
        if fits:
            pm.UpdateMapByRange(s, p, f[0], f[1])
        else:
            pm.ExtendSize(sx, sy, pre, (ox, oy))
        assert pm.geometry() == (sx, sy, ox, oy)
        val, cnt, hit, _ = pm.download_all()
        h = hashlib.sha256()
        for a in (val, cnt, hit):
            h.update(a.tobytes())
        assert h.hexdigest() == str(z["shas"][k]), "step %d" % k
    pm.refresh_occupancy(float(z["knobs"][0]), float(z["knobs"][1]))
    occ = pm.download_all()[3]
    want_occ = np.unpackbits(z["occ_packed"])[: occ.size].reshape(occ.shape)
    assert np.array_equal(occ, want_occ) and occ.sum() > 1000
    got = pm.MapCheckPenalize(pts[-1], z["check_poses"], 100, 2.5, 0.015, True)
    assert np.array_equal(got, z["coeff"]) and len(np.unique(got)) >= 10
    pm.close()


def test_map_rebuilds_from_the_scan_store(ctx, oracle, rng):
    """CorrectPoseAndMap's rebuilds (slam_processor.cpp:329-371) from a device scan store: the scan-match map
    against the oracle's InitMapWithRangeVec, the publishing map against the scan-by-scan device path (itself
    pinned to the reference by test_publishing_map_on_device), after the pose graph moved every scan."""
    sc = synth.config4(1)[0]
    g = sc.grid
    store = matcher.ScanStore(ctx)
    ids = [store.AddRangeData(p, pose) for p, pose in zip(sc.base_pts, sc.base_poses)]
    new_poses = sc.base_poses + rng.uniform(-0.05, 0.05, sc.base_poses.shape)
    store.UpdateRangeData(ids, new_poses)
    order = ids[::-1] + [ids[0]] * 3          # any order, repeats allowed (the reference appends id 0 min_passthrough times)
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    dg.InitMapWithStore(store, order, g.default_prob, g.sigma, g.occu_offset, True)
    want = oracle.build_grid(g, [sc.base_pts[i] for i in order], new_poses[order])
    assert np.array_equal(dg.download(), want)
    dg.InitMapWithStore(store, [], g.default_prob, g.sigma, g.occu_offset, True)      # Reset only
    assert (dg.download() == np.float32(g.default_prob)).all()
    a = matcher.PubMap(ctx, g.res, g.size_x, g.size_y, g.off_x, g.off_y)
    b = matcher.PubMap(ctx, g.res, g.size_x, g.size_y, g.off_x, g.off_y)
    b.UpdateMapByRange(sc.scan_pts, sc.seed_pose, 4.0, 8.0)                            # stale content the rebuild must wipe
    b.InitMapWithRangeVec(store, order, 0.3, 0.7)
    for i in order:
        a.UpdateMapByRange(sc.base_pts[i], new_poses[i], 0.3, 0.7)
    for x, y in zip(a.download_all()[:3], b.download_all()[:3]):
        assert np.array_equal(x, y)
    assert a.download_all()[1].max() >= len(order) - 1
    for m_ in (a, b, dg):
        m_.close()
    store.close()


def test_patch_kernel_variant(ctx, oracle, monkeypatch):
    """The shared-memory patch kernel (batches of small unit-step windows): forced on single matches so that every
    window width 3..16, boxes that fit and that do not (far beams, wide angle fans), beams on cell boundaries and
    patches at the grid border are compared score by score; then a batch large enough to select it by itself."""
    rng = np.random.default_rng(13579)
    m = matcher.BasedCorrelationScanMatch(ctx)
    levels = np.array([0.3, 0.5642, 0.6666, 0.7046, 0.7875, 0.8324, 1.0], dtype=np.float32)
    monkeypatch.setenv("RSM_FORCE_PATCH", "1")
    for trial in range(60):
        size = int(rng.choice([120, 300, 544, 700]))
        res = 0.05
        n_xy = int(rng.choice([3, 5, 7, 8, 9, 12, 13, 14, 15, 16]))
        window = res * (n_xy - 1)
        n_ang = int(rng.choice([1, 5, 8, 9, 21, 30]))
        ares = float(rng.choice([0.0349, 0.1, 0.005]))
        aoff = ares * (n_ang - 1) / 2 + 1e-9
        P = int(rng.choice([1, 31, 32, 33, 100, 360, 940]))
        reach = size / 2 - n_xy / 2 - (1 if trial % 5 == 0 else 6)      # every fifth problem brushes the grid border
        radius = rng.uniform(0.05, 1.0, P) ** float(rng.choice([0.3, 1.0, 3.0])) * reach
        ang = np.sort(rng.uniform(-np.pi, np.pi, P)) if trial % 2 else rng.uniform(-np.pi, np.pi, P)
        pts = np.stack([np.cos(ang) * radius, np.sin(ang) * radius], axis=1)
        snap = rng.random(P) < rng.choice([0.0, 0.2, 1.0])
        pts[snap] = np.round(pts[snap] * 2) / 2
        g = synth.GridSpec(res, 0.15, size, size, float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        grid = levels[rng.integers(0, len(levels), (size, size))]
        centre_cells = np.array([size / 2, size / 2]) + (rng.integers(-1, 2, 2) if rng.random() < 0.5 else rng.uniform(-1, 1, 2))
        theta = float(rng.choice([0.0, aoff, rng.uniform(-3, 3)]))
        seed = np.array([centre_cells[0] * res - g.off_x, centre_cells[1] * res - g.off_y, theta])
        p = synth.pass_param(window, res, aoff, ares, 0.3, int(rng.choice([100000, 50])), bool(rng.integers(0, 2)), int(rng.integers(0, 3)))
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.upload(grid)
        centre = oracle.world_to_map(g, seed)
        try:
            sd = m.scores(dg, pts, p, seed)
        except matcher.RsmError as e:            # a patch really left the grid: the reference would read out of bounds
            assert "RSM_ERR_WINDOW" in str(e) and trial % 5 == 0
            dg.close()
            continue
        so = oracle.scores(grid, g, pts, p, centre)
        assert np.array_equal(so, sd), (trial, size, n_xy, n_ang, P, int((so != sd).sum()))
        want = oracle.match(grid, g, pts, p, seed)
        pose, cov = seed.copy(), np.eye(3)
        assert_pass_equal(m.ScanMatch(dg, pts, p, pose, cov), pose, cov, want)
        dg.close()
    monkeypatch.delenv("RSM_FORCE_PATCH")
    # a batch that selects the patch kernel by itself (150 chains x 4 angle groups >= 592 CTAs), against the oracle
    pairs = synth.config4(150, seed=2468)
    packed = matcher.pack_loop_closure(pairs)
    st0 = ctx.stats()["score_launches"]
    scores, poses, covs, resp = matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
    for i, sc in enumerate(pairs):
        if i % 3:
            continue                      # the oracle chain costs 20 ms per pair: check every third
        w = oracle.match_chain(oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses), sc.grid, sc.scan_pts, sc.passes, sc.seed_pose)
        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
        assert np.array_equal(resp[i], w["responses"])
    monkeypatch.setenv("RSM_NO_PATCH", "1")     # and the whole batch against the tiled kernel
    scores2, poses2, covs2, resp2 = matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
    assert np.array_equal(scores, scores2) and np.array_equal(poses, poses2) and np.array_equal(covs, covs2) and np.array_equal(resp, resp2)
    assert ctx.stats()["score_launches"] >= st0


# ---- round 2: non-blur update, pipelined lanes, heterogeneous batches, strict ties -----------------------
def test_non_blur_update_paths(ctx, oracle):
    """SET_CELL_OCCUPIED (occu_grid_map.h:317-321, 499-516): UpdateMapByRange with use_blur off, and with blur
    parameters the reference rejects, on a constructed front-end map scan by scan, as a rebuild from the scan store
    and inside the batched back-end step -- every cell and every chain result against the oracle (whose non-blur path
    is pinned against the reference: tests/golden/nonblur_*.npz, tests/test_oracle.py)."""
    import dataclasses
    sc = synth.config4(1, seed=77)[0]
    for g in (dataclasses.replace(sc.grid, use_blur=False), dataclasses.replace(sc.grid, sigma=5.0)):
        want = oracle.build_grid(g, sc.base_pts, sc.base_poses)
        assert set(np.unique(want)) <= {np.float32(0.3), np.float32(0.8), np.float32(1.0)} and (want == np.float32(0.8)).sum() > 50
        # (1) one call
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
        assert np.array_equal(dg.download(), want)
        # (2) scan by scan on a constructed map (cell 0 = 0.3, the others 0.5): 0.5 -> 1.0 on the first hit
        dg.fill(0.5, 0.3)
        host = np.full((g.size_y, g.size_x), np.float32(0.5)); host[0, 0] = np.float32(0.3)
        for pts, pose in zip(sc.base_pts, sc.base_poses):
            dg.UpdateMapByRange(pts, pose, g.sigma, g.occu_offset, g.use_blur)
            oracle.grid_stamp(host, g, pts, pose)
        assert np.array_equal(dg.download(), host) and (host == np.float32(1.0)).sum() > 100
        # (3) a float32-held map takes the same path (one non-representable cell forces float cells)
        bumped = np.full((g.size_y, g.size_x), np.float32(0.3)); bumped[0, 0] = np.float32(0.1)
        dg.upload(bumped)
        assert not dg.is_fixed_point()
        hostf = bumped.copy()
        for pts, pose in zip(sc.base_pts[:3], sc.base_poses[:3]):
            dg.UpdateMapByRange(pts, pose, g.sigma, g.occu_offset, g.use_blur)
            oracle.grid_stamp(hostf, g, pts, pose)
        assert np.array_equal(dg.download(), hostf)
        # (4) rebuild from the scan store
        store = matcher.ScanStore(ctx)
        ids = [store.AddRangeData(p, q) for p, q in zip(sc.base_pts, sc.base_poses)]
        mid = store.AddRangeData(sc.scan_pts, sc.seed_pose)
        dg.InitMapWithStore(store, ids, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
        assert np.array_equal(dg.download(), want)
        dg.close()
        # (5) the batched back-end step only has the sigma to go by (use_blur is on in the call, as shipped)
        if g.use_blur:
            w = oracle.match_chain(want, g, sc.scan_pts, sc.passes, sc.seed_pose)
            s_, p_, c_, r_ = matcher.scan_match_interface_batch(ctx, store, g, [sc.grid_centre] * 3, [ids] * 3, [mid] * 3,
                                                                [sc.seed_pose] * 3, sc.passes)
            for i in range(3):
                assert s_[i] == w["score"] and np.array_equal(p_[i], w["pose"]) and cov_close(c_[i], w["cov"]) and np.array_equal(r_[i], w["responses"])
        store.close()


def test_pipelined_lanes_equal_one_lane(ctx, oracle):
    """A batched chain call cut into sub-batches over several streams (RSM_OPT_LANES) returns what the single-lane
    call and the oracle return, item for item -- including items whose exact-tie path or gather round trip makes one
    lane's host stage longer, an empty chain, and a batch smaller than the lane count."""
    n = 37
    pairs = synth.config4(n, seed=2468)
    want = []
    for sc in pairs:
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        want.append(oracle.match_chain(grid, sc.grid, sc.scan_pts, sc.passes, sc.seed_pose))
    packed = matcher.pack_loop_closure(pairs)
    store = matcher.ScanStore(ctx)
    chains, mids = [], []
    for sc in pairs:
        chains.append([store.AddRangeData(p, q) for p, q in zip(sc.base_pts, sc.base_poses)])
        mids.append(store.AddRangeData(sc.scan_pts, sc.seed_pose))
    try:
        for lanes in (1, 2, 3, 8):
            ctx.set_option(matcher.RSM_OPT_LANES, lanes)
            for rep in range(2):      # the second call reuses the lanes' buffers
                res = [matcher.loop_closure_batch(ctx, packed, pairs[0].passes),
                       matcher.scan_match_interface_batch(ctx, store, pairs[0].grid, [sc.grid_centre for sc in pairs], chains, mids,
                                                          [sc.seed_pose for sc in pairs], pairs[0].passes)]
                for scores, poses, covs, resp in res:
                    for i, w in enumerate(want):
                        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"]), (lanes, rep, i)
                        assert np.array_equal(resp[i], w["responses"])
        # fewer items than lanes, and an item without base scans in the middle of a sub-batch
        ctx.set_option(matcher.RSM_OPT_LANES, 4)
        ch2 = [chains[0], [], chains[2]]
        s_, p_, c_, r_ = matcher.scan_match_interface_batch(ctx, store, pairs[0].grid, [pairs[i].grid_centre for i in range(3)], ch2, mids[:3],
                                                            [pairs[i].seed_pose for i in range(3)], pairs[0].passes)
        assert s_[1] == 0.0 and np.array_equal(p_[1], pairs[1].seed_pose) and np.array_equal(c_[1], np.eye(3))
        for i in (0, 2):
            assert s_[i] == want[i]["score"] and np.array_equal(p_[i], want[i]["pose"])
        # existing grids through rsm_match_batch on several lanes
        grids = [device_grid(ctx, sc) for sc in pairs[:9]]
        sm = matcher.ScanMatchers(ctx, pairs[0].passes)
        ctx.set_option(matcher.RSM_OPT_LANES, 3)
        scores, poses, covs, resp = sm.ScanMatchBatch(grids, [sc.scan_pts for sc in pairs[:9]], [sc.seed_pose for sc in pairs[:9]])
        for i in range(9):
            assert scores[i] == want[i]["score"] and np.array_equal(poses[i], want[i]["pose"]) and cov_close(covs[i], want[i]["cov"])
        for dg in grids:
            dg.close()
    finally:
        ctx.set_option(matcher.RSM_OPT_LANES, 0)
        store.close()


def test_heterogeneous_batch_is_grouped_by_plan(ctx, oracle, rng):
    """rsm_match_batch with per-pair parameters (shared_params = 0): a 49-wide unit-step window (staged kernel), a
    13-wide one (tiled / patch) and a 5-wide half-cell one (flat) in ONE call -- every item must get the kernel plan of
    its own window, not the first item's."""
    import ctypes
    scs = [random_scenario(rng, n_points=150 + 40 * k, size=260) for k in range(5)]
    wins = [synth.pass_param(2.4, 0.05, 0.07, 0.0349, 0.1, 100000, True, 0),    # 49 x 49, f = 1
            synth.pass_param(0.6, 0.05, 0.21, 0.0349, 0.1, 100000, True, 0),    # 13 x 13, f = 1
            synth.pass_param(0.1, 0.025, 0.07, 0.0349, 0.1, 60, True, 0),       # 5 x 5, f = 0.5
            synth.pass_param(0.8, 0.1, 0.14, 0.0349, 0.1, 100000, False, 0),    # 9 x 9, f = 2
            synth.pass_param(2.4, 0.05, 0.035, 0.0349, 0.1, 100000, True, 0)]   # 49 x 49 again
    fine = synth.pass_param(0.2, 0.02, 0.07, 0.0349, 0.1, 100000, True, 1)
    sup = synth.pass_param(0.02, 0.01, 0.0349, 0.00349, 0.1, 100000, True, 2)
    grids, want = [], []
    for sc, w0 in zip(scs, wins):
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        want.append(oracle.match_chain(grid, sc.grid, sc.scan_pts, [w0, fine, sup], sc.seed_pose))
        grids.append(device_grid(ctx, sc))
    n = len(scs)
    params = (matcher.PassParamStruct * (3 * n))(*[matcher._as_param(p).struct() for w0 in wins for p in (w0, fine, sup)])
    offs = np.zeros(n + 1, dtype=np.int64)
    for i, sc in enumerate(scs):
        offs[i + 1] = offs[i] + len(sc.scan_pts)
    pts = np.ascontiguousarray(np.concatenate([sc.scan_pts for sc in scs], axis=0))
    poses = np.ascontiguousarray(np.array([sc.seed_pose for sc in scs]))
    covs = np.tile(np.eye(3), (n, 1, 1))
    scores, resp = np.zeros(n), np.zeros((n, 3))
    handles = (ctypes.c_void_p * n)(*[dg.h for dg in grids])
    ctx.check(ctx.lib.rsm_match_batch(ctx.h, n, handles, pts.ctypes.data, offs.ctypes.data, params, 0, 1, poses.ctypes.data,
                                      covs.ctypes.data, scores.ctypes.data, resp.ctypes.data))
    for i, w in enumerate(want):
        assert scores[i] == w["score"] and np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"]), i
        assert np.array_equal(resp[i], w["responses"])
    for dg in grids:
        dg.close()


def test_strict_ties_option(ctx, oracle, rng):
    """RSM_OPT_STRICT_TIES: on a binary grid the centre penalty leaves mirror-image candidates exactly tied inside the
    20-element covariance prefix.  Default: same consumed set, covariance within 1e-6 (the contract).  With the option
    the pass takes the reference's own sort and the covariance is bit-equal."""
    sc, _ = load_golden("ties_icra")
    dg = device_grid(ctx, sc)
    grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
    m = matcher.BasedCorrelationScanMatch(ctx)
    try:
        for strict in (0, 1):
            ctx.set_option(matcher.RSM_OPT_STRICT_TIES, strict)
            exact = 0
            for k, p in enumerate(sc.passes + [synth.pass_param(0.6, 0.05, 0.349, 0.0349, 0.2, 40, True, 0)]):
                want = oracle.match(grid, sc.grid, sc.scan_pts, p, sc.seed_pose)
                pose, cov = sc.seed_pose.copy(), np.eye(3)
                r = m.ScanMatch(dg, sc.scan_pts, p, pose, cov)
                assert_pass_equal(r, pose, cov, want)
                if strict:
                    assert np.array_equal(cov, want["cov"]), k
                exact += m.last_detail.exact_sort_used
            if strict:
                assert exact >= 1
    finally:
        ctx.set_option(matcher.RSM_OPT_STRICT_TIES, 0)
    with pytest.raises(matcher.RsmError):
        ctx.set_option(99, 1)
    dg.close()


def test_in_library_sliced_match_one_rank(ctx, oracle):
    """rsm_match_sliced with a one-rank communicator (no NCCL needed): device-side packing of the partial, merge, the
    column gather into the exchange buffer, finish -- and the gathered exact path on the tie-heavy fixture."""
    sm = matcher.CommSlicedScanMatch(ctx, 0, 1)
    fallbacks = 0
    for sc in (synth.config1(), load_golden("ties_icra")[0], synth.config3(True)):
        grid = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        dg = device_grid(ctx, sc)
        pose_in = sc.seed_pose.copy()
        for p in sc.passes:
            want = oracle.match(grid, sc.grid, sc.scan_pts, p, pose_in)
            pose, cov = pose_in.copy(), np.eye(3)
            r = sm.ScanMatch(dg, sc.scan_pts, p, pose, cov)
            assert_pass_equal(r, pose, cov, want)
            assert sm.last_detail.n_avg == want["n_avg"]
            if sm.exact_fallback:
                assert np.array_equal(cov, want["cov"])
                fallbacks += 1
            pose_in = want["pose"]
        dg.close()
    assert fallbacks >= 1
    sm.close()


def test_in_library_sliced_match_over_nccl():
    """The same over real ranks: torchrun with one process per visible GPU (needs two or more), every rank compares
    its result with the oracle (tools/sliced_check.py)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tools", "sliced_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("sliced matches equal the oracle") == n


def test_batched_backend_step_with_optimiser(ctx, oracle):
    """rsm_scan_match_interface_batch_opt: the coarse map + Gauss-Newton pre-step inside the batched back-end step
    (slam_processor.cpp:282-308, scan_matchers.h:205-232) against the oracle's chain with the optimiser on, candidate by
    candidate -- seeds close enough for the optimiser to succeed (its pose feeds the fine pass), seeds it fails on (the
    coarse correlative pass runs from the seed), use_fine off, and the map check on top."""
    from helpers import load_mapcheck
    pairs = synth.config4(5)
    fine, coarse, pub = matcher.ScanStore(ctx), matcher.ScanStore(ctx), matcher.ScanStore(ctx)
    chains, mids = [], []
    for sc in pairs:
        ids = []
        for p, pose in zip(sc.base_pts, sc.base_poses):
            ids.append(fine.AddRangeData(p, pose))
            assert coarse.AddRangeData(p * 0.5, pose) == ids[-1] and pub.AddRangeData(p, pose) == ids[-1]
        mid = fine.AddRangeData(sc.scan_pts, sc.seed_pose)
        assert coarse.AddRangeData(sc.scan_pts * 0.5, sc.seed_pose) == mid and pub.AddRangeData(sc.scan_pts, sc.seed_pose) == mid
        chains.append(ids)
        mids.append(mid)
    gf = pairs[0].grid
    gc = synth.backend_grid(gf.res * 2, gf.sigma * 2, 10.0, pairs[0].grid_centre)
    op = (10, 0.1, 0.5, 0.5, 0.5)
    deltas = ([0.02, 0.01, 0.02], [0.12, -0.07, 0.1], [0.5, 0.4, -0.3], [0.0, 0.0, 0.0], [-0.2, 0.15, -0.12])
    cands = [(i, pairs[i].truth_pose + np.array(deltas[(i + k) % len(deltas)])) for i in range(len(pairs)) for k in range(3)]
    gpub, occ, _ = load_mapcheck("mapcheck_pair0")
    pm = matcher.ScanMatchMap.from_spec(ctx, gpub)
    pm.upload_occupancy(occ)
    check = (100, 2.5, 0.015, True)
    branches = set()
    for failed_cost, use_fine, with_check in ((2.0, True, False), (20.0, True, True), (100.0, True, False), (90.0, False, False)):
        want = []
        for who, seed in cands:
            sc = pairs[who]
            gci = synth.backend_grid(gc.res, gc.sigma, 10.0, sc.grid_centre)
            grid_f = oracle.build_grid(sc.grid, sc.base_pts, sc.base_poses)
            grid_c = oracle.build_grid(gci, [p * 0.5 for p in sc.base_pts], sc.base_poses)
            w = oracle.match_chain_opt(grid_c, gci, sc.scan_pts * 0.5, grid_f, sc.grid, sc.scan_pts, sc.passes, op, failed_cost, seed,
                                       use_fine=use_fine)
            branches.add((bool(w["optimize_cost"] > failed_cost), use_fine))
            if with_check:
                c = oracle.map_check_penalize(occ, gpub, sc.scan_pts, w["pose"], *check)
                s = w["score"] * c
                w = dict(w, score=1.0 if s > 1.0 else s)
            want.append(w)
        scores, poses, covs, resp = matcher.scan_match_interface_batch_opt(
            ctx, fine, coarse, gf, gc, [pairs[w_].grid_centre for w_, _ in cands], [chains[w_] for w_, _ in cands],
            [mids[w_] for w_, _ in cands], [s_ for _, s_ in cands], pairs[0].passes, op, failed_cost, use_fine,
            pm if with_check else None, pub if with_check else None, check if with_check else None)
        for i, w in enumerate(want):
            assert scores[i] == w["score"], (failed_cost, use_fine, i, scores[i], w["score"])
            assert np.array_equal(poses[i], w["pose"]) and cov_close(covs[i], w["cov"])
            assert resp[i][0] == w["optimize_cost"] and np.array_equal(resp[i][1:], w["responses"])
    assert {(False, True), (True, True), (False, False)} <= branches      # optimiser kept / dropped with the fine passes on; fine passes off
    pm.close()
    for st in (fine, coarse, pub):
        st.close()
