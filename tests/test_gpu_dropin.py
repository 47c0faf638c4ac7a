"""GPU test of the drop-in boundary: the C++ adapter class (scan_matcher_adapter.hpp) is handed
the reference's own live ScanMatchMap / RangeDataContainer2d / CorrelationScanMatchParam objects
and must return what the reference's BasedCorrelationScanMatch returns for them.  Runs only where
oracle/_ref was built (the build container had /root/reference; the built files travel)."""
import numpy as np
import pytest

from helpers import cov_close, load_golden
from roborts_edu_slam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin():
    from oracle.oracle_py import DropIn, dropin_available, ref_available
    if not (dropin_available() and ref_available()):
        pytest.skip("oracle/_ref not built")
    d = DropIn(0)
    yield d
    d.close()


def _check(ref, dropin, sc, chain):
    m = ref.create_map(sc.grid)
    try:
        ref.build_map(m, sc.grid, sc.base_pts, sc.base_poses)
        for p in sc.passes:
            want = ref.match(m, sc.scan_pts, p, sc.seed_pose)
            got = dropin.match_chain(m, sc.scan_pts, [p], sc.seed_pose)
            assert got["score"] == want["response"]
            assert np.array_equal(got["pose"], want["pose"])
            assert cov_close(got["cov"], want["cov"])
        if chain:
            want = ref.match_chain(m, sc.scan_pts, sc.passes, sc.seed_pose)
            got = dropin.match_chain(m, sc.scan_pts, sc.passes, sc.seed_pose)
            assert got["score"] == want["score"]
            assert np.array_equal(got["pose"], want["pose"])
            assert cov_close(got["cov"], want["cov"])
            assert np.array_equal(got["responses"], want["responses"])
    finally:
        ref.destroy_map(m)


def test_adapter_on_live_reference_objects(ref, dropin):
    _check(ref, dropin, synth.config1(), False)
    _check(ref, dropin, synth.config3(True), True)
    for sc in synth.config4(3, seed=7):
        _check(ref, dropin, sc, True)
    sc, _ = load_golden("ties_icra")
    _check(ref, dropin, sc, True)


def test_adapter_follows_map_updates(ref, dropin):
    """The adapter re-uploads the grid when the live map changes (map_update_index)."""
    a, b = synth.config4(2, seed=11)
    m = ref.create_map(a.grid)
    try:
        ref.build_map(m, a.grid, a.base_pts, a.base_poses)
        w1 = ref.match(m, a.scan_pts, a.passes[0], a.seed_pose)
        g1 = dropin.match_chain(m, a.scan_pts, [a.passes[0]], a.seed_pose)
        assert g1["score"] == w1["response"] and np.array_equal(g1["pose"], w1["pose"])
        # same map object, new contents and offset (what ResetScanMatchMapWithRangeVec does)
        ref.L.ref_map_set_offset(m, b.grid.off_x, b.grid.off_y)
        ref.build_map(m, b.grid, b.base_pts, b.base_poses)
        w2 = ref.match(m, b.scan_pts, b.passes[0], b.seed_pose)
        g2 = dropin.match_chain(m, b.scan_pts, [b.passes[0]], b.seed_pose)
        assert g2["score"] == w2["response"] and np.array_equal(g2["pose"], w2["pose"])
        assert cov_close(g2["cov"], w2["cov"])
    finally:
        ref.destroy_map(m)


def test_optimize_adapter_on_live_reference_objects(ref, dropin):
    """rsm_adapter::BasedOptimizeScanMatch handed a live reference map, scan and OptimizeScanMatchParam returns what
    the reference's BasedOptimizeScanMatch returns for them (cost and pose, bit for bit), including the invalid-input
    behaviour (kMaxCost, pose untouched)."""
    ops = ((10, 0.1, 0.5, 0.5, 0.5), (10, 1.0, 2.0, 0.5, 0.2), (3, 1e-9, 0.0, 0.02, 0.01), (1, 0.1, 0.5, 0.5, 0.5))
    deltas = ([0.12, -0.07, 0.1], [0.02, 0.01, 0.02], [-0.2, 0.15, -0.12], [0.5, 0.4, -0.3], [0.0, 0.0, 0.0])
    for sc in [synth.config1()] + synth.config4(2, seed=5):
        m = ref.create_map(sc.grid)
        try:
            seed = sc.truth_pose + np.array(deltas[0])
            got = dropin.optimize(m, sc.scan_pts, ops[0], seed)          # map not initialised yet
            assert got["cost"] == 1000.0 and np.array_equal(got["pose"], seed)
            ref.build_map(m, sc.grid, sc.base_pts, sc.base_poses)
            for op in ops:
                for d in deltas:
                    seed = sc.truth_pose + np.array(d)
                    want = ref.optimize(m, sc.scan_pts, op, seed)
                    got = dropin.optimize(m, sc.scan_pts, op, seed)
                    assert got["cost"] == want["cost"], (sc.name, op, d)
                    assert np.array_equal(got["pose"], want["pose"])
            got = dropin.optimize(m, sc.scan_pts[:0], ops[0], seed)      # empty scan
            assert got["cost"] == 1000.0 and np.array_equal(got["pose"], seed)
        finally:
            ref.destroy_map(m)


def frontend_sequence(n_scans=14):
    """A front-end run on the willow map: (fine GridSpec, [(pose, scan in fine-map cells)])."""
    occ = synth.load_map("willow")
    tr = [np.array([14.375 + 0.06 * k, 28.625 + 0.025 * k, 0.3 + 0.012 * k]) for k in range(n_scans)]
    g = synth.backend_grid(0.01, 0.03, 10.0, tr[0][:2])
    scans = [synth.raycast(occ, p[0], p[1], p[2], 1081, np.deg2rad(270.25), 10.0) * (1 / 0.01) for p in tr]
    return g, tr, scans


def test_frontend_through_the_adapter(ref, dropin):
    """The shipped front-end call sequence on live reference objects -- match the new scan against the fine map
    (coarse / fine / super chain, scan_matchers.h:224-263), then insert it at the matched pose (slam_processor.cpp:558) --
    with the adapter's matcher and rsm_adapter::UpdateMapByRange against the reference's own classes, step by step.
    The device mirror is uploaded ONCE and then follows the host map through the incremental stamps."""
    g, tr, scans = frontend_sequence()
    passes = synth.chain_defaults((100, 100, 200))
    ma, mr = ref.frontend_map_create(g, 0.2), ref.frontend_map_create(g, 0.2)
    full0, inc0 = dropin.sync_counts()
    try:
        for m in (ma, mr):
            ref.frontend_map_update(m, scans[0], tr[0], True)
        stamped = 0
        for k in range(1, len(scans)):
            seed = tr[k] + np.array([0.03, -0.02, 0.015])
            want = ref.match_chain(mr, scans[k], passes, seed)
            got = dropin.match_chain(ma, scans[k], passes, seed)
            assert got["score"] == want["score"] and np.array_equal(got["pose"], want["pose"]) and cov_close(got["cov"], want["cov"]), k
            ok_r = ref.frontend_map_update(mr, scans[k], want["pose"], True)[0]
            ok_a = dropin.update_map(ma, scans[k], got["pose"], True, g.sigma, g.occu_offset)
            assert ok_r == ok_a
            stamped += int(ok_a)
        assert dropin.mirror_equals_host(ma) == 1
        sx, sy = ref.map_size(ma)
        assert np.array_equal(ref.read_map_sized(ma, sx, sy), ref.read_map_sized(mr, sx, sy))
        full1, inc1 = dropin.sync_counts()
        assert full1 - full0 == 1 and inc1 - inc0 == stamped and stamped >= 10, (full1 - full0, inc1 - inc0, stamped)
    finally:
        ref.destroy_map(ma)
        ref.destroy_map(mr)


def test_batched_loop_closure_on_a_live_sensor_data_manager(ref):
    """rsm_adapter::BackEndBatcher (csrc/backend_batcher.hpp) compiled against the unmodified reference headers: scans live
    in a reference SensorDataManager (metres + per-map copies, as SlamProcessor stores them); TryCloseLoop's two-stage test
    over several candidate chains of one scan runs as two batched calls and must pick the chain, pose and covariance that
    the reference's sequential loop (range_scan_pose_graph.cpp:299-352, ScanMatchInterface with its own classes) picks."""
    from oracle.oracle_py import Batcher, batcher_available
    if not batcher_available():
        pytest.skip("oracle/_ref/libbatcher.so not built")
    occ = synth.load_map("willow")
    base = np.array([14.375, 28.625, 0.3])
    poses = [base + np.array([0.25 * k, 0.1 * k, 0.04 * k]) for k in range(24)]
    scans_m = [synth.raycast(occ, p[0], p[1], p[2], 1081, np.deg2rad(270.25), 10.0) for p in poses]
    passes = synth.chain_yaml((100, 100, 200))
    g = synth.backend_grid(0.05, 0.15, 10.0, base[:2])
    B = Batcher((0.05, 0.1, 0.05), (0.15, 0.3), (g.size_x, g.size_x // 2), 0.88, 0.3, passes, (0, 10, 0.1, 0.5, 0.5, 0.5, 20.0))
    try:
        ids = [B.add_scan(s, p) for s, p in zip(scans_m, poses)]
        assert ids == list(range(24))
        query = 23
        q_pose = poses[query] + np.array([0.08, -0.05, 0.03])        # the scan's (drifted) pose
        B.set_pose(query, q_pose)
        # chain 0 lies far from the query scan (fails stage 1), chain 1 near but sparse, chain 2 the good one
        chains = [list(range(0, 6)), list(range(10, 14)), list(range(15, 23)), list(range(16, 22))]
        centre = q_pose[:2]
        th = (0.6, 0.05, 0.7)         # chain 0 fails the coarse test, chain 1 the fine one, chain 2 closes the loop

        def reference_try_close_loop():
            stage1, stage2 = [], []
            hit, best, bcov = -1, None, None
            for ci, ch in enumerate(chains):
                gi = synth.backend_grid(0.05, 0.15, 10.0, centre)
                m = ref.create_map(gi)
                ref.build_map(m, gi, [B.fine_scan(i) for i in ch], np.array([poses[i] if i != query else q_pose for i in ch]))
                w1 = ref.match_chain(m, B.fine_scan(query), passes, q_pose)
                stage1.append(w1["score"])
                s2 = -1.0
                if w1["score"] > th[0] and w1["cov"][0, 0] < th[1] and w1["cov"][1, 1] < th[1]:
                    w2 = ref.match_chain(m, B.fine_scan(query), passes, w1["pose"], w1["cov"])
                    s2 = w2["score"]
                    if hit < 0 and s2 >= th[2]:
                        hit, best, bcov = ci, w2["pose"], w2["cov"]
                stage2.append(s2)
                ref.destroy_map(m)
            return hit, best, bcov, stage1, stage2

        want = reference_try_close_loop()
        hit, best, cov, s1, s2 = B.try_close_loop(query, chains, q_pose, centre, th)
        assert hit == want[0] == 2, (hit, want[0], s1, want[3], s2, want[4])
        assert np.array_equal(s1, want[3]) and np.array_equal(s2, want[4])
        assert np.array_equal(best, want[1]) and cov_close(cov, want[2])
        assert sum(1 for v in want[4] if v < 0) >= 1      # at least one chain stopped at stage 1
    finally:
        B.close()
