"""Step-by-step GPU bring-up: grid parity, score parity, match parity vs the oracle, with timing.
Run on the GPU box:  python tools/gpu_debug.py > gpurun_out/debug.log 2>&1"""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
from oracle.oracle_py import Oracle, Ref, ref_available

O = Oracle()
ctx = matcher.Context(0)
ctx.set_profiling(True)
M = matcher.BasedCorrelationScanMatch(ctx)

def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print("[ok] %s (%.3fs) %s" % (name, time.time() - t, "" if r is None else r), flush=True)
        return r
    except Exception:
        print("[FAIL] %s\n%s" % (name, traceback.format_exc()), flush=True)

def check_scenario(sc, chain=False):
    g = sc.grid
    ogrid = O.build_grid(g, sc.base_pts, sc.base_poses)
    dg = matcher.ScanMatchMap.from_spec(ctx, g)
    def raster():
        dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
        got = dg.download()
        return "fixed=%s equal=%s ndiff=%d" % (dg.is_fixed_point(), np.array_equal(got, ogrid), int((got != ogrid).sum()))
    step(sc.name + " rasterize", raster)
    ug = matcher.ScanMatchMap.from_spec(ctx, g)
    step(sc.name + " upload", lambda: (ug.upload(ogrid), "equal=%s" % np.array_equal(ug.download(), ogrid))[1])
    cm = O.world_to_map(g, sc.seed_pose)
    step(sc.name + " world_to_map", lambda: "equal=%s" % np.array_equal(dg.GetMapCoordsPose(sc.seed_pose), cm))
    pose = sc.seed_pose.copy()
    for pi, p in enumerate(sc.passes):
        def scores():
            so = O.scores(ogrid, g, sc.scan_pts, p, O.world_to_map(g, pose))
            sd = M.scores(dg, sc.scan_pts, p, pose)
            nd = int((so != sd).sum())
            return "n=%d equal=%s ndiff=%d maxabs=%.3g" % (len(so), np.array_equal(so, sd), nd, float(np.abs(so - sd).max()))
        step("%s pass%d scores" % (sc.name, pi), scores)
        def match():
            ro = O.match(ogrid, g, sc.scan_pts, p, pose)
            pd = pose.copy(); cd = np.eye(3)
            ctx.reset_stats()
            t0 = time.time(); r = M.ScanMatch(dg, sc.scan_pts, p, pd, cd); dt = time.time() - t0
            st = ctx.stats(); d = M.last_detail
            ok = (r == ro["response"], np.array_equal(pd, ro["pose"]), np.allclose(cd, ro["cov"], rtol=1e-6, atol=0))
            return "resp %.6f/%.6f pose_eq=%s cov_ok=%s covbit=%s exact=%d navg=%d wall=%.2fms score_k=%.3fms sel_k=%.3fms evals=%.3g -> %.3g evals/s(kernel)" % (
                r, ro["response"], ok[1], ok[2], np.array_equal(cd, ro["cov"]), d.exact_sort_used, d.n_avg, dt * 1e3,
                st["score_kernel_ms"], st["select_kernel_ms"], st["evals"], st["evals"] / max(st["score_kernel_ms"], 1e-9) * 1e3)
        step("%s pass%d match" % (sc.name, pi), match)
        pose = O.match(ogrid, g, sc.scan_pts, p, pose)["pose"]
    if chain and len(sc.passes) == 3:
        def ch():
            ro = O.match_chain(ogrid, g, sc.scan_pts, sc.passes, sc.seed_pose)
            sm = matcher.ScanMatchers(ctx, sc.passes)
            pd = sc.seed_pose.copy(); cd = np.eye(3)
            t0 = time.time(); s = sm.ScanMatch(sc.scan_pts, dg, pd, cd); dt = time.time() - t0
            return "score_eq=%s pose_eq=%s cov_ok=%s resp_eq=%s wall=%.2fms" % (s == ro["score"], np.array_equal(pd, ro["pose"]),
                np.allclose(cd, ro["cov"], rtol=1e-6, atol=0), np.array_equal(sm.last_responses, ro["responses"]), dt * 1e3)
        step(sc.name + " chain", ch)
    dg.close(); ug.close()

check_scenario(synth.config1())
check_scenario(synth.config3(True), chain=True)
check_scenario(synth.config3(False), chain=True)
pairs = synth.config4(8)
check_scenario(pairs[0], chain=True)
check_scenario(synth.config2())

def batch():
    packed = matcher.pack_loop_closure(pairs)
    ctx.reset_stats()
    t0 = time.time(); scores, poses, covs, resp = matcher.loop_closure_batch(ctx, packed, pairs[0].passes); dt = time.time() - t0
    ok = 0
    for i, sc in enumerate(pairs):
        og = O.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        ro = O.match_chain(og, sc.grid, sc.scan_pts, sc.passes, sc.seed_pose)
        good = scores[i] == ro["score"] and np.array_equal(poses[i], ro["pose"]) and np.allclose(covs[i], ro["cov"], rtol=1e-6, atol=0)
        ok += int(good)
        if not good: print("   pair", i, scores[i], ro["score"], poses[i], ro["pose"])
    return "%d/%d pairs identical wall=%.2fms stats=%s" % (ok, len(pairs), dt * 1e3, ctx.stats())
step("loop_closure_batch(8)", batch)
print("done", flush=True)
