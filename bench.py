#!/usr/bin/env python
"""bench.py -- throughput of the correlative scan matcher hot path.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU matcher on the host cores

Metric (BASELINE.json): candidate-beam evaluations per second.  One evaluation = one
(angle, x, y) candidate x one visited beam (reference: the body of GetResponse's loop,
scan_match/correlate_scan_matcher.h:645-654).

Workload (BASELINE.json configs[1]): a 720-beam scan ray-cast from rm.pgm
matched over a +-1 m / +-45 deg window at 0.025 m (81 x 81 x 181 = 1 187 541 candidates x ~707
beams = 8.4e8 evaluations per step) against a 1120^2 lookup grid rasterised from 8 base scans.
A "step" is ONE such pass (BasedCorrelationScanMatch::ScanMatch) per GPU.  With N GPUs every rank
matches its own independent scan/seed (weak scaling, no data-path collective); `value` is the
whole-job aggregate: N * evals / max-over-ranks device time.

  value   grid and scan resident in HBM; CUDA events on the library's stream around each step;
          L2 flushed (256 MB streamed through) before every timed step.
  e2e     same step through the public host-buffer call (rsm_match): scan points come from
          pinned host memory every step, response/pose/covariance land in host memory.
  roofline  scoring kernel only: 4 algorithmic bytes per evaluation / its CUDA-event time, against
          the shared-memory row-gather bandwidth measured by the in-library micro-benchmark.
  cpu_baseline  the reference's own header (oracle/_ref, else the oracle port) on one host core.
  loop_closure  extra: batched loop-closure chains (BASELINE configs[3] shape) per second.

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate_beam_evals_per_s"
UNIT = "evals/s"
ALGO_BYTES_PER_EVAL = 4  # one float32 prob_value_ per evaluation


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=512,
                    help="loop-closure extra: pairs per GPU (BASELINE configs[3] is 4096 pairs over 8 GPUs = 512 each; 0 = skip)")
    ap.add_argument("--cpu-reps", type=int, default=4, help="full config-2 passes timed for cpu_baseline")
    ap.add_argument("--no-widened", action="store_true", help="skip the map-check / Gauss-Newton extras (N = 1 only)")
    ap.add_argument("--lc-contexts", type=int, default=0,
                    help="host threads / contexts per GPU for the loop-closure extra (0 = min(4, host cores / ranks - 1))")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the angle-sliced wide-window extra (BASELINE configs[4])")
    return ap.parse_args()


def rank_scenario(rank):
    """Config 2 with a rank-specific seed so that ranks match independent problems."""
    from roborts_edu_slam_b200 import synth
    sc = synth.config2()
    sc.seed_pose = sc.truth_pose + np.array([0.12 - 0.03 * rank, -0.07 + 0.02 * rank, 0.1 - 0.01 * rank])
    return sc


def workload_config(sc, geo):
    return {
        "workload": "BASELINE configs[1]: RPLidar-class 720-beam scan, +-1 m / +-45 deg window at 0.025 m over maps/rm.pgm",
        "step": "one BasedCorrelationScanMatch::ScanMatch pass per GPU",
        "candidates": geo["n_ang"] * geo["n_xy"] ** 2, "n_ang": geo["n_ang"], "n_xy": geo["n_xy"],
        "beams_visited": geo["visited"], "grid": "%dx%d @ %.3f m" % (sc.grid.size_x, sc.grid.size_y, sc.grid.res),
        "evals_per_step_per_gpu": geo["n_ang"] * geo["n_xy"] ** 2 * geo["visited"],
        "use_point_size": "all beams", "use_center_penalty": True,
    }


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed regions run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.bits, self.max_mhz, self.ok = [], 0, None, False
        self.active = threading.Event()
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.h = h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            if self.active.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        reasons = [n for b, n in self.REASONS.items() if self.bits & b]
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(s)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_handle():
    """(kind, callable(sc, param, pose) -> seconds for one pass) using the reference build if present."""
    from oracle.oracle_py import Oracle, Ref, ref_available
    if ref_available():
        R = Ref()

        def make(sc):
            m = R.create_map(sc.grid)
            R.build_map(m, sc.grid, sc.base_pts, sc.base_poses)
            return lambda param, pose: R.match(m, sc.scan_pts, param, pose)
        return "reference", make
    O = Oracle()

    def make(sc):
        grid = O.build_grid(sc.grid, sc.base_pts, sc.base_poses)
        return lambda param, pose: O.match(grid, sc.grid, sc.scan_pts, param, pose)
    return "port", make


def run_reference_arm(args):
    """The reference's CPU matcher on all host cores: one independent match per thread per step.
    Each step is a bounded sample of the workload: the full 81 x 81 translation window and all
    beams, but 23 of the 181 search angles (+-5.5 deg), so that K steps finish in minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle
    kind, make = cpu_reference_handle()
    threads = os.cpu_count() or 1
    scs = [rank_scenario(t % 8) for t in range(threads)]
    fns = [make(sc) for sc in scs]
    O = Oracle()
    params, evals = [], 0
    for sc in scs:
        p = sc.passes[0].copy()
        p[2] = 11 * p[3]          # aoff = 11 * ares -> 23 angles
        params.append(p)
        geo = O.geometry(sc.grid, p, len(sc.scan_pts), O.world_to_map(sc.grid, sc.seed_pose))
        evals += geo["n_ang"] * geo["n_xy"] ** 2 * geo["visited"]
    full_geo = O.geometry(scs[0].grid, scs[0].passes[0], len(scs[0].scan_pts), O.world_to_map(scs[0].grid, scs[0].seed_pose))

    def one_step():
        ts = [threading.Thread(target=fns[t], args=(params[t], scs[t].seed_pose)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(min(args.warmup, 2)):
        one_step()
    total = sum(one_step() for _ in range(args.steps))
    value = evals * args.steps / total
    sample = "per step and thread: config-2 pass restricted to 23 of 181 angles (81x81 translations, all beams), %d threads" % threads
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(scs[0], full_geo),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from roborts_edu_slam_b200 import matcher, synth
    from roborts_edu_slam_b200.sharding import contiguous_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the matcher has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # the library's host worker pool shares the box's cores with the other ranks
    if args.lc_contexts <= 0:
        # one host thread per context spins in the stream synchronisations: leave a core per rank for the rest
        # (measured at N = 8 on 32 cores: 2 / 3 / 4 / 6 contexts -> 690 k / 729 k / 611 k / 419 k matches/s)
        args.lc_contexts = max(1, min(4, (os.cpu_count() or 1) // max(1, world) - 1))
    os.environ.setdefault("RSM_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world) // max(1, args.lc_contexts))))
    ctx = matcher.Context(local_rank)
    sc = rank_scenario(rank)
    g = sc.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g)
    grid.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
    param = sc.passes[0]
    m = matcher.BasedCorrelationScanMatch(ctx)
    scan_dev = matcher.RangeDataContainer2d(ctx, sc.scan_pts)
    pinned = torch.empty((len(sc.scan_pts), 2), dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(np.ascontiguousarray(sc.scan_pts)))
    scan_host = pinned.numpy()

    # roofline denominators, measured on this GPU
    smem_row = ctx.microbench_gather(0, 160 * 1024)
    smem_rand = ctx.microbench_gather(1, 160 * 1024)
    glob_row = ctx.microbench_gather(2, g.size_x * g.size_y * 4)
    glob_rand = ctx.microbench_gather(3, g.size_x * g.size_y * 4)

    sampler = ClockSampler(local_rank)
    sampler.start()

    def one_step(scan):
        pose, cov = sc.seed_pose.copy(), np.eye(3)
        if not args.no_flush:
            ctx.flush_l2()
        ctx.timer_start()
        m.ScanMatch(grid, scan, param, pose, cov)
        return ctx.timer_stop(), pose

    for _ in range(args.warmup):
        one_step(scan_dev)
    det = m.last_detail
    evals_step = det.n_candidates * det.visited

    # ---- value: inputs resident ---------------------------------------------------------------
    import gc
    ctx.set_profiling(True)
    for _ in range(2):
        one_step(scan_dev)     # the profiled path has its own first-use costs (event pool)
    barrier()
    ctx.reset_stats()
    gc.collect()
    gc.disable()               # a collection inside the timed region shows up as an outlier of a 0.25 ms step
    sampler.active.set()
    ms_value = 0.0
    for _ in range(args.steps):
        ms_value += one_step(scan_dev)[0]
    sampler.active.clear()
    gc.enable()
    barrier()
    st_value = ctx.stats()
    t_value = max_over_ranks(ms_value)
    evals_all = sum_over_ranks(evals_step * args.steps)
    value = evals_all / (t_value * 1e-3)
    ctx.set_profiling(False)

    # ---- e2e: host buffers in, host results out -------------------------------------------------
    for _ in range(max(min(args.warmup, 5), 3)):
        one_step(scan_host)
    barrier()
    ctx.reset_stats()
    gc.collect()
    gc.disable()          # a collection inside the timed region shows up as a 30 % outlier of a 0.25 ms step
    sampler.active.set()
    ms_e2e = 0.0
    for _ in range(args.steps):
        ms_e2e += one_step(scan_host)[0]
    sampler.active.clear()
    gc.enable()
    barrier()
    st_e2e = ctx.stats()
    t_e2e = max_over_ranks(ms_e2e)
    e2e_value = evals_all / (t_e2e * 1e-3)

    # ---- loop-closure extra (config 4 shape): batched chains with device-side grid construction ----
    loop = None
    if args.pairs_per_gpu > 0:
        # The pair list of this rank is cut over a few host threads, each with its own context
        # (the library's rule: one context per caller thread), so that one thread's host->device
        # copies and host finalisation overlap another thread's kernels on the same GPU.
        b, e = contiguous_range(args.pairs_per_gpu * world, rank, world)
        pairs = synth.config4(e - b, first=b)
        nctx = max(1, min(args.lc_contexts, e - b))
        parts = []
        for t in range(nctx):
            pb, pe = contiguous_range(e - b, t, nctx)
            packed = matcher.pack_loop_closure(pairs[pb:pe])
            for key in ("base_pts", "pts", "base_poses", "centres", "poses"):   # inputs live in pinned host memory
                tt = torch.from_numpy(packed[key]).pin_memory()
                packed[key] = tt.numpy()
                packed["_pin_" + key] = tt
            parts.append((ctx if t == 0 else matcher.Context(local_rank), packed))
        # the same pairs with every scan already in a device-resident scan store (the back end adds each accepted
        # scan once; loop-closure candidates then name chains by id: rsm_scan_match_interface_batch)
        stores = []
        for t in range(nctx):
            pb, pe = contiguous_range(e - b, t, nctx)
            c = parts[t][0]
            st = matcher.ScanStore(c)
            chains, mids = [], []
            for sc in pairs[pb:pe]:
                chains.append([st.AddRangeData(p_, q_) for p_, q_ in zip(sc.base_pts, sc.base_poses)])
                mids.append(st.AddRangeData(sc.scan_pts, sc.seed_pose))
            stores.append((st, chains, mids, [sc.grid_centre for sc in pairs[pb:pe]], [sc.seed_pose for sc in pairs[pb:pe]]))
        results = [None] * nctx

        def lc_worker_host(t):
            c, packed = parts[t]
            results[t] = matcher.loop_closure_batch(c, packed, pairs[0].passes)

        def lc_worker_store(t):
            st, chains, mids, centres, seeds = stores[t]
            results[t] = matcher.scan_match_interface_batch(parts[t][0], st, pairs[0].grid, centres, chains, mids, seeds,
                                                            pairs[0].passes)

        def lc_measure(worker):
            def lc_run():
                ts = [threading.Thread(target=worker, args=(t,)) for t in range(nctx)]
                t0 = time.perf_counter()
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) * 1e3

            lc_run()   # warm-up
            barrier()
            for c, _ in parts:
                c.reset_stats()
            reps, ms_lc = 3, 0.0
            sampler.active.set()
            for _ in range(reps):
                ms_lc += lc_run()
            sampler.active.clear()
            barrier()
            st_lc = {k: sum(c.stats()[k] for c, _ in parts) for k in ("evals", "exact_sort_passes", "h2d_bytes", "d2h_bytes")}
            t_lc = max_over_ranks(ms_lc)
            n_matches = sum_over_ranks((e - b) * reps)
            accepted = sum(int((r[0] > 0.6).sum()) for r in results)
            return {
                "matches_per_s": n_matches / (t_lc * 1e-3), "ms_per_batch": t_lc / reps,
                "evals_per_s": sum_over_ranks(st_lc["evals"]) / (t_lc * 1e-3),
                "exact_sort_passes": int(sum_over_ranks(st_lc["exact_sort_passes"])),
                "accepted": int(sum_over_ranks(accepted)),
                "h2d_bytes_per_batch": st_lc["h2d_bytes"] / reps, "d2h_bytes_per_batch": st_lc["d2h_bytes"] / reps,
            }, [(r[0].copy(), r[1].copy()) for r in results]

        lc_host, res_host = lc_measure(lc_worker_host)
        lc_store, res_store = lc_measure(lc_worker_store)
        same = all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(res_host, res_store))
        loop = {
            "workload": "BASELINE configs[3] shape: 1081-beam scan vs 480^2 grid rasterised from 8 base scans, coarse/fine/super chain (YAML values)",
            "pairs": int(args.pairs_per_gpu * world), "contexts_per_gpu": nctx,
            "scans": "resident in a device scan store, chains named by id (rsm_scan_match_interface_batch)",
            **lc_store,
            "host_scans": dict(lc_host, scans="every base scan shipped from pinned host memory per call (rsm_loop_closure_batch)"),
            "store_equals_host_scans": bool(same),
            "timing": "host wall clock around the batched calls (host inputs -> host results), max over ranks",
        }
        if world == 1 and args.cpu_reps > 0:
            # the reference's own ScanMatchInterface step on one core: grid rebuild from the chain + coarse/fine/super chain
            from oracle.oracle_py import Oracle, Ref, ref_available
            n_cpu = 6
            if ref_available():
                R_ = Ref()
                t0 = time.perf_counter()
                for sc_ in pairs[:n_cpu]:
                    m_ = R_.create_map(sc_.grid)
                    R_.build_map(m_, sc_.grid, sc_.base_pts, sc_.base_poses)
                    R_.match_chain(m_, sc_.scan_pts, sc_.passes, sc_.seed_pose)
                    R_.destroy_map(m_)
                dtc, kind_ = time.perf_counter() - t0, "reference"
            else:
                O_ = Oracle()
                t0 = time.perf_counter()
                for sc_ in pairs[:n_cpu]:
                    O_.match_chain(O_.build_grid(sc_.grid, sc_.base_pts, sc_.base_poses), sc_.grid, sc_.scan_pts, sc_.passes, sc_.seed_pose)
                dtc, kind_ = time.perf_counter() - t0, "port"
            loop["cpu_baseline"] = {"value": n_cpu / dtc, "unit": "matches/s", "cores": 1, "kind": kind_,
                                    "sample": "%d of the pairs: grid rebuild from the 8 base scans + coarse/fine/super chain each" % n_cpu}
        for st, *_ in stores:
            st.close()
        for c, _ in parts[1:]:
            c.close()
    # ---- widened rows, N = 1 only: map check and Gauss-Newton matcher ------
    widened = None
    if world == 1 and not args.no_widened:
        widened = {}
        orc = None
        if args.cpu_reps > 0:
            from oracle.oracle_py import Oracle      # CPU baseline leg only (the checker, never the product path)
            orc = Oracle()
        # (1) MapCheckPenalize for a batch of candidate poses: 1081-beam scan, 100 check rays per pose
        zc = np.load(os.path.join(ROOT, "tests", "golden", "mapcheck_pair0.npz"), allow_pickle=False)
        gsp = zc["grid_spec"]
        gpub = synth.GridSpec(float(gsp[0]), 0.0, int(gsp[1]), int(gsp[2]), float(gsp[3]), float(gsp[4]), 0.5, 0.88, False)
        occ = np.unpackbits(zc["occ_packed"])[: gpub.size_x * gpub.size_y].reshape(gpub.size_y, gpub.size_x)
        pm = matcher.ScanMatchMap.from_spec(ctx, gpub)
        pm.upload_occupancy(occ)
        rng = np.random.default_rng(7)
        n_poses = 4096
        poses_mc = zc["poses"][0] + rng.uniform(-1.0, 1.0, (n_poses, 3)) * np.array([1.5, 1.5, 0.5])
        scan_mc = zc["scan_pts"]
        pm.MapCheckPenalize(scan_mc, poses_mc, 100, 2.5, 0.015, True)
        t0 = time.perf_counter()
        for _ in range(5):
            coeff = pm.MapCheckPenalize(scan_mc, poses_mc, 100, 2.5, 0.015, True)
        dt = (time.perf_counter() - t0) / 5
        entry = {"workload": "MapCheckPenalize, %d candidate poses x one %d-point scan, check_point_num 100, logistic (loop-closure form)" % (n_poses, len(scan_mc)),
                 "poses_per_s": n_poses / dt, "ms_per_batch": dt * 1e3, "timing": "host wall clock, host poses in -> host coefficients out"}
        if args.cpu_reps > 0:
            n_cpu = 256
            t0 = time.perf_counter()
            want = np.array([orc.map_check_penalize(occ, gpub, scan_mc, p_, 100, 2.5, 0.015, True) for p_ in poses_mc[:n_cpu]])
            dtc = time.perf_counter() - t0
            entry["cpu_baseline"] = {"value": n_cpu / dtc, "unit": "poses/s", "cores": 1, "kind": "port", "sample": "%d of the poses" % n_cpu}
            entry["equals_cpu"] = bool(np.array_equal(want, coeff[:n_cpu]))
        widened["map_check"] = entry
        pm.close()
        # (2) BasedOptimizeScanMatch for a batch of problems (config 4 pairs, yaml knobs)
        n_opt = 128
        pairs_o = synth.config4(n_opt)
        grids_o = []
        for sc_ in pairs_o:
            dg_ = matcher.ScanMatchMap.from_spec(ctx, sc_.grid)
            dg_.InitMapWithRangeVec(sc_.base_pts, sc_.base_poses, sc_.grid.default_prob, sc_.grid.sigma, sc_.grid.occu_offset, True)
            grids_o.append(dg_)
        opt = matcher.BasedOptimizeScanMatch(ctx)
        knobs = (10, 0.1, 0.5, 0.5, 0.5)
        scans_o = [sc_.scan_pts for sc_ in pairs_o]
        seeds_o = np.array([sc_.seed_pose for sc_ in pairs_o])
        opt.ScanMatchBatch(grids_o, scans_o, knobs, seeds_o)
        ctx.reset_stats()
        t0 = time.perf_counter()
        for _ in range(3):
            costs_o, poses_o, iters_o = opt.ScanMatchBatch(grids_o, scans_o, knobs, seeds_o)
        dt = (time.perf_counter() - t0) / 3
        entry = {"workload": "BasedOptimizeScanMatch, %d problems (config-4 pairs, 1081-beam scans, 480^2 grids), yaml knobs" % n_opt,
                 "problems_per_s": n_opt / dt, "ms_per_batch": dt * 1e3, "mean_iterations": float(iters_o.mean()),
                 "launches_per_batch": ctx.stats()["kernel_launches"] / 3,
                 "timing": "host wall clock, host scans + seeds in -> host poses + costs out"}
        if args.cpu_reps > 0:
            n_cpu = 16
            cpu_grids = [orc.build_grid(sc_.grid, sc_.base_pts, sc_.base_poses) for sc_ in pairs_o[:n_cpu]]
            t0 = time.perf_counter()
            want = [orc.optimize(cpu_grids[i], pairs_o[i].grid, scans_o[i], knobs, seeds_o[i]) for i in range(n_cpu)]
            dtc = time.perf_counter() - t0
            entry["cpu_baseline"] = {"value": n_cpu / dtc, "unit": "problems/s", "cores": 1, "kind": "port", "sample": "%d of the problems" % n_cpu}
            entry["equals_cpu"] = bool(all(want[i]["cost"] == costs_o[i] and np.array_equal(want[i]["pose"], poses_o[i]) for i in range(n_cpu)))
        widened["optimize"] = entry
        for dg_ in grids_o:
            dg_.close()
        # (3) front-end map maintenance: one accepted scan into the fine scan-match map (0.01 m, 2400^2, 5x5 stamps)
        #     and into the publishing map (0.05 m, ray-traced free space), maps resident on the device
        occ_w = synth.load_map("willow")
        tr = [np.array([14.375 + 0.05 * k, 28.625 + 0.02 * k, 0.3 + 0.01 * k]) for k in range(24)]
        scans_m = [synth.raycast(occ_w, p_[0], p_[1], p_[2], 1081, np.deg2rad(270.25), 10.0) for p_ in tr]
        gfine = synth.backend_grid(0.01, 0.03, 10.0, tr[0][:2])
        gpubm = synth.backend_grid(0.05, 0.15, 10.0, tr[0][:2])
        fine = matcher.ScanMatchMap.from_spec(ctx, gfine)
        fine.fill(0.5, 0.3)
        pubm = matcher.PubMap(ctx, gpubm.res, gpubm.size_x, gpubm.size_y, gpubm.off_x, gpubm.off_y)
        fine_pts = [s_ * (1 / 0.01) for s_ in scans_m]
        pub_pts = [s_ * (1 / 0.05) for s_ in scans_m]
        for k in range(4):
            fine.UpdateMapByRange(fine_pts[k], tr[k], gfine.sigma, gfine.occu_offset, True)
            pubm.UpdateMapByRange(pub_pts[k], tr[k], 0.0, 0.0)
        t0 = time.perf_counter()
        for k in range(4, 24):
            fine.UpdateMapByRange(fine_pts[k], tr[k], gfine.sigma, gfine.occu_offset, True)
            pubm.UpdateMapByRange(pub_pts[k], tr[k], 0.0, 0.0)
        pubm.refresh_occupancy(0.2, 4.0)
        dt = (time.perf_counter() - t0) / 20
        host_copy = fine.download()
        t0 = time.perf_counter()
        for _ in range(3):
            fine.upload(host_copy)
        dtu = (time.perf_counter() - t0) / 3
        entry = {"workload": "per accepted scan (1081 beams): UpdateMapByRange on the fine scan-match map (0.01 m, 2400^2, blur) "
                             "and on the publishing map (0.05 m, 480^2, ray-traced free space), both resident on the device",
                 "scans_per_s": 1.0 / dt, "ms_per_scan": dt * 1e3,
                 "whole_map_upload_ms": dtu * 1e3,
                 "note": "whole_map_upload_ms is what an adapter that keeps the maps on the host pays per scan instead (rsm_grid_upload_f32 of the fine map)",
                 "timing": "host wall clock, host scan in"}
        if args.cpu_reps > 0:
            from oracle.oracle_py import Ref, ref_available
            if ref_available():
                R_ = Ref()
                mfine = R_.frontend_map_create(gfine, 0.2)
                mpub = R_.pubmap_create_frontend(gpubm, 0.2)
                R_.pubmap_set_factors(mpub, 0.0, 0.0, 0.2, 4.0)
                for k in range(4):
                    R_.frontend_map_update(mfine, fine_pts[k], tr[k], True)
                    R_.pubmap_update_geom(mpub, pub_pts[k], tr[k])
                t0 = time.perf_counter()
                for k in range(4, 24):
                    R_.frontend_map_update(mfine, fine_pts[k], tr[k], True)
                    R_.pubmap_update_geom(mpub, pub_pts[k], tr[k])
                dtc = (time.perf_counter() - t0) / 20
                same_fine = bool(np.array_equal(R_.read_map_sized(mfine, gfine.size_x, gfine.size_y), host_copy))
                val_r = R_.pubmap_read_all(mpub, gpubm.size_x, gpubm.size_y)[0]
                same_pub = bool(np.array_equal(val_r, pubm.download_all()[0]))
                entry["cpu_baseline"] = {"value": 1.0 / dtc, "unit": "scans/s", "cores": 1, "kind": "reference",
                                         "sample": "the same 20 scans through the reference's own OccuGridMap::UpdateMapByRange (both maps)"}
                entry["equals_cpu"] = same_fine and same_pub
                R_.destroy_map(mfine)
                R_.pubmap_destroy(mpub)
        widened["frontend_update"] = entry
        fine.close()
        pubm.close()
    # ---- the other single-match configurations of BASELINE.json (N = 1 only): latency of one call, host in -> host out --
    small = None
    if world == 1 and not args.no_widened:
        small = {}
        ref_ok = False
        if args.cpu_reps > 0:
            from oracle.oracle_py import Oracle, Ref, ref_available
            ref_ok = ref_available()
            cpu_ = Ref() if ref_ok else Oracle()
        for tag, scx, label in (("configs0", synth.config1(), "BASELINE configs[0]: 360-beam scan vs icra submap, +-0.3 m / +-20 deg, 0.05 m, one coarse pass"),
                                ("configs2", synth.config3(), "BASELINE configs[2]: Hokuyo 1081-beam, coarse 0.1 m + fine + super chain with covariance on the 0.01 m map (2400^2), all beams"),
                                ("configs2_shipped", synth.config3(shipped_points=True), "the same chain with the shipped use_point_size 100/100/200")):
            gx = scx.grid
            dgx = matcher.ScanMatchMap.from_spec(ctx, gx)
            dgx.InitMapWithRangeVec(scx.base_pts, scx.base_poses, gx.default_prob, gx.sigma, gx.occu_offset, gx.use_blur)
            chain = len(scx.passes) == 3
            smx = matcher.ScanMatchers(ctx, scx.passes) if chain else None

            def call():
                pose_, cov_ = scx.seed_pose.copy(), np.eye(3)
                r_ = smx.ScanMatch(scx.scan_pts, dgx, pose_, cov_) if chain else m.ScanMatch(dgx, scx.scan_pts, scx.passes[0], pose_, cov_)
                return r_, pose_, cov_
            for _ in range(5):
                call()
            ctx.reset_stats()
            n_rep = 50
            t0 = time.perf_counter()
            for _ in range(n_rep):
                got = call()
            dt = (time.perf_counter() - t0) / n_rep
            stx = ctx.stats()
            entry = {"workload": label, "ms_per_match": dt * 1e3, "evals_per_match": stx["evals"] / n_rep,
                     "evals_per_s": stx["evals"] / n_rep / dt, "kernel_launches_per_match": stx["kernel_launches"] / n_rep,
                     "timing": "host wall clock per call, host scan in -> host pose / covariance out, grid resident"}
            if args.cpu_reps > 0:
                if ref_ok:
                    mref = cpu_.create_map(gx)
                    cpu_.build_map(mref, gx, scx.base_pts, scx.base_poses)
                    fn_ = (lambda: cpu_.match_chain(mref, scx.scan_pts, scx.passes, scx.seed_pose)) if chain else \
                          (lambda: cpu_.match(mref, scx.scan_pts, scx.passes[0], scx.seed_pose))
                else:
                    gcpu = cpu_.build_grid(gx, scx.base_pts, scx.base_poses)
                    fn_ = (lambda: cpu_.match_chain(gcpu, gx, scx.scan_pts, scx.passes, scx.seed_pose)) if chain else \
                          (lambda: cpu_.match(gcpu, gx, scx.scan_pts, scx.passes[0], scx.seed_pose))
                fn_()
                t0 = time.perf_counter()
                for _ in range(3):
                    w_ = fn_()
                dtc = (time.perf_counter() - t0) / 3
                entry["cpu_baseline"] = {"value": 1e3 * dtc, "unit": "ms per match", "cores": 1, "kind": "reference" if ref_ok else "port",
                                         "sample": "3 matches, grid built beforehand"}
                entry["equals_cpu"] = bool((got[0] == (w_["score"] if chain else w_["response"])) and np.array_equal(got[1], w_["pose"])
                                           and np.allclose(got[2], w_["cov"], rtol=1e-6, atol=0.0))
                if ref_ok:
                    cpu_.destroy_map(mref)
            small[tag] = entry
            dgx.close()
    # ---- wide relocalisation extra (config 5): ONE window angle-sliced over the ranks -----------
    wide = None
    if not args.no_wide:
        sc5 = synth.config5()
        g5 = sc5.grid
        grid5 = matcher.ScanMatchMap.from_spec(ctx, g5)
        grid5.InitMapWithRangeVec(sc5.base_pts, sc5.base_poses, g5.default_prob, g5.sigma, g5.occu_offset, g5.use_blur)

        def all_gather_bytes(buf):
            if world == 1:
                return [buf]
            t = torch.from_numpy(buf).cuda()
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return [o.cpu().numpy() for o in outs]

        sm5 = matcher.SlicedScanMatch(ctx, rank, world, all_gather_bytes)
        p5 = sc5.passes[0]
        res5 = None
        for _ in range(2):
            pose5, cov5 = sc5.seed_pose.copy(), np.eye(3)
            sm5.ScanMatch(grid5, sc5.scan_pts, p5, pose5, cov5)
        barrier()
        ctx.reset_stats()
        reps5, ms5 = 3, 0.0
        sampler.active.set()
        for _ in range(reps5):
            pose5, cov5 = sc5.seed_pose.copy(), np.eye(3)
            barrier()
            t0 = time.perf_counter()
            res5 = sm5.ScanMatch(grid5, sc5.scan_pts, p5, pose5, cov5)
            torch.cuda.synchronize()
            ms5 += (time.perf_counter() - t0) * 1e3
        sampler.active.clear()
        barrier()
        d5 = sm5.last_detail
        t5 = max_over_ranks(ms5)
        evals5 = float(d5.n_candidates) * d5.visited
        wide = {
            "workload": "BASELINE configs[4]: +-8 m / 360 deg window at 0.05 m on the full willow map, angle-sliced over the ranks "
                        "(partial -> all-gather -> merge -> all-gather -> finish)",
            "candidates": int(d5.n_candidates), "beams_visited": int(d5.visited), "evals_per_match": evals5,
            "ms_per_match": t5 / reps5, "evals_per_s": evals5 * reps5 / (t5 * 1e-3), "scaling": "strong",
            "timing": "host wall clock around the whole sliced call, barrier before, max over ranks",
            "response": res5, "pose_error_m": float(np.hypot(pose5[0] - sc5.truth_pose[0], pose5[1] - sc5.truth_pose[1])),
        }
        grid5.close()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- cpu baseline (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if world == 1 and args.cpu_reps > 0:
        kind, make = cpu_reference_handle()
        fn = make(sc)
        t0 = time.perf_counter()
        for _ in range(args.cpu_reps):
            fn(param, sc.seed_pose)
        dt = time.perf_counter() - t0
        cpu = {"value": evals_step * args.cpu_reps / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "%d full config-2 passes (%.3g evaluations each), single thread as in the reference" % (args.cpu_reps, evals_step)}

    if rank == 0:
        score_launches = max(1, st_value["score_launches"])
        k_ms = st_value["score_kernel_ms"] / score_launches
        achieved = ALGO_BYTES_PER_EVAL * evals_step / (k_ms * 1e-3) / 1e9
        geo = {"n_ang": det.n_ang, "n_xy": det.n_xy, "visited": det.visited}
        cfg = workload_config(sc, geo)
        cfg["l2"] = "flushed before every timed step (256 MB streamed)" if not args.no_flush else "not flushed"
        cfg["parallelism"] = "one independent match per GPU, no data-path collective"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "score_kernel_dram_bytes_per_launch.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": cfg,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": t_e2e / args.steps,
                    "h2d_bytes_per_step": st_e2e["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] / args.steps},
            "gpu_launches": int(st_value["kernel_launches"]),
            "roofline": {"bound": "smem", "kernel": "staged::score_staged_kernel<3,6> (all score launches of a step: 148 unsplit CTAs + 132 CTAs as clusters of 4)",
                         "achieved": achieved, "peak": smem_row, "unit": "GB/s",
                         "frac": achieved / smem_row, "traffic": traffic,
                         "kernel_ms": k_ms, "evals_per_s_kernel": evals_step / (k_ms * 1e-3),
                         "peak_source": "rsm_microbench_gather mode 0 (shared-memory row segments) measured in this run",
                         "other_peaks_gbs": {"smem_random": smem_rand, "global_row_l1l2": glob_row, "global_random_l1l2": glob_rand},
                         "frac_of_global_row": achieved / glob_row,
                         "hbm": {"peak": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                                 "achieved": (traffic / (k_ms * 1e-3) / 1e9) if traffic else None}},
            "kernel_share_of_step": {"score_ms": k_ms, "select_ms": st_value["select_kernel_ms"] / score_launches,
                                     "step_ms": t_value / args.steps},
            "exact_sort_passes": int(st_value["exact_sort_passes"]),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if loop:
            line["loop_closure"] = loop
        if widened is not None:
            line["widened"] = widened
        if small is not None:
            line["other_configs"] = small
        if wide:
            line["wide_window"] = wide
        print(json.dumps(line), flush=True)
    grid.close()
    scan_dev.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
