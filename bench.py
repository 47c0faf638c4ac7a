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

  value   grid and scan resident in HBM; CUDA events on the library's stream around each step; inputs larger than L2:
          every step matches against the next of N_MAP_COPIES resident copies of the map (--l2 flush: the round-1
          protocol, 256 MB streamed through L2 before every step; reported as value_l2_flushed either way).
  e2e     same step through the public host-buffer call (rsm_match): scan points come from
          pinned host memory every step, response/pose/covariance land in host memory.
  roofline  scoring kernel only: 4 algorithmic bytes per evaluation / its CUDA-event time -- from a second timed pass over
          the same steps with the library's per-kernel events on -- against the shared-memory row-gather bandwidth
          measured by the in-library micro-benchmark.
  cpu_baseline  the reference's own header (oracle/_ref, else the oracle port) on one host core.
  loop_closure  the second headline metric: batched back-end steps (BASELINE configs[3] shape) per second,
          ONE context per GPU (the library pipelines sub-batches over its own streams), with its own roofline
          block and an all-core CPU figure.
  wide_window   BASELINE configs[4]: one window angle-sliced over the ranks, checked against the committed golden.

Rank 0 prints ONE compact JSON line (the loop_closure / wide_window blocks last, so that they survive a
truncated tail) and writes everything it measured, with the long descriptions, to bench_details_n<N>.json.
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
CONFIG2_GRID = "1120x1120 @ 0.025 m"


N_MAP_COPIES = 32      # x 5.0 MB of cells touched per match = 160 MB > the 126 MB L2


def compact(x):
    """Floats to 6 significant digits: the line stays short enough for a log tail without losing what is measured."""
    if isinstance(x, float):
        return float("%.6g" % x)
    if isinstance(x, dict):
        return {k: compact(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [compact(v) for v in x]
    return x


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=512,
                    help="loop-closure extra: pairs per GPU (BASELINE configs[3] is 4096 pairs over 8 GPUs = 512 each; 0 = skip)")
    ap.add_argument("--cpu-reps", type=int, default=8, help="full config-2 passes timed for cpu_baseline (0 = no CPU legs)")
    ap.add_argument("--no-widened", action="store_true", help="skip the map-check / Gauss-Newton / small-config extras (N = 1 only)")
    ap.add_argument("--lanes", type=int, default=0, help="RSM_OPT_LANES for the loop-closure extra (0 = the library's choice)")
    ap.add_argument("--no-flush", action="store_true", help="(with --l2 flush) do not flush")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush"],
                    help="how a step is kept from finding its inputs in L2: rotate over N_MAP_COPIES resident maps (default) or stream "
                         "256 MB through L2 before every step (the round-1 protocol; reported as value_l2_flushed either way)")
    ap.add_argument("--no-wide", action="store_true", help="skip the angle-sliced wide-window extra (BASELINE configs[4])")
    ap.add_argument("--details", default=None, help="where to write the detailed record (default bench_details_n<N>.json)")
    return ap.parse_args()


def rank_scenario(rank):
    """Config 2 with a rank-specific seed so that ranks match independent problems."""
    from roborts_edu_slam_b200 import synth
    sc = synth.config2()
    sc.seed_pose = sc.truth_pose + np.array([0.12 - 0.03 * rank, -0.07 + 0.02 * rank, 0.1 - 0.01 * rank])
    return sc


def workload_config(grid_spec, geo):
    """The `config` block: identical in the GPU arm and the reference arm (the driver compares them)."""
    grid = "%dx%d @ %.3f m" % (grid_spec.size_x, grid_spec.size_y, grid_spec.res)
    assert grid == CONFIG2_GRID, "the headline workload is BASELINE configs[1]; got a %s grid" % grid
    return {
        "workload": "BASELINE configs[1]: RPLidar-class 720-beam scan, +-1 m / +-45 deg window at 0.025 m over maps/rm.pgm",
        "step": "one BasedCorrelationScanMatch::ScanMatch pass per GPU",
        "candidates": geo["n_ang"] * geo["n_xy"] ** 2, "n_ang": geo["n_ang"], "n_xy": geo["n_xy"],
        "beams_visited": geo["visited"], "grid": grid,
        "evals_per_step_per_gpu": geo["n_ang"] * geo["n_xy"] ** 2 * geo["visited"],
        "use_point_size": "all beams", "use_center_penalty": True,
        "l2": "GPU arm: inputs larger than L2 -- every step matches against the next of %d resident copies of the map (%d MB of cells "
              "read per cycle, L2 is 126 MB), no copy is matched twice within 126 MB of other traffic; CPU arm: n/a" % (
                  N_MAP_COPIES, N_MAP_COPIES * grid_spec.size_x * grid_spec.size_y * 4 // 2 ** 20),
        "parallelism": "one independent match per GPU (per host thread in the CPU arm), no data-path collective",
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
    """(kind, make(scenario) -> callable(param, pose) running one pass) using the reference build if present."""
    from oracle.oracle_py import Oracle, Ref, ref_available
    if ref_available():
        R = Ref()

        def make(scn):
            m = R.create_map(scn.grid)
            R.build_map(m, scn.grid, scn.base_pts, scn.base_poses)
            return lambda param, pose: R.match(m, scn.scan_pts, param, pose)
        return "reference", make
    O = Oracle()

    def make(scn):
        grid = O.build_grid(scn.grid, scn.base_pts, scn.base_poses)
        return lambda param, pose: O.match(grid, scn.grid, scn.scan_pts, param, pose)
    return "port", make


def cpu_chain_workers():
    """(kind, fn(scenario)) running the reference's whole back-end step for one pair: grid rebuild from the chain's base
    scans + coarse / fine / super chain (slam_processor.cpp:250-326)."""
    from oracle.oracle_py import Oracle, Ref, ref_available
    if ref_available():
        R = Ref()

        def one(scn):
            m = R.create_map(scn.grid)
            R.build_map(m, scn.grid, scn.base_pts, scn.base_poses)
            R.match_chain(m, scn.scan_pts, scn.passes, scn.seed_pose)
            R.destroy_map(m)
        return "reference", one
    O = Oracle()
    return "port", lambda scn: O.match_chain(O.build_grid(scn.grid, scn.base_pts, scn.base_poses), scn.grid, scn.scan_pts,
                                             scn.passes, scn.seed_pose)


def run_threads(fn, n_threads):
    ts = [threading.Thread(target=fn, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def run_reference_arm(args):
    """The reference's CPU matcher on all host cores: every step, every thread runs ONE FULL config-2 pass (all 181
    angles, 81 x 81 translations, all beams) on its own independent scan / seed -- the same step the GPU arm times."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle
    kind, make = cpu_reference_handle()
    threads = os.cpu_count() or 1
    scs = [rank_scenario(t % 8) for t in range(min(threads, 8))]
    fns = [make(scn) for scn in scs]
    O = Oracle()
    geo = O.geometry(scs[0].grid, scs[0].passes[0], len(scs[0].scan_pts), O.world_to_map(scs[0].grid, scs[0].seed_pose))
    evals_pass = geo["n_ang"] * geo["n_xy"] ** 2 * geo["visited"]

    def work(t):
        scn = scs[t % len(scs)]
        fns[t % len(scs)](scn.passes[0], scn.seed_pose)

    for _ in range(min(args.warmup, 1)):
        run_threads(work, threads)
    total = sum(run_threads(work, threads) for _ in range(args.steps))
    value = evals_pass * threads * args.steps / total
    sample = "per step: %d host threads x one full config-2 pass each (181 angles x 81 x 81 translations x %d beams)" % (threads, geo["visited"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(scs[0].grid, geo),
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

    import gc
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
    cores = os.cpu_count() or 1
    os.environ.setdefault("RSM_HOST_THREADS", str(max(1, cores // max(1, world))))
    ctx = matcher.Context(local_rank)
    sc2 = rank_scenario(rank)           # the headline scenario; nothing below rebinds it
    g2 = sc2.grid
    grid = matcher.ScanMatchMap.from_spec(ctx, g2)
    grid.InitMapWithRangeVec(sc2.base_pts, sc2.base_poses, g2.default_prob, g2.sigma, g2.occu_offset, g2.use_blur)
    # the same map N_MAP_COPIES times in device memory: a step takes the next copy, so its cells come from HBM, not from L2
    grids = [grid]
    for _ in range(N_MAP_COPIES - 1):
        gcopy = matcher.ScanMatchMap.from_spec(ctx, g2)
        gcopy.InitMapWithRangeVec(sc2.base_pts, sc2.base_poses, g2.default_prob, g2.sigma, g2.occu_offset, g2.use_blur)
        grids.append(gcopy)
    step_no = [0]
    param2 = sc2.passes[0]
    m = matcher.BasedCorrelationScanMatch(ctx)
    scan_dev = matcher.RangeDataContainer2d(ctx, sc2.scan_pts)
    pinned = torch.empty((len(sc2.scan_pts), 2), dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(np.ascontiguousarray(sc2.scan_pts)))
    scan_host = pinned.numpy()

    # roofline denominators, measured on this GPU
    smem_row = ctx.microbench_gather(0, 160 * 1024)
    smem_rand = ctx.microbench_gather(1, 160 * 1024)
    glob_row = ctx.microbench_gather(2, g2.size_x * g2.size_y * 4)
    glob_rand = ctx.microbench_gather(3, g2.size_x * g2.size_y * 4)
    glob_row_lc = ctx.microbench_gather(2, 480 * 544 * 4)      # a loop-closure grid's footprint (L1/L2-resident row gathers)

    sampler = ClockSampler(local_rank)
    sampler.start()
    details = {}

    def one_step(scan, protocol=None):
        pose, cov = sc2.seed_pose.copy(), np.eye(3)
        protocol = protocol or args.l2
        if protocol == "flush":
            if not args.no_flush:
                ctx.flush_l2()
            target = grid
        else:
            step_no[0] += 1
            target = grids[step_no[0] % N_MAP_COPIES]
        ctx.timer_start()
        m.ScanMatch(target, scan, param2, pose, cov)
        return ctx.timer_stop(), pose

    for _ in range(max(args.warmup, N_MAP_COPIES if args.l2 == "rotate" else 0)):    # every copy once: tensor maps, first touch
        one_step(scan_dev)
    det = m.last_detail
    evals_step = det.n_candidates * det.visited
    geo2 = {"n_ang": det.n_ang, "n_xy": det.n_xy, "visited": det.visited}

    # ---- value: inputs resident ---------------------------------------------------------------
    # Two timed passes over the same steps.  (1) the product path -- one CUDA graph per step -- gives `value`; (2) the same
    # steps with the library's per-kernel CUDA events switched on give the kernel durations of the roofline block: the
    # event brackets replace the graph by separate launches and cost ~15 us per step, which is instrumentation, not work.
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
    st_launch = ctx.stats()
    t_value = max_over_ranks(ms_value)
    evals_all = sum_over_ranks(evals_step * args.steps)
    value = evals_all / (t_value * 1e-3)
    ctx.set_profiling(True)
    for _ in range(2):
        one_step(scan_dev)     # the profiled path has its own first-use costs (event pool)
    barrier()
    ctx.reset_stats()
    gc.collect()
    gc.disable()
    sampler.active.set()
    ms_profiled = 0.0
    for _ in range(args.steps):
        ms_profiled += one_step(scan_dev)[0]
    sampler.active.clear()
    gc.enable()
    barrier()
    st_value = ctx.stats()
    ms_profiled = max_over_ranks(ms_profiled)
    ctx.set_profiling(False)
    # the round-1 protocol beside it (explicit flush; cold instruction and descriptor fetches included), untimed otherwise
    ms_flushed = 0.0
    for k in range(13):
        t_ = one_step(scan_dev, "flush")[0]
        if k >= 3:
            ms_flushed += t_ / 10
    ms_flushed = max_over_ranks(ms_flushed)

    # ---- e2e: host buffers in, host results out -------------------------------------------------
    for _ in range(max(min(args.warmup, 5), 3)):
        one_step(scan_host)
    barrier()
    ctx.reset_stats()
    gc.collect()
    gc.disable()
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

    # ---- loop closure (BASELINE configs[3] shape): batched back-end steps, ONE context per GPU -------------------
    loop = None
    if args.pairs_per_gpu > 0:
        b, e = contiguous_range(args.pairs_per_gpu * world, rank, world)
        pairs = synth.config4(e - b, first=b)
        n_lc = e - b
        if args.lanes:
            ctx.set_option(matcher.RSM_OPT_LANES, args.lanes)
        # (a) every scan already in a device-resident scan store (the back end adds each accepted scan once); loop-closure
        #     candidates name their chains by id: only ids, centres and seeds travel per call
        store = matcher.ScanStore(ctx)
        chains, mids = [], []
        for pr in pairs:
            chains.append([store.AddRangeData(p_, q_) for p_, q_ in zip(pr.base_pts, pr.base_poses)])
            mids.append(store.AddRangeData(pr.scan_pts, pr.seed_pose))
        packed_chains = matcher.pack_chains(chains)
        mids = np.array(mids, dtype=np.int32)
        centres = np.array([pr.grid_centre for pr in pairs])
        seeds = np.array([pr.seed_pose for pr in pairs])
        # (b) the same pairs with every base scan shipped from pinned host memory per call
        packed = matcher.pack_loop_closure(pairs)
        for key in ("base_pts", "pts", "base_poses", "centres", "poses"):
            tt = torch.from_numpy(packed[key]).pin_memory()
            packed[key] = tt.numpy()
            packed["_pin_" + key] = tt
        lc_passes = pairs[0].passes
        lc_grid = pairs[0].grid

        def call_store():
            return matcher.scan_match_interface_batch(ctx, store, lc_grid, centres, packed_chains, mids, seeds, lc_passes)

        def call_host():
            return matcher.loop_closure_batch(ctx, packed, lc_passes)

        def lc_measure(call, reps=20):
            for _ in range(2):
                call()
            barrier()
            ctx.reset_stats()
            ms_lc, res = 0.0, None
            sampler.active.set()
            for _ in range(reps):
                t0 = time.perf_counter()
                res = call()
                ms_lc += (time.perf_counter() - t0) * 1e3      # the call returns with its results on the host
            sampler.active.clear()
            barrier()
            st = ctx.stats()
            t_lc = max_over_ranks(ms_lc)
            return {
                "matches_per_s": sum_over_ranks(n_lc * reps) / (t_lc * 1e-3), "ms_per_batch": t_lc / reps,
                "evals_per_s": sum_over_ranks(st["evals"]) / (t_lc * 1e-3),
                "exact_sort_passes": int(sum_over_ranks(st["exact_sort_passes"])) // reps,
                "accepted": int(sum_over_ranks(int((res[0] > 0.6).sum()))),
                "h2d_bytes_per_batch": st["h2d_bytes"] / reps, "d2h_bytes_per_batch": st["d2h_bytes"] / reps,
                "kernel_launches_per_batch": st["kernel_launches"] / reps,
                "host_phase_ms_per_batch": [round(v / reps, 3) for v in st["phase_ms"][:6]],
            }, res

        lc_store, res_store = lc_measure(call_store)
        lc_host, res_host = lc_measure(call_host)
        same = bool(np.array_equal(res_store[0], res_host[0]) and np.array_equal(res_store[1], res_host[1]))
        # kernel times of one batch, one lane so that the event brackets do not overlap other sub-batches' kernels
        ctx.set_option(matcher.RSM_OPT_LANES, 1)
        ctx.set_profiling(True)
        call_store()
        ctx.reset_stats()
        call_store()
        st_k = ctx.stats()
        ctx.set_profiling(False)
        ctx.set_option(matcher.RSM_OPT_LANES, args.lanes)
        lc_roof = None
        if st_k["score_kernel_ms"] > 0:
            ach = ALGO_BYTES_PER_EVAL * st_k["evals"] / (st_k["score_kernel_ms"] * 1e-3) / 1e9
            lc_roof = {"bound": "smem", "kernels": "patch + flat (see details)",
                       "achieved": ach, "peak": smem_row, "unit": "GB/s", "frac": ach / smem_row,
                       "frac_of_global_row_l1l2": ach / glob_row_lc, "global_row_l1l2_peak": glob_row_lc,
                       "score_kernel_ms": st_k["score_kernel_ms"], "select_kernel_ms": st_k["select_kernel_ms"],
                       "raster_kernel_ms": st_k["raster_kernel_ms"], "evals_per_batch": st_k["evals"],
                       }
        loop = dict(lc_store, pairs=int(args.pairs_per_gpu * world), contexts_per_gpu=1,
                    lanes=args.lanes if args.lanes else "auto", store_equals_host_scans=same, roofline=lc_roof,
                    host_scans={k: lc_host[k] for k in ("matches_per_s", "ms_per_batch", "h2d_bytes_per_batch")})
        details["loop_closure"] = {
            "roofline": "score launches of the chain: patch (coarse, shared-memory tiles) + flat (fine, super-fine; L1-resident grid); "
                        "one lane, profiled batch; 4 B x evaluations / summed score-kernel time",
            "workload": "BASELINE configs[3] shape: 1081-beam scan vs 480^2 grid rasterised from 8 base scans, coarse/fine/super chain (YAML values)",
            "scans": "resident in a device scan store, chains named by id (rsm_scan_match_interface_batch); host_scans: every base scan shipped "
                     "from pinned host memory per call (rsm_loop_closure_batch)",
            "timing": "host wall clock around the batched call (host ids / seeds in -> host results out), max over ranks", "host_scans_full": lc_host}
        if world == 1 and args.cpu_reps > 0:
            # the reference's own ScanMatchInterface step, pair-parallel on every host core (one matcher + map per thread)
            kind_, one = cpu_chain_workers()
            per_thread = 3

            def work(t):
                for k in range(per_thread):
                    one(pairs[(t * per_thread + k) % n_lc])
            work(0)
            dt1 = run_threads(work, 1)
            dtc = run_threads(work, cores)
            loop["cpu_baseline"] = {"value": cores * per_thread / dtc, "unit": "matches/s", "cores": cores, "kind": kind_,
                                    "one_core": per_thread / dt1,
                                    "sample": "%d pairs per thread: grid rebuild from the 8 base scans + coarse/fine/super chain each" % per_thread}
        store.close()

    # ---- widened rows, N = 1 only ------------------------------------------------------------
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
        entry = {"poses_per_s": n_poses / dt, "ms_per_batch": dt * 1e3}
        if args.cpu_reps > 0:
            n_cpu = 256
            t0 = time.perf_counter()
            want = np.array([orc.map_check_penalize(occ, gpub, scan_mc, p_, 100, 2.5, 0.015, True) for p_ in poses_mc[:n_cpu]])
            dtc = time.perf_counter() - t0
            entry["cpu_poses_per_s_1core"] = n_cpu / dtc
            entry["equals_cpu"] = bool(np.array_equal(want, coeff[:n_cpu]))
        widened["map_check"] = entry
        pm.close()
        # (2) BasedOptimizeScanMatch for a batch of problems (config 4 pairs, yaml knobs)
        n_opt = 128
        pairs_o = synth.config4(n_opt)
        grids_o = []
        for pr in pairs_o:
            dg_ = matcher.ScanMatchMap.from_spec(ctx, pr.grid)
            dg_.InitMapWithRangeVec(pr.base_pts, pr.base_poses, pr.grid.default_prob, pr.grid.sigma, pr.grid.occu_offset, True)
            grids_o.append(dg_)
        opt = matcher.BasedOptimizeScanMatch(ctx)
        knobs = (10, 0.1, 0.5, 0.5, 0.5)
        scans_o = [pr.scan_pts for pr in pairs_o]
        seeds_o = np.array([pr.seed_pose for pr in pairs_o])
        opt.ScanMatchBatch(grids_o, scans_o, knobs, seeds_o)
        t0 = time.perf_counter()
        for _ in range(3):
            costs_o, poses_o, iters_o = opt.ScanMatchBatch(grids_o, scans_o, knobs, seeds_o)
        dt = (time.perf_counter() - t0) / 3
        entry = {"problems_per_s": n_opt / dt, "ms_per_batch": dt * 1e3, "mean_iterations": float(iters_o.mean())}
        if args.cpu_reps > 0:
            n_cpu = 16
            cpu_grids = [orc.build_grid(pr.grid, pr.base_pts, pr.base_poses) for pr in pairs_o[:n_cpu]]
            t0 = time.perf_counter()
            want = [orc.optimize(cpu_grids[i], pairs_o[i].grid, scans_o[i], knobs, seeds_o[i]) for i in range(n_cpu)]
            dtc = time.perf_counter() - t0
            entry["cpu_problems_per_s_1core"] = n_cpu / dtc
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
        entry = {"scans_per_s": 1.0 / dt, "ms_per_scan": dt * 1e3}
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
                entry["cpu_scans_per_s_1core"] = 1.0 / dtc
                entry["equals_cpu"] = same_fine and same_pub
                R_.destroy_map(mfine)
                R_.pubmap_destroy(mpub)
        widened["frontend_update"] = entry
        fine.close()
        pubm.close()
        # (4) the 4-line drop-in on LIVE reference objects (oracle/_ref/libdropin.so = the product's C++ adapter compiled
        #     against the unmodified reference headers): per accepted scan, the shipped front-end sequence -- coarse / fine /
        #     super chain on the fine map (0.01 m, 2400^2), then UpdateMapByRange at the matched pose -- through
        #     rsm_adapter::BasedCorrelationScanMatch + rsm_adapter::UpdateMapByRange, beside the reference's own classes
        if args.cpu_reps > 0:
            from oracle.oracle_py import DropIn, Ref, dropin_available, ref_available
            if dropin_available() and ref_available():
                R_, D_ = Ref(), DropIn(local_rank)
                passes_f = synth.chain_defaults((100, 100, 200))
                ma, mr = R_.frontend_map_create(gfine, 0.2), R_.frontend_map_create(gfine, 0.2)
                for m_ in (ma, mr):
                    R_.frontend_map_update(m_, fine_pts[0], tr[0], True)
                seeds_f = [tr[k] + np.array([0.03, -0.02, 0.015]) for k in range(24)]
                same_f = True

                def run_ref(k):
                    w_ = R_.match_chain(mr, fine_pts[k], passes_f, seeds_f[k])
                    R_.frontend_map_update(mr, fine_pts[k], w_["pose"], True)
                    return w_

                def run_adapter(k):
                    g_ = D_.match_chain(ma, fine_pts[k], passes_f, seeds_f[k])
                    D_.update_map(ma, fine_pts[k], g_["pose"], True, gfine.sigma, gfine.occu_offset)
                    return g_
                for k in range(1, 4):        # warm-up: first sync uploads the whole plane once
                    w_, g_ = run_ref(k), run_adapter(k)
                    same_f = same_f and g_["score"] == w_["score"] and np.array_equal(g_["pose"], w_["pose"])
                full0, inc0 = D_.sync_counts()
                t0 = time.perf_counter()
                outs_a = [run_adapter(k) for k in range(4, 24)]
                dta = (time.perf_counter() - t0) / 20
                t0 = time.perf_counter()
                outs_r = [run_ref(k) for k in range(4, 24)]
                dtr = (time.perf_counter() - t0) / 20
                full1, inc1 = D_.sync_counts()
                same_f = same_f and all(a_["score"] == b_["score"] and np.array_equal(a_["pose"], b_["pose"]) and
                                        np.allclose(a_["cov"], b_["cov"], rtol=1e-6, atol=0.0) for a_, b_ in zip(outs_a, outs_r))
                widened["dropin_frontend"] = {"ms_per_scan": dta * 1e3, "cpu_ms_per_scan_1core": dtr * 1e3, "speedup": dtr / dta,
                                              "whole_map_uploads": full1 - full0, "incremental_updates": inc1 - inc0,
                                              "mirror_equals_host": D_.mirror_equals_host(ma) == 1, "equals_cpu": bool(same_f)}
                R_.destroy_map(ma)
                R_.destroy_map(mr)
                D_.close()
        details["widened"] = {
            "map_check": "MapCheckPenalize, 4096 candidate poses x one 1081-point scan, check_point_num 100, logistic (loop-closure form); host wall clock",
            "optimize": "BasedOptimizeScanMatch, 128 problems (config-4 pairs, 1081-beam scans, 480^2 grids), yaml knobs; host wall clock",
            "frontend_update": "per accepted scan (1081 beams): UpdateMapByRange on the fine scan-match map (0.01 m, 2400^2, blur) and on the "
                               "publishing map (0.05 m, 480^2, ray-traced free space), both resident on the device; host wall clock, host scan in",
            "dropin_frontend": "per accepted scan on LIVE reference objects: coarse/fine/super chain (shipped use_point_size 100/100/200) on the "
                               "fine map (0.01 m, 2400^2) + UpdateMapByRange at the matched pose, through the C++ adapter "
                               "(csrc/scan_matcher_adapter.hpp compiled against the unmodified reference headers) vs the reference's own classes"}

    # ---- the other single-match configurations of BASELINE.json (N = 1 only): latency of one call, host in -> host out --
    small = None
    if world == 1 and not args.no_widened:
        small = {}
        ref_ok = False
        if args.cpu_reps > 0:
            from oracle.oracle_py import Oracle, Ref, ref_available
            ref_ok = ref_available()
            cpu_ = Ref() if ref_ok else Oracle()
        for tag, scx in (("configs0", synth.config1()), ("configs2", synth.config3()), ("configs2_shipped", synth.config3(shipped_points=True))):
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
            entry = {"ms_per_match": dt * 1e3, "evals_per_match": stx["evals"] / n_rep, "kernel_launches_per_match": stx["kernel_launches"] / n_rep}
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
                entry["cpu_ms_per_match_1core"] = 1e3 * dtc
                entry["cpu_kind"] = "reference" if ref_ok else "port"
                entry["equals_cpu"] = bool((got[0] == (w_["score"] if chain else w_["response"])) and np.array_equal(got[1], w_["pose"])
                                           and np.allclose(got[2], w_["cov"], rtol=1e-6, atol=0.0))
                if ref_ok:
                    cpu_.destroy_map(mref)
            small[tag] = entry
            dgx.close()
        details["other_configs"] = {
            "configs0": "BASELINE configs[0]: 360-beam scan vs icra submap, +-0.3 m / +-20 deg, 0.05 m, one coarse pass",
            "configs2": "BASELINE configs[2]: Hokuyo 1081-beam, coarse 0.1 m + fine + super chain with covariance on the 0.01 m map (2400^2), all beams",
            "configs2_shipped": "the same chain with the shipped use_point_size 100/100/200",
            "timing": "host wall clock per call, host scan in -> host pose / covariance out, grid resident"}

    # ---- wide relocalisation (config 5): ONE window angle-sliced over the ranks ---------------------------------------
    wide = None
    if not args.no_wide:
        sc5 = synth.config5()
        g5 = sc5.grid
        grid5 = matcher.ScanMatchMap.from_spec(ctx, g5)
        grid5.InitMapWithRangeVec(sc5.base_pts, sc5.base_poses, g5.default_prob, g5.sigma, g5.occu_offset, g5.use_blur)
        sm5 = matcher.make_sliced_matcher(ctx, rank, world, dist if world > 1 else None)
        p5 = sc5.passes[0]
        res5 = None
        for _ in range(4):          # (the in-library NCCL communicator settles over its first calls)
            pose5, cov5 = sc5.seed_pose.copy(), np.eye(3)
            sm5.ScanMatch(grid5, sc5.scan_pts, p5, pose5, cov5)
        barrier()
        ctx.reset_stats()
        reps5, ms5 = 10, 0.0
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
        equals_golden = None
        gpath = os.path.join(ROOT, "tests", "golden", "config5_full.npz")
        if os.path.exists(gpath):
            z5 = np.load(gpath, allow_pickle=False)
            equals_golden = bool(res5 == float(z5["response"]) and np.array_equal(pose5, z5["pose"])
                                 and np.allclose(cov5, z5["cov"], rtol=1e-6, atol=0.0) and d5.n_avg == int(z5["n_avg"]))
        wide = {"ms_per_match": t5 / reps5, "evals_per_s": evals5 * reps5 / (t5 * 1e-3), "scaling": "strong",
                "candidates": int(d5.n_candidates), "exchange": sm5.exchange, "equals_golden": equals_golden,
                "exact_fallback": bool(sm5.exact_fallback),
                "pose_error_m": float(np.hypot(pose5[0] - sc5.truth_pose[0], pose5[1] - sc5.truth_pose[1]))}
        details["wide_window"] = {
            "workload": "BASELINE configs[4]: +-8 m / 360 deg window at 0.05 m on the full willow map, angle-sliced over the ranks",
            "timing": "host wall clock around the whole sliced call, barrier before, max over ranks",
            "golden": "tests/golden/config5_full.npz (the reference's own code on the full window): response, pose, covariance, n_avg"}
        sm5.close()
        grid5.close()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- cpu baseline of the headline workload (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if world == 1 and args.cpu_reps > 0:
        kind, make = cpu_reference_handle()
        fn = make(sc2)
        t0 = time.perf_counter()
        for _ in range(args.cpu_reps):
            fn(param2, sc2.seed_pose)
        dt = time.perf_counter() - t0
        cpu = {"value": evals_step * args.cpu_reps / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "%d full config-2 passes (%.3g evaluations each), single thread as in the reference" % (args.cpu_reps, evals_step)}

    if rank == 0:
        score_launches = max(1, st_value["score_launches"])
        k_ms = st_value["score_kernel_ms"] / score_launches
        achieved = ALGO_BYTES_PER_EVAL * evals_step / (k_ms * 1e-3) / 1e9
        cfg = workload_config(g2, geo2)
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
            "gpu_launches": int(st_launch["kernel_launches"]),
            "roofline": {"bound": "smem", "kernel": "staged::score_stream_kernel<MapPaired> (the one persistent score launch of a step)",
                         "achieved": achieved, "peak": smem_row, "unit": "GB/s",
                         "frac": achieved / smem_row, "traffic": traffic,
                         "kernel_ms": k_ms, "select_ms": st_value["select_kernel_ms"] / score_launches,
                         "profiled_ms_per_step": ms_profiled / args.steps,
                         "peak_source": "rsm_microbench_gather mode 0, this run",
                         "hbm_peak": hbm_peak, "hbm_achieved": (traffic / (k_ms * 1e-3) / 1e9) if traffic else None},
            "exact_sort_passes": int(st_value["exact_sort_passes"]),
            "value_l2_flushed": {"value": evals_all / args.steps / (ms_flushed * 1e-3), "ms_per_step": ms_flushed,
                                 "protocol": "256 MB streamed through L2 before every step, as in round 1"},
        }
        details["roofline_other_peaks_gbs"] = {"smem_random": smem_rand, "global_row_l1l2": glob_row, "global_random_l1l2": glob_rand,
                                               "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
        if cpu:
            line["cpu_baseline"] = cpu
        if widened is not None:
            line["widened"] = widened
        if small is not None:
            line["other_configs"] = small
        # the two other headline workloads last, so that a truncated tail of the line still shows them
        if wide:
            line["wide_window"] = wide
        if loop:
            line["loop_closure"] = loop
        print(json.dumps(compact(line)), flush=True)
        try:
            dpath = args.details or os.path.join(os.getcwd(), "bench_details_n%d.json" % world)
            json.dump(dict(line, details=details), open(dpath, "w"), indent=1)
        except Exception:
            pass
    grid.close()
    scan_dev.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
