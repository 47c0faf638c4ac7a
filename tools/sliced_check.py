"""Multi-GPU check of the in-library angle-sliced match (rsm_comm_init / rsm_match_sliced over NCCL): run as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sliced_check.py

Every rank matches the same windows (config 1 chain, the tie-heavy fixture, a quarter-scale config 5) sliced over the
N ranks and compares its result with the CPU oracle; exits non-zero on any difference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from roborts_edu_slam_b200 import matcher, synth  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402  (the checker)
from helpers import load_golden  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = matcher.Context(local)
    sm = matcher.make_sliced_matcher(ctx, rank, world, dist if world > 1 else None)
    orc = Oracle()
    fallbacks = 0
    cases = [synth.config1(), load_golden("ties_icra")[0], synth.config4(1, seed=5)[0], synth.config5(0.25)]
    for sc in cases:
        g = sc.grid
        grid = orc.build_grid(g, sc.base_pts, sc.base_poses)
        dg = matcher.ScanMatchMap.from_spec(ctx, g)
        dg.InitMapWithRangeVec(sc.base_pts, sc.base_poses, g.default_prob, g.sigma, g.occu_offset, g.use_blur)
        pose_in = sc.seed_pose.copy()
        for p in sc.passes:
            if sc.name.startswith("cfg5"):
                want = orc.finish_scores(orc.scores_threaded(grid, g, sc.scan_pts, p, orc.world_to_map(g, pose_in)), g, len(sc.scan_pts), p, pose_in)
            else:
                want = orc.match(grid, g, sc.scan_pts, p, pose_in)
            pose, cov = pose_in.copy(), np.eye(3)
            r = sm.ScanMatch(dg, sc.scan_pts, p, pose, cov)
            ok = r == want["response"] and np.array_equal(pose, want["pose"]) and np.allclose(cov, want["cov"], rtol=1e-6, atol=0.0) \
                and sm.last_detail.n_avg == want["n_avg"]
            if sm.exact_fallback:
                ok = ok and np.array_equal(cov, want["cov"])
                fallbacks += 1
            if not ok:
                print("rank %d: %s differs: %r %r vs %r %r" % (rank, sc.name, r, pose, want["response"], want["pose"]), flush=True)
                sys.exit(1)
            pose_in = want["pose"]
        dg.close()
    if fallbacks < 1:
        print("rank %d: the tie-heavy case did not take the exact path" % rank, flush=True)
        sys.exit(1)
    print("rank %d/%d: sliced matches equal the oracle (%d through the gathered exact path), exchange: %s" % (rank, world, fallbacks, sm.exchange), flush=True)
    sm.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
