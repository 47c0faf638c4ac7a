"""Per-scan time of the front-end sequence through the C++ drop-in adapter on live reference objects (needs oracle/_ref)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth
from oracle.oracle_py import DropIn, Ref
occ_w = synth.load_map("willow")
tr = [np.array([14.375 + 0.05 * k, 28.625 + 0.02 * k, 0.3 + 0.01 * k]) for k in range(24)]
scans_m = [synth.raycast(occ_w, p_[0], p_[1], p_[2], 1081, np.deg2rad(270.25), 10.0) for p_ in tr]
gfine = synth.backend_grid(0.01, 0.03, 10.0, tr[0][:2])
fine_pts = [s_ * (1 / 0.01) for s_ in scans_m]
print("points per scan", sorted(set(len(p) for p in fine_pts)))
R_, D_ = Ref(), DropIn(0)
passes_f = synth.chain_defaults((100, 100, 200))
ma = R_.frontend_map_create(gfine, 0.2)
R_.frontend_map_update(ma, fine_pts[0], tr[0], True)
seeds_f = [tr[k] + np.array([0.03, -0.02, 0.015]) for k in range(24)]
for rep in range(2):
    ts = []
    for k in range(1, 24):
        t0 = time.perf_counter()
        g_ = D_.match_chain(ma, fine_pts[k], passes_f, seeds_f[k])
        t1 = time.perf_counter()
        D_.update_map(ma, fine_pts[k], g_["pose"], True, gfine.sigma, gfine.occu_offset)
        ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    print("rep", rep, "chain ms", [round(a, 2) for a, _ in ts], "\n   update ms", [round(b, 2) for _, b in ts], D_.sync_counts())
