"""TEST INFRASTRUCTURE ONLY -- ctypes front ends for the two CPU checkers.

  Oracle : oracle/liboracle.so   the from-scratch restatement (rsm_oracle.cpp)
  Ref    : oracle/_ref/libref.so the reference's own headers compiled in place (ref_driver.cpp);
           present only where /root/reference was available at build time (it travels to the
           GPU box as a built file).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Both classes expose the same calls on the same flat numpy inputs so a test
can swap one for the other.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

c_d = ctypes.c_double
c_i = ctypes.c_int
c_l = ctypes.c_long
c_p = ctypes.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref.so"))


def pack_scans(pts_list):
    n_pts = np.array([len(p) for p in pts_list], dtype=np.int32)
    if len(pts_list):
        pts = _f64(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in pts_list], axis=0))
    else:
        pts = np.zeros((0, 2), dtype=np.float64)
    return n_pts, pts


class Oracle:
    """The restatement.  A 'map' here is (float32 grid [size_y, size_x], GridSpec)."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run __graft_entry__.build()")
        L = ctypes.CDLL(path)
        L.orc_blur_kernel.restype = c_i
        L.orc_blur_kernel.argtypes = [c_d, c_d, c_p, c_i]
        L.orc_grid_build.restype = c_i
        L.orc_grid_build.argtypes = [c_p, c_i, c_i, ctypes.c_float, c_d, c_d, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_i]
        L.orc_world_to_map.argtypes = [c_d, c_d, c_d, c_p, c_p]
        L.orc_map_to_world.argtypes = [c_d, c_d, c_d, c_p, c_p]
        L.orc_pass_geometry.argtypes = [c_p, c_i, c_d, c_p, c_p, c_p]
        L.orc_scores.restype = c_i
        L.orc_scores.argtypes = [c_p, c_i, c_d, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
        L.orc_scores_slice.restype = c_i
        L.orc_scores_slice.argtypes = [c_p, c_i, c_d, c_i, c_p, c_p, c_p, c_i, c_i, c_p]
        L.orc_finish_scores.restype = c_d
        L.orc_finish_scores.argtypes = [c_p, c_l, c_i, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p]
        L.orc_match.restype = c_d
        L.orc_match.argtypes = [c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
        L.orc_match_chain.restype = c_d
        L.orc_match_chain.argtypes = [c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p]
        L.orc_grid_stamp.restype = c_i
        L.orc_grid_stamp.argtypes = [c_p, c_i, c_i, ctypes.c_float, c_d, c_d, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_i, c_i]
        L.orc_grid_extend.argtypes = [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i, ctypes.c_float, ctypes.c_float]
        L.orc_pubmap_update.argtypes = [c_p, c_p, c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, ctypes.c_float,
                                        ctypes.c_float, ctypes.POINTER(c_i)]
        L.orc_optimize.restype = c_d
        L.orc_optimize.argtypes = [c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_p, c_p]
        L.orc_optimize_cost.argtypes = [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]
        L.orc_ldlt3_solve.argtypes = [c_p, c_p, c_p]
        L.orc_match_chain_opt.restype = c_d
        L.orc_match_chain_opt.argtypes = [c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p,
                                          c_p, c_p, c_d, c_i, c_p, c_p, c_p]
        L.orc_map_check_penalize.restype = c_d
        L.orc_map_check_penalize.argtypes = [c_p, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_i, c_d, c_d, c_i]
        self.L = L

    def map_check_penalize(self, occupied, g, pts_cells, pose_world, check_point_num, bound_tolerance, penalty_gain,
                           use_logistic=False, origin=(0.0, 0.0)):
        """MapFeedbackResponsePenalty (+ the loop-closure logistic) over an occupancy mask [size_y, size_x]."""
        occ = np.ascontiguousarray(occupied, dtype=np.uint8)
        pts = np.ascontiguousarray(pts_cells, dtype=np.float64)
        pose = np.ascontiguousarray(pose_world, dtype=np.float64)
        org = np.ascontiguousarray(origin, dtype=np.float64)
        return self.L.orc_map_check_penalize(occ.ctypes.data, g.size_x, g.size_y, 1.0 / g.res, g.off_x, g.off_y, len(pts),
                                             pts.ctypes.data, pose.ctypes.data, org.ctypes.data, int(check_point_num),
                                             float(bound_tolerance), float(penalty_gain), 1 if use_logistic else 0)

    def blur_kernel(self, sigma, res):
        k = np.zeros(21 * 21, dtype=np.float64)
        h = self.L.orc_blur_kernel(sigma, res, k.ctypes.data, k.size)
        if h < 0:
            return -1, None
        n = 2 * h + 1
        return h, k[: n * n].reshape(n, n).copy()

    def build_grid(self, g, base_pts, base_poses):
        grid = np.empty((g.size_y, g.size_x), dtype=np.float32)
        n_pts, pts = pack_scans(base_pts)
        poses = _f64(base_poses)
        rc = self.L.orc_grid_build(grid.ctypes.data, g.size_x, g.size_y, g.default_prob, g.sigma, g.res,
                                   g.occu_offset, g.off_x, g.off_y, len(base_pts), n_pts.ctypes.data,
                                   pts.ctypes.data, poses.ctypes.data, int(g.use_blur))
        if rc:
            raise RuntimeError("orc_grid_build rc=%d" % rc)
        return grid

    def world_to_map(self, g, w):
        out = np.zeros(3)
        w = _f64(w)
        self.L.orc_world_to_map(1.0 / g.res, g.off_x, g.off_y, w.ctypes.data, out.ctypes.data)
        return out

    def map_to_world(self, g, m):
        out = np.zeros(3)
        m = _f64(m)
        self.L.orc_map_to_world(1.0 / g.res, g.off_x, g.off_y, m.ctypes.data, out.ctypes.data)
        return out

    def geometry(self, g, param, P, center_map):
        out = np.zeros(5, dtype=np.int64)
        dout = np.zeros(4)
        cell_len = 1 / (1.0 / g.res)
        param, center_map = _f64(param), _f64(center_map)
        self.L.orc_pass_geometry(param.ctypes.data, P, cell_len, center_map.ctypes.data, out.ctypes.data, dout.ctypes.data)
        return dict(n_ang=int(out[0]), n_xy=int(out[1]), step=int(out[2]), divisor=int(out[3]), visited=int(out[4]),
                    start_x=dout[0], start_y=dout[1], factor=dout[2], start_angle=dout[3])

    def scores(self, grid, g, pts, param, center_map, dump_indices=False):
        pts, param, center_map = _f64(pts), _f64(param), _f64(center_map)
        geo = self.geometry(g, param, len(pts), center_map)
        n = geo["n_ang"] * geo["n_xy"] ** 2
        score = np.empty(n, dtype=np.float64)
        gx = gy = None
        if dump_indices:
            gx = np.empty((n, geo["visited"]), dtype=np.int32)
            gy = np.empty((n, geo["visited"]), dtype=np.int32)
        cell_len = 1 / (1.0 / g.res)
        self.L.orc_scores(grid.ctypes.data, g.size_x, cell_len, len(pts), pts.ctypes.data, param.ctypes.data,
                          center_map.ctypes.data, score.ctypes.data, _ptr(gx), _ptr(gy))
        return (score, gx, gy) if dump_indices else score

    def scores_threaded(self, grid, g, pts, param, center_map, threads=None):
        """The same bits as scores(), angle slices on `threads` host threads (ctypes releases the GIL)."""
        import threading
        pts, param, center_map = _f64(pts), _f64(param), _f64(center_map)
        geo = self.geometry(g, param, len(pts), center_map)
        plane = geo["n_xy"] ** 2
        score = np.empty(geo["n_ang"] * plane, dtype=np.float64)
        threads = max(1, min(threads or (os.cpu_count() or 1), geo["n_ang"]))
        cell_len = 1 / (1.0 / g.res)
        cuts = [geo["n_ang"] * t // threads for t in range(threads + 1)]
        rcs = [0] * threads

        def work(t):
            out = score[cuts[t] * plane: cuts[t + 1] * plane]
            rcs[t] = self.L.orc_scores_slice(grid.ctypes.data, g.size_x, cell_len, len(pts), pts.ctypes.data, param.ctypes.data,
                                             center_map.ctypes.data, cuts[t], cuts[t + 1], out.ctypes.data)
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not any(rcs), rcs
        return score

    def finish_scores(self, score, g, n_pts, param, pose_world, cov=None):
        """orc_match's tail on a score array computed by scores() / scores_threaded()."""
        score, param = _f64(score), _f64(param)
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        best = np.zeros(4)
        navg = c_l(0)
        r = self.L.orc_finish_scores(score.ctypes.data, len(score), int(n_pts), 1.0 / g.res, g.off_x, g.off_y, param.ctypes.data,
                                     pose.ctypes.data, cov.ctypes.data, best.ctypes.data, ctypes.byref(navg))
        assert r >= 0.0, "score array does not match the window"
        return dict(response=r, pose=pose, cov=cov, best_map=best, n_avg=navg.value)

    def match(self, grid, g, pts, param, pose_world, cov=None):
        pts, param = _f64(pts), _f64(param)
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        best = np.zeros(4)
        navg = c_l(0)
        sec = c_d(0)
        r = self.L.orc_match(grid.ctypes.data, g.size_x, g.size_y, 1.0 / g.res, g.off_x, g.off_y, len(pts),
                             pts.ctypes.data, param.ctypes.data, pose.ctypes.data, cov.ctypes.data,
                             best.ctypes.data, ctypes.byref(navg), ctypes.byref(sec))
        return dict(response=r, pose=pose, cov=cov, best_map=best, n_avg=navg.value, seconds=sec.value)

    def match_chain(self, grid, g, pts, params, pose_world, cov=None, use_fine=True):
        pts = _f64(pts)
        params = _f64(np.concatenate(params))
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        resp = np.zeros(3)
        sec = c_d(0)
        r = self.L.orc_match_chain(grid.ctypes.data, g.size_x, g.size_y, 1.0 / g.res, g.off_x, g.off_y, len(pts),
                                   pts.ctypes.data, params.ctypes.data, int(use_fine), pose.ctypes.data,
                                   cov.ctypes.data, resp.ctypes.data, ctypes.byref(sec))
        return dict(score=r, pose=pose, cov=cov, responses=resp, seconds=sec.value)


    # ---- front-end maps: incremental UpdateMapByRange, ExtendSize copy, publishing map -------------------
    def grid_stamp(self, grid, g, pts_cells, pose_world):
        """UpdateMapByRange(scan, use_blur) on a scan-match map as it is (no reset); grid is modified in place."""
        n_pts, pts = pack_scans([pts_cells])
        pose = _f64(pose_world)
        rc = self.L.orc_grid_stamp(grid.ctypes.data, g.size_x, g.size_y, g.default_prob, g.sigma, g.res, g.occu_offset,
                                   g.off_x, g.off_y, 1, n_pts.ctypes.data, pts.ctypes.data, pose.ctypes.data, int(g.use_blur), 0)
        if rc:
            raise RuntimeError("orc_grid_stamp rc=%d" % rc)

    def grid_extend(self, grid, new_sx, new_sy, pre, fill=0.5, first=0.3):
        old_sy, old_sx = grid.shape
        out = np.empty((new_sy, new_sx), dtype=np.float32)
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        self.L.orc_grid_extend(grid.ctypes.data, old_sx, old_sy, out.ctypes.data, new_sx, new_sy, int(pre[0]), int(pre[1]),
                               float(fill), float(first))
        return out

    def pubmap_new(self, size_x, size_y, default_prob=0.5):
        shape = (size_y, size_x)
        return dict(hit=np.zeros(shape, np.float32), passc=np.zeros(shape, np.float32),
                    value=np.full(shape, np.float32(default_prob), np.float32), index=np.full(shape, -1, np.int32), cur=0)

    def pubmap_update(self, pm, g_res, off_x, off_y, pts_cells, pose_world, free_factor, occu_factor):
        pts, pose = _f64(pts_cells), _f64(pose_world)
        cur = c_i(pm["cur"])
        sy, sx = pm["value"].shape
        self.L.orc_pubmap_update(pm["hit"].ctypes.data, pm["passc"].ctypes.data, pm["value"].ctypes.data, pm["index"].ctypes.data,
                                 sx, sy, 1.0 / g_res, off_x, off_y, len(pts), pts.ctypes.data, pose.ctypes.data,
                                 float(free_factor), float(occu_factor), ctypes.byref(cur))
        pm["cur"] = cur.value

    def pubmap_extend(self, pm, new_sx, new_sy, pre, default_prob=0.5):
        out = self.pubmap_new(new_sx, new_sy, default_prob)
        sy, sx = pm["value"].shape
        for k in ("hit", "passc", "value", "index"):
            out[k][pre[1]:pre[1] + sy, pre[0]:pre[0] + sx] = pm[k]
        out["cur"] = pm["cur"] + 3            # the update that triggered the extension still advances the index (:296-299)
        return out

    # ---- BasedOptimizeScanMatch (optimize_scan_matcher.h) ------------------------------------------
    # op = (iterate_max_times, cost_decrease_threshold, cost_min_threshold, max_update_distance, max_update_angle)
    def optimize(self, grid, g, pts, op, pose_world):
        pts, op = _f64(pts), _f64(op)
        pose = _f64(pose_world).copy()
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        iters, oob = c_i(0), c_l(0)
        cost = self.L.orc_optimize(grid.ctypes.data, g.size_x, g.size_y, 1.0 / g.res, g.off_x, g.off_y, len(pts),
                                   pts.ctypes.data, op.ctypes.data, pose.ctypes.data, ctypes.byref(iters), ctypes.byref(oob))
        return dict(cost=cost, pose=pose, iterations=iters.value, oob_reads=oob.value)

    def optimize_cost(self, grid, g, pts, pose_map):
        pts, pm = _f64(pts), _f64(pose_map)
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        H, b, cost = np.zeros((3, 3)), np.zeros(3), c_d(0)
        self.L.orc_optimize_cost(grid.ctypes.data, g.size_x, g.size_y, len(pts), pts.ctypes.data, pm.ctypes.data,
                                 H.ctypes.data, b.ctypes.data, ctypes.byref(cost))
        return H, b, cost.value

    def ldlt3_solve(self, H, b):
        Hc = _f64(np.asarray(H).T)      # column-major
        b = _f64(b)
        x = np.zeros(3)
        self.L.orc_ldlt3_solve(Hc.ctypes.data, b.ctypes.data, x.ctypes.data)
        return x

    def match_chain_opt(self, grid_c, g_c, pts_c, grid_f, g_f, pts_f, params, op, failed_cost, pose_world, cov=None,
                        use_fine=True):
        pts_c, pts_f, op = _f64(pts_c), _f64(pts_f), _f64(op)
        params = _f64(np.concatenate(params))
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        grid_c = np.ascontiguousarray(grid_c, dtype=np.float32)
        grid_f = np.ascontiguousarray(grid_f, dtype=np.float32)
        resp = np.zeros(4)
        r = self.L.orc_match_chain_opt(grid_c.ctypes.data, g_c.size_x, g_c.size_y, 1.0 / g_c.res, g_c.off_x, g_c.off_y,
                                       len(pts_c), pts_c.ctypes.data, grid_f.ctypes.data, g_f.size_x, g_f.size_y,
                                       1.0 / g_f.res, g_f.off_x, g_f.off_y, len(pts_f), pts_f.ctypes.data,
                                       params.ctypes.data, op.ctypes.data, float(failed_cost), int(use_fine),
                                       pose.ctypes.data, cov.ctypes.data, resp.ctypes.data)
        return dict(score=r, pose=pose, cov=cov, optimize_cost=resp[0], responses=resp[1:].copy())


class Ref:
    """The reference's own code.  A 'map' here is an opaque handle to a live ScanMatchMap."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libref.so missing (reference not built here)")
        L = ctypes.CDLL(path)
        L.ref_map_create.restype = c_p
        L.ref_map_create.argtypes = [c_d, c_i, c_i, c_d, c_d, c_d, ctypes.c_float, c_d]
        L.ref_map_destroy.argtypes = [c_p]
        L.ref_map_set_offset.argtypes = [c_p, c_d, c_d]
        L.ref_map_build.restype = c_i
        L.ref_map_build.argtypes = [c_p, c_i, c_p, c_p, c_p, c_i]
        L.ref_map_size.argtypes = [c_p, c_p, c_p]
        L.ref_map_cell_length.restype = c_d
        L.ref_map_cell_length.argtypes = [c_p]
        L.ref_map_read.argtypes = [c_p, c_p]
        L.ref_map_write.argtypes = [c_p, c_p]
        L.ref_world_to_map.argtypes = [c_p, c_p, c_p]
        L.ref_map_to_world.argtypes = [c_p, c_p, c_p]
        L.ref_candidate_count.restype = c_l
        L.ref_candidate_count.argtypes = [c_p]
        L.ref_match_candidates.restype = c_d
        L.ref_match_candidates.argtypes = [c_p, c_i, c_p, c_p, c_p, c_l, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
        L.ref_match.restype = c_d
        L.ref_match.argtypes = [c_p, c_i, c_p, c_p, c_p, c_p, c_p]
        L.ref_match_chain.restype = c_d
        L.ref_match_chain.argtypes = [c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p]
        L.ref_blur_kernel.restype = c_i
        L.ref_blur_kernel.argtypes = [c_d, c_d, c_p, c_i]
        L.ref_pubmap_create_frontend.restype = c_p
        L.ref_pubmap_create_frontend.argtypes = [c_d, c_i, c_i, c_d, c_d, c_d]
        L.ref_pubmap_set_factors.argtypes = [c_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
        L.ref_pubmap_update_geom.restype = c_i
        L.ref_pubmap_update_geom.argtypes = [c_p, c_i, c_p, c_p, c_p]
        L.ref_pubmap_read_all.argtypes = [c_p, c_p, c_p, c_p, c_p]
        L.ref_frontend_map_create.restype = c_p
        L.ref_frontend_map_create.argtypes = [c_d, c_i, c_i, c_d, c_d, c_d, ctypes.c_float, c_d, c_d]
        L.ref_frontend_map_update.restype = c_i
        L.ref_frontend_map_update.argtypes = [c_p, c_i, c_p, c_p, c_i, c_p]
        L.ref_frontend_map_size_check.restype = c_i
        L.ref_frontend_map_size_check.argtypes = [c_p, c_p, c_d, c_d, c_p]
        L.ref_optimize.restype = c_d
        L.ref_optimize.argtypes = [c_p, c_i, c_p, c_p, c_p]
        L.ref_match_chain_opt.restype = c_d
        L.ref_match_chain_opt.argtypes = [c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_d, c_i, c_p, c_p, c_p]
        L.ref_pubmap_create.restype = c_p
        L.ref_pubmap_create.argtypes = [c_d, c_i, c_i, c_d, c_d, ctypes.c_float]
        L.ref_pubmap_destroy.argtypes = [c_p]
        L.ref_pubmap_update.restype = c_i
        L.ref_pubmap_update.argtypes = [c_p, c_i, c_p, c_p]
        L.ref_pubmap_read.argtypes = [c_p, c_p, c_p, c_p]
        L.ref_pubmap_penalty.restype = c_d
        L.ref_pubmap_penalty.argtypes = [c_p, c_i, c_p, c_p, c_p, c_i, c_d, c_d, c_i, c_i]
        self.L = L

    # ---- publishing map (CountCell) and MapFeedbackResponsePenalty -------------------------------
    def pubmap_create(self, g, default_prob=0.5):
        return self.L.ref_pubmap_create(g.res, g.size_x, g.size_y, g.off_x, g.off_y, default_prob)

    def pubmap_destroy(self, m):
        self.L.ref_pubmap_destroy(m)

    def pubmap_update(self, m, pts_cells, pose_world):
        pts = np.ascontiguousarray(pts_cells, dtype=np.float64)
        pose = np.ascontiguousarray(pose_world, dtype=np.float64)
        return self.L.ref_pubmap_update(m, len(pts), pts.ctypes.data, pose.ctypes.data)

    def pubmap_read(self, m, g):
        n = g.size_x * g.size_y
        val = np.zeros(n, dtype=np.float32); cnt = np.zeros(n, dtype=np.float32); occ = np.zeros(n, dtype=np.uint8)
        self.L.ref_pubmap_read(m, val.ctypes.data, cnt.ctypes.data, occ.ctypes.data)
        shape = (g.size_y, g.size_x)
        return val.reshape(shape), cnt.reshape(shape), occ.reshape(shape)

    def pubmap_penalty(self, m, pts_cells, pose_world, check_point_num, bound_tolerance, penalty_gain, use_blur=False,
                       use_logistic=False, origin=None):
        pts = np.ascontiguousarray(pts_cells, dtype=np.float64)
        pose = np.ascontiguousarray(pose_world, dtype=np.float64)
        org = np.ascontiguousarray(origin, dtype=np.float64) if origin is not None else None
        return self.L.ref_pubmap_penalty(m, len(pts), pts.ctypes.data, pose.ctypes.data,
                                         org.ctypes.data if org is not None else None, int(check_point_num),
                                         float(bound_tolerance), float(penalty_gain), 1 if use_blur else 0,
                                         1 if use_logistic else 0)

    # ---- publishing map as the front end keeps it -----------------------------------------------------
    def pubmap_create_frontend(self, g, extend_factor=0.2):
        return self.L.ref_pubmap_create_frontend(g.res, g.size_x, g.size_y, g.off_x, g.off_y, float(extend_factor))

    def pubmap_set_factors(self, m, free_factor, occu_factor, occu_threshold, min_pass_through):
        self.L.ref_pubmap_set_factors(m, free_factor, occu_factor, occu_threshold, min_pass_through)

    def pubmap_update_geom(self, m, pts_cells, pose_world):
        pts, pose = _f64(pts_cells), _f64(pose_world)
        geom = np.zeros(4)
        ok = self.L.ref_pubmap_update_geom(m, len(pts), pts.ctypes.data, pose.ctypes.data, geom.ctypes.data)
        return bool(ok), (int(geom[0]), int(geom[1]), float(geom[2]), float(geom[3]))

    def pubmap_read_all(self, m, size_x, size_y):
        n = size_x * size_y
        val, cnt, hit = (np.zeros(n, dtype=np.float32) for _ in range(3))
        occ = np.zeros(n, dtype=np.uint8)
        self.L.ref_pubmap_read_all(m, val.ctypes.data, cnt.ctypes.data, hit.ctypes.data, occ.ctypes.data)
        shape = (size_y, size_x)
        return val.reshape(shape), cnt.reshape(shape), hit.reshape(shape), occ.reshape(shape)

    # ---- front-end scan-match map (auto-resize, incremental UpdateMapByRange) ---------------------
    def frontend_map_create(self, g, extend_factor=0.2):
        return self.L.ref_frontend_map_create(g.res, g.size_x, g.size_y, g.off_x, g.off_y, g.sigma, g.default_prob,
                                              g.occu_offset, float(extend_factor))

    def frontend_map_update(self, m, pts_cells, pose_world, use_blur=True):
        """-> (stamped?, (size_x, size_y, off_x, off_y) afterwards)"""
        pts, pose = _f64(pts_cells), _f64(pose_world)
        geom = np.zeros(4)
        ok = self.L.ref_frontend_map_update(m, len(pts), pts.ctypes.data, pose.ctypes.data, int(use_blur), geom.ctypes.data)
        return bool(ok), (int(geom[0]), int(geom[1]), float(geom[2]), float(geom[3]))

    def map_size(self, m):
        sx, sy = c_i(0), c_i(0)
        self.L.ref_map_size(m, ctypes.byref(sx), ctypes.byref(sy))
        return sx.value, sy.value

    def frontend_map_size_check(self, m, pose_world, range_max, offset):
        pose = _f64(pose_world)
        geom = np.zeros(4)
        ok = self.L.ref_frontend_map_size_check(m, pose.ctypes.data, float(range_max), float(offset), geom.ctypes.data)
        return bool(ok), (int(geom[0]), int(geom[1]), float(geom[2]), float(geom[3]))

    def read_map_sized(self, m, size_x, size_y):
        out = np.empty((size_y, size_x), dtype=np.float32)
        self.L.ref_map_read(m, out.ctypes.data)
        return out

    def optimize(self, m, pts, op, pose_world):
        """The reference's BasedOptimizeScanMatch::ScanMatch (LDLT solve = the stand-in's restatement of Eigen's)."""
        pts, op = _f64(pts), _f64(op)
        pose = _f64(pose_world).copy()
        cost = self.L.ref_optimize(m, len(pts), pts.ctypes.data, op.ctypes.data, pose.ctypes.data)
        return dict(cost=cost, pose=pose)

    def match_chain_opt(self, m_c, pts_c, m_f, pts_f, params, op, failed_cost, pose_world, cov=None, use_fine=True):
        pts_c, pts_f, op = _f64(pts_c), _f64(pts_f), _f64(op)
        params = _f64(np.concatenate(params))
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        resp = np.zeros(4)
        r = self.L.ref_match_chain_opt(m_c, len(pts_c), pts_c.ctypes.data, m_f, len(pts_f), pts_f.ctypes.data,
                                       params.ctypes.data, op.ctypes.data, float(failed_cost), int(use_fine),
                                       pose.ctypes.data, cov.ctypes.data, resp.ctypes.data)
        return dict(score=r, pose=pose, cov=cov, optimize_cost=resp[0], responses=resp[1:].copy())

    def blur_kernel(self, sigma, res):
        k = np.zeros(21 * 21, dtype=np.float64)
        h = self.L.ref_blur_kernel(sigma, res, k.ctypes.data, k.size)
        if h < 0:
            return -1, None
        n = 2 * h + 1
        return h, k[: n * n].reshape(n, n).copy()

    def create_map(self, g):
        return self.L.ref_map_create(g.res, g.size_x, g.size_y, g.off_x, g.off_y, g.sigma, g.default_prob, g.occu_offset)

    def destroy_map(self, m):
        self.L.ref_map_destroy(m)

    def build_map(self, m, g, base_pts, base_poses):
        n_pts, pts = pack_scans(base_pts)
        poses = _f64(base_poses)
        rc = self.L.ref_map_build(m, len(base_pts), n_pts.ctypes.data, pts.ctypes.data, poses.ctypes.data, int(g.use_blur))
        if rc:
            raise RuntimeError("ref_map_build rc=%d" % rc)

    def read_map(self, m, g):
        out = np.empty((g.size_y, g.size_x), dtype=np.float32)
        self.L.ref_map_read(m, out.ctypes.data)
        return out

    def write_map(self, m, grid):
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        self.L.ref_map_write(m, grid.ctypes.data)

    def world_to_map(self, m, w):
        out = np.zeros(3)
        w = _f64(w)
        self.L.ref_world_to_map(m, w.ctypes.data, out.ctypes.data)
        return out

    def map_to_world(self, m, p):
        out = np.zeros(3)
        p = _f64(p)
        self.L.ref_map_to_world(m, p.ctypes.data, out.ctypes.data)
        return out

    def candidates(self, m, pts, param, center_map):
        """Sorted candidate list exactly as the reference leaves it after a pass."""
        pts, param, center_map = _f64(pts), _f64(param), _f64(center_map)
        n = self.L.ref_candidate_count(param.ctypes.data)
        cx, cy, ca, cs = (np.empty(n) for _ in range(4))
        ci = np.empty(n, dtype=np.int32)
        nout = c_l(0)
        best = np.zeros(4)
        self.L.ref_match_candidates(m, len(pts), pts.ctypes.data, param.ctypes.data, center_map.ctypes.data, n,
                                    cx.ctypes.data, cy.ctypes.data, ca.ctypes.data, ci.ctypes.data, cs.ctypes.data,
                                    ctypes.byref(nout), best.ctypes.data)
        assert nout.value == n, (nout.value, n)
        return dict(x=cx, y=cy, angle=ca, angle_index=ci, score=cs, best=best)

    def match(self, m, pts, param, pose_world, cov=None):
        pts, param = _f64(pts), _f64(param)
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        sec = c_d(0)
        r = self.L.ref_match(m, len(pts), pts.ctypes.data, param.ctypes.data, pose.ctypes.data, cov.ctypes.data, ctypes.byref(sec))
        return dict(response=r, pose=pose, cov=cov, seconds=sec.value)

    def match_chain(self, m, pts, params, pose_world, cov=None, use_fine=True):
        pts = _f64(pts)
        params = _f64(np.concatenate(params))
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        resp = np.zeros(3)
        sec = c_d(0)
        r = self.L.ref_match_chain(m, len(pts), pts.ctypes.data, params.ctypes.data, int(use_fine), pose.ctypes.data,
                                   cov.ctypes.data, resp.ctypes.data, ctypes.byref(sec))
        return dict(score=r, pose=pose, cov=cov, responses=resp, seconds=sec.value)


def dropin_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdropin.so"))


class DropIn:
    """The product's C++ adapter (scan_matcher_adapter.hpp) running on live reference objects.

    Built only where the reference was available (oracle/_ref/libdropin.so); needs a GPU."""

    def __init__(self, device=0):
        L = ctypes.CDLL(os.path.join(_HERE, "_ref", "libdropin.so"))
        L.dropin_create.restype = c_p
        L.dropin_create.argtypes = [c_i]
        L.dropin_destroy.argtypes = [c_p]
        L.dropin_match_chain.restype = c_d
        L.dropin_match_chain.argtypes = [c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p, c_p]
        L.dropin_update_map.restype = c_i
        L.dropin_update_map.argtypes = [c_p, c_i, c_p, c_p, c_i, c_d, c_d]
        L.dropin_mirror_equals_host.restype = c_i
        L.dropin_mirror_equals_host.argtypes = [c_p]
        L.dropin_sync_counts.argtypes = [ctypes.POINTER(c_l), ctypes.POINTER(c_l)]
        L.dropin_opt_create.restype = c_p
        L.dropin_opt_create.argtypes = [c_i]
        L.dropin_opt_destroy.argtypes = [c_p]
        L.dropin_optimize.restype = c_d
        L.dropin_optimize.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p]
        self.L = L
        self.h = L.dropin_create(device)
        self.opt = L.dropin_opt_create(device)
        if not self.h or not self.opt:
            raise RuntimeError("adapter could not create a CUDA context")

    def close(self):
        if self.h:
            self.L.dropin_destroy(self.h)
            self.h = None
        if self.opt:
            self.L.dropin_opt_destroy(self.opt)
            self.opt = None

    def optimize(self, ref_map, pts, op, pose_world):
        """rsm_adapter::BasedOptimizeScanMatch::ScanMatch on a live reference map."""
        pts, op = _f64(pts), _f64(op)
        pose = _f64(pose_world).copy()
        cost = self.L.dropin_optimize(self.opt, ref_map, len(pts), pts.ctypes.data, op.ctypes.data, pose.ctypes.data)
        return dict(cost=cost, pose=pose)

    def match_chain(self, ref_map, pts, params, pose_world, cov=None):
        """params: list of 1 or 3 pass-parameter blocks; ref_map: handle from Ref.create_map."""
        pts = _f64(pts)
        n_pass = len(params)
        params = _f64(np.concatenate(params))
        pose = _f64(pose_world).copy()
        cov = np.eye(3) if cov is None else _f64(cov).copy()
        resp = np.zeros(3)
        used = c_i(0)
        s = self.L.dropin_match_chain(self.h, ref_map, len(pts), pts.ctypes.data, params.ctypes.data, n_pass,
                                      pose.ctypes.data, cov.ctypes.data, resp.ctypes.data, ctypes.byref(used))
        return dict(score=s, pose=pose, cov=cov, responses=resp, exact_used=used.value)

    def update_map(self, ref_map, pts_cells, pose_world, use_blur, deviation, occu_offset):
        """rsm_adapter::UpdateMapByRange: the reference's own update on the host map + the stamp on the device mirror."""
        pts, pose = _f64(pts_cells), _f64(pose_world)
        return bool(self.L.dropin_update_map(ref_map, len(pts), pts.ctypes.data, pose.ctypes.data, int(use_blur), float(deviation),
                                             float(occu_offset)))

    def mirror_equals_host(self, ref_map):
        return self.L.dropin_mirror_equals_host(ref_map)

    def sync_counts(self):
        a, b = c_l(0), c_l(0)
        self.L.dropin_sync_counts(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


def batcher_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libbatcher.so"))


class Batcher:
    """The product's back-end batcher (csrc/backend_batcher.hpp) fed from a LIVE reference SensorDataManager
    (oracle/_ref/libbatcher.so); needs a GPU."""

    def __init__(self, res, dev, sizes, blur_offset, default_prob, passes, opt):
        L = ctypes.CDLL(os.path.join(_HERE, "_ref", "libbatcher.so"))
        L.batcher_create.restype = c_p
        L.batcher_create.argtypes = [c_p, c_p, c_p, c_d, ctypes.c_float, c_p, c_p]
        L.batcher_destroy.argtypes = [c_p]
        L.batcher_add_scan.restype = c_i
        L.batcher_add_scan.argtypes = [c_p, c_i, c_p, c_p]
        L.batcher_get_fine_scan.restype = c_i
        L.batcher_get_fine_scan.argtypes = [c_p, c_i, c_p, c_i]
        L.batcher_set_pose.argtypes = [c_p, c_i, c_p]
        L.batcher_try_close_loop.restype = c_i
        L.batcher_try_close_loop.argtypes = [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
        self.L = L
        res, dev, passes, opt = _f64(res), _f64(dev), _f64(np.concatenate(passes)), _f64(opt)
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        self.h = L.batcher_create(res.ctypes.data, dev.ctypes.data, sizes.ctypes.data, float(blur_offset), float(default_prob),
                                  passes.ctypes.data, opt.ctypes.data)
        if not self.h:
            raise RuntimeError("batcher could not create a CUDA context")

    def close(self):
        if self.h:
            self.L.batcher_destroy(self.h)
            self.h = None

    def add_scan(self, pts_metres, pose_world):
        pts, pose = _f64(pts_metres), _f64(pose_world)
        return self.L.batcher_add_scan(self.h, len(pts), pts.ctypes.data, pose.ctypes.data)

    def fine_scan(self, scan_id, cap=4096):
        out = np.zeros((cap, 2))
        n = self.L.batcher_get_fine_scan(self.h, int(scan_id), out.ctypes.data, cap)
        return out[:n].copy()

    def set_pose(self, scan_id, pose_world):
        pose = _f64(pose_world)
        self.L.batcher_set_pose(self.h, int(scan_id), pose.ctypes.data)

    def try_close_loop(self, range_id, chains, scan_pose, centre, thresholds):
        off = np.zeros(len(chains) + 1, dtype=np.int32)
        for i, ch in enumerate(chains):
            off[i + 1] = off[i] + len(ch)
        ids = np.ascontiguousarray(np.concatenate([np.asarray(ch, dtype=np.int32) for ch in chains]), dtype=np.int32)
        pose, cen, th = _f64(scan_pose), _f64(centre), _f64(thresholds)
        best, cov = np.zeros(3), np.zeros(9)
        s1, s2 = np.zeros(len(chains)), np.zeros(len(chains))
        hit = self.L.batcher_try_close_loop(self.h, int(range_id), len(chains), off.ctypes.data, ids.ctypes.data, pose.ctypes.data,
                                            cen.ctypes.data, th.ctypes.data, best.ctypes.data, cov.ctypes.data, s1.ctypes.data,
                                            s2.ctypes.data)
        return hit, best, cov.reshape(3, 3), s1, s2
