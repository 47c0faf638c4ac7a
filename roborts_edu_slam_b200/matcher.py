"""Host-side mirror of the reference's matcher interface over the C ABI (include/rsm.h).

The reference is C++ (its drop-in adapter is roborts_edu_slam_b200/csrc/scan_matcher_adapter.hpp);
this module is the same surface for Python callers -- tests, bench.py and torch.distributed
launchers -- with the reference's names and argument meaning:

  CorrelationScanMatchParam      scan_match/correlate_scan_matcher.h:41-86
  ScanMatchMap                   map/slam_map.h:32-34 (device-resident lookup grid)
  BasedCorrelationScanMatch      .ScanMatch(map, scan, param, pose, cov) -> response   (:784-875)
  ScanMatchers                   .ScanMatch(scan, map, pose, cov, use_fine) -> score   (scan_matchers.h:179-289)

Every compute call goes to librsm.so (CUDA, sm_100a).  There is no CPU path: importing works
without a GPU (so the ABI can be inspected), creating a Context does not.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSM_LIB_PATH") or os.path.join(_HERE, "librsm.so")   # override: debug builds only

COARSE_CORRELATION_SCAN_MATCH = 0
FINE_CORRELATION_SCAN_MATCH = 1
SUPER_CORRELATION_SCAN_MATCH = 2
FAST_CORRELATION_SCAN_MATCH = 3

RSM_OK = 0
RSM_NEED_EXACT = 7
RSM_OPT_STRICT_TIES = 1
RSM_OPT_LANES = 2
STATUS_NAMES = {0: "RSM_OK", 1: "RSM_ERR_NO_DEVICE", 2: "RSM_ERR_CUDA", 3: "RSM_ERR_INVALID",
                4: "RSM_ERR_WINDOW", 5: "RSM_ERR_UNSUPPORTED", 6: "RSM_ERR_NOT_INIT", 7: "RSM_NEED_EXACT"}

c_d, c_i, c_p, c_i64 = ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64


class RsmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


class PassParamStruct(ctypes.Structure):
    _fields_ = [("search_space_size", c_d), ("search_space_resolution", c_d),
                ("search_angle_offset", c_d), ("search_angle_resolution", c_d),
                ("response_threshold", c_d), ("use_point_size", ctypes.c_int32),
                ("use_center_penalty", ctypes.c_int32), ("type", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class PassDetail(ctypes.Structure):
    _fields_ = [("best_score", c_d), ("best_pose_map", c_d * 3), ("n_candidates", c_i64),
                ("n_avg", ctypes.c_int32), ("exact_sort_used", ctypes.c_int32),
                ("pose_updated", ctypes.c_int32), ("n_ang", ctypes.c_int32), ("n_xy", ctypes.c_int32),
                ("visited", ctypes.c_int32), ("divisor", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", c_i64), ("score_launches", c_i64), ("evals", c_i64), ("passes", c_i64),
                ("exact_sort_passes", c_i64), ("h2d_bytes", c_i64), ("d2h_bytes", c_i64),
                ("score_kernel_ms", c_d), ("raster_kernel_ms", c_d), ("select_kernel_ms", c_d),
                ("phase_ms", c_d * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_ms"] = [float(v) for v in self.phase_ms]
        return d


class MapGeometryStruct(ctypes.Structure):
    """rsm_map_geometry"""
    _fields_ = [("size_x", ctypes.c_int32), ("size_y", ctypes.c_int32), ("pre_grid_offset_x", ctypes.c_int32),
                ("pre_grid_offset_y", ctypes.c_int32), ("offset_x", c_d), ("offset_y", c_d)]


class MapCheckParamStruct(ctypes.Structure):
    """rsm_map_check_param"""
    _fields_ = [("bound_tolerance", c_d), ("penalty_gain", c_d), ("check_point_num", ctypes.c_int32),
                ("use_logistic", ctypes.c_int32)]


class OptimizeParamStruct(ctypes.Structure):
    """rsm_optimize_param (OptimizeScanMatchParam, optimize_scan_matcher.h:33-58)"""
    _fields_ = [("cost_decrease_threshold", c_d), ("cost_min_threshold", c_d), ("max_update_distance", c_d),
                ("max_update_angle", c_d), ("iterate_max_times", ctypes.c_int32), ("reserved", ctypes.c_int32)]


# every symbol include/rsm.h declares: (restype, argtypes)
_PPARAM = ctypes.POINTER(PassParamStruct)
ABI = {
    "rsm_version": (ctypes.c_char_p, []),
    "rsm_create": (c_i, [c_i, ctypes.POINTER(c_p)]),
    "rsm_destroy": (None, [c_p]),
    "rsm_last_error": (ctypes.c_char_p, [c_p]),
    "rsm_set_profiling": (c_i, [c_p, c_i]),
    "rsm_set_option": (c_i, [c_p, c_i, c_i]),
    "rsm_get_stats": (c_i, [c_p, ctypes.POINTER(Stats)]),
    "rsm_reset_stats": (c_i, [c_p]),
    "rsm_synchronize": (c_i, [c_p]),
    "rsm_timer_start": (c_i, [c_p]),
    "rsm_timer_stop": (c_i, [c_p, ctypes.POINTER(c_d)]),
    "rsm_flush_l2": (c_i, [c_p]),
    "rsm_grid_create": (c_i, [c_p, c_i, c_i, c_d, c_d, c_d, ctypes.POINTER(c_p)]),
    "rsm_grid_create_from_scale": (c_i, [c_p, c_i, c_i, c_d, c_d, c_d, ctypes.POINTER(c_p)]),
    "rsm_grid_destroy": (None, [c_p, c_p]),
    "rsm_grid_set_offset": (c_i, [c_p, c_p, c_d, c_d]),
    "rsm_grid_upload_f32": (c_i, [c_p, c_p, c_p]),
    "rsm_grid_rasterize": (c_i, [c_p, c_p, ctypes.c_float, c_d, c_d, c_i, c_i, c_p, c_p, c_p]),
    "rsm_grid_download_f32": (c_i, [c_p, c_p, c_p]),
    "rsm_pubmap_create": (c_i, [c_p, c_i, c_i, c_d, c_d, c_d, ctypes.c_float, ctypes.POINTER(c_p)]),
    "rsm_pubmap_destroy": (None, [c_p, c_p]),
    "rsm_pubmap_update_by_range": (c_i, [c_p, c_p, c_p, c_i, c_p, ctypes.c_float, ctypes.c_float]),
    "rsm_pubmap_extend": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_d]),
    "rsm_pubmap_refresh_occupancy": (c_i, [c_p, c_p, ctypes.c_float, ctypes.c_float]),
    "rsm_pubmap_check_grid": (c_p, [c_p]),
    "rsm_pubmap_download": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p]),
    "rsm_grid_rebuild": (c_i, [c_p, c_p, c_p, c_i, c_p, ctypes.c_float, c_d, c_d, c_i]),
    "rsm_pubmap_rebuild": (c_i, [c_p, c_p, c_p, c_i, c_p, ctypes.c_float, ctypes.c_float]),
    "rsm_blur_half_size": (c_i, [c_d, c_d]),
    "rsm_map_bounds_create": (c_i, [c_i, c_i, c_d, c_d, c_d, c_d, ctypes.POINTER(c_p)]),
    "rsm_map_bounds_destroy": (None, [c_p]),
    "rsm_map_bounds_update_scan": (c_i, [c_p, c_p, c_i, c_p, c_i, c_i, ctypes.POINTER(c_i), ctypes.POINTER(MapGeometryStruct)]),
    "rsm_map_bounds_size_check": (c_i, [c_p, c_p, c_d, c_d, ctypes.POINTER(c_i), ctypes.POINTER(MapGeometryStruct)]),
    "rsm_grid_fill": (c_i, [c_p, c_p, ctypes.c_float, ctypes.c_float]),
    "rsm_grid_update_by_range": (c_i, [c_p, c_p, c_d, c_d, c_i, c_p, c_i, c_p]),
    "rsm_grid_update_by_range_map": (c_i, [c_p, c_p, c_d, c_d, c_i, c_p, c_i, c_p]),
    "rsm_grid_extend": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_d, ctypes.c_float, ctypes.c_float]),
    "rsm_grid_geometry": (c_i, [c_p, ctypes.POINTER(c_i), ctypes.POINTER(c_i), ctypes.POINTER(c_d), ctypes.POINTER(c_d)]),
    "rsm_grid_upload_occupancy": (c_i, [c_p, c_p, c_p]),
    "rsm_map_check_penalize": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_i, c_p]),
    "rsm_grid_is_fixed_point": (c_i, [c_p]),
    "rsm_world_to_map": (c_i, [c_p, c_p, c_p]),
    "rsm_map_to_world": (c_i, [c_p, c_p, c_p]),
    "rsm_match": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_p, c_p, ctypes.POINTER(c_d), ctypes.POINTER(PassDetail)]),
    "rsm_match_map": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_p, c_p, ctypes.POINTER(c_d), c_p, ctypes.POINTER(PassDetail)]),
    "rsm_scan_create": (c_i, [c_p, c_p, c_i, ctypes.POINTER(c_p)]),
    "rsm_scan_destroy": (None, [c_p, c_p]),
    "rsm_match_resident": (c_i, [c_p, c_p, c_p, _PPARAM, c_p, c_p, ctypes.POINTER(c_d), ctypes.POINTER(PassDetail)]),
    "rsm_microbench_gather": (c_i, [c_p, c_i, c_i64, c_i, ctypes.POINTER(c_d)]),
    "rsm_stream_plan": (c_i, [c_i, c_p, c_i, c_i, c_p, c_i, ctypes.POINTER(c_i64), ctypes.POINTER(c_i), ctypes.POINTER(c_i)]),
    "rsm_match_chain": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_i, c_p, c_p, ctypes.POINTER(c_d), c_p]),
    "rsm_match_batch": (c_i, [c_p, c_i, c_p, c_p, c_p, _PPARAM, c_i, c_i, c_p, c_p, c_p, c_p]),
    "rsm_loop_closure_batch": (c_i, [c_p, c_i, c_i, c_d, ctypes.c_float, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                     _PPARAM, c_i, c_p, c_p, c_p, c_p]),
    "rsm_scan_store_create": (c_i, [c_p, ctypes.POINTER(c_p)]),
    "rsm_scan_store_destroy": (None, [c_p, c_p]),
    "rsm_scan_store_add": (c_i, [c_p, c_p, c_p, c_i, c_p, ctypes.POINTER(ctypes.c_int32)]),
    "rsm_scan_store_set_poses": (c_i, [c_p, c_p, c_i, c_p, c_p]),
    "rsm_scan_store_get_pose": (c_i, [c_p, ctypes.c_int32, c_p]),
    "rsm_scan_store_size": (c_i, [c_p]),
    "rsm_scan_match_interface_batch": (c_i, [c_p, c_p, c_i, c_i, c_d, ctypes.c_float, c_d, c_d, c_p, c_p, c_p, c_p,
                                             _PPARAM, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "rsm_scan_match_interface_batch_opt": (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_d, c_i, c_d, c_d, ctypes.c_float, c_d, c_p, c_p, c_p, c_p,
                                                 _PPARAM, ctypes.POINTER(OptimizeParamStruct), c_d, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "rsm_optimize": (c_i, [c_p, c_p, c_p, c_i, ctypes.POINTER(OptimizeParamStruct), c_p, ctypes.POINTER(c_d),
                           ctypes.POINTER(ctypes.c_int32)]),
    "rsm_optimize_map": (c_i, [c_p, c_p, c_p, c_i, ctypes.POINTER(OptimizeParamStruct), c_p, ctypes.POINTER(c_d),
                               ctypes.POINTER(ctypes.c_int32)]),
    "rsm_optimize_batch": (c_i, [c_p, c_i, c_p, c_p, c_p, ctypes.POINTER(OptimizeParamStruct), c_p, c_p, c_p]),
    "rsm_match_chain_opt": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, c_i, _PPARAM, ctypes.POINTER(OptimizeParamStruct), c_d, c_i,
                                  c_p, c_p, ctypes.POINTER(c_d), c_p]),
    "rsm_pass_scores": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_p, c_i, c_i, c_p, c_i64, ctypes.POINTER(c_i64)]),
    "rsm_match_partial": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_p, c_i, c_i, c_p]),
    "rsm_match_merge": (c_i, [c_p, c_p, c_i, c_p]),
    "rsm_match_finish": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, ctypes.POINTER(c_d), ctypes.POINTER(PassDetail)]),
    "rsm_match_slice_scores": (c_i, [c_p, c_p, c_i64, ctypes.POINTER(c_i64)]),
    "rsm_match_finish_exact": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, ctypes.POINTER(c_d), ctypes.POINTER(PassDetail)]),
    "rsm_comm_unique_id": (c_i, [c_p]),
    "rsm_comm_init": (c_i, [c_p, c_i, c_i, c_p]),
    "rsm_comm_destroy": (c_i, [c_p]),
    "rsm_match_sliced": (c_i, [c_p, c_p, c_p, c_i, _PPARAM, c_p, c_p, ctypes.POINTER(c_d), ctypes.POINTER(PassDetail)]),
}

_lib = None


def load_library():
    """dlopen librsm.so and bind every declared symbol; raises if the extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librsm.so is missing (%s): build it with `python -m roborts_edu_slam_b200.build`; "
                               "there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def stream_plan(runs, variant=-1, max_ctas=148):
    """The staged scoring kernel's stream plan for jobs (beams, n_xy, angles): host arithmetic only, no device.
    Returns (shares [n, 12] int32, n_items, n_tickets, n_slots); see rsm_stream_plan in include/rsm.h."""
    lib = load_library()
    r = np.ascontiguousarray(np.asarray(runs, dtype=np.int32).reshape(-1, 3))
    out = np.zeros((max(1, int(max_ctas)), 12), dtype=np.int32)
    n_items, n_t, n_s = c_i64(0), c_i(0), c_i(0)
    n = lib.rsm_stream_plan(len(r), r.ctypes.data, int(variant), int(max_ctas), out.ctypes.data, len(out), ctypes.byref(n_items),
                            ctypes.byref(n_t), ctypes.byref(n_s))
    if n <= 0:
        raise ValueError("rsm_stream_plan: bad arguments (%d)" % n)
    return out[:n].copy(), n_items.value, n_t.value, n_s.value


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class CorrelationScanMatchParam:
    """Same fields and accessors as the reference class (correlate_scan_matcher.h:41-86)."""

    def __init__(self, search_space_size=0.0, search_space_resolution=0.0, search_angle_offset=0.0,
                 search_angle_resolution=0.0, response_threshold=0.0, use_point_size=0, max_depth=0,
                 use_center_penalty=True, correlation_scan_match_type=COARSE_CORRELATION_SCAN_MATCH):
        self._v = dict(search_space_size=search_space_size, search_space_resolution=search_space_resolution,
                       search_angle_offset=search_angle_offset, search_angle_resolution=search_angle_resolution,
                       response_threshold=response_threshold, use_point_size=use_point_size, max_depth=max_depth,
                       use_center_penalty=use_center_penalty, correlation_scan_match_type=correlation_scan_match_type)

    def __getattr__(self, name):
        v = self.__dict__.get("_v", {})
        if name.startswith("set_") and name[4:] in v:
            key = name[4:]

            def setter(value, v=v, key=key, d=self.__dict__):
                v[key] = value
                d.pop("_struct", None)          # the cached rsm_pass_param is stale
            return setter
        if name in v:
            return lambda: v[name]
        raise AttributeError(name)

    @classmethod
    def from_array(cls, p):
        """p = the 8-double block used by synth.pass_param / the oracle."""
        return cls(p[0], p[1], p[2], p[3], p[4], int(p[5]), 0, bool(p[6]), int(p[7]))

    def struct(self):
        cached = self.__dict__.get("_struct")
        if cached is not None:
            return cached
        v = self._v
        self.__dict__["_struct"] = st = PassParamStruct(v["search_space_size"], v["search_space_resolution"], v["search_angle_offset"],
                               v["search_angle_resolution"], v["response_threshold"], int(v["use_point_size"]),
                               1 if v["use_center_penalty"] else 0, int(v["correlation_scan_match_type"]), 0)
        return st


def _as_param(p):
    return p if isinstance(p, CorrelationScanMatchParam) else CorrelationScanMatchParam.from_array(p)


_STRUCTS = {}


def _param_struct(p):
    """rsm_pass_param of a parameter object or of the 8-double block of synth.pass_param; the conversion is cached (a
    front end passes the same three parameter sets for every scan, and building the ctypes struct costs 4 us)."""
    if isinstance(p, CorrelationScanMatchParam):
        return p.struct()
    key = np.asarray(p, dtype=np.float64).tobytes()
    st = _STRUCTS.get(key)
    if st is None:
        if len(_STRUCTS) > 256:
            _STRUCTS.clear()
        st = _STRUCTS[key] = CorrelationScanMatchParam.from_array(p).struct()
    return st


class Context:
    """One per GPU / caller thread (rsm_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = c_p()
        rc = self.lib.rsm_create(int(device), ctypes.byref(h))
        if rc != RSM_OK:
            raise RsmError(rc, "rsm_create(device=%d) failed; this library has no CPU path" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != RSM_OK:
            raise RsmError(rc, (self.lib.rsm_last_error(self.h) or b"").decode())

    def set_profiling(self, on):
        self.check(self.lib.rsm_set_profiling(self.h, int(on)))

    def set_option(self, option, value):
        """RSM_OPT_STRICT_TIES (bit-equal covariance under every tie pattern), RSM_OPT_LANES (pipelined sub-batches)."""
        self.check(self.lib.rsm_set_option(self.h, int(option), int(value)))

    def stats(self):
        s = Stats()
        self.check(self.lib.rsm_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self.check(self.lib.rsm_reset_stats(self.h))

    def synchronize(self):
        self.check(self.lib.rsm_synchronize(self.h))

    def timer_start(self):
        self.check(self.lib.rsm_timer_start(self.h))

    def timer_stop(self):
        ms = c_d(0)
        self.check(self.lib.rsm_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def flush_l2(self):
        self.check(self.lib.rsm_flush_l2(self.h))

    def microbench_gather(self, mode, footprint_bytes, iters=4096):
        """GB/s of 4-byte gathers: mode 0/1 shared memory (row segments / random), 2/3 global."""
        g = c_d(0)
        self.check(self.lib.rsm_microbench_gather(self.h, int(mode), int(footprint_bytes), int(iters), ctypes.byref(g)))
        return g.value


class RangeDataContainer2d:
    """A scan resident on the device (rsm_scan): points (P,2) in cells, sensor frame."""

    def __init__(self, ctx, pts):
        self.ctx = ctx
        pts = _f64(pts).reshape(-1, 2)
        self.n = len(pts)
        h = c_p()
        ctx.check(ctx.lib.rsm_scan_create(ctx.h, pts.ctypes.data, self.n, ctypes.byref(h)))
        self.h = h

    def GetSize(self):
        return self.n

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rsm_scan_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ScanMatchMap:
    """Device-resident lookup grid with the reference map's geometry (OccuGridMap<ProbabilityCell>)."""

    def __init__(self, ctx, resolution, size_x, size_y, offset_x, offset_y):
        self.ctx = ctx
        self.resolution, self.size_x, self.size_y = float(resolution), int(size_x), int(size_y)
        h = c_p()
        ctx.check(ctx.lib.rsm_grid_create(ctx.h, self.size_x, self.size_y, self.resolution,
                                          float(offset_x), float(offset_y), ctypes.byref(h)))
        self.h = h

    @classmethod
    def from_spec(cls, ctx, g):
        return cls(ctx, g.res, g.size_x, g.size_y, g.off_x, g.off_y)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rsm_grid_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map_offset(self, off_x, off_y):
        self.ctx.check(self.ctx.lib.rsm_grid_set_offset(self.ctx.h, self.h, float(off_x), float(off_y)))

    def upload(self, prob):
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        assert prob.shape == (self.size_y, self.size_x), prob.shape
        self.ctx.check(self.ctx.lib.rsm_grid_upload_f32(self.ctx.h, self.h, prob.ctypes.data))

    def InitMapWithRangeVec(self, base_pts, base_poses, default_prob=0.3, sigma=0.15, occu_offset=0.88, use_blur=True):
        """Device-side construction from base scans (points in cells, sensor frame; poses world)."""
        n_pts = np.array([len(p) for p in base_pts], dtype=np.int32)
        pts = _f64(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in base_pts], axis=0)) \
            if len(base_pts) else np.zeros((0, 2))
        poses = _f64(base_poses)
        self.ctx.check(self.ctx.lib.rsm_grid_rasterize(self.ctx.h, self.h, float(default_prob), float(sigma),
                                                       float(occu_offset), int(use_blur), len(base_pts),
                                                       n_pts.ctypes.data, pts.ctypes.data, poses.ctypes.data))

    def InitMapWithStore(self, store, ids, default_prob=0.3, sigma=0.15, occu_offset=0.88, use_blur=True):
        """InitMapWithRangeVec over scans of a device scan store (CorrectPoseAndMap's rebuild of a scan-match map)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.ctx.check(self.ctx.lib.rsm_grid_rebuild(self.ctx.h, self.h, store.h, len(ids), ids.ctypes.data, float(default_prob),
                                                     float(sigma), float(occu_offset), int(use_blur)))

    # ---- front-end maintenance: the map stays on the device between scans ----------------------------
    def fill(self, fill_prob=0.5, first_cell_prob=0.3):
        """A freshly constructed map: cell 0 = default_prob, the others kDefaultCellProb (grid_map_base.h:150-163)."""
        self.ctx.check(self.ctx.lib.rsm_grid_fill(self.ctx.h, self.h, float(fill_prob), float(first_cell_prob)))

    def UpdateMapByRange(self, pts_cells, sensor_pose, sigma=0.15, occu_offset=0.88, use_blur=True):
        """Stamp one scan (the reference's UpdateMapByRange returned true for it); no reset."""
        pts = _f64(np.asarray(pts_cells).reshape(-1, 2))
        pose = _f64(sensor_pose)
        self.ctx.check(self.ctx.lib.rsm_grid_update_by_range(self.ctx.h, self.h, float(sigma), float(occu_offset), int(use_blur),
                                                             pts.ctypes.data, len(pts), pose.ctypes.data))

    def ExtendSize(self, new_size_x, new_size_y, pre_grid_offset, new_map_offset, fill_prob=0.5, first_cell_prob=0.3):
        """Mirror of GridMapBase::ExtendSize (grid_map_base.h:188-254) once the caller's policy decided it."""
        self.ctx.check(self.ctx.lib.rsm_grid_extend(self.ctx.h, self.h, int(new_size_x), int(new_size_y), int(pre_grid_offset[0]),
                                                    int(pre_grid_offset[1]), float(new_map_offset[0]), float(new_map_offset[1]),
                                                    float(fill_prob), float(first_cell_prob)))
        self.size_x, self.size_y = int(new_size_x), int(new_size_y)

    def geometry(self):
        sx, sy, ox, oy = c_i(0), c_i(0), c_d(0), c_d(0)
        self.ctx.check(self.ctx.lib.rsm_grid_geometry(self.h, ctypes.byref(sx), ctypes.byref(sy), ctypes.byref(ox), ctypes.byref(oy)))
        return sx.value, sy.value, ox.value, oy.value

    def upload_occupancy(self, occupied):
        """Occupancy mask of a publishing map (PubMap) for MapCheckPenalize: nonzero where the
        reference's CheckOccuLineVisitorCallback counts the cell (occu_grid_map.h:447-471)."""
        occ = np.ascontiguousarray(occupied, dtype=np.uint8)
        assert occ.shape == (self.size_y, self.size_x), occ.shape
        self.ctx.check(self.ctx.lib.rsm_grid_upload_occupancy(self.ctx.h, self.h, occ.ctypes.data))

    def MapCheckPenalize(self, scans, poses_world, check_point_num=100, bound_tolerance=2.5, penalty_gain=0.015,
                         use_logistic=False, sensor_origin=None):
        """SlamProcessor::MapCheckPenalize (slam_processor.cpp:573-595) for many candidate poses at
        once.  scans: one array of points (cells of this map, sensor frame) shared by every pose, or
        a list with one array per pose.  Returns the response coefficients."""
        poses = _f64(np.asarray(poses_world, dtype=np.float64).reshape(-1, 3))
        n = len(poses)
        if isinstance(scans, (list, tuple)):
            assert len(scans) == n
            arrs = [np.asarray(a, dtype=np.float64).reshape(-1, 2) for a in scans]
            count = np.array([len(a) for a in arrs], dtype=np.int32)
            begin = np.concatenate([[0], np.cumsum(count)[:-1]]).astype(np.int64) if n else np.zeros(0, dtype=np.int64)
            pts = _f64(np.concatenate(arrs, axis=0)) if n else np.zeros((0, 2))
        else:
            pts = _f64(np.asarray(scans, dtype=np.float64).reshape(-1, 2))
            count = np.full(n, len(pts), dtype=np.int32)
            begin = np.zeros(n, dtype=np.int64)
        org = _f64(sensor_origin) if sensor_origin is not None else None
        out = np.zeros(n, dtype=np.float64)
        self.ctx.check(self.ctx.lib.rsm_map_check_penalize(
            self.ctx.h, self.h, n, poses.ctypes.data, pts.ctypes.data if len(pts) else None, begin.ctypes.data,
            count.ctypes.data, org.ctypes.data if org is not None else None, int(check_point_num),
            float(bound_tolerance), float(penalty_gain), 1 if use_logistic else 0, out.ctypes.data))
        return out

    def download(self):
        out = np.empty((self.size_y, self.size_x), dtype=np.float32)
        self.ctx.check(self.ctx.lib.rsm_grid_download_f32(self.ctx.h, self.h, out.ctypes.data))
        return out

    def is_fixed_point(self):
        return bool(self.ctx.lib.rsm_grid_is_fixed_point(self.h))

    def GetMapCoordsPose(self, pose_world):
        w, out = _f64(pose_world), np.zeros(3)
        self.ctx.check(self.ctx.lib.rsm_world_to_map(self.h, w.ctypes.data, out.ctypes.data))
        return out

    def GetWorldCoordsPose(self, pose_map):
        m, out = _f64(pose_map), np.zeros(3)
        self.ctx.check(self.ctx.lib.rsm_map_to_world(self.h, m.ctypes.data, out.ctypes.data))
        return out


class MapBounds:
    """GridMapBase's resize policy (UpdateBound / ExtendSize), host only: no GPU needed."""

    def __init__(self, size_x, size_y, resolution, offset_x, offset_y, extend_factor=1.0):
        self.lib = load_library()
        h = c_p()
        rc = self.lib.rsm_map_bounds_create(int(size_x), int(size_y), 1.0 / float(resolution), float(offset_x), float(offset_y),
                                            float(extend_factor), ctypes.byref(h))
        if rc != RSM_OK:
            raise RsmError(rc, "rsm_map_bounds_create")
        self.h = h

    def _ret(self, rc, fits, g):
        if rc != RSM_OK:
            raise RsmError(rc, "rsm_map_bounds")
        return bool(fits.value), (g.size_x, g.size_y, g.offset_x, g.offset_y), (g.pre_grid_offset_x, g.pre_grid_offset_y)

    def UpdateMapByRange(self, pts_cells, sensor_pose, half_kernel=0, use_blur=False):
        """-> (stamp this scan?, (size_x, size_y, off_x, off_y), pre_grid_offset of the last extension)"""
        pts = _f64(np.asarray(pts_cells).reshape(-1, 2))
        pose = _f64(sensor_pose)
        fits, g = c_i(0), MapGeometryStruct()
        return self._ret(self.lib.rsm_map_bounds_update_scan(self.h, pts.ctypes.data, len(pts), pose.ctypes.data, int(half_kernel),
                                                             int(use_blur), ctypes.byref(fits), ctypes.byref(g)), fits, g)

    def MapSizeCheck(self, pose_world, range_max, offset):
        pose = _f64(pose_world)
        fits, g = c_i(0), MapGeometryStruct()
        return self._ret(self.lib.rsm_map_bounds_size_check(self.h, pose.ctypes.data, float(range_max), float(offset),
                                                            ctypes.byref(fits), ctypes.byref(g)), fits, g)

    def close(self):
        if self.h:
            self.lib.rsm_map_bounds_destroy(self.h)
            self.h = None


class PubMap(ScanMatchMap):
    """Device-resident publishing map (OccuGridMap<CountCell>, slam_map.h:35).  Inherits the map-check call:
    self.h is the map's check grid, whose occupancy mask refresh_occupancy() computes on the device."""

    def __init__(self, ctx, resolution, size_x, size_y, offset_x, offset_y, default_prob=0.5):
        self.ctx = ctx
        self.resolution, self.size_x, self.size_y = float(resolution), int(size_x), int(size_y)
        pm = c_p()
        ctx.check(ctx.lib.rsm_pubmap_create(ctx.h, self.size_x, self.size_y, self.resolution, float(offset_x), float(offset_y),
                                            float(default_prob), ctypes.byref(pm)))
        self.pm = pm
        self.h = c_p(ctx.lib.rsm_pubmap_check_grid(pm))

    def close(self):
        if getattr(self, "pm", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rsm_pubmap_destroy(self.ctx.h, self.pm)
        self.pm = None
        self.h = None

    def UpdateMapByRange(self, pts_cells, sensor_pose, update_free_factor=0.0, update_occu_factor=0.0):
        pts = _f64(np.asarray(pts_cells).reshape(-1, 2))
        pose = _f64(sensor_pose)
        self.ctx.check(self.ctx.lib.rsm_pubmap_update_by_range(self.ctx.h, self.pm, pts.ctypes.data, len(pts), pose.ctypes.data,
                                                               float(update_free_factor), float(update_occu_factor)))

    def ExtendSize(self, new_size_x, new_size_y, pre_grid_offset, new_map_offset):
        self.ctx.check(self.ctx.lib.rsm_pubmap_extend(self.ctx.h, self.pm, int(new_size_x), int(new_size_y), int(pre_grid_offset[0]),
                                                      int(pre_grid_offset[1]), float(new_map_offset[0]), float(new_map_offset[1])))
        self.size_x, self.size_y = int(new_size_x), int(new_size_y)
        self.h = c_p(self.ctx.lib.rsm_pubmap_check_grid(self.pm))

    def InitMapWithRangeVec(self, store, ids, update_free_factor=0.0, update_occu_factor=0.0):
        """Reset + one update per stored scan, in list order (CorrectPoseAndMap's rebuild)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.ctx.check(self.ctx.lib.rsm_pubmap_rebuild(self.ctx.h, self.pm, store.h, len(ids), ids.ctypes.data,
                                                       float(update_free_factor), float(update_occu_factor)))

    def refresh_occupancy(self, occu_threshold=0.5, min_pass_through=2.0):
        self.ctx.check(self.ctx.lib.rsm_pubmap_refresh_occupancy(self.ctx.h, self.pm, float(occu_threshold), float(min_pass_through)))

    def download_all(self):
        shape = (self.size_y, self.size_x)
        prob, cnt, hit = (np.zeros(shape, dtype=np.float32) for _ in range(3))
        occ = np.zeros(shape, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.rsm_pubmap_download(self.ctx.h, self.pm, prob.ctypes.data, cnt.ctypes.data, hit.ctypes.data,
                                                        occ.ctypes.data))
        return prob, cnt, hit, occ


_D3, _D9 = ctypes.c_double * 3, ctypes.c_double * 9


class BasedCorrelationScanMatch:
    """One pass: ScanMatch(map, range_data, param, current_pose, cov_matrix) -> response.

    current_pose (3,) and cov_matrix (3,3) are numpy arrays updated IN PLACE, like the reference's
    Eigen references (correlate_scan_matcher.h:784-788)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.last_detail = None

    def ScanMatch(self, map_, range_data, scan_match_param, current_pose, cov_matrix):
        ctx = self.ctx
        assert current_pose.dtype == np.float64 and cov_matrix.dtype == np.float64
        assert current_pose.flags.c_contiguous and cov_matrix.flags.c_contiguous and current_pose.size == 3 and cov_matrix.size == 9
        ps = _param_struct(scan_match_param)
        resp = c_d(0)
        det = PassDetail()
        pose_p, cov_p = _D3.from_buffer(current_pose), _D9.from_buffer(cov_matrix)     # (cheaper than ndarray.ctypes.data)
        if isinstance(range_data, RangeDataContainer2d):
            ctx.check(ctx.lib.rsm_match_resident(ctx.h, map_.h, range_data.h, ctypes.byref(ps), pose_p, cov_p,
                                                 ctypes.byref(resp), ctypes.byref(det)))
        else:
            pts = _f64(range_data).reshape(-1, 2)
            ctx.check(ctx.lib.rsm_match(ctx.h, map_.h, pts.ctypes.data, len(pts), ctypes.byref(ps), pose_p, cov_p,
                                        ctypes.byref(resp), ctypes.byref(det)))
        self.last_detail = det
        return resp.value

    def scores(self, map_, range_data, scan_match_param, pose_world, angle_begin=0, angle_end=-1):
        """Penalised score of every candidate, candidate order (parity tests)."""
        ctx = self.ctx
        pts = _f64(range_data).reshape(-1, 2)
        p = _as_param(scan_match_param)
        ps = p.struct()
        n_ang = int(np.floor(ps.search_angle_offset * 2 / ps.search_angle_resolution) + 1)
        r = ps.search_space_size / ps.search_space_resolution
        n_xy = int((np.floor(r + 0.5) if r >= 0 else np.ceil(r - 0.5)) + 1)
        a1 = n_ang if angle_end < 0 else angle_end
        cap = max(1, (a1 - angle_begin) * n_xy * n_xy)
        out = np.empty(cap, dtype=np.float64)
        w = c_i64(0)
        pose = _f64(pose_world)
        ctx.check(ctx.lib.rsm_pass_scores(ctx.h, map_.h, pts.ctypes.data, len(pts), ctypes.byref(ps), pose.ctypes.data,
                                          int(angle_begin), int(angle_end), out.ctypes.data, cap, ctypes.byref(w)))
        return out[: w.value]


RSM_PARTIAL_BYTES = 65536
RSM_COLUMNS_BYTES = 131072


class OptimizeScanMatchParam:
    """Same fields as the reference class (optimize_scan_matcher.h:33-58)."""

    def __init__(self, iterate_max_times=10, cost_decrease_threshold=1.0, cost_min_threshold=2.0,
                 max_update_distance=0.5, max_update_angle=0.2):
        self.iterate_max_times = int(iterate_max_times)
        self.cost_decrease_threshold = float(cost_decrease_threshold)
        self.cost_min_threshold = float(cost_min_threshold)
        self.max_update_distance = float(max_update_distance)
        self.max_update_angle = float(max_update_angle)

    def struct(self):
        return OptimizeParamStruct(self.cost_decrease_threshold, self.cost_min_threshold, self.max_update_distance,
                                   self.max_update_angle, self.iterate_max_times, 0)


def _as_opt(p):
    return p if isinstance(p, OptimizeScanMatchParam) else OptimizeScanMatchParam(*p)


class BasedOptimizeScanMatch:
    """ScanMatch(map, range_data, optimize_scan_match_param, best_pose) -> cost; best_pose (numpy, world) is
    updated in place like the reference's Eigen::Vector3d& (optimize_scan_matcher.h:68-131)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.last_iterations = 0

    def ScanMatch(self, map_, range_data, optimize_scan_match_param, best_pose):
        ctx = self.ctx
        pts = _f64(range_data).reshape(-1, 2)
        st = _as_opt(optimize_scan_match_param).struct()
        cost, it = c_d(0), ctypes.c_int32(0)
        ctx.check(ctx.lib.rsm_optimize(ctx.h, map_.h, pts.ctypes.data, len(pts), ctypes.byref(st), best_pose.ctypes.data,
                                       ctypes.byref(cost), ctypes.byref(it)))
        self.last_iterations = it.value
        return cost.value

    def ScanMatchBatch(self, maps, scans, optimize_scan_match_param, poses):
        """n independent problems (rsm_optimize_batch) -> (costs, poses, iterations)."""
        ctx = self.ctx
        n = len(maps)
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, s_ in enumerate(scans):
            offs[i + 1] = offs[i] + len(s_)
        pts = _f64(np.concatenate([np.asarray(s_).reshape(-1, 2) for s_ in scans], axis=0)) if n else np.zeros((0, 2))
        poses = _f64(poses).copy()
        costs = np.zeros(n)
        iters = np.zeros(n, dtype=np.int32)
        st = _as_opt(optimize_scan_match_param).struct()
        handles = (c_p * n)(*[m.h for m in maps])
        ctx.check(ctx.lib.rsm_optimize_batch(ctx.h, n, handles, pts.ctypes.data, offs.ctypes.data, ctypes.byref(st),
                                             poses.ctypes.data, costs.ctypes.data, iters.ctypes.data))
        return costs, poses, iters


class SlicedScanMatch:
    """One large search window cut along the angle index over several GPUs.

    Every rank owns a Context and a copy of the grid.  `all_gather(buf)` must return the list of
    every rank's uint8 buffer in rank order (torch.distributed.all_gather over NCCL in bench.py,
    a plain list in the single-process tests).  All ranks return identical results."""

    def __init__(self, ctx, rank, world_size, all_gather):
        self.ctx, self.rank, self.world, self.all_gather = ctx, rank, world_size, all_gather
        self.last_detail = None
        self.exact_fallback = False     # the last match needed the gathered exact-tie path
        self.exchange = "caller-supplied all_gather"

    def close(self):
        pass

    def ScanMatch(self, map_, range_data, scan_match_param, current_pose, cov_matrix):
        from .sharding import contiguous_range
        ctx = self.ctx
        pts = _f64(range_data).reshape(-1, 2)
        ps = _as_param(scan_match_param).struct()
        n_ang = int(np.floor(ps.search_angle_offset * 2 / ps.search_angle_resolution) + 1)
        a0, a1 = contiguous_range(n_ang, self.rank, self.world)
        partial = np.zeros(RSM_PARTIAL_BYTES, dtype=np.uint8)
        pose = _f64(current_pose)
        ctx.check(ctx.lib.rsm_match_partial(ctx.h, map_.h, pts.ctypes.data, len(pts), ctypes.byref(ps),
                                            pose.ctypes.data, a0, a1, partial.ctypes.data))
        partials = [np.ascontiguousarray(p, dtype=np.uint8) for p in self.all_gather(partial)]
        pp = (c_p * len(partials))(*[p.ctypes.data for p in partials])
        columns = np.zeros(RSM_COLUMNS_BYTES, dtype=np.uint8)
        ctx.check(ctx.lib.rsm_match_merge(ctx.h, pp, len(partials), columns.ctypes.data))
        cols = [np.ascontiguousarray(c, dtype=np.uint8) for c in self.all_gather(columns)]
        cp = (c_p * len(cols))(*[c.ctypes.data for c in cols])
        resp = c_d(0)
        det = PassDetail()
        rc = ctx.lib.rsm_match_finish(ctx.h, pp, len(partials), cp, current_pose.ctypes.data,
                                      cov_matrix.ctypes.data, ctypes.byref(resp), ctypes.byref(det))
        self.exact_fallback = rc == RSM_NEED_EXACT
        if rc == RSM_NEED_EXACT:
            # Exact score ties in a consumed set (every rank takes this branch: the decision only reads gathered
            # data).  The reference's answer is then defined by its unstable sort of the WHOLE candidate array
            # (correlate_scan_matcher.h:607-611): gather the slices and run that sort on every rank.
            n_xy = det.n_xy if det.n_xy > 0 else int(np.floor(ps.search_space_size / ps.search_space_resolution + 0.5) + 1)
            plane = n_xy * n_xy
            counts = [(e - b) * plane for b, e in (contiguous_range(n_ang, r, self.world) for r in range(self.world))]
            cap = max(max(counts), 1)
            mine = np.zeros(cap, dtype=np.float64)
            n_mine = c_i64(0)
            ctx.check(ctx.lib.rsm_match_slice_scores(ctx.h, mine.ctypes.data, cap, ctypes.byref(n_mine)))
            assert n_mine.value == counts[self.rank], (n_mine.value, counts[self.rank])
            slices = [np.ascontiguousarray(sl).view(np.float64) if sl.dtype != np.float64 else np.ascontiguousarray(sl)
                      for sl in self.all_gather(mine.view(np.uint8))]
            sp = (c_p * len(slices))(*[sl.ctypes.data for sl in slices])
            cnt = np.array(counts, dtype=np.int64)
            rc = ctx.lib.rsm_match_finish_exact(ctx.h, sp, cnt.ctypes.data, len(slices), current_pose.ctypes.data,
                                                cov_matrix.ctypes.data, ctypes.byref(resp), ctypes.byref(det))
        ctx.check(rc)
        self.last_detail = det
        return resp.value


RSM_COMM_ID_BYTES = 128


class CommSlicedScanMatch:
    """One large window cut along the angle index over the ranks of a job, exchange inside the library over NCCL
    (rsm_comm_init / rsm_match_sliced).  `broadcast_id(buf)` must hand rank 0's 128-byte id (a uint8 numpy array,
    filled on rank 0) to every rank and return it; world_size 1 needs neither NCCL nor a broadcast."""

    def __init__(self, ctx, rank, world_size, broadcast_id=None):
        self.ctx, self.rank, self.world = ctx, rank, world_size
        self.last_detail = None
        self.exact_fallback = False
        self.exchange = "none (one rank)" if world_size == 1 else "in-library ncclAllGather on the context's stream (device buffers)"
        ident = np.zeros(RSM_COMM_ID_BYTES, dtype=np.uint8)
        if world_size > 1:
            if rank == 0:
                rc = ctx.lib.rsm_comm_unique_id(ident.ctypes.data)
                if rc != RSM_OK:
                    raise RsmError(rc, "rsm_comm_unique_id (libnccl.so.2 not loadable?)")
            ident = np.ascontiguousarray(broadcast_id(ident), dtype=np.uint8)
        ctx.check(ctx.lib.rsm_comm_init(ctx.h, int(rank), int(world_size), ident.ctypes.data))

    def ScanMatch(self, map_, range_data, scan_match_param, current_pose, cov_matrix):
        ctx = self.ctx
        pts = _f64(range_data).reshape(-1, 2)
        ps = _as_param(scan_match_param).struct()
        resp, det = c_d(0), PassDetail()
        ctx.check(ctx.lib.rsm_match_sliced(ctx.h, map_.h, pts.ctypes.data, len(pts), ctypes.byref(ps), current_pose.ctypes.data,
                                           cov_matrix.ctypes.data, ctypes.byref(resp), ctypes.byref(det)))
        self.last_detail = det
        self.exact_fallback = bool(det.exact_sort_used)
        return resp.value

    def close(self):
        if getattr(self.ctx, "h", None):
            self.ctx.lib.rsm_comm_destroy(self.ctx.h)


def scan_match_interface_batch_opt(ctx, fine_store, coarse_store, fine_spec, coarse_spec, centres, chains, match_ids, seed_poses,
                                   params, optimize_param, optimize_failed_cost, use_fine=True, pub_map=None, pub_store=None,
                                   check=None):
    """rsm_scan_match_interface_batch_opt: the batched back-end step with the Gauss-Newton pre-step on the coarse map
    (ScanMatchers::ScanMatch with use_optimize_scan_match, scan_matchers.h:205-232).  coarse_store holds the scans of
    fine_store (same ids) in coarse-map cells.  Returns (scores, poses, covs, responses[n, 4] = optimiser cost + the
    three pass responses)."""
    off, ids = chains if isinstance(chains, tuple) else pack_chains(chains)
    n = len(off) - 1
    mids = np.ascontiguousarray(match_ids, dtype=np.int32)
    arr = (PassParamStruct * 3)(*[_as_param(p).struct() for p in params])
    st = _as_opt(optimize_param).struct()
    poses = _f64(np.asarray(seed_poses).reshape(-1, 3)).copy()
    covs = np.tile(np.eye(3), (n, 1, 1))
    scores = np.zeros(n)
    resp = np.zeros((n, 4))
    c = _f64(np.asarray(centres).reshape(-1, 2))
    gf, gc = fine_spec, coarse_spec
    chk = None
    if pub_map is not None:
        chk = MapCheckParamStruct(float(check[1]), float(check[2]), int(check[0]), int(bool(check[3])))
    ctx.check(ctx.lib.rsm_scan_match_interface_batch_opt(
        ctx.h, fine_store.h, coarse_store.h, n, int(gf.size_x), float(gf.res), float(gf.sigma), int(gc.size_x), float(gc.res),
        float(gc.sigma), float(gf.default_prob), float(gf.occu_offset), c.ctypes.data, off.ctypes.data, ids.ctypes.data,
        mids.ctypes.data, arr, ctypes.byref(st), float(optimize_failed_cost), int(use_fine), poses.ctypes.data, covs.ctypes.data,
        scores.ctypes.data, resp.ctypes.data, pub_map.h if pub_map is not None else None,
        pub_store.h if pub_store is not None else None, ctypes.byref(chk) if chk is not None else None))
    return scores, poses, covs, resp


def make_sliced_matcher(ctx, rank, world_size, dist=None, in_library=True):
    """The angle-sliced matcher of a torch.distributed job (one process per GPU).  in_library: the exchange runs
    inside librsm.so over its own NCCL communicator (the id travels through dist.broadcast once); else the three-call
    protocol with `dist.all_gather` on host-staged byte tensors.  dist = None (world_size 1): no exchange."""
    if world_size == 1 or dist is None:
        if in_library:
            return CommSlicedScanMatch(ctx, rank, 1)
        sm = SlicedScanMatch(ctx, rank, world_size, lambda buf: [buf])
        sm.exchange = "none (one rank)"
        return sm
    import torch
    if in_library:
        def broadcast_id(buf):
            t = torch.from_numpy(buf.copy()).cuda()
            dist.broadcast(t, src=0)
            return t.cpu().numpy()
        return CommSlicedScanMatch(ctx, rank, world_size, broadcast_id)

    def all_gather(buf):
        t = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8)).cuda()
        outs = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(outs, t)
        return [o.cpu().numpy() for o in outs]

    sm = SlicedScanMatch(ctx, rank, world_size, all_gather)
    sm.exchange = "torch.distributed all_gather (NCCL) of host-staged buffers"
    return sm


class ScanMatchers:
    """Coarse -> fine -> super-fine chain: ScanMatch(range_data, map, best_pose, cov, use_fine) -> score.

    The reference's signature also carries the coarse-resolution map and scan, which only its
    optional Gauss-Newton pre-step reads (scan_matchers.h:205-213); all three correlative passes
    run on the fine map (:238-259), which is what is passed here."""

    def __init__(self, ctx, params):
        self.ctx = ctx
        self.SetScanMatchParam(params)
        self.last_responses = np.zeros(3)

    def SetScanMatchParam(self, params):
        assert len(params) == 3
        self.params = [_as_param(p) for p in params]
        self._arr = (PassParamStruct * 3)(*[p.struct() for p in self.params])

    def ScanMatch(self, range_data, map_, best_pose, cov_matrix, use_fine_scan_match=True):
        ctx = self.ctx
        pts = _f64(range_data).reshape(-1, 2)
        score = c_d(0)
        self.last_responses = np.zeros(3)
        ctx.check(ctx.lib.rsm_match_chain(ctx.h, map_.h, pts.ctypes.data, len(pts), self._arr, int(use_fine_scan_match),
                                          best_pose.ctypes.data, cov_matrix.ctypes.data, ctypes.byref(score),
                                          self.last_responses.ctypes.data))
        return score.value

    def ScanMatchWithOptimize(self, coarse_map_range_data, fine_map_range_data, coarse_map, fine_map, best_pose, cov_matrix,
                              optimize_scan_match_param, optimize_failed_cost, use_fine_scan_match=True):
        """The reference's full signature with use_optimize_scan_match on (scan_matchers.h:179-289):
        Gauss-Newton on the coarse map first, the correlative passes on the fine map.  -> score;
        self.last_optimize_cost / self.last_responses hold the step results."""
        ctx = self.ctx
        pc = _f64(coarse_map_range_data).reshape(-1, 2)
        pf = _f64(fine_map_range_data).reshape(-1, 2)
        st = _as_opt(optimize_scan_match_param).struct()
        score = c_d(0)
        resp = np.zeros(4)
        ctx.check(ctx.lib.rsm_match_chain_opt(ctx.h, coarse_map.h, pc.ctypes.data, len(pc), fine_map.h, pf.ctypes.data, len(pf),
                                              self._arr, ctypes.byref(st), float(optimize_failed_cost), int(use_fine_scan_match),
                                              best_pose.ctypes.data, cov_matrix.ctypes.data, ctypes.byref(score),
                                              resp.ctypes.data))
        self.last_optimize_cost = resp[0]
        self.last_responses = resp[1:].copy()
        return score.value

    def ScanMatchBatch(self, maps, scans, poses, covs=None, use_fine_scan_match=True):
        """n independent chains in batched launches (rsm_match_batch). poses (n,3) in/out."""
        ctx = self.ctx
        n = len(maps)
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, s in enumerate(scans):
            offs[i + 1] = offs[i] + len(s)
        pts = _f64(np.concatenate([np.asarray(s).reshape(-1, 2) for s in scans], axis=0))
        poses = _f64(poses).copy()
        covs = np.tile(np.eye(3), (n, 1, 1)) if covs is None else _f64(covs).copy()
        scores = np.zeros(n)
        resp = np.zeros((n, 3))
        handles = (c_p * n)(*[m.h for m in maps])
        ctx.check(ctx.lib.rsm_match_batch(ctx.h, n, handles, pts.ctypes.data, offs.ctypes.data, self._arr, 1,
                                          int(use_fine_scan_match), poses.ctypes.data, covs.ctypes.data,
                                          scores.ctypes.data, resp.ctypes.data))
        return scores, poses, covs, resp


def pack_loop_closure(scenarios):
    """Flatten a list of synth.Scenario (same grid size / resolution / params) for rsm_loop_closure_batch."""
    n = len(scenarios)
    g0 = scenarios[0].grid
    centres = np.zeros((n, 2))
    scan_off = np.zeros(n + 1, dtype=np.int64)
    pts_off = np.zeros(n + 1, dtype=np.int64)
    base_n, base_pts, base_poses, pts = [], [], [], []
    for i, sc in enumerate(scenarios):
        g = sc.grid
        assert (g.size_x, g.size_y, g.res) == (g0.size_x, g0.size_x, g0.res)
        assert sc.grid_centre is not None, "loop-closure batches need back-end grids (centred on a pose)"
        centres[i] = sc.grid_centre
        scan_off[i + 1] = scan_off[i] + len(sc.base_pts)
        for bp, pose in zip(sc.base_pts, sc.base_poses):
            base_n.append(len(bp))
            base_pts.append(np.asarray(bp).reshape(-1, 2))
            base_poses.append(pose)
        pts_off[i + 1] = pts_off[i] + len(sc.scan_pts)
        pts.append(np.asarray(sc.scan_pts).reshape(-1, 2))
    return dict(n=n, grid_size=g0.size_x, resolution=g0.res, default_prob=g0.default_prob, sigma=g0.sigma,
                occu_offset=g0.occu_offset, centres=_f64(centres), scan_off=scan_off,
                base_n=np.array(base_n, dtype=np.int32), base_pts=_f64(np.concatenate(base_pts, axis=0)),
                base_poses=_f64(np.array(base_poses)), pts=_f64(np.concatenate(pts, axis=0)), pts_off=pts_off,
                poses=_f64(np.array([sc.seed_pose for sc in scenarios])))


def loop_closure_batch(ctx, packed, params, use_fine=True, centres=None):
    """rsm_loop_closure_batch on arrays from pack_loop_closure; returns (scores, poses, covs, responses)."""
    n = packed["n"]
    arr = (PassParamStruct * 3)(*[_as_param(p).struct() for p in params])
    poses = packed["poses"].copy()
    covs = np.tile(np.eye(3), (n, 1, 1))
    scores = np.zeros(n)
    resp = np.zeros((n, 3))
    c = packed["centres"] if centres is None else _f64(centres)
    ctx.check(ctx.lib.rsm_loop_closure_batch(
        ctx.h, n, int(packed["grid_size"]), float(packed["resolution"]), float(packed["default_prob"]),
        float(packed["sigma"]), float(packed["occu_offset"]), c.ctypes.data, packed["scan_off"].ctypes.data,
        packed["base_n"].ctypes.data, packed["base_pts"].ctypes.data, packed["base_poses"].ctypes.data,
        packed["pts"].ctypes.data, packed["pts_off"].ctypes.data, arr, int(use_fine), poses.ctypes.data,
        covs.ctypes.data, scores.ctypes.data, resp.ctypes.data))
    return scores, poses, covs, resp


class ScanStore:
    """Device-resident scans addressed by id: one resolution of the reference's
    SensorDataManager::multiresolution_range_data_[name] (sensor_data_manager.h:514-525)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = c_p()
        ctx.check(ctx.lib.rsm_scan_store_create(ctx.h, ctypes.byref(h)))
        self.h = h

    def AddRangeData(self, pts_cells, sensor_pose):
        pts = _f64(np.asarray(pts_cells).reshape(-1, 2))
        pose = _f64(sensor_pose)
        out = ctypes.c_int32(-1)
        self.ctx.check(self.ctx.lib.rsm_scan_store_add(self.ctx.h, self.h, pts.ctypes.data, len(pts), pose.ctypes.data,
                                                       ctypes.byref(out)))
        return out.value

    def UpdateRangeData(self, ids, sensor_poses):
        """SlamProcessor::UpdateRangeData (slam_processor.cpp:597-603) for a list of ids."""
        ids = np.ascontiguousarray(np.atleast_1d(ids), dtype=np.int32)
        poses = _f64(np.asarray(sensor_poses).reshape(-1, 3))
        assert len(ids) == len(poses)
        self.ctx.check(self.ctx.lib.rsm_scan_store_set_poses(self.ctx.h, self.h, len(ids), ids.ctypes.data, poses.ctypes.data))

    def sensor_pose(self, scan_id):
        out = np.zeros(3)
        if self.ctx.lib.rsm_scan_store_get_pose(self.h, int(scan_id), out.ctypes.data) != 0:
            raise RsmError(3, "unknown scan id %d" % scan_id)
        return out

    def __len__(self):
        return int(self.ctx.lib.rsm_scan_store_size(self.h))

    def close(self):
        if self.h:
            self.ctx.lib.rsm_scan_store_destroy(self.ctx.h, self.h)
            self.h = None


def pack_chains(chains):
    """[[ids of chain 0], [ids of chain 1], ...] -> (chain_offset int64[n + 1], chain_ids int32[...]) for
    scan_match_interface_batch; a caller that matches the same chains repeatedly packs them once."""
    n = len(chains)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, ch in enumerate(chains):
        off[i + 1] = off[i] + len(ch)
    ids = np.ascontiguousarray(np.concatenate([np.asarray(ch, dtype=np.int32) for ch in chains]) if n else np.zeros(0), dtype=np.int32)
    return off, ids


def scan_match_interface_batch(ctx, store, grid_spec, centres, chains, match_ids, seed_poses, params, use_fine=True,
                               pub_map=None, pub_store=None, check=None):
    """SlamProcessor::ScanMatchInterface (slam_processor.cpp:250-326) for many loop-closure candidates:
    candidate i matches scan match_ids[i] of `store` against the chain `chains[i]` (list of scan ids; or the
    (offsets, ids) pair of pack_chains) on a back-end grid of grid_spec's size / resolution / blur centred on
    centres[i].  With pub_map / pub_store / check = (check_point_num, bound_tolerance, penalty_gain, use_logistic)
    the scores end with the map check.  Returns (scores, poses, covs, responses)."""
    if isinstance(chains, tuple):
        off, ids = chains
    else:
        off, ids = pack_chains(chains)
    n = len(off) - 1
    mids = np.ascontiguousarray(match_ids, dtype=np.int32)
    arr = (PassParamStruct * 3)(*[_as_param(p).struct() for p in params])
    poses = _f64(np.asarray(seed_poses).reshape(-1, 3)).copy()
    covs = np.tile(np.eye(3), (n, 1, 1))
    scores = np.zeros(n)
    resp = np.zeros((n, 3))
    c = _f64(np.asarray(centres).reshape(-1, 2))
    g = grid_spec
    chk = None
    if pub_map is not None:
        chk = MapCheckParamStruct(float(check[1]), float(check[2]), int(check[0]), int(bool(check[3])))
    ctx.check(ctx.lib.rsm_scan_match_interface_batch(
        ctx.h, store.h, n, int(g.size_x), float(g.res), float(g.default_prob), float(g.sigma), float(g.occu_offset),
        c.ctypes.data, off.ctypes.data, ids.ctypes.data, mids.ctypes.data, arr, int(use_fine), poses.ctypes.data,
        covs.ctypes.data, scores.ctypes.data, resp.ctypes.data,
        pub_map.h if pub_map is not None else None, pub_store.h if pub_store is not None else None,
        ctypes.byref(chk) if chk is not None else None))
    return scores, poses, covs, resp
