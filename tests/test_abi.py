"""CPU tests: the C-ABI library loads and exports exactly what include/rsm.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from roborts_edu_slam_b200 import matcher, synth
from roborts_edu_slam_b200.sharding import angle_slices, contiguous_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_library_agree():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(matcher.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "librsm.so does not export %s" % n
    assert sorted(matcher.ABI.keys()) == names, "matcher.ABI and include/rsm.h differ"


def test_library_binds_and_reports_version():
    lib = matcher.load_library()
    assert lib.rsm_version().startswith(b"rsm ")


def test_struct_layouts():
    assert ctypes.sizeof(matcher.PassParamStruct) == 56
    assert ctypes.sizeof(matcher.PassDetail) == 72
    assert ctypes.sizeof(matcher.Stats) == 144


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(matcher.RsmError) as e:
        matcher.Context(0)
    assert e.value.status == 1   # RSM_ERR_NO_DEVICE


def test_param_mirror():
    p = matcher.CorrelationScanMatchParam()
    p.set_search_space_size(0.6)
    p.set_search_space_resolution(0.05)
    p.set_search_angle_offset(0.349)
    p.set_search_angle_resolution(0.0349)
    p.set_response_threshold(0.6)
    p.set_use_point_size(100)
    p.set_use_center_penalty(True)
    p.set_correlation_scan_match_type(matcher.FINE_CORRELATION_SCAN_MATCH)
    s = p.struct()
    assert (s.search_space_size, s.search_space_resolution, s.use_point_size, s.type) == (0.6, 0.05, 100, 1)
    assert p.search_angle_offset() == 0.349
    q = matcher.CorrelationScanMatchParam.from_array(synth.pass_param(2.0, 0.025, 0.7854, 0.0087266, 0.6, 100000, True, 0))
    assert q.struct().use_point_size == 100000 and q.struct().use_center_penalty == 1


def test_synth_is_deterministic():
    a, b = synth.config1(), synth.config1()
    assert np.array_equal(a.scan_pts, b.scan_pts) and len(a.scan_pts) == 360
    assert all(np.array_equal(x, y) for x, y in zip(a.base_pts, b.base_pts))
    assert (a.grid.size_x, a.grid.size_y) == (480, 480)
    p4 = synth.config4(3)
    q4 = synth.config4(2, first=1)
    assert np.array_equal(p4[1].scan_pts, q4[0].scan_pts) and np.array_equal(p4[2].seed_pose, q4[1].seed_pose)


def test_pack_loop_closure_layout():
    pairs = synth.config4(3)
    pk = matcher.pack_loop_closure(pairs)
    assert pk["n"] == 3 and pk["grid_size"] == 480
    assert pk["scan_off"].tolist() == [0, 8, 16, 24]
    assert pk["pts_off"][-1] == sum(len(s.scan_pts) for s in pairs)
    assert pk["base_n"].sum() == len(pk["base_pts"])
    assert np.array_equal(pk["centres"][1], pairs[1].truth_pose[:2])


def test_sharding_ranges():
    for n in (0, 1, 7, 512, 4096, 4099):
        for w in (1, 2, 4, 8):
            r = [contiguous_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert angle_slices(721, 8)[-1] == (631, 721)
