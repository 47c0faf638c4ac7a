"""B200-native correlative scan matcher: the RoboRTS-Edu-SLAM hot path (src/scan_match +
lookup-grid construction) as hand-written sm_100a CUDA behind a C ABI (include/rsm.h).

    from roborts_edu_slam_b200 import matcher
    ctx = matcher.Context(0)                      # needs a CUDA device; there is no CPU path
    grid = matcher.ScanMatchMap(ctx, res, nx, ny, off_x, off_y); grid.InitMapWithRangeVec(...)
    response = matcher.BasedCorrelationScanMatch(ctx).ScanMatch(grid, scan, param, pose, cov)
"""
from . import matcher  # noqa: F401

__all__ = ["matcher"]
