"""How the path shards over the GPUs of one box.

Batched loop closure: pairs are independent, so the pair list is cut into contiguous ranges, one
per rank; no data-path collective, results are only gathered at the end.  A single large window
is cut along the angle index instead.
"""


def contiguous_range(n_items, rank, world_size):
    """Items [begin, end) owned by `rank`: sizes differ by at most one, earlier ranks get the extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def angle_slices(n_ang, world_size):
    """[(begin, end)] per rank over angle indices [0, n_ang)."""
    return [contiguous_range(n_ang, r, world_size) for r in range(world_size)]
