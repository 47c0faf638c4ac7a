"""Derive the occupancy fixtures from the reference's Stage floor plans.

Run in the build container only (needs /root/reference):
    python tests/golden/make_maps.py
Reads maps/{icra,rm,willow-full-0.05}.pgm (0.05 m/px, maps/*.yaml:2), thresholds them the way
SURVEY.md section 8(d) defines the synthetic workloads (pixel < 128 = occupied), flips rows so
that row 0 is the bottom of the map (origin bottom-left) and stores the bit-packed masks in
tests/golden/maps_occ.npz.  Only the occupancy mask is kept; the grey levels are not needed to
ray-cast scans.
"""
import os
import sys

import numpy as np

REF = os.environ.get("RSM_REFERENCE", "/root/reference")
NAMES = {"icra": "icra.pgm", "rm": "rm.pgm", "willow": "willow-full-0.05.pgm"}


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    # P5 header: magic, width, height, maxval separated by whitespace, comments start with '#'
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    pos += 1
    assert tokens[0] == b"P5", tokens
    w, h, maxval = int(tokens[1]), int(tokens[2]), int(tokens[3])
    assert maxval < 256
    return np.frombuffer(data, dtype=np.uint8, count=w * h, offset=pos).reshape(h, w)


def main():
    out = {}
    for key, fn in NAMES.items():
        img = read_pgm(os.path.join(REF, "maps", fn))
        occ = (img < 128)[::-1, :]  # row 0 = bottom
        out[key + "_shape"] = np.array(occ.shape, dtype=np.int32)
        out[key + "_bits"] = np.packbits(occ, axis=None)
        print(key, occ.shape, "occupied fraction %.4f" % occ.mean())
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps_occ.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
