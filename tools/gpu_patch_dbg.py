"""Debug aid (needs a -DRSM_STAGED_DEBUG build, RSM_LIB_PATH): how often does a 32-beam chunk of the patch kernel fit its box?"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roborts_edu_slam_b200 import synth, matcher
ctx = matcher.Context(0)
lib = ctypes.CDLL(matcher.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
pairs = synth.config4(64)
packed = matcher.pack_loop_closure(pairs)
lib.rsm_debug_patch(out, 1)
matcher.loop_closure_batch(ctx, packed, pairs[0].passes)
lib.rsm_debug_patch(out, 1)
v = list(out)
print("chunks", v[0], "fit", v[1], "= %.1f%%" % (100.0 * v[1] / max(1, v[0])), "mean hull %.1f x %.1f" % (v[2] / max(1, v[0]), v[3] / max(1, v[0])))
