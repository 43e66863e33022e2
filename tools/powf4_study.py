import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200 import _lib
L = _lib.load()
for tau in (0, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
    out = np.zeros(4, np.uint64)
    _lib.check(L.te_test_powf4_exhaustive(0, tau, out.ctypes.data))
    print("tau=2^%s: differ=%d max_dist=%d (2^%.2f) declined=%d (%.4f%%) accepted_but_wrong=%d" % (
        np.log2(tau) if tau else "-inf", out[0], out[1], np.log2(max(int(out[1]), 1)), out[2], 100.0 * out[2] / 0x7f800000, out[3]))
