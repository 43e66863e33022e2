import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200 import _lib
L = _lib.load()
for c in (13.89, 20.0, 3.0, 7.77, 1.0, 33.3):
    out = np.zeros(3, np.uint64)
    _lib.check(L.te_test_fdiv_const_exhaustive(0, c, out.ctypes.data))
    print("c=%g: mismatches=%d normal-quotient mismatches=%d largest mismatching v=%g" % (
        c, out[0], out[1], np.array([out[2]], np.uint64).astype(np.uint32).view(np.float32)[0]))
