"""Developer tool: the arithmetic-only micro-kernel (te_idm_peak) at w warps per SM -> updates per clock per SM and
the implied latency of one dependent IDM update (w = 4: one warp per scheduler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traffic_env_b200.vec_env import idm_arithmetic_peak
for w in (4, 8, 12, 16, 20, 24, 28, 32):
    os.environ["TE_PEAK_WARPS_PER_SM"] = str(w)
    r = idm_arithmetic_peak(iters=4000)
    per_clk_sm = r / 148 / 1.965e9
    # w/4 warps per scheduler, each issuing one update of 32 cars per L cycles: r = 148 * w * 32 / L * clk
    L = 148 * w * 32 * 1.965e9 / r
    print("warps/SM %2d  %.3e updates/s  %.3f updates/clk/SM  cycles per warp-update %.0f" % (w, r, per_clk_sm, L))
for mode, what in (("1", "two idm_update calls per lane"), ("2", "split fast path, one car per lane"),
                   ("3", "split fast path, two cars per lane (first parts in one basic block)"),
                   ("4", "idm_update<FA = true> (the step kernel's form), one car per lane")):
    os.environ["TE_PEAK_ILP2"] = mode
    for w in (4, 8, 12, 16, 20, 24, 32):
        os.environ["TE_PEAK_WARPS_PER_SM"] = str(w)
        r = idm_arithmetic_peak(iters=4000)
        print("%s, warps/SM %2d  %.3e updates/s" % (what, w, r))
del os.environ["TE_PEAK_ILP2"]
del os.environ["TE_PEAK_WARPS_PER_SM"]
print("full occupancy: %.3e" % idm_arithmetic_peak(iters=4000))
