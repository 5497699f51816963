"""Launch-shape sweep on the config-2 batch (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
n_inst, n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 100, 0
m = int(round(4.26 * n))
eng = pkg.Engine(0)
eng.generate_ksat(0x5EED000000000000 if n == 100 else 0x5EED000100000000, n_inst, n, m, 3)
flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT
for wpb, bps in [(1, 2), (1, 4), (1, 6), (1, 8), (1, 12), (1, 16), (1, 24), (2, 16), (1, 32)]:
    eng.configure(warps_per_block=wpb, blocks_per_sm=bps)
    best = 1e9
    for it in range(2):
        eng.solve(kind=pkg.CORE, flags=flags)
        info = eng.launch_info()
        best = min(best, info["last_kernel_ms"])
    res, _ = eng.download(want_models=False)
    print("wpb %d bps %2d warps %5d grid %5d: %8.2f ms  %9.0f inst/s  props/s %.3e" % (
        wpb, bps, info["warps_total"], info["grid"], best, n_inst / best * 1e3, res["stats"][:, 5].sum() / best * 1e3), flush=True)
