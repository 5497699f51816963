"""One solve of a config-2/3-style batch with a fixed launch shape (profiling aid).
usage: gpu_one.py n_inst n_vars warps_per_block blocks_per_sm [conflict_budget]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200pkg
pkg = _b200pkg.load()
n_inst, n, wpb, bps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
budget = int(sys.argv[5]) if len(sys.argv) > 5 else -1
m = int(round(4.26 * n))
eng = pkg.Engine(0)
eng.generate_ksat(0x5EED000000000000 if n == 100 else 0x5EED000100000000, n_inst, n, m, 3)
eng.configure(warps_per_block=wpb, blocks_per_sm=bps)
if budget >= 0:
    eng.set_budget(conflict_budget=budget)
for it in range(2):
    eng.solve(kind=pkg.CORE, flags=pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT)
info = eng.launch_info()
res, _ = eng.download(want_models=False)
print(info, "props/s %.3e" % (res["stats"][:, 5].sum() / info["search_ms"] * 1e3), "conflicts/s %.3e" % (res["stats"][:, 4].sum() / info["search_ms"] * 1e3),
      "interrupted %d" % int((res["verdict"] == 2).sum()))
