"""Where does SimpSolver-mode time go?  preprocess-only vs full solve on the config-2 batch (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200pkg
pkg = _b200pkg.load()
n_inst, n = int(sys.argv[1]), int(sys.argv[2])
m = int(round(4.26 * n))
eng = pkg.Engine(0)
eng.generate_ksat(0x5EED000000000000 if n == 100 else 0x5EED000100000000, n_inst, n, m, 3)
for name, kind, flags in (("core full", pkg.CORE, pkg.F_PREPROCESS), ("core ingest only", pkg.CORE, pkg.F_PREPROCESS | pkg.F_NO_SOLVE),
                          ("simp preprocess only", pkg.SIMP, pkg.F_PREPROCESS | pkg.F_NO_SOLVE), ("simp full", pkg.SIMP, pkg.F_PREPROCESS)):
    best = 1e9
    for it in range(2):
        eng.solve(kind=kind, flags=flags)
        best = min(best, eng.launch_info()["last_kernel_ms"])
    res, _ = eng.download(want_models=False)
    print("%-22s %8.2f ms   conflicts max %d mean %.1f  props %.3e" % (name, best, res["stats"][:, 4].max(), res["stats"][:, 4].mean(), res["stats"][:, 5].sum()), flush=True)
