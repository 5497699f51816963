"""Quick GPU sanity + timing script (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from oracle import pyoracle as po

def run(n_inst, n, m, seed, check=200, kind=0):
    eng = pkg.Engine(0)
    eng.generate_ksat(seed, n_inst, n, m, 3)
    flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT | pkg.F_TRACE_HASH
    for it in range(3):
        t = time.time(); rc = eng.solve(kind=kind, flags=flags); dt = time.time() - t
        info = eng.launch_info()
        print("solve rc", rc, "wall %.4f s" % dt, info, flush=True)
    res, models = eng.download(model_stride=n)
    st = res["stats"]
    print("n_inst", n_inst, "n", n, "inst/s (kernel)", n_inst / (info["last_kernel_ms"] * 1e-3), "props/s", st[:, 5].sum() / (info["last_kernel_ms"] * 1e-3),
          "sat", int((res["verdict"] == 1).sum()), "status!=0", int((res["status"] != 0).sum()), "attempts max", int(res["attempts"].max()))
    ioff, coff, lits = eng.download_batch()
    k = min(check, n_inst)
    sub_ioff = ioff[:k + 1]
    ores, _, secs = po.solve_batch(sub_ioff, coff, lits, kind=kind, threads=4)
    bad = 0
    for i in range(k):
        if res[i]["verdict"] != ores[i].verdict or list(st[i]) != list(ores[i].stats) or int(res[i]["trace_hash"]) != ores[i].trace_hash:
            bad += 1
            if bad < 5: print("MISMATCH", i, res[i]["verdict"], ores[i].verdict, list(st[i]), list(ores[i].stats))
        if res[i]["verdict"] == 1:
            cb, ce = ioff[i], ioff[i + 1]
            if not po.check_model(coff[cb:ce + 1], lits, models[i]): print("BAD MODEL", i); bad += 1
    print("checked", k, "bad", bad, "oracle %.1f inst/s (4 threads)" % (k / secs), flush=True)
    eng.close()
    return bad

if __name__ == "__main__":
    bad = run(256, 50, 213, 0x5EED000200000000)
    bad += run(10000, 100, 426, 0x5EED000000000000, check=300)
    bad += run(64, 250, 1065, 0x5EED000100000000, check=0)
    sys.exit(1 if bad else 0)
