"""BASELINE.json configs[3]: SimpSolver preprocessing (subsumption + BVE) of structured tests/cnf-hard instances on the
GPU (reference-order device simplifier, --no-solve) vs the oracle's elimination result stored in
tests/golden/corpus_hard.npz.  Prints one JSON line per instance."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from common import batch_of
g = np.load(os.path.join(ROOT, "tests", "golden", "corpus_hard.npz"))
only = sys.argv[1:]
eng = pkg.Engine(0)
for i, name in enumerate(g["names"]):
    if only and not any(o in str(name) for o in only):
        continue
    cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
    off = g["clause_off"][cb:ce + 1].astype(np.int64)
    lits = g["lits"][off[0]:off[-1]]
    ioff, coff, li = batch_of([((off - off[0]).astype(np.uint32), lits)])
    t0 = time.perf_counter()
    res, _ = eng.solve_batch(ioff, coff, li, kind=pkg.SIMP, flags=pkg.F_PREPROCESS | pkg.F_NO_SOLVE, model_stride=0)
    dt = time.perf_counter() - t0
    r, row = res[0], [int(x) for x in g["rows"][i]]
    mine = [int(r["n_vars"]), int(r["n_clauses_ingest"]), int(r["n_clauses_pre"]), int(r["eliminated_vars"]), int(r["pre_ok"])] + [int(x) for x in r["stats"]]
    print(json.dumps({"instance": str(name), "status": int(r["status"]), "attempts": int(r["attempts"]), "gpu_seconds": round(dt, 2),
                      "kernel_ms": eng.launch_info()["last_kernel_ms"], "vars": mine[0], "clauses_in": mine[1], "clauses_after": mine[2],
                      "eliminated_vars": mine[3], "equals_oracle": mine == row, "oracle": row[:5]}), flush=True)
eng.close()
