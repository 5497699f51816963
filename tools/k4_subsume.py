"""K4 multi-CTA subsumption kernels on the structured tests/cnf-hard fixture: clauses removed + kernel time per instance,
beside the oracle's full SimpSolver preprocessing (subsumption + strengthening + BVE) result."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from common import batch_of
g = np.load(os.path.join(ROOT, "tests", "golden", "corpus_hard.npz"))
eng = pkg.Engine(0)
for i, name in enumerate(g["names"]):
    cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
    off = g["clause_off"][cb:ce + 1].astype(np.int64)
    lits = g["lits"][off[0]:off[-1]]
    ioff, coff, li = batch_of([((off - off[0]).astype(np.uint32), lits)])
    eng.upload(ioff, coff, li)
    best = None
    for it in range(3):
        r = eng.subsume_instance(0, len(off) - 1, len(lits))
        best = r["kernel_ms"] if best is None else min(best, r["kernel_ms"])
    r2 = eng.simplify_instance(0, len(off) - 1, len(lits))
    r3 = eng.bve_instance(0, len(off) - 1, len(lits))
    print(json.dumps({"instance": str(name), "k4_bve": {"eliminated": int(r3["eliminated"]), "rounds": int(r3["rounds"]), "kept": int(r3["kept"]),
                                                        "kept_lits": int(r3["kept_lits"]), "kernel_ms": round(r3["kernel_ms"], 3)},
                      "oracle_eliminated_vars": int(g["rows"][i][3]), "oracle_clauses_left": int(g["rows"][i][2])}), flush=True)
    print(json.dumps({"instance": str(name), "k4_strengthen": {k: (int(v) if not isinstance(v, (bool, float)) else v) for k, v in r2.items() if k in ("kept", "kept_lits", "rounds", "strengthened", "unsat", "removed_count", "kernel_ms")}, "clauses": len(off) - 1, "literals": int(len(lits)), "k4_removed": int(r["removed_count"]),
                      "k4_kept": int(r["kept"]), "subset_tests": int(r["tests"]), "k4_kernels_ms": round(best, 3),
                      "oracle_full_preprocess_clauses_left": int(g["rows"][i][2]), "oracle_eliminated_vars": int(g["rows"][i][3])}), flush=True)
eng.close()
