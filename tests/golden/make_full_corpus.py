"""Generates tests/golden/corpus_full_cnf.npz and corpus_full_hard.npz: EVERY file of the reference's bundled corpus
(/root/reference/tests/cnf: 3,023 files; tests/cnf-hard: 198 files) as CSR with the ORACLE's verdict, the 8 stats and the
trace hash for CoreSolver and SimpSolver at default settings -- the reference's own test walks every 4th file of the
same directories (tests/minisat.rs:12-37,173-210).

Run in the build container (needs /root/reference):  python tests/golden/make_full_corpus.py [--jobs 8]

Time box: a cnf-hard file is packed with stats only when the oracle needs at most MAX_CONFLICTS conflicts with BOTH
solvers (one warp has to finish it inside the GPU test's time); the others are listed in `excluded` with the reason
and stay covered by the preprocessing-only fixture (corpus_hard.npz).  SATLIB names pin the verdict (uf* SAT, uuf* UNSAT).
"""
import glob
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/tests"
MAX_CONFLICTS = 500000


def solve_one(rel):
    from oracle import pyoracle as po
    off, lits, hv, hc = po.parse_dimacs(os.path.join(REF, rel))
    po.set_budget(MAX_CONFLICTS, -1, 0)
    rows = {}
    for kind, nm in ((po.CORE, "core"), (po.SIMP, "simp")):
        r, m = po.solve_csr(off, lits, kind=kind)
        if r.verdict == po.SAT:
            assert po.check_model(off, lits, m), rel
        rows[nm] = (r.verdict, list(r.stats), r.trace_hash, r.eliminated_vars, r.n_clauses_pre, r.n_vars)
    return rel, off, lits, rows


def pack(rels, out_name, jobs):
    with mp.Pool(jobs) as pool:
        results = pool.map(solve_one, rels, chunksize=4)
    names, insts, excluded = [], [], []
    out = {k: [] for k in ("core_verdict", "core_stats", "core_hash", "simp_verdict", "simp_stats", "simp_hash", "simp_elim",
                           "simp_nclauses_pre", "n_vars", "name_verdict")}
    for rel, off, lits, rows in results:
        b = os.path.basename(rel)
        if rows["core"][0] == 2 or rows["simp"][0] == 2:
            excluded.append("%s: more than %d conflicts (core %d, simp %d at the cut)" % (b, MAX_CONFLICTS, rows["core"][1][4], rows["simp"][1][4]))
            continue
        nv_name = 1 if b.startswith("uf") else 0 if b.startswith("uuf") else -1
        if nv_name >= 0:
            assert rows["core"][0] == nv_name and rows["simp"][0] == nv_name, b   # SATLIB file name pins the verdict
        names.append(b); insts.append((off, lits))
        out["core_verdict"].append(rows["core"][0]); out["core_stats"].append(rows["core"][1]); out["core_hash"].append(rows["core"][2])
        out["simp_verdict"].append(rows["simp"][0]); out["simp_stats"].append(rows["simp"][1]); out["simp_hash"].append(rows["simp"][2])
        out["simp_elim"].append(rows["simp"][3]); out["simp_nclauses_pre"].append(rows["simp"][4]); out["n_vars"].append(rows["core"][5])
        out["name_verdict"].append(nv_name)
    ioff, coffs, alll, base = [0], [np.zeros(1, np.uint32)], [], 0
    for off, lits in insts:
        coffs.append((off[1:].astype(np.uint64) + base).astype(np.uint32))
        alll.append(lits)
        base += int(off[-1])
        ioff.append(ioff[-1] + len(off) - 1)
    lits_all = np.concatenate(alll)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), out_name),
                        names=np.array(names), excluded=np.array(excluded), inst_clause_off=np.array(ioff, np.uint32),
                        clause_off=np.concatenate(coffs),
                        lits=lits_all.astype(np.uint16) if lits_all.max() < 65536 else lits_all,
                        **{k: np.array(v, dtype=np.uint64 if "hash" in k or "stats" in k else np.int64) for k, v in out.items()})
    print("wrote", len(names), "instances to", out_name, "; excluded", len(excluded))
    for e in excluded:
        print("  excluded", e)


if __name__ == "__main__":
    jobs = int(sys.argv[sys.argv.index("--jobs") + 1]) if "--jobs" in sys.argv else 8
    cnf = sorted(os.path.relpath(f, REF) for f in glob.glob(os.path.join(REF, "cnf", "*.cnf.gz")))
    hard = sorted(os.path.relpath(f, REF) for f in glob.glob(os.path.join(REF, "cnf-hard", "*.cnf.gz")))
    if "--hard-only" not in sys.argv:
        pack(cnf, "corpus_full_cnf.npz", jobs)
    pack(hard, "corpus_full_hard.npz", jobs)
