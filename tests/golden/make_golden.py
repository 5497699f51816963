"""Generates tests/golden/corpus_sample.npz: a small sample of the reference's bundled
instances (/root/reference/tests/cnf*, SATLIB data) as CSR, with the ORACLE's verdict,
stats and trace hash for CoreSolver and SimpSolver runs (default settings).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The reference stores no golden answers (its test compares against a live `minisat`
binary), so the stats here are the oracle's own output -- "parity unpinned" for the
counters; the SATLIB-name verdict (uf* SAT, uuf* UNSAT) is the pinned part.
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference/tests"
PICK = (["cnf/uf20-0%d.cnf.gz" % i for i in (1, 2, 3, 4, 5, 6, 7, 8)] +
        ["cnf/uf50-0%d.cnf.gz" % i for i in (1, 2, 3, 4, 5, 6, 7, 8)] +
        ["cnf/uuf50-0%d.cnf.gz" % i for i in (1, 2, 3, 4, 5, 6, 7, 8)] +
        ["cnf-hard/uf250-083.cnf.gz", "cnf-hard/uf250-021.cnf.gz", "cnf-hard/uf250-068.cnf.gz",
         "cnf-hard/uf250-037.cnf.gz", "cnf-hard/uf250-04.cnf.gz",
         "cnf/2bitcomp_5.cnf.gz", "cnf/2bitmax_6.cnf.gz", "cnf/3blocks.cnf.gz", "cnf/2bitadd_11.cnf.gz",
         "cnf/2bitadd_12.cnf.gz", "cnf/4blocksb.cnf.gz"])


def build(pick, out_name):
    names, insts = [], []
    out = {k: [] for k in ("core_verdict", "core_stats", "core_hash", "core_traffic", "simp_verdict", "simp_stats",
                           "simp_hash", "simp_elim", "simp_nclauses_pre", "n_vars", "name_verdict")}
    for rel in pick:
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            print("skip", rel)
            continue
        off, lits, hv, hc = po.parse_dimacs(path)
        names.append(os.path.basename(rel))
        insts.append((off, lits))
        rc, mc = po.solve_csr(off, lits, kind=po.CORE)
        rs, ms = po.solve_csr(off, lits, kind=po.SIMP)
        for r, m in ((rc, mc), (rs, ms)):
            if r.verdict == po.SAT:
                assert po.check_model(off, lits, m), rel
        out["core_verdict"].append(rc.verdict); out["core_stats"].append(list(rc.stats))
        out["core_hash"].append(rc.trace_hash); out["core_traffic"].append(list(rc.traffic))
        out["simp_verdict"].append(rs.verdict); out["simp_stats"].append(list(rs.stats))
        out["simp_hash"].append(rs.trace_hash); out["simp_elim"].append(rs.eliminated_vars)
        out["simp_nclauses_pre"].append(rs.n_clauses_pre); out["n_vars"].append(rc.n_vars)
        b = os.path.basename(rel)
        out["name_verdict"].append(1 if b.startswith("uf") else 0 if b.startswith("uuf") else -1)
    ioff = [0]
    coffs, alll, base = [np.zeros(1, np.uint32)], [], 0
    for off, lits in insts:
        coffs.append((off[1:].astype(np.uint64) + base).astype(np.uint32))
        alll.append(lits)
        base += int(off[-1])
        ioff.append(ioff[-1] + len(off) - 1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), out_name),
                        names=np.array(names), inst_clause_off=np.array(ioff, np.uint32),
                        clause_off=np.concatenate(coffs), lits=np.concatenate(alll),
                        **{k: np.array(v, dtype=np.uint64 if "hash" in k or "stats" in k or "traffic" in k else np.int64)
                           for k, v in out.items()})
    print("wrote", len(names), "instances to", out_name)


def main():
    build(PICK, "corpus_sample.npz")
    # config 1 at reduced size: every 12th bundled uf20/uf50/uuf50 instance (tests/minisat.rs walks every 4th)
    wide = sorted(os.path.relpath(f, REF) for f in glob.glob(os.path.join(REF, "cnf", "u*.cnf.gz")))[::12]
    build(wide, "corpus_wide.npz")


if __name__ == "__main__" and "--hard" not in sys.argv:
    main()


def build_hard(pick, out_name):
    """BASELINE.json configs[3]: structured tests/cnf-hard instances, SimpSolver PREPROCESSING ONLY
    (--no-solve): the oracle's elimination result (clauses left, eliminated vars, propagations)."""
    names, insts, rows = [], [], []
    for rel in pick:
        off, lits, hv, hc = po.parse_dimacs(os.path.join(REF, rel))
        r, _ = po.solve_csr(off, lits, kind=po.SIMP, flags=po.F_PREPROCESS | po.F_NO_SOLVE)
        names.append(os.path.basename(rel))
        insts.append((off, lits))
        rows.append([r.n_vars, r.n_clauses_ingest, r.n_clauses_pre, r.eliminated_vars, r.pre_ok] + list(r.stats))
    ioff = [0]
    coffs, alll, base = [np.zeros(1, np.uint32)], [], 0
    for off, lits in insts:
        coffs.append((off[1:].astype(np.uint64) + base).astype(np.uint32))
        alll.append(lits)
        base += int(off[-1])
        ioff.append(ioff[-1] + len(off) - 1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), out_name), names=np.array(names),
                        inst_clause_off=np.array(ioff, np.uint32), clause_off=np.concatenate(coffs),
                        lits=np.concatenate(alll), rows=np.array(rows, dtype=np.uint64))
    print("wrote", len(names), "instances to", out_name)


if __name__ == "__main__" and "--hard" in sys.argv:
    build_hard(["cnf-hard/2bitadd_10.cnf.gz", "cnf-hard/3bitadd_32.cnf.gz", "cnf-hard/grieu-vmpc-s05-24s.cnf.gz",
                "cnf-hard/e0ddr2-10-by-5-1.cnf.gz", "cnf-hard/manol-pipe-f6b.cnf.gz", "cnf-hard/stric-bmc-ibm-12.cnf.gz"],
               "corpus_hard.npz")
