"""K4 one-preprocessor soundness probe (development aid): random instances, reduced formula vs original."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from oracle import pyoracle as po
from common import batch_of
from test_gpu_parity import _extend_model_host
e = pkg.Engine(0)
bad = 0
def probe(name, off, lits, **kw):
    global bad
    ioff, coff, li = batch_of([(off, lits)])
    e.upload(ioff, coff, li)
    r = e.preprocess_instance(0, len(off) - 1, len(lits), **kw)
    nv = int(lits.max() >> 1) + 1
    o, om = po.solve_csr(off, lits, kind=po.CORE)
    if r["unsat"]:
        ok = o.verdict == 0
        print(name, kw, "derived UNSAT", "ok" if ok else "WRONG"); bad += not ok; return
    red, rm = po.solve_csr(r["clause_off"], r["lits"] if len(r["lits"]) else np.zeros(1, np.uint32), kind=po.CORE) if r["kept"] else (None, None)
    rv = 1 if red is None else red.verdict
    msg = "kept %d elim %d sub %d str %d outer %d" % (r["kept"], r["eliminated"], r["subsumed"], r["strengthened"], r["outer"])
    if rv != o.verdict:
        print(name, kw, "VERDICT MISMATCH", rv, o.verdict, msg); bad += 1; return
    if rv == 1:
        m0 = np.full(nv, 2, np.uint8)
        if rm is not None: m0[:len(rm)] = rm
        # is the reduced model itself a model of the ORIGINAL clauses that survive?  extension:
        dev = e.k4_extend_model(m0); host = _extend_model_host(r["estack"], m0)
        full = dev.copy(); full[full == 2] = 0
        ok = po.check_model(off, lits, full)
        if not ok or not (dev == host).all():
            # which original clauses are violated?
            viol = []
            for c in range(len(off) - 1):
                cl = lits[off[c]:off[c + 1]]
                if not any((full[l >> 1] == 1) == ((l & 1) == 0) for l in cl): viol.append([int(x) for x in cl])
            print(name, kw, "MODEL BAD dev==host", bool((dev == host).all()), msg, "violated", len(viol), viol[:3]); bad += 1; return
    print(name, kw, "ok", msg)
rng = np.random.RandomState(1)
for t in range(6):
    n, m = 40, 150
    lits = po.gen_ksat(1234 + t, n, m); off = np.arange(0, m + 1, dtype=np.uint32) * 3
    for kw in (dict(cl_lim=0, max_outer=1), dict(cl_lim=0), dict()):
        probe("rnd%d" % t, off, lits, **kw)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "corpus_hard.npz"))
i = list(g["names"]).index("grieu-vmpc-s05-24s.cnf.gz")
cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
off = g["clause_off"][cb:ce + 1].astype(np.int64); lits = g["lits"][off[0]:off[-1]].astype(np.uint32); off = (off - off[0]).astype(np.uint32)
for kw in (dict(cl_lim=0, max_outer=1), dict(cl_lim=0), dict()):
    probe("grieu", off, lits, **kw)
print("bad", bad)
