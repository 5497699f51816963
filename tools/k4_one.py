"""K4 one-preprocessor on one corpus_hard instance (profiling aid): python tools/k4_one.py [name]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from common import batch_of
name = sys.argv[1] if len(sys.argv) > 1 else "stric-bmc-ibm-12.cnf.gz"
g = np.load(os.path.join(ROOT, "tests", "golden", "corpus_hard.npz"))
i = list(g["names"]).index(name)
cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
off = g["clause_off"][cb:ce + 1].astype(np.int64); lits = g["lits"][off[0]:off[-1]].astype(np.uint32); off = (off - off[0]).astype(np.uint32)
e = pkg.Engine(0)
ioff, coff, li = batch_of([(off, lits)])
e.upload(ioff, coff, li)
for _ in range(2):
    r = e.preprocess_instance(0, len(off) - 1, len(lits))
print(name, len(off) - 1, len(lits), {k: v for k, v in r.items() if not isinstance(v, np.ndarray)})
