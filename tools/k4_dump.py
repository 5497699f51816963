"""K4 development aid: run the one-preprocessor on one corpus_hard instance and save its outputs for offline analysis."""
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
out = {}
for tag, kw in (("full", {}), ("outer1", dict(max_outer=1)), ("nolim_outer1", dict(max_outer=1, cl_lim=-1))):
    r = e.preprocess_instance(0, len(off) - 1, len(lits), **kw)
    print(tag, {k: v for k, v in r.items() if not isinstance(v, np.ndarray)})
    out[tag + "_off"] = r["clause_off"]; out[tag + "_lits"] = r["lits"]; out[tag + "_estack"] = r["estack"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "k4_dump.npz"), **out)
