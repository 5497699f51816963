"""Replay one instance of a batch dumped by tools/gpu_fuzz.py (ioff/coff/lits/kind/settings .npz) on the GPU and
compare it with the oracle:  python tools/gpu_replay.py dump.npz instance_index   (development aid)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _b200pkg  # noqa: E402

pkg = _b200pkg.load()
from oracle import pyoracle as po  # noqa: E402

f = np.load(sys.argv[1])
i = int(sys.argv[2])
ioff, coff, lits, kind = f["ioff"], f["coff"], f["lits"], int(f["kind"])
cb, ce = int(ioff[i]), int(ioff[i + 1])
off = coff[cb:ce + 1].astype(np.int64)
sub_l = lits[off[0]:off[-1]].astype(np.uint32)
sub_c = (off - off[0]).astype(np.uint32)
sub_i = np.array([0, ce - cb], np.uint32)
st = pkg.default_settings()
C.memmove(C.addressof(st), f["settings"].tobytes(), C.sizeof(st))
ost = po.default_settings()
C.memmove(C.addressof(ost), C.addressof(st), C.sizeof(st))
flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT | pkg.F_TRACE_HASH | pkg.F_COUNT_TRAFFIC
eng = pkg.Engine(0)
t = time.time()
res, _ = eng.solve_batch(sub_i, sub_c, sub_l, settings=st, kind=kind, flags=flags, model_stride=int(sub_l.max() >> 1) + 1)
t_gpu = time.time() - t
t = time.time()
ores, _, _ = po.solve_batch(sub_i, sub_c, sub_l, kind=kind, settings=ost, threads=1)
t_cpu = time.time() - t
same = (res[0]["status"] == 0 and res[0]["verdict"] == ores[0].verdict and list(res[0]["stats"]) == list(ores[0].stats)
        and int(res[0]["trace_hash"]) == ores[0].trace_hash and list(res[0]["traffic"]) == list(ores[0].traffic))
print("gpu %.1f s (attempts %d, status %d) stats %s | oracle %.1f s stats %s | %s" % (
    t_gpu, res[0]["attempts"], res[0]["status"], list(res[0]["stats"]), t_cpu, list(ores[0].stats), "IDENTICAL" if same else "MISMATCH"))
sys.exit(0 if same else 1)
