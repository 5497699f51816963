"""BASELINE.json configs[2] shape (uniform 3-SAT n=250, m=1065) on ONE GPU for a sample of the 100k batch:
device-resident throughput + the oracle on the host cores for a sub-sample (development / BASELINE.md aid)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _b200pkg
pkg = _b200pkg.load()
from oracle import pyoracle as po
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cpu_n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n, m = 250, 1065
eng = pkg.Engine(0)
eng.generate_ksat(0x5EED000100000000, n_inst, n, m, 3)
t0 = time.perf_counter()
eng.solve(kind=pkg.CORE, flags=pkg.F_PREPROCESS | pkg.F_COUNT_TRAFFIC)
wall = time.perf_counter() - t0
info = eng.launch_info()
res, models = eng.download(model_stride=n)
ioff, coff, lits = eng.download_batch()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import check_models
bad = check_models(ioff, coff, lits, res["verdict"], models)
ms = info["last_kernel_ms"]
props = float(res["stats"][:, 5].sum())
alg = float(pkg.traffic_bytes(res["traffic"].sum(axis=0)))
cores = len(os.sched_getaffinity(0))
ores, _, secs = po.solve_batch(ioff[:cpu_n + 1], coff, lits, kind=po.CORE, threads=cores)
same = all(list(res[i]["stats"]) == list(ores[i].stats) for i in range(cpu_n))
print(json.dumps({"workload": "uniform 3-SAT n=250 m=1065, first %d instances of the config-3 batch, CoreSolver" % n_inst,
                  "kernel_ms": ms, "instances_per_sec": n_inst / ms * 1e3, "propagations_per_sec": props / ms * 1e3,
                  "conflicts_mean": float(res["stats"][:, 4].mean()), "conflicts_max": int(res["stats"][:, 4].max()),
                  "sat": int((res["verdict"] == 1).sum()), "failed": int((res["status"] != 0).sum()),
                  "attempts_max": int(res["attempts"].max()), "model_violations": bad,
                  "algorithmic_GBps": alg / ms / 1e6, "bytes_per_propagation": alg / props, "launch": info,
                  "cpu_oracle": {"instances": cpu_n, "threads": cores, "seconds": secs, "instances_per_sec": cpu_n / secs,
                                 "stats_equal_gpu": bool(same)}}))
