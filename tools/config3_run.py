#!/usr/bin/env python
"""BASELINE.json configs[2]: 100,000 uniform random 3-SAT instances, n=250, m=1065 (ratio 4.26), SplitMix64 seeds
0x5EED0001_00000000 + global index, sharded over the GPUs of one box with no data-path collective.

  python tools/config3_run.py [--instances 100000] [--cpu-sample 64]          (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/config3_run.py ...                                               (N GPUs, one rank each)

Every rank generates its contiguous shard on the device, solves it in trace mode (CoreSolver, the reference's
configuration; conflict slices on so that long instances yield to waiting ones), verifies every SAT model on the
device, and reports.  Rank 0 times the oracle (the C++ port of the reference algorithm) on a sample of the same
batch with all host threads -- the reported CPU baseline -- and checks that sample's stats against the device.
One JSON line on stdout (rank 0)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

SEED3 = 0x5EED000100000000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=100000, help="instances of the whole job (all GPUs)")
    ap.add_argument("--nvars", type=int, default=250)
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--slice", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rescue", action="store_true", help="free-running mode: idle warps help instances that are still running")
    ap.add_argument("--max-seconds", type=float, default=0.0, help="raise the async interrupt after this long (0 = never): "
                    "instances still running then are reported as interrupted")
    a = ap.parse_args()
    import torch
    import _b200pkg
    pkg = _b200pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.nvars; m = int(round(4.26 * n))
    lo, hi = pkg.shard_range(a.instances, rank, world)
    B = hi - lo
    eng = pkg.Engine(local)
    eng.generate_ksat(SEED3 + lo, B, n, m, 3)
    eng.set_slice(a.slice)
    eng.set_rescue(a.rescue)
    flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT | (0 if a.rescue else pkg.F_COUNT_TRAFFIC)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    import threading
    killer = threading.Timer(a.max_seconds, lambda: eng.interrupt(True)) if a.max_seconds > 0 else None
    if killer:
        killer.start()
    t0 = time.perf_counter()
    eng.solve(kind=pkg.CORE, flags=flags)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if killer:
        killer.cancel()
        eng.interrupt(False)
    info = eng.launch_info()
    res, _ = eng.download(want_models=False)
    verified, n_bad = eng.verify_models()
    st = res["stats"].astype(np.float64)
    alg = float(sum(pkg.traffic_bytes(t) for t in res["traffic"].sum(axis=0, keepdims=True)))
    mine = dict(ms=info["last_kernel_ms"], search_ms=info["search_ms"], tail_ms=info["tail_ms"], span_ms=info["span_ms"],
                props=float(st[:, 5].sum()), confl=float(st[:, 4].sum()), sat=int((res["verdict"] == pkg.SAT).sum()),
                unsat=int((res["verdict"] == pkg.UNSAT).sum()), failed=int((res["status"] != 0).sum()),
                interrupted=int((res["verdict"] == pkg.INTERRUPTED).sum()),
                bad_models=int(n_bad), max_confl=float(st[:, 4].max()), alg=alg, suspends=info["suspends"],
                attempts=int((res["attempts"] & 0xFF).max()), wall=wall, warps=info["warps_total"],
                rescued=int(((res["attempts"] & 0x100) != 0).sum()))
    keys = sorted(mine)
    t = torch.tensor([mine[k] for k in keys], dtype=torch.float64, device="cuda")
    allr = [t.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    if rank == 0:
        per = [dict(zip(keys, [float(x) for x in r.tolist()])) for r in allr]
        ms = max(p["ms"] for p in per)
        out = {
            "workload": "uniform random 3-SAT, n=%d m=%d, %d instances over %d GPU(s), CoreSolver, %s" % (
                n, m, a.instances, world, "free-running mode with straggler rescue" if a.rescue else "trace mode"),
            "mode": "rescue" if a.rescue else "trace", "decided_by_helper_replicas": int(sum(p["rescued"] for p in per)),
            "n_gpus": world, "instances": a.instances, "seed_base": hex(SEED3), "slice_conflicts": a.slice,
            "ms": ms, "instances_per_sec": a.instances / ms * 1e3,
            "propagations_per_sec": sum(p["props"] for p in per) / ms * 1e3,
            "conflicts_per_sec": sum(p["confl"] for p in per) / ms * 1e3,
            "conflicts_mean": sum(p["confl"] for p in per) / a.instances, "conflicts_max": max(p["max_confl"] for p in per),
            "sat": int(sum(p["sat"] for p in per)), "unsat": int(sum(p["unsat"] for p in per)),
            "failed": int(sum(p["failed"] for p in per)), "interrupted_by_time_limit": int(sum(p["interrupted"] for p in per)),
            "model_violations": int(sum(p["bad_models"] for p in per)),
            "all_models_verified_on_device": all(p["bad_models"] == 0 for p in per),
            "tail_fraction": max(p["tail_ms"] / max(p["span_ms"], 1e-9) for p in per),
            "tail_ms_per_gpu": [p["tail_ms"] for p in per], "ms_per_gpu": [p["ms"] for p in per],
            "suspends": int(sum(p["suspends"] for p in per)), "retry_attempts_max": int(max(p["attempts"] for p in per)),
            "algorithmic_GBps": sum(p["alg"] for p in per) / ms / 1e6,
            "bytes_per_propagation": sum(p["alg"] for p in per) / max(1.0, sum(p["props"] for p in per)),
            "warps_per_gpu": int(per[0]["warps"]), "wall_s_rank0": wall,
        }
        if not a.no_cpu:
            from oracle import pyoracle as po
            try:
                po.use_native()
            except Exception:
                pass
            cores = len(os.sched_getaffinity(0))
            k = min(a.cpu_sample, B)
            lits = np.concatenate([po.gen_ksat(SEED3 + i, n, m) for i in range(k)])
            coff = np.arange(0, k * m + 1, dtype=np.uint32) * 3
            ioff = np.arange(0, k + 1, dtype=np.uint32) * m
            ores, _, secs = po.solve_batch(ioff, coff, lits, kind=po.CORE, threads=cores)
            same = all(res[i]["verdict"] == ores[i].verdict and ((res[i]["attempts"] & 0x100) != 0 or list(res[i]["stats"]) == list(ores[i].stats))
                       for i in range(k))
            out["cpu_baseline"] = {"kind": "port", "cores": cores, "sample": "first %d instances of the batch" % k, "seconds": secs,
                                   "instances_per_sec": k / secs, "stats_equal_device": bool(same),
                                   "check": "verdicts of the sample equal; counters equal for every instance its original replica decided",
                                   "build": po.build_info()}
            out["speedup_vs_host_cores"] = out["instances_per_sec"] / (k / secs)
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
