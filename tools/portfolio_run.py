"""Portfolio mode driver (BASELINE.json configs[4]): diversified replicas of ONE tests/cnf-hard instance on every GPU
of the box, short learnt clauses exchanged through per-GPU rings that peers map with CUDA IPC (NVLink P2P), first
finisher cancels the rest.  Run directly (1 GPU) or under torchrun (one rank per GPU).

  python tools/portfolio_run.py --instance grieu-vmpc-s05-24s.cnf.gz --replicas 740 --arms share,noshare,lbd2,single --cpu
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
         tools/portfolio_run.py --instance grieu-vmpc-s05-24s.cnf.gz --arms share,noshare

The instance comes from the committed fixture tests/golden/corpus_hard.npz (the reference tree does not exist on the GPU
box); `--synthetic N` takes instance N of the config-3 seed sequence (n=250) instead.  One JSON line per run (rank 0):
for every arm the time to the first verdict (max over ranks of the wall time of the blocking call, all ranks start
behind a barrier), the winner, the verdict, model validity, exported / imported clause counts; `--cpu` adds the oracle
(the C++ port of the reference algorithm) on ONE host core, `--single` the reference configuration (replica 0) alone on one warp.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def load_instance(a):
    if a.synthetic >= 0:
        return None, "uniform 3-SAT n=250 m=1065, instance %d of the config-3 sequence" % a.synthetic
    if a.instance == "uuf250-0100.cnf.gz":   # a hard UNSAT tests/cnf-hard instance (more than 500k conflicts for the reference configuration)
        g = np.load(os.path.join(ROOT, "tests", "golden", "portfolio_uuf250_0100.npz"))
        return (g["clause_off"].astype(np.uint32), g["lits"].astype(np.uint32)), "tests/cnf-hard/uuf250-0100.cnf.gz (%d clauses)" % (len(g["clause_off"]) - 1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "corpus_hard.npz"))
    i = list(g["names"]).index(a.instance)
    cb, ce = int(g["inst_clause_off"][i]), int(g["inst_clause_off"][i + 1])
    off = g["clause_off"][cb:ce + 1].astype(np.int64)
    lits = g["lits"][off[0]:off[-1]].astype(np.uint32)
    return ((off - off[0]).astype(np.uint32), lits), "tests/cnf-hard/%s (%d clauses)" % (a.instance, ce - cb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instance", default="grieu-vmpc-s05-24s.cnf.gz")
    ap.add_argument("--synthetic", type=int, default=-1)
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (0 = one per resident warp)")
    ap.add_argument("--arms", default="share,noshare")
    ap.add_argument("--import-cap", type=int, default=64)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--budget-s", type=float, default=600.0, help="give up on an arm after this many seconds (async interrupt)")
    a = ap.parse_args()
    import threading
    import torch
    import _b200pkg
    pkg = _b200pkg.load()
    from common import batch_of, check_models
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inst, desc = load_instance(a)
    eng = pkg.Engine(local)
    if inst is None:
        eng.generate_ksat(0x5EED000100000000 + a.synthetic, 1, 250, 1065, 3)
        ioff, coff, lits = eng.download_batch()
    else:
        ioff, coff, lits = batch_of([inst])
        eng.upload(ioff, coff, lits)
    nv = int(lits.max() >> 1) + 1
    eng.ring_create(1 << 22, world, rank)
    eng.set_import_cap(a.import_cap)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.ring_ipc_handle())
        for r in range(world):
            if r != rank:
                eng.ring_open_peer(r, handles[r])
        dist.barrier()
    # replicas per GPU: one per resident warp of the search kernel (probe with a 1-replica launch)
    n_rep = a.replicas
    if n_rep <= 0:
        eng.set_budget(conflict_budget=0)
        eng.portfolio_solve(1, share=pkg.PF_IGNORE_DONE)
        n_rep = int(eng.launch_info()["max_warps"])
        eng.set_budget()
    out = {"workload": "portfolio: " + desc, "gpus": world, "replicas_per_gpu": n_rep, "n_vars": nv,
           "import_cap": a.import_cap, "arms": {}}

    def run_arm(name, replicas, share, only_rank0=False):
        """all ranks clear their ring, meet at a barrier, then call the blocking portfolio solve"""
        eng.ring_reset()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        mine = None
        if not only_rank0 or rank == 0:
            killer = threading.Timer(a.budget_s, lambda: eng.interrupt(True))
            killer.start()
            t0 = time.perf_counter()
            res, models = eng.portfolio_solve(replicas, replica_base=rank * replicas, share=share | pkg.PF_KEEP_RING, model_stride=nv)
            secs = time.perf_counter() - t0
            killer.cancel()
            eng.interrupt(False)
            done = np.nonzero((res["verdict"] != pkg.INTERRUPTED) & (res["status"] == 0))[0]
            model_ok = True
            if len(done) and res[done[0]]["verdict"] == pkg.SAT:
                model_ok = check_models(ioff, coff, lits, np.array([1]), models[done[:1]]) == 0
            info = eng.launch_info()
            mine = {"rank": rank, "seconds": secs, "kernel_ms": info["last_kernel_ms"],
                    "finished_replicas": [int(rank * replicas + k) for k in done[:4]],
                    "verdicts": sorted(set(int(v) for v in res["verdict"][done])), "model_ok": bool(model_ok),
                    "winner_conflicts": int(res[done[0]]["stats"][4]) if len(done) else None,
                    "conflicts_all_replicas": int(res["stats"][:, 4].sum()),
                    "exported": int(res["exported"].sum()), "imported": int(res["imported"].sum()),
                    "dropped_out_of_memory": int((res["status"] < 0).sum())}
        allr = [mine]
        if dist is not None:
            allr = [None] * world
            dist.all_gather_object(allr, mine)
        allr = [r for r in allr if r is not None]
        if rank == 0:
            verd = sorted(set(v for r in allr for v in r["verdicts"]))
            out["arms"][name] = {"time_to_verdict_s": max(r["seconds"] for r in allr), "verdicts": verd,
                                 "models_ok": all(r["model_ok"] for r in allr), "replicas": replicas * len(allr),
                                 "exported": sum(r["exported"] for r in allr), "imported": sum(r["imported"] for r in allr),
                                 "ranks": allr}

    for arm in a.arms.split(","):
        if arm == "share":
            run_arm("share_units_binaries", n_rep, pkg.PF_SHARE)
        elif arm == "lbd2":
            run_arm("share_units_binaries_lbd2", n_rep, pkg.PF_SHARE | pkg.PF_SHARE_LBD2)
        elif arm == "noshare":
            run_arm("no_sharing", n_rep, 0)
        elif arm == "single":
            # the reference configuration alone on one warp: the plain trace-mode solve (with the larger-slab retry ladder)
            if rank == 0:
                killer = threading.Timer(a.budget_s, lambda: eng.interrupt(True))
                killer.start()
                t0 = time.perf_counter()
                eng.solve(kind=pkg.CORE, flags=pkg.F_PREPROCESS)
                secs = time.perf_counter() - t0
                killer.cancel(); eng.interrupt(False)
                r1, _ = eng.download(want_models=False)
                out["arms"]["single_replica_reference_configuration"] = {
                    "time_to_verdict_s": secs, "verdicts": [int(r1[0]["verdict"])], "conflicts": int(r1[0]["stats"][4]),
                    "status": int(r1[0]["status"]), "note": "verdict 2 = stopped by --budget-s"}
            if dist is not None:
                dist.barrier()
    if a.cpu and rank == 0:
        from oracle import pyoracle as po
        try:
            po.use_native()
        except Exception:
            pass
        off = coff[ioff[0]:ioff[1] + 1]
        t0 = time.perf_counter()
        r, m = po.solve_csr((off - off[0]).astype(np.uint32), lits, kind=po.CORE)
        out["cpu_one_core"] = {"seconds": time.perf_counter() - t0, "verdict": int(r.verdict), "conflicts": int(r.stats[4]),
                               "kind": "port (oracle, CoreSolver, default settings)", "build": po.build_info()}
    if rank == 0:
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
