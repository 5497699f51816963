"""Portfolio mode driver (BASELINE.json configs[4]): diversified replicas of ONE instance on every GPU of
the box, short learnt clauses exchanged through per-GPU rings that peers map with CUDA IPC (NVLink P2P),
first finisher cancels the rest.  Run directly (1 GPU) or under torchrun (one rank per GPU).

  python tools/portfolio_run.py --nvars 250 --index 3 --replicas 2048 [--no-share]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nvars", type=int, default=250)
    ap.add_argument("--ratio", type=float, default=4.26)
    ap.add_argument("--index", type=int, default=0, help="instance index in the config-3 seed sequence")
    ap.add_argument("--replicas", type=int, default=2048, help="replicas per GPU")
    ap.add_argument("--no-share", action="store_true")
    ap.add_argument("--lbd2", action="store_true", help="also exchange LBD<=2 clauses of <= 8 literals")
    ap.add_argument("--single", action="store_true", help="also time replica 0 alone (the reference configuration)")
    a = ap.parse_args()
    import torch
    import _b200pkg
    pkg = _b200pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, m = a.nvars, int(round(a.ratio * a.nvars))
    seed = 0x5EED000100000000 + a.index
    eng = pkg.Engine(local_rank)
    eng.generate_ksat(seed, 1, n, m, 3)
    eng.ring_create(1 << 20, world, rank)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.ring_ipc_handle())
        for r in range(world):
            if r != rank:
                eng.ring_open_peer(r, handles[r])
        dist.barrier()
    share = 0 if a.no_share else (pkg.PF_SHARE | (pkg.PF_SHARE_LBD2 if a.lbd2 else 0))
    out = {}
    if a.single and rank == 0:
        t0 = time.perf_counter()
        res, _ = eng.portfolio_solve(1, share=pkg.PF_IGNORE_DONE, model_stride=n)
        out["single_replica"] = {"seconds": time.perf_counter() - t0, "verdict": int(res[0]["verdict"]),
                                 "conflicts": int(res[0]["stats"][4])}
        eng.ring_reset()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    res, models = eng.portfolio_solve(a.replicas, replica_base=rank * a.replicas, share=share, model_stride=n)
    secs = time.perf_counter() - t0
    done = np.nonzero(res["verdict"] != pkg.INTERRUPTED)[0]
    ioff, coff, lits = eng.download_batch()
    ok_models = True
    if len(done) and res[done[0]]["verdict"] == pkg.SAT:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from common import check_models
        ok_models = check_models(ioff, coff, lits, np.array([1]), models[done[:1]]) == 0
    mine = {"rank": rank, "seconds": secs, "finished_replicas": [int(rank * a.replicas + k) for k in done[:4]],
            "verdicts": sorted(set(int(v) for v in res["verdict"][done])), "model_ok": bool(ok_models),
            "winner_conflicts": int(res[done[0]]["stats"][4]) if len(done) else None,
            "exported": int(res["exported"].sum()), "imported": int(res["imported"].sum()),
            "kernel_ms": eng.launch_info()["last_kernel_ms"]}
    allr = [mine]
    if dist is not None:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank == 0:
        out.update({"workload": "portfolio: uniform 3-SAT n=%d m=%d, instance %d" % (n, m, a.index), "gpus": world,
                    "replicas_per_gpu": a.replicas, "share": not a.no_share,
                    "time_to_verdict_s": max(r["seconds"] for r in allr), "ranks": allr})
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
