#!/usr/bin/env python
"""bench.py -- batched CDCL throughput on B200 (BASELINE.json metric: instances/sec).

A "step" is one pass of the hot path (ingest + CDCL search of every instance) over one
batch of synthetic uniform random 3-SAT.  N=1 workload = BASELINE.json configs[1]:
10,000 instances, n=100, m=426 (ratio 4.26), SplitMix64 seeds 0x5EED0000_00000000+i
(SURVEY.md 8d).  Under torchrun every rank solves its own batch of the same size with
disjoint seeds (weak scaling, no data-path collective).

  value      instances/s with the CSR batch already resident in HBM (CUDA events on the launch stream)
  e2e        instances/s through b200sat_solve_batch with pinned HOST buffers (H2D + solve + D2H)
  roofline   algorithmic bytes (SURVEY.md 8d counters) of the cdcl_warp kernel / its event time
  cpu_baseline  the oracle (C++ port of the reference algorithm) on a bounded sample, host cores

`--impl reference` times the oracle port on the host cores instead (the reference is
Rust; no toolchain exists in this image, so oracle/_ref cannot be built -- DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED_BASE = 0x5EED000000000000
METRIC = "instances/sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--nvars", type=int, default=100)
    ap.add_argument("--ratio", type=float, default=4.26)
    ap.add_argument("--solver", default="core", choices=["core", "simp"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=3, help="batches in flight (engines/streams); 1 = strictly serial steps")
    ap.add_argument("--config3", type=int, default=-1,
                    help="instances per GPU of the bounded config-3 (n=250) sub-record; -1 = 2368 when --gpus > 1 else off")
    ap.add_argument("--no-bcp-replay", action="store_true")
    return ap.parse_args()


def workload_name(a, m):
    return "uniform random 3-SAT batch: %d instances/GPU, n=%d, m=%d (ratio %.2f), %s solver" % (
        a.instances, a.nvars, m, a.ratio, "CoreSolver (--core)" if a.solver == "core" else "SimpSolver (default)")


def shared_config(a, m):
    """`config` of the JSON line: identical keys and values in both arms (b200 and --impl reference)."""
    return {"workload": workload_name(a, m), "instances_per_gpu": a.instances, "n_vars": a.nvars, "n_clauses": m,
            "seed_base": hex(SEED_BASE), "solver": a.solver,
            "l2": "256 MiB memset between steps (L2 flush); the batch (51 MB CSR + 8.5 GB of per-instance slabs) exceeds L2"}


def load_oracle():
    """The oracle for the CPU legs: rebuilt on THIS host with -O3 -march=native (-flto where the image has it),
    SURVEY.md 8d; falls back to the portable build that travels with the repo."""
    from oracle import pyoracle as po
    try:
        po.use_native()
    except Exception:
        po.build()
    return po


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.smax = None
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.smax = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_sample_rate(po, n, m, kind, n_sample, threads, seed0):
    lits = np.concatenate([po.gen_ksat(seed0 + i, n, m) for i in range(n_sample)])
    coff = np.arange(0, n_sample * m + 1, dtype=np.uint32) * 3
    ioff = np.arange(0, n_sample + 1, dtype=np.uint32) * m
    res, _, secs = po.solve_batch(ioff, coff, lits, kind=kind, threads=threads)
    props = sum(int(r.stats[5]) for r in res)
    oracle_sample_rate.last = res
    return n_sample / secs, props / secs, secs


def run_reference(a):
    """--impl reference: the oracle port on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    po = load_oracle()
    m = int(round(a.ratio * a.nvars))
    kind = po.CORE if a.solver == "core" else po.SIMP
    cores = cpu_cores()
    # calibrate the per-step sample to ~10 s of wall time per step, at most 60 s in total
    r0, _, _ = oracle_sample_rate(po, a.nvars, m, kind, max(4 * cores, 16), cores, SEED_BASE)
    budget = min(10.0, 60.0 / max(1, a.steps + a.warmup))
    n_sample = int(max(cores, min(a.instances, r0 * budget)))
    for w in range(a.warmup):
        oracle_sample_rate(po, a.nvars, m, kind, max(cores, n_sample // 8), cores, SEED_BASE)
    t_tot, n_tot, p_tot = 0.0, 0, 0.0
    for s in range(a.steps):
        r, p, secs = oracle_sample_rate(po, a.nvars, m, kind, n_sample, cores, SEED_BASE)
        t_tot += secs
        n_tot += n_sample
        p_tot += p * secs
    value = n_tot / t_tot
    sample = "first %d instances of the batch per step, %d threads" % (n_sample, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "instances/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_tot / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": shared_config(a, m),
        "details": {"sample": sample, "oracle_build": po.build_info(),
                    "note": "reference arm = C++ oracle port of minisat-rust (Rust toolchain absent); CPU only"},
        "propagations_per_sec": p_tot / t_tot,
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": cores, "kind": "port", "sample": sample,
                         "build": po.build_info()},
        "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import _b200pkg
    pkg = _b200pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, m, B = a.nvars, int(round(a.ratio * a.nvars)), a.instances
    kind = pkg.CORE if a.solver == "core" else pkg.SIMP
    flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT
    main_stream = torch.cuda.current_stream()
    sptr = main_stream.cuda_stream
    # Pipeline depth: consecutive steps run on alternating engines/streams so that the tail of one
    # batch (its few hardest instances) overlaps the start of the next one.  Every step still is one
    # complete pass over one complete batch.
    NS = max(1, a.pipeline)
    engs = [pkg.Engine(local_rank) for _ in range(NS)]
    streams = [torch.cuda.Stream() for _ in range(NS)]
    seed0 = SEED_BASE + rank * B
    for e in engs:
        e.generate_ksat(seed0, B, n, m, 3, stream=sptr)
    torch.cuda.synchronize()
    eng = engs[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic bytes of this batch (deterministic trace): one counting pass, untimed
    eng.solve(kind=kind, flags=flags | pkg.F_COUNT_TRAFFIC, stream=sptr)
    res_c, _ = eng.download(want_models=False)
    assert int((res_c["status"] != 0).sum()) == 0, "instances failed"
    alg_bytes = float(sum(pkg.traffic_bytes(t) for t in res_c["traffic"].sum(axis=0, keepdims=True)))
    props = float(res_c["stats"][:, 5].sum())
    confl = float(res_c["stats"][:, 4].sum())
    n_sat = int((res_c["verdict"] == pkg.SAT).sum())

    def run_steps(k_steps):
        """k_steps passes, pipelined over the engines; returns (sum of per-launch event ms, #kernel launches)."""
        kms, nl = 0.0, 0
        for k in range(k_steps):
            e, st = engs[k % NS], streams[k % NS]
            if k >= NS:
                e.finish()
                li = e.launch_info(); kms += li["search_ms"]; nl += li["launches"] + 3
            flush.zero_()                       # L2 flush between steps (256 MiB > 126 MB L2), on the main stream
            st.wait_stream(main_stream)
            e.solve_async(kind=kind, flags=flags, stream=st.cuda_stream)
        for k in range(min(NS, k_steps)):
            e = engs[k]
            e.finish()
            li = e.launch_info(); kms += li["search_ms"]; nl += li["launches"] + 3
        for st in streams:
            main_stream.wait_stream(st)
        return kms, nl

    # ---- device-resident timing
    run_steps(a.warmup)
    # one un-overlapped launch for the per-launch duration the roofline is quoted on
    torch.cuda.synchronize()
    eng.solve(kind=kind, flags=flags, stream=sptr)
    solo_info = eng.launch_info()
    solo_ms = solo_info["search_ms"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main_stream)
    kernel_ms, launches = run_steps(a.steps)
    ev1.record(main_stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    sampler.stop_flag.set()
    sampler.join()
    info = eng.launch_info()
    for e in engs[:min(NS, a.steps)]:
        res, _ = e.download(want_models=False)
        assert (res["stats"] == res_c["stats"]).all(), "timed pass diverged from the counting pass"

    # ---- end to end through the C-ABI batch call with pinned host buffers (H2D + solve + D2H per step)
    ioff, coff, lits = eng.download_batch()
    h_ioff = torch.from_numpy(ioff.astype(np.int32)).pin_memory()
    h_coff = torch.from_numpy(coff.astype(np.int32)).pin_memory()
    h_lits = torch.from_numpy(lits.astype(np.int32)).pin_memory()
    np_ioff, np_coff, np_lits = (h_ioff.numpy().view(np.uint32), h_coff.numpy().view(np.uint32),
                                 h_lits.numpy().view(np.uint32))
    outs = []
    for _ in range(NS):
        hr = torch.zeros(B * pkg.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        hm = torch.zeros((B, n), dtype=torch.uint8).pin_memory()
        outs.append((hr, hm, hr.numpy().view(pkg.RESULT_DTYPE), hm.numpy()))
    h2d = 4 * (len(np_ioff) + len(np_coff) + len(np_lits))
    d2h = B * pkg.RESULT_DTYPE.itemsize + B * n

    def run_e2e(k_steps):
        for k in range(k_steps):
            e, st, o = engs[k % NS], streams[k % NS], outs[k % NS]
            if k >= NS:
                e.solve_batch_finish()
            e.solve_batch_async(np_ioff, np_coff, np_lits, o[2], o[3], kind=kind, flags=flags, stream=st.cuda_stream)
        for k in range(min(NS, k_steps)):
            engs[k].solve_batch_finish()

    run_e2e(max(NS, a.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    run_e2e(a.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for o in outs[:min(NS, a.steps)]:
        assert (o[2]["stats"] == res_c["stats"]).all()
        assert (o[2]["verdict"] == res_c["verdict"]).all()

    # ---- K2: BCP-only replay of the recorded search prefixes (SURVEY.md 8d "BCP-only figure"), rank 0, untimed region
    bcp = None
    if rank == 0 and a.solver == "core" and not a.no_bcp_replay:
        bcp = eng.bcp_replay(reps=3, stream=sptr)

    # ---- bounded config-3 sub-record (BASELINE.json configs[2] shape: n=250, m=1065): every rank solves a slice of the
    # 100k batch under a conflict budget, so it finishes in seconds; the full-size run is tools/config3_run.py
    c3 = None
    c3_n = a.config3 if a.config3 >= 0 else (2368 if world > 1 else 0)
    if c3_n > 0:
        e3 = engs[-1]
        e3.generate_ksat(0x5EED000100000000 + rank * c3_n, c3_n, 250, 1065, 3, stream=sptr)
        e3.set_budget(conflict_budget=20000)
        for _ in range(2):
            e3.solve(kind=pkg.CORE, flags=flags, stream=sptr)
        e3.set_budget()
        li3 = e3.launch_info()
        r3, _ = e3.download(want_models=False)
        c3 = [float(li3["last_kernel_ms"]), float(r3["stats"][:, 5].sum()), float(r3["stats"][:, 4].sum()),
              float((r3["verdict"] != pkg.INTERRUPTED).sum()), float((r3["status"] != 0).sum())]
        t3 = torch.tensor(c3, dtype=torch.float64, device="cuda")
        if dist is not None:
            g3 = [t3.clone() for _ in range(world)]
            dist.all_gather(g3, t3)
            c3 = [[float(x) for x in g.tolist()] for g in g3]
        else:
            c3 = [c3]

    # ---- max over ranks
    tt = torch.tensor([dev_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([props, alg_bytes, confl], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, kernel_ms_max = [float(x) for x in tt.tolist()]
    props_all, bytes_all, confl_all = [float(x) for x in agg.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if a.solver == "core" and n == 100 and B == 10000:
                # NOT measured in this run (ncu cannot run inside the bench): the committed `ncu --set full` capture of
                # this exact workload and kernel; `traffic_static` says where it came from
                traffic = float(tj["dram_bytes_per_launch"])
                traffic_src = {"static": True, "capture": tj.get("capture"), "kernel": tj.get("kernel"), "commit": tj.get("commit"),
                               "issue": tj.get("issue")}
        except Exception:
            pass
        value = world * B * a.steps / (dev_ms * 1e-3)
        # roofline of the dominant kernel (cdcl_warp) on rank 0.  With NS launches in flight the GPU is shared,
        # so `achieved` is the aggregate: algorithmic bytes of all K launches / device time of the timed region;
        # the per-launch event durations (overlapped and solo) are reported beside it.
        k_ms = kernel_ms / a.steps
        achieved = alg_bytes * a.steps / (dev_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "instances/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": shared_config(a, m),
            "details": {"pipeline": "%d batches in flight on %d engines/streams" % (NS, NS), "satisfiable": n_sat,
                        "launch": info, "solo_launch": solo_info},
            "propagations_per_sec": props_all * a.steps / (dev_ms * 1e-3),
            "conflicts_per_sec": confl_all * a.steps / (dev_ms * 1e-3),
            "e2e": {"value": world * B * a.steps / (e2e_ms * 1e-3), "unit": "instances/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_static": traffic_src, "kernel": "search_smem<128,0,0>", "launches_in_flight": NS,
                         "kernel_ms_per_launch_overlapped": k_ms, "kernel_ms_per_launch_solo": solo_ms,
                         "achieved_solo_launch": alg_bytes / (solo_ms * 1e-3) / 1e9,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_propagation": alg_bytes / max(props, 1.0),
                         "peak_source": peak_src},
            "clocks": sampler.summary(),
        }
        if bcp is not None:
            # BCP only: the replay kernel re-executes the recorded prefixes (up to the first reduceDB) with unit
            # propagation as the only real work; bytes = SURVEY.md 8d terms (1)-(5) of exactly those prefixes
            bsec = bcp["ms_min"] * 1e-3
            line["roofline"]["bcp_only"] = {
                "achieved": bcp["bcp_bytes"] / bsec / 1e9 if bsec > 0 else None, "unit": "GB/s",
                "frac": bcp["bcp_bytes"] / bsec / 1e9 / peak if bsec > 0 else None,
                "kernel": "replay_smem<128,0>", "kernel_ms": bcp["ms_min"], "kernel_ms_avg": bcp["ms_avg"],
                "algorithmic_bytes": bcp["bcp_bytes"], "propagations": bcp["propagations"],
                "propagations_per_sec": bcp["propagations"] / bsec if bsec > 0 else None,
                "instances": bcp["instances"], "replay_mismatches": bcp["mismatches"], "dram_bytes": None,
                "note": "mismatches must be 0"}
            try:   # committed ncu capture of replay_smem on this exact workload (static, like `traffic`)
                kj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("bcp_only")
                if kj and a.solver == "core" and n == 100 and B == 10000:
                    line["roofline"]["bcp_only"]["dram_bytes"] = float(kj["dram_bytes_per_launch"])
                    line["roofline"]["bcp_only"]["dram_bytes_static"] = {"static": True, "capture": kj.get("capture")}
            except Exception:
                pass
            assert bcp["mismatches"] == 0, "bcp_replay left the recorded trace"
        if c3 is not None:
            ms3 = max(x[0] for x in c3)
            line["config3"] = {
                "workload": "uniform random 3-SAT n=250 m=1065 (BASELINE.json configs[2] shape), %d instances/GPU, "
                            "conflict budget 20000 per instance (bounded slice; full size: tools/config3_run.py, profiles/r02_config3_*.json)" % c3_n,
                "ms": ms3, "propagations_per_sec": sum(x[1] for x in c3) / ms3 * 1e3,
                "conflicts_per_sec": sum(x[2] for x in c3) / ms3 * 1e3, "decided_within_budget": int(sum(x[3] for x in c3)),
                "instances": c3_n * world, "failed": int(sum(x[4] for x in c3))}
        if not a.no_cpu_baseline:
            po = load_oracle()
            cores = cpu_cores()
            okind = po.CORE if a.solver == "core" else po.SIMP
            r0, _, _ = oracle_sample_rate(po, n, m, okind, max(2 * cores, 16), cores, SEED_BASE)
            ns = a.cpu_sample or int(max(cores, min(B, r0 * 12.0)))
            r, p, secs = oracle_sample_rate(po, n, m, okind, ns, cores, SEED_BASE)
            ores = oracle_sample_rate.last
            same = all(list(res_c[i]["stats"]) == list(ores[i].stats) and int(res_c[i]["verdict"]) == ores[i].verdict for i in range(ns))
            assert same, "oracle sample and device disagree on verdict / stats"
            line["cpu_baseline"] = {"value": r, "unit": "instances/s", "cores": cores, "kind": "port",
                                    "sample": "first %d instances of the same batch, %d threads, %.1f s" % (ns, cores, secs),
                                    "propagations_per_sec": p, "build": po.build_info(),
                                    "verdicts_and_stats_equal_device": bool(same)}
        print(json.dumps(line), flush=True)
    for e in engs:
        e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
