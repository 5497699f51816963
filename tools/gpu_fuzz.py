"""Randomised parity fuzz on the GPU box: ragged batches of mixed instance families and random
solver settings, CUDA path (through the C ABI) vs the oracle -- verdict, the 8 stats, trace hash,
traffic counters, preprocessing counts and models must all be identical.

    python tools/gpu_fuzz.py [seconds=300] [seed=1] [--emu] [--force=rcheck,asymm]

`--emu` runs the same device source on the test-only CPU SIMT emulator (tests/emu) instead of the GPU.

On the first mismatch the batch is dumped to gpurun_out/fuzz_fail_<iter>.npz and the exit code is 1.
Development / validation aid (uses the oracle, so it is not part of the product path).
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _b200pkg  # noqa: E402
from common import batch_of, check_models, csr_of  # noqa: E402

pkg = _b200pkg.load()
from oracle import pyoracle as po  # noqa: E402

FORCE = [a[8:] for a in sys.argv if a.startswith("--force=")]   # e.g. --force=rcheck,asymm : always switch these on
sys.argv = [a for a in sys.argv if not a.startswith("--force=")]
FORCE = FORCE[0].split(",") if FORCE else []
EMU = "--emu" in sys.argv          # run the device source on the CPU SIMT emulator instead of the GPU (slow)
if EMU:
    sys.argv.remove("--emu")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import pyemu  # noqa: E402


def ksat(rng, n, m, k):
    cls = []
    for _ in range(m):
        vs = rng.choice(n, size=k, replace=False) + 1
        cls.append([int(v) * (1 if rng.random() < 0.5 else -1) for v in vs])
    return cls


def fam_uniform(rng):
    k = int(rng.choice([3, 3, 3, 4, 5]))
    if k == 3:
        n, r = int(rng.integers(10, 60 if EMU else 180)), rng.uniform(3.6, 4.8)
    elif k == 4:
        n, r = int(rng.integers(10, 70)), rng.uniform(8.5, 10.5)
    else:
        n, r = int(rng.integers(10, 40)), rng.uniform(18, 22)
    return ksat(rng, n, int(n * r), k)


def fam_mixed(rng):
    """Ragged lengths 1..8 with duplicate literals and tautologies sprinkled in."""
    n = int(rng.integers(5, 120))
    m = int(n * rng.uniform(2.0, 4.5))
    cls = []
    for _ in range(m):
        ln = int(rng.choice([1, 2, 2, 3, 3, 3, 3, 4, 5, 8], p=None))
        if ln == 1 and rng.random() < 0.8:
            ln = 3
        c = [int(rng.integers(1, n + 1)) * (1 if rng.random() < 0.5 else -1) for _ in range(ln)]
        if rng.random() < 0.03 and c:
            c.append(-c[0])
        if rng.random() < 0.05 and c:
            c.append(c[-1])
        cls.append(c)
    if rng.random() < 0.02:
        cls.insert(int(rng.integers(0, len(cls) + 1)), [])
    return cls


def fam_two_plus_long(rng):
    n = int(rng.integers(20, 60 if EMU else 200))
    cls = ksat(rng, n, int(n * rng.uniform(0.6, 1.1)), 2)
    for _ in range(int(n * rng.uniform(0.2, 1.0))):
        ln = int(rng.integers(6, min(n, 48)))
        cls += ksat(rng, n, 1, ln)
    rng.shuffle(cls)
    return cls


def fam_planted(rng):
    n = int(rng.integers(30, 70 if EMU else 250))
    sol = rng.integers(0, 2, n)
    cls = []
    m = int(n * rng.uniform(3.5, 5.5))
    while len(cls) < m:
        c = ksat(rng, n, 1, 3)[0]
        if any((x > 0) == bool(sol[abs(x) - 1]) for x in c):
            cls.append(c)
    return cls


def fam_php(rng):
    h = int(rng.integers(2, 6))
    p = h + 1
    var = lambda i, j: i * h + j + 1
    cls = [[var(i, j) for j in range(h)] for i in range(p)]
    for j in range(h):
        for a in range(p):
            for b in range(a + 1, p):
                cls.append([-var(a, j), -var(b, j)])
    if rng.random() < 0.5:
        rng.shuffle(cls)
    return cls


def fam_xor_chain(rng):
    """Parity chain x1^x2^..^xn = b encoded with auxiliaries, twice with contradicting/agreeing parity."""
    n = int(rng.integers(4, 24))
    cls = []
    nxt = [n + 1]

    def xor3(a, b, c, par):  # a ^ b ^ c = par
        for sa in (1, -1):
            for sb in (1, -1):
                for sc_ in (1, -1):
                    neg = (sa < 0) + (sb < 0) + (sc_ < 0)
                    if (neg % 2 == 0) != bool(par):
                        cls.append([sa * a, sb * b, sc_ * c])

    def chain(order, par):
        acc = order[0]
        for x in order[1:-1]:
            t = nxt[0]
            nxt[0] += 1
            xor3(acc, x, t, 0)
            acc = t
        a, b = acc, order[-1]
        if par:
            cls.extend([[a, b], [-a, -b]])
        else:
            cls.extend([[a, -b], [-a, b]])

    vs = list(range(1, n + 1))
    chain(vs, 1)
    rng.shuffle(vs)
    chain(vs, int(rng.integers(0, 2)))
    return cls


def fam_coloring(rng):
    nv, k = int(rng.integers(8, 40)), 3
    ne = int(nv * rng.uniform(1.8, 2.6))
    var = lambda v, c: v * k + c + 1
    cls = [[var(v, c) for c in range(k)] for v in range(nv)]
    for v in range(nv):
        for a in range(k):
            for b in range(a + 1, k):
                cls.append([-var(v, a), -var(v, b)])
    for _ in range(ne):
        u, w = rng.choice(nv, 2, replace=False)
        for c in range(k):
            cls.append([-var(int(u), c), -var(int(w), c)])
    return cls


FAMILIES = [fam_uniform, fam_uniform, fam_mixed, fam_two_plus_long, fam_planted, fam_php, fam_xor_chain, fam_coloring]


def random_settings(rng, kind):
    st = pkg.default_settings()
    desc = []

    def maybe(p, name, fn):
        if rng.random() < p:
            fn()
            desc.append(name)

    maybe(0.2, "rnd_freq", lambda: setattr(st, "random_var_freq", float(rng.choice([0.02, 0.1, 0.5]))))
    maybe(0.2, "seed", lambda: setattr(st, "random_seed", float(rng.integers(1, 1 << 30))))
    maybe(0.15, "rnd_init", lambda: setattr(st, "rnd_init_act", 1))
    maybe(0.15, "rnd_pol", lambda: setattr(st, "rnd_pol", 1))
    maybe(0.3, "ccmin", lambda: setattr(st, "ccmin_mode", int(rng.integers(0, 3))))
    maybe(0.3, "phase", lambda: setattr(st, "phase_saving", int(rng.integers(0, 3))))
    # (restart_first below ~10 together with a tiny learnt DB makes some instances take millions of conflicts:
    #  correct but minutes long for one warp, so the fuzz keeps away from that corner)
    maybe(0.2, "no_luby", lambda: (setattr(st, "luby_restart", 0), setattr(st, "restart_first", float(rng.choice([10, 50])))))
    maybe(0.2, "restart", lambda: (setattr(st, "restart_first", float(rng.choice([20, 100]))),
                                   setattr(st, "restart_inc", float(rng.choice([1.5, 2.0, 3.0])))))
    maybe(0.25, "tiny_db", lambda: (setattr(st, "size_factor", float(rng.choice([0.02, 0.1]))),
                                    setattr(st, "size_adjust_start_confl", int(rng.choice([8, 30]))),
                                    setattr(st, "min_learnts_lim", int(rng.choice([0, 10])))))
    maybe(0.2, "decay", lambda: (setattr(st, "var_decay", float(rng.choice([0.8, 0.9, 0.99]))),
                                 setattr(st, "clause_decay", float(rng.choice([0.9, 0.999])))))
    maybe(0.2, "garbage", lambda: setattr(st, "garbage_frac", float(rng.choice([0.05, 0.5]))))
    maybe(0.1, "keep_sat", lambda: setattr(st, "remove_satisfied", 0))
    maybe(1.0 if "rcheck" in FORCE else 0.2, "rcheck", lambda: setattr(st, "use_rcheck", 1))
    if kind == pkg.SIMP:
        maybe(1.0 if "asymm" in FORCE else 0.2, "asymm", lambda: setattr(st, "use_asymm", 1))
        maybe(0.15, "no_elim", lambda: setattr(st, "use_elim", 0))
        maybe(0.2, "grow", lambda: setattr(st, "grow", int(rng.choice([1, 4, 16]))))
        maybe(0.2, "clause_lim", lambda: setattr(st, "clause_lim", int(rng.choice([-1, 6, 12]))))
        maybe(0.2, "subsumption_lim", lambda: setattr(st, "subsumption_lim", int(rng.choice([-1, 4, 20]))))
        maybe(0.15, "simp_garbage", lambda: setattr(st, "simp_garbage_frac", float(rng.choice([0.1, 0.9]))))
        maybe(0.1, "no_extend", lambda: setattr(st, "extend_model", 0))
    return st, "+".join(desc) or "default"


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    eng = None if EMU else pkg.Engine(0)
    flags = pkg.F_PREPROCESS | pkg.F_ZERO_STATS_ON_PRE_UNSAT | pkg.F_TRACE_HASH | pkg.F_COUNT_TRAFFIC
    t0 = time.time()
    it = total = 0
    counts = {"sat": 0, "unsat": 0}
    while time.time() - t0 < budget:
        kind = pkg.CORE if rng.random() < 0.5 else pkg.SIMP
        st, desc = random_settings(rng, kind)
        ost = po.default_settings()
        C.memmove(C.addressof(ost), C.addressof(st), C.sizeof(st))
        n_inst = int(rng.integers(8, 24)) if EMU else int(rng.integers(40, 200))
        insts = [csr_of(FAMILIES[int(rng.integers(0, len(FAMILIES)))](rng)) for _ in range(n_inst)]
        ioff, coff, lits = batch_of(insts)
        stride = max(1, max((int(l.max()) >> 1) + 1 if len(l) else 1 for _, l in insts))
        if EMU:
            eres, models = pyemu.solve_batch(ioff, coff, lits, st, kind=kind, flags=flags, model_stride=stride,
                                             sched_seed=int(rng.integers(0, 1 << 30)))
            res = [{"status": r.status, "verdict": r.verdict, "stats": list(r.stats), "trace_hash": r.trace_hash,
                    "traffic": list(r.traffic), "n_clauses_pre": r.n_clauses_pre, "eliminated_vars": r.eliminated_vars}
                   for r in eres]
            verdicts = np.array([r.verdict for r in eres])
        else:
            res, models = eng.solve_batch(ioff, coff, lits, settings=st, kind=kind, flags=flags, model_stride=stride)
            verdicts = res["verdict"]
        ores, omodels, _ = po.solve_batch(ioff, coff, lits, kind=po.CORE if kind == pkg.CORE else po.SIMP, settings=ost,
                                          threads=os.cpu_count() or 8, model_stride=stride)
        bad = None
        for i in range(n_inst):
            o = ores[i]
            ok = (res[i]["status"] == 0 and res[i]["verdict"] == o.verdict and list(res[i]["stats"]) == list(o.stats)
                  and int(res[i]["trace_hash"]) == o.trace_hash and list(res[i]["traffic"]) == list(o.traffic)
                  and res[i]["n_clauses_pre"] == o.n_clauses_pre and res[i]["eliminated_vars"] == o.eliminated_vars)
            if ok and o.verdict == 1 and st.extend_model:
                ok = bool((models[i][:o.n_vars] == omodels[i][:o.n_vars]).all())
            if not ok:
                bad = i
                break
            counts["sat" if o.verdict == 1 else "unsat"] += 1
        if bad is None and st.extend_model and check_models(ioff, coff, lits, verdicts, models) != 0:
            bad = -1
        if bad is not None:
            os.makedirs("gpurun_out", exist_ok=True)
            raw = (C.c_char * C.sizeof(st)).from_buffer_copy(st).raw
            np.savez("gpurun_out/fuzz_fail_%d.npz" % it, ioff=ioff, coff=coff, lits=lits, kind=kind, bad=bad,
                     settings=np.frombuffer(raw, np.uint8))
            i = max(bad, 0)
            print("MISMATCH iter %d kind %s settings %s instance %d: device status %d verdict %d stats %s | oracle verdict %d stats %s"
                  % (it, "core" if kind == pkg.CORE else "simp", desc, bad, res[i]["status"], res[i]["verdict"],
                     list(res[i]["stats"]), ores[i].verdict, list(ores[i].stats)), flush=True)
            sys.exit(1)
        it += 1
        total += n_inst
        print("iter %3d %-4s %4d instances ok  [%s]  %.0fs" % (it, "core" if kind == pkg.CORE else "simp", n_inst, desc,
                                                              time.time() - t0), flush=True)
    print("FUZZ OK: %d batches, %d instances (%d sat, %d unsat), seed %d, %.0f s" % (it, total, counts["sat"], counts["unsat"],
                                                                                    seed, time.time() - t0))


if __name__ == "__main__":
    main()
