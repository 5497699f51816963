"""`minisat-b200` -- command-line front end with the reference's flag surface and output
(reference: src/main.rs:15-279 flags and ranges, src/lib.rs:47-157 driver and strings,
src/sat/minisat/search.rs:398-405 search table frame).

  python tools/minisat_b200.py [flags] input.cnf[.gz] [output]

Log lines go to stderr (the reference logs through env_logger to stderr), the verdict line to
stdout.  Out-of-range flag values are silently ignored, `--phase-saving` is a no-op and
`--rsync` does not exist, exactly as in the reference.  The per-learnt-limit progress rows of
the search table (search.rs:289-301) are recorded on the device by the instrumented kernel and
printed after the solve (`--verb 1`); `--no-pre` seeds the SimplifyGuard as lib.rs:36-41 does.
Solving runs on the GPU through the C ABI; there is no CPU path.
"""
import argparse
import sys
import time

import numpy as np

from . import (CORE, SIMP, SAT, UNSAT, INTERRUPTED, F_PREPROCESS, F_NO_SOLVE, F_SIMP_OFF_EMPTY, F_COUNT_TRAFFIC, Engine, EngineError,
               default_settings, dimacs, lib, mk_lit, SolveRes, Stats, STAT_NAMES)


def build_parser():
    ap = argparse.ArgumentParser(prog="minisat-b200", description="Minisat reimplementation in Rust -- B200 batched engine front end")
    ap.add_argument("--verb", choices=["0", "1", "2"], default="1", help="Verbosity level (0=silent, 1=some, 2=more)")
    ap.add_argument("--core", action="store_true", help="Use core solver")
    ap.add_argument("--strict", action="store_true", help="Validate DIMACS header during parsing")
    ap.add_argument("--pre", action="store_true")
    ap.add_argument("--no-pre", action="store_true")
    ap.add_argument("--solve", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--dimacs")
    ap.add_argument("input")
    ap.add_argument("output", nargs="?")
    for f in ("var-decay", "cla-decay", "rnd-freq", "rnd-seed", "rfirst", "rinc", "gc-frac", "simp-gc-frac"):
        ap.add_argument("--" + f, type=float)
    for f in ("ccmin-mode", "phase-saving", "min-learnts", "grow", "cl-lim", "sub-lim"):
        ap.add_argument("--" + f, type=int)
    for f in ("rnd-init", "luby", "rcheck", "asymm", "elim"):
        ap.add_argument("--" + f, action="store_true")
        ap.add_argument("--no-" + f, action="store_true")
    return ap


def settings_from_args(a):
    """main.rs:85-279: same defaults, same accepted ranges (out-of-range values are ignored)."""
    s = default_settings()
    if a.var_decay is not None and 0.0 < a.var_decay < 1.0: s.var_decay = a.var_decay
    if a.cla_decay is not None and 0.0 < a.cla_decay < 1.0: s.clause_decay = a.cla_decay
    if a.rnd_freq is not None and 0.0 <= a.rnd_freq <= 1.0: s.random_var_freq = a.rnd_freq
    if a.rnd_seed is not None and a.rnd_seed > 0.0: s.random_seed = a.rnd_seed
    if a.ccmin_mode is not None and a.ccmin_mode in (0, 1, 2): s.ccmin_mode = a.ccmin_mode
    # --phase-saving: parsed but never applied (main.rs:144 reads "phase_saving")
    if a.rnd_init: s.rnd_init_act = 1
    if a.no_rnd_init: s.rnd_init_act = 0
    if a.luby: s.luby_restart = 1
    if a.no_luby: s.luby_restart = 0
    if a.rfirst is not None and a.rfirst > 0: s.restart_first = a.rfirst
    if a.rinc is not None and a.rinc > 1.0: s.restart_inc = a.rinc
    if a.gc_frac is not None and 0.0 < a.gc_frac <= 1.0: s.garbage_frac = a.gc_frac
    if a.min_learnts is not None and a.min_learnts >= 0: s.min_learnts_lim = a.min_learnts
    if a.rcheck: s.use_rcheck = 1
    if a.no_rcheck: s.use_rcheck = 0
    if not a.core:
        if a.asymm: s.use_asymm = 1
        if a.no_asymm: s.use_asymm = 0
        if a.elim: s.use_elim = 1
        if a.no_elim: s.use_elim = 0
        if a.grow is not None and a.grow >= 0: s.grow = a.grow
        if a.cl_lim is not None and a.cl_lim >= -1: s.clause_lim = a.cl_lim
        if a.sub_lim is not None and a.sub_lim >= -1: s.subsumption_lim = a.sub_lim
        if a.simp_gc_frac is not None and 0.0 < a.simp_gc_frac <= 1.0: s.simp_garbage_frac = a.simp_gc_frac
    return s


def _mem_used_peak_mb():
    """src/util.rs:11-22: VmPeak of /proc/self/status, kB -> MB"""
    try:
        for line in open("/proc/self/status"):
            if line.startswith("VmPeak:"):
                return int(line.split()[1]) / 1024.0
    except OSError:
        pass
    return None


def _pct(num, den):
    return "NaN" if den == 0 else "%4.2f" % (num * 100.0 / den)


def print_stats(info, st, cpu_time, mem_mb):
    """lib.rs:128-157"""
    info("restarts              : %-12d" % st["restarts"])
    info("conflicts             : %-12d   (%.0f /sec)" % (st["conflicts"], st["conflicts"] / cpu_time))
    info("decisions             : %-12d   (%s %% random) (%.0f /sec)" % (
        st["decisions"], _pct(st["rnd_decisions"], st["decisions"]), st["decisions"] / cpu_time))
    info("propagations          : %-12d   (%.0f /sec)" % (st["propagations"], st["propagations"] / cpu_time))
    info("conflict literals     : %-12d   (%s %% deleted)" % (
        st["tot_literals"], _pct(st["del_literals"], st["del_literals"] + st["tot_literals"])))
    if mem_mb is not None:
        info("Memory used           : %.2f MB" % mem_mb)
    info("CPU time              : %s s" % repr(cpu_time))
    info("")


def run(argv=None, out=sys.stdout, errout=sys.stderr):
    a = build_parser().parse_args(argv)
    verb = int(a.verb)

    def info(msg):
        if verb >= 1:
            print(msg, file=errout)

    pre = not a.no_pre
    solve = not a.no_solve
    st = settings_from_args(a)
    kind = CORE if a.core else SIMP
    info("============================[ Problem Statistics ]=============================")
    info("|                                                                             |")
    t0 = time.perf_counter()
    off, lits, hv, hc = dimacs.read_csr(a.input, validate=a.strict)
    ioff = np.array([0, len(off) - 1], dtype=np.uint32)
    n_vars = int(lits.max() >> 1) + 1 if len(lits) else 0
    t_parse = time.perf_counter()
    eng = Engine(0)
    # --no-pre on a SimpSolver = preprocess() on the empty formula first, which switches the simplifier off
    # (lib.rs:38-40): from then on the instance is handled exactly like by the core solver
    run_kind = kind if (kind == CORE or pre) else CORE
    no_pre_flag = F_SIMP_OFF_EMPTY if (kind == SIMP and not pre) else 0
    res_ing, _ = eng.solve_batch(ioff, off, lits, settings=st, kind=run_kind, flags=F_NO_SOLVE | no_pre_flag, model_stride=max(n_vars, 1))
    info("|  Number of variables:  %12d                                         |" % int(res_ing[0]["n_vars"]))
    info("|  Number of clauses:    %12d                                         |" % int(res_ing[0]["n_clauses_ingest"]))
    info("|  Parse time:           %12.2f s                                       |" % (t_parse - t0))
    flags = F_PREPROCESS | (0 if solve else F_NO_SOLVE) | no_pre_flag
    if verb >= 1 and solve:
        eng.set_progress_rows(256)   # the search table's progress rows come from the instrumented kernel
        flags |= F_COUNT_TRAFFIC
    res, models = eng.solve_batch(ioff, off, lits, settings=st, kind=run_kind, flags=flags, model_stride=max(n_vars, 1))
    r = res[0]
    if int(r["status"]) != 0:
        raise EngineError(int(r["status"]), "instance did not fit the device slabs")
    info("|  Simplification time:  %12.2f s                                       |" % (time.perf_counter() - t_parse))
    info("|                                                                             |")
    stats = Stats(zip(STAT_NAMES, [int(x) for x in r["stats"]]))
    if not r["pre_ok"]:
        info("===============================================================================")
        info("Solved by simplification")
        verdict, stats = UNSAT, Stats(zip(STAT_NAMES, [0] * 8))
    elif not solve:
        info("===============================================================================")
        verdict = INTERRUPTED
    else:
        info("============================[ Search Statistics ]==============================")
        info("| Conflicts |          ORIGINAL         |          LEARNT          | Progress |")
        info("|           |    Vars  Clauses Literals |    Limit  Clauses Lit/Cl |          |")
        info("===============================================================================")
        for row in eng.progress_rows(0):   # search.rs:289-301, one row per learnt-limit bump
            info("| %9d | %7d %8d %8d | %8d %8d %6.0f | %6.3f %% |" % (int(row[0]), int(row[1]), int(row[2]), int(row[3]),
                                                                       int(row[4]), int(row[5]), row[6], row[7]))
        info("===============================================================================")
        verdict = int(r["verdict"])
    cpu_time = time.perf_counter() - t0
    print_stats(info, stats, cpu_time, _mem_used_peak_mb())
    model_lits = None
    if verdict == UNSAT:
        print("UNSATISFIABLE", file=out)
    elif verdict == SAT:
        print("SATISFIABLE", file=out)
        m = models[0]
        model_lits = [mk_lit(v, m[v] == 0) for v in range(int(r["n_vars"])) if m[v] != 2]
        ver, n_bad = eng.verify_models()          # lib.rs:114-117 self-check, on the device against the original CNF
        assert n_bad == 0 and ver[0] == 1, "SELF-CHECK FAILED"
    else:
        print("INDETERMINATE", file=out)
    if a.output:
        with open(a.output, "w") as f:
            dimacs.write_result(f, SolveRes(verdict, stats, model_lits))
    eng.close()
    return verdict


def main(argv=None):
    try:
        run(argv)
    except (EngineError, IOError) as e:
        print("minisat-b200: %s" % e, file=sys.stderr)
        return 1
    return 0            # the reference exits 0 on any verdict (main.rs:281)
