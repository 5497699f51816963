"""CPU tests: the DEVICE solver source (cdcl_warp.cuh), run under the test-only SIMT
emulator, is trace-exact against the oracle: verdict, the 8 stats, the per-conflict
trace hash, the traffic counters and the model."""
import ctypes as C

import numpy as np
import pytest

from common import EDGE_CASES, SEED2, batch_of, csr_of, golden_instance, load_golden, synth_batch

FLAGS = 1 | 4 | 8 | 16


@pytest.fixture(scope="module")
def emu():
    import pyemu
    pyemu.build()
    return pyemu


def compare(po, emu, ioff, coff, lits, kind, settings=None, nv=None, **kw):
    st = settings if settings is not None else po.default_settings()
    n_inst = len(ioff) - 1
    nv = nv or (int(lits.max() >> 1) + 1 if len(lits) else 1)
    ores, omodels, _ = po.solve_batch(ioff, coff, lits, kind=kind, settings=st, threads=2, model_stride=nv)
    eres, emodels = emu.solve_batch(ioff, coff, lits, st, kind=kind, flags=FLAGS, model_stride=nv, **kw)
    for i in range(n_inst):
        o, e = ores[i], eres[i]
        assert e.status == 0, i
        assert e.verdict == o.verdict, i
        assert list(e.stats) == list(o.stats), (i, list(e.stats), list(o.stats))
        assert e.trace_hash == o.trace_hash, i
        assert list(e.traffic) == list(o.traffic), (i, list(e.traffic), list(o.traffic))
        assert e.n_vars == o.n_vars and e.n_clauses_ingest == o.n_clauses_ingest, i
        assert e.n_clauses_pre == o.n_clauses_pre and e.eliminated_vars == o.eliminated_vars and e.pre_ok == o.pre_ok, i
        if o.verdict == 1:
            assert (emodels[i][:o.n_vars] == omodels[i][:o.n_vars]).all(), i


def test_core_edge_cases(po, emu):
    ioff, coff, lits = batch_of([csr_of(EDGE_CASES[k]) for k in sorted(EDGE_CASES)])
    compare(po, emu, ioff, coff, lits, po.CORE, nv=64)


@pytest.mark.parametrize("tier,seed", [(-1, 0), (-1, 11), (0, 5)])
def test_core_golden_small(po, emu, tier, seed):
    g = load_golden()
    idx = [i for i, n in enumerate(g["names"]) if n.startswith(("uf20", "uf50-01", "uf50-02", "uuf50-01"))]
    ioff, coff, lits = batch_of([golden_instance(g, i) for i in idx])
    compare(po, emu, ioff, coff, lits, po.CORE, tier_nv=tier, sched_seed=seed)


def test_core_synthetic_adversarial_schedule(po, emu):
    ioff, coff, lits = synth_batch(po, SEED2, 24, 40, 170)
    compare(po, emu, ioff, coff, lits, po.CORE, sched_seed=1234)


@pytest.mark.parametrize("variant", ["rnd_freq", "rnd_init", "rnd_pol", "ccmin_basic", "ccmin_none", "phase_none",
                                     "phase_limited", "no_luby", "small_restart", "rcheck"])
def test_core_settings_variants(po, emu, variant):
    st = po.default_settings()
    if variant == "rcheck": st.use_rcheck = 1          # search/util.rs:16-42 on every ingested clause
    if variant == "rnd_freq": st.random_var_freq = 0.1
    if variant == "rnd_init": st.rnd_init_act = 1
    if variant == "rnd_pol": st.rnd_pol = 1
    if variant == "ccmin_basic": st.ccmin_mode = 1
    if variant == "ccmin_none": st.ccmin_mode = 0
    if variant == "phase_none": st.phase_saving = 0
    if variant == "phase_limited": st.phase_saving = 1
    if variant == "no_luby": st.luby_restart = 0; st.restart_first = 10.0; st.restart_inc = 2.0
    if variant == "small_restart": st.restart_first = 5.0; st.size_adjust_start_confl = 10; st.size_factor = 0.05
    ioff, coff, lits = synth_batch(po, SEED2 + 1000, 10, 50, 213)
    compare(po, emu, ioff, coff, lits, po.CORE, settings=st, sched_seed=3)


def test_edge_cases_with_rcheck(po, emu):
    """Duplicate literals, tautologies, ingest-time units and the empty clause through is_implied first."""
    st = po.default_settings()
    st.use_rcheck = 1
    st.use_asymm = 1
    ioff, coff, lits = batch_of([csr_of(EDGE_CASES[k]) for k in sorted(EDGE_CASES)])
    compare(po, emu, ioff, coff, lits, po.CORE, settings=st, nv=64, sched_seed=5)
    compare(po, emu, ioff, coff, lits, po.SIMP, settings=st, nv=64, sched_seed=6)


def test_core_reduce_and_gc_paths(po, emu):
    # tiny learnt limit + tiny restart interval: reduceDB, garbage collection and ground
    # simplification run many times inside a short search
    st = po.default_settings()
    st.size_factor = 0.02
    st.size_adjust_start_confl = 8
    st.restart_first = 7.0
    ioff, coff, lits = synth_batch(po, SEED2 + 5000, 6, 60, 256)
    compare(po, emu, ioff, coff, lits, po.CORE, settings=st, sched_seed=9)


def test_core_slab_overflow_retry(po, emu):
    # an arena far too small for the instance forces the larger-slab retry ladder
    ioff, coff, lits = synth_batch(po, SEED2 + 7000, 3, 50, 213)
    st = po.default_settings()
    eres, _ = emu.solve_batch(ioff, coff, lits, st, kind=0, flags=FLAGS, arena_words=1200, wpool_entries=700)
    ores, _, _ = po.solve_batch(ioff, coff, lits, kind=po.CORE)
    for i in range(3):
        assert eres[i].status == 0 and eres[i].attempts > 1
        assert list(eres[i].stats) == list(ores[i].stats)


# ---------------------------------------------------------------- SimpSolver path (simp_warp.cuh)
def test_simp_edge_cases(po, emu):
    ioff, coff, lits = batch_of([csr_of(EDGE_CASES[k]) for k in sorted(EDGE_CASES)])
    compare(po, emu, ioff, coff, lits, po.SIMP, nv=64)


@pytest.mark.parametrize("tier,seed", [(-1, 0), (0, 13)])
def test_simp_golden_small(po, emu, tier, seed):
    g = load_golden()
    idx = [i for i, n in enumerate(g["names"]) if n.startswith(("uf20", "uf50-01", "uf50-02", "uuf50-01"))]
    ioff, coff, lits = batch_of([golden_instance(g, i) for i in idx])
    compare(po, emu, ioff, coff, lits, po.SIMP, tier_nv=tier, sched_seed=seed)


@pytest.mark.parametrize("name", ["2bitcomp_5.cnf.gz", "2bitmax_6.cnf.gz", "2bitadd_11.cnf.gz", "3blocks.cnf.gz"])
def test_simp_structured(po, emu, name):
    # heavy variable elimination (50..307 eliminated vars), simplifier-time garbage collection, model extension
    g = load_golden()
    i = list(g["names"]).index(name)
    ioff, coff, lits = batch_of([golden_instance(g, i)])
    compare(po, emu, ioff, coff, lits, po.SIMP, nv=int(g["n_vars"][i]), sched_seed=7)
    e, _ = emu.solve_batch(ioff, coff, lits, po.default_settings(), kind=po.SIMP, flags=FLAGS)
    assert list(e[0].stats) == [int(x) for x in g["simp_stats"][i]]
    assert e[0].eliminated_vars == int(g["simp_elim"][i]) and e[0].n_clauses_pre == int(g["simp_nclauses_pre"][i])


def test_simp_without_preprocess_call(po, emu):
    # solve_limited on a SimpSolver whose preprocess() was never called (simplify.rs:165-211)
    ioff, coff, lits = synth_batch(po, SEED2 + 50, 5, 50, 213)
    st = po.default_settings()
    ores, om, _ = po.solve_batch(ioff, coff, lits, kind=po.SIMP, flags=0, threads=2, model_stride=50)
    eres, em = emu.solve_batch(ioff, coff, lits, st, kind=po.SIMP, flags=8 | 16, model_stride=50)
    for o, e, a, b in zip(ores, eres, om, em):
        assert e.status == 0 and e.verdict == o.verdict and list(e.stats) == list(o.stats) and e.trace_hash == o.trace_hash
        if o.verdict == 1:
            assert (a[:o.n_vars] == b[:o.n_vars]).all()


@pytest.mark.parametrize("variant", ["grow", "no_elim", "cl_lim", "sub_lim", "simp_gc", "rcheck", "asymm",
                                     "asymm_no_elim", "rcheck_asymm"])
def test_simp_settings_variants(po, emu, variant):
    st = po.default_settings()
    if "rcheck" in variant: st.use_rcheck = 1          # also applies to the resolvents BVE adds
    if "asymm" in variant: st.use_asymm = 1            # simplify.rs:303-336 + search/util.rs:44-73
    if variant == "asymm_no_elim": st.use_elim = 0
    if variant == "grow": st.grow = 8
    if variant == "no_elim": st.use_elim = 0
    if variant == "cl_lim": st.clause_lim = 4
    if variant == "sub_lim": st.subsumption_lim = 3
    if variant == "simp_gc": st.simp_garbage_frac = 0.01
    g = load_golden()
    idx = [list(g["names"]).index(n) for n in ("2bitcomp_5.cnf.gz", "uf50-03.cnf.gz", "uuf50-02.cnf.gz")]
    ioff, coff, lits = batch_of([golden_instance(g, i) for i in idx])
    compare(po, emu, ioff, coff, lits, po.SIMP, settings=st, nv=int(max(g["n_vars"][i] for i in idx)), sched_seed=2)


# ---------------------------------------------------------------- portfolio mode
def test_portfolio_replica0_exact_and_sharing_sound(po, emu):
    """Replica 0 keeps the reference trace; diversified replicas that import the short learnt clauses of
    earlier replicas still return the oracle's verdict and valid models (sharing is sound)."""
    g = load_golden()
    st = po.default_settings()
    st.restart_first = 6.0          # many restart boundaries -> many import points
    for name in ("uuf50-01.cnf.gz", "uf50-03.cnf.gz", "uf50-07.cnf.gz", "uuf50-05.cnf.gz"):
        i = list(g["names"]).index(name)
        off, lits = golden_instance(g, i)
        nv = int(g["n_vars"][i])
        o, _ = po.solve_csr(off, lits, kind=po.CORE, settings=st)
        for share in (2, 3):        # PF_IGNORE_DONE alone / with PF_SHARE
            res, models = emu.portfolio(off, lits, st, 8, share=share, model_stride=nv, sched_seed=5)
            assert list(res[0].stats) == list(o.stats) and res[0].trace_hash == o.trace_hash, name
            for k in range(8):
                assert res[k].status == 0 and res[k].verdict == o.verdict, (name, k)
                if res[k].verdict == 1:
                    assert po.check_model(off, lits, models[k]), (name, k)
            if share == 3:
                assert sum(r.exported for r in res) > 0
                assert sum(r.imported for r in res) > 0, name
            else:
                assert sum(r.exported for r in res) == 0 and sum(r.imported for r in res) == 0


def test_portfolio_first_finisher_cancels_the_rest(po, emu):
    g = load_golden()
    i = list(g["names"]).index("uuf50-02.cnf.gz")
    off, lits = golden_instance(g, i)
    res, _ = emu.portfolio(off, lits, po.default_settings(), 4, share=1)   # done flag honoured
    assert res[0].status == 0 and res[0].verdict == 0
    # the emulator runs replicas one after another: replica 0 finishes, the others see `done` at their first conflict
    assert all(r.status == 1 and r.verdict == 2 for r in res[1:])


# ---------------------------------------------------------------- resident instances: conflict slices, budgets (round 2)
@pytest.mark.parametrize("kind,tier,slice_c", [(0, -1, 3), (0, 0, 5), (1, -1, 4)])
def test_conflict_slices_are_trace_neutral(po, emu, kind, tier, slice_c):
    """An instance that is parked after every few conflicts and resumed from its slab by a warp with wiped shared
    memory produces the oracle's trace: suspension is not visible in any counter, hash or model."""
    ioff, coff, lits = synth_batch(po, SEED2 + 9000, 6, 50, 213)
    st = po.default_settings()
    st.restart_first = 9.0
    st.size_factor = 0.05          # reduceDB + garbage collection between the pick-ups
    st.size_adjust_start_confl = 10
    compare(po, emu, ioff, coff, lits, kind, settings=st, tier_nv=tier, sched_seed=17, slice_conflicts=slice_c)
    _, _, n_susp = emu.solve_batch(ioff, coff, lits, st, kind=kind, flags=FLAGS, tier_nv=tier, slice_conflicts=slice_c,
                                   want_suspends=True)
    assert n_susp > 20


@pytest.mark.parametrize("kind", [0, 1])
def test_budget_interrupts_and_the_solver_resumes(po, emu, kind):
    """budget.rs:5-34 / search.rs:473-477: a conflict or propagation budget ends solve_limited with Interrupted at
    the first check after it is used up (backtracked to ground, stats kept); calling solve_limited again on the
    returned solver starts a new search (solves += 1, restart sequence from 0) on the same clause database."""
    ioff, coff, lits = synth_batch(po, SEED2 + 9100, 8, 50, 213)
    st = po.default_settings()
    for cb, pb in ((4, -1), (-1, 150), (0, -1)):
        po.set_budget(cb, pb, 0)
        try:
            ores, _, _ = po.solve_batch(ioff, coff, lits, kind=kind, settings=st, threads=1, model_stride=50)
        finally:
            po.set_budget()
        eres, _ = emu.solve_batch(ioff, coff, lits, st, kind=kind, flags=FLAGS, model_stride=50, sched_seed=3,
                                  conflict_budget=cb, propagation_budget=pb)
        n_int = 0
        for o, e in zip(ores, eres):
            assert e.status == 0 and e.verdict == o.verdict and list(e.stats) == list(o.stats)
            assert e.trace_hash == o.trace_hash
            assert e.progress == o.progress      # the f64 of SolveRes::Interrupted (search/util.rs:5-14); 0 unless interrupted
            assert o.verdict == po.INTERRUPTED or o.progress == 0.0
            n_int += o.verdict == po.INTERRUPTED
        assert n_int >= 3
        # ... and resumed without a budget: same as the oracle's second solve_limited call
        po.set_budget(cb, pb, 1)
        try:
            ores, om, _ = po.solve_batch(ioff, coff, lits, kind=kind, settings=st, threads=1, model_stride=50)
        finally:
            po.set_budget()
        eres, em = emu.solve_batch(ioff, coff, lits, st, kind=kind, flags=FLAGS, model_stride=50, sched_seed=4,
                                   conflict_budget=cb, propagation_budget=pb, resume_rounds=1, slice_conflicts=11)
        for o, e, a, b in zip(ores, eres, om, em):
            assert e.status == 0 and e.verdict == o.verdict and o.verdict != po.INTERRUPTED
            assert list(e.stats) == list(o.stats) and e.trace_hash == o.trace_hash
            if o.verdict == 1:
                assert (a[:o.n_vars] == b[:o.n_vars]).all()


def test_bcp_replay_reproduces_the_recorded_prefix(po, emu):
    """K2 (SURVEY.md 8d BCP-only figure): the replay of the recorded decisions / learnt clauses / restarts of a
    search prefix makes exactly the recorded number of propagations with exactly the recorded BCP traffic."""
    ioff, coff, lits = synth_batch(po, SEED2 + 9200, 8, 60, 256)
    st = po.default_settings()
    st.restart_first = 15.0
    for cap in (8192, 300):   # a full prefix (up to the first reduceDB / end of search) and a record that fills up
        out = emu.bcp_replay(ioff, coff, lits, st, trace_cap=cap, sched_seed=21)
        assert (out[:, 0] > 0).sum() >= 6
        for words, ok, rec_p, got_p, same in out:
            if words == 0:
                continue
            assert ok == 1 and rec_p == got_p and same == 1
            assert words <= cap


def test_hand_traces_device_source(po, emu):
    """tests/golden/hand_traces.md against the device source (emulated): a pin that does not pass through oracle.cpp."""
    from common import HAND_TRACES
    for name, (clauses, kind, over, verdict, stats, model, elim, ncl) in HAND_TRACES.items():
        st = po.default_settings()
        for k, v in over.items():
            setattr(st, k, v)
        ioff, coff, lits = batch_of([csr_of(clauses)])
        nv = max(abs(x) for c in clauses for x in c)
        res, models = emu.solve_batch(ioff, coff, lits, st, kind=kind, flags=FLAGS, model_stride=nv, sched_seed=31)
        assert res[0].status == 0 and res[0].verdict == verdict, name
        assert list(res[0].stats) == stats, (name, list(res[0].stats))
        assert res[0].eliminated_vars == elim and res[0].n_clauses_pre == ncl, name
        if model is not None:
            assert [int(x) for x in models[0][:nv]] == model, name


def test_hand_traces_call_semantics_device_source(po, emu):
    """H6 / H7 of tests/golden/hand_traces.md (false assumption; conflict budget -> Interrupted -> second solve_limited)
    against the device source under the emulator."""
    from common import HAND_TRACES_CALLS, HAND_TRACES_PROGRESS
    for name, (clauses, assumps, (cb, pb), first, second) in HAND_TRACES_CALLS.items():
        ioff, coff, lits = batch_of([csr_of(clauses)])
        nv = max(abs(x) for c in clauses for x in c)
        for rounds, want in ((0, first), (1, second)):
            if want is None:
                continue
            res, models = emu.solve_batch(ioff, coff, lits, po.default_settings(), kind=0, flags=FLAGS, model_stride=nv, sched_seed=17,
                                          conflict_budget=cb, propagation_budget=pb, resume_rounds=rounds,
                                          assumptions=assumps if assumps else None)
            assert res[0].status == 0 and res[0].verdict == want[0], (name, rounds)
            assert list(res[0].stats) == want[1], (name, rounds, list(res[0].stats))
            if len(want) > 2:
                assert [int(x) for x in models[0][:nv]] == want[2], name
            if rounds == 0 and name in HAND_TRACES_PROGRESS:
                assert res[0].progress == HAND_TRACES_PROGRESS[name], (name, res[0].progress)


@pytest.mark.parametrize("kind", [0, 1])
def test_new_var_flags_and_assumptions(po, emu, kind):
    """Solver::new_var(upol, dvar) (decision_heuristic.rs:70-102,181-192) and solve_limited's assumptions
    (search.rs:216-232,431-435) on the device, against the oracle: user polarities, non-decision vars, vars that occur
    in no clause, true / undefined / false assumptions."""
    rng = np.random.RandomState(5 + kind)
    for trial in range(6):
        n, m = 40, 150 + 5 * trial
        lits = po.gen_ksat(SEED2 + 9300 + trial, n, m)
        coff = np.arange(0, m + 1, dtype=np.uint32) * 3
        ioff = np.array([0, m], np.uint32)
        nv = n + 3                                   # three vars no clause mentions
        vf = np.zeros(nv, np.uint8)
        for v in range(nv):
            r = rng.rand()
            if r < 0.3: vf[v] |= 1 | (2 if rng.rand() < 0.5 else 0)      # user polarity
            if rng.rand() < 0.15: vf[v] |= 4                            # not a decision var
        assumps = [int((v << 1) | rng.randint(2)) for v in rng.choice(n, 3, replace=False)] if trial % 2 else []
        po.set_var_flags(vf); po.set_assumptions(assumps)
        try:
            o, om = po.solve_csr(coff, lits, kind=kind)
        finally:
            po.set_var_flags(); po.set_assumptions()
        e, em = emu.solve_batch(ioff, coff, lits, po.default_settings(), kind=kind, flags=FLAGS, model_stride=nv, sched_seed=9,
                                var_flags=vf, inst_var_off=np.array([0, nv], np.uint32), assumptions=assumps)
        assert e[0].status == 0 and e[0].verdict == o.verdict, trial
        assert list(e[0].stats) == list(o.stats), (trial, list(e[0].stats), list(o.stats))
        assert e[0].trace_hash == o.trace_hash and e[0].n_vars == o.n_vars == nv, trial
        assert e[0].eliminated_vars == o.eliminated_vars


def test_new_var_interleaved_with_add_clause(po, emu):
    """new_var calls interleaved with add_clause (the reference's DIMACS loader, dimacs.rs:146-167, and any caller that
    creates vars on demand): SimpSolver's elimination heap sees every occurrence as it arrives (elim_queue.rs:26-76), so
    the interleaving changes the preprocessing result.  The per-clause watermark (clause_nvars) reproduces it: with
    the lazy watermark the flagged run equals the plain fixture run, and on 2bitcomp_5 the up-front order differs."""
    g = load_golden()
    i = list(g["names"]).index("2bitcomp_5.cnf.gz")
    off, lits = golden_instance(g, i)
    nv = int(g["n_vars"][i])
    ioff = np.array([0, len(off) - 1], np.uint32)
    vf = np.zeros(nv, np.uint8)
    lazy, seen = [], 0
    for c in range(len(off) - 1):
        seen = max(seen, int(lits[off[c]:off[c + 1]].max() >> 1) + 1)
        lazy.append(seen)
    st = po.default_settings()
    outs = {}
    for tag, cnv in (("upfront", None), ("lazy", np.array(lazy, np.uint32)), ("early", np.minimum(np.array(lazy, np.uint32) + 7, nv))):
        po.set_var_flags(vf); po.set_clause_nvars(cnv)
        try:
            o, _ = po.solve_csr(off, lits, kind=po.SIMP)
        finally:
            po.set_var_flags(); po.set_clause_nvars()
        e, _ = emu.solve_batch(ioff, off, lits, st, kind=po.SIMP, flags=FLAGS, model_stride=nv, sched_seed=3,
                               var_flags=vf, inst_var_off=np.array([0, nv], np.uint32), clause_nvars=cnv)
        assert e[0].status == 0 and e[0].verdict == o.verdict, tag
        assert list(e[0].stats) == list(o.stats) and e[0].trace_hash == o.trace_hash, tag
        assert e[0].eliminated_vars == o.eliminated_vars and e[0].n_clauses_pre == o.n_clauses_pre, tag
        outs[tag] = (o.n_clauses_pre, list(o.stats))
    assert outs["lazy"][0] == int(g["simp_nclauses_pre"][i]) and outs["lazy"][1] == [int(x) for x in g["simp_stats"][i]]
    assert outs["upfront"] != outs["lazy"]


def test_no_pre_simp_solver_is_a_core_solver_with_a_seeded_simplify_guard(po, emu):
    """--no-pre (lib.rs:36-41): SimpSolver::preprocess runs on the EMPTY formula before parsing; the simplifier is off
    afterwards but SimplifyGuard holds (Some(0), 0), so try_simplify fires at different moments than in a fresh
    CoreSolver.  Device: kind CORE + B200SAT_F_SIMP_OFF_EMPTY == oracle SimpSolver with ORACLE_F_NO_PRE."""
    ioff, coff, lits = synth_batch(po, SEED2 + 9400, 16, 80, 341)
    st = po.default_settings()
    ores, om, _ = po.solve_batch(ioff, coff, lits, kind=po.SIMP, flags=po.F_PREPROCESS | po.F_NO_PRE, threads=2, model_stride=80)
    eres, em = emu.solve_batch(ioff, coff, lits, st, kind=po.CORE, flags=1 | 8 | 32, model_stride=80, sched_seed=4)
    for o, e in zip(ores, eres):
        assert e.status == 0 and e.verdict == o.verdict and list(e.stats) == list(o.stats) and e.trace_hash == o.trace_hash
    # (on uniform 3-SAT the seeded guard practically never changes a counter -- 0 of 400 n=100 instances -- so the test
    # pins the flag's plumbing and the oracle's reading of lib.rs:36-41, not a large effect)
