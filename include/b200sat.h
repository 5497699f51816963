/*
 * b200sat.h -- C ABI of the B200-native batched CDCL engine.
 *
 * This is the drop-in boundary for minisat-rust's solver seam.  The reference
 * has no FFI layer of its own; the seam is `trait Solver` (src/sat.rs:28-36)
 * with its two implementors CoreSolver (src/sat/minisat.rs:26-103) and
 * SimpSolver (src/sat/minisat.rs:123-279).  Every entry point below names the
 * reference item it replaces.  INTEGRATION.md shows the Rust `extern "C"`
 * block + `impl Solver for B200Solver` a maintainer would add.
 *
 * Conventions
 *   - Lit / Var encodings are the reference's: Var 0-based u32,
 *     Lit = var<<1 | sign, sign 1 = negative (src/sat/formula.rs:63-75,137-154).
 *   - bool results keep the reference meaning: 0 == "formula already UNSAT".
 *   - negative int returns are engine errors (B200SAT_E_*); nothing throws.
 *   - all solving runs in sm_100a CUDA kernels; there is NO CPU fallback: every
 *     solving call fails with B200SAT_E_NO_DEVICE when no CUDA device exists.
 */
#ifndef B200SAT_H
#define B200SAT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ settings
 * One flat POD for the nested settings structs of the reference
 * (CoreSettings/SimpSettings src/sat/minisat.rs:16-23,106-120).  Field
 * defaults: SURVEY.md Appendix B.  Layout is identical to oracle_settings. */
typedef struct b200sat_settings {
    double   var_decay;               /* decision_heuristic.rs:25   0.95       */
    double   clause_decay;            /* clause_db.rs:14            0.999      */
    double   random_seed;             /* decision_heuristic.rs:26   91648253.0 */
    double   random_var_freq;         /* decision_heuristic.rs:27   0.0        */
    double   restart_first;           /* search.rs:31               100.0      */
    double   restart_inc;             /* search.rs:32               2.0        */
    double   size_factor;             /* search.rs:63               1/3        */
    double   size_inc;                /* search.rs:64               1.1        */
    double   size_adjust_inc;         /* search.rs:66               1.5        */
    double   garbage_frac;            /* search.rs:176              0.20       */
    double   simp_garbage_frac;       /* simplify.rs:28             0.5        */
    uint64_t grow;                    /* simplify.rs:25             0          */
    int32_t  size_adjust_start_confl; /* search.rs:65               100        */
    int32_t  min_learnts_lim;         /* search.rs:62               0          */
    int32_t  clause_lim;              /* simplify.rs:26             20         */
    int32_t  subsumption_lim;         /* simplify.rs:27             1000       */
    int32_t  ccmin_mode;              /* conflict.rs:5-15  0 none 1 basic 2 deep */
    int32_t  phase_saving;            /* decision_heuristic.rs:6-10 0/1/2      */
    uint8_t  luby_restart;            /* search.rs:30               1          */
    uint8_t  rnd_pol;                 /* decision_heuristic.rs:29   0          */
    uint8_t  rnd_init_act;            /* decision_heuristic.rs:30   0          */
    uint8_t  use_rcheck;              /* search.rs:177              0   is_implied before add_clause, search/util.rs:16-42 */
    uint8_t  use_asymm;               /* simplify.rs:29             0   SimpSolver only, search/util.rs:44-73 */
    uint8_t  use_elim;                /* simplify.rs:30             1          */
    uint8_t  extend_model;            /* minisat.rs:117             1          */
    uint8_t  remove_satisfied;        /* clause_db.rs:13            1          */
} b200sat_settings;

/* src/sat/minisat/budget.rs:5-34.  The reference never sets a budget (no
 * public setter exists); the struct shape is kept.  conflict/propagation
 * budgets < 0 mean "off".  `interrupt` may point at a host int another thread
 * sets non-zero to cancel (Budget::asynch_interrupt): it is forwarded to the
 * device while the kernel runs; every running instance stops at its next
 * conflict and returns B200SAT_INTERRUPTED. */
typedef struct b200sat_budget {
    int64_t       conflict_budget;
    int64_t       propagation_budget;
    volatile int *interrupt;
} b200sat_budget;

/* src/sat.rs:8-18, same field order. */
typedef struct b200sat_stats {
    uint64_t solves, restarts, decisions, rnd_decisions, conflicts, propagations,
             tot_literals, del_literals;
} b200sat_stats;

enum { B200SAT_UNSAT = 0, B200SAT_SAT = 1, B200SAT_INTERRUPTED = 2 };   /* SolveRes src/sat.rs:21-25 */
enum { B200SAT_CORE = 0, B200SAT_SIMP = 1 };                            /* CoreSolver / SimpSolver   */

enum {
    B200SAT_OK            = 0,
    B200SAT_E_NO_DEVICE   = -1,   /* no CUDA device / driver: the product has no CPU path */
    B200SAT_E_CUDA        = -2,   /* a CUDA runtime call failed (b200sat_last_error())    */
    B200SAT_E_ARG         = -3,   /* bad argument / unsupported setting                   */
    B200SAT_E_SPENT       = -4,   /* handle already consumed by solve_limited             */
    B200SAT_E_MEM         = -5    /* an instance outgrew the largest slab the device can give it */
};

/* flags of the batch calls (driver behaviour of src/lib.rs:47-126) */
enum {
    B200SAT_F_PREPROCESS = 1,     /* Solver::preprocess before solve_limited (lib.rs:66)      */
    B200SAT_F_NO_SOLVE   = 2,     /* stop after preprocess (--no-solve, lib.rs:83-86)         */
    B200SAT_F_ZERO_STATS_ON_PRE_UNSAT = 4, /* lib.rs:75-78: "Solved by simplification" reports Stats::default() */
    B200SAT_F_TRACE_HASH = 8,     /* fold (bt level, learnt lits) of every conflict into result.trace_hash */
    B200SAT_F_COUNT_TRAFFIC = 16, /* fill result.traffic (SURVEY.md 8d counters)              */
    B200SAT_F_SIMP_OFF_EMPTY = 32 /* with kind CORE: the solver is a SimpSolver whose preprocess() ran on the EMPTY
                                     formula first (--no-pre, lib.rs:36-41): the simplifier is off, but SimplifyGuard
                                     already holds (Some(0), 0) (search.rs:114-135), which changes when try_simplify fires */
};

void        b200sat_default_settings(b200sat_settings *s);   /* the Default impls, SURVEY App. B */
const char *b200sat_last_error(void);                        /* thread-local message             */
int         b200sat_device_count(void);                      /* <0: B200SAT_E_NO_DEVICE          */

/* ------------------------------------------------------------------ single solver
 * Mirrors `trait Solver` (src/sat.rs:28-36) one call per method. */
typedef struct b200sat_solver b200sat_solver;

/* CoreSolver::new (minisat.rs:90-102) / SimpSolver::new (minisat.rs:263-271) */
b200sat_solver *b200sat_new(const b200sat_settings *s, int kind);
void            b200sat_free(b200sat_solver *h);
/* Solver::n_vars / n_clauses (sat.rs:29-30).  n_clauses is exact after
 * preprocess/solve; before that it is the count of clauses handed in that are
 * not trivially consumed (tautology / duplicate-literal folding is applied). */
size_t   b200sat_n_vars(const b200sat_solver *h);
size_t   b200sat_n_clauses(const b200sat_solver *h);
/* Solver::new_var (sat.rs:31): upol -1 none / 0 false / 1 true; dvar bool. */
uint32_t b200sat_new_var(b200sat_solver *h, int upol, int dvar);
/* Solver::add_clause (sat.rs:32).  Returns 0 only when the clause is empty
 * (after duplicate removal) or the handle is already known UNSAT; UNSAT found
 * by unit propagation is reported by preprocess / solve_limited, because
 * propagation runs on the device at solve time.  Adding a
 * clause after preprocess() drops the resident device state: the next call
 * ingests the whole clause list again. */
int      b200sat_add_clause(b200sat_solver *h, const uint32_t *lits, size_t n);
/* Solver::preprocess (sat.rs:33; minisat.rs:50-55,161-191): 1 ok, 0 UNSAT, <0 error */
int      b200sat_preprocess(b200sat_solver *h, const b200sat_budget *b);
/* Solver::solve_limited (sat.rs:34).  Returns B200SAT_UNSAT/SAT/INTERRUPTED or
 * <0.  model (may be NULL) receives n_vars bytes: 1 true, 0 false, 2 absent
 * (model in var-index order, minisat.rs:68,235-239).  After preprocess() the
 * search continues from the state resident on the device (nothing is ingested
 * twice).  Budgets are honoured exactly where the reference checks them
 * (search.rs:473-477); B200SAT_INTERRUPTED leaves the handle usable -- the
 * next solve_limited() resumes it (SolveRes::Interrupted, sat.rs:24) -- while
 * SAT / UNSAT consume it.  `b->interrupt` is forwarded to the device while the
 * kernel runs.  Assumptions: search.rs:216-232 (a false one ends UNSAT). */
int      b200sat_solve_limited(b200sat_solver *h, const b200sat_budget *b,
                               const uint32_t *assumps, size_t n_assumps,
                               uint8_t *model, size_t model_cap);
/* Solver::stats (sat.rs:35; search.rs:590-601) */
void     b200sat_get_stats(const b200sat_solver *h, b200sat_stats *out);
/* the progress estimate carried by the last SolveRes::Interrupted of this handle (sat.rs:24); 0 before any */
double   b200sat_progress_estimate(const b200sat_solver *h);
/* sizeof of the structs that cross this boundary, for bindings that mirror them by hand:
 * out[0] = b200sat_result, out[1] = b200sat_settings, out[2] = b200sat_stats, out[3] = b200sat_launch_info */
void     b200sat_abi_sizes(uint32_t out[4]);

/* ------------------------------------------------------------------ batch (new: no reference analogue)
 * Many independent solvers advanced by one persistent kernel, one warp per
 * instance.  Instances are given as CSR and are ingested on the device exactly
 * as the equivalent DIMACS stream would be (dimacs.rs:146-167: lazy var
 * creation in id order, then Solver::add_clause per clause in file order). */
typedef struct b200sat_batch {
    uint32_t        n_inst;
    const uint32_t *inst_clause_off;  /* n_inst+1: clause index range of each instance */
    const uint32_t *clause_off;       /* total_clauses+1: absolute literal offsets      */
    const uint32_t *lits;             /* Lit-encoded literals                           */
} b200sat_batch;

typedef struct b200sat_result {
    int32_t  verdict;          /* B200SAT_UNSAT / SAT / INTERRUPTED                     */
    int32_t  status;           /* B200SAT_OK or B200SAT_E_MEM                           */
    uint32_t n_vars;           /* Solver::n_vars after ingest                           */
    uint32_t n_clauses_ingest; /* Solver::n_clauses after ingest (lib.rs:56)            */
    uint32_t n_clauses_pre;    /* Solver::n_clauses after preprocess                    */
    uint32_t eliminated_vars;  /* simplify.rs:388                                       */
    uint32_t pre_ok;           /* result of preprocess()                                */
    uint32_t attempts;         /* 1 + number of larger-slab retries                     */
    uint64_t stats[8];         /* b200sat_stats order                                   */
    uint64_t traffic[8];       /* visited, kept, deref, tail, moves, implied, analysis words, learnt words */
    uint64_t trace_hash;       /* FNV-1a over conflicts when B200SAT_F_TRACE_HASH       */
    uint32_t exported;         /* portfolio: clauses this replica appended to its ring  */
    uint32_t imported;         /* portfolio: clauses this replica took from the rings   */
    double   progress;         /* B200SAT_INTERRUPTED by a budget / interrupt: the f64 of SolveRes::Interrupted(progress, _),
                                  search/util.rs:5-14, taken before the trail is undone (search.rs:473-477); else 0 */
} b200sat_result;

typedef struct b200sat_engine b200sat_engine;

/* One engine per GPU (device index as in cudaSetDevice). */
b200sat_engine *b200sat_engine_create(int device);
void            b200sat_engine_destroy(b200sat_engine *e);
/* Copy a host CSR batch to the device (pinned staging + async copies on
 * `stream`, a cudaStream_t passed as void*; NULL = the legacy default stream)
 * and size the per-warp working slabs for it. */
int b200sat_engine_upload(b200sat_engine *e, const b200sat_batch *batch, void *stream);
/* Generate the SURVEY.md 8d synthetic uniform k-SAT batch directly on the
 * device (kernel gen_random_ksat; bit-identical to the host generator). */
int b200sat_engine_generate_ksat(b200sat_engine *e, uint64_t seed_base, uint32_t n_inst,
                                 uint32_t n_vars, uint32_t n_clauses, uint32_t k, void *stream);
/* Solve the resident batch (kernel cdcl_warp).  Asynchronous on `stream`
 * unless a slab overflow forces a retry pass (which synchronises). */
int b200sat_engine_solve(b200sat_engine *e, const b200sat_settings *s, int kind, int flags,
                         void *stream);
/* Asynchronous form: enqueue the solve on `stream` and return; b200sat_engine_finish() waits for it
 * and runs the larger-slab retry passes.  Engines on different streams overlap (one batch's
 * hardest instances no longer idle the GPU while the next batch starts). */
int b200sat_engine_solve_async(b200sat_engine *e, const b200sat_settings *s, int kind, int flags,
                               void *stream);
int b200sat_engine_finish(b200sat_engine *e);
/* Copy results (and models: n_inst*model_stride bytes, may be NULL) to the host; synchronises. */
int b200sat_engine_download(b200sat_engine *e, b200sat_result *results, uint8_t *models,
                            uint32_t model_stride, void *stream);
/* Kernel verify_models: check every SAT model of the last solve against the ORIGINAL resident CNF on the
 * device (dimacs.rs:90-128).  verified (n_inst bytes, may be NULL): 1 ok, 0 violated, 2 instance not SAT. */
int b200sat_engine_verify_models(b200sat_engine *e, uint8_t *verified, uint32_t *n_violated, void *stream);
/* K4: multi-CTA backward subsumption of ONE (large) instance of the resident batch with the
 * occurrence-list kernel set of csrc/preprocess_kernels.cu (signature filter, segmented sort of the occurrence
 * lists, prefix-sum compaction).  Not the reference's sequential job order: the result is the order-independent
 * set {D : some other clause C is a subset of D}; equivalence-preserving.  removed: one byte per clause of the
 * instance (may be NULL); out_clause_off (kept+1 entries, relative offsets) / out_lits: the compacted CSR (may be
 * NULL; size them with the instance's clause / literal counts); counts = {kept clauses, kept literals,
 * full subset tests, removed clauses}. */
int b200sat_engine_subsume_instance(b200sat_engine *e, uint32_t inst, uint8_t *removed, uint32_t *out_clause_off,
                                    uint32_t *out_lits, uint32_t counts[4], float *kernel_ms, void *stream);
/* Same kernel set plus self-subsuming strengthening, iterated to a fixpoint (strengthen != 0): clause D loses the
 * literal !l when some clause C with l in C has (C minus l) inside (D minus !l) (simplify.rs:276-301 semantics, not its
 * job order).  counts = {kept clauses, kept literals, rounds, strengthened literals, empty clause derived (UNSAT),
 * removed clauses}. */
int b200sat_engine_simplify_instance(b200sat_engine *e, uint32_t inst, int strengthen, uint8_t *removed,
                                     uint32_t *out_clause_off, uint32_t *out_lits, uint32_t counts[6],
                                     float *kernel_ms, void *stream);
/* K4 bounded variable elimination (simplify.rs:338-420 semantics, not its order): rounds of independent-set
 * selection + warp-per-variable elimination.  Buffers: out_clause_off 2*nc+1026, out_lits / out_estack 3*nlits+4096
 * words.  out_estack holds records [var, default unit lit, round, k, (len, lits...)*k] (the variable's literal first)
 * for model extension (elim_clauses.rs:43-63): walk the rounds backwards.  counts = {kept clauses, kept literals,
 * eliminated vars, rounds, estack words, overflow}. */
int b200sat_engine_bve_instance(b200sat_engine *e, uint32_t inst, uint32_t grow, int cl_lim, uint32_t max_rounds,
                                uint32_t *out_clause_off, uint32_t *out_lits, uint32_t *out_estack,
                                uint32_t counts[6], float *kernel_ms, void *stream);
/* Copy the resident CSR back to the host (sizes via b200sat_engine_batch_dims). */
int b200sat_engine_batch_dims(b200sat_engine *e, uint32_t *n_inst, uint64_t *n_clauses, uint64_t *n_lits);
int b200sat_engine_download_batch(b200sat_engine *e, uint32_t *inst_clause_off, uint32_t *clause_off,
                                  uint32_t *lits);
/* Tunables: warps per block / blocks per SM (0 = auto).  Launch facts of the last solve. */
int b200sat_engine_configure(b200sat_engine *e, int warps_per_block, int blocks_per_sm,
                             uint64_t arena_words, uint64_t wpool_entries);
/* Resident-instance scheduling and budgets of the following solve calls of this engine (round 2).
 *   slice:  conflicts an instance may run before it yields its warp to instances waiting in the device work ring
 *           (it is parked in its slab and re-queued; trace-neutral).  0 = never yield.  Default 20000.
 *   budget: Budget{conflict_budget, propagation_budget} of budget.rs:5-34 for every instance of the batch, checked
 *           where the reference checks it (search.rs:473-477); an instance that runs out returns
 *           B200SAT_INTERRUPTED with status OK, backtracked to ground level, its state still resident.  < 0 = none.
 *   interrupt: Budget::asynch_interrupt -- may be raised from another host thread while a solve call is running;
 *           every running instance returns B200SAT_INTERRUPTED at its next check. */
int b200sat_engine_set_slice(b200sat_engine *e, uint32_t slice_conflicts);
/* Solver::new_var(upol, dvar) (sat.rs:31) for the resident batch: flags[inst_var_off[i] + v] = B200SAT_VAR_* bits of
 * var v of instance i; inst_var_off[i+1] - inst_var_off[i] = number of vars created (may exceed the largest var a
 * clause mentions).  NULL = every var (None, true).  Call after upload, before solve; an upload resets it. */
enum { B200SAT_VAR_UPOL_SET = 1, B200SAT_VAR_UPOL_VAL = 2 /* user polarity = sign: 1 picks the negative literal */,
       B200SAT_VAR_NOT_DECISION = 4 };
int b200sat_engine_set_var_flags(b200sat_engine *e, const uint8_t *flags, const uint32_t *inst_var_off);
/* With var flags set: clause_nvars[c] = how many vars the caller had created when it added clause c (global clause index
 * of the resident batch) -- for callers that interleave new_var with add_clause, as the reference's own DIMACS loader
 * does (dimacs.rs:146-167).  SimpSolver's elimination order depends on the interleaving (elim_queue.rs:26-76); CoreSolver
 * is indifferent to it.  NULL = every var was created before the first clause (sat.rs:31-32 call order).  Call after
 * set_var_flags; an upload resets it.  The b200sat_add_clause / b200sat_new_var handle records it by itself. */
int b200sat_engine_set_clause_nvars(b200sat_engine *e, const uint32_t *clause_nvars);
/* assumptions of the following solve calls (Solver::solve_limited's &[Lit], sat.rs:34; search.rs:216-232), applied to
 * every instance of the batch; n = 0 clears them.  A false assumption ends the search UNSAT, as in the reference
 * (search.rs:431-435 drops analyze_final's result). */
int b200sat_engine_set_assumptions(b200sat_engine *e, const uint32_t *lits, uint32_t n);
/* Search again on the instances the last solve left parked on the device (B200SAT_INTERRUPTED by budget / interrupt,
 * or ingested with B200SAT_F_NO_SOLVE): nothing is ingested again.  A new search() call in the reference's sense
 * (sat.rs:24: solves += 1, restart sequence from 0).  The batch must be resident in one chunk. */
int b200sat_engine_resume(b200sat_engine *e, void *stream);
int b200sat_engine_resume_async(b200sat_engine *e, void *stream);
int b200sat_engine_set_budget(b200sat_engine *e, int64_t conflict_budget, int64_t propagation_budget);
int b200sat_engine_interrupt(b200sat_engine *e, int raised);
typedef struct b200sat_launch_info {
    uint32_t grid, block, smem_bytes, warps_total, tier /*0 smem var state, 1 global*/, launches;
    uint64_t slab_bytes;
    float    last_kernel_ms;   /* CUDA-event time of all launches (ingest + search) of the last solve */
    float    search_ms;        /* CUDA-event time of the search kernel(s) alone (first chunk of each attempt) */
    uint32_t suspends;         /* times an instance was parked at the end of a conflict slice and re-queued */
    uint32_t resident_items;   /* instances whose state is resident at once (slabs allocated) */
    float    tail_ms;          /* search kernel: time between the first warp that found the work ring empty and the
                                  last warp's exit (device %globaltimer) -- the part of the launch that is bound by
                                  its longest-running instances */
    float    span_ms;          /* search kernel: first block's start to last warp's exit (device %globaltimer) */
    uint32_t max_warps;        /* warps of the search kernel that can be resident on this GPU at once (SMs x blocks/SM) */
    uint32_t rescued;          /* free-running mode: instances decided by a helper replica (b200sat_engine_set_rescue) */
} b200sat_launch_info;
int b200sat_engine_launch_info(b200sat_engine *e, b200sat_launch_info *out);

/* K4 as ONE preprocessor (round 2, csrc/preprocess_coop.cu): backward subsumption, self-subsuming strengthening and
 * bounded variable elimination of instance `inst` of the resident batch, interleaved until nothing changes
 * (simplify.rs:213-274 as passes) inside one persistent cooperative kernel: multi-block scans, grid-wide barriers
 * instead of launches (3 launches per instance).  Reference semantics (subsumes.rs:10-49, simplify.rs:338-420,567-591,
 * formula/util.rs:45-77), not its job order: the result is equisatisfiable and every model of it extends to a model of
 * the input.  out_clause_off / out_lits: the reduced CNF as CSR (capacity 2*nc+1024 / 3*nlits+4096); out_estack: the
 * eliminated-clause records; counts[16]: [0] kept clauses [1] kept literals [2] eliminated vars [3] subsumed clauses
 * [4] strengthened literals [5] outer iterations [6] BVE rounds [7] subsume/strengthen rounds [8] stack words
 * [9] overflow [10] empty clause derived (UNSAT) [11] kernel launches. */
int b200sat_engine_preprocess_instance(b200sat_engine *e, uint32_t inst, uint32_t grow, int cl_lim, uint32_t max_outer,
                                       uint32_t *out_clause_off, uint32_t *out_lits, uint32_t *out_estack,
                                       uint32_t counts[16], float *kernel_ms, void *stream);
/* ElimClauses::extend_model (elim_clauses.rs:43-63) ON THE DEVICE over the stack the last preprocess_instance left there */
int b200sat_engine_k4_extend_model(b200sat_engine *e, uint8_t *model, uint32_t n_vars, void *stream);

/* Progress rows of the reference's search table (search.rs:289-301: printed whenever the learnt-clause limit is bumped),
 * recorded on the device for every instance solved with B200SAT_F_COUNT_TRAFFIC while rows > 0.  One row = 8 doubles:
 * conflicts, free decision vars, original clauses, their literals, learnt limit, learnts, literals per learnt, progress %. */
int b200sat_engine_set_progress_rows(b200sat_engine *e, uint32_t rows);
int b200sat_engine_download_progress(b200sat_engine *e, uint32_t inst, double *out, uint32_t cap_rows, uint32_t *n_rows);

/* Free-running mode with straggler rescue -- NOT the reference trace (default: off).  Once the device work ring is empty,
 * warps with nothing left to do clone the post-ingest state of instances that are still running into their own slabs
 * and search them as diversified helper replicas (the portfolio machinery: units / binaries exchanged through a ring
 * per running warp, first finisher claims the instance).  The original replica exports but never imports, so an
 * instance it decides itself still carries the reference's counters; an instance a helper decided first has
 * `attempts & 0x100` set and the helper's counters.  Verdicts are exact either way and every SAT model satisfies the
 * original formula (b200sat_engine_verify_models).  This is what removes the tail of a batch: in trace mode a launch
 * lasts as long as its hardest instance on ONE warp. */
int b200sat_engine_set_rescue(b200sat_engine *e, int on);

/* K2 bcp_replay (measurement; SURVEY.md 8d "BCP-only figure", reference function watches.rs:64-152): the resident
 * batch is solved once with the instrumented kernels while every instance's search prefix is recorded (decisions,
 * learnt clauses + backtrack levels, restarts, up to the first database reduction or ground simplification); the
 * replay kernel then re-executes exactly those prefixes with unit propagation as the only real work.  The counted
 * replay must reproduce the recorded propagation count and BCP traffic counters of every instance (`mismatches`). */
typedef struct b200sat_replay_info {
    uint32_t instances;     /* instances with a recorded prefix                                   */
    uint32_t mismatches;    /* replays that left the recorded trace (must be 0)                   */
    uint64_t propagations;  /* literals dequeued by the replayed prefixes                         */
    uint64_t bcp_bytes;     /* SURVEY.md 8d algorithmic bytes, BCP terms (1)-(5) only             */
    float    ms_min, ms_avg;/* CUDA-event time of the plain replay kernel over `reps` launches    */
    uint32_t warps;         /* resident warps of the replay launch                                */
} b200sat_replay_info;
int b200sat_engine_bcp_replay(b200sat_engine *e, const b200sat_settings *s, int reps,
                              b200sat_replay_info *out, void *stream);

/* upload + solve + download with HOST buffers (the end-to-end call). */
int b200sat_solve_batch(b200sat_engine *e, const b200sat_batch *batch, const b200sat_settings *s,
                        int kind, int flags, b200sat_result *results, uint8_t *models,
                        uint32_t model_stride);

/* ------------------------------------------------------------------ portfolio (new: SURVEY.md 8e)
 * Many seed/heuristic-diversified replicas of ONE instance (instance 0 of the resident batch), one
 * warp each.  Replicas exchange short learnt clauses (units, binaries, LBD <= 2) through one
 * append-only ring per GPU; rings of other GPUs are mapped with CUDA IPC and read/written over
 * NVLink P2P.  The first replica that finishes raises a `done` word in every ring and the others
 * stop at their next conflict (the Budget / interrupt analogue, budget.rs:5-34).  Replica 0 keeps the
 * base settings: with sharing off its trace is bit-exact the single-instance one. */
enum { B200SAT_PF_SHARE = 1 /* units + binaries */, B200SAT_PF_IGNORE_DONE = 2,
       B200SAT_PF_SHARE_LBD2 = 4 /* also LBD <= 2 clauses of <= 8 literals */,
       B200SAT_PF_KEEP_RING = 8 /* do not clear the local ring first: the caller cleared all rings and synchronised the ranks */ };
/* settings of global replica `replica` derived from `base` (replica 0 = base) -- SURVEY.md T19 */
void b200sat_diversify(const b200sat_settings *base, uint32_t replica, b200sat_settings *out);
int  b200sat_engine_ring_create(b200sat_engine *e, uint32_t payload_words, uint32_t n_rings, uint32_t my_slot);
int  b200sat_engine_ring_reset(b200sat_engine *e, void *stream);            /* zero the local ring */
int  b200sat_engine_ring_ipc_handle(b200sat_engine *e, uint8_t handle[64]); /* cudaIpcMemHandle_t of the local ring */
int  b200sat_engine_ring_open_peer(b200sat_engine *e, uint32_t slot, const uint8_t handle[64]);
/* Launch n_replicas replicas (global ids replica_base ...) of instance 0 of the resident batch (CoreSolver).
 * Results (one b200sat_result + model per REPLICA) are fetched with b200sat_engine_download_replicas. */
/* map the ring of another engine of the SAME process (another GPU) into slot `slot` through CUDA peer access */
int  b200sat_engine_ring_open_local_peer(b200sat_engine *e, uint32_t slot, b200sat_engine *owner);
/* clauses one replica imports per restart boundary (0 = no cap; default 64).  Every replica also skips clauses whose
 * order-independent hash it has imported before (8192-bit filter per replica). */
int  b200sat_engine_set_import_cap(b200sat_engine *e, uint32_t cap);
int  b200sat_portfolio_solve(b200sat_engine *e, const b200sat_settings *base, uint32_t n_replicas,
                             uint32_t replica_base, uint32_t share_flags, int flags, void *stream);
int  b200sat_engine_download_replicas(b200sat_engine *e, b200sat_result *results, uint8_t *models,
                                      uint32_t model_stride, void *stream);

/* ------------------------------------------------------------------ DIMACS I/O (host side)
 * DIMACS(.gz) file -> CSR, following the reference's parser and its quirks (dimacs.rs:16-43,171-365).
 * Arrays are malloc'ed; release them with b200sat_free_buffer.  Returns 0 or B200SAT_E_ARG + message. */
int  b200sat_read_dimacs(const char *path, int strict, uint32_t **clause_off, uint32_t **lits, uint32_t *n_clauses,
                         uint32_t *hdr_vars, uint32_t *hdr_clauses, char *err, size_t errcap);
void b200sat_free_buffer(void *p);
/* dimacs::write_result (dimacs.rs:46-70): "UNSAT" | "INDET" | "SAT\n<lits> 0" */
int  b200sat_write_result(const char *path, int verdict, const uint8_t *model, uint32_t n_vars);

/* Asynchronous end-to-end form: H2D + solve + D2H enqueued on `stream`; buffers must stay valid
 * (and should be page-locked) until b200sat_solve_batch_finish() returns. */
int b200sat_solve_batch_async(b200sat_engine *e, const b200sat_batch *batch, const b200sat_settings *s,
                              int kind, int flags, b200sat_result *results, uint8_t *models,
                              uint32_t model_stride, void *stream);
int b200sat_solve_batch_finish(b200sat_engine *e);

#ifdef __cplusplus
}
#endif
#endif
