/*
 * oracle.h -- C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a single-threaded C++ restatement of mishun/minisat-rust's
 * CoreSolver / SimpSolver (reference files cited function by function in
 * oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (minisat-rust_b200/) never links, imports or executes anything in oracle/.
 *
 * PARITY PIN STATUS: the reference tree holds no stored golden vectors
 * (tests/minisat.rs compares against a live `minisat` binary, absent here, and
 * no Rust toolchain exists in this image).  Pinned offline: (1) SAT/UNSAT
 * verdicts of the 3,200 SATLIB uf and uuf instances by file name, (2) every SAT model
 * checked against the original CNF, (3) the Lit/LBool truth tables of
 * src/sat/formula.rs:183-261.  The five trace counters (restarts, conflicts,
 * decisions, propagations, conflict literals) are "parity unpinned": they are
 * the output of this restatement, not of the Rust binary.
 */
#ifndef B200SAT_ORACLE_H
#define B200SAT_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order / meaning as b200sat_settings in include/b200sat.h
 * (tests assert the two layouts are identical). Defaults: SURVEY.md App. B. */
typedef struct oracle_settings {
    double   var_decay;              /* decision_heuristic.rs:25  0.95        */
    double   clause_decay;           /* clause_db.rs:14           0.999       */
    double   random_seed;            /* decision_heuristic.rs:26  91648253.0  */
    double   random_var_freq;        /* decision_heuristic.rs:27  0.0         */
    double   restart_first;          /* search.rs:31              100.0       */
    double   restart_inc;            /* search.rs:32              2.0         */
    double   size_factor;            /* search.rs:63              1/3         */
    double   size_inc;               /* search.rs:64              1.1         */
    double   size_adjust_inc;        /* search.rs:66              1.5         */
    double   garbage_frac;           /* search.rs:176             0.20        */
    double   simp_garbage_frac;      /* simplify.rs:28            0.5         */
    uint64_t grow;                   /* simplify.rs:25            0           */
    int32_t  size_adjust_start_confl;/* search.rs:65              100         */
    int32_t  min_learnts_lim;        /* search.rs:62              0           */
    int32_t  clause_lim;             /* simplify.rs:26            20          */
    int32_t  subsumption_lim;        /* simplify.rs:27            1000        */
    int32_t  ccmin_mode;             /* conflict.rs:5-15  0 none 1 basic 2 deep */
    int32_t  phase_saving;           /* decision_heuristic.rs:6-10 0/1/2 (2=full) */
    uint8_t  luby_restart;           /* search.rs:30              1           */
    uint8_t  rnd_pol;                /* decision_heuristic.rs:29  0           */
    uint8_t  rnd_init_act;           /* decision_heuristic.rs:30  0           */
    uint8_t  use_rcheck;             /* search.rs:177             0           */
    uint8_t  use_asymm;              /* simplify.rs:29            0           */
    uint8_t  use_elim;               /* simplify.rs:30            1           */
    uint8_t  extend_model;           /* minisat.rs:117            1           */
    uint8_t  remove_satisfied;       /* clause_db.rs:13           1           */
} oracle_settings;

enum { ORACLE_UNSAT = 0, ORACLE_SAT = 1, ORACLE_INTERRUPTED = 2 };
enum { ORACLE_CORE = 0, ORACLE_SIMP = 1 };

/* flags for oracle_solve_csr */
enum {
    ORACLE_F_PREPROCESS = 1,  /* call Solver::preprocess before solve_limited (lib.rs:66)            */
    ORACLE_F_NO_SOLVE   = 2,  /* stop after preprocess (--no-solve, lib.rs:83-86) -> INTERRUPTED     */
    ORACLE_F_ZERO_STATS_ON_PRE_UNSAT = 4, /* lib.rs:75-78 / tests/minisat.rs:116-117: Stats::default() */
    ORACLE_F_NO_PRE = 32      /* --no-pre (lib.rs:36-41): SimpSolver::preprocess on the empty formula before parsing  */
};

typedef struct oracle_result {
    int32_t  verdict;         /* ORACLE_*                                                 */
    uint32_t n_vars;          /* Solver::n_vars after ingest                              */
    uint32_t n_clauses_ingest;/* Solver::n_clauses after ingest (lib.rs:56)               */
    uint32_t n_clauses_pre;   /* Solver::n_clauses after preprocess                       */
    uint32_t eliminated_vars; /* simplify.rs:388                                          */
    uint32_t pre_ok;          /* result of preprocess()                                   */
    uint64_t stats[8];        /* sat.rs:8-18 order: solves restarts decisions rnd_decisions
                                 conflicts propagations tot_literals del_literals         */
    uint64_t traffic[8];      /* SURVEY 8d counters: 0 watchers visited, 1 watchers kept,
                                 2 clause derefs (step iii), 3 tail literals examined,
                                 4 watch moves, 5 implied assignments,
                                 6 analysis words read, 7 learnt words written            */
    uint64_t trace_hash;      /* FNV-1a over (bt level, learnt lits) of every conflict    */
    double   seconds;         /* wall time of preprocess + solve                          */
    double   progress;        /* ORACLE_INTERRUPTED: the f64 of SolveRes::Interrupted (search/util.rs:5-14), else 0 */
} oracle_result;

void oracle_default_settings(oracle_settings *s);

/* Ingest exactly as dimacs.rs:146-167 would (lazy var creation in id order up
 * to the largest id seen so far), then preprocess / solve.  `lits` are
 * Lit-encoded (var<<1|sign, formula.rs:63-75).  model (may be NULL) receives
 * n_vars bytes: 1 = true, 0 = false, 2 = absent from the model. */
int oracle_solve_csr(const uint32_t *clause_off, const uint32_t *lits, uint32_t n_clauses,
                     const oracle_settings *s, int kind, int flags,
                     oracle_result *out, uint8_t *model, uint32_t model_cap);

/* Many instances, `threads` worker threads pulling indices from an atomic
 * counter.  inst_clause_off has n_inst+1 entries indexing
 * clause_off (which holds absolute offsets into lits, total_clauses+1 entries).
 * models (may be NULL) is n_inst*model_stride bytes. Returns wall seconds. */
double oracle_solve_batch_csr(const uint32_t *inst_clause_off, const uint32_t *clause_off,
                              const uint32_t *lits, uint32_t n_inst,
                              const oracle_settings *s, int kind, int flags, int threads,
                              oracle_result *out, uint8_t *models, uint32_t model_stride);

/* DIMACS(.gz) reader following dimacs.rs:171-365. Returns 0 on success, else
 * writes a message into err. Arrays are malloc'ed; free with oracle_free. */
int oracle_parse_dimacs(const char *path, int strict, uint32_t **clause_off, uint32_t **lits,
                        uint32_t *n_clauses, uint32_t *hdr_vars, uint32_t *hdr_clauses,
                        char *err, size_t errcap);
void oracle_free(void *p);

/* Synthetic generator of SURVEY 8d (SplitMix64, 3 distinct vars, sign p=1/2).
 * Writes m*k lits; clause_off is implicit (k per clause). */
void oracle_gen_ksat(uint64_t seed, uint32_t n, uint32_t m, uint32_t k, uint32_t *lits);

/* check model (bytes 0/1/2) against CSR clauses: 1 ok, 0 violated */
int oracle_check_model(const uint32_t *clause_off, const uint32_t *lits, uint32_t n_clauses,
                       const uint8_t *model, uint32_t n_vars);

/* Budget of the solve_limited call(s) made by oracle_solve_csr / _batch_csr from now on (process-wide; budget.rs:5-34,
 * < 0 = none).  An instance that runs out of budget returns ORACLE_INTERRUPTED; with resume_rounds > 0 solve_limited is
 * then called again on the returned solver, without a budget, up to that many times (sat.rs:24).  Exact for CoreSolver and
 * for SimpSolver after preprocess() (simp == None). */
void oracle_set_budget(int64_t conflict_budget, int64_t propagation_budget, int resume_rounds);

/* Solver::new_var(upol, dvar) records of the NEXT solve calls (process-wide): var v is created with flags[v]
 * (bit0 upol set, bit1 upol value = sign, bit2 not a decision var) before any clause is added, as a caller of the
 * trait does (sat.rs:31-32); n = 0 restores the DIMACS behaviour (lazy creation, (None, true)). */
void oracle_set_var_flags(const uint8_t *flags, uint32_t n);
/* with var flags: clause_nvars[c] = vars created before clause c was added (interleaved new_var / add_clause) */
void oracle_set_clause_nvars(const uint32_t *clause_nvars, uint32_t n);
/* assumptions (&[Lit]) of the NEXT solve_limited calls; n = 0 clears them (search.rs:216-232,431-435) */
void oracle_set_assumptions(const uint32_t *lits, uint32_t n);

/* rows of the reference's search table (search.rs:289-301), recorded by the following oracle_solve_csr calls
 * (single-threaded use); 8 doubles per row */
void oracle_record_progress(int on);
uint32_t oracle_get_progress(double *out, uint32_t cap_rows);

/* dump per-conflict trace (debugging divergences): up to cap records of
 * (bt_level, learnt_len, hash). Enabled per call through oracle_set_trace. */
void oracle_set_trace(uint64_t *buf, uint32_t cap_records);

#ifdef __cplusplus
}
#endif
#endif
