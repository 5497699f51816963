"""minisat-rust_b200 -- host-side mirror of minisat-rust's solver interface over the
B200 CUDA engine (libb200sat.so, C ABI in include/b200sat.h).

Mirrors, with the same names and argument meaning (reference paths relative to
/root/reference/src):
  * ``Solver`` trait               sat.rs:28-36       -> class ``Solver`` (``CoreSolver`` / ``SimpSolver``)
  * ``SolveRes`` / ``Stats``       sat.rs:8-25        -> ``SolveRes``, ``Stats``
  * ``Lit`` / ``Var`` encoding     sat/formula.rs:63-75,137-154 -> ``mk_lit`` / ``lit_var`` / ``lit_sign``
  * settings with Default impls    SURVEY.md App. B   -> ``Settings`` / ``default_settings()``
  * DIMACS reader / result writer  sat/dimacs.rs      -> ``dimacs.parse`` / ``dimacs.write_result``
plus the batch engine the reference does not have (``Engine``).

All solving happens in the CUDA library.  If the library is missing or no CUDA
device exists, solving calls raise ``EngineError`` -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SAT_LIB") or os.path.join(_HERE, "libb200sat.so")

CORE, SIMP = 0, 1
UNSAT, SAT, INTERRUPTED = 0, 1, 2
F_PREPROCESS, F_NO_SOLVE, F_ZERO_STATS_ON_PRE_UNSAT, F_TRACE_HASH, F_COUNT_TRAFFIC = 1, 2, 4, 8, 16
F_SIMP_OFF_EMPTY = 32
E_NO_DEVICE, E_CUDA, E_ARG, E_SPENT, E_MEM = -1, -2, -3, -4, -5
STAT_NAMES = ("solves", "restarts", "decisions", "rnd_decisions", "conflicts", "propagations",
              "tot_literals", "del_literals")
TRAFFIC_NAMES = ("visited", "kept", "deref", "tail", "moves", "implied", "an_words", "learnt_words")


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200sat error %d: %s" % (code, msg))
        self.code = code


class Settings(C.Structure):
    """b200sat_settings (include/b200sat.h); defaults = the reference's Default impls."""
    _fields_ = [
        ("var_decay", C.c_double), ("clause_decay", C.c_double), ("random_seed", C.c_double),
        ("random_var_freq", C.c_double), ("restart_first", C.c_double), ("restart_inc", C.c_double),
        ("size_factor", C.c_double), ("size_inc", C.c_double), ("size_adjust_inc", C.c_double),
        ("garbage_frac", C.c_double), ("simp_garbage_frac", C.c_double), ("grow", C.c_uint64),
        ("size_adjust_start_confl", C.c_int32), ("min_learnts_lim", C.c_int32),
        ("clause_lim", C.c_int32), ("subsumption_lim", C.c_int32), ("ccmin_mode", C.c_int32),
        ("phase_saving", C.c_int32), ("luby_restart", C.c_uint8), ("rnd_pol", C.c_uint8),
        ("rnd_init_act", C.c_uint8), ("use_rcheck", C.c_uint8), ("use_asymm", C.c_uint8),
        ("use_elim", C.c_uint8), ("extend_model", C.c_uint8), ("remove_satisfied", C.c_uint8),
    ]


class Budget(C.Structure):
    _fields_ = [("conflict_budget", C.c_int64), ("propagation_budget", C.c_int64),
                ("interrupt", C.POINTER(C.c_int))]


class StatsStruct(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in STAT_NAMES]


class Result(C.Structure):
    """b200sat_result"""
    _fields_ = [
        ("verdict", C.c_int32), ("status", C.c_int32), ("n_vars", C.c_uint32),
        ("n_clauses_ingest", C.c_uint32), ("n_clauses_pre", C.c_uint32), ("eliminated_vars", C.c_uint32),
        ("pre_ok", C.c_uint32), ("attempts", C.c_uint32),
        ("stats", C.c_uint64 * 8), ("traffic", C.c_uint64 * 8), ("trace_hash", C.c_uint64),
        ("exported", C.c_uint32), ("imported", C.c_uint32), ("progress", C.c_double),
    ]

    def stats_dict(self):
        return dict(zip(STAT_NAMES, [int(x) for x in self.stats]))


RESULT_DTYPE = np.dtype([
    ("verdict", np.int32), ("status", np.int32), ("n_vars", np.uint32), ("n_clauses_ingest", np.uint32),
    ("n_clauses_pre", np.uint32), ("eliminated_vars", np.uint32), ("pre_ok", np.uint32), ("attempts", np.uint32),
    ("stats", np.uint64, (8,)), ("traffic", np.uint64, (8,)), ("trace_hash", np.uint64),
    ("exported", np.uint32), ("imported", np.uint32), ("progress", np.float64)])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class BatchStruct(C.Structure):
    _fields_ = [("n_inst", C.c_uint32), ("inst_clause_off", C.POINTER(C.c_uint32)),
                ("clause_off", C.POINTER(C.c_uint32)), ("lits", C.POINTER(C.c_uint32))]


class LaunchInfo(C.Structure):
    _fields_ = [("grid", C.c_uint32), ("block", C.c_uint32), ("smem_bytes", C.c_uint32),
                ("warps_total", C.c_uint32), ("tier", C.c_uint32), ("launches", C.c_uint32),
                ("slab_bytes", C.c_uint64), ("last_kernel_ms", C.c_float), ("search_ms", C.c_float),
                ("suspends", C.c_uint32), ("resident_items", C.c_uint32), ("tail_ms", C.c_float),
                ("span_ms", C.c_float), ("max_warps", C.c_uint32), ("rescued", C.c_uint32)]


class ReplayInfo(C.Structure):
    _fields_ = [("instances", C.c_uint32), ("mismatches", C.c_uint32), ("propagations", C.c_uint64),
                ("bcp_bytes", C.c_uint64), ("ms_min", C.c_float), ("ms_avg", C.c_float), ("warps", C.c_uint32)]


def traffic_bytes(t):
    """algorithmic bytes from the eight traffic counters."""
    visited, kept, deref, tail, moves, implied, an_words, learnt_words = [int(x) for x in t]
    return (8 * visited + 8 * kept + 4 * deref + 4 * (2 * deref + tail) + 16 * moves + 12 * implied
            + 4 * an_words + 4 * learnt_words)


# every symbol include/b200sat.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "b200sat_default_settings", "b200sat_last_error", "b200sat_device_count",
    "b200sat_new", "b200sat_free", "b200sat_n_vars", "b200sat_n_clauses", "b200sat_new_var",
    "b200sat_add_clause", "b200sat_preprocess", "b200sat_solve_limited", "b200sat_get_stats",
    "b200sat_progress_estimate", "b200sat_abi_sizes",
    "b200sat_engine_create", "b200sat_engine_destroy", "b200sat_engine_upload",
    "b200sat_engine_generate_ksat", "b200sat_engine_solve", "b200sat_engine_download",
    "b200sat_engine_batch_dims", "b200sat_engine_download_batch", "b200sat_engine_configure",
    "b200sat_engine_launch_info", "b200sat_solve_batch",
    "b200sat_engine_solve_async", "b200sat_engine_finish", "b200sat_solve_batch_async", "b200sat_solve_batch_finish",
    "b200sat_diversify", "b200sat_engine_ring_create", "b200sat_engine_ring_reset", "b200sat_engine_ring_ipc_handle",
    "b200sat_engine_ring_open_peer", "b200sat_portfolio_solve", "b200sat_engine_download_replicas",
    "b200sat_engine_verify_models", "b200sat_engine_subsume_instance", "b200sat_engine_simplify_instance", "b200sat_engine_bve_instance", "b200sat_read_dimacs", "b200sat_free_buffer", "b200sat_write_result",
    "b200sat_engine_set_slice", "b200sat_engine_set_budget", "b200sat_engine_interrupt",
    "b200sat_engine_bcp_replay", "b200sat_engine_set_var_flags", "b200sat_engine_set_clause_nvars", "b200sat_engine_set_assumptions",
    "b200sat_engine_resume", "b200sat_engine_resume_async", "b200sat_engine_set_import_cap",
    "b200sat_engine_ring_open_local_peer", "b200sat_engine_set_progress_rows", "b200sat_engine_download_progress",
    "b200sat_engine_set_rescue", "b200sat_engine_preprocess_instance", "b200sat_engine_k4_extend_model",
)
PF_SHARE, PF_IGNORE_DONE, PF_SHARE_LBD2, PF_KEEP_RING = 1, 2, 4, 8

_lib = None


def lib():
    """Load libb200sat.so (the CUDA product library).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(E_NO_DEVICE, "CUDA library %s is not built (run __graft_entry__.build()); "
                                       "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    u32p, u8p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_void_p
    L.b200sat_default_settings.argtypes = [C.POINTER(Settings)]
    L.b200sat_last_error.restype = C.c_char_p
    L.b200sat_device_count.restype = C.c_int
    L.b200sat_new.argtypes = [C.POINTER(Settings), C.c_int]
    L.b200sat_new.restype = vp
    L.b200sat_free.argtypes = [vp]
    L.b200sat_n_vars.argtypes = [vp]
    L.b200sat_n_vars.restype = C.c_size_t
    L.b200sat_n_clauses.argtypes = [vp]
    L.b200sat_n_clauses.restype = C.c_size_t
    L.b200sat_new_var.argtypes = [vp, C.c_int, C.c_int]
    L.b200sat_new_var.restype = C.c_uint32
    L.b200sat_add_clause.argtypes = [vp, u32p, C.c_size_t]
    L.b200sat_add_clause.restype = C.c_int
    L.b200sat_preprocess.argtypes = [vp, C.POINTER(Budget)]
    L.b200sat_preprocess.restype = C.c_int
    L.b200sat_solve_limited.argtypes = [vp, C.POINTER(Budget), u32p, C.c_size_t, u8p, C.c_size_t]
    L.b200sat_solve_limited.restype = C.c_int
    L.b200sat_get_stats.argtypes = [vp, C.POINTER(StatsStruct)]
    L.b200sat_progress_estimate.argtypes = [vp]
    L.b200sat_progress_estimate.restype = C.c_double
    L.b200sat_abi_sizes.argtypes = [C.POINTER(C.c_uint32)]
    L.b200sat_abi_sizes.restype = None
    L.b200sat_engine_create.argtypes = [C.c_int]
    L.b200sat_engine_create.restype = vp
    L.b200sat_engine_destroy.argtypes = [vp]
    L.b200sat_engine_upload.argtypes = [vp, C.POINTER(BatchStruct), vp]
    L.b200sat_engine_generate_ksat.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.b200sat_engine_solve.argtypes = [vp, C.POINTER(Settings), C.c_int, C.c_int, vp]
    L.b200sat_engine_download.argtypes = [vp, vp, u8p, C.c_uint32, vp]
    L.b200sat_engine_batch_dims.argtypes = [vp, u32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.b200sat_engine_download_batch.argtypes = [vp, u32p, u32p, u32p]
    L.b200sat_engine_configure.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
    L.b200sat_engine_launch_info.argtypes = [vp, C.POINTER(LaunchInfo)]
    L.b200sat_engine_bcp_replay.argtypes = [vp, C.POINTER(Settings), C.c_int, C.POINTER(ReplayInfo), vp]
    L.b200sat_engine_set_var_flags.argtypes = [vp, u8p, u32p]
    L.b200sat_engine_set_clause_nvars.argtypes = [vp, u32p]
    L.b200sat_engine_set_assumptions.argtypes = [vp, u32p, C.c_uint32]
    L.b200sat_engine_resume.argtypes = [vp, vp]
    L.b200sat_engine_resume_async.argtypes = [vp, vp]
    L.b200sat_engine_set_progress_rows.argtypes = [vp, C.c_uint32]
    L.b200sat_engine_download_progress.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double), C.c_uint32, u32p]
    L.b200sat_engine_set_rescue.argtypes = [vp, C.c_int]
    L.b200sat_engine_preprocess_instance.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, u32p, u32p, u32p, u32p,
                                                     C.POINTER(C.c_float), vp]
    L.b200sat_engine_k4_extend_model.argtypes = [vp, u8p, C.c_uint32, vp]
    L.b200sat_engine_set_import_cap.argtypes = [vp, C.c_uint32]
    L.b200sat_engine_ring_open_local_peer.argtypes = [vp, C.c_uint32, vp]
    L.b200sat_engine_set_slice.argtypes = [vp, C.c_uint32]
    L.b200sat_engine_set_budget.argtypes = [vp, C.c_int64, C.c_int64]
    L.b200sat_engine_interrupt.argtypes = [vp, C.c_int]
    L.b200sat_solve_batch.argtypes = [vp, C.POINTER(BatchStruct), C.POINTER(Settings), C.c_int, C.c_int,
                                      vp, u8p, C.c_uint32]
    L.b200sat_engine_solve_async.argtypes = [vp, C.POINTER(Settings), C.c_int, C.c_int, vp]
    L.b200sat_engine_finish.argtypes = [vp]
    L.b200sat_solve_batch_async.argtypes = [vp, C.POINTER(BatchStruct), C.POINTER(Settings), C.c_int, C.c_int,
                                            vp, u8p, C.c_uint32, vp]
    L.b200sat_solve_batch_finish.argtypes = [vp]
    L.b200sat_diversify.argtypes = [C.POINTER(Settings), C.c_uint32, C.POINTER(Settings)]
    L.b200sat_diversify.restype = None
    L.b200sat_engine_ring_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.b200sat_engine_ring_reset.argtypes = [vp, vp]
    L.b200sat_engine_ring_ipc_handle.argtypes = [vp, u8p]
    L.b200sat_engine_ring_open_peer.argtypes = [vp, C.c_uint32, u8p]
    L.b200sat_portfolio_solve.argtypes = [vp, C.POINTER(Settings), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp]
    L.b200sat_engine_download_replicas.argtypes = [vp, vp, u8p, C.c_uint32, vp]
    L.b200sat_engine_verify_models.argtypes = [vp, u8p, u32p, vp]
    L.b200sat_engine_subsume_instance.argtypes = [vp, C.c_uint32, u8p, u32p, u32p, u32p, C.POINTER(C.c_float), vp]
    L.b200sat_engine_bve_instance.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, u32p, u32p, u32p, u32p,
                                              C.POINTER(C.c_float), vp]
    L.b200sat_engine_simplify_instance.argtypes = [vp, C.c_uint32, C.c_int, u8p, u32p, u32p, u32p, C.POINTER(C.c_float), vp]
    L.b200sat_read_dimacs.argtypes = [C.c_char_p, C.c_int, C.POINTER(u32p), C.POINTER(u32p), u32p, u32p, u32p,
                                      C.c_char_p, C.c_size_t]
    L.b200sat_free_buffer.argtypes = [vp]
    L.b200sat_free_buffer.restype = None
    L.b200sat_write_result.argtypes = [C.c_char_p, C.c_int, u8p, C.c_uint32]
    # the hand-written mirrors below must have the library's struct sizes: fail loudly instead of reading garbage
    sizes = (C.c_uint32 * 4)()
    L.b200sat_abi_sizes(sizes)
    mine = (C.sizeof(Result), C.sizeof(Settings), C.sizeof(StatsStruct), C.sizeof(LaunchInfo))
    if tuple(sizes) != mine:
        raise EngineError(E_ARG, "libb200sat.so struct sizes %s != binding %s (stale build?)" % (tuple(sizes), mine))
    _lib = L
    return L


def _check(rc, allow=()):
    if rc < 0 and rc not in allow:
        raise EngineError(rc, lib().b200sat_last_error().decode())
    return rc


def default_settings():
    s = Settings()
    lib().b200sat_default_settings(C.byref(s))
    return s


# ---- Lit / Var algebra, sat/formula.rs:63-75,137-154
def mk_lit(var, sign):
    return (var << 1) | (1 if sign else 0)


def lit_var(lit):
    return lit >> 1


def lit_sign(lit):
    return (lit & 1) != 0


def lit_not(lit):
    return lit ^ 1


class Stats(dict):
    """sat.rs:8-18"""
    def __getattr__(self, k):
        return self[k]


class SolveRes:
    """sat.rs:21-25: UnSAT(stats) | SAT(model, stats) | Interrupted(progress, solver)"""
    def __init__(self, kind, stats, model=None, progress=0.0):
        self.kind, self.stats, self.model, self.progress = kind, stats, model, progress

    def is_sat(self):
        return self.kind == SAT

    def is_unsat(self):
        return self.kind == UNSAT


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


class Solver:
    """`trait Solver` (sat.rs:28-36) over one C-ABI handle.  Use CoreSolver / SimpSolver."""
    KIND = CORE

    def __init__(self, settings=None):
        self.settings = settings if settings is not None else default_settings()
        self._h = lib().b200sat_new(C.byref(self.settings), self.KIND)
        if not self._h:
            raise EngineError(E_ARG, lib().b200sat_last_error().decode())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.b200sat_free(h)

    def n_vars(self):
        return lib().b200sat_n_vars(self._h)

    def n_clauses(self):
        return lib().b200sat_n_clauses(self._h)

    def new_var(self, upol=None, dvar=True):
        return lib().b200sat_new_var(self._h, -1 if upol is None else int(bool(upol)), int(dvar))

    def add_clause(self, clause):
        a, p = _u32(list(clause) if len(clause) else [0])
        return bool(lib().b200sat_add_clause(self._h, p, len(clause)))

    @staticmethod
    def _budget(budget):
        """budget: None, a Budget struct, or (conflict_budget, propagation_budget)"""
        if budget is None:
            return None
        if isinstance(budget, Budget):
            return C.byref(budget)
        b = Budget(int(budget[0]), int(budget[1]), None)
        return C.byref(b)

    def preprocess(self, budget=None):
        return bool(_check(lib().b200sat_preprocess(self._h, self._budget(budget))))

    def solve_limited(self, budget=None, assumptions=()):
        """-> SolveRes.  An INTERRUPTED result leaves the solver usable (SolveRes::Interrupted(progress, solver),
        sat.rs:24): call solve_limited again to continue; SAT / UNSAT consume it."""
        n = self.n_vars()
        model = np.full(max(n, 1), 2, dtype=np.uint8)
        a, ap = _u32(list(assumptions) if len(assumptions) else [0])
        rc = _check(lib().b200sat_solve_limited(self._h, self._budget(budget), ap, len(assumptions),
                                                model.ctypes.data_as(C.POINTER(C.c_uint8)), n))
        st = self.stats()
        if rc == SAT:
            # model as Vec<Lit> in var order (minisat.rs:68,235-239)
            lits = [mk_lit(v, model[v] == 0) for v in range(n) if model[v] != 2]
            return SolveRes(SAT, st, lits)
        return SolveRes(rc, st, progress=float(lib().b200sat_progress_estimate(self._h)) if rc == INTERRUPTED else 0.0)

    def stats(self):
        s = StatsStruct()
        lib().b200sat_get_stats(self._h, C.byref(s))
        return Stats((k, int(getattr(s, k))) for k in STAT_NAMES)


class CoreSolver(Solver):
    """sat/minisat.rs:26-103"""
    KIND = CORE


class SimpSolver(Solver):
    """sat/minisat.rs:123-279"""
    KIND = SIMP


class Engine:
    """The batch engine (no reference analogue): many independent solvers advanced by one
    persistent warp-per-instance kernel on one GPU."""

    def __init__(self, device=0):
        self._e = lib().b200sat_engine_create(device)
        if not self._e:
            raise EngineError(E_NO_DEVICE, lib().b200sat_last_error().decode())
        self.device = device
        self._keep = None

    def close(self):
        e, self._e = self._e, None
        if e and _lib is not None:
            _lib.b200sat_engine_destroy(e)

    def __del__(self):
        self.close()

    def configure(self, warps_per_block=0, blocks_per_sm=0, arena_words=0, wpool_entries=0):
        _check(lib().b200sat_engine_configure(self._e, warps_per_block, blocks_per_sm, arena_words, wpool_entries))

    def bcp_replay(self, settings=None, reps=3, stream=None):
        """K2: BCP-only replay of the recorded search prefixes of the resident batch (measurement)"""
        s = settings if settings is not None else default_settings()
        ri = ReplayInfo()
        _check(lib().b200sat_engine_bcp_replay(self._e, C.byref(s), reps, C.byref(ri), stream))
        return {k: getattr(ri, k) for k, _ in ReplayInfo._fields_}

    def set_var_flags(self, flags, inst_var_off):
        """Solver::new_var(upol, dvar) records (B200SAT_VAR_* bits) of the resident batch; None clears them"""
        if flags is None:
            _check(lib().b200sat_engine_set_var_flags(self._e, None, None))
            return
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        o, op = _u32(inst_var_off)
        _check(lib().b200sat_engine_set_var_flags(self._e, f.ctypes.data_as(C.POINTER(C.c_uint8)), op))

    def set_clause_nvars(self, clause_nvars):
        """vars that existed when each clause of the resident batch was added (new_var interleaved with add_clause);
        None = every var created before the first clause.  Needs set_var_flags."""
        if clause_nvars is None:
            _check(lib().b200sat_engine_set_clause_nvars(self._e, None))
            return
        a, ap = _u32(clause_nvars)
        _check(lib().b200sat_engine_set_clause_nvars(self._e, ap))

    def set_assumptions(self, lits):
        a, ap = _u32(lits if lits is not None and len(lits) else np.zeros(1, np.uint32))
        _check(lib().b200sat_engine_set_assumptions(self._e, ap, 0 if lits is None else len(lits)))

    def resume(self, stream=None):
        """search again on the instances the last solve left parked (INTERRUPTED / NO_SOLVE)"""
        _check(lib().b200sat_engine_resume(self._e, stream))

    def ring_open_local_peer(self, slot, owner):
        """map the ring of `owner` (an Engine of this process on another GPU) through CUDA peer access"""
        _check(lib().b200sat_engine_ring_open_local_peer(self._e, slot, owner._e))

    def set_progress_rows(self, rows):
        """record the search-table progress rows (search.rs:289-301) of solves run with F_COUNT_TRAFFIC"""
        _check(lib().b200sat_engine_set_progress_rows(self._e, rows))

    def progress_rows(self, inst=0, cap=256):
        out = np.zeros((cap, 8), dtype=np.float64)
        n = C.c_uint32()
        _check(lib().b200sat_engine_download_progress(self._e, inst, out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(n)))
        return out[:n.value]

    def preprocess_instance(self, inst, n_clauses, n_lits, grow=0, cl_lim=20, max_outer=64, stream=None):
        """K4 as one preprocessor (subsumption + strengthening + BVE interleaved, one cooperative kernel) ->
        dict(clause_off, lits, estack, kept, kept_lits, eliminated, subsumed, strengthened, outer, bve_rounds, ss_rounds,
        launches, unsat, kernel_ms)"""
        NC, NL = 2 * n_clauses + 1024, 3 * n_lits + 4096
        off = np.zeros(NC + 2, np.uint32); li = np.zeros(NL + 2, np.uint32); es = np.zeros(NL + 2, np.uint32)
        cnt = np.zeros(16, np.uint32); ms = C.c_float()
        u32p = C.POINTER(C.c_uint32)
        _check(lib().b200sat_engine_preprocess_instance(self._e, inst, grow, cl_lim, max_outer, off.ctypes.data_as(u32p),
                                                        li.ctypes.data_as(u32p), es.ctypes.data_as(u32p), cnt.ctypes.data_as(u32p),
                                                        C.byref(ms), stream))
        kept, kl = int(cnt[0]), int(cnt[1])
        return dict(clause_off=off[:kept + 1].copy(), lits=li[:kl].copy(), estack=es[:int(cnt[8])].copy(), kept=kept, kept_lits=kl,
                    eliminated=int(cnt[2]), subsumed=int(cnt[3]), strengthened=int(cnt[4]), outer=int(cnt[5]), bve_rounds=int(cnt[6]),
                    ss_rounds=int(cnt[7]), launches=int(cnt[11]), unsat=bool(cnt[10]), kernel_ms=float(ms.value))

    def k4_extend_model(self, model, stream=None):
        """elim_clauses.rs:43-63 on the device over the stack of the last preprocess_instance; model: uint8[n_vars] in/out"""
        m = np.ascontiguousarray(model, dtype=np.uint8).copy()
        _check(lib().b200sat_engine_k4_extend_model(self._e, m.ctypes.data_as(C.POINTER(C.c_uint8)), len(m), stream))
        return m

    def set_rescue(self, on=True):
        """free-running mode: idle warps help instances that are still running (NOT the reference trace for those)"""
        _check(lib().b200sat_engine_set_rescue(self._e, 1 if on else 0))

    def set_import_cap(self, cap):
        _check(lib().b200sat_engine_set_import_cap(self._e, cap))

    def set_slice(self, slice_conflicts):
        """conflicts an instance runs before it yields to waiting ones (0 = never); trace-neutral"""
        _check(lib().b200sat_engine_set_slice(self._e, slice_conflicts))

    def set_budget(self, conflict_budget=-1, propagation_budget=-1):
        """Budget (budget.rs:5-34) of every instance of the following solve calls; < 0 = none"""
        _check(lib().b200sat_engine_set_budget(self._e, conflict_budget, propagation_budget))

    def interrupt(self, raised=True):
        """Budget::asynch_interrupt: may be called from another thread while solve() runs"""
        _check(lib().b200sat_engine_interrupt(self._e, 1 if raised else 0))

    @staticmethod
    def _batch(inst_clause_off, clause_off, lits):
        ioff, ioffp = _u32(inst_clause_off)
        off, offp = _u32(clause_off)
        li, lip = _u32(lits if len(lits) else np.zeros(1, np.uint32))
        b = BatchStruct(len(ioff) - 1, ioffp, offp, lip)
        return b, (ioff, off, li)

    def upload(self, inst_clause_off, clause_off, lits, stream=None):
        b, keep = self._batch(inst_clause_off, clause_off, lits)
        _check(lib().b200sat_engine_upload(self._e, C.byref(b), stream))

    def generate_ksat(self, seed_base, n_inst, n_vars, n_clauses, k=3, stream=None):
        _check(lib().b200sat_engine_generate_ksat(self._e, seed_base & 0xFFFFFFFFFFFFFFFF, n_inst, n_vars,
                                                  n_clauses, k, stream))

    def solve(self, settings=None, kind=CORE, flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT, stream=None):
        s = settings if settings is not None else default_settings()
        return _check(lib().b200sat_engine_solve(self._e, C.byref(s), kind, flags, stream), allow=(E_MEM,))

    def solve_async(self, settings=None, kind=CORE, flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT, stream=None):
        self._settings_keep = settings if settings is not None else default_settings()
        _check(lib().b200sat_engine_solve_async(self._e, C.byref(self._settings_keep), kind, flags, stream))

    def finish(self):
        return _check(lib().b200sat_engine_finish(self._e), allow=(E_MEM,))

    def solve_batch_async(self, inst_clause_off, clause_off, lits, results, models, settings=None, kind=CORE,
                          flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT, stream=None):
        """results: uint8/RESULT_DTYPE ndarray of n_inst records, models: (n_inst, stride) uint8 ndarray or None;
        all buffers should be page-locked and must stay alive until solve_batch_finish()."""
        s = settings if settings is not None else default_settings()
        b, keep = self._batch(inst_clause_off, clause_off, lits)
        self._keep = (b, keep, s, results, models)
        mp = models.ctypes.data_as(C.POINTER(C.c_uint8)) if models is not None else None
        stride = models.shape[1] if models is not None else 0
        _check(lib().b200sat_solve_batch_async(self._e, C.byref(b), C.byref(s), kind, flags, results.ctypes.data, mp,
                                               stride, stream))

    def solve_batch_finish(self):
        return _check(lib().b200sat_solve_batch_finish(self._e), allow=(E_MEM,))

    # ---- portfolio mode
    def ring_create(self, payload_words=1 << 20, n_rings=1, my_slot=0):
        _check(lib().b200sat_engine_ring_create(self._e, payload_words, n_rings, my_slot))

    def ring_reset(self, stream=None):
        _check(lib().b200sat_engine_ring_reset(self._e, stream))

    def ring_ipc_handle(self):
        buf = (C.c_uint8 * 64)()
        _check(lib().b200sat_engine_ring_ipc_handle(self._e, buf))
        return bytes(buf)

    def ring_open_peer(self, slot, handle):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(lib().b200sat_engine_ring_open_peer(self._e, slot, buf))

    def portfolio_solve(self, n_replicas, replica_base=0, share=PF_SHARE, settings=None, flags=0, stream=None,
                        model_stride=None):
        """Run n_replicas diversified replicas of instance 0 of the resident batch; returns (results, models)
        with one row per replica.  The winner is any replica whose verdict is not INTERRUPTED."""
        s = settings if settings is not None else default_settings()
        _check(lib().b200sat_portfolio_solve(self._e, C.byref(s), n_replicas, replica_base, share, flags, stream))
        res = np.zeros(n_replicas, dtype=RESULT_DTYPE)
        models = None
        mp = None
        if model_stride:
            models = np.full((n_replicas, model_stride), 2, dtype=np.uint8)
            mp = models.ctypes.data_as(C.POINTER(C.c_uint8))
        _check(lib().b200sat_engine_download_replicas(self._e, res.ctypes.data, mp, model_stride or 0, stream))
        return res, models

    def dims(self):
        n, nc, nl = C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(lib().b200sat_engine_batch_dims(self._e, C.byref(n), C.byref(nc), C.byref(nl)))
        return n.value, nc.value, nl.value

    def download(self, want_models=True, model_stride=None, stream=None):
        n, _, _ = self.dims()
        res = np.zeros(n, dtype=RESULT_DTYPE)
        models = None
        mp = None
        if want_models:
            if model_stride is None:
                raise ValueError("model_stride required")
            models = np.full((n, model_stride), 2, dtype=np.uint8)
            mp = models.ctypes.data_as(C.POINTER(C.c_uint8))
        _check(lib().b200sat_engine_download(self._e, res.ctypes.data, mp, model_stride or 0, stream))
        return res, models

    def verify_models(self, stream=None):
        """On-device check of every SAT model against the original CNF -> (verified uint8[n_inst], n_violated)."""
        n, _, _ = self.dims()
        ver = np.zeros(max(n, 1), dtype=np.uint8)
        nv = C.c_uint32()
        _check(lib().b200sat_engine_verify_models(self._e, ver.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nv), stream))
        return ver[:n], nv.value

    def subsume_instance(self, inst, n_clauses, n_lits, stream=None):
        """K4 multi-CTA backward subsumption of instance `inst` of the resident batch ->
        dict(removed uint8[n_clauses], clause_off, lits (compacted CSR), kept, kept_lits, tests, removed_count, kernel_ms)."""
        u32p = C.POINTER(C.c_uint32)
        removed = np.zeros(max(n_clauses, 1), np.uint8)
        ooff = np.zeros(n_clauses + 2, np.uint32)
        olits = np.zeros(max(n_lits, 1), np.uint32)
        counts = (C.c_uint32 * 4)()
        ms = C.c_float()
        _check(lib().b200sat_engine_subsume_instance(self._e, inst, removed.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                     ooff.ctypes.data_as(u32p), olits.ctypes.data_as(u32p), counts, C.byref(ms), stream))
        kept, kept_lits = counts[0], counts[1]
        return {"removed": removed[:n_clauses], "clause_off": ooff[:kept + 1], "lits": olits[:kept_lits], "kept": kept,
                "kept_lits": kept_lits, "tests": counts[2], "removed_count": counts[3], "kernel_ms": ms.value}

    def simplify_instance(self, inst, n_clauses, n_lits, strengthen=True, stream=None):
        """K4 subsumption (+ self-subsuming strengthening to fixpoint) of instance `inst` -> dict like subsume_instance
        plus rounds / strengthened / unsat."""
        u32p = C.POINTER(C.c_uint32)
        removed = np.zeros(max(n_clauses, 1), np.uint8)
        ooff = np.zeros(n_clauses + 2, np.uint32)
        olits = np.zeros(max(n_lits, 1), np.uint32)
        counts = (C.c_uint32 * 6)()
        ms = C.c_float()
        _check(lib().b200sat_engine_simplify_instance(self._e, inst, int(strengthen), removed.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                      ooff.ctypes.data_as(u32p), olits.ctypes.data_as(u32p), counts, C.byref(ms), stream))
        kept, kept_lits = counts[0], counts[1]
        return {"removed": removed[:n_clauses], "clause_off": ooff[:kept + 1], "lits": olits[:kept_lits], "kept": kept,
                "kept_lits": kept_lits, "rounds": counts[2], "strengthened": counts[3], "unsat": bool(counts[4]),
                "removed_count": counts[5], "kernel_ms": ms.value}

    def bve_instance(self, inst, n_clauses, n_lits, grow=0, cl_lim=20, max_rounds=512, stream=None):
        """K4 multi-CTA bounded variable elimination of instance `inst` -> dict(clause_off, lits, estack records, counts)."""
        u32p = C.POINTER(C.c_uint32)
        ooff = np.zeros(2 * n_clauses + 1030, np.uint32)
        olits = np.zeros(3 * n_lits + 4100, np.uint32)
        est = np.zeros(3 * n_lits + 4100, np.uint32)
        counts = (C.c_uint32 * 6)()
        ms = C.c_float()
        _check(lib().b200sat_engine_bve_instance(self._e, inst, grow, cl_lim, max_rounds, ooff.ctypes.data_as(u32p),
                                                 olits.ctypes.data_as(u32p), est.ctypes.data_as(u32p), counts, C.byref(ms), stream))
        kept, kept_lits, ew = counts[0], counts[1], counts[4]
        recs, w = [], 0
        while w < ew:                      # [var, unit lit, round, k, (len, lits...)*k]
            var, unit, rnd, k = (int(x) for x in est[w:w + 4])
            w += 4
            cls = []
            for _ in range(k):
                n = int(est[w]); cls.append([int(x) for x in est[w + 1:w + 1 + n]]); w += 1 + n
            recs.append((rnd, var, unit, cls))
        return {"clause_off": ooff[:kept + 1], "lits": olits[:kept_lits], "kept": kept, "kept_lits": kept_lits,
                "eliminated": counts[2], "rounds": counts[3], "records": recs, "kernel_ms": ms.value}

    def download_batch(self):
        n, nc, nl = self.dims()
        ioff = np.zeros(n + 1, np.uint32)
        coff = np.zeros(nc + 1, np.uint32)
        lits = np.zeros(max(nl, 1), np.uint32)
        u32p = C.POINTER(C.c_uint32)
        _check(lib().b200sat_engine_download_batch(self._e, ioff.ctypes.data_as(u32p), coff.ctypes.data_as(u32p),
                                                   lits.ctypes.data_as(u32p)))
        return ioff, coff, lits[:nl]

    def launch_info(self):
        li = LaunchInfo()
        _check(lib().b200sat_engine_launch_info(self._e, C.byref(li)))
        return {k: getattr(li, k) for k, _ in LaunchInfo._fields_}

    def solve_batch(self, inst_clause_off, clause_off, lits, settings=None, kind=CORE,
                    flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT, model_stride=0):
        """upload + solve + download with HOST buffers (b200sat_solve_batch)."""
        s = settings if settings is not None else default_settings()
        b, keep = self._batch(inst_clause_off, clause_off, lits)
        n = b.n_inst
        res = np.zeros(n, dtype=RESULT_DTYPE)
        models = None
        mp = None
        if model_stride:
            models = np.full((n, model_stride), 2, dtype=np.uint8)
            mp = models.ctypes.data_as(C.POINTER(C.c_uint8))
        _check(lib().b200sat_solve_batch(self._e, C.byref(b), C.byref(s), kind, flags, res.ctypes.data, mp,
                                         model_stride), allow=(E_MEM,))
        return res, models


def extend_model_k4(model, records):
    """Model extension over the K4 eliminated-clause records (elim_clauses.rs:43-63): later rounds first; per record the
    default unit, then every saved clause that is not satisfied flips the variable to its own literal."""
    def is_true(l):
        m = model[l >> 1]
        return m != 2 and (m == 1) == ((l & 1) == 0)
    for v in range(len(model)):     # free variables take `false` before the walk (the reference's search assigns every variable)
        if model[v] == 2:
            model[v] = 0
    for rnd, var, unit, cls in sorted(records, key=lambda r: -r[0]):
        model[var] = 0 if (unit & 1) else 1
        for c in cls:
            if not any(is_true(l) for l in c):
                model[c[0] >> 1] = 0 if (c[0] & 1) else 1
    return model


def diversify(base, replica):
    """Settings of global replica `replica` (b200sat_diversify)."""
    out = Settings()
    lib().b200sat_diversify(C.byref(base), replica, C.byref(out))
    return out


def shard_range(n_items, rank, world_size):
    """Contiguous, near-equal index range of `rank` (SURVEY.md 8e batch mode: no data-path collective)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


from . import dimacs  # noqa: E402,F401
