"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.  See oracle/oracle.h for the parity-pin status
("parity unpinned" for the trace counters).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

CORE, SIMP = 0, 1
UNSAT, SAT, INTERRUPTED = 0, 1, 2
F_PREPROCESS, F_NO_SOLVE, F_ZERO_STATS_ON_PRE_UNSAT = 1, 2, 4
F_NO_PRE = 32
STAT_NAMES = ("solves", "restarts", "decisions", "rnd_decisions", "conflicts",
              "propagations", "tot_literals", "del_literals")
TRAFFIC_NAMES = ("visited", "kept", "deref", "tail", "moves", "implied", "an_words", "learnt_words")


class Settings(C.Structure):
    """Mirror of oracle_settings (identical layout to b200sat_settings)."""
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


class Result(C.Structure):
    _fields_ = [
        ("verdict", C.c_int32), ("n_vars", C.c_uint32), ("n_clauses_ingest", C.c_uint32),
        ("n_clauses_pre", C.c_uint32), ("eliminated_vars", C.c_uint32), ("pre_ok", C.c_uint32),
        ("stats", C.c_uint64 * 8), ("traffic", C.c_uint64 * 8), ("trace_hash", C.c_uint64),
        ("seconds", C.c_double), ("progress", C.c_double),
    ]

    def stats_dict(self):
        return dict(zip(STAT_NAMES, [int(x) for x in self.stats]))

    def traffic_dict(self):
        return dict(zip(TRAFFIC_NAMES, [int(x) for x in self.traffic]))


def traffic_bytes(t):
    """algorithmic bytes from the eight traffic counters."""
    visited, kept, deref, tail, moves, implied, an_words, learnt_words = [int(x) for x in t]
    return (8 * visited + 8 * kept + 4 * deref + 4 * (2 * deref + tail) + 16 * moves + 12 * implied
            + 4 * an_words + 4 * learnt_words)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None
_NATIVE_PATH = os.path.join(_HERE, "_native", "liboracle_native.so")
_use_native = False
_native_flags = ""


def _cpu_signature():
    try:
        model, flags = "", ""
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name") and not model:
                model = ln.split(":", 1)[1].strip()
            elif ln.startswith("flags") and not flags:
                flags = " ".join(sorted(ln.split(":", 1)[1].split()))
            if model and flags:
                break
        import hashlib
        return model + " " + hashlib.sha1(flags.encode()).hexdigest()[:16]
    except Exception:
        return "unknown"


def build_native():
    """CPU-baseline build for THIS machine: -O3 -march=native -flto (mirrors the reference's opt-level=3, lto=true,
    Cargo.toml:28-31).  Kept apart from the portable liboracle.so that travels with the repo."""
    src = os.path.join(_HERE, "oracle.cpp")
    # a -march=native library is only valid on the CPU it was built on: the build records the host's CPU signature and
    # a copy that travelled to another machine is rebuilt there
    sig = _cpu_signature()
    sig_path = _NATIVE_PATH + ".cpu"
    if (os.path.exists(_NATIVE_PATH) and os.path.getmtime(_NATIVE_PATH) >= os.path.getmtime(src)
            and os.path.exists(sig_path) and open(sig_path).read() == sig):
        return _NATIVE_PATH
    os.makedirs(os.path.dirname(_NATIVE_PATH), exist_ok=True)
    global _native_flags
    for lto in (["-flto"], []):   # the oracle is ONE translation unit, so -flto changes nothing; some images lack lto-wrapper
        cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O3", "-march=native"] + lto + [
            "-ffp-contract=off", "-fPIC", "-shared", "-o", _NATIVE_PATH, src, "-lz", "-lpthread"]
        if subprocess.run(cmd, capture_output=True).returncode == 0:
            _native_flags = "-O3 -march=native" + (" -flto" if lto else " (single translation unit; -flto unavailable in this image)")
            open(_NATIVE_PATH + ".flags", "w").write(_native_flags)
            open(sig_path, "w").write(sig)
            return _NATIVE_PATH
    raise RuntimeError("native oracle build failed")


def use_native():
    """Switch this process to the -march=native -flto build (compiled here on first use).  Must be called before
    the first oracle call; raises when the compiler is missing or fails (callers fall back to the portable build)."""
    global _use_native, _lib
    if _lib is not None and not _use_native:
        raise RuntimeError("oracle library already loaded")
    build_native()
    _use_native = True


def build_info():
    if _use_native:
        try:
            fl = open(_NATIVE_PATH + ".flags").read()
        except Exception:
            fl = "-O3 -march=native"
        return "g++ %s -ffp-contract=off, built on this host" % fl
    return "g++ -O3 -march=x86-64-v2 -ffp-contract=off (portable build)"


def lib():
    global _lib
    if _lib is None:
        if not _use_native and not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_NATIVE_PATH if _use_native else _LIB_PATH)
        u32p = C.POINTER(C.c_uint32)
        L.oracle_default_settings.argtypes = [C.POINTER(Settings)]
        L.oracle_solve_csr.argtypes = [u32p, u32p, C.c_uint32, C.POINTER(Settings), C.c_int, C.c_int,
                                       C.POINTER(Result), C.POINTER(C.c_uint8), C.c_uint32]
        L.oracle_solve_csr.restype = C.c_int
        L.oracle_solve_batch_csr.argtypes = [u32p, u32p, u32p, C.c_uint32, C.POINTER(Settings), C.c_int,
                                             C.c_int, C.c_int, C.POINTER(Result), C.POINTER(C.c_uint8),
                                             C.c_uint32]
        L.oracle_solve_batch_csr.restype = C.c_double
        L.oracle_parse_dimacs.argtypes = [C.c_char_p, C.c_int, C.POINTER(u32p), C.POINTER(u32p),
                                          u32p, u32p, u32p, C.c_char_p, C.c_size_t]
        L.oracle_parse_dimacs.restype = C.c_int
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_gen_ksat.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
        L.oracle_check_model.argtypes = [u32p, u32p, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32]
        L.oracle_check_model.restype = C.c_int
        L.oracle_set_trace.argtypes = [C.POINTER(C.c_uint64), C.c_uint32]
        L.oracle_set_budget.argtypes = [C.c_int64, C.c_int64, C.c_int]
        L.oracle_record_progress.argtypes = [C.c_int]
        L.oracle_get_progress.argtypes = [C.POINTER(C.c_double), C.c_uint32]
        L.oracle_get_progress.restype = C.c_uint32
        L.oracle_set_var_flags.argtypes = [C.POINTER(C.c_uint8), C.c_uint32]
        L.oracle_set_clause_nvars.argtypes = [C.POINTER(C.c_uint32), C.c_uint32]
        L.oracle_set_assumptions.argtypes = [C.POINTER(C.c_uint32), C.c_uint32]
        _lib = L
    return _lib


def default_settings():
    s = Settings()
    lib().oracle_default_settings(C.byref(s))
    return s


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


def solve_csr(clause_off, lits, kind=SIMP, flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT,
              settings=None, want_model=True, trace_cap=0):
    """Returns (Result, model ndarray(uint8) or None[, trace ndarray])."""
    s = settings if settings is not None else default_settings()
    off, offp = _u32(clause_off)
    li, lip = _u32(lits if len(lits) else np.zeros(1, np.uint32))
    n_clauses = len(off) - 1
    cap = int(li.max() >> 1) + 1 if len(lits) else 1
    model = np.full(cap, 2, dtype=np.uint8)
    res = Result()
    trace = None
    if trace_cap:
        trace = np.zeros((trace_cap, 3), dtype=np.uint64)
        lib().oracle_set_trace(trace.ctypes.data_as(C.POINTER(C.c_uint64)), trace_cap)
    lib().oracle_solve_csr(offp, lip, n_clauses, C.byref(s), kind, flags, C.byref(res),
                           model.ctypes.data_as(C.POINTER(C.c_uint8)), cap)
    if trace_cap:
        lib().oracle_set_trace(None, 0)
    m = model[:res.n_vars] if (want_model and res.verdict == SAT) else None
    return (res, m, trace) if trace_cap else (res, m)


def solve_batch(inst_clause_off, clause_off, lits, kind=SIMP,
                flags=F_PREPROCESS | F_ZERO_STATS_ON_PRE_UNSAT, settings=None, threads=1,
                model_stride=0):
    s = settings if settings is not None else default_settings()
    ioff, ioffp = _u32(inst_clause_off)
    off, offp = _u32(clause_off)
    li, lip = _u32(lits)
    n = len(ioff) - 1
    results = (Result * n)()
    models = None
    mp = None
    if model_stride:
        models = np.full((n, model_stride), 2, dtype=np.uint8)
        mp = models.ctypes.data_as(C.POINTER(C.c_uint8))
    secs = lib().oracle_solve_batch_csr(ioffp, offp, lip, n, C.byref(s), kind, flags, threads,
                                        results, mp, model_stride)
    return results, models, secs


def solve_csr_with_progress(clause_off, lits, kind=SIMP, settings=None, cap=256):
    """solve_csr + the rows of the search table (search.rs:289-301) as an (n, 8) float64 array"""
    lib().oracle_record_progress(1)
    try:
        r, m = solve_csr(clause_off, lits, kind=kind, settings=settings)
        out = np.zeros((cap, 8), dtype=np.float64)
        n = lib().oracle_get_progress(out.ctypes.data_as(C.POINTER(C.c_double)), cap)
    finally:
        lib().oracle_record_progress(0)
    return r, m, out[:n]


def set_budget(conflict_budget=-1, propagation_budget=-1, resume_rounds=0):
    """Budget of the following solve calls (process-wide); set_budget() resets it."""
    lib().oracle_set_budget(conflict_budget, propagation_budget, resume_rounds)


def set_var_flags(flags=None):
    """new_var(upol, dvar) records of the following solves (bit0 upol set, bit1 upol value, bit2 not decision)"""
    if flags is None or len(flags) == 0:
        lib().oracle_set_var_flags(None, 0)
        return
    f = np.ascontiguousarray(flags, dtype=np.uint8)
    lib().oracle_set_var_flags(f.ctypes.data_as(C.POINTER(C.c_uint8)), len(f))


def set_clause_nvars(clause_nvars=None):
    """with var flags: vars that existed when each clause was added (new_var interleaved with add_clause); None clears"""
    if clause_nvars is None or len(clause_nvars) == 0:
        lib().oracle_set_clause_nvars(None, 0)
        return
    a = np.ascontiguousarray(clause_nvars, dtype=np.uint32)
    lib().oracle_set_clause_nvars(a.ctypes.data_as(C.POINTER(C.c_uint32)), len(a))


def set_assumptions(lits=None):
    if lits is None or len(lits) == 0:
        lib().oracle_set_assumptions(None, 0)
        return
    a = np.ascontiguousarray(lits, dtype=np.uint32)
    lib().oracle_set_assumptions(a.ctypes.data_as(C.POINTER(C.c_uint32)), len(a))


def parse_dimacs(path, strict=False):
    """-> (clause_off uint32[m+1], lits uint32[L], hdr_vars, hdr_clauses)."""
    u32p = C.POINTER(C.c_uint32)
    offp, litp = u32p(), u32p()
    n, hv, hc = C.c_uint32(), C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(512)
    rc = lib().oracle_parse_dimacs(os.fsencode(path), int(strict), C.byref(offp), C.byref(litp),
                                   C.byref(n), C.byref(hv), C.byref(hc), err, 512)
    if rc != 0:
        raise IOError(err.value.decode())
    off = np.ctypeslib.as_array(offp, shape=(n.value + 1,)).copy()
    nl = int(off[-1])
    lits = np.ctypeslib.as_array(litp, shape=(max(nl, 1),))[:nl].copy()
    lib().oracle_free(offp)
    lib().oracle_free(litp)
    return off, lits, hv.value, hc.value


def gen_ksat(seed, n, m, k=3):
    lits = np.zeros(m * k, dtype=np.uint32)
    lib().oracle_gen_ksat(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), n, m, k,
                          lits.ctypes.data_as(C.POINTER(C.c_uint32)))
    return lits


def check_model(clause_off, lits, model):
    off, offp = _u32(clause_off)
    li, lip = _u32(lits)
    m = np.ascontiguousarray(model, dtype=np.uint8)
    return bool(lib().oracle_check_model(offp, lip, len(off) - 1,
                                         m.ctypes.data_as(C.POINTER(C.c_uint8)), len(m)))
