"""ctypes binding of the TEST-ONLY SIMT emulator harness (tests/emu/libemu.so).

Runs the device solver source on the CPU, one emulated warp per instance.
Used by the `not gpu` tests to check trace parity of the device algorithm
against the oracle without a GPU.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libemu.so")


class Settings(C.Structure):
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
        ("verdict", C.c_int32), ("status", C.c_int32), ("n_vars", C.c_uint32),
        ("n_clauses_ingest", C.c_uint32), ("n_clauses_pre", C.c_uint32), ("eliminated_vars", C.c_uint32),
        ("pre_ok", C.c_uint32), ("attempts", C.c_uint32),
        ("stats", C.c_uint64 * 8), ("traffic", C.c_uint64 * 8), ("trace_hash", C.c_uint64),
        ("exported", C.c_uint32), ("imported", C.c_uint32), ("progress", C.c_double),
    ]


def build(force=False):
    subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u32p = C.POINTER(C.c_uint32)
        L.emu_solve_batch.argtypes = [u32p, u32p, u32p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Result),
                                      C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, C.c_int]
        L.emu_solve_batch.restype = C.c_int
        L.emu_portfolio.argtypes = [u32p, u32p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64,
                                    C.POINTER(Result), C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32]
        L.emu_portfolio.restype = C.c_int
        L.emu_set_extras.argtypes = [C.POINTER(C.c_uint8), u32p, u32p, C.c_uint32]
        L.emu_set_clause_nvars.argtypes = [u32p]
        L.emu_bcp_replay.argtypes = [u32p, u32p, u32p, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64,
                                     C.POINTER(C.c_uint64)]
        L.emu_bcp_replay.restype = C.c_int
        _lib = L
    return _lib


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


def solve_batch(inst_clause_off, clause_off, lits, settings, kind=0, flags=1 | 4 | 8 | 16, tier_nv=-1,
                arena_words=0, wpool_entries=0, sched_seed=0, model_stride=0, slice_conflicts=0,
                conflict_budget=-1, propagation_budget=-1, resume_rounds=0, want_suspends=False,
                var_flags=None, inst_var_off=None, assumptions=None, clause_nvars=None):
    """settings: any ctypes struct with the b200sat_settings layout (e.g. oracle Settings).
    slice_conflicts > 0 parks and resumes every instance each time it used that many conflicts."""
    ioff, ioffp = _u32(inst_clause_off)
    off, offp = _u32(clause_off)
    li, lip = _u32(lits if len(lits) else np.zeros(1, np.uint32))
    n = len(ioff) - 1
    res = (Result * n)()
    models = None
    mp = None
    if model_stride:
        models = np.full((n, model_stride), 2, dtype=np.uint8)
        mp = models.ctypes.data_as(C.POINTER(C.c_uint8))
    vf = vo = asm = None
    if var_flags is not None or assumptions is not None:
        vfp = vop = asp = None
        if var_flags is not None:
            vf = np.ascontiguousarray(var_flags, dtype=np.uint8); vfp = vf.ctypes.data_as(C.POINTER(C.c_uint8))
            vo, vop = _u32(inst_var_off)
        if assumptions is not None and len(assumptions):
            asm, asp = _u32(assumptions)
        lib().emu_set_extras(vfp, vop, asp, 0 if asm is None else len(asm))
    cnv = None
    if clause_nvars is not None:
        cnv, cnvp = _u32(clause_nvars)
        lib().emu_set_clause_nvars(cnvp)
    rc = lib().emu_solve_batch(ioffp, offp, lip, n, C.addressof(settings), kind, flags, tier_nv,
                               arena_words, wpool_entries, sched_seed, res, mp, model_stride, slice_conflicts,
                               conflict_budget, propagation_budget, resume_rounds)
    lib().emu_set_extras(None, None, None, 0)
    lib().emu_set_clause_nvars(None)
    if rc < 0:
        raise RuntimeError("emulator: solver kind not compiled in")
    if want_suspends:
        return res, models, rc
    return res, models


def portfolio(clause_off, lits, settings, n_replicas, share=1, flags=8, sched_seed=0, model_stride=0, import_cap=0):
    """Replicas of one instance run one after another against a shared ring (see emu_harness.cpp)."""
    off, offp = _u32(clause_off)
    li, lip = _u32(lits)
    res = (Result * n_replicas)()
    models = np.full((n_replicas, max(model_stride, 1)), 2, dtype=np.uint8)
    lib().emu_portfolio(offp, lip, len(off) - 1, C.addressof(settings), n_replicas, share, flags, sched_seed, res,
                        models.ctypes.data_as(C.POINTER(C.c_uint8)), model_stride, import_cap)
    return res, models


def bcp_replay(inst_clause_off, clause_off, lits, settings, tier_nv=-1, trace_cap=8192, sched_seed=0):
    """K2 under the emulator: rows of (recorded words, replay ok, recorded propagations, replayed propagations,
    BCP traffic counters equal) per instance."""
    ioff, ioffp = _u32(inst_clause_off)
    off, offp = _u32(clause_off)
    li, lip = _u32(lits)
    n = len(ioff) - 1
    out = np.zeros((n, 5), dtype=np.uint64)
    lib().emu_bcp_replay(ioffp, offp, lip, n, C.addressof(settings), tier_nv, trace_cap, sched_seed,
                         out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out
