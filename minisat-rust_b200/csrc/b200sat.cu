/*
 * b200sat.cu -- kernels + C-ABI host of the B200-native batched CDCL engine.
 *
 * Kernels (sm_100a):
 *   cdcl_warp_smem<NV,COUNT>  persistent, one warp = one instance, var state in shared memory
 *   cdcl_warp_glob<COUNT>     same solver, var state in the HBM slab (large instances)
 *   gen_random_ksat           synthetic uniform k-SAT batch (SURVEY.md 8d generator)
 *   collect_failed            builds the retry work list of instances that outgrew their slab
 *
 * The host side mirrors `trait Solver` (reference src/sat.rs:28-36) and adds
 * the batch engine.  There is no CPU solving path in this file: without a CUDA
 * device every solving call returns B200SAT_E_NO_DEVICE.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <string>
#include <vector>

#include "cdcl_plan.h"

using namespace b200sat;

/* the cdcl_warp kernels are instantiated in kern_*.cu (one translation unit per COUNT x SIMP combination) */
#include "cdcl_kernels.cuh"

/* SplitMix64 generator of SURVEY.md 8d: bit-identical to the host generator the CPU baseline uses */
__global__ void gen_random_ksat(u64 seed_base, u32 n_inst, u32 n, u32 m, u32 k, u32 *inst_clause_off,
                                u32 *clause_off, u32 *lits) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_inst) return;
    inst_clause_off[i] = i * m;
    if (i == n_inst) { clause_off[(size_t)n_inst * m] = n_inst * m * k; return; }
    u64 state = seed_base + i;
    for (u32 c = 0; c < m; c++) {
        size_t cbase = ((size_t)i * m + c);
        clause_off[cbase] = (u32)(cbase * k);
        u32 *cl = lits + cbase * k;
        for (u32 j = 0; j < k;) {
            state += 0x9E3779B97F4A7C15ull;
            u64 z = state;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            u64 r = z ^ (z >> 31);
            u32 v = (u32)(((r >> 32) * (u64)n) >> 32);
            bool dup = false;
            for (u32 t = 0; t < j; t++) if ((cl[t] >> 1) == v) dup = true;
            if (dup) continue;
            cl[j++] = (v << 1) | (u32)(r & 1);
        }
    }
}

/* batch shape (largest instance) computed on the device right after the H2D copy: out = {max var id,
 * max clause length, max clauses per instance, max literals per instance, monotonicity violations} */
__global__ void batch_shape(const u32 *ioff, const u32 *coff, const u32 *lits, u32 n_inst, u64 nc, u64 nl, u32 *out) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nt = (u64)gridDim.x * blockDim.x;
    u32 mv = 0, ml = 0, mc = 0, mL = 0, bad = 0;
    for (u64 k = tid; k < nl; k += nt) { u32 v = lits[k] >> 1; mv = v > mv ? v : mv; }
    for (u64 c = tid; c < nc; c += nt) {
        u32 a = coff[c], b = coff[c + 1];
        if (b < a || b > nl) bad = 1; else { u32 l = b - a; ml = l > ml ? l : ml; }
    }
    for (u64 i = tid; i < n_inst; i += nt) {
        u32 a = ioff[i], b = ioff[i + 1];
        if (b < a || b > nc) bad = 1;
        else {
            u32 c = b - a; mc = c > mc ? c : mc;
            u32 la = coff[a], lb = coff[b];
            if (lb >= la) { u32 l = lb - la; mL = l > mL ? l : mL; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        u32 t;
        t = __shfl_xor_sync(0xFFFFFFFFu, mv, o); mv = t > mv ? t : mv;
        t = __shfl_xor_sync(0xFFFFFFFFu, ml, o); ml = t > ml ? t : ml;
        t = __shfl_xor_sync(0xFFFFFFFFu, mc, o); mc = t > mc ? t : mc;
        t = __shfl_xor_sync(0xFFFFFFFFu, mL, o); mL = t > mL ? t : mL;
        bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, mv); atomicMax(out + 1, ml); atomicMax(out + 2, mc); atomicMax(out + 3, mL);
        if (bad) atomicOr(out + 4, 1u);
    }
}

/* K5: model verification for a whole batch -- one warp per instance walks the ORIGINAL
 * CNF and checks that every clause holds a literal that is true under the instance's model
 * (dimacs.rs:90-128 validate_model, formula/util.rs:24-32).  verified[i]: 1 ok, 0 violated, 2 not SAT. */
__global__ void verify_models(const u32 *ioff, const u32 *coff, const u32 *lits, const b200sat_result *results,
                              const u8 *models, u32 model_stride, u32 n_inst, u8 *verified, u32 *n_violated) {
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_inst) return;
    if (results[warp].status != B200SAT_OK || results[warp].verdict != B200SAT_SAT) { if (lane == 0) verified[warp] = 2; return; }
    const u8 *m = models + (size_t)warp * model_stride;
    const u32 cb = ioff[warp], ce = ioff[warp + 1];
    bool bad = false;
    for (u32 c = cb + lane; c < ce; c += 32) {
        bool sat = false;
        for (u32 k = coff[c]; k < coff[c + 1]; k++) {
            const u32 l = lits[k], v = l >> 1;
            const u8 val = v < model_stride ? m[v] : (u8)2;
            if (val != 2 && (val == 1) == ((l & 1) == 0)) { sat = true; break; }
        }
        if (!sat) bad = true;
    }
    const u32 any = __ballot_sync(0xFFFFFFFFu, bad);
    if (lane == 0) { verified[warp] = any ? 0 : 1; if (any) atomicAdd(n_violated, 1u); }
}

__global__ void collect_failed(const b200sat_result *results, const u32 *work, u32 n_work, bool portfolio, u32 *out_list, u32 *out_count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    u32 inst = (work && !portfolio) ? work[i] : i;
    if (results[inst].status < 0) out_list[atomicAdd(out_count, 1u)] = inst; /* > 0: cancelled portfolio replica */
}

/* ------------------------------------------------------------------ host: errors */
static thread_local std::string g_err;
static int set_err(int code, const std::string &m) { g_err = m; return code; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return set_err(B200SAT_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

extern "C" const char *b200sat_last_error(void) { return g_err.c_str(); }

extern "C" void b200sat_default_settings(b200sat_settings *s) { /* SURVEY.md Appendix B */
    memset(s, 0, sizeof *s);
    s->var_decay = 0.95; s->clause_decay = 0.999; s->random_seed = 91648253.0; s->random_var_freq = 0.0;
    s->restart_first = 100.0; s->restart_inc = 2.0; s->size_factor = 1.0 / 3.0; s->size_inc = 1.1;
    s->size_adjust_inc = 1.5; s->garbage_frac = 0.20; s->simp_garbage_frac = 0.5; s->grow = 0;
    s->size_adjust_start_confl = 100; s->min_learnts_lim = 0; s->clause_lim = 20; s->subsumption_lim = 1000;
    s->ccmin_mode = 2; s->phase_saving = 2; s->luby_restart = 1; s->rnd_pol = 0; s->rnd_init_act = 0;
    s->use_rcheck = 0; s->use_asymm = 0; s->use_elim = 1; s->extend_model = 1; s->remove_satisfied = 1;
}

extern "C" int b200sat_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return set_err(B200SAT_E_NO_DEVICE, "no CUDA device: the engine has no CPU path"); }
    return n;
}

/* ------------------------------------------------------------------ host: engine */
struct b200sat_engine {
    int device = 0;
    int num_sms = 0;
    size_t smem_optin = 0;
    /* resident batch */
    u32 n_inst = 0;
    u64 n_clauses = 0, n_lits = 0;
    u32 *d_ioff = nullptr, *d_coff = nullptr, *d_lits = nullptr;
    size_t cap_ioff = 0, cap_coff = 0, cap_lits = 0;
    BatchShape shape = {0, 0, 0, 0};
    /* outputs */
    b200sat_result *d_results = nullptr;
    size_t cap_results = 0;
    u8 *d_models = nullptr;
    size_t cap_models = 0;
    u8 *d_verified = nullptr;
    size_t cap_verified = 0;
    u32 *d_pp = nullptr;          /* scratch of the K4 preprocessing kernels */
    u32 *k4_estack = nullptr;     /* eliminated-clause stack of the last preprocess_instance (inside d_pp) */
    u32 k4_estack_words = 0, k4_nvars = 0;
    size_t cap_pp = 0;
    u32 model_stride = 0;
    /* work */
    u8 *d_slabs = nullptr;
    size_t cap_slabs = 0;
    u32 *d_queue = nullptr;      /* [0] queue counter, [1] failed counter, [32..40) work-ring control words */
    unsigned long long *d_ring = nullptr; /* device work ring of the search kernel */
    size_t cap_ring = 0;
    cudaEvent_t ev2 = nullptr, ev3 = nullptr; /* around the search kernel alone */
    u32 slice_conflicts = 20000;  /* conflicts an instance may run before it yields to waiting ones */
    long long conflict_budget = -1, propagation_budget = -1;
    int *h_interrupt = nullptr;   /* pinned staging of the async interrupt word (Budget::asynch_interrupt) */
    int *d_interrupt = nullptr;   /* the word the kernels poll (device memory, d_queue + 48)               */
    cudaStream_t side_stream = nullptr; /* carries the interrupt word to the device while a solve is running */
    bool rescue = false;          /* free-running mode: idle warps help instances that are still running */
    u8 *d_pristine = nullptr;     /* copy of the slabs as the ingest kernel left them (rescue mode) */
    size_t cap_pristine = 0;
    u32 *d_run_table = nullptr, *d_rrings = nullptr;
    size_t cap_run_table = 0, cap_rrings = 0;
    u8 *d_parked = nullptr;       /* [item] resumable state resident in its slab */
    size_t cap_parked = 0;
    u8 *d_var_flags = nullptr;    /* Solver::new_var(upol, dvar) records of the resident batch (optional) */
    size_t cap_var_flags = 0;
    u32 *d_var_off = nullptr;
    size_t cap_var_off = 0;
    bool have_var_flags = false;
    u32 *d_clause_nvars = nullptr; /* vars that existed when each clause was added (optional, with the var flags) */
    size_t cap_clause_nvars = 0;
    bool have_clause_nvars = false;
    u32 *d_assumps = nullptr;     /* assumptions of the following solve calls */
    size_t cap_assumps = 0;
    u32 n_assumps = 0;
    double *d_progress = nullptr; /* CLI search-table rows of the instrumented kernel */
    size_t cap_progress = 0;
    u32 progress_rows = 0;
    u32 *d_trace = nullptr;       /* K2: per-instance event record of the search prefix */
    size_t cap_trace = 0;
    u32 trace_cap = 0;            /* words per instance; 0 = recording off */
    int replay_mode = 0;          /* 1: launch_attempt runs ingest + bcp_replay (count kernel), 2: same, plain kernel */
    u32 resident_items = 0;       /* items whose state is parked in d_slabs (chunk 0 of the last solve) */
    Layout resident_L;            /* layout the parked states were written with */
    u32 resident_attempt = 0;
    int resident_tier_nv = 0;
    u32 *d_work[2] = {nullptr, nullptr};
    size_t cap_work = 0, cap_work1 = 0;
    u32 *h_pinned = nullptr;     /* small pinned staging for counters */
    void *h_stage = nullptr;     /* pinned staging for uploads */
    size_t cap_stage = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    /* portfolio: clause-exchange rings (slot my_ring is ours, the others are peer mappings) */
    u32 *rings[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ring_is_peer[8] = {false, false, false, false, false, false, false, false};
    bool ring_is_local_peer[8] = {false, false, false, false, false, false, false, false};
    u32 n_rings = 0, my_ring = 0, ring_words = 0;
    b200sat_settings *d_replica_settings = nullptr;
    size_t cap_replica_settings = 0;
    u32 n_replicas = 0;
    /* in-flight asynchronous solve */
    struct Pending {
        bool active = false, retried = false, searched = false, resumed = false;
        float search_ms = 0.f;
        b200sat_settings st;
        int kind = 0, flags = 0, tier_nv = 0;
        cudaStream_t stream = nullptr;
        u32 attempt = 0, n_work = 0;
        const u32 *work = nullptr;
        u32 *next_list = nullptr;
        float total_ms = 0.f;
        b200sat_result *out_results = nullptr;   /* batch-async destination buffers */
        uint8_t *out_models = nullptr;
        uint32_t out_stride = 0;
        bool batch = false;
        u32 portfolio = 0, replica_base = 0, share = 0;
    } pend;
    /* tunables */
    int cfg_wpb = 0, cfg_bps = 0;
    u64 cfg_arena = 0, cfg_wpool = 0;
    size_t cfg_slab_budget = 0;   /* bytes; 0 = 60% of the free HBM */
    u32 import_cap = 64;          /* portfolio: imported clauses per replica and restart boundary */
    b200sat_launch_info info;
};

template <class T> static int ensure(T *&p, size_t &cap, size_t need_elems) {
    if (need_elems <= cap && p) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t n = need_elems + need_elems / 8 + 64;
    cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
    if (e != cudaSuccess) return set_err(B200SAT_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = n;
    return 0;
}

extern "C" b200sat_engine *b200sat_engine_create(int device) {
    if (b200sat_device_count() <= 0) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { set_err(B200SAT_E_CUDA, "cudaSetDevice failed"); return nullptr; }
    b200sat_engine *e = new b200sat_engine();
    e->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; set_err(B200SAT_E_CUDA, "cudaGetDeviceProperties failed"); return nullptr; }
    e->num_sms = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    memset(&e->info, 0, sizeof e->info);
    memset(&e->resident_L, 0, sizeof e->resident_L);
    if (cudaMalloc((void **)&e->d_queue, 256) != cudaSuccess || cudaMallocHost((void **)&e->h_pinned, 1024) != cudaSuccess ||
        cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreate(&e->ev2) != cudaSuccess || cudaEventCreate(&e->ev3) != cudaSuccess ||
        cudaMallocHost((void **)&e->h_interrupt, 64) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMemset(e->d_queue, 0, 256) != cudaSuccess) {
        set_err(B200SAT_E_CUDA, "engine allocation failed");
        delete e;
        return nullptr;
    }
    *e->h_interrupt = 0;
    e->d_interrupt = (int *)(e->d_queue + 48);
    return e;
}

extern "C" void b200sat_engine_destroy(b200sat_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_ioff); cudaFree(e->d_coff); cudaFree(e->d_lits); cudaFree(e->d_results); cudaFree(e->d_models);
    cudaFree(e->d_slabs); cudaFree(e->d_queue); cudaFree(e->d_work[0]); cudaFree(e->d_work[1]);
    cudaFree(e->d_replica_settings); cudaFree(e->d_verified); cudaFree(e->d_pp);
    for (u32 r = 0; r < 8; r++)
        if (e->rings[r]) { if (e->ring_is_local_peer[r]) {} else if (e->ring_is_peer[r]) cudaIpcCloseMemHandle(e->rings[r]); else cudaFree(e->rings[r]); }
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev2) cudaEventDestroy(e->ev2);
    if (e->ev3) cudaEventDestroy(e->ev3);
    if (e->h_interrupt) cudaFreeHost(e->h_interrupt);
    if (e->side_stream) cudaStreamDestroy(e->side_stream);
    cudaFree(e->d_ring);
    cudaFree(e->d_trace); cudaFree(e->d_progress); cudaFree(e->d_pristine); cudaFree(e->d_run_table); cudaFree(e->d_rrings); cudaFree(e->d_parked); cudaFree(e->d_var_flags); cudaFree(e->d_clause_nvars); cudaFree(e->d_var_off); cudaFree(e->d_assumps);
    delete e;
}

extern "C" int b200sat_engine_configure(b200sat_engine *e, int warps_per_block, int blocks_per_sm, uint64_t arena_words,
                                        uint64_t wpool_entries) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->cfg_wpb = warps_per_block; e->cfg_bps = blocks_per_sm; e->cfg_arena = arena_words; e->cfg_wpool = wpool_entries;
    return 0;
}

static int alloc_outputs(b200sat_engine *e);
/* Solver::new_var(upol, dvar) records for the resident batch (sat.rs:31; decision_heuristic.rs:70-102,181-192) */
/* With var flags set: clause_nvars[c] = how many vars the caller had created when it added clause c of the resident
 * batch (global clause index).  SimpSolver's elimination order depends on it (elim_queue.rs:26-76).  null clears. */
extern "C" int b200sat_engine_set_clause_nvars(b200sat_engine *e, const uint32_t *clause_nvars) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    if (!clause_nvars) { e->have_clause_nvars = false; return 0; }
    if (e->n_clauses == 0) return 0;
    CK(cudaSetDevice(e->device));
    int rc;
    if ((rc = ensure(e->d_clause_nvars, e->cap_clause_nvars, (size_t)e->n_clauses))) return rc;
    CK(cudaMemcpy(e->d_clause_nvars, clause_nvars, (size_t)e->n_clauses * 4, cudaMemcpyHostToDevice));
    e->have_clause_nvars = true;
    return 0;
}
extern "C" int b200sat_engine_set_var_flags(b200sat_engine *e, const uint8_t *flags, const uint32_t *inst_var_off) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    if (!flags || !inst_var_off) { e->have_var_flags = false; return 0; }
    if (e->n_inst == 0) return 0;
    const u32 total = inst_var_off[e->n_inst];
    u32 maxv = 0;
    for (u32 i = 0; i < e->n_inst; i++) {
        if (inst_var_off[i + 1] < inst_var_off[i]) return set_err(B200SAT_E_ARG, "inst_var_off is not monotone");
        maxv = std::max(maxv, inst_var_off[i + 1] - inst_var_off[i]);
    }
    int rc;
    if ((rc = ensure(e->d_var_flags, e->cap_var_flags, (size_t)total + 1))) return rc;
    if ((rc = ensure(e->d_var_off, e->cap_var_off, (size_t)e->n_inst + 1))) return rc;
    CK(cudaMemcpy(e->d_var_flags, flags, total, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_var_off, inst_var_off, ((size_t)e->n_inst + 1) * 4, cudaMemcpyHostToDevice));
    if (maxv > e->shape.max_vars) { /* vars created but never used in a clause still exist (n_vars, elimination, model) */
        e->shape.max_vars = maxv;
        if ((rc = alloc_outputs(e))) return rc;
    }
    e->have_var_flags = true;
    return 0;
}
/* assumptions of the following solve calls (Solver::solve_limited, sat.rs:34), applied to every instance */
extern "C" int b200sat_engine_set_assumptions(b200sat_engine *e, const uint32_t *lits, uint32_t n) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    e->n_assumps = 0;
    if (!lits || n == 0) return 0;
    int rc;
    if ((rc = ensure(e->d_assumps, e->cap_assumps, (size_t)n))) return rc;
    CK(cudaMemcpy(e->d_assumps, lits, (size_t)n * 4, cudaMemcpyHostToDevice));
    e->n_assumps = n;
    return 0;
}

/* Rows of the reference's search table (search.rs:289-301: one per learnt-limit bump) are recorded for every instance
 * of the following solves that run with B200SAT_F_COUNT_TRAFFIC (the instrumented kernels); rows = 0 switches it off. */
extern "C" int b200sat_engine_set_progress_rows(b200sat_engine *e, uint32_t rows) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->progress_rows = rows;
    return 0;
}
/* out: rows x 8 doubles {conflicts, vars, clauses, literals, learnt limit, learnts, literals per learnt, progress %} */
extern "C" int b200sat_engine_download_progress(b200sat_engine *e, uint32_t inst, double *out, uint32_t cap_rows, uint32_t *n_rows) {
    if (!e || !out || !n_rows || inst >= e->n_inst) return set_err(B200SAT_E_ARG, "bad argument");
    *n_rows = 0;
    if (!e->progress_rows || !e->d_progress) return 0;
    CK(cudaSetDevice(e->device));
    const size_t stride = 1 + 8 * (size_t)e->progress_rows;
    std::vector<double> h(stride);
    CK(cudaMemcpy(h.data(), e->d_progress + inst * stride, stride * sizeof(double), cudaMemcpyDeviceToHost));
    u32 n = (u32)h[0];
    if (n > e->progress_rows) n = e->progress_rows;
    if (n > cap_rows) n = cap_rows;
    memcpy(out, h.data() + 1, (size_t)n * 8 * sizeof(double));
    *n_rows = n;
    return 0;
}
/* Free-running mode with straggler rescue (default off = trace mode): once the work ring is empty, idle warps clone the
 * post-ingest state of instances that are still running and search them as diversified helper replicas that exchange
 * units / binaries with each other; the first replica to finish decides the instance.  Verdicts stay exact and every
 * SAT model is one of the original formula; the counters of an instance a helper decided (result.attempts & 0x100) are
 * the helper's, not the reference trace's.  Needs the whole batch resident at once (2 x the slab memory); not
 * combined with budgets, traffic counting or the portfolio call. */
extern "C" int b200sat_engine_set_rescue(b200sat_engine *e, int on) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->rescue = on != 0;
    return 0;
}
/* portfolio: clauses one replica imports per restart boundary (0 = no cap); default 64 */
extern "C" int b200sat_engine_set_import_cap(b200sat_engine *e, uint32_t cap) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->import_cap = cap;
    return 0;
}
extern "C" int b200sat_engine_set_slice(b200sat_engine *e, uint32_t slice_conflicts) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->slice_conflicts = slice_conflicts;
    return 0;
}
extern "C" int b200sat_engine_set_budget(b200sat_engine *e, int64_t conflict_budget, int64_t propagation_budget) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    e->conflict_budget = conflict_budget < 0 ? -1 : conflict_budget;
    e->propagation_budget = propagation_budget < 0 ? -1 : propagation_budget;
    return 0;
}
extern "C" int b200sat_engine_interrupt(b200sat_engine *e, int raised) {
    if (!e || !e->h_interrupt) return set_err(B200SAT_E_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    *e->h_interrupt = raised ? 1 : 0;
    CK(cudaMemcpyAsync(e->d_interrupt, e->h_interrupt, 4, cudaMemcpyHostToDevice, e->side_stream));
    CK(cudaStreamSynchronize(e->side_stream));
    return 0;
}

static int alloc_outputs(b200sat_engine *e) {
    int rc;
    if ((rc = ensure(e->d_results, e->cap_results, (size_t)e->n_inst))) return rc;
    e->model_stride = e->shape.max_vars ? e->shape.max_vars : 1;
    if ((rc = ensure(e->d_models, e->cap_models, (size_t)e->n_inst * e->model_stride))) return rc;
    if ((rc = ensure(e->d_work[0], e->cap_work, (size_t)e->n_inst))) return rc;
    if ((rc = ensure(e->d_work[1], e->cap_work1, (size_t)e->n_inst))) return rc;
    return 0;
}

static int device_shape(b200sat_engine *e, cudaStream_t stream) {
    e->shape = BatchShape{0, 0, 0, 0};
    if (e->n_inst == 0) return 0;
    u32 *d_out = e->d_queue + 8;
    CK(cudaMemsetAsync(d_out, 0, 20, stream));
    batch_shape<<<e->num_sms * 4, 256, 0, stream>>>(e->d_ioff, e->d_coff, e->d_lits, e->n_inst, e->n_clauses, e->n_lits, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_pinned + 8, d_out, 20, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const u32 *h = e->h_pinned + 8;
    if (h[4]) return set_err(B200SAT_E_ARG, "CSR offsets are not monotone / out of range");
    e->shape.max_vars = e->n_lits ? h[0] + 1 : 0;
    e->shape.max_clause_len = h[1];
    e->shape.max_clauses = h[2];
    e->shape.max_lits = h[3];
    return 0;
}

extern "C" int b200sat_engine_upload(b200sat_engine *e, const b200sat_batch *b, void *stream_) {
    if (!e || !b) return set_err(B200SAT_E_ARG, "null argument");
    e->have_var_flags = false; e->have_clause_nvars = false; e->resident_items = 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    const u32 n = b->n_inst;
    const u64 nc = n ? b->inst_clause_off[n] : 0;
    const u64 nl = nc ? b->clause_off[nc] : 0;
    BatchShape sh = {0, 0, 0, 0};
    int rc;
    if ((rc = ensure(e->d_ioff, e->cap_ioff, (size_t)n + 1))) return rc;
    if ((rc = ensure(e->d_coff, e->cap_coff, (size_t)nc + 1))) return rc;
    if ((rc = ensure(e->d_lits, e->cap_lits, (size_t)nl + 1))) return rc;
    /* caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)
     * are copied from directly; pageable ones are staged through pinned memory so the copies are async DMA */
    {
        cudaPointerAttributes a0, a1, a2;
        bool pinned = cudaPointerGetAttributes(&a0, b->inst_clause_off) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
                      cudaPointerGetAttributes(&a1, b->clause_off) == cudaSuccess && a1.type == cudaMemoryTypeHost &&
                      (nl == 0 || (cudaPointerGetAttributes(&a2, b->lits) == cudaSuccess && a2.type == cudaMemoryTypeHost));
        cudaGetLastError();
        if (pinned) {
            CK(cudaMemcpyAsync(e->d_ioff, b->inst_clause_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(e->d_coff, b->clause_off, ((size_t)nc + 1) * 4, cudaMemcpyHostToDevice, stream));
            if (nl) CK(cudaMemcpyAsync(e->d_lits, b->lits, (size_t)nl * 4, cudaMemcpyHostToDevice, stream));
            e->n_inst = n; e->n_clauses = nc; e->n_lits = nl;
            if ((rc = device_shape(e, stream))) return rc;
            return alloc_outputs(e);
        }
    }
    size_t bytes = ((size_t)n + 1 + nc + 1 + nl) * 4;
    if (bytes > e->cap_stage) {
        if (e->h_stage) cudaFreeHost(e->h_stage);
        e->h_stage = nullptr; e->cap_stage = 0;
        CK(cudaMallocHost(&e->h_stage, bytes + bytes / 8));
        e->cap_stage = bytes + bytes / 8;
    }
    u32 *hs = (u32 *)e->h_stage;
    memcpy(hs, b->inst_clause_off, ((size_t)n + 1) * 4);
    if (n) {
        memcpy(hs + n + 1, b->clause_off, ((size_t)nc + 1) * 4);
        if (nl) memcpy(hs + n + 1 + nc + 1, b->lits, (size_t)nl * 4);
    } else hs[1] = 0;
    CK(cudaMemcpyAsync(e->d_ioff, hs, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(e->d_coff, hs + n + 1, ((size_t)nc + 1) * 4, cudaMemcpyHostToDevice, stream));
    if (nl) CK(cudaMemcpyAsync(e->d_lits, hs + n + 1 + nc + 1, (size_t)nl * 4, cudaMemcpyHostToDevice, stream));
    e->n_inst = n; e->n_clauses = nc; e->n_lits = nl;
    (void)sh;
    if ((rc = device_shape(e, stream))) return rc;
    return alloc_outputs(e);
}

extern "C" int b200sat_engine_generate_ksat(b200sat_engine *e, uint64_t seed_base, uint32_t n_inst, uint32_t n_vars,
                                            uint32_t n_clauses, uint32_t k, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    if (k == 0 || k > n_vars) return set_err(B200SAT_E_ARG, "k must be in 1..n_vars");
    e->have_var_flags = false; e->have_clause_nvars = false; e->resident_items = 0;
    if ((u64)n_inst * n_clauses * k > 0xFFFFFFF0ull) return set_err(B200SAT_E_ARG, "batch exceeds 2^32 literals");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    const u64 nc = (u64)n_inst * n_clauses, nl = nc * k;
    int rc;
    if ((rc = ensure(e->d_ioff, e->cap_ioff, (size_t)n_inst + 1))) return rc;
    if ((rc = ensure(e->d_coff, e->cap_coff, (size_t)nc + 1))) return rc;
    if ((rc = ensure(e->d_lits, e->cap_lits, (size_t)nl + 1))) return rc;
    gen_random_ksat<<<(n_inst + 1 + 127) / 128, 128, 0, stream>>>(seed_base, n_inst, n_vars, n_clauses, k, e->d_ioff, e->d_coff, e->d_lits);
    CK(cudaGetLastError());
    e->n_inst = n_inst; e->n_clauses = nc; e->n_lits = nl;
    e->shape.max_vars = n_vars; e->shape.max_clauses = n_clauses; e->shape.max_lits = (u64)n_clauses * k; e->shape.max_clause_len = k;
    return alloc_outputs(e);
}

static kernel_fn pick_ingest(int tier_nv, bool count, int kind, size_t &smem_per_warp, size_t &save) {
#ifdef B200SAT_HAVE_SIMP
    if (kind == B200SAT_SIMP)
        return count ? b200sat_pick_ingest_c1s1(tier_nv, &smem_per_warp, &save) : b200sat_pick_ingest_c0s1(tier_nv, &smem_per_warp, &save);
#endif
    return count ? b200sat_pick_ingest_c1s0(tier_nv, &smem_per_warp, &save) : b200sat_pick_ingest_c0s0(tier_nv, &smem_per_warp, &save);
}
static kernel_fn pick_search(int tier_nv, bool count, bool portfolio, size_t &smem_per_warp, size_t &save) {
    if (portfolio) return b200sat_pick_search_pf(tier_nv, &smem_per_warp, &save);
    return count ? b200sat_pick_search_c1(tier_nv, &smem_per_warp, &save) : b200sat_pick_search_c0(tier_nv, &smem_per_warp, &save);
}

extern "C" int b200sat_engine_finish(b200sat_engine *e);

static int check_settings(const b200sat_settings *s, int kind) {
    if (!s) return set_err(B200SAT_E_ARG, "null settings");
    if (!(s->var_decay > 0 && s->clause_decay > 0)) return set_err(B200SAT_E_ARG, "decay must be > 0");
    if (kind != B200SAT_CORE && kind != B200SAT_SIMP) return set_err(B200SAT_E_ARG, "kind must be CORE or SIMP");
#ifndef B200SAT_HAVE_SIMP
    if (kind == B200SAT_SIMP) return set_err(B200SAT_E_ARG, "SimpSolver path not built into this library");
#endif
    return 0;
}

__global__ void init_qctl(u32 *qctl, u32 total) {
    if (threadIdx.x < QC_WORDS) qctl[threadIdx.x] = threadIdx.x == QC_TOTAL ? total : 0u;
}
/* between the ingest and the search kernel: the first min(runnable, warps) ring slots are handed out statically */
__global__ void start_ring(u32 *qctl, u32 n_warps) {
    const u32 t = qctl[QC_TAIL];
    const u32 h = t < n_warps ? t : n_warps;
    qctl[QC_HEAD] = h;
    qctl[QC_HEAD0] = h;
}

/* resume: every parked item goes back on the work ring, everything else counts as done */
__global__ void requeue_parked(const u8 *parked, u32 n_items, u32 *qctl, unsigned long long *ring, u32 ring_mask) {
    const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_items) return;
    if (parked[w]) {
        const u32 s = atomicAdd(qctl + QC_TAIL, 1u);
        ring[s & ring_mask] = ((unsigned long long)(s + 1) << 32) | w;
    } else atomicAdd(qctl + QC_DONE, 1u);
}

/* launch shape of one persistent kernel: as many resident warps per SM as shared memory and registers allow */
struct LaunchShape { int wpb = 1, bps = 1; u32 grid = 0; size_t smem = 0; };
static int shape_kernel(b200sat_engine *e, kernel_fn fn, size_t spw, u32 n_items, int max_wpb, LaunchShape &out) {
    int best_wpb = 1, best_bps = 1, best_warps = 0;
    const int cand[] = {1, 2};
    for (int ci = 0; ci < max_wpb; ci++) {
        int wpb = (e->cfg_wpb > 0 && e->cfg_wpb <= max_wpb) ? e->cfg_wpb : cand[ci];
        if (wpb > 2) return set_err(B200SAT_E_ARG, "warps_per_block must be 1 or 2 (launch bounds)");
        size_t smem = spw * wpb;
        if (smem > e->smem_optin) { if (e->cfg_wpb > 0) return set_err(B200SAT_E_ARG, "warps_per_block needs too much shared memory"); continue; }
        CK(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int bps = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void *)fn, wpb * 32, smem));
        if (e->cfg_bps > 0 && bps > e->cfg_bps) bps = e->cfg_bps;
        if (bps * wpb > best_warps) { best_warps = bps * wpb; best_wpb = wpb; best_bps = bps; }
        if (e->cfg_wpb > 0 && e->cfg_wpb <= max_wpb) break;
    }
    if (best_warps == 0) return set_err(B200SAT_E_CUDA, "kernel does not fit on the device");
    u32 grid = (u32)(e->num_sms * best_bps);
    u32 need_blocks = (n_items + best_wpb - 1) / best_wpb;
    if (grid > need_blocks) grid = need_blocks;
    if (grid == 0) grid = 1;
    out.wpb = best_wpb; out.bps = best_bps; out.grid = grid; out.smem = spw * best_wpb;
    CK(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out.smem));
    {   /* give shared memory only what the resident blocks need; the rest of the 256 KB stays L1 for the slab traffic */
        int per_sm = (int)((grid + e->num_sms - 1) / e->num_sms);
        if (per_sm > best_bps) per_sm = best_bps;
        size_t need = (size_t)per_sm * (out.smem + 1024);
        int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    return 0;
}

static void fill_params(b200sat_engine *e, KParams &P, const Layout &L) {
    b200sat_engine::Pending &pd = e->pend;
    memset(&P, 0, sizeof P);
    P.inst_clause_off = e->d_ioff; P.clause_off = e->d_coff; P.lits = e->d_lits;
    P.work = pd.work; P.queue = e->d_queue;
    P.results = e->d_results; P.models = e->d_models; P.model_stride = e->model_stride;
    P.slabs = e->d_slabs; P.L = L; P.st = pd.st; P.kind = pd.kind; P.flags = pd.flags; P.attempt = pd.attempt + 1;
    P.portfolio = pd.portfolio; P.replica_base = pd.replica_base; P.share = pd.share;
    P.n_rings = e->n_rings; P.my_ring = e->my_ring; P.ring_words = e->ring_words;
    for (u32 r = 0; r < 8; r++) P.rings[r] = e->rings[r];
    P.replica_settings = e->d_replica_settings;
    P.qctl = e->d_queue + 32; P.ring = e->d_ring; P.slice_conflicts = e->slice_conflicts;
    P.conflict_budget = e->conflict_budget; P.propagation_budget = e->propagation_budget;
    P.interrupt = e->d_interrupt;
    P.import_cap = e->import_cap;
    P.trace_buf = e->trace_cap ? e->d_trace : nullptr; P.trace_cap = e->trace_cap;
    P.parked = e->d_parked;
    if (e->progress_rows && e->d_progress && !pd.portfolio) { P.progress_buf = e->d_progress; P.progress_cap = e->progress_rows; }
    if (e->have_var_flags && !pd.portfolio) { P.var_flags = e->d_var_flags; P.inst_var_off = e->d_var_off; P.clause_nvars = e->have_clause_nvars ? e->d_clause_nvars : nullptr; }
    P.assumps = e->d_assumps; P.n_assumps = e->n_assumps;
}

/* enqueue one attempt on the pending stream, no host sync: per chunk of resident items the ingest kernel (every
 * item's state ends parked in its own slab), then the search kernel (device work ring, conflict slices), then the
 * failure collection + count read-back */
static int launch_attempt(b200sat_engine *e) {
    b200sat_engine::Pending &pd = e->pend;
    cudaStream_t stream = pd.stream;
    int rc;
    const bool count = (pd.flags & B200SAT_F_COUNT_TRAFFIC) != 0;
    if (pd.attempt == 3) pd.tier_nv = 0; /* last resort: var state in HBM too, biggest slabs */
    size_t spw_i = 0, spw_s = 0, save = 0, save2 = 0;
    kernel_fn fn_i = pick_ingest(pd.tier_nv, count, pd.kind, spw_i, save);
    kernel_fn fn_s = pick_search(pd.tier_nv, count, pd.portfolio != 0, spw_s, save2);
    const bool rescue = e->rescue && !pd.portfolio && !e->replay_mode && !count && !(pd.flags & B200SAT_F_NO_SOLVE);
    if (rescue) fn_s = b200sat_pick_search_rescue(pd.tier_nv, &spw_s, &save2);
    if (e->replay_mode) fn_s = e->replay_mode == 1 ? b200sat_pick_replay_c1(pd.tier_nv, &spw_s, &save2) : b200sat_pick_replay_c0(pd.tier_nv, &spw_s, &save2);
    Layout L = plan_layout(e->shape, pd.tier_nv, pd.kind, pd.attempt, e->cfg_arena, e->cfg_wpool, (u32)save);
    /* resident items per chunk: every item owns a slab; bounded by ~60% of the free HBM */
    u64 slots = pd.n_work;
    if (slots * L.slab_bytes > e->cap_slabs) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = (size_t)((double)(free_b + e->cap_slabs) * (rescue ? 0.3 : 0.6));
        if (e->cfg_slab_budget && e->cfg_slab_budget < budget) budget = e->cfg_slab_budget;
        if (slots * L.slab_bytes > budget) {
            slots = budget / L.slab_bytes;
            if (slots == 0) return set_err(B200SAT_E_MEM, "an instance needs a slab larger than the free device memory");
        }
    }
    LaunchShape shi, shs;
    if ((rc = shape_kernel(e, fn_i, spw_i, (u32)slots, 2, shi))) return rc;
    if ((rc = shape_kernel(e, fn_s, spw_s, (u32)slots, 1, shs))) return rc; /* the search kernels are one warp per block */
    if (slots * L.slab_bytes > e->cap_slabs) {
        CK(cudaStreamSynchronize(stream)); /* the old slabs may still be in use by an earlier launch on this stream */
        if ((rc = ensure(e->d_slabs, e->cap_slabs, (size_t)(slots * L.slab_bytes)))) return rc;
    }
    u32 ring_cap = 64;
    while (ring_cap < 2 * slots) ring_cap <<= 1;
    if (ring_cap > e->cap_ring) {
        CK(cudaStreamSynchronize(stream));
        if ((rc = ensure(e->d_ring, e->cap_ring, (size_t)ring_cap))) return rc;
    }
    const u32 rr_words = 32768; /* 128 KB per running warp: the last stragglers collect thousands of helpers */
    if (rescue) {
        if (pd.n_work > slots) return set_err(B200SAT_E_MEM, "straggler rescue needs the whole batch resident at once");
        if ((rc = ensure(e->d_pristine, e->cap_pristine, (size_t)(slots * L.slab_bytes)))) return rc;
        if ((rc = ensure(e->d_run_table, e->cap_run_table, (size_t)shs.grid + 32))) return rc;
        if ((rc = ensure(e->d_rrings, e->cap_rrings, (size_t)shs.grid * (RING_PAYLOAD + rr_words)))) return rc;
        CK(cudaMemsetAsync(e->d_run_table, 0, ((size_t)shs.grid + 32) * 4, stream));
        CK(cudaMemsetAsync(e->d_rrings, 0, (size_t)shs.grid * (RING_PAYLOAD + rr_words) * 4, stream));
        /* diversified settings of the helper replicas (entry 0 = the reference configuration) */
        const u32 n_div = 64;
        std::vector<b200sat_settings> hs(n_div);
        for (u32 i = 0; i < n_div; i++) diversify_settings(&pd.st, i, &hs[i]);
        if ((rc = ensure(e->d_replica_settings, e->cap_replica_settings, (size_t)n_div))) return rc;
        CK(cudaMemcpyAsync(e->d_replica_settings, hs.data(), n_div * sizeof(b200sat_settings), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream)); /* hs is a local */
    }
    if ((rc = ensure(e->d_parked, e->cap_parked, (size_t)std::max<u64>(pd.n_work, 1)))) return rc;
    CK(cudaMemsetAsync(e->d_parked, 0, (size_t)std::max<u64>(pd.n_work, 1), stream));
    KParams P;
    fill_params(e, P, L);
    P.ring_mask = ring_cap - 1;
    if (rescue) {
        P.rescue = 1; P.run_table = e->d_run_table; P.rrings = e->d_rrings; P.rr_words = rr_words; P.pristine = e->d_pristine;
        P.n_div = 64; P.replica_settings = e->d_replica_settings; P.share = PF_SHARE; P.import_cap = e->import_cap;
    }
    CK(cudaEventRecord(e->ev0, stream));
    float dummy = 0.f; (void)dummy;
    bool first_chunk = true;
    for (u64 base = 0; base < pd.n_work; base += slots) {
        const u32 n_chunk = (u32)std::min<u64>(slots, pd.n_work - base);
        P.chunk_base = (u32)base; P.n_work = n_chunk;
        CK(cudaMemsetAsync(e->d_queue, 0, 4, stream));
        CK(cudaMemsetAsync(e->d_ring, 0, (size_t)ring_cap * 8, stream));
        init_qctl<<<1, 32, 0, stream>>>(P.qctl, n_chunk);
        fn_i<<<shi.grid, shi.wpb * 32, shi.smem, stream>>>(P);
        CK(cudaGetLastError());
        if (e->replay_mode) {
            CK(cudaMemsetAsync(e->d_queue, 0, 4, stream));
            if (first_chunk) CK(cudaEventRecord(e->ev2, stream));
            fn_s<<<shs.grid, shs.wpb * 32, shs.smem, stream>>>(P);
            CK(cudaGetLastError());
            if (first_chunk) CK(cudaEventRecord(e->ev3, stream));
            e->info.launches += 1;
        } else if (!(pd.flags & B200SAT_F_NO_SOLVE)) {
            if (rescue) CK(cudaMemcpyAsync(e->d_pristine, e->d_slabs, (size_t)n_chunk * L.slab_bytes, cudaMemcpyDeviceToDevice, stream));
            if (first_chunk) CK(cudaEventRecord(e->ev2, stream));
            start_ring<<<1, 1, 0, stream>>>(P.qctl, shs.grid * shs.wpb);
            fn_s<<<shs.grid, shs.wpb * 32, shs.smem, stream>>>(P);
            CK(cudaGetLastError());
            if (first_chunk) CK(cudaEventRecord(e->ev3, stream));
            e->info.launches += 1;
        }
        e->info.launches += 1;
        first_chunk = false;
    }
    CK(cudaEventRecord(e->ev1, stream));
    pd.searched = !(pd.flags & B200SAT_F_NO_SOLVE) || e->replay_mode != 0;
    pd.next_list = e->d_work[pd.attempt & 1];
    CK(cudaMemsetAsync(e->d_queue + 1, 0, 4, stream));
    collect_failed<<<(pd.n_work + 255) / 256, 256, 0, stream>>>(e->d_results, pd.work, pd.n_work, pd.portfolio != 0, pd.next_list, e->d_queue + 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_pinned, e->d_queue + 1, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(e->h_pinned + 32, e->d_queue + 32, QC_WORDS * 4, cudaMemcpyDeviceToHost, stream));
    e->info.grid = shs.grid; e->info.block = shs.wpb * 32; e->info.smem_bytes = (u32)shs.smem;
    e->info.warps_total = shs.grid * shs.wpb; e->info.tier = L.tier; e->info.slab_bytes = L.slab_bytes;
    e->info.resident_items = (u32)slots;
    e->info.max_warps = (u32)(e->num_sms * shs.bps * shs.wpb);
    e->resident_items = (pd.n_work <= slots && pd.work == nullptr) ? pd.n_work : 0;
    e->resident_L = L; e->resident_tier_nv = pd.tier_nv; e->resident_attempt = pd.attempt;
    return 0;
}

/* Asynchronous solve of the resident batch: enqueues the first attempt on `stream` and returns.
 * b200sat_engine_finish() waits for it and runs the larger-slab retry passes if any instance needs one. */
extern "C" int b200sat_engine_solve_async(b200sat_engine *e, const b200sat_settings *s, int kind, int flags, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    int rc = check_settings(s, kind);
    if (rc) return rc;
    if (e->pend.active) { rc = b200sat_engine_finish(e); if (rc && rc != B200SAT_E_MEM) return rc; }
    CK(cudaSetDevice(e->device));
    e->info.launches = 0;
    e->info.last_kernel_ms = 0.f; e->info.search_ms = 0.f; e->info.suspends = 0; e->info.tail_ms = 0.f; e->info.span_ms = 0.f; e->info.rescued = 0;
    b200sat_engine::Pending &pd = e->pend;
    pd.stream = (cudaStream_t)stream_;
    if (e->n_inst == 0) return 0;
    pd.st = *s; pd.kind = kind; pd.flags = flags;
    pd.attempt = 0; pd.n_work = e->n_inst; pd.work = nullptr; pd.total_ms = 0.f; pd.search_ms = 0.f; pd.retried = false; pd.resumed = false;
    if (e->progress_rows && (flags & B200SAT_F_COUNT_TRAFFIC)) {
        const size_t words = (size_t)e->n_inst * (1 + 8 * (size_t)e->progress_rows);
        if ((rc = ensure(e->d_progress, e->cap_progress, words))) return rc;
        CK(cudaMemsetAsync(e->d_progress, 0, words * sizeof(double), pd.stream));
    }
    /* model rows of instances that end without a model read "undefined" */
    if (e->d_models) CK(cudaMemsetAsync(e->d_models, 2, (size_t)e->n_inst * e->model_stride, pd.stream));
    pd.portfolio = 0; pd.replica_base = 0; pd.share = 0;
    pd.tier_nv = pick_tier_nv(e->shape.max_vars, e->shape.max_clauses);
    rc = launch_attempt(e);
    if (rc) return rc;
    pd.active = true;
    return 0;
}

/* Search again on the instances the last solve left parked: the ones that returned B200SAT_INTERRUPTED (budget or
 * interrupt: SolveRes::Interrupted hands the solver back, sat.rs:24) or were only ingested (B200SAT_F_NO_SOLVE:
 * preprocess() without solve_limited()).  Nothing is ingested again; settings, kind and flags are those of that solve,
 * budget / slice / assumptions are the engine's current ones.  Requires the whole batch to be resident (one chunk). */
extern "C" int b200sat_engine_resume_async(b200sat_engine *e, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    int rc;
    if (e->pend.active) { rc = b200sat_engine_finish(e); if (rc && rc != B200SAT_E_MEM) return rc; }
    if (e->n_inst == 0) return 0;
    if (e->resident_items != e->n_inst) return set_err(B200SAT_E_ARG, "nothing resident to resume (solve or preprocess first; the batch must fit in one chunk)");
    CK(cudaSetDevice(e->device));
    b200sat_engine::Pending &pd = e->pend;
    cudaStream_t stream = (cudaStream_t)stream_;
    pd.stream = stream;
    pd.flags &= ~B200SAT_F_NO_SOLVE;
    pd.work = nullptr; pd.n_work = e->n_inst; pd.total_ms = 0.f; pd.search_ms = 0.f; pd.retried = false;
    pd.attempt = e->resident_attempt;
    e->info.launches = 0; e->info.last_kernel_ms = 0.f; e->info.search_ms = 0.f; e->info.suspends = 0; e->info.tail_ms = 0.f; e->info.span_ms = 0.f;
    const bool count = (pd.flags & B200SAT_F_COUNT_TRAFFIC) != 0;
    size_t spw_s = 0, save2 = 0;
    kernel_fn fn_s = pick_search(e->resident_tier_nv, count, false, spw_s, save2);
    LaunchShape shs;
    if ((rc = shape_kernel(e, fn_s, spw_s, e->n_inst, 1, shs))) return rc;
    KParams P;
    fill_params(e, P, e->resident_L);
    u32 ring_cap = 64;
    while (ring_cap < 2 * (u64)e->n_inst) ring_cap <<= 1;
    P.ring_mask = ring_cap - 1;
    P.chunk_base = 0; P.n_work = e->n_inst;
    CK(cudaEventRecord(e->ev0, stream));
    CK(cudaMemsetAsync(e->d_ring, 0, (size_t)ring_cap * 8, stream));
    init_qctl<<<1, 32, 0, stream>>>(P.qctl, e->n_inst);
    requeue_parked<<<(e->n_inst + 255) / 256, 256, 0, stream>>>(e->d_parked, e->n_inst, P.qctl, e->d_ring, P.ring_mask);
    CK(cudaEventRecord(e->ev2, stream));
    start_ring<<<1, 1, 0, stream>>>(P.qctl, shs.grid * shs.wpb);
    fn_s<<<shs.grid, shs.wpb * 32, shs.smem, stream>>>(P);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev3, stream));
    CK(cudaEventRecord(e->ev1, stream));
    e->info.launches = 1;
    pd.searched = true;
    pd.next_list = e->d_work[pd.attempt & 1];
    CK(cudaMemsetAsync(e->d_queue + 1, 0, 4, stream));
    collect_failed<<<(pd.n_work + 255) / 256, 256, 0, stream>>>(e->d_results, nullptr, pd.n_work, false, pd.next_list, e->d_queue + 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_pinned, e->d_queue + 1, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(e->h_pinned + 32, e->d_queue + 32, QC_WORDS * 4, cudaMemcpyDeviceToHost, stream));
    pd.active = true;
    pd.resumed = true;
    return 0;
}
extern "C" int b200sat_engine_resume(b200sat_engine *e, void *stream_) {
    int rc = b200sat_engine_resume_async(e, stream_);
    if (rc) return rc;
    return b200sat_engine_finish(e);
}

extern "C" int b200sat_engine_finish(b200sat_engine *e) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    b200sat_engine::Pending &pd = e->pend;
    if (!pd.active) return 0;
    CK(cudaSetDevice(e->device));
    for (;;) {
        CK(cudaStreamSynchronize(pd.stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        pd.total_ms += ms;
        if (pd.searched) { CK(cudaEventElapsedTime(&ms, e->ev2, e->ev3)); pd.search_ms += ms; }
        e->info.suspends += e->h_pinned[32 + QC_SUSPENDS];
        e->info.rescued += e->h_pinned[32 + QC_RESCUED];
        if (pd.searched && !e->replay_mode) { /* time from the moment the work ring ran dry to the last warp's exit */
            const unsigned long long *tq = (const unsigned long long *)(e->h_pinned + 32 + QC_T_START);
            const unsigned long long t0 = tq[0], td = tq[1], t1 = tq[2];
            if (t1 > t0 && td >= t0) { e->info.tail_ms = (float)((double)(t1 - td) * 1e-6); e->info.span_ms = (float)((double)(t1 - t0) * 1e-6); }
        }
        u32 left = e->h_pinned[0];
        if (left == 0 || pd.attempt >= 4) { pd.n_work = left; break; }
        if (pd.resumed) pd.resumed = false; /* instances that outgrew their slab while resumed start over in a larger one */
        pd.retried = true;
        pd.n_work = left;
        pd.work = pd.next_list;
        pd.attempt++;
        int rc = launch_attempt(e);
        if (rc) { pd.active = false; return rc; }
    }
    pd.active = false;
    e->info.last_kernel_ms = pd.total_ms;
    e->info.search_ms = pd.search_ms;
    if (pd.n_work > 0) return set_err(B200SAT_E_MEM, "instances outgrew the largest slab (status in results)");
    return 0;
}

extern "C" int b200sat_engine_solve(b200sat_engine *e, const b200sat_settings *s, int kind, int flags, void *stream_) {
    int rc = b200sat_engine_solve_async(e, s, kind, flags, stream_);
    if (rc) return rc;
    return b200sat_engine_finish(e);
}

/* K2 (SURVEY.md section 7 / 8d "BCP-only figure"): solve the resident batch once with the instrumented kernels while
 * recording every instance's search prefix (decisions, learnt clauses, restarts up to the first database reduction),
 * then re-ingest and run the bcp_replay kernel, which re-executes exactly those prefixes with BCP as the only real
 * work: once with the traffic counters (validation: every replay must reproduce the recorded propagation count and
 * BCP counters; result: algorithmic BCP bytes) and `reps` times plain, CUDA-event timed. */
extern "C" int b200sat_engine_bcp_replay(b200sat_engine *e, const b200sat_settings *s, int reps, b200sat_replay_info *out, void *stream_) {
    if (!e || !out) return set_err(B200SAT_E_ARG, "null argument");
    memset(out, 0, sizeof *out);
    if (e->n_inst == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    int rc;
    const u32 cap = 8192;
    if ((rc = ensure(e->d_trace, e->cap_trace, (size_t)e->n_inst * cap))) return rc;
    CK(cudaMemsetAsync(e->d_trace, 0, (size_t)e->n_inst * cap * 4, stream));
    const int base = B200SAT_F_PREPROCESS | B200SAT_F_ZERO_STATS_ON_PRE_UNSAT;
    e->trace_cap = cap;
    rc = b200sat_engine_solve(e, s, B200SAT_CORE, base | B200SAT_F_COUNT_TRAFFIC, stream_);
    e->trace_cap = 0;
    if (rc) return rc;
    if (e->resident_items != e->n_inst) return set_err(B200SAT_E_MEM, "bcp_replay needs the whole batch resident at once");
    /* the record: header of every instance */
    std::vector<u32> hdr((size_t)e->n_inst * 32);
    CK(cudaMemcpy2D(hdr.data(), 128, e->d_trace, (size_t)cap * 4, 128, e->n_inst, cudaMemcpyDeviceToHost));
    u64 want_props = 0, want_t[6] = {0, 0, 0, 0, 0, 0};
    u32 recorded = 0;
    for (u32 i = 0; i < e->n_inst; i++) {
        const u32 *h = hdr.data() + (size_t)i * 32;
        if (h[TR_WORDS] == 0) continue;
        recorded++;
    }
    std::vector<b200sat_result> res(e->n_inst);
    float best = 0.f, sum = 0.f;
    e->trace_cap = cap; /* the replay kernels read the record */
    for (int rep = 0; rep <= reps; rep++) {
        e->replay_mode = rep == 0 ? 1 : 2;
        rc = b200sat_engine_solve(e, s, B200SAT_CORE, base | B200SAT_F_NO_SOLVE | (rep == 0 ? B200SAT_F_COUNT_TRAFFIC : 0), stream_);
        e->replay_mode = 0;
        if (rc) { e->trace_cap = 0; return rc; }
        if (rep == 0) {
            CK(cudaMemcpy(res.data(), e->d_results, (size_t)e->n_inst * sizeof(b200sat_result), cudaMemcpyDeviceToHost));
            u64 got_t[6] = {0, 0, 0, 0, 0, 0};
            for (u32 i = 0; i < e->n_inst; i++) {
                const u32 *h = hdr.data() + (size_t)i * 32;
                if (h[TR_WORDS] == 0) continue;
                if (res[i].status != B200SAT_OK) { out->mismatches++; continue; }
                const u64 wp = (u64)h[TR_PROPS_LO] | ((u64)h[TR_PROPS_HI] << 32);
                const u64 wp0 = (u64)h[TR_PROPS0_LO] | ((u64)h[TR_PROPS0_HI] << 32);
                bool same = res[i].stats[5] == wp;
                for (int k = 0; k < 6; k++) {
                    const u64 wt = (u64)h[TR_TRAFFIC + 2 * k] | ((u64)h[TR_TRAFFIC + 2 * k + 1] << 32);
                    const u64 wt0 = (u64)h[TR_TRAFFIC0 + 2 * k] | ((u64)h[TR_TRAFFIC0 + 2 * k + 1] << 32);
                    if (res[i].traffic[k] != wt) same = false;
                    want_t[k] += wt; got_t[k] += res[i].traffic[k] - wt0; /* the replay's own share: ingest-time BCP excluded */
                }
                if (!same) out->mismatches++;
                want_props += wp;
                out->propagations += wp - wp0;
            }
            /* SURVEY.md 8d formula, BCP terms only (items 1-5): watchers visited / kept, clause derefs + tails, moves, implications */
            out->bcp_bytes = 8 * got_t[0] + 8 * got_t[1] + 4 * got_t[2] + 4 * (2 * got_t[2] + got_t[3]) + 16 * got_t[4] + 12 * got_t[5];
        } else {
            const float ms = e->info.search_ms;
            sum += ms;
            if (best == 0.f || ms < best) best = ms;
        }
    }
    e->trace_cap = 0;
    out->instances = recorded;
    out->ms_min = best;
    out->ms_avg = reps > 0 ? sum / (float)reps : 0.f;
    out->warps = e->info.warps_total;
    return 0;
}

extern "C" int b200sat_engine_launch_info(b200sat_engine *e, b200sat_launch_info *out) {
    if (!e || !out) return set_err(B200SAT_E_ARG, "null argument");
    *out = e->info;
    return 0;
}

extern "C" int b200sat_engine_download(b200sat_engine *e, b200sat_result *results, uint8_t *models, uint32_t model_stride,
                                       void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    if (e->n_inst == 0) return 0;
    if (results) CK(cudaMemcpyAsync(results, e->d_results, (size_t)e->n_inst * sizeof(b200sat_result), cudaMemcpyDeviceToHost, stream));
    if (models) {
        if (model_stride == e->model_stride)
            CK(cudaMemcpyAsync(models, e->d_models, (size_t)e->n_inst * model_stride, cudaMemcpyDeviceToHost, stream));
        else
            CK(cudaMemcpy2DAsync(models, model_stride, e->d_models, e->model_stride, std::min(model_stride, e->model_stride),
                                 e->n_inst, cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    if (models && model_stride > e->model_stride)
        for (u32 i = 0; i < e->n_inst; i++) memset(models + (size_t)i * model_stride + e->model_stride, 2, model_stride - e->model_stride);
    return 0;
}

extern "C" int b200sat_engine_verify_models(b200sat_engine *e, uint8_t *verified, uint32_t *n_violated, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    if (n_violated) *n_violated = 0;
    if (e->n_inst == 0) return 0;
    int rc;
    if ((rc = ensure(e->d_verified, e->cap_verified, (size_t)e->n_inst))) return rc;
    CK(cudaMemsetAsync(e->d_queue + 2, 0, 4, stream));
    verify_models<<<(e->n_inst * 32 + 255) / 256, 256, 0, stream>>>(e->d_ioff, e->d_coff, e->d_lits, e->d_results, e->d_models,
                                                                     e->model_stride, e->n_inst, e->d_verified, e->d_queue + 2);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_pinned + 2, e->d_queue + 2, 4, cudaMemcpyDeviceToHost, stream));
    if (verified) CK(cudaMemcpyAsync(verified, e->d_verified, e->n_inst, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (n_violated) *n_violated = e->h_pinned[2];
    return 0;
}

/* K4: multi-CTA subsumption of one large instance (csrc/preprocess_kernels.cu) */
extern "C" size_t b200sat_pp_scratch_words(size_t nc, size_t nvars, size_t nlits);
extern "C" int b200sat_pp_subsume_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 *scratch,
                                         u8 *d_removed, u32 *d_out_coff, u32 *d_out_lits, u32 *d_counts, cudaStream_t stream);

extern "C" size_t b200sat_bve_words(size_t nc, size_t nvars, size_t nlits);
extern "C" int b200sat_pp_bve_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base, u32 *mem,
                                     u32 grow, int cl_lim, u32 max_rounds, u32 *d_out_coff, u32 *d_out_lits, u32 **d_estack_out,
                                     u32 *d_counts, u32 *h_pin, cudaStream_t stream);

/* K4: multi-CTA bounded variable elimination of one large instance */
extern "C" int b200sat_engine_bve_instance(b200sat_engine *e, uint32_t inst, uint32_t grow, int cl_lim, uint32_t max_rounds,
                                           uint32_t *out_clause_off, uint32_t *out_lits, uint32_t *out_estack, uint32_t counts[6],
                                           float *kernel_ms, void *stream_) {
    if (!e || inst >= e->n_inst) return set_err(B200SAT_E_ARG, "bad instance index");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    u32 io[2], lo[2];
    CK(cudaMemcpy(io, e->d_ioff + inst, 8, cudaMemcpyDeviceToHost));
    const u32 cb = io[0], nc = io[1] - io[0];
    if (nc == 0) { if (counts) for (int i = 0; i < 6; i++) counts[i] = 0; return 0; }
    CK(cudaMemcpy(&lo[0], e->d_coff + cb, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&lo[1], e->d_coff + cb + nc, 4, cudaMemcpyDeviceToHost));
    const u32 nlits = lo[1] - lo[0], nvars = e->shape.max_vars;
    const size_t NC = 2 * (size_t)nc + 1024, NL = 3 * (size_t)nlits + 4096;
    int rc;
    const size_t bw = b200sat_bve_words(nc, nvars, nlits);
    if ((rc = ensure(e->d_pp, e->cap_pp, bw + NC + 2 + NL + 64))) return rc;
    u32 *mem = e->d_pp;
    u32 *d_out_coff = mem + bw;
    u32 *d_out_lits = d_out_coff + NC + 2;
    u32 *d_counts = e->d_queue + 16;
    u32 *d_estack = nullptr;
    CK(cudaEventRecord(e->ev0, stream));
    if (b200sat_pp_bve_launch(e->d_coff, e->d_lits, cb, nc, nvars, nlits, lo[0], mem, grow, cl_lim, max_rounds, d_out_coff, d_out_lits,
                              &d_estack, d_counts, e->h_pinned + 24, stream))
        return set_err(B200SAT_E_CUDA, "bve kernel launch failed");
    CK(cudaEventRecord(e->ev1, stream));
    CK(cudaMemcpyAsync(e->h_pinned + 16, d_counts, 24, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const u32 kept = e->h_pinned[16], kept_lits = e->h_pinned[17], ewords = e->h_pinned[20];
    if (counts) for (int i = 0; i < 6; i++) counts[i] = e->h_pinned[16 + i];
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, e->ev0, e->ev1));
    if (e->h_pinned[21]) return set_err(B200SAT_E_MEM, "bve: clause store overflow");
    if (out_clause_off) {
        CK(cudaMemcpy(out_clause_off, d_out_coff, (size_t)kept * 4, cudaMemcpyDeviceToHost));
        out_clause_off[kept] = kept_lits;
    }
    if (out_lits && kept_lits) CK(cudaMemcpy(out_lits, d_out_lits, (size_t)kept_lits * 4, cudaMemcpyDeviceToHost));
    if (out_estack && ewords) CK(cudaMemcpy(out_estack, d_estack, (size_t)ewords * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" size_t b200sat_k4_words(size_t nc, size_t nvars, size_t nlits);
extern "C" int b200sat_k4_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base, u32 *mem,
                                 u32 grow, int cl_lim, u32 max_outer, u32 *d_out_coff, u32 *d_out_lits, u32 **d_estack_out,
                                 u32 *h_counts, int num_sms, cudaStream_t stream);
extern "C" int b200sat_k4_extend_model_launch(const u32 *d_estack, u32 n_words, u8 *d_model, u32 n_vars, cudaStream_t stream);

/* K4 as ONE preprocessor (csrc/preprocess_coop.cu): subsumption + strengthening + BVE of instance `inst` of the resident
 * batch, interleaved to a fixpoint inside one cooperative kernel.  counts[16]: [0] kept clauses [1] kept literals
 * [2] eliminated vars [3] subsumed clauses [4] strengthened literals [5] outer iterations [6] BVE rounds
 * [7] subsume/strengthen rounds [8] eliminated-clause stack words [9] overflow [10] empty clause derived (UNSAT)
 * [11] kernel launches.  The eliminated-clause stack stays on the device for b200sat_engine_k4_extend_model. */
extern "C" int b200sat_engine_preprocess_instance(b200sat_engine *e, uint32_t inst, uint32_t grow, int cl_lim, uint32_t max_outer,
                                                  uint32_t *out_clause_off, uint32_t *out_lits, uint32_t *out_estack, uint32_t counts[16],
                                                  float *kernel_ms, void *stream_) {
    if (!e || inst >= e->n_inst) return set_err(B200SAT_E_ARG, "bad instance index");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    u32 io[2], lo[2];
    CK(cudaMemcpy(io, e->d_ioff + inst, 8, cudaMemcpyDeviceToHost));
    const u32 cb = io[0], nc = io[1] - io[0];
    if (counts) for (int i = 0; i < 16; i++) counts[i] = 0;
    if (nc == 0) return 0;
    CK(cudaMemcpy(&lo[0], e->d_coff + cb, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&lo[1], e->d_coff + cb + nc, 4, cudaMemcpyDeviceToHost));
    const u32 nlits = lo[1] - lo[0], nvars = e->shape.max_vars;
    const size_t NC = 2 * (size_t)nc + 1024, NL = 3 * (size_t)nlits + 4096;
    int rc;
    const size_t kw = b200sat_k4_words(nc, nvars, nlits);
    if ((rc = ensure(e->d_pp, e->cap_pp, kw + NC + 2 + NL + 64))) return rc;
    u32 *mem = e->d_pp;
    u32 *d_out_coff = mem + kw;
    u32 *d_out_lits = d_out_coff + NC + 2;
    u32 *d_estack = nullptr;
    u32 *hc = e->h_pinned + 64; /* 32 pinned words */
    CK(cudaEventRecord(e->ev0, stream));
    rc = b200sat_k4_launch(e->d_coff, e->d_lits, cb, nc, nvars, nlits, lo[0], mem, grow, cl_lim, max_outer ? max_outer : 64, d_out_coff,
                           d_out_lits, &d_estack, hc, e->num_sms, stream);
    if (rc == -3) return set_err(B200SAT_E_CUDA, "cooperative launch not available");
    if (rc) return set_err(B200SAT_E_CUDA, std::string("k4 preprocess launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    CK(cudaEventRecord(e->ev1, stream));
    CK(cudaStreamSynchronize(stream));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, e->ev0, e->ev1));
    const u32 kept = hc[13], kept_lits = hc[14], ewords = hc[12];
    if (counts) {
        counts[0] = kept; counts[1] = kept_lits; counts[2] = hc[2]; counts[3] = hc[5]; counts[4] = hc[6]; counts[5] = hc[7];
        counts[6] = hc[8]; counts[7] = hc[9]; counts[8] = ewords; counts[9] = hc[3]; counts[10] = hc[4]; counts[11] = 3;
    }
    e->k4_estack = d_estack; e->k4_estack_words = ewords; e->k4_nvars = nvars;
    if (hc[3]) return set_err(B200SAT_E_MEM, "k4: clause store overflow");
    if (out_clause_off) {
        CK(cudaMemcpy(out_clause_off, d_out_coff, (size_t)kept * 4, cudaMemcpyDeviceToHost));
        out_clause_off[kept] = kept_lits;
    }
    if (out_lits && kept_lits) CK(cudaMemcpy(out_lits, d_out_lits, (size_t)kept_lits * 4, cudaMemcpyDeviceToHost));
    if (out_estack && ewords) CK(cudaMemcpy(out_estack, d_estack, (size_t)ewords * 4, cudaMemcpyDeviceToHost));
    return 0;
}
/* ElimClauses::extend_model (elim_clauses.rs:43-63) on the device, over the stack the last preprocess_instance left
 * there: model bytes in / out (1 true, 0 false, 2 undefined) */
extern "C" int b200sat_engine_k4_extend_model(b200sat_engine *e, uint8_t *model, uint32_t n_vars, void *stream_) {
    if (!e || !model) return set_err(B200SAT_E_ARG, "null argument");
    if (!e->k4_estack) return set_err(B200SAT_E_ARG, "no eliminated-clause stack resident (call preprocess_instance first)");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    int rc;
    if ((rc = ensure(e->d_verified, e->cap_verified, (size_t)n_vars + 16))) return rc;
    CK(cudaMemcpyAsync(e->d_verified, model, n_vars, cudaMemcpyHostToDevice, stream));
    if (b200sat_k4_extend_model_launch(e->k4_estack, e->k4_estack_words, e->d_verified, n_vars, stream)) return set_err(B200SAT_E_CUDA, "extend_model launch failed");
    CK(cudaMemcpyAsync(model, e->d_verified, n_vars, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int b200sat_engine_subsume_instance(b200sat_engine *e, uint32_t inst, uint8_t *removed, uint32_t *out_clause_off,
                                               uint32_t *out_lits, uint32_t counts[4], float *kernel_ms, void *stream_);
extern "C" int b200sat_pp_simplify_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base,
                                          u32 *scratch, u32 *d_work, u8 *d_removed, u32 *d_out_coff, u32 *d_out_lits, u32 *d_counts,
                                          u32 *h_flag, u32 max_rounds, cudaStream_t stream);

/* K4: subsumption + self-subsuming strengthening to fixpoint (mode 1) or subsumption only (mode 0) */
extern "C" int b200sat_engine_simplify_instance(b200sat_engine *e, uint32_t inst, int strengthen, uint8_t *removed, uint32_t *out_clause_off,
                                                uint32_t *out_lits, uint32_t counts[6], float *kernel_ms, void *stream_) {
    if (!strengthen) {
        uint32_t c4[4] = {0, 0, 0, 0};
        int rc0 = b200sat_engine_subsume_instance(e, inst, removed, out_clause_off, out_lits, c4, kernel_ms, stream_);
        if (counts) { counts[0] = c4[0]; counts[1] = c4[1]; counts[2] = 1; counts[3] = 0; counts[4] = 0; counts[5] = c4[3]; }
        return rc0;
    }
    if (!e || inst >= e->n_inst) return set_err(B200SAT_E_ARG, "bad instance index");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    u32 io[2], lo[2];
    CK(cudaMemcpy(io, e->d_ioff + inst, 8, cudaMemcpyDeviceToHost));
    const u32 cb = io[0], nc = io[1] - io[0];
    if (nc == 0) { if (counts) for (int i = 0; i < 6; i++) counts[i] = 0; return 0; }
    CK(cudaMemcpy(&lo[0], e->d_coff + cb, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&lo[1], e->d_coff + cb + nc, 4, cudaMemcpyDeviceToHost));
    const u32 nlits = lo[1] - lo[0], nvars = e->shape.max_vars;
    int rc;
    const size_t sw = b200sat_pp_scratch_words(nc, nvars, nlits);
    size_t words = sw + ((size_t)nc + 2 + nlits + 8) + ((size_t)2 * nc + nlits + 16);
    if ((rc = ensure(e->d_pp, e->cap_pp, words))) return rc;
    if ((rc = ensure(e->d_verified, e->cap_verified, (size_t)nc))) return rc;
    u32 *scratch = e->d_pp;
    u32 *d_out_coff = scratch + sw;
    u32 *d_out_lits = d_out_coff + nc + 2;
    u32 *d_work = d_out_lits + nlits + 6;
    u32 *d_counts = e->d_queue + 16;
    CK(cudaEventRecord(e->ev0, stream));
    if (b200sat_pp_simplify_launch(e->d_coff, e->d_lits, cb, nc, nvars, nlits, lo[0], scratch, d_work, e->d_verified, d_out_coff, d_out_lits,
                                   d_counts, e->h_pinned + 24, 64, stream))
        return set_err(B200SAT_E_CUDA, "preprocess kernel launch failed");
    CK(cudaEventRecord(e->ev1, stream));
    CK(cudaMemcpyAsync(e->h_pinned + 16, d_counts, 24, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const u32 kept = e->h_pinned[16], kept_lits = e->h_pinned[17];
    if (counts) { for (int i = 0; i < 5; i++) counts[i] = e->h_pinned[16 + i]; counts[5] = nc - kept; }
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, e->ev0, e->ev1));
    if (removed) CK(cudaMemcpy(removed, e->d_verified, nc, cudaMemcpyDeviceToHost));
    if (out_clause_off) {
        CK(cudaMemcpy(out_clause_off, d_out_coff, (size_t)kept * 4, cudaMemcpyDeviceToHost));
        out_clause_off[kept] = kept_lits;
    }
    if (out_lits && kept_lits) CK(cudaMemcpy(out_lits, d_out_lits, (size_t)kept_lits * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int b200sat_engine_subsume_instance(b200sat_engine *e, uint32_t inst, uint8_t *removed, uint32_t *out_clause_off,
                                               uint32_t *out_lits, uint32_t counts[4], float *kernel_ms, void *stream_) {
    if (!e || inst >= e->n_inst) return set_err(B200SAT_E_ARG, "bad instance index");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    u32 io[2], lo[2];
    CK(cudaMemcpy(io, e->d_ioff + inst, 8, cudaMemcpyDeviceToHost));
    const u32 cb = io[0], nc = io[1] - io[0];
    if (nc == 0) { if (counts) counts[0] = counts[1] = counts[2] = counts[3] = 0; return 0; }
    CK(cudaMemcpy(&lo[0], e->d_coff + cb, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&lo[1], e->d_coff + cb + nc, 4, cudaMemcpyDeviceToHost));
    const u32 nlits = lo[1] - lo[0], nvars = e->shape.max_vars;
    int rc;
    size_t words = b200sat_pp_scratch_words(nc, nvars, nlits) + (size_t)nc + 2 + nlits + 8;
    if ((rc = ensure(e->d_pp, e->cap_pp, words))) return rc;
    if ((rc = ensure(e->d_verified, e->cap_verified, (size_t)nc))) return rc;
    u32 *scratch = e->d_pp;
    u32 *d_out_coff = scratch + b200sat_pp_scratch_words(nc, nvars, nlits);
    u32 *d_out_lits = d_out_coff + nc + 2;
    u32 *d_counts = e->d_queue + 16;
    CK(cudaEventRecord(e->ev0, stream));
    if (b200sat_pp_subsume_launch(e->d_coff, e->d_lits, cb, nc, nvars, nlits, scratch, e->d_verified, d_out_coff, d_out_lits, d_counts, stream))
        return set_err(B200SAT_E_CUDA, "preprocess kernel launch failed");
    CK(cudaEventRecord(e->ev1, stream));
    CK(cudaMemcpyAsync(e->h_pinned + 16, d_counts, 16, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const u32 kept = e->h_pinned[16], kept_lits = e->h_pinned[17];
    if (counts) { counts[0] = kept; counts[1] = kept_lits; counts[2] = e->h_pinned[18]; counts[3] = nc - kept; }
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, e->ev0, e->ev1));
    if (removed) CK(cudaMemcpy(removed, e->d_verified, nc, cudaMemcpyDeviceToHost));
    if (out_clause_off) {
        CK(cudaMemcpy(out_clause_off, d_out_coff, (size_t)kept * 4, cudaMemcpyDeviceToHost));
        out_clause_off[kept] = kept_lits;
    }
    if (out_lits && kept_lits) CK(cudaMemcpy(out_lits, d_out_lits, (size_t)kept_lits * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int b200sat_engine_batch_dims(b200sat_engine *e, uint32_t *n_inst, uint64_t *n_clauses, uint64_t *n_lits) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    if (n_inst) *n_inst = e->n_inst;
    if (n_clauses) *n_clauses = e->n_clauses;
    if (n_lits) *n_lits = e->n_lits;
    return 0;
}

extern "C" int b200sat_engine_download_batch(b200sat_engine *e, uint32_t *ioff, uint32_t *coff, uint32_t *lits) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaDeviceSynchronize());
    if (ioff) CK(cudaMemcpy(ioff, e->d_ioff, ((size_t)e->n_inst + 1) * 4, cudaMemcpyDeviceToHost));
    if (coff) CK(cudaMemcpy(coff, e->d_coff, ((size_t)e->n_clauses + 1) * 4, cudaMemcpyDeviceToHost));
    if (lits && e->n_lits) CK(cudaMemcpy(lits, e->d_lits, (size_t)e->n_lits * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int enqueue_download(b200sat_engine *e, b200sat_result *results, uint8_t *models, uint32_t model_stride, cudaStream_t stream) {
    if (e->n_inst == 0) return 0;
    if (results) CK(cudaMemcpyAsync(results, e->d_results, (size_t)e->n_inst * sizeof(b200sat_result), cudaMemcpyDeviceToHost, stream));
    if (models) {
        if (model_stride == e->model_stride)
            CK(cudaMemcpyAsync(models, e->d_models, (size_t)e->n_inst * model_stride, cudaMemcpyDeviceToHost, stream));
        else
            CK(cudaMemcpy2DAsync(models, model_stride, e->d_models, e->model_stride, std::min(model_stride, e->model_stride),
                                 e->n_inst, cudaMemcpyDeviceToHost, stream));
    }
    return 0;
}

/* Asynchronous end-to-end call: H2D of the CSR batch, solve, D2H of results/models, all enqueued on
 * `stream`.  The caller's buffers must stay valid (and should be page-locked) until
 * b200sat_solve_batch_finish() returns.  Two engines on two streams pipeline consecutive batches. */
extern "C" int b200sat_solve_batch_async(b200sat_engine *e, const b200sat_batch *batch, const b200sat_settings *s, int kind, int flags,
                                         b200sat_result *results, uint8_t *models, uint32_t model_stride, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    int rc;
    if (e->pend.active) { rc = b200sat_engine_finish(e); if (rc && rc != B200SAT_E_MEM) return rc; }
    rc = b200sat_engine_upload(e, batch, stream_);
    if (rc) return rc;
    rc = b200sat_engine_solve_async(e, s, kind, flags, stream_);
    if (rc) return rc;
    e->pend.batch = true; e->pend.out_results = results; e->pend.out_models = models; e->pend.out_stride = model_stride;
    return enqueue_download(e, results, models, model_stride, (cudaStream_t)stream_);
}

extern "C" int b200sat_solve_batch_finish(b200sat_engine *e) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    b200sat_engine::Pending &pd = e->pend;
    const bool was_batch = pd.batch;
    pd.batch = false;
    int src = b200sat_engine_finish(e);
    if (src && src != B200SAT_E_MEM) return src;
    if (was_batch) {
        if (pd.retried) { /* results changed after the first download: fetch them again */
            int rc = enqueue_download(e, pd.out_results, pd.out_models, pd.out_stride, pd.stream);
            if (rc) return rc;
        }
        CK(cudaStreamSynchronize(pd.stream));
        if (pd.out_models && pd.out_stride > e->model_stride)
            for (u32 i = 0; i < e->n_inst; i++)
                memset(pd.out_models + (size_t)i * pd.out_stride + e->model_stride, 2, pd.out_stride - e->model_stride);
    }
    return src;
}

extern "C" int b200sat_solve_batch(b200sat_engine *e, const b200sat_batch *batch, const b200sat_settings *s, int kind, int flags,
                                   b200sat_result *results, uint8_t *models, uint32_t model_stride) {
    int rc = b200sat_solve_batch_async(e, batch, s, kind, flags, results, models, model_stride, nullptr);
    if (rc) return rc;
    return b200sat_solve_batch_finish(e);
}

/* ------------------------------------------------------------------ host: portfolio */
extern "C" void b200sat_diversify(const b200sat_settings *base, uint32_t g, b200sat_settings *out) {
    diversify_settings(base, g, out);
}

extern "C" int b200sat_engine_ring_create(b200sat_engine *e, uint32_t payload_words, uint32_t n_rings, uint32_t my_slot) {
    if (!e || n_rings == 0 || n_rings > 8 || my_slot >= n_rings) return set_err(B200SAT_E_ARG, "bad ring arguments");
    CK(cudaSetDevice(e->device));
    if (e->rings[my_slot] && !e->ring_is_peer[my_slot]) cudaFree(e->rings[my_slot]);
    e->rings[my_slot] = nullptr;
    CK(cudaMalloc((void **)&e->rings[my_slot], ((size_t)payload_words + RING_PAYLOAD) * 4));
    CK(cudaMemset(e->rings[my_slot], 0, ((size_t)payload_words + RING_PAYLOAD) * 4));
    e->ring_is_peer[my_slot] = false;
    e->n_rings = n_rings; e->my_ring = my_slot; e->ring_words = payload_words;
    return 0;
}
extern "C" int b200sat_engine_ring_reset(b200sat_engine *e, void *stream) {
    if (!e || !e->rings[e->my_ring]) return set_err(B200SAT_E_ARG, "no ring");
    CK(cudaSetDevice(e->device));
    CK(cudaMemsetAsync(e->rings[e->my_ring], 0, ((size_t)e->ring_words + RING_PAYLOAD) * 4, (cudaStream_t)stream));
    return 0;
}
extern "C" int b200sat_engine_ring_ipc_handle(b200sat_engine *e, uint8_t handle[64]) {
    if (!e || !e->rings[e->my_ring]) return set_err(B200SAT_E_ARG, "no ring");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->rings[e->my_ring]));
    memcpy(handle, &h, 64);
    return 0;
}
extern "C" int b200sat_engine_ring_open_peer(b200sat_engine *e, uint32_t slot, const uint8_t handle[64]) {
    if (!e || slot >= e->n_rings || slot == e->my_ring) return set_err(B200SAT_E_ARG, "bad peer slot");
    CK(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->rings[slot] = (u32 *)p;
    e->ring_is_peer[slot] = true;
    return 0;
}

/* single-process multi-GPU use (tests, embedding hosts): the ring of another engine of THIS process is mapped through
 * CUDA peer access instead of an IPC handle (cudaIpcOpenMemHandle refuses handles of the calling process) */
extern "C" int b200sat_engine_ring_open_local_peer(b200sat_engine *e, uint32_t slot, b200sat_engine *owner) {
    if (!e || !owner || slot >= e->n_rings || slot == e->my_ring || !owner->rings[owner->my_ring]) return set_err(B200SAT_E_ARG, "bad peer slot");
    if (owner->ring_words != e->ring_words) return set_err(B200SAT_E_ARG, "rings must have the same size");
    CK(cudaSetDevice(e->device));
    if (owner->device != e->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, e->device, owner->device));
        if (!can) return set_err(B200SAT_E_CUDA, "no peer access between the two devices");
        cudaError_t pe = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return set_err(B200SAT_E_CUDA, cudaGetErrorString(pe));
        cudaGetLastError();
    }
    e->rings[slot] = owner->rings[owner->my_ring];
    e->ring_is_peer[slot] = true;
    e->ring_is_local_peer[slot] = true;
    return 0;
}

extern "C" int b200sat_portfolio_solve(b200sat_engine *e, const b200sat_settings *base, uint32_t n_replicas, uint32_t replica_base,
                                       uint32_t share_flags, int flags, void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    int rc = check_settings(base, B200SAT_CORE);
    if (rc) return rc;
    if (e->n_inst < 1 || n_replicas == 0) return set_err(B200SAT_E_ARG, "portfolio needs a resident instance and >= 1 replica");
    if (e->pend.active) { rc = b200sat_engine_finish(e); if (rc && rc != B200SAT_E_MEM) return rc; }
    CK(cudaSetDevice(e->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (e->n_rings == 0) { /* single-GPU portfolio without an explicit ring: make a private one */
        if ((rc = b200sat_engine_ring_create(e, 1u << 20, 1, 0))) return rc;
    }
    for (u32 r = 0; r < e->n_rings; r++) if (!e->rings[r]) return set_err(B200SAT_E_ARG, "a ring slot is not mapped");
    std::vector<b200sat_settings> hs(n_replicas);
    for (u32 i = 0; i < n_replicas; i++) b200sat_diversify(base, replica_base + i, &hs[i]);
    if ((rc = ensure(e->d_replica_settings, e->cap_replica_settings, (size_t)n_replicas))) return rc;
    CK(cudaMemcpyAsync(e->d_replica_settings, hs.data(), n_replicas * sizeof(b200sat_settings), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream)); /* hs is a local */
    /* outputs are per replica */
    if ((rc = ensure(e->d_results, e->cap_results, (size_t)std::max(n_replicas, e->n_inst)))) return rc;
    if ((rc = ensure(e->d_models, e->cap_models, (size_t)std::max(n_replicas, e->n_inst) * e->model_stride))) return rc;
    if ((rc = ensure(e->d_work[0], e->cap_work, (size_t)std::max(n_replicas, e->n_inst)))) return rc;
    if ((rc = ensure(e->d_work[1], e->cap_work1, (size_t)std::max(n_replicas, e->n_inst)))) return rc;
    e->n_replicas = n_replicas;
    e->info.launches = 0;
    e->info.last_kernel_ms = 0.f; e->info.search_ms = 0.f; e->info.suspends = 0;
    b200sat_engine::Pending &pd = e->pend;
    pd.st = *base; pd.kind = B200SAT_CORE; pd.flags = flags; pd.stream = stream;
    pd.attempt = 0; pd.n_work = n_replicas; pd.work = nullptr; pd.total_ms = 0.f; pd.search_ms = 0.f; pd.retried = false; pd.batch = false;
    CK(cudaMemsetAsync(e->d_models, 2, (size_t)n_replicas * e->model_stride, stream));
    /* a ring that still holds the clauses / the done flag of an earlier run would cancel every replica at once or
     * feed this instance the learnt clauses of another one: the local ring is cleared here.  Multi-GPU callers clear
     * every ring and synchronise the ranks themselves, then pass B200SAT_PF_KEEP_RING (tools/portfolio_run.py). */
    if (!(share_flags & B200SAT_PF_KEEP_RING))
        CK(cudaMemsetAsync(e->rings[e->my_ring], 0, ((size_t)e->ring_words + RING_PAYLOAD) * 4, stream));
    pd.portfolio = n_replicas; pd.replica_base = replica_base; pd.share = share_flags;
    /* the slab is sized for the one instance; BatchShape of a 1-instance batch is that instance */
    pd.tier_nv = pick_tier_nv(e->shape.max_vars, e->shape.max_clauses);
    rc = launch_attempt(e);
    if (rc) return rc;
    pd.active = true;
    /* no larger-slab retry ladder in portfolio mode: a replica that runs out of memory just drops out */
    CK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->info.last_kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, e->ev2, e->ev3));
    e->info.search_ms = ms;
    e->info.suspends = e->h_pinned[32 + QC_SUSPENDS];
    pd.active = false;
    return 0;
}

extern "C" int b200sat_engine_download_replicas(b200sat_engine *e, b200sat_result *results, uint8_t *models, uint32_t model_stride,
                                                void *stream_) {
    if (!e) return set_err(B200SAT_E_ARG, "null engine");
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaSetDevice(e->device));
    if (e->n_replicas == 0) return 0;
    if (results) CK(cudaMemcpyAsync(results, e->d_results, (size_t)e->n_replicas * sizeof(b200sat_result), cudaMemcpyDeviceToHost, stream));
    if (models) CK(cudaMemcpy2DAsync(models, model_stride, e->d_models, e->model_stride, std::min(model_stride, e->model_stride),
                                     e->n_replicas, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

/* ------------------------------------------------------------------ host: single-solver mirror of `trait Solver` */
struct b200sat_solver {
    b200sat_settings st;
    int kind;
    bool ok = true, spent = false, preprocessed = false, resident = false;
    u32 n_vars = 0;
    std::vector<u8> upol;      /* 0 none, 2 false, 3 true */
    std::vector<u8> dvar;
    std::vector<u32> clause_off{0};
    std::vector<u32> lits;
    std::vector<u32> clause_nvars;   /* n_vars at each add_clause */
    size_t n_clauses_known = 0;
    b200sat_stats stats;
    double progress = 0.0;           /* of the last SolveRes::Interrupted */
    b200sat_engine *eng = nullptr;
};

extern "C" b200sat_solver *b200sat_new(const b200sat_settings *s, int kind) {
    if (check_settings(s, kind)) return nullptr;
    b200sat_solver *h = new b200sat_solver();
    h->st = *s;
    h->kind = kind;
    memset(&h->stats, 0, sizeof h->stats);
    return h;
}
extern "C" void b200sat_free(b200sat_solver *h) {
    if (!h) return;
    if (h->eng) b200sat_engine_destroy(h->eng);
    delete h;
}
extern "C" size_t b200sat_n_vars(const b200sat_solver *h) { return h ? h->n_vars : 0; }
extern "C" size_t b200sat_n_clauses(const b200sat_solver *h) { return h ? h->n_clauses_known : 0; }
extern "C" uint32_t b200sat_new_var(b200sat_solver *h, int upol, int dvar) {
    if (!h) return 0xFFFFFFFFu;
    h->upol.push_back(upol < 0 ? 0 : (upol ? 3 : 2));
    h->dvar.push_back(dvar ? 1 : 0);
    return h->n_vars++;
}
extern "C" int b200sat_add_clause(b200sat_solver *h, const uint32_t *lits, size_t n) {
    if (!h || h->spent) return 0;
    for (size_t i = 0; i < n; i++)
        if ((lits[i] >> 1) >= h->n_vars) { set_err(B200SAT_E_ARG, "literal of an unknown variable"); return 0; }
    if (h->kind == B200SAT_CORE && !h->ok) return 0; /* minisat.rs:41-48 */
    /* host-visible part of Searcher::add_clause (search.rs:348-367): duplicate literals and tautologies */
    std::vector<u32> ps(lits, lits + n);
    std::sort(ps.begin(), ps.end());
    ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
    bool taut = false;
    for (size_t i = 1; i < ps.size(); i++) if (ps[i] == (ps[i - 1] ^ 1)) taut = true;
    h->lits.insert(h->lits.end(), lits, lits + n);
    h->clause_off.push_back((u32)h->lits.size());
    h->clause_nvars.push_back(h->n_vars);
    h->resident = false; /* the next solve ingests the whole clause list again */
    if (!taut) {
        if (ps.empty()) { h->ok = false; return 0; }
        if (ps.size() > 1) h->n_clauses_known++;
    }
    return 1;
}

/* wait for the pending launch of the handle's engine, forwarding the caller's interrupt word (Budget::asynch_interrupt)
 * to the device while the kernel runs */
static int solver_wait(b200sat_solver *h, const b200sat_budget *b) {
    b200sat_engine *e = h->eng;
    if (b && b->interrupt) {
        bool raised = false;
        while (cudaEventQuery(e->ev1) == cudaErrorNotReady) {
            if (!raised && *b->interrupt) { b200sat_engine_interrupt(e, 1); raised = true; }
            struct timespec ts = {0, 200000};
            nanosleep(&ts, nullptr);
        }
        cudaGetLastError();
        int rc = b200sat_engine_finish(e);
        if (raised) b200sat_engine_interrupt(e, 0);
        return rc;
    }
    return b200sat_engine_finish(e);
}

/* ingest (+ preprocess) on the device; with `search` the search kernel follows in the same call */
static int solver_ingest(b200sat_solver *h, int flags, const b200sat_budget *b, b200sat_result *res) {
    if (!h->eng) {
        h->eng = b200sat_engine_create(0);
        if (!h->eng) return B200SAT_E_NO_DEVICE;
    }
    u32 ioff[2] = {0, (u32)(h->clause_off.size() - 1)};
    b200sat_batch bt;
    bt.n_inst = 1; bt.inst_clause_off = ioff; bt.clause_off = h->clause_off.data();
    u32 dummy = 0;
    bt.lits = h->lits.empty() ? &dummy : h->lits.data();
    int rc = b200sat_engine_upload(h->eng, &bt, nullptr);
    if (rc) return rc;
    /* Solver::new_var(upol, dvar): one record per created var, also for vars no clause mentions */
    std::vector<u8> vf(h->n_vars ? h->n_vars : 1, 0);
    for (u32 v = 0; v < h->n_vars; v++)
        vf[v] = (u8)((h->upol[v] ? (VARF_UPOL_SET | (h->upol[v] == 3 ? VARF_UPOL_VAL : 0)) : 0) | (h->dvar[v] ? 0 : VARF_NOT_DECISION));
    u32 voff[2] = {0, h->n_vars};
    if ((rc = b200sat_engine_set_var_flags(h->eng, vf.data(), voff))) return rc;
    if ((rc = b200sat_engine_set_clause_nvars(h->eng, h->clause_nvars.empty() ? nullptr : h->clause_nvars.data()))) return rc;
    if ((rc = b200sat_engine_set_budget(h->eng, b ? b->conflict_budget : -1, b ? b->propagation_budget : -1))) return rc;
    if ((rc = b200sat_engine_solve_async(h->eng, &h->st, h->kind, flags, nullptr))) return rc;
    if ((rc = solver_wait(h, b))) return rc;
    return b200sat_engine_download(h->eng, res, nullptr, 0, nullptr);
}

extern "C" int b200sat_preprocess(b200sat_solver *h, const b200sat_budget *b) {
    if (!h) return set_err(B200SAT_E_ARG, "null handle");
    if (h->spent) return set_err(B200SAT_E_SPENT, "handle already consumed");
    if (!h->ok) return 0;
    if (b && b->interrupt && *b->interrupt) return set_err(B200SAT_E_ARG, "interrupt is already raised");
    b200sat_result r;
    int rc = solver_ingest(h, B200SAT_F_PREPROCESS | B200SAT_F_NO_SOLVE, b, &r);
    if (rc) return rc;
    if (r.status != B200SAT_OK) return set_err(B200SAT_E_MEM, "instance outgrew the largest slab");
    h->preprocessed = true;
    h->resident = true;      /* solve_limited() continues from this state: nothing is ingested twice */
    h->n_clauses_known = r.n_clauses_pre;
    memcpy(&h->stats, r.stats, sizeof h->stats);
    if (!r.pre_ok) h->ok = false;
    return h->ok ? 1 : 0;
}

extern "C" int b200sat_solve_limited(b200sat_solver *h, const b200sat_budget *b, const uint32_t *assumps, size_t n_assumps, uint8_t *model,
                                     size_t model_cap) {
    if (!h) return set_err(B200SAT_E_ARG, "null handle");
    if (h->spent) return set_err(B200SAT_E_SPENT, "handle already consumed");
    if (n_assumps && !assumps) return set_err(B200SAT_E_ARG, "null assumptions");
    for (size_t i = 0; i < n_assumps; i++)
        if ((assumps[i] >> 1) >= h->n_vars) return set_err(B200SAT_E_ARG, "assumption on an unknown variable");
    if (b && b->interrupt && *b->interrupt) return B200SAT_INTERRUPTED;
    if (!h->ok && h->preprocessed) { h->spent = true; return B200SAT_UNSAT; } /* minisat.rs:80-82: stats stay as they are */
    b200sat_result r;
    int rc;
    if (!h->eng) {
        h->eng = b200sat_engine_create(0);
        if (!h->eng) return B200SAT_E_NO_DEVICE;
    }
    if ((rc = b200sat_engine_set_assumptions(h->eng, assumps, (uint32_t)n_assumps))) return rc;
    if (h->resident) { /* state parked by preprocess() or by an Interrupted solve_limited(): search only */
        if ((rc = b200sat_engine_set_budget(h->eng, b ? b->conflict_budget : -1, b ? b->propagation_budget : -1))) return rc;
        if ((rc = b200sat_engine_resume_async(h->eng, nullptr))) return rc;
        if ((rc = solver_wait(h, b))) return rc;
        rc = b200sat_engine_download(h->eng, &r, nullptr, 0, nullptr);
    } else rc = solver_ingest(h, h->preprocessed ? B200SAT_F_PREPROCESS : 0, b, &r);
    b200sat_engine_set_assumptions(h->eng, nullptr, 0);
    if (rc) return rc;
    if (r.status != B200SAT_OK) return set_err(B200SAT_E_MEM, "instance outgrew the largest slab");
    memcpy(&h->stats, r.stats, sizeof h->stats);
    h->n_clauses_known = r.n_clauses_pre;
    if (r.verdict == B200SAT_INTERRUPTED) { /* SolveRes::Interrupted(progress, solver): the handle stays usable */
        h->resident = true;
        h->progress = r.progress;
        return B200SAT_INTERRUPTED;
    }
    h->spent = true;
    if (r.verdict == B200SAT_SAT && model) {
        std::vector<u8> m(h->eng->model_stride ? h->eng->model_stride : 1, 2);
        if ((rc = b200sat_engine_download(h->eng, nullptr, m.data(), (u32)m.size(), nullptr))) return rc;
        size_t n = std::min(model_cap, (size_t)h->n_vars);
        for (size_t v = 0; v < n; v++) model[v] = v < m.size() ? m[v] : 2;
    }
    return r.verdict;
}

extern "C" void b200sat_get_stats(const b200sat_solver *h, b200sat_stats *out) {
    if (h && out) *out = h->stats;
}
extern "C" double b200sat_progress_estimate(const b200sat_solver *h) { return h ? h->progress : 0.0; }
extern "C" void b200sat_abi_sizes(uint32_t out[4]) {
    out[0] = (uint32_t)sizeof(b200sat_result); out[1] = (uint32_t)sizeof(b200sat_settings);
    out[2] = (uint32_t)sizeof(b200sat_stats); out[3] = (uint32_t)sizeof(b200sat_launch_info);
}
