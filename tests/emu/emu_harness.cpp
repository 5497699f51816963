/*
 * emu_harness.cpp -- TEST-ONLY: runs the device solver source under the
 * single-warp SIMT emulator (simt_emu.h) on the CPU.  Built by
 * tests/emu/Makefile into tests/emu/libemu.so and loaded only by tests.
 */
#define B200SAT_EMU 1
#include <string.h>
#include <vector>
#include "simt_emu.h"
#include "../../minisat-rust_b200/csrc/cdcl_plan.h"
#include "../../minisat-rust_b200/csrc/cdcl_warp.cuh"
#ifdef B200SAT_HAVE_SIMP
#include "../../minisat-rust_b200/csrc/simp_warp.cuh"
#endif

namespace simt {
thread_local Warp *g_warp_tls = nullptr;
Warp *g_warp = nullptr;

static void trampoline(int lane) {
    Warp *w = g_warp;
    w->entry(w->user, lane);
    w->done[lane] = true;
    w->kind[lane] = K_EXIT;
    swapcontext(&w->lanes[lane], &w->sched);
}

static inline uint64_t next_rng(uint64_t &s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void run_warp(void (*entry)(void *, int), void *user, uint64_t seed) {
    static const size_t STACK = 512 * 1024;
    Warp *w = new Warp();
    memset(w->kind, 0, sizeof w->kind);
    w->entry = entry;
    w->user = user;
    w->rng = seed;
    w->n_collectives = 0;
    g_warp = w;
    for (int l = 0; l < 32; l++) {
        w->stacks[l] = (char *)malloc(STACK);
        w->done[l] = false;
        getcontext(&w->lanes[l]);
        w->lanes[l].uc_stack.ss_sp = w->stacks[l];
        w->lanes[l].uc_stack.ss_size = STACK;
        w->lanes[l].uc_link = &w->sched;
        makecontext(&w->lanes[l], (void (*)())trampoline, 1, l);
    }
    int order[32];
    for (;;) {
        for (int i = 0; i < 32; i++) order[i] = i;
        if (seed != 0) {
            uint64_t mode = next_rng(w->rng) % 3;
            if (mode == 0) for (int i = 31; i > 0; i--) { int j = (int)(next_rng(w->rng) % (uint64_t)(i + 1)); int t = order[i]; order[i] = order[j]; order[j] = t; }
            else if (mode == 1) for (int i = 0; i < 32; i++) order[i] = 31 - i;
        }
        for (int i = 0; i < 32; i++) {
            int l = order[i];
            if (w->done[l]) continue;
            w->cur = l;
            swapcontext(&w->sched, &w->lanes[l]);
        }
        int k0 = w->kind[0];
        for (int l = 1; l < 32; l++)
            if (w->kind[l] != k0) {
                fprintf(stderr, "simt_emu: DIVERGENCE at collective %llu: lane0 kind %d, lane %d kind %d\n",
                        (unsigned long long)w->n_collectives, k0, l, w->kind[l]);
                abort();
            }
        if (k0 == K_EXIT) break;
        w->n_collectives++;
        switch (k0) {
        case K_BALLOT: {
            uint32_t m = 0;
            for (int l = 0; l < 32; l++) if (w->arg[l]) m |= 1u << l;
            for (int l = 0; l < 32; l++) w->res[l] = m;
            break;
        }
        case K_SHFL:
            for (int l = 0; l < 32; l++) w->res[l] = w->arg[w->arg2[l]];
            break;
        case K_SHFL_XOR:
            for (int l = 0; l < 32; l++) w->res[l] = w->arg[(l ^ w->arg2[l]) & 31];
            break;
        case K_SHFL_UP:
            for (int l = 0; l < 32; l++) { int s = l - w->arg2[l]; w->res[l] = w->arg[s < 0 ? l : s]; }
            break;
        case K_MATCH:
            for (int l = 0; l < 32; l++) { uint32_t m = 0; for (int o = 0; o < 32; o++) if (w->arg[o] == w->arg[l]) m |= 1u << o; w->res[l] = m; }
            break;
        case K_SYNC: break;
        default: fprintf(stderr, "simt_emu: bad collective %d\n", k0); abort();
        }
    }
    for (int l = 0; l < 32; l++) free(w->stacks[l]);
    g_warp = nullptr;
    delete w;
}
} // namespace simt

using namespace b200sat;

namespace {
struct Job {
    KParams P;
    void *smem;
    int tier_nv;
    u32 inst;       /* work index: instance id, or replica id in portfolio mode */
    bool count;
    bool replica = false;
    int phase = 0;  /* 0: ingest kernel body, 1: one search-kernel pick-up (load, run a slice, park or finish) */
    u32 runnable = 0, suspended = 0;
};

template <class WS> void bind_job(WS &S, const Job *j) {
    S.slab = j->P.slabs;
    S.res_index = j->inst;
    S.stp = j->replica ? j->P.replica_settings + j->inst : &j->P.st;
}

/* the body of ingest_loop / search_loop_ring (cdcl_kernels.cuh) for ONE work item, without the device queue */
template <bool SIMP, class WS> void ingest_one(WS &S, Job *j) {
    bind_job(S, j);
    const u32 inst = j->replica ? 0u : j->inst;
    bool r;
    if constexpr (SIMP) {
#ifdef B200SAT_HAVE_SIMP
        r = ingest_simp(S, inst);
#else
        r = false;
#endif
    } else r = S.ingest_core(inst);
    if (S.lane == 0) j->runnable = (r && !(j->P.flags & B200SAT_F_NO_SOLVE)) ? 1 : 0;
}
template <class WS> void search_one(WS &S, Job *j) {
    bind_job(S, j);
    S.load_state();
    const u32 r = S.run_search();
    if (r == WS::SR_SUSPEND) {
        S.save_state();
        if (S.lane == 0) j->suspended = 1;
        return;
    }
    S.finish_search(r, j->P.kind == B200SAT_SIMP);
    if (S.lane == 0) j->suspended = 0;
}

/* the body of replay_loop (cdcl_kernels.cuh) for one parked item */
template <class WS> void replay_one(WS &S, Job *j) {
    bind_job(S, j);
    S.load_state();
    const bool ok = S.bcp_replay(S.res_index);
    if (S.lane == 0) {
        b200sat_result *r = j->P.results + S.res_index;
        r->status = ok ? B200SAT_OK : -(i32)(100 + ST_INTERNAL);
        r->stats[5] = S.sc().propagations;
        for (int i = 0; i < 8; i++) r->traffic[i] = S.sc().traffic[i];
    }
}

template <int NV, bool COUNT, bool PF> void entry_s(void *user, int lane) {
    Job *j = (Job *)user;
    if (j->phase == 0) {
        if (j->P.kind == B200SAT_SIMP) { /* SimpSolver ingest: elimination heap + occurrence counts in shared memory */
            WarpSolver<VarsS<NV, true>, COUNT> S(j->P, (u32)lane);
            S.V.s = (SmemVars<NV, true> *)j->smem;
            ingest_one<true>(S, j);
        } else {
            WarpSolver<VarsS<NV>, COUNT> S(j->P, (u32)lane);
            S.V.s = (SmemVars<NV> *)j->smem;
            ingest_one<false>(S, j);
        }
        return;
    }
    WarpSolver<VarsS<NV>, COUNT, PF> S(j->P, (u32)lane);
    S.V.s = (SmemVars<NV> *)j->smem;
    if (j->phase == 2) { replay_one(S, j); return; }
    search_one(S, j);
}
template <bool COUNT, bool PF> void entry_g(void *user, int lane) {
    Job *j = (Job *)user;
    if (j->phase == 0) {
        WarpSolver<VarsG, COUNT> S(j->P, (u32)lane);
        S.V.s = (SmemSmall *)j->smem;
        S.V.bind(j->P.slabs + j->P.L.off_vars, j->P.L.nv_cap);
        if (j->P.kind == B200SAT_SIMP) ingest_one<true>(S, j); else ingest_one<false>(S, j);
        return;
    }
    WarpSolver<VarsG, COUNT, PF> S(j->P, (u32)lane);
    S.V.s = (SmemSmall *)j->smem;
    S.V.bind(j->P.slabs + j->P.L.off_vars, j->P.L.nv_cap);
    if (j->phase == 2) { replay_one(S, j); return; }
    search_one(S, j);
}

u32 save_bytes_of(int tier_nv) {
    switch (tier_nv) {
    case 128: return SmemSave<128>::BYTES;
    case 256: return SmemSave<256>::BYTES;
    case 1024: return SmemSave<1024>::BYTES;
    default: return (u32)sizeof(SmemSmall);
    }
}
template <bool PF> void (*pick_entry(int tier_nv, bool count))(void *, int) {
    switch (tier_nv) {
    case 128: return count ? entry_s<128, true, PF> : entry_s<128, false, PF>;
    case 256: return count ? entry_s<256, true, PF> : entry_s<256, false, PF>;
    case 1024: return count ? entry_s<1024, true, PF> : entry_s<1024, false, PF>;
    default: return count ? entry_g<true, PF> : entry_g<false, PF>;
    }
}

/* One work item through both phases.  The shared-memory block is wiped between pick-ups (a different warp resumes
 * the item on the GPU), so everything a resumed search needs has to come out of the slab. */
template <bool PF> void run_item(Job &j, uint64_t sched_seed, u32 *n_suspends) {
    void (*fn)(void *, int) = pick_entry<PF>(j.tier_nv, j.count);
    j.phase = 0; j.runnable = 0;
    simt::run_warp(fn, &j, sched_seed);
    u32 k = 0;
    while (j.runnable) {
        memset(j.smem, 0xA5, 32 * 1024 * 8);
        j.phase = 1; j.suspended = 0;
        simt::run_warp(fn, &j, sched_seed ? sched_seed + 77 * (++k) : 0);
        if (!j.suspended) break;
        if (n_suspends) (*n_suspends)++;
    }
}
} // namespace

extern "C" {

/* Portfolio mode under the emulator: the replicas of ONE instance run one after another (a replica sees the
 * clauses earlier replicas exported to the shared ring).  share: PF_SHARE | PF_IGNORE_DONE bits. */
int emu_portfolio(const uint32_t *clause_off, const uint32_t *lits, uint32_t n_clauses, const b200sat_settings *base,
                  uint32_t n_replicas, uint32_t share, int flags, uint64_t sched_seed, b200sat_result *results,
                  uint8_t *models, uint32_t model_stride, uint32_t import_cap) {
    BatchShape sh = {0, n_clauses, clause_off[n_clauses], 0};
    for (u32 c = 0; c < n_clauses; c++) { u32 len = clause_off[c + 1] - clause_off[c]; if (len > sh.max_clause_len) sh.max_clause_len = len; }
    for (u32 k = 0; k < clause_off[n_clauses]; k++) if ((lits[k] >> 1) + 1 > sh.max_vars) sh.max_vars = (lits[k] >> 1) + 1;
    int tier_nv = pick_tier_nv(sh.max_vars, sh.max_clauses);
    Layout L = plan_layout(sh, tier_nv, B200SAT_CORE, 0, 0, 0, save_bytes_of(tier_nv));
    std::vector<u64> slab((L.slab_bytes + 7) / 8 + 2);
    std::vector<u64> smem(32 * 1024);
    const u32 ring_words = 1u << 18;
    std::vector<u32> ring(ring_words + RING_PAYLOAD, 0);
    std::vector<b200sat_settings> rs(n_replicas);
    for (u32 i = 0; i < n_replicas; i++) diversify_settings(base, i, &rs[i]);
    u32 ioff[2] = {0, n_clauses};
    u32 queue = 0;
    Job j;
    memset(&j.P, 0, sizeof j.P);
    j.P.inst_clause_off = ioff; j.P.clause_off = clause_off; j.P.lits = lits;
    j.P.n_work = n_replicas; j.P.queue = &queue; j.P.results = results; j.P.models = models; j.P.model_stride = model_stride;
    j.P.slabs = (u8 *)slab.data(); j.P.L = L; j.P.st = *base; j.P.kind = B200SAT_CORE; j.P.flags = flags; j.P.attempt = 1;
    j.P.portfolio = n_replicas; j.P.replica_base = 0; j.P.share = share; j.P.n_rings = 1; j.P.my_ring = 0;
    j.P.ring_words = ring_words; j.P.rings[0] = ring.data(); j.P.replica_settings = rs.data();
    j.P.conflict_budget = -1; j.P.propagation_budget = -1; j.P.import_cap = import_cap;
    j.smem = smem.data(); j.tier_nv = tier_nv; j.count = false;
    for (u32 r = 0; r < n_replicas; r++) {
        j.inst = r; /* replica id: instance 0 with the replica's diversified settings */
        j.replica = true;
        run_item<true>(j, sched_seed ? sched_seed + r : 0, nullptr);
    }
    return 0;
}

/* optional inputs of the following emu_solve_batch calls: new_var(upol, dvar) records and assumptions */
static const uint8_t *g_var_flags = nullptr;
static const uint32_t *g_inst_var_off = nullptr, *g_assumps = nullptr, *g_clause_nvars = nullptr;
static uint32_t g_n_assumps = 0;
void emu_set_extras(const uint8_t *var_flags, const uint32_t *inst_var_off, const uint32_t *assumps, uint32_t n_assumps) {
    g_var_flags = var_flags; g_inst_var_off = inst_var_off; g_assumps = assumps; g_n_assumps = n_assumps;
}
void emu_set_clause_nvars(const uint32_t *clause_nvars) { g_clause_nvars = clause_nvars; }

/* tier_nv: 128/256/1024 = shared-memory tier, 0 = global tier, -1 = auto.
 * sched_seed: 0 = lanes run in ascending order, else adversarial ordering.
 * slice_conflicts: 0 = an instance runs to the end in one pick-up; else it is parked and resumed (with wiped shared
 * memory) every time it has used that many conflicts -- the result must not depend on it.
 * conflict_budget / propagation_budget: budget.rs:5-34 (< 0 = none); an instance that runs out returns
 * INTERRUPTED and, when resume_rounds > 0, is searched again (a new solve_limited call) that many times without a budget.
 * Returns the number of park/resume cycles, or -1 when `kind` is not compiled in. */
int emu_solve_batch(const uint32_t *inst_clause_off, const uint32_t *clause_off, const uint32_t *lits,
                    uint32_t n_inst, const b200sat_settings *st, int kind, int flags, int tier_nv,
                    uint64_t arena_words, uint64_t wpool_entries, uint64_t sched_seed,
                    b200sat_result *results, uint8_t *models, uint32_t model_stride,
                    uint32_t slice_conflicts, int64_t conflict_budget, int64_t propagation_budget, int resume_rounds) {
#ifndef B200SAT_HAVE_SIMP
    if (kind == B200SAT_SIMP) return -1;
#endif
    BatchShape sh = {0, 0, 0, 0};
    for (u32 i = 0; i < n_inst; i++) {
        u32 cb = inst_clause_off[i], ce = inst_clause_off[i + 1];
        u32 nc = ce - cb;
        u64 nl = clause_off[ce] - clause_off[cb];
        if (nc > sh.max_clauses) sh.max_clauses = nc;
        if (nl > sh.max_lits) sh.max_lits = nl;
        for (u32 c = cb; c < ce; c++) {
            u32 len = clause_off[c + 1] - clause_off[c];
            if (len > sh.max_clause_len) sh.max_clause_len = len;
        }
        for (u32 k = clause_off[cb]; k < clause_off[ce]; k++) if ((lits[k] >> 1) + 1 > sh.max_vars) sh.max_vars = (lits[k] >> 1) + 1;
        if (g_var_flags && g_inst_var_off[i + 1] - g_inst_var_off[i] > sh.max_vars) sh.max_vars = g_inst_var_off[i + 1] - g_inst_var_off[i];
    }
    if (tier_nv < 0) tier_nv = pick_tier_nv(sh.max_vars, sh.max_clauses);
    std::vector<u32> todo(n_inst);
    for (u32 i = 0; i < n_inst; i++) todo[i] = i;
    u32 n_suspends = 0;
    u32 qctl[QC_WORDS];
    memset(qctl, 0, sizeof qctl);
    qctl[QC_TAIL] = 1; /* "somebody is always waiting": every expired slice parks the instance */
    for (u32 attempt = 0; attempt < 4 && !todo.empty(); attempt++) {
        if (attempt == 3) tier_nv = 0;
        Layout L = plan_layout(sh, tier_nv, kind, attempt, arena_words, wpool_entries, save_bytes_of(tier_nv));
        std::vector<u64> slab((L.slab_bytes + 7) / 8 + 2);
        std::vector<u64> smem(32 * 1024);
        u32 queue = 0;
        Job j;
        memset(&j.P, 0, sizeof j.P);
        j.P.inst_clause_off = inst_clause_off; j.P.clause_off = clause_off; j.P.lits = lits;
        j.P.work = nullptr; j.P.n_work = n_inst; j.P.queue = &queue;
        j.P.results = results; j.P.models = models; j.P.model_stride = model_stride;
        j.P.slabs = (u8 *)slab.data(); j.P.L = L; j.P.st = *st; j.P.kind = kind; j.P.flags = flags;
        j.P.attempt = attempt + 1;
        j.P.qctl = qctl; j.P.slice_conflicts = slice_conflicts;
        j.P.var_flags = g_var_flags; j.P.inst_var_off = g_inst_var_off; j.P.clause_nvars = g_var_flags ? g_clause_nvars : nullptr; j.P.assumps = g_assumps; j.P.n_assumps = g_n_assumps;
        j.smem = smem.data(); j.tier_nv = tier_nv;
        j.count = (flags & B200SAT_F_COUNT_TRAFFIC) != 0;
        std::vector<u32> again;
        for (u32 inst : todo) {
            j.inst = inst;
            j.P.conflict_budget = conflict_budget; j.P.propagation_budget = propagation_budget;
            const uint64_t seed = sched_seed ? sched_seed + inst : 0;
            run_item<false>(j, seed, &n_suspends);
            for (int r = 0; r < resume_rounds; r++) {
                if (results[inst].status != B200SAT_OK || results[inst].verdict != B200SAT_INTERRUPTED) break;
                if (flags & B200SAT_F_NO_SOLVE) break;
                j.P.conflict_budget = -1; j.P.propagation_budget = -1;
                j.runnable = 1; /* the interrupted search left the instance parked (PH_READY) */
                void (*fn)(void *, int) = pick_entry<false>(j.tier_nv, j.count);
                u32 k = 0;
                while (j.runnable) {
                    memset(j.smem, 0x5A, 32 * 1024 * 8);
                    j.phase = 1; j.suspended = 0;
                    simt::run_warp(fn, &j, seed ? seed + 991 * (++k) : 0);
                    if (!j.suspended) break;
                    n_suspends++;
                }
            }
            if (results[inst].status < 0) again.push_back(inst);
        }
        todo.swap(again);
    }
    return (int)n_suspends;
}

/* K2 under the emulator: record the search prefix of every instance (instrumented kernels), ingest again, replay.
 * out[i] = {recorded words, replay ok, recorded propagations, replayed propagations, recorded BCP traffic equal}. */
int emu_bcp_replay(const uint32_t *inst_clause_off, const uint32_t *clause_off, const uint32_t *lits, uint32_t n_inst,
                   const b200sat_settings *st, int tier_nv, uint32_t trace_cap, uint64_t sched_seed, uint64_t *out) {
    BatchShape sh = {0, 0, 0, 0};
    for (u32 i = 0; i < n_inst; i++) {
        u32 cb = inst_clause_off[i], ce = inst_clause_off[i + 1];
        if (ce - cb > sh.max_clauses) sh.max_clauses = ce - cb;
        u64 nl = clause_off[ce] - clause_off[cb];
        if (nl > sh.max_lits) sh.max_lits = nl;
        for (u32 c = cb; c < ce; c++) { u32 len = clause_off[c + 1] - clause_off[c]; if (len > sh.max_clause_len) sh.max_clause_len = len; }
        for (u32 k = clause_off[cb]; k < clause_off[ce]; k++) if ((lits[k] >> 1) + 1 > sh.max_vars) sh.max_vars = (lits[k] >> 1) + 1;
    }
    if (tier_nv < 0) tier_nv = pick_tier_nv(sh.max_vars, sh.max_clauses);
    Layout L = plan_layout(sh, tier_nv, B200SAT_CORE, 1, 0, 0, save_bytes_of(tier_nv));
    std::vector<u64> slab((L.slab_bytes + 7) / 8 + 2), smem(32 * 1024);
    std::vector<u32> trace((size_t)n_inst * trace_cap, 0);
    std::vector<b200sat_result> res(n_inst);
    u32 queue = 0;
    Job j;
    memset(&j.P, 0, sizeof j.P);
    j.P.inst_clause_off = inst_clause_off; j.P.clause_off = clause_off; j.P.lits = lits;
    j.P.n_work = n_inst; j.P.queue = &queue; j.P.results = res.data();
    j.P.slabs = (u8 *)slab.data(); j.P.L = L; j.P.st = *st; j.P.kind = B200SAT_CORE; j.P.attempt = 2;
    j.P.conflict_budget = -1; j.P.propagation_budget = -1;
    j.P.trace_buf = trace.data(); j.P.trace_cap = trace_cap;
    j.smem = smem.data(); j.tier_nv = tier_nv; j.count = true;
    void (*fn)(void *, int) = pick_entry<false>(tier_nv, true);
    for (u32 i = 0; i < n_inst; i++) {
        j.inst = i;
        j.P.flags = B200SAT_F_PREPROCESS | B200SAT_F_COUNT_TRAFFIC;
        run_item<false>(j, sched_seed ? sched_seed + i : 0, nullptr);        /* records */
        const u32 *tb = trace.data() + (size_t)i * trace_cap;
        out[5 * i + 0] = tb[TR_WORDS];
        out[5 * i + 2] = (u64)tb[TR_PROPS_LO] | ((u64)tb[TR_PROPS_HI] << 32);
        out[5 * i + 1] = 0; out[5 * i + 3] = 0; out[5 * i + 4] = 0;
        if (tb[TR_WORDS] == 0) continue;
        j.P.flags = B200SAT_F_PREPROCESS | B200SAT_F_COUNT_TRAFFIC | B200SAT_F_NO_SOLVE;
        j.phase = 0; j.runnable = 0;
        simt::run_warp(fn, &j, sched_seed ? sched_seed + 5 * i : 0);          /* ingest again: parked, not searched */
        j.phase = 2;
        simt::run_warp(fn, &j, sched_seed ? sched_seed + 7 * i : 0);          /* replay */
        out[5 * i + 1] = res[i].status == B200SAT_OK ? 1 : 0;
        out[5 * i + 3] = res[i].stats[5];
        bool same = true;
        for (int k = 0; k < 6; k++) {
            const u64 wt = (u64)tb[TR_TRAFFIC + 2 * k] | ((u64)tb[TR_TRAFFIC + 2 * k + 1] << 32);
            if (res[i].traffic[k] != wt) same = false;
        }
        out[5 * i + 4] = same ? 1 : 0;
    }
    return 0;
}
}
