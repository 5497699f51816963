/*
 * cdcl_warp.cuh -- warp-per-instance CDCL search for sm_100a.
 *
 * One warp owns one SAT instance.  Control flow is warp-uniform; the 32 lanes
 * cooperate inside the data-parallel steps (watch-list scanning with ordered
 * commit, clause loads for conflict analysis and minimisation, trail undo,
 * reduceDB sort, garbage collection) and lane 0 runs the strictly sequential
 * data-structure updates (VSIDS heap, activity bumps, list heads).
 *
 * The ALGORITHM follows mishun/minisat-rust so that decisions / conflicts /
 * propagations are bit-exact against it.  Reference
 * anchors (paths relative to /root/reference/src/sat/minisat):
 *   propagate()        search/watches.rs:64-152
 *   analyze()          search/conflict.rs:70-176
 *   lit_redundant()    search/conflict.rs:195-249
 *   cancel_until()     search.rs:253-261 + formula/assignment.rs:116-130
 *   decide()           search.rs:216-237, search/decision_heuristic.rs:158-192
 *   heap_*()           formula/index_map.rs:183-309
 *   reduce_db()        search/clause_db.rs:156-215
 *   remove_satisfied() search/clause_db.rs:238-254,283-303
 *   garbage_collect()  formula/clause.rs:192-223, watches.rs:154-169, clause_db.rs:256-280
 *   try_simplify()     search.rs:522-568
 *   search_loop()      search.rs:452-517
 *   add_clause()       search.rs:342-386
 * The IMPLEMENTATION (data layout, warp cooperation, memory management) is new.
 *
 * Memory discipline: scalar state in memory is written by lane 0 only; a
 * __syncwarp() separates such writes from reads by other lanes.
 */
#ifndef B200SAT_CDCL_WARP_CUH
#define B200SAT_CDCL_WARP_CUH

#include "cdcl_types.h"

#ifdef B200SAT_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <stddef.h>
#define DEV __device__ __forceinline__
#define DEVNI __device__ __noinline__
#endif

namespace b200sat {

/* ------------------------------------------------------------------ per-warp scalar state (shared memory) */
struct Scal {
    u64 solves, starts, decisions, rnd_decisions, conflicts, propagations, max_literals, tot_literals;
    u64 traffic[8];
    u64 clauses_literals, learnts_literals, lc_size, lc_wasted, simp_db_props, trace_hash;
    double var_inc, cla_inc, max_learnts, size_adjust_confl, seed, inv_var_decay, inv_cla_decay;
    u32 arena_top, wpool_top, n_learnts, n_clauses, num_clauses, num_learnts;
    i32 size_adjust_cnt;
    u32 simp_db_assigns, simp_db_assigns_set, status, dec_vars, heap_n, cur_arena, cur_wpool, eliminated_vars;
    u32 remove_satisfied, extra_clause_field, elim_n;
    u32 ring_cursor[8], exported, imported, simp_cur_opool;
    /* ---- resident-instance state (round 2): what a warp needs to park an instance and any warp to resume it */
    u32 phase;                                   /* PH_* life cycle                                              */
    u32 sv_nvars, sv_trail_n, sv_qhead, sv_nlevels; /* the solver's register state while parked                   */
    u32 curr_restarts;                           /* search.rs:413 loop counter                                   */
    u32 n_ingest, n_pre, pre_ok;                 /* clause counts for the result record                          */
    u32 in_loop;                                 /* 1: parked inside a search_loop (confl_limit valid)           */
    u64 confl_limit;                             /* conflict count that ends the current search_loop             */
    double progress;                             /* progress estimate of the last Interrupted return             */
    u32 rec_n, rec_on;                           /* K2 event record: words used, still recording                 */
};

/* ------------------------------------------------------------------ var-state tiers */
/* SimpSolver kernels keep the elimination heap and the literal occurrence counts (the
 * simplifier's most frequently chased words, elim_queue.rs:13-76) in shared memory too */
template <int NV, bool SIMP> struct SimpSmem {};
template <int NV> struct SimpSmem<NV, true> {
    i32 nocc[2 * NV];
    u16 eheap[NV];
    u16 eidx[NV];
};

template <int NV, bool SIMP = false> struct SmemVars {
    double act[NV];
    u32 reason[NV];
    u32 wstart[2 * NV];
    u16 wlen[2 * NV];
    u16 level[NV];
    u16 trail[NV];
    u16 lim[NV + 2];
    u16 heap[NV];
    u16 hidx[NV];
    u8 assign[NV];
    u8 vflags[NV];
    Scal sc;
    u32 bl[32];
    SimpSmem<NV, SIMP> simp;
};
/* bytes of a warp's shared-memory block that make up an instance's parked state: everything in front of the
 * simplifier's scratch (which is dead once the ingest phase is over) */
template <int NV> struct SmemSave {
    typedef SmemVars<NV, false> SV_;
    static const u32 BYTES = (u32)((offsetof(SV_, simp) + 7) & ~(size_t)7);
};

template <int NV, bool SIMP = false> struct VarsS {
    typedef u16 idx_t;
    static const u32 ABSENT = 0xFFFFu;
    static const u32 MAXLEN = 0xFFF0u;
    static const bool SIMP_SMEM = SIMP;
    static const bool GLOBAL = false;
    SmemVars<NV, SIMP> *s;
    DEV double *act() const { return s->act; }
    DEV u32 *reason() const { return s->reason; }
    DEV u32 *wstart() const { return s->wstart; }
    DEV u16 *wlen() const { return s->wlen; }
    DEV u16 *level() const { return s->level; }
    DEV u16 *trail() const { return s->trail; }
    DEV u16 *lim() const { return s->lim; }
    DEV u16 *heap() const { return s->heap; }
    DEV u16 *hidx() const { return s->hidx; }
    DEV u8 *assign() const { return s->assign; }
    DEV u8 *vflags() const { return s->vflags; }
    DEV Scal &sc() const { return s->sc; }
    DEV u32 *bl() const { return s->bl; }
    DEV u32 cap() const { return NV; }
    DEV void *simp_smem() const { return (void *)&s->simp; }
};

struct SmemSmall {
    Scal sc;
    u32 bl[32];
};

struct VarsG {
    typedef u32 idx_t;
    static const u32 ABSENT = 0xFFFFFFFFu;
    static const u32 MAXLEN = 0x3FFFFFF0u;
    static const bool SIMP_SMEM = false;
    static const bool GLOBAL = true;
    DEV void *simp_smem() const { return nullptr; }
    SmemSmall *s;
    double *act_;
    u32 *reason_, *wstart_, *wlen_, *level_, *trail_, *lim_, *heap_, *hidx_;
    u8 *assign_, *vflags_;
    u32 cap_;
    DEV double *act() const { return act_; }
    DEV u32 *reason() const { return reason_; }
    DEV u32 *wstart() const { return wstart_; }
    DEV u32 *wlen() const { return wlen_; }
    DEV u32 *level() const { return level_; }
    DEV u32 *trail() const { return trail_; }
    DEV u32 *lim() const { return lim_; }
    DEV u32 *heap() const { return heap_; }
    DEV u32 *hidx() const { return hidx_; }
    DEV u8 *assign() const { return assign_; }
    DEV u8 *vflags() const { return vflags_; }
    DEV Scal &sc() const { return s->sc; }
    DEV u32 *bl() const { return s->bl; }
    DEV u32 cap() const { return cap_; }
    DEV void bind(u8 *base, u32 nv) {
        u64 n = nv, o = 0;
        act_ = (double *)(base + o); o += layout_align(8 * n);
        reason_ = (u32 *)(base + o); o += layout_align(4 * n);
        wstart_ = (u32 *)(base + o); o += layout_align(4 * 2 * n);
        wlen_ = (u32 *)(base + o); o += layout_align(4 * 2 * n);
        level_ = (u32 *)(base + o); o += layout_align(4 * (n + 2));
        trail_ = (u32 *)(base + o); o += layout_align(4 * (n + 2));
        lim_ = (u32 *)(base + o); o += layout_align(4 * (n + 2));
        heap_ = (u32 *)(base + o); o += layout_align(4 * (n + 2));
        hidx_ = (u32 *)(base + o); o += layout_align(4 * (n + 2));
        assign_ = base + o; o += layout_align(n);
        vflags_ = base + o;
        cap_ = nv;
    }
};

/* ------------------------------------------------------------------ small helpers */
DEV u32 ceil_log2(u32 x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }
DEV u32 round4(u32 x) { return (x + 3) & ~3u; }
DEV u32 lanemask_lt(u32 lane) { return (1u << lane) - 1; }
DEV u32 bcast(u32 x) { return __shfl_sync(FULL_MASK, x, 0); }
DEV u32 warp_sum(u32 x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL_MASK, x, o);
    return x;
}
DEV u32 warp_max(u32 x) {
    for (int o = 16; o > 0; o >>= 1) { u32 y = __shfl_xor_sync(FULL_MASK, x, o); x = y > x ? y : x; }
    return x;
}
DEV u32 warp_excl_scan(u32 x, u32 lane, u32 &total) {
    u32 v = x;
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(FULL_MASK, v, o); if (lane >= (u32)o) v += y; }
    total = __shfl_sync(FULL_MASK, v, 31);
    return v - x;
}
DEV uint4 ld4(const u32 *p) { return *reinterpret_cast<const uint4 *>(p); }
#if defined(__CUDACC__) && defined(B200SAT_PREFETCH_REASONS)
DEV void prefetch_l1(const void *p) { asm volatile("prefetch.L1 [%0];" ::"l"(p)); }
#else
DEV void prefetch_l1(const void *) {}
#endif
DEV float u2f(u32 x) { return __uint_as_float(x); }
DEV u32 f2u(float x) { return __float_as_uint(x); }

enum : u32 { OUT_NONE = 0, OUT_KEEPW = 1, OUT_KEEPCW = 2, OUT_MOVE = 3, OUT_UNIT = 4, OUT_CONFL = 5 };
enum : u32 { LR_RESTART = 0, LR_UNSAT = 1, LR_SAT = 2, LR_ABORT = 3 };

/* ================================================================== the solver */
template <class VT, bool COUNT, bool PF = false> struct WarpSolver {
    typedef typename VT::idx_t idx_t;
    static const u32 VT_ABSENT = VT::ABSENT;
    static const u32 VT_MAXLEN = VT::MAXLEN;
    static const bool VT_SIMP_SMEM = VT::SIMP_SMEM;
    static const bool VT_GLOBAL = VT::GLOBAL;
    VT V;
    const KParams &P;
    const b200sat_settings *stp;   /* P.st, or this replica's diversified settings in portfolio mode */
    u32 res_index;                 /* slot in results/models (= instance id, or replica id)          */
    u32 item_index = 0;            /* work index of the item (slab / parked slot)                     */
    u32 *rx_ring = nullptr;        /* straggler rescue: the ring of the instance this replica works on */
    u32 rx_tag = 0;                /* ... its item tag (RING_ITEM word)                                */
    bool rx_main = false;          /* ... true: the original replica (exports only, stays on the trace) */
    u32 result_flags = 0;          /* OR-ed into result.attempts (0x100: decided by a helper replica)  */
    u32 lane;
    u8 *slab;
    u32 *arena;   /* active arena half */
    uint2 *wpool; /* active watcher pool half */
    u32 nvars, trail_n, qhead, nlevels;

    DEV WarpSolver(const KParams &p, u32 ln) : P(p), stp(&p.st), res_index(0), lane(ln) {}
    DEV const b200sat_settings &st() const { return *stp; }

    DEV Scal &sc() const { return V.sc(); }
    DEV u32 *learnts() const { return (u32 *)(slab + P.L.off_learnts); }
    DEV u32 *clauses() const { return (u32 *)(slab + P.L.off_clauses); }
    DEV u32 *ol() const { return (u32 *)(slab + P.L.off_ol); }
    DEV u32 *tc() const { return (u32 *)(slab + P.L.off_tc); }
    DEV u32 *stk() const { return (u32 *)(slab + P.L.off_stk); }
    DEV u32 *tmp() const { return (u32 *)(slab + P.L.off_tmp); }
    DEV u32 *arena_half(u32 h) const { return (u32 *)(slab + P.L.off_arena[h]); }
    DEV uint2 *wpool_half(u32 h) const { return (uint2 *)(slab + P.L.off_wpool[h]); }

    DEV u32 lval(u32 lit) const { return (u32)V.assign()[lit >> 1] ^ (lit & 1); }
    DEV bool is_true(u32 lit) const { return lval(lit) == 1; }
    DEV bool is_false(u32 lit) const { return lval(lit) == 0; }
    DEV bool failed() const { return sc().status != ST_OK; }
    DEV void fail(u32 code) {
        __syncwarp(); /* lagging lanes may still be reading status in failed() */
        if (lane == 0 && sc().status == ST_OK) sc().status = code;
        __syncwarp();
    }
    /* reference byte accounting, clause.rs:226-251 (LegacyCounter) */
    DEV static u32 ref_size(u32 hdr) {
        u32 len = hdr >> H_SHIFT;
        return (hdr & (H_LEARNT | H_EXTRA)) ? 4 * (2 + len) : 4 * (1 + len);
    }
    DEV static u32 block_words(u32 hdr) { return round4(1 + (hdr >> H_SHIFT) + ((hdr & H_EXTRA) ? 1 : 0)); }

    /* -------------------------------------------------------------- VSIDS heap (lane 0 only), index_map.rs:183-309 */
    DEV void heap_sift_up(u32 i) {
        idx_t *heap = V.heap(); idx_t *hidx = V.hidx(); const double *act = V.act();
        u32 x = heap[i];
        double ax = act[x];
        while (i > 0) {
            u32 p = (i - 1) >> 1;
            u32 y = heap[p];
            if (ax > act[y]) { heap[i] = (idx_t)y; hidx[y] = (idx_t)i; i = p; } else break;
        }
        heap[i] = (idx_t)x; hidx[x] = (idx_t)i;
    }
    DEV void heap_sift_down(u32 i) {
        idx_t *heap = V.heap(); idx_t *hidx = V.hidx(); const double *act = V.act();
        u32 n = sc().heap_n;
        u32 x = heap[i];
        double ax = act[x];
        for (;;) {
            u32 l = 2 * i + 1;
            if (l >= n) break;
            u32 r = l + 1;
            u32 yl = heap[l];
            double al = act[yl];
            u32 c = l, yc = yl;
            double ac = al;
            if (r < n) {
                u32 yr = heap[r];
                double ar = act[yr];
                if (ar > al) { c = r; yc = yr; ac = ar; }
            }
            if (ac > ax) { heap[i] = (idx_t)yc; hidx[yc] = (idx_t)i; i = c; } else break;
        }
        heap[i] = (idx_t)x; hidx[x] = (idx_t)i;
    }
    DEV void heap_insert(u32 v) {
        if (V.hidx()[v] != (idx_t)VT::ABSENT) return;
        u32 place = sc().heap_n;
        sc().heap_n = place + 1;
        V.heap()[place] = (idx_t)v;
        heap_sift_up(place);
    }
    DEV u32 heap_pop() {
        idx_t *heap = V.heap();
        u32 x = heap[0];
        u32 n = sc().heap_n - 1;
        sc().heap_n = n;
        u32 last = heap[n];
        V.hidx()[x] = (idx_t)VT::ABSENT;
        if (n > 0) { heap[0] = (idx_t)last; heap_sift_down(0); }
        return x;
    }
    /* decision_heuristic.rs:126-140 (lane 0).  sift_down of IdxHeap::update is a
     * no-op after an activity increase (children <= old key <= new key), so only
     * the sift_up half is executed. */
    DEV void bump_var(u32 v) {
        double *act = V.act();
        double nw = act[v] + sc().var_inc;
        if (nw > 1e100) {
            sc().var_inc *= 1e-100;
            for (u32 i = 0; i < nvars; i++) act[i] *= 1e-100;
            act[v] = nw * 1e-100;
        } else act[v] = nw;
        u32 place = V.hidx()[v];
        if (place != VT::ABSENT) heap_sift_up(place);
    }
    /* random.rs:12-26 (lane 0) */
    DEV double drand() {
        double seed = sc().seed;
        seed *= 1389796.0;
        i32 q = (i32)(seed / 2147483647.0);
        seed -= (double)q * 2147483647.0;
        sc().seed = seed;
        return seed / 2147483647.0;
    }

    /* -------------------------------------------------------------- trail (assignment.rs:99-130) */
    DEV void assign_lit(u32 l, u32 reason) {
        if (lane == 0) {
            u32 v = l >> 1;
            V.assign()[v] = (u8)((l & 1) ^ 1);
            V.level()[v] = (idx_t)nlevels;
            V.reason()[v] = reason;
            V.trail()[trail_n] = (idx_t)l;
        }
        trail_n++;
        __syncwarp();
    }
    DEV void new_decision_level() {
        if (lane == 0) V.lim()[nlevels] = (idx_t)trail_n;
        nlevels++;
        __syncwarp();
    }

    /* -------------------------------------------------------------- watch lists */
    /* relocate list `l` to the pool end with doubled capacity; false on failure */
    DEV bool grow_list(u32 l) {
        u32 word = V.wstart()[l];
        u32 len = V.wlen()[l];
        u32 capc = word & W_CAP_MASK;
        u32 ncap = 2u << capc;
        if (ncap > VT::MAXLEN) { fail(ST_MEM_LIST); return false; }
        if (sc().wpool_top + ncap > P.L.wpool_cap) {
            if (!compact_wpool(ncap + 64)) return false;
            word = V.wstart()[l];
        }
        u32 top = sc().wpool_top;
        uint2 *src = wpool + (word >> W_SHIFT);
        uint2 *dst = wpool + top;
        for (u32 i = lane; i < len; i += 32) dst[i] = src[i];
        __syncwarp();
        if (lane == 0) {
            V.wstart()[l] = (top << W_SHIFT) | (word & W_DIRTY) | (capc + 1);
            sc().wpool_top = top + ncap;
        }
        __syncwarp();
        return true;
    }
    /* copy every list into the other pool half with fresh capacities */
    DEV bool compact_wpool(u32 reserve) {
        u32 nl = 2 * nvars;
        u32 h = sc().cur_wpool ^ 1;
        uint2 *dst = wpool_half(h);
        u32 top = 0;
        for (u32 l0 = 0; l0 < nl; l0 += 32) {
            u32 l = l0 + lane;
            u32 word = 0, len = 0, ncapc = 0, ncap = 0;
            if (l < nl) {
                word = V.wstart()[l]; len = V.wlen()[l];
                u32 want = len + (len >> 1) + 2;
                ncapc = ceil_log2(want < 4 ? 4 : want);
                ncap = 1u << ncapc;
            }
            u32 tot;
            u32 off = top + warp_excl_scan(ncap, lane, tot);
            if (top + tot + reserve > P.L.wpool_cap) { fail(ST_MEM_WPOOL); return false; }
            if (l < nl) {
                const uint2 *src = wpool + (word >> W_SHIFT);
                for (u32 i = 0; i < len; i++) dst[off + i] = src[i];
                V.wstart()[l] = (off << W_SHIFT) | (word & W_DIRTY) | ncapc;
            }
            top += tot;
        }
        __syncwarp();
        if (lane == 0) { sc().wpool_top = top; sc().cur_wpool = h; }
        wpool = dst;
        __syncwarp();
        return true;
    }
    DEV bool push_watch(u32 l, u32 cref, u32 blocker) {
        u32 word = V.wstart()[l];
        u32 len = V.wlen()[l];
        if (len == (1u << (word & W_CAP_MASK))) {
            __syncwarp();
            if (!grow_list(l)) return false;
            word = V.wstart()[l];
        }
        __syncwarp();
        if (lane == 0) {
            wpool[(word >> W_SHIFT) + len] = make_uint2(cref, blocker);
            V.wlen()[l] = (idx_t)(len + 1);
        }
        __syncwarp();
        return true;
    }
    /* watches.rs:39-44 */
    DEV bool attach(u32 cref) {
        u32 c0 = arena[cref + 1], c1 = arena[cref + 2];
        if (!push_watch(c0 ^ 1, cref, c1)) return false;
        return push_watch(c1 ^ 1, cref, c0);
    }

    /* -------------------------------------------------------------- BCP, watches.rs:64-152
     * Lanes evaluate up to 32 watchers of watches[p] at once against the current
     * assignment.  A watcher's verdict can only be changed by an implication made
     * by an EARLIER watcher of the same list, and only when the implied variable
     * occurs among the literals the verdict depends on (blocker, first literal,
     * chosen replacement watch).  So the lanes are committed in list order up to
     * the first lane that such an implication touches (or a conflict); that lane
     * and the ones behind it are re-evaluated (their watcher and clause words are
     * already in registers).  Several implications commit in one round. */

    /* ordered parallel append: lane i (if `moving`) appends (cref, blocker) to list tgt; lanes
     * with the same target keep their lane order.  One step per distinct target list. */
    DEV bool push_watch_multi(u32 movemask, bool moving, u32 tgt, u32 cref, u32 blocker) {
        const u32 lt = lanemask_lt(lane);
        while (movemask) {
            const u32 l = (u32)__ffs(movemask) - 1;
            const u32 t = __shfl_sync(FULL_MASK, tgt, l);
            const bool mine = moving && tgt == t;
            const u32 same = __ballot_sync(FULL_MASK, mine);
            const u32 cnt = __popc(same);
            movemask &= ~same;
            u32 word = V.wstart()[t];
            const u32 len = V.wlen()[t];
            if (len + cnt > (1u << (word & W_CAP_MASK))) { /* the list has to grow: one by one */
                u32 mm = same;
                while (mm) {
                    u32 ml = (u32)__ffs(mm) - 1;
                    mm &= mm - 1;
                    u32 c = __shfl_sync(FULL_MASK, cref, ml);
                    u32 bk = __shfl_sync(FULL_MASK, blocker, ml);
                    if (!push_watch(t, c, bk)) return false;
                }
                continue;
            }
            __syncwarp();
            if (mine) wpool[(word >> W_SHIFT) + len + __popc(same & lt)] = make_uint2(cref, blocker);
            if (lane == l) V.wlen()[t] = (idx_t)(len + cnt);
        }
        __syncwarp();
        return true;
    }

    DEV u32 propagate() {
        u32 confl = CREF_UNDEF;
        u32 nprops = 0;
        u32 t_vis = 0, t_kept = 0, t_deref = 0, t_tail = 0, t_moves = 0, t_impl = 0;
        const u32 lt = lanemask_lt(lane);
        bool have_pf = false;
        u32 pf_word = 0, pf_n = 0;
        uint2 pf_w = make_uint2(0, 0);
        while (qhead < trail_n && confl == CREF_UNDEF) {
            u32 p = V.trail()[qhead];
            qhead++;
            nprops++;
            const u32 not_p = p ^ 1;
            u32 word, n;
            uint2 w0 = make_uint2(0, 0);
            if (have_pf) { word = pf_word; n = pf_n; w0 = pf_w; }
            else {
                word = V.wstart()[p];
                n = V.wlen()[p];
                if (lane < n) w0 = wpool[(word >> W_SHIFT) + lane];
            }
            /* software pipelining: fetch the head of the NEXT literal's list now.  That list cannot
             * change while this one is processed (a true literal is never a move target). */
#ifdef B200SAT_PREFETCH /* measured slower on B200 (r01: 248 vs 208 ms at 2 warps/SM): off by default */
            have_pf = qhead < trail_n;
#else
            have_pf = false;
#endif
            if (have_pf) {
                const u32 p2 = V.trail()[qhead];
                pf_word = V.wstart()[p2];
                pf_n = V.wlen()[p2];
                pf_w = make_uint2(0, 0);
                if (lane < pf_n) pf_w = wpool[(pf_word >> W_SHIFT) + lane];
            }
            const bool dirty = (word & W_DIRTY) != 0;
            uint2 *ws = wpool + (word >> W_SHIFT);
            u32 j = 0;
            for (u32 i = 0; i < n; i += 32) {
                u32 idx = i + lane;
                bool valid = idx < n;
                uint2 w = w0;
                if (i != 0) { w = make_uint2(0, 0); if (valid) w = ws[idx]; }
                uint4 hd = make_uint4(0, 0, 0, 0);
                bool loaded = false, dead = false;
                if (valid && dirty) { hd = ld4(arena + w.x); loaded = true; dead = (hd.x & H_DELETED) != 0; }
                const bool live = valid && !dead;
                if (confl != CREF_UNDEF) {
                    /* conflict already found: copy the remainder unchanged (watches.rs:123-137) */
                    u32 m = __ballot_sync(FULL_MASK, live);
                    if (live) ws[j + __popc(m & lt)] = w;
                    j += __popc(m);
                    continue;
                }
                u32 kf = 2, lbase = 0;
                for (;;) {
                    const bool pend = live && lane >= lbase;
                    u32 out = OUT_NONE, first = 0, lk = 0, len = 0;
                    if (pend) {
                        if (is_true(w.y)) out = OUT_KEEPW;
                        else {
                            if (!loaded) { hd = ld4(arena + w.x); loaded = true; }
                            first = (hd.y == not_p) ? hd.z : hd.y;
                            const u32 fv = lval(first);
                            if (fv == 1) out = OUT_KEEPCW;
                            else {
                                len = hd.x >> H_SHIFT;
                                out = (fv == 0) ? OUT_CONFL : OUT_UNIT;
                                /* tail scan, 128 bits (4 literals) per load; literal k lives in word k+1 of the block */
                                if (kf == 2 && len > 2) {
                                    if (!is_false(hd.w)) { lk = hd.w; out = OUT_MOVE; } else kf = 3;
                                }
                                while (out != OUT_MOVE && kf < len) {
                                    const u32 gb = (kf + 1) & ~3u;
                                    const uint4 q = ld4(arena + w.x + gb);
                                    const u32 qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                    for (u32 t = 0; t < 4; t++) {
                                        const u32 k2 = gb + t - 1;
                                        if (out != OUT_MOVE && k2 >= kf && k2 < len && !is_false(qs[t])) { lk = qs[t]; kf = k2; out = OUT_MOVE; }
                                    }
                                    if (out != OUT_MOVE) kf = gb + 3;
                                }
                            }
                        }
                    }
                    const u32 pendmask = __ballot_sync(FULL_MASK, pend);
                    if (!pendmask) break;
                    u32 umask = __ballot_sync(FULL_MASK, out >= OUT_UNIT);
                    u32 stop = 32;     /* first lane that is NOT committed this round */
                    u32 cfl_lane = 32; /* lane whose clause is conflicting, if inside the committed prefix */
                    while (umask) {
                        const u32 u = (u32)__ffs(umask) - 1;
                        if (u >= stop) break;
                        umask &= umask - 1;
                        const u32 uout = __shfl_sync(FULL_MASK, out, u);
                        if (uout == OUT_CONFL) { cfl_lane = u; stop = u + 1; break; }
                        const u32 xv = __shfl_sync(FULL_MASK, first, u) >> 1;
                        const bool aff = pend && lane > u && out != OUT_KEEPW &&
                                         ((w.y >> 1) == xv || (first >> 1) == xv || (out == OUT_MOVE && (lk >> 1) == xv));
                        const u32 am = __ballot_sync(FULL_MASK, aff);
                        if (am) { const u32 s2 = (u32)__ffs(am) - 1; if (s2 < stop) stop = s2; }
                    }
                    const bool commit = pend && lane < stop;
                    const bool keep = commit && out != OUT_MOVE;
                    const u32 keepmask = __ballot_sync(FULL_MASK, keep);
                    if (keep) ws[j + __popc(keepmask & lt)] = (out == OUT_KEEPW) ? w : make_uint2(w.x, first);
                    j += __popc(keepmask);
                    const bool mv = commit && out == OUT_MOVE;
                    if (commit && out != OUT_KEEPW) {
                        if (hd.y == not_p || mv) {
                            arena[w.x + 1] = first;
                            arena[w.x + 2] = mv ? lk : not_p;
                        }
                        if (mv) arena[w.x + 1 + kf] = not_p;
                    }
                    /* implications of the committed prefix, in lane order (assignment.rs:104-113) */
                    const bool imp = commit && out == OUT_UNIT;
                    const u32 impmask = __ballot_sync(FULL_MASK, imp);
                    if (imp) {
                        const u32 v = first >> 1;
                        V.assign()[v] = (u8)((first & 1) ^ 1);
                        V.level()[v] = (idx_t)nlevels;
                        V.reason()[v] = w.x;
                        V.trail()[trail_n + __popc(impmask & lt)] = (idx_t)first;
                    }
                    trail_n += __popc(impmask);
                    const u32 movemask = __ballot_sync(FULL_MASK, mv);
                    if (COUNT) {
                        t_vis += __popc(__ballot_sync(FULL_MASK, commit));
                        t_kept += __popc(keepmask);
                        t_deref += __popc(__ballot_sync(FULL_MASK, commit && out != OUT_KEEPW));
                        u32 tl = 0;
                        if (commit && out >= OUT_MOVE) tl = mv ? kf - 1 : len - 2;
                        t_tail += warp_sum(tl);
                        t_moves += __popc(movemask);
                        t_impl += __popc(impmask);
                    }
                    if (movemask) {
                        if (!push_watch_multi(movemask, mv, lk ^ 1, w.x, first)) return CREF_UNDEF;
                        if (wpool + (V.wstart()[p] >> W_SHIFT) != ws) { /* the pool was compacted: descriptors moved */
                            ws = wpool + (V.wstart()[p] >> W_SHIFT);
                            have_pf = false;
                        }
                    }
                    __syncwarp();
                    if (cfl_lane != 32) {
                        confl = __shfl_sync(FULL_MASK, w.x, cfl_lane);
                        const bool rem = live && lane >= stop;
                        u32 m = __ballot_sync(FULL_MASK, rem);
                        if (rem) ws[j + __popc(m & lt)] = w;
                        j += __popc(m);
                        break;
                    }
                    if (stop == 32) break;
                    lbase = stop;
                }
            }
            __syncwarp();
            if (lane == 0) {
                V.wlen()[p] = (idx_t)j;
                if (dirty) V.wstart()[p] = V.wstart()[p] & ~W_DIRTY;
            }
            __syncwarp();
        }
        if (lane == 0) {
            sc().propagations += nprops;
            if (COUNT) {
                u64 *t = sc().traffic;
                t[0] += t_vis; t[1] += t_kept; t[2] += t_deref; t[3] += t_tail; t[4] += t_moves; t[5] += t_impl;
            }
        }
        __syncwarp();
        return confl;
    }

    /* -------------------------------------------------------------- clause arena */
    /* returns cref or CREF_UNDEF; lits are written by the caller */
    DEV u32 alloc_clause(u32 len, u32 flags) {
        u32 hdr = (len << H_SHIFT) | flags;
        u32 blk = block_words(hdr);
        u32 top = sc().arena_top;
        if (top + blk > P.L.arena_cap) { fail(ST_MEM_ARENA); return CREF_UNDEF; }
        __syncwarp();
        if (lane == 0) {
            arena[top] = hdr;
            sc().arena_top = top + blk;
            sc().lc_size += ref_size(hdr);
        }
        __syncwarp();
        return top;
    }
    /* clause_db.rs:119-143 */
    DEV void bump_cla(u32 cr, u32 len) {
        u32 resc = 0;
        if (lane == 0) {
            double nw = (double)u2f(arena[cr + 1 + len]) + sc().cla_inc;
            arena[cr + 1 + len] = f2u((float)nw);
            if (nw > 1e20) { sc().cla_inc *= 1e-20; resc = 1; }
        }
        resc = bcast(resc);
        if (resc) {
            __syncwarp();
            u32 nl = sc().n_learnts;
            const u32 *ls = learnts();
            for (u32 i = lane; i < nl; i += 32) {
                u32 c = ls[i];
                u32 l = arena[c] >> H_SHIFT;
                arena[c + 1 + l] = f2u((float)((double)u2f(arena[c + 1 + l]) * 1e-20));
            }
        }
        __syncwarp();
    }

    /* -------------------------------------------------------------- conflict analysis, conflict.rs:70-249 */
    /* iterative redundancy check; frames are (lit, cref, pos, len); the scan of one
     * reason clause is lane-parallel, the first non-trivial literal in clause order
     * decides what happens next, exactly as the sequential walk would. */
    DEV bool lit_redundant(u32 literal, u32 &tcn, u64 &an_words) {
        u32 fcr = V.reason()[literal >> 1];
        if (fcr == CREF_UNDEF) return false;
        u32 *st = stk();
        u32 *tcl = tc();
        const u32 amax = P.L.arena_cap - 1;
        u32 sp = 0;
        u32 fp = literal, fpos = 1, flen = 0;
        bool fresh_frame = true; /* header not read yet: fetch it together with the first literals */
        for (;;) {
            bool pushed = false;
            for (;;) {
                /* words [wb, wb+32) of the clause block; word 0 is the header */
                const u32 wb = fresh_frame ? 0 : fpos + 1;
                const u32 wi = wb + lane;
                u32 x = arena[(fcr + wi) < amax ? (fcr + wi) : amax];
                if (fresh_frame) {
                    flen = bcast(x) >> H_SHIFT;
                    if (COUNT) an_words += 1 + flen;
                    fresh_frame = false;
                }
                if (fpos >= flen) break;
                const u32 k = wi - 1;
                u32 code = 0;
                if (wi >= 1 && k >= fpos && k < flen) {
                    u32 v = x >> 1;
                    u32 s = V.vflags()[v] & VF_SEEN;
                    if (!(V.level()[v] == GROUND || s == SEEN_SOURCE || s == SEEN_REMOVABLE))
                        code = (V.reason()[v] != CREF_UNDEF && s == SEEN_UNDEF) ? 1 : 2;
                }
                u32 nm = __ballot_sync(FULL_MASK, code != 0);
                if (!nm) { fpos = wb + 31; if (fpos >= flen) break; continue; }
                u32 f = (u32)__ffs(nm) - 1;
                u32 fc = __shfl_sync(FULL_MASK, code, f);
                u32 fl = __shfl_sync(FULL_MASK, x, f);
                if (fc == 2) {
                    /* failure: every literal on the stack is not redundant either (conflict.rs:232-240) */
                    __syncwarp();
                    if (lane == 0) {
                        u8 *vf = V.vflags();
                        for (u32 t = 0; t <= sp; t++) {
                            u32 q = (t == sp) ? fp : st[4 * t];
                            if ((vf[q >> 1] & VF_SEEN) == SEEN_UNDEF) { vf[q >> 1] |= SEEN_FAILED; tcl[tcn++] = q; }
                        }
                    }
                    tcn = bcast(tcn);
                    __syncwarp();
                    return false;
                }
                if (lane == 0) { st[4 * sp] = fp; st[4 * sp + 1] = fcr; st[4 * sp + 2] = wb + f; st[4 * sp + 3] = flen; }
                sp++;
                fp = fl;
                fcr = V.reason()[fl >> 1];
                fpos = 1;
                fresh_frame = true;
                pushed = true;
                break;
            }
            if (pushed) continue;
            /* frame exhausted: fp is implied by redundant literals (conflict.rs:241-246) */
            __syncwarp();
            if (lane == 0) {
                u8 *vf = V.vflags();
                if ((vf[fp >> 1] & VF_SEEN) == SEEN_UNDEF) { vf[fp >> 1] |= SEEN_REMOVABLE; tcl[tcn++] = fp; }
            }
            tcn = bcast(tcn);
            __syncwarp();
            if (sp == 0) return true;
            sp--;
            fp = st[4 * sp]; fcr = st[4 * sp + 1]; fpos = st[4 * sp + 2]; flen = st[4 * sp + 3];
        }
    }
    DEV bool lit_redundant_basic(u32 literal, u64 &an_words) { /* conflict.rs:178-192 */
        u32 cr = V.reason()[literal >> 1];
        if (cr == CREF_UNDEF) return false;
        u32 len = arena[cr] >> H_SHIFT;
        if (COUNT) an_words += 1 + len;
        bool bad = false;
        for (u32 k = 1 + lane; k < len; k += 32) {
            u32 v = arena[cr + 1 + k] >> 1;
            if ((V.vflags()[v] & VF_SEEN) == SEEN_UNDEF && V.level()[v] > GROUND) bad = true;
        }
        return __ballot_sync(FULL_MASK, bad) == 0;
    }

    /* returns number of literals of the learnt clause (in ol()), sets bt */
    DEV u32 analyze(u32 confl0, u32 &bt) {
        u32 *out = ol();
        u32 *tcl = tc();
        u8 *vf = V.vflags();
        const u32 lt = lanemask_lt(lane);
        const u32 cur = nlevels;
        i32 path_c = 0;
        u32 index = trail_n;
        u32 n_out = 1;
        u32 confl = confl0;
        u64 an_words = 0;
        const u32 amax = P.L.arena_cap - 1;
        for (;;) {
            /* one round trip: lane 0 reads the header, lanes 1..31 the first 31 literals */
            u32 x = arena[(confl + lane) < amax ? (confl + lane) : amax];
            const u32 h = bcast(x);
            const u32 len = h >> H_SHIFT;
            if (h & H_LEARNT) bump_cla(confl, len);
            if (COUNT) an_words += 1 + len;
            const u32 kbase = (confl == confl0) ? 0 : 1;
            for (u32 w0 = 0; w0 <= len; w0 += 32) {
                const u32 wi = w0 + lane;          /* word index inside the clause block, 0 = header */
                const u32 k = wi - 1;              /* literal index */
                if (w0 != 0) x = (wi <= len) ? arena[confl + wi] : 0;
                bool fresh = false, iscur = false;
                u32 q = 0, v = 0;
                if (wi >= 1 && k < len && k >= kbase) {
                    q = x;
                    v = q >> 1;
                    u32 f = vf[v];
                    if ((f & VF_SEEN) == SEEN_UNDEF) {
                        u32 lv = V.level()[v];
                        if (lv > GROUND) {
                            fresh = true; iscur = lv >= cur; vf[v] = (u8)(f | SEEN_SOURCE);
                            /* its reason clause is read soon (current level: by this walk, lower levels: by minimisation) */
                            const u32 rr = V.reason()[v];
                            if (rr != CREF_UNDEF) prefetch_l1(arena + rr);
                        }
                    }
                }
                u32 fm = __ballot_sync(FULL_MASK, fresh);
                u32 cm = __ballot_sync(FULL_MASK, fresh && iscur);
                path_c += __popc(cm);
                u32 lowm = fm & ~cm;
                if (fresh && !iscur) out[n_out + __popc(lowm & lt)] = q;
                n_out += __popc(lowm);
                if (fm) {
                    u32 *bl = V.bl();
                    if (fresh) bl[__popc(fm & lt)] = v;
                    __syncwarp();
                    if (lane == 0) {
                        u32 cnt = __popc(fm);
                        for (u32 t = 0; t < cnt; t++) bump_var(bl[t]);
                    }
                    __syncwarp();
                }
            }
            /* next literal to expand: last seen literal on the trail (conflict.rs:112-121) */
            __syncwarp();
            for (;;) {
                i32 idx = (i32)index - 1 - (i32)lane;
                bool s = idx >= 0 && (vf[V.trail()[idx] >> 1] & VF_SEEN) != 0;
                u32 m = __ballot_sync(FULL_MASK, s);
                if (m) { index = index - (u32)__ffs(m); break; }
                if (index <= 32) { fail(ST_INTERNAL); return 0; }
                index -= 32;
            }
            u32 pl = V.trail()[index];
            __syncwarp();
            if (lane == 0) vf[pl >> 1] &= (u8)~VF_SEEN;
            __syncwarp();
            path_c -= 1;
            if (path_c <= 0) {
                if (lane == 0) out[0] = pl ^ 1;
                break;
            }
            confl = V.reason()[pl >> 1];
        }
        __syncwarp();
        /* minimisation (conflict.rs:139-152); retain() visits position 0 too */
        u32 tcn = 0;
        u32 n_min = n_out;
        const i32 ccmin = st().ccmin_mode;
        if (ccmin != 0) {
            u32 j = 0;
            for (u32 i = 0; i < n_out; i++) {
                u32 l = out[i];
                bool red = (ccmin == 2) ? lit_redundant(l, tcn, an_words) : lit_redundant_basic(l, an_words);
                if (lane == 0) {
                    if (red) tcl[tcn] = l; else out[j] = l;
                }
                if (red) tcn++; else j++;
                __syncwarp();
            }
            n_min = j;
        }
        /* clear seen[] (conflict.rs:154-156) */
        for (u32 i = lane; i < n_min; i += 32) vf[out[i] >> 1] &= (u8)~VF_SEEN;
        for (u32 i = lane; i < tcn; i += 32) vf[tcl[i] >> 1] &= (u8)~VF_SEEN;
        /* backtrack level: first literal of the highest level among positions >= 1 (conflict.rs:158-175) */
        if (n_min == 1) bt = GROUND;
        else {
            u32 best_lv = 0, best_i = 0xFFFFFFFFu;
            for (u32 i = 1 + lane; i < n_min; i += 32) {
                u32 lv = V.level()[out[i] >> 1];
                if (lv > best_lv) { best_lv = lv; best_i = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                u32 olv = __shfl_xor_sync(FULL_MASK, best_lv, o);
                u32 oi = __shfl_xor_sync(FULL_MASK, best_i, o);
                if (olv > best_lv || (olv == best_lv && oi < best_i)) { best_lv = olv; best_i = oi; }
            }
            if (lane == 0 && best_i != 1) { u32 t = out[1]; out[1] = out[best_i]; out[best_i] = t; }
            bt = best_lv;
        }
        __syncwarp();
        if (lane == 0) {
            sc().max_literals += n_out;
            sc().tot_literals += n_min;
            if (COUNT) sc().traffic[6] += an_words;
            if (P.flags & B200SAT_F_TRACE_HASH) {
                u64 hsh = sc().trace_hash;
                for (u32 i = 0; i < n_min + 2; i++) {
                    u32 x = (i == 0) ? bt : (i == 1) ? n_min : out[i - 2];
                    for (int b = 0; b < 4; b++) { hsh ^= (x >> (8 * b)) & 0xff; hsh *= 1099511628211ull; }
                }
                sc().trace_hash = hsh;
            }
        }
        __syncwarp();
        return n_min;
    }

    /* -------------------------------------------------------------- backtracking: search.rs:253-261 + assignment.rs:116-130 */
    DEV void cancel_until(u32 target_level) {
        if (nlevels > target_level) {
            const u32 lt = lanemask_lt(lane);
            const u32 target = V.lim()[target_level];
            const u32 top = nlevels;
            const i32 ps = st().phase_saving;
            u8 *vf = V.vflags();
            for (i32 hi = (i32)trail_n; hi > (i32)target; hi -= 32) {
                i32 k = hi - 1 - (i32)lane;
                bool ins = false;
                u32 v = 0;
                if (k >= (i32)target) {
                    u32 l = V.trail()[k];
                    v = l >> 1;
                    u32 f = vf[v];
                    if (ps == 2 || (ps == 1 && V.level()[v] == top)) f = (f & ~VF_POL) | ((l & 1) ? VF_POL : 0);
                    vf[v] = (u8)f;
                    V.assign()[v] = LB_UNDEF;
                    ins = (f & VF_DEC) && V.hidx()[v] == (idx_t)VT::ABSENT;
                }
                u32 im = __ballot_sync(FULL_MASK, ins);
                if (im) {
                    u32 *bl = V.bl();
                    if (ins) bl[__popc(im & lt)] = v;
                    __syncwarp();
                    if (lane == 0) {
                        u32 cnt = __popc(im);
                        for (u32 t = 0; t < cnt; t++) heap_insert(bl[t]);
                    }
                    __syncwarp();
                }
            }
            trail_n = target;
            nlevels = target_level;
        }
        if (qhead > trail_n) qhead = trail_n;
        __syncwarp();
    }

    /* -------------------------------------------------------------- decision: search.rs:216-237, decision_heuristic.rs:158-192 */
    DEV u32 pick_branch_lit() {
        u32 next = LIT_NONE;
        if (lane == 0) {
            bool got = false;
            u32 v = 0;
            const bool use_rand = (st().random_var_freq != 0.0) || st().rnd_pol;
            if (use_rand) {
                if (drand() < st().random_var_freq && sc().heap_n > 0) {
                    v = V.heap()[(u32)(drand() * (double)sc().heap_n)];
                    if (V.assign()[v] == LB_UNDEF && (V.vflags()[v] & VF_DEC)) { sc().rnd_decisions++; got = true; }
                }
            }
            if (!got) {
                while (sc().heap_n > 0) {
                    v = heap_pop();
                    if (V.assign()[v] == LB_UNDEF && (V.vflags()[v] & VF_DEC)) { got = true; break; }
                }
            }
            if (got) {
                u32 f = V.vflags()[v];
                u32 sign;
                if (f & VF_UPOL_SET) sign = (f & VF_UPOL_VAL) ? 1 : 0;
                else if (st().rnd_pol) sign = drand() < 0.5 ? 1 : 0;
                else sign = (f & VF_POL) ? 1 : 0;
                next = (v << 1) | sign;
            }
        }
        return bcast(next);
    }

    /* -------------------------------------------------------------- learnt DB */
    /* clause_db.rs:156-215.  The stable sort is done as a bitonic sort of the
     * unique keys (class/activity bits << 32 | list position). */
    DEV void reduce_db() {
        const u32 L = sc().n_learnts;
        if (L == 0) return;
        u32 *ls = learnts();
        u32 *scratch = arena_half(sc().cur_arena ^ 1);
        u32 P2 = 1u << ceil_log2(L < 2 ? 2 : L);
        if ((u64)P2 * 2 + L > P.L.arena_cap) { fail(ST_MEM_ARENA); return; }
        u64 *keys = (u64 *)scratch;
        u32 *oldl = scratch + 2 * (size_t)P2;
        for (u32 i = lane; i < P2; i += 32) {
            u64 k = ~0ull;
            if (i < L) {
                u32 cr = ls[i];
                u32 len = arena[cr] >> H_SHIFT;
                u32 kb = (len == 2) ? 0x7F800000u : arena[cr + 1 + len];
                k = ((u64)kb << 32) | i;
                oldl[i] = cr;
            }
            keys[i] = k;
        }
        __syncwarp();
        for (u32 k = 2; k <= P2; k <<= 1) {
            for (u32 jj = k >> 1; jj > 0; jj >>= 1) {
                for (u32 t = lane; t < (P2 >> 1); t += 32) {
                    u32 i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    u32 ix = i | jj;
                    u64 a = keys[i], b = keys[ix];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ix] = a; }
                }
                __syncwarp();
            }
        }
        for (u32 i = lane; i < L; i += 32) ls[i] = oldl[(u32)(keys[i] & 0xFFFFFFFFu)];
        __syncwarp();
        const u32 index_lim = L / 2;
        const double extra_lim = sc().cla_inc / (double)L;
        const u32 lt = lanemask_lt(lane);
        u32 j = 0, removed = 0, rem_lits = 0, rem_bytes = 0;
        for (u32 i0 = 0; i0 < L; i0 += 32) {
            u32 i = i0 + lane;
            bool keep = false;
            u32 cr = 0, rl = 0, rb = 0;
            if (i < L) {
                cr = ls[i];
                u32 h = arena[cr];
                if (!(h & H_DELETED)) {
                    u32 len = h >> H_SHIFT;
                    u32 c0 = arena[cr + 1];
                    bool locked = is_true(c0) && V.reason()[c0 >> 1] == cr;
                    bool rm = len > 2 && !locked && (i < index_lim || (double)u2f(arena[cr + 1 + len]) < extra_lim);
                    if (rm) {
                        u32 c1 = arena[cr + 2];
                        atomicOr(&V.wstart()[c0 ^ 1], (u32)W_DIRTY);
                        atomicOr(&V.wstart()[c1 ^ 1], (u32)W_DIRTY);
                        arena[cr] = h | H_DELETED;
                        rl = len; rb = ref_size(h);
                    } else keep = true;
                }
            }
            u32 km = __ballot_sync(FULL_MASK, keep);
            if (keep) ls[j + __popc(km & lt)] = cr;
            j += __popc(km);
            removed += __popc(__ballot_sync(FULL_MASK, rl != 0));
            rem_lits += warp_sum(rl);
            rem_bytes += warp_sum(rb);
        }
        __syncwarp();
        if (lane == 0) {
            sc().n_learnts = j;
            sc().num_learnts -= removed;
            sc().learnts_literals -= rem_lits;
            sc().lc_wasted += rem_bytes;
        }
        __syncwarp();
    }

    /* clause_db.rs:217-254,283-303 over one list */
    DEV u32 remove_satisfied_list(u32 *list, u32 n, bool is_learnt) {
        const u32 lt = lanemask_lt(lane);
        u32 j = 0, removed = 0, rem_lits = 0, rem_bytes = 0;
        for (u32 i0 = 0; i0 < n; i0 += 32) {
            u32 i = i0 + lane;
            bool keep = false;
            u32 cr = 0, rl = 0, rb = 0, rmv = 0;
            if (i < n) {
                cr = list[i];
                u32 h = arena[cr];
                if (!(h & H_DELETED)) {
                    u32 len = h >> H_SHIFT;
                    bool sat = false;
                    for (u32 k = 0; k < len; k++) if (is_true(arena[cr + 1 + k])) { sat = true; break; }
                    if (sat) {
                        atomicOr(&V.wstart()[arena[cr + 1] ^ 1], (u32)W_DIRTY);
                        atomicOr(&V.wstart()[arena[cr + 2] ^ 1], (u32)W_DIRTY);
                        arena[cr] = h | H_DELETED;
                        rl = len; rb = ref_size(h); rmv = 1;
                    } else {
                        u32 l = 2, r = len;
                        while (l < r) {
                            if (!is_false(arena[cr + 1 + l])) l++;
                            else { r--; arena[cr + 1 + l] = arena[cr + 1 + r]; }
                        }
                        if (r != len) {
                            if (h & H_EXTRA) arena[cr + 1 + r] = arena[cr + 1 + len];
                            arena[cr] = (r << H_SHIFT) | (h & ((1u << H_SHIFT) - 1));
                        }
                        keep = true;
                    }
                }
            }
            u32 km = __ballot_sync(FULL_MASK, keep);
            if (keep) list[j + __popc(km & lt)] = cr;
            j += __popc(km);
            removed += __popc(__ballot_sync(FULL_MASK, rmv != 0));
            rem_lits += warp_sum(rl);
            rem_bytes += warp_sum(rb);
        }
        __syncwarp();
        if (lane == 0) {
            if (is_learnt) { sc().num_learnts -= removed; sc().learnts_literals -= rem_lits; }
            else { sc().num_clauses -= removed; sc().clauses_literals -= rem_lits; }
            sc().lc_wasted += rem_bytes;
        }
        __syncwarp();
        return j;
    }
    DEV void remove_satisfied() {
        u32 j = remove_satisfied_list(learnts(), sc().n_learnts, true);
        if (lane == 0) sc().n_learnts = j;
        __syncwarp();
        if (sc().remove_satisfied) {
            j = remove_satisfied_list(clauses(), sc().n_clauses, false);
            if (lane == 0) sc().n_clauses = j;
            __syncwarp();
        }
    }

    /* -------------------------------------------------------------- copying GC into the other arena half
     * (clause.rs:136-155,192-223).  One call relocates up to 32 distinct clauses. */
    struct GcState { u32 *dst; u32 top; u64 size; };
    DEV bool gc_relocate(GcState &g, bool valid, u32 cr, u32 &ncr) {
        bool alive = false, need = false;
        u32 h = 0, blk = 0;
        if (valid) {
            h = arena[cr];
            if (!(h & H_DELETED)) {
                alive = true;
                if (h & H_RELOCED) ncr = arena[cr + 1];
                else {
                    need = true;
                    /* ClauseAllocator::reloc drops the signature when extra_clause_field is off (clause.rs:143-149) */
                    if ((h & H_EXTRA) && !(h & H_LEARNT) && !sc().extra_clause_field) h &= ~(u32)H_EXTRA;
                    blk = block_words(h);
                }
            }
        }
        u32 tot;
        u32 off = warp_excl_scan(blk, lane, tot);
        if (need) {
            ncr = g.top + off;
            u32 words = 1 + (h >> H_SHIFT) + ((h & H_EXTRA) ? 1 : 0);
            g.dst[ncr] = h;
            for (u32 k = 1; k < words; k++) g.dst[ncr + k] = arena[cr + k];
            arena[cr + 1] = ncr;
            arena[cr] = arena[cr] | H_RELOCED;
        }
        g.top += tot;
        g.size += warp_sum(need ? ref_size(h) : 0);
        return alive;
    }
    DEV void garbage_collect() {
        const u32 lt = lanemask_lt(lane);
        GcState g;
        u32 hnew = sc().cur_arena ^ 1;
        g.dst = arena_half(hnew);
        g.top = 0;
        g.size = 0;
        /* watches.rs:154-169 */
        for (u32 l = 0; l < 2 * nvars; l++) {
            u32 n = V.wlen()[l];
            uint2 *ws = wpool + (V.wstart()[l] >> W_SHIFT);
            u32 j = 0;
            for (u32 i0 = 0; i0 < n; i0 += 32) {
                u32 i = i0 + lane;
                uint2 w = make_uint2(0, 0);
                if (i < n) w = ws[i];
                u32 ncr = 0;
                bool alive = gc_relocate(g, i < n, w.x, ncr);
                u32 m = __ballot_sync(FULL_MASK, alive);
                if (alive) ws[j + __popc(m & lt)] = make_uint2(ncr, w.y);
                j += __popc(m);
            }
            __syncwarp();
            if (lane == 0) { V.wlen()[l] = (idx_t)j; V.wstart()[l] = V.wstart()[l] & ~W_DIRTY; }
        }
        __syncwarp();
        /* assignment.rs:236-243 */
        for (u32 i0 = 0; i0 < trail_n; i0 += 32) {
            u32 i = i0 + lane;
            u32 v = 0, r = CREF_UNDEF;
            if (i < trail_n) { v = V.trail()[i] >> 1; r = V.reason()[v]; }
            u32 ncr = 0;
            bool has = r != CREF_UNDEF;
            bool alive = gc_relocate(g, has, r, ncr);
            if (has) V.reason()[v] = alive ? ncr : CREF_UNDEF;
        }
        /* clause_db.rs:256-280 */
        for (int pass = 0; pass < 2; pass++) {
            u32 *list = pass == 0 ? learnts() : clauses();
            u32 n = pass == 0 ? sc().n_learnts : sc().n_clauses;
            u32 j = 0;
            for (u32 i0 = 0; i0 < n; i0 += 32) {
                u32 i = i0 + lane;
                u32 cr = (i < n) ? list[i] : 0;
                u32 ncr = 0;
                bool alive = gc_relocate(g, i < n, cr, ncr);
                u32 m = __ballot_sync(FULL_MASK, alive);
                if (alive) list[j + __popc(m & lt)] = ncr;
                j += __popc(m);
            }
            __syncwarp();
            if (lane == 0) { if (pass == 0) sc().n_learnts = j; else sc().n_clauses = j; }
            __syncwarp();
        }
        gc_finish(g, hnew);
    }
    DEV void gc_finish(GcState &g, u32 hnew) {
        __syncwarp();
        if (lane == 0) {
            sc().cur_arena = hnew;
            sc().arena_top = g.top;
            sc().lc_size = g.size;
            sc().lc_wasted = 0;
        }
        arena = g.dst;
        __syncwarp();
    }
    /* search.rs:577-581 (+ a capacity-driven trigger of our own; GC is trace-neutral during search) */
    DEV void try_garbage_collect() {
        bool ref_trigger = (double)sc().lc_wasted > (double)sc().lc_size * st().garbage_frac;
        bool cap_trigger = sc().lc_wasted > 0 && sc().arena_top > (P.L.arena_cap / 4) * 3;
        if (ref_trigger || cap_trigger) garbage_collect();
    }

    /* decision_heuristic.rs:146-156 + index_map.rs:257-264 */
    DEV void rebuild_order_heap() {
        const u32 lt = lanemask_lt(lane);
        u32 cnt = 0;
        for (u32 v0 = 0; v0 < nvars; v0 += 32) {
            u32 v = v0 + lane;
            bool in = v < nvars && (V.vflags()[v] & VF_DEC) && V.assign()[v] == LB_UNDEF;
            u32 m = __ballot_sync(FULL_MASK, in);
            u32 pos = cnt + __popc(m & lt);
            if (v < nvars) V.hidx()[v] = (idx_t)(in ? pos : VT::ABSENT);
            if (in) V.heap()[pos] = (idx_t)v;
            cnt += __popc(m);
        }
        __syncwarp();
        if (lane == 0) {
            sc().heap_n = cnt;
            for (i32 i = (i32)(cnt / 2) - 1; i >= 0; i--) heap_sift_down((u32)i);
        }
        __syncwarp();
    }

    /* search.rs:522-568 */
    DEV void try_simplify() {
        if (nlevels != 1) return;
        if ((sc().simp_db_assigns_set && sc().simp_db_assigns == trail_n) || sc().propagations < sc().simp_db_props) return;
        remove_satisfied();
        try_garbage_collect();
        rebuild_order_heap();
        if (lane == 0) {
            sc().simp_db_assigns_set = 1;
            sc().simp_db_assigns = trail_n;
            sc().simp_db_props = sc().propagations + sc().clauses_literals + sc().learnts_literals;
        }
        __syncwarp();
    }

    /* -------------------------------------------------------------- portfolio: clause exchange + first-finisher cancel
     * (no reference analogue).  Each GPU owns an append-only ring in its HBM:
     * [head][done][...][payload: hdr = 0x80000000 | replica<<8 | len, lits...].  Replicas append short learnt
     * clauses (units, binaries, LBD <= 2) and, at restart boundaries (ground level), import what the other
     * replicas -- on this GPU and, through peer-mapped pointers over NVLink, on the other GPUs -- appended. */
    DEV static u32 ld_vol(const u32 *p) { return *(const volatile u32 *)p; }
    DEV bool cancelled() {
        if (!PF || (P.share & PF_IGNORE_DONE)) return false;
        u32 v = 0; /* one lane reads, all lanes agree: another GPU may raise the flag at any moment */
        if (lane == 0) {
            if (rx_ring) v = (ld_vol(rx_ring + RING_DONE) == rx_tag || ld_vol(rx_ring + RING_ITEM) != rx_tag) ? 1u : 0u;
            else v = ld_vol(P.rings[P.my_ring] + RING_DONE);
        }
        return bcast(v) != 0;
    }
    DEV void announce_done() {
        if (!PF || rx_ring) return; /* rescue mode: the instance is claimed by a CAS on its ring (rescue_claim) */
        if (lane == 0) {
            __threadfence_system();
            for (u32 r = 0; r < P.n_rings; r++) atomicCAS(P.rings[r] + RING_DONE, 0u, 1u + P.replica_base + res_index);
        }
        __syncwarp();
    }
    /* called with the learnt clause in ol()[0..n) BEFORE backtracking (levels still valid) */
    DEV void export_clause(u32 n) {
        /* PF_SHARE alone: units and binaries; PF_SHARE_LBD2 adds LBD <= 2 clauses of at most 8 literals.
         * (Round-1 measurement: exporting every LBD <= 2 clause up to 30 literals floods the importers --
         * 1,024 replicas each took ~80 k clauses -- and a second GPU made the run slower, not faster.) */
        const u32 max_len = (P.share & PF_SHARE_LBD2) ? 8u : 2u;
        if (n > max_len) return;
        const u32 *out = ol();
        u32 lit = lane < n ? out[lane] : 0;
        u32 lbd = 0;
        if (n > 2) {
            u32 lv = lane < n ? (u32)V.level()[lit >> 1] : (0x80000000u | lane);
            u32 peers = __match_any_sync(FULL_MASK, lv);
            lbd = __popc(__ballot_sync(FULL_MASK, lane < n && (u32)__ffs(peers) - 1 == lane));
            if (lbd > 2) return;
        }
        u32 *ring = rx_ring ? rx_ring : P.rings[P.my_ring];
        const u32 ring_words = rx_ring ? P.rr_words : P.ring_words;
        if (rx_ring && rx_main) { /* the original replica feeds its helpers, if it has any */
            u32 nh = 0;
            if (lane == 0) nh = ld_vol(rx_ring + RING_HELPERS);
            if (bcast(nh) == 0) return;
        }
        u32 pos = 0;
        if (lane == 0) pos = atomicAdd(ring + RING_HEAD, n + 1);
        pos = bcast(pos);
        if (pos + n + 1 > ring_words) return; /* ring full: stop exporting */
        u32 *pay = ring + RING_PAYLOAD + pos;
        if (lane < n) pay[1 + lane] = lit;
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            *(volatile u32 *)pay = 0x80000000u | (((rx_ring ? rx_tag : P.replica_base + res_index) & 0x7FFFFFu) << 8) | n;
            sc().exported++;
        }
        __syncwarp();
    }
    /* at ground level with an empty propagation queue; false = the formula is UNSAT.
     * At most P.import_cap clauses are taken per call (the cursors stay where the call stopped, so nothing is lost)
     * and a clause whose order-independent hash this replica has already seen is skipped: round 1 measured that
     * without cap and filter 1,024 replicas each swallowed ~80k mostly duplicate clauses and got slower. */
    DEV bool import_clauses() {
        const u32 me = (P.replica_base + res_index) & 0x7FFFFFu;
        const u32 cap = P.import_cap ? P.import_cap : 0xFFFFFFFFu;
        u32 *filter = (u32 *)(slab + P.L.off_pf);
        u32 got = 0;
        const u32 n_rings = rx_ring ? 1u : P.n_rings;
        const u32 ring_words = rx_ring ? P.rr_words : P.ring_words;
        for (u32 r = 0; r < n_rings && got < cap; r++) {
            const u32 *ring = rx_ring ? rx_ring : P.rings[r];
            u32 cur = sc().ring_cursor[r];
            for (;;) {
                if (got >= cap || cur + 1 > ring_words) break;
                u32 hdr = ld_vol(ring + RING_PAYLOAD + cur);
                if (!(hdr & 0x80000000u)) break; /* not committed yet */
                u32 n = hdr & 0xFF;
                u32 src = (hdr >> 8) & 0x7FFFFFu;
                /* portfolio: everything but my own exports; rescue: only clauses of the instance I work on (the ring is
                 * re-used by the next instance of its warp; the duplicate filter drops my own exports) */
                if (rx_ring ? src == (rx_tag & 0x7FFFFFu) : src != me) {
                    __threadfence();
                    u32 *t = tmp();
                    u32 lit = lane < n ? ld_vol(ring + RING_PAYLOAD + cur + 1 + lane) : 0u;
                    /* order-independent clause hash (the same clause arrives from many replicas, literals in any order) */
                    u32 hm = lane < n ? (lit + 1u) * 0x9E3779B1u : 0u;
                    hm ^= hm >> 15;
                    u32 h = warp_sum(hm) + n * 0x85EBCA6Bu;
                    h ^= h >> 13;
                    const u32 b0 = h & 8191u, b1 = (h >> 13) & 8191u;
                    u32 seen = 0;
                    if (lane == 0) {
                        const u32 w0 = filter[b0 >> 5], w1 = filter[b1 >> 5];
                        seen = ((w0 >> (b0 & 31)) & (w1 >> (b1 & 31)) & 1u);
                        if (!seen) { filter[b0 >> 5] = w0 | (1u << (b0 & 31)); filter[b1 >> 5] |= 1u << (b1 & 31); }
                    }
                    seen = bcast(seen);
                    if (!seen) {
                        if (lane < n) t[lane] = lit;
                        __syncwarp();
                        u32 cr = CREF_UNDEF;
                        u32 res = add_clause(t, n, cr, true);
                        if (failed()) return true;
                        got++;
                        if (res == 0) { /* the imported clause is falsified at ground level: UNSAT */
                            if (lane == 0) sc().imported += got;
                            __syncwarp();
                            return false;
                        }
                    }
                }
                cur += n + 1;
            }
            __syncwarp();
            if (lane == 0) sc().ring_cursor[r] = cur;
            __syncwarp();
        }
        if (lane == 0) sc().imported += got;
        __syncwarp();
        return true;
    }

    DEV static double powi(double b, i32 e) {
        double r = 1.0;
        while (e > 0) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    DEV static double luby(double y, u32 x) { /* luby.rs:2-20 */
        u32 size = 1;
        i32 seq = 0;
        while (size < x + 1) { seq += 1; size = 2 * size + 1; }
        while (size - 1 != x) { size = (size - 1) >> 1; seq -= 1; x = x % size; }
        return powi(y, seq);
    }

    /* ============================================================== round-2 search path
     * One flat, resumable loop (search.rs:409-517 flattened): the instance's whole state lives in its slab plus the
     * shared-memory var block, so a warp can park it at a conflict-free point and any warp can pick it up again.
     * The hot functions below are written for instruction economy (ncu: the round-1 BCP spent ~360 warp
     * instructions per dequeued literal); semantics are those of the functions they replace. */

    /* BCP, watches.rs:64-152 -- same ordered-commit scheme as propagate() (see the comment there), leaner:
     * register copies of the queue cursors, one 32-watcher chunk is the common case, moved watchers are appended
     * with one __match_any_sync instead of a loop over target lists, and list growth / pool compaction is the
     * slow path. */
    DEV u32 propagate2() {
        u32 confl = CREF_UNDEF;
        const u32 lt = lanemask_lt(lane);
        u32 qh = qhead, tn = trail_n;
        const u32 q0 = qh;
        u32 *const ar = arena;
        u32 t_vis = 0, t_kept = 0, t_deref = 0, t_tail = 0, t_moves = 0, t_impl = 0;
        while (qh < tn) {
            const u32 p = V.trail()[qh];
            qh++;
            const u32 not_p = p ^ 1;
            u32 word = V.wstart()[p];
            const u32 n = V.wlen()[p];
            const bool dirty = (word & W_DIRTY) != 0;
            uint2 *ws = wpool + (word >> W_SHIFT);
            u32 j = 0, i = 0;
            for (; i < n; i += 32) {
                /* Evaluation is written branch-free (safe addresses + selects): on sm_100a every lane-divergent `if`
                 * costs a BSSY / BRA / BSYNC triple plus register moves -- 43 % of the executed instructions of the
                 * first version of this function were such control overhead (profiles/r02_ncu_search_c.md). */
                const u32 idx = i + lane;
                const bool valid = idx < n;
                const uint2 w = ws[valid ? idx : i];
                /* a blocker that is true now stays true for the whole call: such lanes never need their clause */
                const bool need = valid && (dirty || !is_true(w.y));
                const uint4 hd = ld4(ar + (need ? w.x : 0u));
                const bool live = valid && !(dirty && (hd.x & H_DELETED));
                const bool len3 = need && hd.x >= (3u << H_SHIFT);
                const u32 first = need ? ((hd.y == not_p) ? hd.z : hd.y) : 0u;
                const u32 third = len3 ? hd.w : 0u;
                u32 kf = 2, lbase = 0;
                for (;;) {
                    const bool pend = live && lane >= lbase;
                    const u32 bv = lval(w.y), fv = lval(first), tv = lval(third);
                    const bool can_move = len3 && kf == 2 && tv != 0;
                    u32 out = bv == 1 ? OUT_KEEPW : fv == 1 ? OUT_KEEPCW : can_move ? OUT_MOVE : fv == 0 ? OUT_CONFL : OUT_UNIT;
                    if (!pend) out = OUT_NONE;
                    u32 lk = third;
                    const bool scan = out >= OUT_UNIT && hd.x >= (4u << H_SHIFT);
                    if (__ballot_sync(FULL_MASK, scan)) {
                        /* clauses longer than three literals whose third literal is false: 128-bit tail scan */
                        if (scan) {
                            const u32 len = hd.x >> H_SHIFT;
                            if (kf == 2) kf = 3;
                            while (out != OUT_MOVE && kf < len) {
                                const u32 gb = (kf + 1) & ~3u;
                                const uint4 q = ld4(ar + w.x + gb);
                                const u32 qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                for (u32 t = 0; t < 4; t++) {
                                    const u32 k2 = gb + t - 1;
                                    if (out != OUT_MOVE && k2 >= kf && k2 < len && !is_false(qs[t])) { lk = qs[t]; kf = k2; out = OUT_MOVE; }
                                }
                                if (out != OUT_MOVE) kf = gb + 3;
                            }
                        }
                    }
                    const u32 umask0 = __ballot_sync(FULL_MASK, out >= OUT_UNIT);
                    u32 stop = 32, cfl_lane = 32;
                    if (umask0) {
                        const u32 cmask = __ballot_sync(FULL_MASK, out == OUT_CONFL);
                        u32 umask = umask0;
                        do {
                            const u32 u = (u32)__ffs(umask) - 1;
                            if (u >= stop) break;
                            umask &= umask - 1;
                            if ((cmask >> u) & 1) { cfl_lane = u; stop = u + 1; break; }
                            const u32 xv = __shfl_sync(FULL_MASK, first, u) >> 1;
                            const bool aff = pend && lane > u && out != OUT_KEEPW &&
                                             ((w.y >> 1) == xv || (first >> 1) == xv || (out == OUT_MOVE && (lk >> 1) == xv));
                            const u32 am = __ballot_sync(FULL_MASK, aff);
                            if (am) { const u32 s2 = (u32)__ffs(am) - 1; if (s2 < stop) stop = s2; }
                        } while (umask);
                    }
                    const bool commit = pend && lane < stop;
                    const bool mv = commit && out == OUT_MOVE;
                    const bool keep = commit && !mv;
                    /* lanes that move to the same list; issued here so that the match (the longest-latency warp
                     * instruction of the loop) overlaps the stores below instead of stalling the append */
                    const u32 tgt = lk ^ 1;
                    const u32 peers = __match_any_sync(FULL_MASK, mv ? tgt : (0x80000000u | lane));
                    const u32 keepmask = __ballot_sync(FULL_MASK, keep);
                    if (keep) ws[j + __popc(keepmask & lt)] = make_uint2(w.x, out == OUT_KEEPW ? w.y : first);
                    j += __popc(keepmask);
                    if (commit && out != OUT_KEEPW) {
                        if (hd.y == not_p || mv) {
                            ar[w.x + 1] = first;
                            ar[w.x + 2] = mv ? lk : not_p;
                        }
                        if (mv) ar[w.x + 1 + kf] = not_p;
                    }
                    /* implications of the committed prefix, in lane order (assignment.rs:104-113) */
                    if (umask0) {
                        const bool imp = commit && out == OUT_UNIT;
                        const u32 impmask = __ballot_sync(FULL_MASK, imp);
                        if (imp) {
                            const u32 v = first >> 1;
                            V.assign()[v] = (u8)((first & 1) ^ 1);
                            V.level()[v] = (idx_t)nlevels;
                            V.reason()[v] = w.x;
                            V.trail()[tn + __popc(impmask & lt)] = (idx_t)first;
                        }
                        tn += __popc(impmask);
                        if (COUNT) t_impl += __popc(impmask);
                    }
                    const u32 movemask = __ballot_sync(FULL_MASK, mv);
                    if (COUNT) {
                        t_vis += __popc(__ballot_sync(FULL_MASK, commit));
                        t_kept += __popc(keepmask);
                        t_deref += __popc(__ballot_sync(FULL_MASK, commit && out != OUT_KEEPW));
                        u32 tl = 0;
                        if (commit && out >= OUT_MOVE) tl = mv ? kf - 1 : (hd.x >> H_SHIFT) - 2;
                        t_tail += warp_sum(tl);
                        t_moves += __popc(movemask);
                    }
                    if (movemask) {
                        /* ordered parallel append: lanes with the same target list keep their lane order */
                        u32 tword = 0, tlen = 0;
                        bool over = false;
                        if (mv) {
                            tword = V.wstart()[tgt];
                            tlen = V.wlen()[tgt];
                            over = tlen + __popc(peers) > (1u << (tword & W_CAP_MASK));
                        }
#ifdef B200SAT_MOVES_LOOP
                        over = true;
#endif
                        if (__ballot_sync(FULL_MASK, over)) {
                            qhead = qh; trail_n = tn;
                            if (!push_watch_multi(movemask, mv, tgt, w.x, first)) return CREF_UNDEF;
                            word = V.wstart()[p]; /* the pool may have been compacted */
                            ws = wpool + (word >> W_SHIFT);
                        } else { /* every lane has read its list length before the ballot above completed */
                            if (mv) {
                                wpool[(tword >> W_SHIFT) + tlen + __popc(peers & lt)] = make_uint2(w.x, first);
                                if ((peers & lt) == 0) V.wlen()[tgt] = (idx_t)(tlen + __popc(peers));
                            }
                        }
                    }
                    __syncwarp();
                    if (cfl_lane != 32) {
                        confl = __shfl_sync(FULL_MASK, w.x, cfl_lane);
                        const bool rem = live && lane >= stop;
                        const u32 m = __ballot_sync(FULL_MASK, rem);
                        if (rem) ws[j + __popc(m & lt)] = w;
                        j += __popc(m);
                        break;
                    }
                    if (stop == 32) break;
                    lbase = stop;
                }
                if (confl != CREF_UNDEF) { i += 32; break; }
            }
            /* conflict: the rest of the list is copied unchanged (watches.rs:123-137), minus dead watchers */
            for (; i < n; i += 32) {
                const u32 idx = i + lane;
                bool live = idx < n;
                uint2 w = make_uint2(0, 0);
                if (live) {
                    w = ws[idx];
                    if (dirty) live = (ar[w.x] & H_DELETED) == 0;
                }
                const u32 m = __ballot_sync(FULL_MASK, live);
                if (live) ws[j + __popc(m & lt)] = w;
                j += __popc(m);
            }
            /* every lane stores the same words (one predicated STS each instead of a lane-0 branch region); nobody
             * reads them before the barrier below */
            V.wlen()[p] = (idx_t)j;
            if (dirty) V.wstart()[p] = word & ~W_DIRTY;
            __syncwarp();
            if (confl != CREF_UNDEF) break;
        }
        qhead = qh; trail_n = tn;
        if (lane == 0) {
            sc().propagations += qh - q0;
            if (COUNT) {
                u64 *t = sc().traffic;
                t[0] += t_vis; t[1] += t_kept; t[2] += t_deref; t[3] += t_tail; t[4] += t_moves; t[5] += t_impl;
            }
        }
        __syncwarp();
        return confl;
    }

    /* -------------------------------------------------------------- K2: event record of the search prefix + BCP replay
     * (SURVEY.md section 7 K2 / 8d "BCP-only figure").  The instrumented search kernel records decisions, learnt
     * clauses with their backtrack levels and restarts until the first database reduction / ground simplification
     * (or until the buffer is full); bcp_replay() then re-executes exactly that prefix with propagate2() as the only
     * real work -- the decisions and learnt clauses are read back instead of being computed -- so its run time is
     * the BCP share of the search and its propagation count must equal the recorded one. */
    DEV u32 *rec_buf(u32 item) const { return P.trace_buf + (size_t)item * P.trace_cap; }
    DEV void rec_close(u32 item) { /* lane 0 */
        if (sc().rec_on) {
            u32 *tb = rec_buf(item);
            sc().rec_on = 0;
            tb[TR_WORDS] = sc().rec_n;
            tb[TR_PROPS_LO] = (u32)sc().propagations; tb[TR_PROPS_HI] = (u32)(sc().propagations >> 32);
            for (int i = 0; i < 6; i++) { tb[TR_TRAFFIC + 2 * i] = (u32)sc().traffic[i]; tb[TR_TRAFFIC + 2 * i + 1] = (u32)(sc().traffic[i] >> 32); }
        }
    }
    DEV void rec_stop(u32 item) {
        if (lane == 0) rec_close(item);
        __syncwarp();
    }
    /* lane 0: reserve `n` event words; closes the record and returns 0 when it is full */
    DEV u32 *rec_reserve(u32 item, u32 n) {
        if (!sc().rec_on) return nullptr;
        if (sc().rec_n + n > P.trace_cap) { rec_close(item); return nullptr; }
        u32 *w = rec_buf(item) + sc().rec_n;
        sc().rec_n += n;
        return w;
    }
    DEV bool recording() const { return COUNT && P.trace_buf != nullptr; }
    /* all lanes call these; lane 0 writes.  `n` = 0 records the final ground-level conflict. */
    DEV void rec_conflict(u32 n, u32 bt) {
        if (lane == 0) {
            u32 *w = rec_reserve(res_index, 2 + n);
            if (w) { w[0] = TR_CONFLICT | n; w[1] = bt; const u32 *out = ol(); for (u32 i = 0; i < n; i++) w[2 + i] = out[i]; }
        }
        __syncwarp();
    }
    DEV void rec_word(u32 word) {
        if (lane == 0) {
            u32 *w = rec_reserve(res_index, 1);
            if (w) w[0] = word;
        }
        __syncwarp();
    }
    DEV bool will_simplify() const { /* the guard of try_simplify(), search.rs:522-530 */
        return nlevels == 1 && !((sc().simp_db_assigns_set && sc().simp_db_assigns == trail_n) || sc().propagations < sc().simp_db_props);
    }
    DEV void cancel_until_light(u32 target_level) { /* assignment.rs:116-130 without the heuristic's callbacks */
        if (nlevels > target_level) {
            const u32 target = V.lim()[target_level];
            for (u32 k = target + lane; k < trail_n; k += 32) V.assign()[V.trail()[k] >> 1] = LB_UNDEF;
            trail_n = target;
            nlevels = target_level;
        }
        if (qhead > trail_n) qhead = trail_n;
        __syncwarp();
    }
    /* returns false when the replay left the recorded trace (a bug) */
    DEV bool bcp_replay(u32 item) {
        const u32 *tb = rec_buf(item);
        const u32 n_words = tb[TR_WORDS];
        u32 pos = TR_EVENTS;
        bool ok = true;
        u32 confl = propagate2();
        while (ok && !failed()) {
            if (pos >= n_words) break; /* end of the record (whether the last propagation conflicts was not recorded) */
            const bool expect_confl = (tb[pos] & 0xC0000000u) == TR_CONFLICT;
            if ((confl != CREF_UNDEF) != expect_confl) { ok = false; break; }
            const u32 ev = tb[pos];
            if (ev == TR_RESTART) { cancel_until_light(GROUND); pos += 1; }
            else if (ev & 0x80000000u) {
                const u32 n = ev & 0x3FFFFFFFu, bt = tb[pos + 1];
                if (n == 0) break; /* the final ground-level conflict */
                const u32 *lits = tb + pos + 2;
                pos += 2 + n;
                cancel_until_light(bt);
                u32 reason = CREF_UNDEF;
                if (n > 1) {
                    reason = alloc_clause(n, H_LEARNT | H_EXTRA);
                    if (reason == CREF_UNDEF) break;
                    for (u32 i = lane; i < n; i += 32) arena[reason + 1 + i] = lits[i];
                    if (lane == 0) arena[reason + 1 + n] = f2u(0.0f);
                    __syncwarp();
                }
                assign_lit(lits[0], reason);
                if (reason != CREF_UNDEF && !attach(reason)) break;
            } else {
                new_decision_level();
                assign_lit(ev, CREF_UNDEF);
                pos += 1;
            }
            confl = propagate2();
        }
        if (ok && !failed()) {
            const u64 want = (u64)tb[TR_PROPS_LO] | ((u64)tb[TR_PROPS_HI] << 32);
            ok = sc().propagations == want;
        }
        return ok && !failed();
    }

    /* -------------------------------------------------------------- straggler rescue (see KParams::rescue) */
    DEV u32 *rescue_ring(u32 warp) const { return P.rrings + (size_t)warp * (RING_PAYLOAD + P.rr_words); }
    /* the original replica of `item` starts on this warp: its ring is handed over to the new instance */
    DEV void rescue_begin_main(u32 warp, u32 item) {
        rx_ring = rescue_ring(warp); rx_tag = 0x80000000u | item; rx_main = true; result_flags = 0;
        if (lane == 0) { *(volatile u32 *)(rx_ring + RING_ITEM) = 0; __threadfence(); } /* helpers of the previous instance stop */
        __syncwarp();
        for (u32 k = lane; k < P.rr_words; k += 32) rx_ring[RING_PAYLOAD + k] = 0; /* no stale entry survives the hand-over */
        __syncwarp();
        if (lane == 0) {
            *(volatile u32 *)(rx_ring + RING_HEAD) = 0; *(volatile u32 *)(rx_ring + RING_DONE) = 0;
            *(volatile u32 *)(rx_ring + RING_HELPERS) = 0;
            __threadfence();
            *(volatile u32 *)(rx_ring + RING_ITEM) = rx_tag;
            *(volatile u32 *)(P.run_table + warp) = rx_tag;
        }
        __syncwarp();
    }
    DEV void rescue_end_main(u32 warp) {
        if (lane == 0) { *(volatile u32 *)(P.run_table + warp) = 0; __threadfence(); }
        __syncwarp();
    }
    /* First finisher takes the instance: RING_DONE receives the instance's tag (a stale value left by a late helper of
     * the ring's previous instance is simply overwritten); all lanes get the answer. */
    DEV bool rescue_claim() {
        u32 won = 0;
        if (lane == 0 && ld_vol(rx_ring + RING_ITEM) == rx_tag) {
            for (;;) {
                const u32 old = ld_vol(rx_ring + RING_DONE);
                if (old == rx_tag) break; /* another replica of this instance was first */
                if (atomicCAS(rx_ring + RING_DONE, old, rx_tag) == old) { won = 1; break; }
            }
        }
        return bcast(won) != 0;
    }
    /* An idle warp looks for a running instance to help: returns the main warp's index or ~0u.  Instances with few
     * helpers are preferred so that the idle warps spread over the stragglers. */
    DEV u32 rescue_pick(u32 n_warps, u32 my_warp, u32 round) {
        for (u32 pass = 0; pass < 2; pass++) {
            const u32 cap = pass == 0 ? 4u + round : 0x7FFFFFFFu;
            const u32 start = (my_warp * 2654435761u + round * 40503u) % n_warps;
            for (u32 k = 0; k < n_warps; k += 32) {
                const u32 idx = (start + k + lane) % n_warps;
                bool ok = false;
                if (k + lane < n_warps && ld_vol(P.run_table + idx) != 0) {
                    const u32 *rg = rescue_ring(idx);
                    ok = ld_vol(rg + RING_DONE) != ld_vol(rg + RING_ITEM) && ld_vol(rg + RING_HELPERS) < cap;
                }
                const u32 m = __ballot_sync(FULL_MASK, ok);
                if (m) return __shfl_sync(FULL_MASK, idx, __ffs(m) - 1);
            }
        }
        return 0xFFFFFFFFu;
    }
    /* become helper replica of the instance main warp `mw` runs: clone its pristine post-ingest state into my slab and
     * diversify.  false = the instance is gone already. */
    DEV bool rescue_join(u32 mw) {
        u32 *rg = rescue_ring(mw);
        u32 tag = 0, hidx = 0;
        if (lane == 0) {
            tag = ld_vol(rg + RING_ITEM);
            if (tag != 0 && ld_vol(rg + RING_DONE) != tag) hidx = atomicAdd(rg + RING_HELPERS, 1u) + 1;
            else tag = 0;
        }
        tag = bcast(tag); hidx = bcast(hidx);
        if (tag == 0) return false;
        const u32 item = tag & 0x7FFFFFFFu;
        const uint4 *src = reinterpret_cast<const uint4 *>(P.pristine + (size_t)(item - P.chunk_base) * P.L.slab_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(slab);
        for (u64 k = lane; k < P.L.slab_bytes / 16; k += 32) dst[k] = src[k];
        __syncwarp();
        rx_ring = rg; rx_tag = tag; rx_main = false; result_flags = 0x100;
        res_index = P.work ? P.work[item] : item;
        item_index = item;
        stp = P.replica_settings + (1 + (hidx + item) % (P.n_div - 1)); /* entry 0 is the reference configuration */
        load_state();
        if (lane == 0) { /* the diversified configuration replaces what setup() derived from the base settings */
            sc().seed = st().random_seed + (double)(hidx % 977);
            sc().inv_var_decay = 1.0 / st().var_decay;
            sc().inv_cla_decay = 1.0 / st().clause_decay;
            for (int i = 0; i < 8; i++) sc().ring_cursor[i] = 0;
            if (st().rnd_init_act) for (u32 v = 0; v < nvars; v++) V.act()[v] = drand() * 0.00001;
        }
        { u32 *filter = (u32 *)(slab + P.L.off_pf); for (u32 k = lane; k < 256; k += 32) filter[k] = 0; }
        __syncwarp();
        if (st().rnd_init_act) rebuild_order_heap();
        return true;
    }

    /* -------------------------------------------------------------- park / resume (no reference analogue) */
    DEV void save_state() {
        __syncwarp();
        if (lane == 0) { sc().sv_nvars = nvars; sc().sv_trail_n = trail_n; sc().sv_qhead = qhead; sc().sv_nlevels = nlevels; }
        __syncwarp();
        {
            const uint2 *src = reinterpret_cast<const uint2 *>(V.s);
            uint2 *dst = reinterpret_cast<uint2 *>(slab + P.L.off_save);
            for (u32 k = lane; k < P.L.save_bytes / 8; k += 32) dst[k] = src[k];
        }
        __syncwarp();
    }
    DEV void load_state() {
        {
            uint2 *dst = reinterpret_cast<uint2 *>(V.s);
            const uint2 *src = reinterpret_cast<const uint2 *>(slab + P.L.off_save);
            for (u32 k = lane; k < P.L.save_bytes / 8; k += 32) dst[k] = src[k];
        }
        __syncwarp();
        nvars = sc().sv_nvars; trail_n = sc().sv_trail_n; qhead = sc().sv_qhead; nlevels = sc().sv_nlevels;
        arena = arena_half(sc().cur_arena);
        wpool = wpool_half(sc().cur_wpool);
        __syncwarp();
    }

    /* search/util.rs:5-14 progress_estimate(): same operation order, so the f64 result is bit-identical */
    DEV double progress_estimate() {
        double progress = 0.0;
        if (lane == 0) {
            const double vars = 1.0 / (double)nvars;
            double factor = vars;
            for (u32 lv = 0; lv < nlevels; lv++) {
                const u32 beg = (u32)V.lim()[lv];
                const u32 end = lv + 1 == nlevels ? trail_n : (u32)V.lim()[lv + 1];
                progress += factor * (double)(end - beg);
                factor *= vars;
            }
        }
        return progress;
    }

    enum : u32 { SR_SAT = 0, SR_UNSAT = 1, SR_ABORT = 2, SR_SUSPEND = 3, SR_INTERRUPTED = 4 };

    /* true while other runnable items wait in the work ring (lane 0 reads, all lanes agree) */
    DEV bool others_waiting() {
        u32 v = 0;
        if (lane == 0 && P.qctl) v = ld_vol(P.qctl + QC_TAIL) > ld_vol(P.qctl + QC_HEAD) ? 1u : 0u;
        return bcast(v) != 0;
    }
    DEV bool interrupted_now() {
        u32 v = 0;
        if (lane == 0 && P.interrupt) v = *P.interrupt != 0 ? 1u : 0u; /* volatile device word */
        return bcast(v) != 0;
    }

    /* Searcher::search_internal + search_loop + propagate_learn_backtrack (search.rs:409-517) as one resumable loop.
     * Entered with phase PH_READY (a new search() call: solves += 1, LearningGuard reset, restart counter 0) or
     * PH_SEARCH (parked mid-loop by SR_SUSPEND: continues exactly where it stopped, so slicing is trace-neutral). */
    DEV u32 run_search() {
        const bool fresh_call = sc().phase == PH_READY;
        __syncwarp(); /* every lane has read the phase before lane 0 changes it */
        if (fresh_call) {
            if (lane == 0) {
                sc().solves++;
                double ml = (double)sc().num_clauses * st().size_factor;       /* LearningGuard::reset search.rs:89-94 */
                double lo = (double)st().min_learnts_lim;
                sc().max_learnts = ml > lo ? ml : lo;
                sc().size_adjust_confl = (double)st().size_adjust_start_confl;
                sc().size_adjust_cnt = st().size_adjust_start_confl;
                sc().curr_restarts = 0;
                sc().in_loop = 0;
                sc().phase = PH_SEARCH;
                if (recording()) {
                    sc().rec_on = 1; sc().rec_n = TR_EVENTS;
                    u32 *tb = rec_buf(res_index);
                    tb[TR_PROPS0_LO] = (u32)sc().propagations; tb[TR_PROPS0_HI] = (u32)(sc().propagations >> 32);
                    for (int i = 0; i < 6; i++) { tb[TR_TRAFFIC0 + 2 * i] = (u32)sc().traffic[i]; tb[TR_TRAFFIC0 + 2 * i + 1] = (u32)(sc().traffic[i] >> 32); }
                }
            }
            __syncwarp();
        }
        u64 slice_end = P.slice_conflicts ? sc().conflicts + P.slice_conflicts : ~0ull;
        for (;;) {
            if (!sc().in_loop) { /* search_loop prologue, search.rs:465-467 */
                if (PF && (P.share & PF_SHARE) && sc().curr_restarts > 0 && !rx_main) {
                    if (!import_clauses()) return SR_UNSAT;
                    if (failed()) return SR_ABORT;
                }
                __syncwarp();
                if (lane == 0) {
                    const u32 cr = sc().curr_restarts;
                    const double base = st().luby_restart ? luby(st().restart_inc, cr) : powi(st().restart_inc, (i32)cr);
                    sc().starts++;
                    sc().confl_limit = sc().conflicts + (u64)(base * st().restart_first);
                    sc().in_loop = 1;
                }
                __syncwarp();
            }
            const u64 confl_limit = sc().confl_limit;
            for (;;) {
                /* propagate_learn_backtrack, search.rs:503-517 + handle_conflict :263-304 */
                bool had_conflict = false;
                for (;;) {
                    const u32 confl = propagate2();
                    if (failed()) return SR_ABORT;
                    if (confl == CREF_UNDEF) break;
                    had_conflict = true;
                    if (lane == 0) sc().conflicts++;
                    if (nlevels == 1) { __syncwarp(); if (recording()) { rec_conflict(0, 0); rec_stop(res_index); } return SR_UNSAT; }
                    u32 bt;
                    const u32 n = analyze(confl, bt);
                    if (failed()) return SR_ABORT;
                    if (recording()) rec_conflict(n, bt);
                    if (PF) {
                        if (P.share & PF_SHARE) export_clause(n);
                        if (cancelled()) { fail(ST_CANCELLED); return SR_ABORT; }
                    }
                    /* the reference prints its progress row inside handle_conflict, i.e. BEFORE the trail is backtracked
                     * (search.rs:263-304 vs :503-517): the estimate is taken here when this conflict bumps the limit */
                    double row_progress = 0.0;
                    if (COUNT && P.progress_buf && sc().size_adjust_cnt == 1) row_progress = progress_estimate() * 100.0;
                    cancel_until(bt);
                    u32 reason = CREF_UNDEF;
                    const u32 *out = ol();
                    const u32 asserting = out[0];
                    if (n > 1) {
                        if (sc().n_learnts >= P.L.learnts_cap) { fail(ST_MEM_ARENA); return SR_ABORT; }
                        reason = alloc_clause(n, H_LEARNT | H_EXTRA);
                        if (reason == CREF_UNDEF) return SR_ABORT;
                        for (u32 i = lane; i < n; i += 32) arena[reason + 1 + i] = out[i];
                        if (lane == 0) {
                            arena[reason + 1 + n] = f2u(0.0f);
                            learnts()[sc().n_learnts] = reason;
                            sc().n_learnts++;
                            sc().num_learnts++;
                            sc().learnts_literals += n;
                            if (COUNT) sc().traffic[7] += 1 + n + 4;
                        }
                        __syncwarp();
                        bump_cla(reason, n);
                    }
                    if (lane == 0) {
                        sc().var_inc *= sc().inv_var_decay;   /* decision_heuristic.rs:142-144 */
                        sc().cla_inc *= sc().inv_cla_decay;   /* clause_db.rs:145-147 */
                        sc().size_adjust_cnt -= 1;            /* LearningGuard::bump search.rs:96-110 */
                        if (sc().size_adjust_cnt == 0) {
                            sc().size_adjust_confl *= st().size_adjust_inc;
                            sc().size_adjust_cnt = (i32)sc().size_adjust_confl;
                            sc().max_learnts *= st().size_inc;
                            if (COUNT && P.progress_buf) { /* the progress row of the search table, search.rs:289-301 */
                                double *pb = P.progress_buf + (size_t)res_index * (1 + 8 * (size_t)P.progress_cap);
                                const u32 row = (u32)pb[0];
                                if (row < P.progress_cap) {
                                    const u32 ground = nlevels > 1 ? (u32)V.lim()[1] : trail_n;
                                    double *r = pb + 1 + 8 * (size_t)row;
                                    r[0] = (double)sc().conflicts; r[1] = (double)(sc().dec_vars - ground);
                                    r[2] = (double)sc().num_clauses; r[3] = (double)sc().clauses_literals;
                                    r[4] = (double)(u64)sc().max_learnts; r[5] = (double)sc().num_learnts;
                                    r[6] = (double)sc().learnts_literals / (double)sc().num_learnts;
                                    r[7] = row_progress;
                                    pb[0] = (double)(row + 1);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    assign_lit(asserting, reason);
                    if (reason != CREF_UNDEF && !attach(reason)) return SR_ABORT;
                }
                const u64 nconf = sc().conflicts;
                /* Budget::within, search.rs:473-477 (budget.rs:21-25).  The asynchronous-interrupt word lives in
                 * device memory and is polled after conflicts only (a host-memory poll per decision costs a PCIe
                 * round trip: measured 28x slower on the n=100 batch). */
                if (P.conflict_budget >= 0 || P.propagation_budget >= 0 || (had_conflict && P.interrupt)) {
                    const bool out_of = (P.conflict_budget >= 0 && nconf >= (u64)P.conflict_budget) ||
                                        (P.propagation_budget >= 0 && sc().propagations >= (u64)P.propagation_budget) ||
                                        (had_conflict && interrupted_now());
                    if (out_of) {
                        const double pe = progress_estimate();
                        cancel_until(GROUND);
                        if (lane == 0) { sc().progress = pe; sc().in_loop = 0; sc().phase = PH_READY; }
                        __syncwarp();
                        return SR_INTERRUPTED;
                    }
                }
                if (nconf >= confl_limit) { /* search.rs:479-482 */
                    if (recording()) rec_word(TR_RESTART);
                    cancel_until(GROUND);
                    if (lane == 0) { sc().curr_restarts++; sc().in_loop = 0; }
                    __syncwarp();
                    break;
                }
                if (nconf >= slice_end) { /* yield to waiting instances; nothing of the trace depends on this */
                    if (!(PF && rx_ring && !rx_main) && others_waiting()) return SR_SUSPEND; /* helper replicas never park */
                    slice_end = nconf + P.slice_conflicts;
                }
                if (recording() && will_simplify()) rec_stop(res_index);
                try_simplify();
                if ((double)sc().n_learnts >= sc().max_learnts + (double)trail_n) {
                    if (recording()) rec_stop(res_index);
                    reduce_db();
                    if (failed()) return SR_ABORT;
                    try_garbage_collect();
                }
                u32 next = LIT_NONE;
                /* SearchCtx::decide, search.rs:216-232: user assumptions first, one (possibly dummy) level each */
                while (nlevels - 1 < P.n_assumps) {
                    const u32 a = P.assumps[nlevels - 1];
                    const u32 av = lval(a);
                    __syncwarp(); /* every lane has read the value before cancel_until() below starts to unassign */
                    if (av == 1) new_decision_level();
                    else if (av == 0) {
                        /* analyze_final (conflict.rs:255-279) computes the set of responsible assumptions, which the
                         * reference then drops ("TODO: implement properly", search.rs:431-435): UNSAT at ground level */
                        cancel_until(GROUND);
                        return SR_UNSAT;
                    } else { next = a; break; }
                }
                if (next == LIT_NONE) {
                    if (lane == 0) sc().decisions++; /* :234-236, counted for real decisions only */
                    __syncwarp();
                    next = pick_branch_lit();
                }
                if (next == LIT_NONE) { if (recording()) rec_stop(res_index); return SR_SAT; }
                if (recording()) rec_word(next);
                new_decision_level();
                assign_lit(next, CREF_UNDEF);
            }
        }
    }

    /* -------------------------------------------------------------- work ring (device side) */
    /* hand work item `w` (whose state is parked in its slab) to whichever warp pops it next */
    DEV void ring_push(u32 w) {
        if (lane == 0) {
            __threadfence(); /* the parked state is visible before the ticket is */
            const u32 s = atomicAdd(P.qctl + QC_TAIL, 1u);
            *(volatile unsigned long long *)(P.ring + (s & P.ring_mask)) = ((unsigned long long)(s + 1) << 32) | w;
        }
        __syncwarp();
    }
    DEV void count_done() {
        if (lane == 0) { __threadfence(); atomicAdd(P.qctl + QC_DONE, 1u); }
        __syncwarp();
    }

    /* elim_clauses.rs:43-63 over the eliminated-clause stack the simplifier left in the slab (lane 0) */
    DEV void extend_model_from_slab(u8 *model) {
        if (lane == 0) {
            const SimpLayout SL = simp_layout(P.L.nv_cap, P.L.clauses_cap, P.L.simp_lits, P.L.tmp_cap);
            const u32 *base = (const u32 *)(slab + P.L.off_simp);
            const u32 *elits = base + SL.o_elits, *esizes = base + SL.o_esizes;
            for (u32 i = sc().elim_n; i-- > 0;) {
                const u32 head = i > 0 ? esizes[i - 1] : 0, tail = esizes[i];
                bool sat = false;
                for (u32 k = head; k < tail; k++) {
                    const u32 l = elits[k];
                    const u8 m = model[l >> 1];
                    if (m != 2 && (m != 0) != ((l & 1) != 0)) { sat = true; break; }
                }
                if (!sat) { const u32 l = elits[head]; model[l >> 1] = (l & 1) ? 0 : 1; }
            }
        }
        __syncwarp();
    }

    /* result record + model of a finished (or interrupted) search */
    DEV void finish_search(u32 r, bool simp_kind) {
        const i32 verdict = (r == SR_SAT) ? B200SAT_SAT : (r == SR_UNSAT) ? B200SAT_UNSAT : B200SAT_INTERRUPTED;
        __syncwarp();
        write_result(verdict, sc().n_ingest, sc().n_pre, sc().pre_ok);
        if (simp_kind && P.models && !failed() && verdict == B200SAT_SAT && st().extend_model && sc().elim_n)
            extend_model_from_slab(P.models + (size_t)res_index * P.model_stride);
        if (PF && !failed() && verdict != B200SAT_INTERRUPTED) announce_done();
        if (r == SR_INTERRUPTED) save_state(); /* SolveRes::Interrupted hands the solver back (sat.rs:24): stays resident */
        else if (lane == 0) sc().phase = PH_DONE;
        if (lane == 0 && P.parked) P.parked[item_index] = r == SR_INTERRUPTED ? 1 : 0;
        __syncwarp();
    }

    /* -------------------------------------------------------------- ingest phase of one instance (CoreSolver):
     * add_clause* + preprocess (sat/minisat.rs:37-55, lib.rs:47-100); leaves the instance parked (PH_READY) when a
     * search has to follow, and always writes the result record of the phase */
    DEV bool ingest_core(u32 inst) {
        u32 cb, ce;
        bool ok = setup(inst, cb, ce);
        u32 n_ingest = 0, n_pre = 0, pre_ok = 1;
        i32 verdict = B200SAT_INTERRUPTED;
        bool runnable = false;
        if (ok) {
            for (u32 c = cb; c < ce && ok; c++) { /* CoreSolver::add_clause minisat.rs:41-48 */
                u32 b = P.clause_off[c], e = P.clause_off[c + 1];
                u32 cr;
                if (st().use_rcheck) { /* search.rs:343-345 */
                    bool imp = is_implied(P.lits + b, e - b);
                    if (failed()) break;
                    if (imp) continue;
                }
                if (add_clause(P.lits + b, e - b, cr) == 0) ok = false;
                if (failed()) break;
            }
            n_ingest = sc().num_clauses;
            n_pre = n_ingest;
            if (!failed()) {
                if (P.flags & B200SAT_F_PREPROCESS) { /* CoreSolver::preprocess minisat.rs:50-55, search.rs:388-395 */
                    if (ok) {
                        u32 confl = propagate();
                        if (confl == CREF_UNDEF) try_simplify(); else ok = false;
                    }
                    pre_ok = ok ? 1 : 0;
                    n_pre = sc().num_clauses;
                }
                if (!failed()) {
                    if (!ok) verdict = B200SAT_UNSAT;
                    else runnable = true;
                }
            }
        }
        return park_after_ingest(verdict, n_ingest, n_pre, pre_ok, runnable);
    }
    /* common tail of the ingest kernels; returns true when the instance waits for the search kernel */
    DEV bool park_after_ingest(i32 verdict, u32 n_ingest, u32 n_pre, u32 pre_ok, bool runnable) {
        __syncwarp();
        runnable = runnable && !failed();
        if (lane == 0) {
            sc().n_ingest = n_ingest; sc().n_pre = n_pre; sc().pre_ok = pre_ok;
            sc().phase = runnable ? PH_READY : PH_DONE;
        }
        __syncwarp();
        write_result(verdict, n_ingest, n_pre, pre_ok);
        if (PF && !runnable && !failed()) announce_done();
        if (runnable) save_state();
        if (lane == 0 && P.parked) P.parked[item_index] = runnable ? 1 : 0;
        return runnable;
    }

    /* -------------------------------------------------------------- per-instance setup + ingest */
    DEV bool setup(u32 inst, u32 &cb, u32 &ce) {
        cb = P.inst_clause_off[inst];
        ce = P.inst_clause_off[inst + 1];
        const u32 lb = P.clause_off[cb], le = P.clause_off[ce];
        u32 mx = 0;
        for (u32 k = lb + lane; k < le; k += 32) { u32 v = P.lits[k] >> 1; mx = v > mx ? v : mx; }
        mx = warp_max(mx);
        nvars = (le > lb) ? mx + 1 : 0;
        const u8 *vfl = nullptr; /* Solver::new_var(upol, dvar) records of this instance (sat.rs:31), if any */
        if (P.var_flags) {
            const u32 vb = P.inst_var_off[inst], nv_given = P.inst_var_off[inst + 1] - vb;
            vfl = P.var_flags + vb;
            if (nv_given > nvars) nvars = nv_given; /* vars that were created but occur in no clause */
            else if (nv_given < nvars) vfl = nullptr; /* the host checks this; never read out of range */
        }
        trail_n = 0; qhead = 0; nlevels = 1;
        if (lane == 0) {
            Scal &s = sc();
            s.solves = s.starts = s.decisions = s.rnd_decisions = s.conflicts = s.propagations = 0;
            s.max_literals = s.tot_literals = 0;
            for (int i = 0; i < 8; i++) s.traffic[i] = 0;
            s.clauses_literals = s.learnts_literals = s.lc_size = s.lc_wasted = s.simp_db_props = 0;
            s.trace_hash = 1469598103934665603ull;
            s.var_inc = 1.0; s.cla_inc = 1.0; s.max_learnts = 0.0; s.size_adjust_confl = 0.0;
            s.seed = st().random_seed;
            s.inv_var_decay = 1.0 / st().var_decay;
            s.inv_cla_decay = 1.0 / st().clause_decay;
            s.arena_top = 0; s.wpool_top = 0; s.n_learnts = 0; s.n_clauses = 0; s.num_clauses = 0; s.num_learnts = 0;
            s.size_adjust_cnt = 0; s.simp_db_assigns = 0; s.status = ST_OK;
            s.simp_db_assigns_set = (P.flags & B200SAT_F_SIMP_OFF_EMPTY) ? 1u : 0u; /* lib.rs:36-41: try_simplify already ran once, on nothing */
            s.dec_vars = nvars; s.heap_n = nvars; s.cur_arena = 0; s.cur_wpool = 0; s.eliminated_vars = 0;
            s.remove_satisfied = st().remove_satisfied; s.extra_clause_field = 0; s.elim_n = 0;
            for (int i = 0; i < 8; i++) s.ring_cursor[i] = 0;
            s.exported = 0; s.imported = 0;
            s.phase = PH_NEW; s.curr_restarts = 0; s.in_loop = 0; s.confl_limit = 0; s.progress = 0.0;
            s.n_ingest = s.n_pre = 0; s.pre_ok = 1; s.rec_n = TR_EVENTS; s.rec_on = 0;
        }
        arena = arena_half(0);
        wpool = wpool_half(0);
        __syncwarp();
        if (nvars > V.cap() || nvars > P.L.nv_cap) { fail(ST_MEM_VARS); return false; }
        for (u32 v = lane; v < nvars; v += 32) {
            V.act()[v] = 0.0;
            V.reason()[v] = CREF_UNDEF;
            V.level()[v] = (idx_t)GROUND;
            V.heap()[v] = (idx_t)v;
            V.hidx()[v] = (idx_t)v;
            V.assign()[v] = LB_UNDEF;
            u32 f = VF_POL | VF_DEC; /* decision_heuristic.rs:70-88: polarity true, user_pol None, decision var */
            if (vfl) {
                const u32 g = vfl[v];
                if (g & VARF_UPOL_SET) f |= VF_UPOL_SET | ((g & VARF_UPOL_VAL) ? VF_UPOL_VAL : 0u);
                if (g & VARF_NOT_DECISION) f &= ~(u32)VF_DEC;
            }
            V.vflags()[v] = (u8)f;
            V.wstart()[2 * v] = 0; V.wstart()[2 * v + 1] = 0;
            V.wlen()[2 * v] = 0; V.wlen()[2 * v + 1] = 0;
        }
        if (lane == 0) V.lim()[0] = 0;
        if (PF || P.portfolio) { u32 *filter = (u32 *)(slab + P.L.off_pf); for (u32 k = lane; k < 256; k += 32) filter[k] = 0; }
        __syncwarp();
        if (st().rnd_init_act || vfl) { /* decision_heuristic.rs:70-102: activity drawn at var creation; only decision
                                           vars enter the heap, in creation order (set_decision_var) */
            if (lane == 0) {
                sc().heap_n = 0;
                u32 nd = 0;
                for (u32 v = 0; v < nvars; v++) {
                    if (st().rnd_init_act) V.act()[v] = drand() * 0.00001;
                    V.hidx()[v] = (idx_t)VT::ABSENT;
                    if (V.vflags()[v] & VF_DEC) { heap_insert(v); nd++; }
                }
                sc().dec_vars = nd;
            }
            __syncwarp();
        }
        /* size the watch lists: list l can never hold more original watchers than
         * the number of clauses containing the complement of l */
        for (u32 k = lb + lane; k < le; k += 32) atomicAdd(&V.wstart()[P.lits[k] ^ 1], 1u);
        __syncwarp();
        u32 top = 0;
        for (u32 l0 = 0; l0 < 2 * nvars; l0 += 32) {
            u32 l = l0 + lane;
            u32 capc = 0, cap = 0;
            if (l < 2 * nvars) {
                u32 cnt = V.wstart()[l];
                capc = ceil_log2(cnt + 4);
                cap = 1u << capc;
            }
            u32 tot;
            u32 off = top + warp_excl_scan(cap, lane, tot);
            if (l < 2 * nvars) V.wstart()[l] = (off << W_SHIFT) | capc;
            top += tot;
        }
        __syncwarp();
        if (top > P.L.wpool_cap) { fail(ST_MEM_WPOOL); return false; }
        if (lane == 0) sc().wpool_top = top;
        __syncwarp();
        return true;
    }

    /* search/util.rs:16-42 (--rcheck): on a scratch decision level assume the negation of every literal
     * and propagate; the clause is implied iff that conflicts (or one literal already holds).  Phases of
     * the scratch level are saved as top-level ones, exactly as cancel_until does for the newest level. */
    DEV bool is_implied(const u32 *c, u32 n) {
        new_decision_level();
        bool hit = false;
        for (u32 i = 0; i < n && !hit; i++) {
            const u32 l = c[i];
            u32 act = 0; /* read on the lane that writes the assignment, so no lane sees a half-updated state */
            if (lane == 0) act = is_true(l) ? 1u : (V.assign()[l >> 1] == LB_UNDEF ? 2u : 0u);
            act = __shfl_sync(FULL_MASK, act, 0);
            if (act == 1) hit = true;
            else if (act == 2) assign_lit(l ^ 1, CREF_UNDEF);
        }
        bool result = hit;
        if (!hit) result = propagate() != CREF_UNDEF;
        cancel_until(GROUND);
        return result && !failed();
    }

    /* Searcher::add_clause search.rs:342-386.  0 = UNSAT, 1 = consumed, 2 = added (cref in out_cr) */
    DEV u32 add_clause(const u32 *src, u32 len, u32 &out_cr, bool as_learnt = false) {
        u32 nk = 0;
        bool consumed = false;
        u32 cr = CREF_UNDEF;
        if (len <= 32) {
            u32 lit = lane < len ? src[lane] : LIT_NONE;
            bool dup = false, taut = false;
            for (u32 j = 0; j < len; j++) {
                u32 o = __shfl_sync(FULL_MASK, lit, j);
                if (o == lit && j < lane) dup = true;
                if (o == (lit ^ 1)) taut = true;
            }
            bool valid = lane < len;
            bool kept = valid && !dup && !is_false(lit);
            consumed = __ballot_sync(FULL_MASK, valid && (taut || is_true(lit))) != 0;
            if (consumed) return 1;
            u32 km = __ballot_sync(FULL_MASK, kept);
            nk = __popc(km);
            if (nk == 0) return 0;
            if (nk == 1) {
                u32 unit = __shfl_sync(FULL_MASK, lit, __ffs(km) - 1);
                assign_lit(unit, CREF_UNDEF);
                u32 confl = propagate();
                if (failed()) return 0;
                return confl == CREF_UNDEF ? 1 : 0;
            }
            u32 key = kept ? lit : LIT_NONE;
            u32 rank = 0;
            for (u32 j = 0; j < len; j++) {
                u32 o = __shfl_sync(FULL_MASK, key, j);
                if (o < key) rank++;
            }
            cr = alloc_clause(nk, as_learnt ? (u32)(H_LEARNT | H_EXTRA) : (sc().extra_clause_field ? (u32)H_EXTRA : 0u));
            if (cr == CREF_UNDEF) return 0;
            if (kept) arena[cr + 1 + rank] = lit;
        } else {
            /* long clause: strided over lanes, O(len^2/32) ranking through scratch memory */
            u32 *t = tmp();
            if (len > P.L.tmp_cap) { fail(ST_MEM_ARENA); return 0; }
            bool cons = false;
            u32 cnt = 0;
            for (u32 i = lane; i < len; i += 32) {
                u32 lit = src[i];
                bool dup = false;
                for (u32 j = 0; j < len; j++) {
                    u32 o = src[j];
                    if (o == lit && j < i) dup = true;
                    if (o == (lit ^ 1)) cons = true;
                }
                if (is_true(lit)) cons = true;
                bool kept = !dup && !is_false(lit);
                t[i] = kept ? lit : LIT_NONE;
                if (kept) cnt++;
            }
            if (__ballot_sync(FULL_MASK, cons)) return 1;
            nk = warp_sum(cnt);
            __syncwarp();
            if (nk == 0) return 0;
            if (nk == 1) {
                u32 unit = LIT_NONE;
                for (u32 i = lane; i < len; i += 32) if (t[i] != LIT_NONE) unit = t[i];
                u32 m = __ballot_sync(FULL_MASK, unit != LIT_NONE);
                unit = __shfl_sync(FULL_MASK, unit, __ffs(m) - 1);
                assign_lit(unit, CREF_UNDEF);
                u32 confl = propagate();
                if (failed()) return 0;
                return confl == CREF_UNDEF ? 1 : 0;
            }
            cr = alloc_clause(nk, sc().extra_clause_field ? (u32)H_EXTRA : 0u);
            if (cr == CREF_UNDEF) return 0;
            for (u32 i = lane; i < len; i += 32) {
                u32 key = t[i];
                if (key == LIT_NONE) continue;
                u32 rank = 0;
                for (u32 j = 0; j < len; j++) if (t[j] < key) rank++;
                arena[cr + 1 + rank] = key;
            }
        }
        __syncwarp();
        if (as_learnt) { /* imported clause: stored like a learnt one (clause_db.rs:93-100) */
            if (sc().n_learnts >= P.L.learnts_cap) { fail(ST_MEM_ARENA); return 0; }
            if (lane == 0) {
                arena[cr + 1 + nk] = f2u(0.0f);
                learnts()[sc().n_learnts] = cr;
                sc().n_learnts++;
                sc().num_learnts++;
                sc().learnts_literals += nk;
            }
            __syncwarp();
            bump_cla(cr, nk);
            if (!attach(cr)) return 0;
            out_cr = cr;
            return 2;
        }
        if (sc().extra_clause_field) { /* clause_db.rs:78-91 + formula/util.rs:5-11 */
            u32 a = 0;
            for (u32 i = lane; i < nk; i += 32) a |= 1u << ((arena[cr + 1 + i] >> 1) & 31);
            for (int o = 16; o > 0; o >>= 1) a |= __shfl_xor_sync(FULL_MASK, a, o);
            if (lane == 0) arena[cr + 1 + nk] = a;
        }
        if (sc().n_clauses >= P.L.clauses_cap) { fail(ST_MEM_ARENA); return 0; }
        if (lane == 0) {
            clauses()[sc().n_clauses] = cr;
            sc().n_clauses++;
            sc().num_clauses++;
            sc().clauses_literals += nk;
        }
        __syncwarp();
        if (!attach(cr)) return 0;
        out_cr = cr;
        return 2;
    }

    DEV void write_result(i32 verdict, u32 n_ingest, u32 n_pre, u32 pre_ok) {
        b200sat_result *r = P.results + res_index;
        const Scal &s = sc();
        bool zero = !pre_ok && (P.flags & B200SAT_F_ZERO_STATS_ON_PRE_UNSAT);
        if (lane == 0) {
            r->verdict = s.status == ST_OK ? verdict : B200SAT_INTERRUPTED;
            r->status = s.status == ST_OK ? B200SAT_OK : (s.status == ST_CANCELLED ? 1 : -(i32)(100 + s.status));
            r->n_vars = nvars;
            r->n_clauses_ingest = n_ingest;
            r->n_clauses_pre = n_pre;
            r->eliminated_vars = s.eliminated_vars;
            r->pre_ok = pre_ok;
            r->attempts = P.attempt | result_flags;
            r->stats[0] = zero ? 0 : s.solves;       /* search.rs:590-601 */
            r->stats[1] = zero ? 0 : s.starts;
            r->stats[2] = zero ? 0 : s.decisions;
            r->stats[3] = zero ? 0 : s.rnd_decisions;
            r->stats[4] = zero ? 0 : s.conflicts;
            r->stats[5] = zero ? 0 : s.propagations;
            r->stats[6] = zero ? 0 : s.tot_literals;
            r->stats[7] = zero ? 0 : s.max_literals - s.tot_literals;
            for (int i = 0; i < 8; i++) r->traffic[i] = s.traffic[i];
            r->trace_hash = s.trace_hash;
            r->exported = s.exported;
            r->imported = s.imported;
            r->progress = (s.status == ST_OK && verdict == B200SAT_INTERRUPTED) ? s.progress : 0.0;
        }
        if (P.models && s.status == ST_OK && verdict == B200SAT_SAT) { /* formula/util.rs:35-41 */
            u8 *m = P.models + (size_t)res_index * P.model_stride;
            for (u32 v = lane; v < P.model_stride; v += 32) m[v] = v < nvars ? V.assign()[v] : (u8)LB_UNDEF;
        }
        __syncwarp();
    }
};

} // namespace b200sat
#endif
