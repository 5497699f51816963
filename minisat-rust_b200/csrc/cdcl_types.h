/*
 * cdcl_types.h -- plain-data definitions shared by the host launcher and the
 * device solver (kernel parameters, per-warp slab layout, clause/watch word
 * encodings).  No reference code is copied here; encodings are this
 * engine's own (DESIGN.md "Data layout in HBM").
 */
#ifndef B200SAT_CDCL_TYPES_H
#define B200SAT_CDCL_TYPES_H
#include <stdint.h>
#include "../../include/b200sat.h"

namespace b200sat {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

static const u32 FULL_MASK = 0xFFFFFFFFu;
static const u32 CREF_UNDEF = 0xFFFFFFFFu;
static const u32 LIT_NONE = 0xFFFFFFFFu;
static const u32 GROUND = 1; /* assignment.rs:10 GROUND_LEVEL */

/* clause header word: len << H_SHIFT | flags.  A clause block is
 *   [hdr][lit0][lit1]...[lit len-1][extra?]   rounded up to 4 words, 16-byte aligned,
 * extra = f32 activity (learnt, clause_header.rs:7) or u32 signature (clause_header.rs:6). */
enum : u32 { H_DELETED = 1, H_RELOCED = 2, H_LEARNT = 4, H_EXTRA = 8, H_TOUCHED = 16, H_SHIFT = 5 };
/* per-literal watch list descriptor: start << W_SHIFT | dirty | log2(capacity) */
enum : u32 { W_CAP_MASK = 31, W_DIRTY = 32, W_SHIFT = 6 };
/* per-var flag byte */
enum : u32 { VF_SEEN = 3, VF_POL = 4, VF_DEC = 8, VF_UPOL_SET = 16, VF_UPOL_VAL = 32, VF_FROZEN = 64, VF_ELIM = 128 };
enum : u32 { SEEN_UNDEF = 0, SEEN_SOURCE = 1, SEEN_REMOVABLE = 2, SEEN_FAILED = 3 }; /* conflict.rs:18-24 */
enum : u8 { LB_FALSE = 0, LB_TRUE = 1, LB_UNDEF = 2 };                                /* formula.rs:14-18 */

/* internal per-instance status */
enum : u32 { ST_OK = 0, ST_MEM_ARENA = 1, ST_MEM_WPOOL = 2, ST_MEM_LIST = 3, ST_MEM_VARS = 4, ST_MEM_SIMP = 5, ST_INTERNAL = 9 };

/* Per-warp working slab in HBM (all sizes in elements; offsets in bytes). */
struct Layout {
    u32 nv_cap;       /* variable capacity                                  */
    u32 arena_cap;    /* words per arena half (multiple of 4)               */
    u32 wpool_cap;    /* watchers per pool half                             */
    u32 learnts_cap;  /* entries of the learnts list                        */
    u32 clauses_cap;  /* entries of the clauses list                        */
    u32 tmp_cap;      /* words of ingest scratch (>= longest input clause)  */
    u32 simp_cap;     /* words of simplifier heap (0 when CORE only)        */
    u32 simp_lits;    /* literal budget the simplifier region was sized for */
    u32 tier;         /* 0: var state in shared memory, 1: in the slab      */
    u64 off_arena[2], off_wpool[2], off_learnts, off_clauses, off_ol, off_tc, off_stk, off_tmp, off_vars,
        off_simp;
    u64 off_save;     /* parked copy of the shared-memory var state of a suspended instance */
    u64 off_pf;       /* portfolio: 8192-bit filter of the clause hashes this replica already imported */
    u32 save_bytes;   /* = sizeof(SmemVars<NV>) of the tier (sizeof(SmemSmall) for the global tier) */
    u64 slab_bytes;
};

#ifdef __CUDACC__
#define B200SAT_HD __host__ __device__
#else
#define B200SAT_HD
#endif
B200SAT_HD static inline u64 layout_align(u64 x) { return (x + 255) & ~(u64)255; }

/* bytes of the var-state block when it lives in the slab (tier 1) */
static inline u64 layout_vars_bytes(u32 nv) {
    u64 n = nv;
    return layout_align(8 * n) + layout_align(4 * n) + layout_align(4 * 2 * n) + layout_align(4 * 2 * n) +
           5 * layout_align(4 * (n + 2)) + 2 * layout_align(n);
}

static inline void layout_finalize(Layout &L) {
    u64 o = 0;
    for (int h = 0; h < 2; h++) { L.off_arena[h] = o; o += layout_align(4ull * L.arena_cap); }
    for (int h = 0; h < 2; h++) { L.off_wpool[h] = o; o += layout_align(8ull * L.wpool_cap); }
    L.off_learnts = o; o += layout_align(4ull * L.learnts_cap);
    L.off_clauses = o; o += layout_align(4ull * L.clauses_cap);
    L.off_ol = o; o += layout_align(4ull * (L.nv_cap + 2));
    L.off_tc = o; o += layout_align(4ull * (L.nv_cap + 2));
    L.off_stk = o; o += layout_align(16ull * (L.nv_cap + 2));
    L.off_tmp = o; o += layout_align(4ull * L.tmp_cap);
    L.off_vars = o; if (L.tier == 1) o += layout_vars_bytes(L.nv_cap);
    L.off_simp = o; o += layout_align(4ull * L.simp_cap);
    L.off_save = o; o += layout_align(L.save_bytes);
    L.off_pf = o; o += layout_align(1024);
    L.slab_bytes = layout_align(o);
}

/* sub-layout of the slab's simplifier region (word offsets), identical on host and device */
struct SimpLayout {
    u32 o_ostart, o_olen, o_ocap, o_eheap, o_eidx, o_nocc, o_flags, o_sq, o_elits, o_esizes, o_cand, o_cls, o_pn,
        o_rbuf, o_roff, o_tmp, o_opool[2];
    u32 sq_cap, elits_cap, esizes_cap, cand_cap, rbuf_cap, roff_cap, opool_cap, tmp_cap;
    u32 total;
};
B200SAT_HD static inline SimpLayout simp_layout(u32 nv, u32 max_clauses, u64 max_lits, u32 max_clause_len) {
    SimpLayout s;
    u32 o = 0;
    const u32 M = max_clauses, L = (u32)max_lits;
    s.o_ostart = o; o += nv;
    s.o_olen = o; o += nv;
    s.o_ocap = o; o += nv;
    s.o_eheap = o; o += nv;
    s.o_eidx = o; o += nv;
    s.o_nocc = o; o += 2 * nv;
    s.o_flags = o; o += (3 * nv + 3) / 4 + 1;
    s.sq_cap = 8 * M + 4096; s.o_sq = o; o += s.sq_cap;
    s.elits_cap = 6 * L + 8 * nv + 1024; s.o_elits = o; o += s.elits_cap;
    s.esizes_cap = 3 * M + 2 * nv + 64; s.o_esizes = o; o += s.esizes_cap;
    s.cand_cap = 3 * M + 256; s.o_cand = o; o += s.cand_cap;
    s.o_cls = o; o += s.cand_cap;
    s.o_pn = o; o += s.cand_cap;
    s.rbuf_cap = 4 * L + 64 * (max_clause_len + 32) + 4096; s.o_rbuf = o; o += s.rbuf_cap;
    s.roff_cap = 3 * M + 256; s.o_roff = o; o += s.roff_cap;
    s.tmp_cap = 4 * max_clause_len + 256; s.o_tmp = o; o += s.tmp_cap;
    s.opool_cap = 6 * L + 16 * nv + 4096;
    s.o_opool[0] = o; o += s.opool_cap;
    s.o_opool[1] = o; o += s.opool_cap;
    s.total = o;
    /* the offsets are u32 words: a layout that does not fit is reported as total = 0 (simp_cap 0 -> ST_MEM_SIMP /
     * B200SAT_E_MEM) instead of silently wrapping and aliasing regions */
    {
        const u64 n64 = nv, M64 = max_clauses, L64 = max_lits, C64 = max_clause_len;
        const u64 tot = 7 * n64 + (3 * n64 + 3) / 4 + 1 + (8 * M64 + 4096) + (6 * L64 + 8 * n64 + 1024) + (3 * M64 + 2 * n64 + 64) +
                        3 * (3 * M64 + 256) + (4 * L64 + 64 * (C64 + 32) + 4096) + (3 * M64 + 256) + (4 * C64 + 256) +
                        2 * (6 * L64 + 16 * n64 + 4096);
        if (tot > 0xFFFFFFF0ull || max_lits > 0xFFFFFFF0ull) s.total = 0;
    }
    return s;
}

struct KParams {
    const u32 *inst_clause_off;
    const u32 *clause_off;
    const u32 *lits;
    const u32 *work;     /* optional list of instance ids to solve (retry pass); NULL = identity */
    u32 n_work;
    u32 *queue;          /* atomic work counter */
    b200sat_result *results;
    u8 *models;
    u32 model_stride;
    u8 *slabs;
    Layout L;
    b200sat_settings st;
    i32 kind;
    i32 flags;
    u32 attempt;
    /* portfolio mode: every work item is a diversified replica of instance 0 */
    u32 portfolio;            /* 0 = batch mode, else number of replicas launched on this GPU          */
    u32 replica_base;         /* global index of this GPU's first replica                              */
    u32 share;                /* PF_SHARE: exchange short learnts, PF_IGNORE_DONE: run every replica out */
    u32 n_rings, my_ring;     /* clause-exchange rings, one per GPU (peer-mapped over NVLink)          */
    u32 ring_words;           /* payload capacity of one ring                                          */
    u32 *rings[8];
    const b200sat_settings *replica_settings; /* [portfolio] diversified settings                      */
    /* instance-resident scheduling (round 2): every work item owns slab `w - chunk_base`; a device work ring hands
     * runnable items to warps; an item that used up its conflict slice while others wait is parked and re-queued */
    u32 chunk_base;           /* first work index of the chunk the slabs hold                          */
    u32 *qctl;                /* [QC_HEAD, QC_TAIL, QC_DONE, QC_TOTAL, ...] work-ring control words     */
    unsigned long long *ring; /* work ring: (ticket + 1) << 32 | work index                            */
    u32 ring_mask;            /* ring capacity - 1 (power of two >= 2 * items)                         */
    u32 slice_conflicts;      /* conflicts an instance may run before it yields to waiting ones (0 = never) */
    long long conflict_budget, propagation_budget; /* budget.rs:5-34; < 0 = none                       */
    const volatile int *interrupt;                 /* device-visible async interrupt flag or NULL      */
    /* Solver::new_var(upol, dvar) (sat.rs:31) for the resident batch, or NULL = (None, true) for every var:
     * var_flags[inst_var_off[i] + v]: VARF_* bits; inst_var_off[i + 1] - inst_var_off[i] = number of vars created */
    const u8 *var_flags;
    const u32 *inst_var_off;
    /* optional, with var_flags: clause_nvars[c] = vars that existed when clause c was added (new_var calls interleaved
     * with add_clause, as the reference's own DIMACS loader does); null = every var was created before the first clause */
    const u32 *clause_nvars;
    const u32 *assumps;       /* assumptions of solve_limited (sat.rs:34), shared by every instance of the batch */
    u32 n_assumps;
    u8 *parked;               /* [item] 1: the item's state is resident and resumable (PH_READY), 0: not */
    /* straggler rescue (free-running mode, no reference analogue): once the work ring is empty, idle warps
     * clone the pristine post-ingest state of a still-running instance into their own slab and search it as diversified
     * helper replicas that exchange short learnt clauses with each other through the running warp's rescue ring; the
     * original ("main") replica exports but never imports, so it stays on the reference trace.  Whoever finishes first
     * claims the instance. */
    u32 rescue;               /* 0 = off (trace mode) */
    u32 *run_table;           /* [search warp] item it runs as main replica | 0x80000000, or 0 */
    u32 *rrings;              /* [search warp][RING_PAYLOAD + rr_words] rescue rings */
    u32 rr_words;
    const u8 *pristine;       /* copy of every slab as the ingest kernel left it */
    u32 n_div;                /* diversified settings available in replica_settings[] */
    double *progress_buf;     /* CLI search-table rows (search.rs:289-301), instrumented kernel only: per item
                                 [n_rows][8 values per row ...], or NULL */
    u32 progress_cap;         /* rows per item */
    u32 *trace_buf;           /* K2: per-item event record of the search prefix (instrumented kernel only), or NULL */
    u32 trace_cap;            /* words per item in trace_buf                                             */
    u32 import_cap;           /* portfolio: clauses one replica imports per restart boundary (0 = no cap) */
    u32 ring_epoch;           /* portfolio: stamps ring entries of this run                             */
};
enum : u32 { QC_HEAD = 0, QC_TAIL = 1, QC_DONE = 2, QC_TOTAL = 3, QC_SUSPENDS = 4, QC_HEAD0 = 5 /* items handed out statically: warp i starts with ring slot i */, QC_RESCUED = 6 /* instances decided by a helper replica */,
             QC_T_START = 8, QC_T_DRAIN = 10, QC_T_END = 12 /* u64 %globaltimer ns: search kernel start, first warp that found the
                                                              ring empty, last warp exit */, QC_WORDS = 16 };
/* life cycle of a resident instance (Scal::phase) */
enum : u32 { PH_NEW = 0, PH_READY = 1, PH_SEARCH = 2, PH_DONE = 3 };
/* K2 event record (bcp_replay): header words, then events: a decision is its literal; a conflict is
 * TR_CONFLICT | n, bt level, the n learnt literals; a restart is TR_RESTART */
enum : u32 { TR_WORDS = 0, TR_PROPS_LO = 1, TR_PROPS_HI = 2, TR_TRAFFIC = 4 /* 6 x u64 at the end of the record */,
             TR_PROPS0_LO = 16, TR_PROPS0_HI = 17, TR_TRAFFIC0 = 18 /* 6 x u64 when the search started */, TR_EVENTS = 32,
             TR_CONFLICT = 0x80000000u, TR_RESTART = 0xC0000000u };

enum : u32 { PF_SHARE = 1, PF_IGNORE_DONE = 2, PF_SHARE_LBD2 = 4 };
enum : u32 { VARF_UPOL_SET = 1, VARF_UPOL_VAL = 2, VARF_NOT_DECISION = 4 };
/* ring header words */
enum : u32 { RING_HEAD = 0, RING_DONE = 1, RING_EXPORTS = 2, RING_ITEM = 3 /* rescue: item the ring belongs to | 0x80000000 */,
             RING_HELPERS = 4 /* rescue: helper replicas that joined */, RING_PAYLOAD = 16 };
static const u32 ST_CANCELLED = 20;

} // namespace b200sat
#endif
