/*
 * cdcl_plan.h -- host-side sizing of the per-warp working slab and choice of
 * the var-state tier.  Shared by the C-ABI host (b200sat_api.cu) and the
 * test-only emulator harness so both size memory identically.
 */
#ifndef B200SAT_CDCL_PLAN_H
#define B200SAT_CDCL_PLAN_H
#include <string.h>
#include "cdcl_types.h"

namespace b200sat {

struct BatchShape {
    u32 max_vars;        /* largest (max var id + 1) over the instances  */
    u32 max_clauses;     /* largest clause count of one instance         */
    u64 max_lits;        /* largest literal count of one instance        */
    u32 max_clause_len;  /* longest clause                               */
};

/* shared-memory tiers compiled into the library (template instantiations) */
static const int kSmemTiers[] = {128, 256, 1024};
static const int kNumSmemTiers = 3;

/* smallest shared-memory tier that holds `nvars`, or 0 for the global tier */
static inline int pick_tier_nv(u32 nvars, u32 max_clauses) {
    /* u16 watch-list lengths: a single list of more than 65,520 watchers fails with ST_MEM_LIST and the instance is
     * retried on the global tier; with up to ~200k clauses spread over the literals that practically never happens */
    if (max_clauses > 200000) return 0;
    for (int i = 0; i < kNumSmemTiers; i++) if (nvars <= (u32)kSmemTiers[i]) return kSmemTiers[i];
    return 0;
}

static inline u32 sat_u32(u64 x) { return x > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (u32)x; }

/* attempt 0 is the first try; every retry multiplies the growable regions by 4 */
static inline Layout plan_layout(const BatchShape &sh, int tier_nv, int kind, u32 attempt,
                                 u64 arena_words_override, u64 wpool_override, u32 save_bytes = 0) {
    Layout L;
    memset(&L, 0, sizeof L);
    const u64 nv = sh.max_vars ? sh.max_vars : 1;
    const u64 grow = 1ull << (2 * attempt);
    const u64 orig_words = sh.max_lits + 4ull * sh.max_clauses;
    L.tier = tier_nv ? 0 : 1;
    L.nv_cap = tier_nv ? (u32)tier_nv : (u32)nv;
    u64 arena = (orig_words * 4 + 64 * nv + 32768) * grow;
    if (kind == B200SAT_SIMP) arena += orig_words * 4;
    if (arena_words_override) arena = arena_words_override * grow;
    L.arena_cap = sat_u32((arena + 3) & ~3ull);
    u64 wpool = (4 * sh.max_lits + 32 * nv + 16384) * grow;
    if (wpool_override) wpool = wpool_override * grow;
    if (wpool > (1ull << 25)) wpool = (wpool < (1ull << 26)) ? wpool : (1ull << 26) - 64; /* start<<W_SHIFT must fit u32 */
    L.wpool_cap = sat_u32(wpool);
    L.learnts_cap = sat_u32((8192 + (u64)sh.max_clauses) * grow);
    L.clauses_cap = sat_u32((u64)sh.max_clauses * (kind == B200SAT_SIMP ? 3 : 1) + 64);
    L.tmp_cap = sh.max_clause_len < 64 ? 64 : sh.max_clause_len;
    L.simp_cap = 0;
    L.simp_lits = 0;
    if (kind == B200SAT_SIMP) {
        L.simp_lits = sat_u32((sh.max_lits + 64) * grow);
        L.simp_cap = simp_layout(L.nv_cap, L.clauses_cap, L.simp_lits, L.tmp_cap).total;
    }
    L.save_bytes = save_bytes;
    layout_finalize(L);
    return L;
}

/* settings of global replica g of a portfolio (the seed alone changes nothing with the
 * default settings, so the diversification switches randomness on and varies decay / restart / phase) */
static inline void diversify_settings(const b200sat_settings *base, u32 g, b200sat_settings *out) {
    *out = *base;
    if (g == 0) return; /* replica 0 = the reference configuration */
    double seed = base->random_seed + 7919.0 * (double)g;
    while (seed >= 2147483646.0) seed -= 2147483646.0;
    out->random_seed = seed < 1.0 ? 1.0 : seed;
    switch (g % 4) {
    case 1: out->random_var_freq = 0.02; break;
    case 2: out->rnd_init_act = 1; break;
    case 3: out->random_var_freq = 0.05; out->rnd_init_act = 1; break;
    default: out->rnd_init_act = 1; out->random_var_freq = 0.01; break;
    }
    static const double decays[4] = {0.0, 0.92, 0.88, 0.97};
    static const double first_scale[4] = {1.0, 0.5, 2.0, 0.3};
    if ((g / 4) % 4 != 0) out->var_decay = decays[(g / 4) % 4];
    out->restart_first = base->restart_first * first_scale[(g / 16) % 4];
    if ((g / 64) % 2 == 1) out->phase_saving = 1;
}

} // namespace b200sat
#endif
