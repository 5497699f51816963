/*
 * simp_warp.cuh -- SimpSolver path: reference-order preprocessor (backward
 * subsumption, self-subsuming strengthening, bounded variable elimination) run
 * on the device in front of the warp-cooperative search.
 *
 * The simplifier's job order is inherently sequential and trace-relevant
 * (queue FIFO order, elimination-heap tie breaking, dirty
 * occurrence lists, GC timing all change which clause is looked at next), so
 * round 1 executes it on lane 0 of the instance's warp, operating directly on
 * the warp's slab (same clause arena, watch lists and var state the search
 * kernel code uses); the other lanes wait at the reconvergence point and the
 * warp-cooperative search starts from the state it leaves behind.  The
 * data-parallel preprocessing kernels for large single instances (SURVEY.md
 * K4) are a separate path.
 *
 * Reference anchors (relative to /root/reference/src/sat/minisat):
 *   SimpSolver::{new_var,add_clause,preprocess,solve_limited}  ../minisat.rs:123-279
 *   Simplificator::{add_clause,solve_limited,eliminate}        search/simplify.rs:139-274
 *   strengthen_clause / eliminate_var / backward_subsumption_check  simplify.rs:276-523
 *   try_garbage_collect / off / on                             simplify.rs:526-549
 *   subsumes / unit_subsumes                                   simplify/subsumes.rs:10-67
 *   SubsumptionQueue                                           simplify/subsumption_queue.rs
 *   ElimQueue / OccLists / ElimOcc                             simplify/elim_queue.rs
 *   ElimClauses                                                simplify/elim_clauses.rs
 *   merge / calc_abstraction                                   ../formula/util.rs:5-11,45-77
 *   Searcher::add_clause / propagate / gc                      search.rs:342-386, watches.rs, clause.rs
 */
#ifndef B200SAT_SIMP_WARP_CUH
#define B200SAT_SIMP_WARP_CUH

#include "cdcl_warp.cuh"

namespace b200sat {

enum : u32 { SUB_DIFFERENT = 0, SUB_EXACT = 1, SUB_LITSIGN = 2 };
enum : u32 { ADD_UNSAT = 0, ADD_CONSUMED = 1, ADD_ADDED = 2 };

/* Everything in this struct runs on ONE lane (lane 0 of the warp). */
template <class WS, bool COUNT> struct SeqSimp {
    typedef typename WS::idx_t idx_t;
    WS &S;
    SimpLayout SL;
    u32 *base;
    u32 *ostart, *olen, *ocap, *sq, *elits, *esizes, *cand, *cls, *pn, *rbuf, *roff, *tmpb, *opool;
    idx_t *eheap, *eidx; /* shared memory (u16) in the shared-memory tiers, slab (u32) in the global tier */
    i32 *nocc;
    u8 *odirty, *opresent, *touched;
    u32 eheap_n, n_touched, sq_head, sq_tail, bwdsub_assigns, elits_n, esizes_n, opool_top, cur_opool, n_created;
    u32 eliminated_vars;
    u64 t_vis, t_kept, t_deref, t_tail, t_moves, t_impl;

    DEV SeqSimp(WS &s) : S(s) {}
    DEV Scal &sc() const { return S.sc(); }
    DEV bool failed() const { return sc().status != ST_OK; }
    DEV void fail(u32 code) { if (sc().status == ST_OK) sc().status = code; }
    DEV u32 *arena() const { return S.arena; }

    DEV void bind_ptrs() {
        base = (u32 *)(S.slab + S.P.L.off_simp);
        SL = simp_layout(S.P.L.nv_cap, S.P.L.clauses_cap, S.P.L.simp_lits, S.P.L.tmp_cap);
        elits = base + SL.o_elits; esizes = base + SL.o_esizes;
    }
    DEV void bind_heap() {
        if (WS::VT_SIMP_SMEM) {
            char *sm = (char *)S.V.simp_smem();
            nocc = (i32 *)sm;
            eheap = (idx_t *)(sm + 8 * (size_t)S.V.cap());
            eidx = eheap + S.V.cap();
        } else {
            eheap = (idx_t *)(base + SL.o_eheap); eidx = (idx_t *)(base + SL.o_eidx); nocc = (i32 *)(base + SL.o_nocc);
        }
    }
    DEV void bind() {
        base = (u32 *)(S.slab + S.P.L.off_simp);
        SL = simp_layout(S.P.L.nv_cap, S.P.L.clauses_cap, S.P.L.simp_lits, S.P.L.tmp_cap);
        ostart = base + SL.o_ostart; olen = base + SL.o_olen; ocap = base + SL.o_ocap;
        bind_heap();
        u8 *fl = (u8 *)(base + SL.o_flags);
        odirty = fl; opresent = fl + S.P.L.nv_cap; touched = fl + 2 * S.P.L.nv_cap;
        sq = base + SL.o_sq; elits = base + SL.o_elits; esizes = base + SL.o_esizes;
        cand = base + SL.o_cand; cls = base + SL.o_cls; pn = base + SL.o_pn;
        rbuf = base + SL.o_rbuf; roff = base + SL.o_roff; tmpb = base + SL.o_tmp;
        opool = base + SL.o_opool[0];
        eheap_n = n_touched = sq_head = sq_tail = bwdsub_assigns = elits_n = esizes_n = opool_top = cur_opool = n_created = 0;
        eliminated_vars = 0;
        if (S.lane == 0) sc().simp_cur_opool = 0;
        t_vis = t_kept = t_deref = t_tail = t_moves = t_impl = 0;
    }

    /* ------------------------------------------------------------ sequential watch-list primitives */
    DEV bool seq_compact_wpool(u32 reserve) {
        const u32 nl = 2 * S.nvars;
        const u32 h = sc().cur_wpool ^ 1;
        uint2 *dst = S.wpool_half(h);
        u32 top = 0;
        for (u32 l = 0; l < nl; l++) {
            u32 word = S.V.wstart()[l], len = S.V.wlen()[l];
            u32 want = len + (len >> 1) + 2;
            u32 ncapc = ceil_log2(want < 4 ? 4 : want);
            u32 ncap = 1u << ncapc;
            if (top + ncap + reserve > S.P.L.wpool_cap) { fail(ST_MEM_WPOOL); return false; }
            const uint2 *src = S.wpool + (word >> W_SHIFT);
            for (u32 i = 0; i < len; i++) dst[top + i] = src[i];
            S.V.wstart()[l] = (top << W_SHIFT) | (word & W_DIRTY) | ncapc;
            top += ncap;
        }
        sc().wpool_top = top;
        sc().cur_wpool = h;
        S.wpool = dst;
        return true;
    }
    DEV bool seq_push_watch(u32 l, u32 cref, u32 blocker) {
        u32 word = S.V.wstart()[l];
        u32 len = S.V.wlen()[l];
        u32 capc = word & W_CAP_MASK;
        if (len == (1u << capc)) {
            u32 ncap = 2u << capc;
            if (ncap > WS::VT_MAXLEN) { fail(ST_MEM_LIST); return false; }
            if (sc().wpool_top + ncap > S.P.L.wpool_cap) {
                if (!seq_compact_wpool(ncap + 64)) return false;
                word = S.V.wstart()[l];
            }
            u32 top = sc().wpool_top;
            const uint2 *src = S.wpool + (word >> W_SHIFT);
            uint2 *dst = S.wpool + top;
            for (u32 i = 0; i < len; i++) dst[i] = src[i];
            word = (top << W_SHIFT) | (word & W_DIRTY) | (capc + 1);
            S.V.wstart()[l] = word;
            sc().wpool_top = top + ncap;
        }
        S.wpool[(word >> W_SHIFT) + len] = make_uint2(cref, blocker);
        S.V.wlen()[l] = (idx_t)(len + 1);
        return true;
    }
    DEV bool seq_attach(u32 cr) { /* watches.rs:39-44 */
        u32 c0 = arena()[cr + 1], c1 = arena()[cr + 2];
        if (!seq_push_watch(c0 ^ 1, cr, c1)) return false;
        return seq_push_watch(c1 ^ 1, cr, c0);
    }
    DEV void unwatch_lazy(u32 cr) { /* watches.rs:52-56 */
        S.V.wstart()[arena()[cr + 1] ^ 1] |= W_DIRTY;
        S.V.wstart()[arena()[cr + 2] ^ 1] |= W_DIRTY;
    }
    DEV void unwatch_strict(u32 cr) { /* watches.rs:46-50 */
        for (int i = 0; i < 2; i++) {
            u32 l = arena()[cr + 1 + i] ^ 1;
            uint2 *ws = S.wpool + (S.V.wstart()[l] >> W_SHIFT);
            u32 n = S.V.wlen()[l], j = 0;
            for (u32 k = 0; k < n; k++) if (ws[k].x != cr) ws[j++] = ws[k];
            S.V.wlen()[l] = (idx_t)j;
        }
    }
    DEV void seq_assign(u32 l, u32 reason) {
        u32 v = l >> 1;
        S.V.assign()[v] = (u8)((l & 1) ^ 1);
        S.V.level()[v] = (idx_t)S.nlevels;
        S.V.reason()[v] = reason;
        S.V.trail()[S.trail_n] = (idx_t)l;
        S.trail_n++;
    }
    /* watches.rs:64-152, scalar */
    DEV u32 seq_propagate() {
        u32 *ar = arena();
        u32 nprops = 0;
        u32 confl = CREF_UNDEF;
        while (S.qhead < S.trail_n) {
            const u32 p = S.V.trail()[S.qhead];
            S.qhead++;
            nprops++;
            const u32 not_p = p ^ 1;
            const u32 n = S.V.wlen()[p];
            const bool dirty = (S.V.wstart()[p] & W_DIRTY) != 0;
            uint2 *ws = S.wpool + (S.V.wstart()[p] >> W_SHIFT);
            u32 head = 0, tail = 0;
            while (head < n) {
                uint2 w = ws[head++];
                const u32 cr = w.x;
                const u32 hdr = ar[cr];
                if (hdr & H_DELETED) continue;
                if (COUNT) t_vis++;
                if (S.is_true(w.y)) { ws[tail++] = w; if (COUNT) t_kept++; continue; }
                if (COUNT) t_deref++;
                if (ar[cr + 1] == not_p) { ar[cr + 1] = ar[cr + 2]; ar[cr + 2] = not_p; }
                const u32 first = ar[cr + 1];
                if (first != w.y && S.is_true(first)) { ws[tail++] = make_uint2(cr, first); if (COUNT) t_kept++; continue; }
                const u32 len = hdr >> H_SHIFT;
                bool moved = false;
                for (u32 k = 2; k < len; k++) {
                    if (COUNT) t_tail++;
                    u32 lk = ar[cr + 1 + k];
                    if (!S.is_false(lk)) {
                        if (!seq_push_watch(lk ^ 1, cr, first)) return CREF_UNDEF;
                        ws = S.wpool + (S.V.wstart()[p] >> W_SHIFT);
                        ar[cr + 2] = lk;
                        ar[cr + 1 + k] = not_p;
                        if (COUNT) t_moves++;
                        moved = true;
                        break;
                    }
                }
                if (moved) continue;
                ws[tail++] = make_uint2(cr, first);
                if (COUNT) t_kept++;
                if (S.is_false(first)) {
                    while (head < n) {
                        uint2 r = ws[head++];
                        if (dirty && (ar[r.x] & H_DELETED)) continue;
                        ws[tail++] = r;
                    }
                    confl = cr;
                    break;
                }
                seq_assign(first, cr);
                if (COUNT) t_impl++;
            }
            S.V.wlen()[p] = (idx_t)tail;
            if (dirty) S.V.wstart()[p] &= ~(u32)W_DIRTY;
            if (confl != CREF_UNDEF) break;
        }
        sc().propagations += nprops;
        return confl;
    }

    /* ------------------------------------------------------------ clause arena (scalar) */
    DEV u32 seq_alloc(const u32 *lits, u32 len, u32 flags) {
        u32 hdr = (len << H_SHIFT) | flags;
        u32 blk = WS::block_words(hdr);
        u32 top = sc().arena_top;
        if (top + blk > S.P.L.arena_cap) { fail(ST_MEM_ARENA); return CREF_UNDEF; }
        u32 *ar = arena();
        ar[top] = hdr;
        for (u32 i = 0; i < len; i++) ar[top + 1 + i] = lits[i];
        sc().arena_top = top + blk;
        sc().lc_size += WS::ref_size(hdr);
        return top;
    }
    DEV void st_add(u32 cr) { /* clause_db.rs:29-41 */
        u32 h = arena()[cr];
        if (h & H_LEARNT) { sc().num_learnts++; sc().learnts_literals += h >> H_SHIFT; }
        else { sc().num_clauses++; sc().clauses_literals += h >> H_SHIFT; }
    }
    DEV void st_del(u32 cr) { /* clause_db.rs:43-55 */
        u32 h = arena()[cr];
        if (h & H_LEARNT) { sc().num_learnts--; sc().learnts_literals -= h >> H_SHIFT; }
        else { sc().num_clauses--; sc().clauses_literals -= h >> H_SHIFT; }
    }
    DEV void remove_clause(u32 cr) { /* clause_db.rs:102-105 + clause.rs:157-162 */
        st_del(cr);
        u32 h = arena()[cr];
        arena()[cr] = h | H_DELETED;
        sc().lc_wasted += WS::ref_size(h);
    }
    DEV u32 abstraction_of(u32 cr, u32 len) {
        u32 a = 0;
        for (u32 i = 0; i < len; i++) a |= 1u << ((arena()[cr + 1 + i] >> 1) & 31);
        return a;
    }
    /* ------------------------------------------------------------ scratch decision level (scalar)
     * assignment.rs:99-101,116-130; phases are saved as top-level ones (search/util.rs:27,38,66).  Nothing has
     * left the order heap while the simplifier runs, so no variable needs returning to it. */
    DEV void seq_new_level() { S.V.lim()[S.nlevels] = (idx_t)S.trail_n; S.nlevels++; }
    DEV void seq_backtrack_ground() {
        const u32 target = (u32)S.V.lim()[GROUND];
        const i32 ps = S.st().phase_saving;
        u8 *vf = S.V.vflags();
        for (u32 k = target; k < S.trail_n; k++) {
            const u32 l = S.V.trail()[k], v = l >> 1;
            if (ps >= 1) vf[v] = (u8)((vf[v] & ~VF_POL) | ((l & 1) ? VF_POL : 0));
            S.V.assign()[v] = LB_UNDEF;
        }
        S.trail_n = target;
        S.nlevels = GROUND;
        if (S.qhead > S.trail_n) S.qhead = S.trail_n;
    }
    /* search/util.rs:16-42 (--rcheck) */
    DEV bool seq_is_implied(const u32 *c, u32 n) {
        seq_new_level();
        bool hit = false;
        for (u32 i = 0; i < n && !hit; i++) {
            const u32 l = c[i];
            if (S.is_true(l)) hit = true;
            else if (S.V.assign()[l >> 1] == LB_UNDEF) seq_assign(l ^ 1, CREF_UNDEF);
        }
        bool result = hit;
        if (!hit) result = seq_propagate() != CREF_UNDEF;
        seq_backtrack_ground();
        return result && !failed();
    }
    /* search/util.rs:44-73 (--asymm): true = the literal of `v` in `cr` can be dropped (returned in `out`) */
    DEV bool seq_asymmetric_branching(u32 v, u32 cr, u32 &out) {
        const u32 *ar = arena();
        if (ar[cr] & H_DELETED) return false;
        const u32 n = ar[cr] >> H_SHIFT;
        for (u32 i = 0; i < n; i++) if (S.is_true(ar[cr + 1 + i])) return false;
        seq_new_level();
        u32 vl = LIT_NONE;
        for (u32 i = 0; i < n; i++) {
            const u32 l = ar[cr + 1 + i];
            if ((l >> 1) == v) vl = l;
            else if (S.V.assign()[l >> 1] == LB_UNDEF) seq_assign(l ^ 1, CREF_UNDEF);
        }
        const u32 confl = seq_propagate();
        seq_backtrack_ground();
        if (failed() || confl == CREF_UNDEF) return false;
        out = vl;
        return true;
    }

    /* Searcher::add_clause search.rs:342-386 (scalar) */
    DEV u32 seq_add_clause(const u32 *src, u32 n, u32 &out_cr) {
        if (S.st().use_rcheck) { /* search.rs:343-345 */
            bool imp = seq_is_implied(src, n);
            if (failed()) return ADD_UNSAT;
            if (imp) return ADD_CONSUMED;
        }
        if (n > SL.tmp_cap) { fail(ST_MEM_SIMP); return ADD_UNSAT; }
        u32 *ps = tmpb;
        for (u32 i = 0; i < n; i++) ps[i] = src[i];
        for (u32 i = 1; i < n; i++) { /* sort */
            u32 x = ps[i], j = i;
            while (j > 0 && ps[j - 1] > x) { ps[j] = ps[j - 1]; j--; }
            ps[j] = x;
        }
        u32 m = 0;
        for (u32 i = 0; i < n; i++) if (i == 0 || ps[i] != ps[i - 1]) ps[m++] = ps[i]; /* dedup */
        u32 k = 0;
        for (u32 i = 0; i < m; i++) if (!S.is_false(ps[i])) ps[k++] = ps[i];           /* drop false */
        for (u32 i = 0; i < k; i++)
            if (S.is_true(ps[i]) || (i > 0 && ps[i - 1] == (ps[i] ^ 1))) return ADD_CONSUMED;
        if (k == 0) return ADD_UNSAT;
        if (k == 1) {
            seq_assign(ps[0], CREF_UNDEF);
            u32 confl = seq_propagate();
            if (failed()) return ADD_UNSAT;
            return confl == CREF_UNDEF ? ADD_CONSUMED : ADD_UNSAT;
        }
        const bool extra = sc().extra_clause_field != 0;
        u32 cr = seq_alloc(ps, k, extra ? (u32)H_EXTRA : 0u);
        if (cr == CREF_UNDEF) return ADD_UNSAT;
        if (extra) arena()[cr + 1 + k] = abstraction_of(cr, k);
        if (sc().n_clauses >= S.P.L.clauses_cap) { fail(ST_MEM_ARENA); return ADD_UNSAT; }
        S.clauses()[sc().n_clauses++] = cr;
        st_add(cr);
        if (!seq_attach(cr)) return ADD_UNSAT;
        out_cr = cr;
        return ADD_ADDED;
    }

    /* ------------------------------------------------------------ elimination heap, elim_queue.rs:13-76 + index_map.rs */
    DEV u64 cost(u32 v) const { return (u64)(u32)nocc[2 * v] * (u64)(u32)nocc[2 * v + 1]; }
    DEV void eh_sift_up(u32 i) {
        u32 x = eheap[i];
        u64 cx = cost(x);
        while (i > 0) {
            u32 p = (i - 1) >> 1;
            u32 y = eheap[p];
            if (cx < cost(y)) { eheap[i] = (idx_t)y; eidx[y] = (idx_t)i; i = p; } else break;
        }
        eheap[i] = (idx_t)x; eidx[x] = (idx_t)i;
    }
    DEV void eh_sift_down(u32 i) {
        u32 x = eheap[i];
        u64 cx = cost(x);
        for (;;) {
            u32 l = 2 * i + 1;
            if (l >= eheap_n) break;
            u32 r = l + 1;
            u32 c = (r < eheap_n && cost(eheap[r]) < cost(eheap[l])) ? r : l;
            if (cost(eheap[c]) < cx) { u32 y = eheap[c]; eheap[i] = (idx_t)y; eidx[y] = (idx_t)i; i = c; } else break;
        }
        eheap[i] = (idx_t)x; eidx[x] = (idx_t)i;
    }
    DEV bool eh_contains(u32 v) const { return (u32)eidx[v] != WS::VT_ABSENT; }
    DEV void eh_insert(u32 v) {
        if (eh_contains(v)) return;
        u32 place = eheap_n++;
        eheap[place] = (idx_t)v;
        eh_sift_up(place);
    }
    DEV void eh_update(u32 v) { /* index_map.rs:244-255: sift_down(place) then sift_up(place), same place */
        if (!eh_contains(v)) return;
        u32 place = eidx[v];
        eh_sift_down(place);
        eh_sift_up(place);
    }
    DEV bool eh_pop(u32 &out) {
        if (eheap_n == 0) return false;
        out = eheap[0];
        u32 last = eheap[--eheap_n];
        eidx[out] = (idx_t)WS::VT_ABSENT;
        if (eheap_n > 0) { eheap[0] = (idx_t)last; eh_sift_down(0); }
        return true;
    }
    DEV bool is_elim(u32 v) const { return (S.V.vflags()[v] & VF_ELIM) != 0; }
    DEV bool is_frozen(u32 v) const { return (S.V.vflags()[v] & VF_FROZEN) != 0; }
    DEV void update_elim_heap(u32 v) { /* elim_queue.rs:46-58 */
        if (eh_contains(v)) eh_update(v);
        else if (!is_frozen(v) && !is_elim(v) && S.V.assign()[v] == LB_UNDEF) eh_insert(v);
    }
    DEV void bump_lit_occ(u32 l, i32 d) { nocc[l] += d; eh_update(l >> 1); } /* elim_queue.rs:64-70 */

    /* ------------------------------------------------------------ occurrence lists, elim_queue.rs:85-151 */
    DEV bool occ_compact(u32 reserve) {
        u32 h = cur_opool ^ 1;
        u32 *dst = base + SL.o_opool[h];
        u32 top = 0;
        for (u32 v = 0; v < n_created; v++) {
            u32 len = olen[v];
            u32 ncap = len + (len >> 1) + 4;
            if (top + ncap + reserve > SL.opool_cap) { fail(ST_MEM_SIMP); return false; }
            const u32 *src = opool + ostart[v];
            for (u32 i = 0; i < len; i++) dst[top + i] = src[i];
            ostart[v] = top; ocap[v] = ncap;
            top += ncap;
        }
        opool = dst; cur_opool = h; opool_top = top;
        sc().simp_cur_opool = h;
        return true;
    }
    DEV bool occ_push(u32 v, u32 cr) {
        if (olen[v] == ocap[v]) {
            u32 ncap = ocap[v] ? 2 * ocap[v] : 8;
            if (opool_top + ncap > SL.opool_cap) { if (!occ_compact(ncap + 64)) return false; }
            const u32 *src = opool + ostart[v];
            u32 *dst = opool + opool_top;
            for (u32 i = 0; i < olen[v]; i++) dst[i] = src[i];
            ostart[v] = opool_top; ocap[v] = ncap;
            opool_top += ncap;
        }
        opool[ostart[v] + olen[v]] = cr;
        olen[v]++;
        return true;
    }
    DEV void occ_lookup(u32 v) { /* elim_queue.rs:118-125: clean lazily */
        if (odirty[v]) {
            u32 *o = opool + ostart[v];
            u32 n = olen[v], j = 0;
            for (u32 i = 0; i < n; i++) if (!(arena()[o[i]] & H_DELETED)) o[j++] = o[i];
            olen[v] = j;
            odirty[v] = 0;
        }
    }
    DEV void smudge_clause(u32 cr) { /* elim_queue.rs:182-188 */
        u32 len = arena()[cr] >> H_SHIFT;
        for (u32 i = 0; i < len; i++) {
            u32 l = arena()[cr + 1 + i];
            bump_lit_occ(l, -1);
            update_elim_heap(l >> 1);
            odirty[l >> 1] = 1;
        }
    }
    DEV void remove_lit(u32 l, u32 cr) { /* elim_queue.rs:190-194 */
        bump_lit_occ(l, -1);
        update_elim_heap(l >> 1);
        u32 v = l >> 1;
        u32 *o = opool + ostart[v];
        u32 n = olen[v], j = 0;
        for (u32 i = 0; i < n; i++) if (o[i] != cr) o[j++] = o[i];
        olen[v] = j;
    }

    /* ------------------------------------------------------------ subsumption queue (FIFO ring), subsumption_queue.rs */
    DEV u32 sq_len() const { return sq_tail - sq_head; }
    DEV bool sq_push(u32 cr) {
        if (sq_len() >= SL.sq_cap) { fail(ST_MEM_SIMP); return false; }
        sq[sq_tail % SL.sq_cap] = cr;
        sq_tail++;
        return true;
    }

    /* ------------------------------------------------------------ var creation + ingest */
    DEV void init_var(u32 v) { /* simplify.rs:134-137 + elim_queue.rs:26-32,162-168 */
        ostart[v] = 0; olen[v] = 0; ocap[v] = 0;
        odirty[v] = 0; opresent[v] = 1; touched[v] = 0;
        nocc[2 * v] = 0; nocc[2 * v + 1] = 0;
        eidx[v] = (idx_t)WS::VT_ABSENT;
        n_created = v + 1;
        eh_insert(v);
    }
    /* Simplificator::add_clause simplify.rs:139-163; false = Err */
    DEV bool add_clause(const u32 *ps, u32 n) {
        u32 cr = CREF_UNDEF;
        u32 r = seq_add_clause(ps, n, cr);
        if (r == ADD_UNSAT) return false;
        if (r == ADD_CONSUMED) return true;
        if (!sq_push(cr)) return false;
        u32 len = arena()[cr] >> H_SHIFT;
        for (u32 i = 0; i < len; i++) { /* ElimOcc::add_clause elim_queue.rs:175-180 */
            u32 l = arena()[cr + 1 + i];
            if (!occ_push(l >> 1, cr)) return false;
            bump_lit_occ(l, 1);
        }
        for (u32 i = 0; i < len; i++) { touched[arena()[cr + 1 + i] >> 1] = 1; n_touched++; } /* simplify.rs:74-79 */
        return true;
    }

    /* Touched::enqueue_touched_clauses simplify.rs:82-111 */
    DEV bool enqueue_touched_clauses() {
        if (n_touched == 0) return true;
        u32 *ar = arena();
        for (u32 q = sq_head; q != sq_tail; q++) ar[sq[q % SL.sq_cap]] |= H_TOUCHED;
        for (u32 v = 0; v < n_created; v++) {
            if (!touched[v] || is_elim(v)) continue;
            occ_lookup(v);
            const u32 n = olen[v];
            for (u32 i = 0; i < n; i++) {
                u32 cr = opool[ostart[v] + i];
                if (!(ar[cr] & H_TOUCHED)) { if (!sq_push(cr)) return false; ar[cr] |= H_TOUCHED; }
            }
            touched[v] = 0;
        }
        for (u32 q = sq_head; q != sq_tail; q++) ar[sq[q % SL.sq_cap]] &= ~(u32)H_TOUCHED;
        n_touched = 0;
        return true;
    }

    DEV bool try_propagate(u32 p) { /* simplify.rs:553-565 */
        if (S.is_false(p)) return false;
        if (S.V.assign()[p >> 1] == LB_UNDEF) seq_assign(p, CREF_UNDEF);
        u32 confl = seq_propagate();
        return !failed() && confl == CREF_UNDEF;
    }
    /* simplify.rs:276-301 + 567-591 */
    DEV bool strengthen_clause(u32 cr, u32 l) {
        u32 *ar = arena();
        u32 len = ar[cr] >> H_SHIFT;
        if (len == 2) {
            smudge_clause(cr);
            unwatch_lazy(cr);
            remove_clause(cr);
            u32 unit = (l == ar[cr + 1]) ? ar[cr + 2] : ar[cr + 1];
            return try_propagate(unit);
        }
        unwatch_strict(cr);
        st_del(cr);
        u32 k = 0;
        while (k < len && ar[cr + 1 + k] != l) k++;
        if (k == len) { fail(ST_INTERNAL); return false; }
        u32 h = ar[cr];
        for (; k + 1 < len; k++) ar[cr + 1 + k] = ar[cr + 2 + k];
        u32 nlen = len - 1;
        h = (nlen << H_SHIFT) | (h & ((1u << H_SHIFT) - 1));
        if (!(h & H_LEARNT)) { h |= H_EXTRA; ar[cr] = h; ar[cr + 1 + nlen] = abstraction_of(cr, nlen); }
        else { ar[cr + 1 + nlen] = ar[cr + 1 + len]; ar[cr] = h; }
        st_add(cr);
        if (!seq_attach(cr)) return false;
        remove_lit(l, cr);
        return true;
    }

    DEV u32 subsumes(u32 a, u32 b, u32 &out) { /* subsumes.rs:10-49 */
        const u32 *ar = arena();
        u32 ha = ar[a], hb = ar[b];
        u32 na = ha >> H_SHIFT, nb = hb >> H_SHIFT;
        if (nb < na) return SUB_DIFFERENT;
        if ((ha & H_EXTRA) && !(ha & H_LEARNT) && (hb & H_EXTRA) && !(hb & H_LEARNT))
            if ((ar[a + 1 + na] & ~ar[b + 1 + nb]) != 0) return SUB_DIFFERENT;
        u32 ret = SUB_EXACT;
        for (u32 i = 0; i < na; i++) {
            u32 lit = ar[a + 1 + i];
            bool found = false;
            for (u32 j = 0; j < nb; j++) {
                u32 cur = ar[b + 1 + j];
                if (lit == cur) { found = true; break; }
                else if (lit == (cur ^ 1)) {
                    if (ret == SUB_EXACT) { ret = SUB_LITSIGN; out = lit; found = true; break; }
                    else return SUB_DIFFERENT;
                }
            }
            if (!found) return SUB_DIFFERENT;
        }
        return ret;
    }
    DEV u32 unit_subsumes(u32 unit, u32 b, u32 &out) { /* subsumes.rs:51-67 */
        const u32 *ar = arena();
        u32 hb = ar[b], nb = hb >> H_SHIFT;
        if ((hb & H_EXTRA) && !(hb & H_LEARNT) && ((1u << ((unit >> 1) & 31)) & ~ar[b + 1 + nb]) != 0) return SUB_DIFFERENT;
        for (u32 j = 0; j < nb; j++) {
            u32 y = ar[b + 1 + j];
            if (unit == y) return SUB_EXACT;
            else if (unit == (y ^ 1)) { out = unit; return SUB_LITSIGN; }
        }
        return SUB_DIFFERENT;
    }

    /* simplify.rs:423-523; false = Err */
    DEV bool backward_subsumption_check() {
        const i32 sub_lim = S.st().subsumption_lim;
        for (;;) {
            bool is_clause = false, have = false;
            u32 sub_cr = CREF_UNDEF, unit = 0;
            for (;;) { /* SubsumptionQueue::pop subsumption_queue.rs:23-42 */
                if (sq_len() > 0) {
                    u32 cr = sq[sq_head % SL.sq_cap];
                    sq_head++;
                    if (!(arena()[cr] & H_DELETED)) { is_clause = true; sub_cr = cr; have = true; break; }
                } else {
                    u32 ground_end = S.nlevels > 1 ? (u32)S.V.lim()[1] : S.trail_n;
                    if (bwdsub_assigns < ground_end) { unit = S.V.trail()[bwdsub_assigns++]; have = true; }
                    break;
                }
            }
            if (!have) break;
            u32 lookup_var;
            if (!is_clause) lookup_var = unit >> 1;
            else { /* simplify.rs:459-476: occurrence list with the fewest entries, dead ones included */
                u32 n = arena()[sub_cr] >> H_SHIFT;
                u32 best = arena()[sub_cr + 1] >> 1;
                for (u32 i = 1; i < n; i++) {
                    u32 v = arena()[sub_cr + 1 + i] >> 1;
                    if (olen[v] < olen[best]) best = v;
                }
                lookup_var = best;
            }
            occ_lookup(lookup_var);
            const u32 nc = olen[lookup_var];
            if (nc > SL.cand_cap) { fail(ST_MEM_SIMP); return false; }
            for (u32 i = 0; i < nc; i++) cand[i] = opool[ostart[lookup_var] + i]; /* snapshot, simplify.rs:478 */
            for (u32 ci = 0; ci < nc; ci++) {
                u32 cj = cand[ci];
                if (is_clause) {
                    if (arena()[sub_cr] & H_DELETED) break;
                    if (cj == sub_cr) continue;
                }
                u32 hj = arena()[cj];
                if ((hj & H_DELETED) || !(sub_lim == -1 || (i32)(hj >> H_SHIFT) < sub_lim)) continue;
                u32 l = 0;
                u32 sub = is_clause ? subsumes(sub_cr, cj, l) : unit_subsumes(unit, cj, l);
                if (sub == SUB_EXACT) {
                    smudge_clause(cj);
                    unwatch_lazy(cj);
                    remove_clause(cj);
                } else if (sub == SUB_LITSIGN) {
                    if (!sq_push(cj)) return false;
                    if (!strengthen_clause(cj, l ^ 1)) return false;
                }
            }
        }
        return true;
    }

    /* formula/util.rs:45-77; false = tautology.  Result in out[0..n_out) */
    DEV bool merge(u32 v, u32 pr, u32 qr, u32 *out, u32 &n_out) {
        const u32 *ar = arena();
        u32 np = ar[pr] >> H_SHIFT, nq = ar[qr] >> H_SHIFT;
        const u32 *longer = ar + pr + 1, *shorter = ar + qr + 1;
        u32 nl = np, ns = nq;
        if (np < nq) { longer = ar + qr + 1; nl = nq; shorter = ar + pr + 1; ns = np; }
        n_out = 0;
        for (u32 i = 0; i < ns; i++) {
            u32 q = shorter[i];
            if ((q >> 1) != v) {
                bool ok = true;
                for (u32 j = 0; j < nl; j++) {
                    if ((longer[j] >> 1) == (q >> 1)) {
                        if (longer[j] == (q ^ 1)) return false;
                        ok = false;
                        break;
                    }
                }
                if (ok) out[n_out++] = q;
            }
        }
        for (u32 j = 0; j < nl; j++) if ((longer[j] >> 1) != v) out[n_out++] = longer[j];
        return true;
    }
    DEV bool mk_elim_clause(u32 v, u32 cr) { /* elim_clauses.rs:25-41 */
        const u32 *ar = arena();
        u32 n = ar[cr] >> H_SHIFT;
        if (elits_n + n > SL.elits_cap || esizes_n + 1 > SL.esizes_cap) { fail(ST_MEM_SIMP); return false; }
        u32 first = elits_n, index = 0;
        for (u32 i = 0; i < n; i++) if ((ar[cr + 1 + i] >> 1) == v) { index = i; break; }
        for (u32 i = 0; i < n; i++) elits[elits_n++] = ar[cr + 1 + i];
        esizes[esizes_n++] = elits_n;
        u32 t = elits[first]; elits[first] = elits[first + index]; elits[first + index] = t;
        return true;
    }
    DEV bool mk_elim_unit(u32 x) { /* elim_clauses.rs:20-23 */
        if (elits_n + 1 > SL.elits_cap || esizes_n + 1 > SL.esizes_cap) { fail(ST_MEM_SIMP); return false; }
        elits[elits_n++] = x;
        esizes[esizes_n++] = elits_n;
        return true;
    }

    /* simplify.rs:338-420; false = Err */
    DEV bool eliminate_var(u32 v) {
        occ_lookup(v);
        const u32 nc = olen[v];
        if (nc > SL.cand_cap) { fail(ST_MEM_SIMP); return false; }
        for (u32 i = 0; i < nc; i++) cls[i] = opool[ostart[v] + i];
        u32 npos = 0, nneg = 0; /* pos grows from the front of pn, neg from the back */
        for (u32 i = 0; i < nc; i++) {
            u32 cr = cls[i];
            u32 n = arena()[cr] >> H_SHIFT;
            for (u32 k = 0; k < n; k++) {
                u32 l = arena()[cr + 1 + k];
                if ((l >> 1) == v) { if (l & 1) { nneg++; pn[SL.cand_cap - nneg] = cr; } else pn[npos++] = cr; break; }
            }
        }
        const u64 max_resolvents = S.st().grow + npos + nneg;
        const i32 cl_lim = S.st().clause_lim;
        u32 nres = 0, rtop = 0;
        for (u32 a = 0; a < npos; a++)
            for (u32 b = 0; b < nneg; b++) {
                u32 pr = pn[a], nr = pn[SL.cand_cap - 1 - b];
                u32 need = (arena()[pr] >> H_SHIFT) + (arena()[nr] >> H_SHIFT);
                if (rtop + need > SL.rbuf_cap || nres + 2 > SL.roff_cap) { fail(ST_MEM_SIMP); return false; }
                u32 n_out = 0;
                if (merge(v, pr, nr, rbuf + rtop, n_out)) {
                    roff[nres] = rtop;
                    rtop += n_out;
                    nres++;
                    roff[nres] = rtop;
                    if ((u64)nres > max_resolvents || !(cl_lim == -1 || (i32)n_out <= cl_lim)) return true;
                }
            }
        S.V.vflags()[v] |= VF_ELIM;
        if (S.V.vflags()[v] & VF_DEC) { S.V.vflags()[v] &= (u8)~VF_DEC; sc().dec_vars--; } /* set_decision_var(v,false) */
        eliminated_vars++;
        if (npos > nneg) {
            for (u32 b = 0; b < nneg; b++) if (!mk_elim_clause(v, pn[SL.cand_cap - 1 - b])) return false;
            if (!mk_elim_unit(2 * v)) return false;
        } else {
            for (u32 a = 0; a < npos; a++) if (!mk_elim_clause(v, pn[a])) return false;
            if (!mk_elim_unit(2 * v + 1)) return false;
        }
        for (u32 i = 0; i < nc; i++) {
            u32 cr = cls[i];
            smudge_clause(cr);
            unwatch_lazy(cr);
            remove_clause(cr);
        }
        for (u32 r = 0; r < nres; r++) {
            /* the resolvent buffer is separate from the scratch seq_add_clause sorts in */
            if (!add_clause(rbuf + roff[r], roff[r + 1] - roff[r])) return false;
        }
        olen[v] = 0; ocap[v] = 0; opresent[v] = 0; /* clear_var */
        return true;
    }

    /* ------------------------------------------------------------ GC (clause.rs:192-223) scalar */
    DEV bool gc_reloc(u32 *dst, u32 &top, u64 &size, u32 cr, u32 &ncr) {
        u32 *ar = arena();
        u32 h = ar[cr];
        if (h & H_DELETED) return false;
        if (h & H_RELOCED) { ncr = ar[cr + 1]; return true; }
        u32 nh = h;
        if ((nh & H_EXTRA) && !(nh & H_LEARNT) && !sc().extra_clause_field) nh &= ~(u32)H_EXTRA;
        u32 words = 1 + (nh >> H_SHIFT) + ((nh & H_EXTRA) ? 1 : 0);
        ncr = top;
        dst[ncr] = nh;
        for (u32 k = 1; k < words; k++) dst[ncr + k] = ar[cr + k];
        top += WS::block_words(nh);
        size += WS::ref_size(nh);
        ar[cr + 1] = ncr;
        ar[cr] = h | H_RELOCED;
        return true;
    }
    DEV void seq_gc(bool with_simp) {
        const u32 hnew = sc().cur_arena ^ 1;
        u32 *dst = S.arena_half(hnew);
        u32 top = 0;
        u64 size = 0;
        for (u32 l = 0; l < 2 * S.nvars; l++) { /* watches.rs:154-169 */
            uint2 *ws = S.wpool + (S.V.wstart()[l] >> W_SHIFT);
            u32 n = S.V.wlen()[l], j = 0;
            for (u32 i = 0; i < n; i++) {
                u32 ncr;
                if (gc_reloc(dst, top, size, ws[i].x, ncr)) { ws[j] = make_uint2(ncr, ws[i].y); j++; }
            }
            S.V.wlen()[l] = (idx_t)j;
            S.V.wstart()[l] &= ~(u32)W_DIRTY;
        }
        for (u32 i = 0; i < S.trail_n; i++) { /* assignment.rs:236-243 */
            u32 v = S.V.trail()[i] >> 1;
            u32 r = S.V.reason()[v];
            if (r != CREF_UNDEF) { u32 ncr; S.V.reason()[v] = gc_reloc(dst, top, size, r, ncr) ? ncr : CREF_UNDEF; }
        }
        { /* clause_db.rs:256-280 */
            u32 *ls = S.learnts(); u32 n = sc().n_learnts, j = 0;
            for (u32 i = 0; i < n; i++) { u32 ncr; if (gc_reloc(dst, top, size, ls[i], ncr)) ls[j++] = ncr; }
            sc().n_learnts = j;
            u32 *cl = S.clauses(); n = sc().n_clauses; j = 0;
            for (u32 i = 0; i < n; i++) { u32 ncr; if (gc_reloc(dst, top, size, cl[i], ncr)) cl[j++] = ncr; }
            sc().n_clauses = j;
        }
        if (with_simp) {
            for (u32 v = 0; v < n_created; v++) { /* OccLists::gc elim_queue.rs:138-150 */
                if (!opresent[v]) continue;
                u32 *o = opool + ostart[v];
                u32 n = olen[v], j = 0;
                for (u32 i = 0; i < n; i++) { u32 ncr; if (gc_reloc(dst, top, size, o[i], ncr)) o[j++] = ncr; }
                olen[v] = j;
                odirty[v] = 0;
            }
            /* SubsumptionQueue::gc */
            u32 w = sq_head;
            for (u32 q = sq_head; q != sq_tail; q++) {
                u32 ncr;
                if (gc_reloc(dst, top, size, sq[q % SL.sq_cap], ncr)) { sq[w % SL.sq_cap] = ncr; w++; }
            }
            sq_tail = w;
        }
        sc().cur_arena = hnew;
        sc().arena_top = top;
        sc().lc_size = size;
        sc().lc_wasted = 0;
        S.arena = dst;
    }
    DEV void simp_try_gc() { /* simplify.rs:526-532 */
        if ((double)sc().lc_wasted > (double)sc().lc_size * S.st().simp_garbage_frac) seq_gc(true);
    }

    /* decision_heuristic.rs:146-156 (scalar) */
    DEV void seq_rebuild_order_heap() {
        u32 cnt = 0;
        for (u32 v = 0; v < S.nvars; v++) {
            bool in = (S.V.vflags()[v] & VF_DEC) && S.V.assign()[v] == LB_UNDEF;
            S.V.hidx()[v] = (idx_t)(in ? cnt : WS::VT_ABSENT);
            if (in) S.V.heap()[cnt++] = (idx_t)v;
        }
        sc().heap_n = cnt;
        for (i32 i = (i32)(cnt / 2) - 1; i >= 0; i--) S.heap_sift_down((u32)i);
    }
    /* clause_db.rs:217-254,283-303 (scalar) over one list */
    DEV u32 seq_remove_satisfied_list(u32 *list, u32 n) {
        u32 *ar = arena();
        u32 j = 0;
        for (u32 i = 0; i < n; i++) {
            u32 cr = list[i];
            u32 h = ar[cr];
            if (h & H_DELETED) continue;
            u32 len = h >> H_SHIFT;
            bool sat = false;
            for (u32 k = 0; k < len; k++) if (S.is_true(ar[cr + 1 + k])) { sat = true; break; }
            if (sat) { unwatch_lazy(cr); remove_clause(cr); continue; }
            u32 l = 2, r = len;
            while (l < r) {
                if (!S.is_false(ar[cr + 1 + l])) l++;
                else { r--; ar[cr + 1 + l] = ar[cr + 1 + r]; }
            }
            if (r != len) {
                if (h & H_EXTRA) ar[cr + 1 + r] = ar[cr + 1 + len];
                ar[cr] = (r << H_SHIFT) | (h & ((1u << H_SHIFT) - 1));
            }
            list[j++] = cr;
        }
        return j;
    }
    /* Searcher::try_simplify search.rs:522-568 (scalar) */
    DEV void seq_try_simplify() {
        if (S.nlevels != 1) return;
        if ((sc().simp_db_assigns_set && sc().simp_db_assigns == S.trail_n) || sc().propagations < sc().simp_db_props) return;
        sc().n_learnts = seq_remove_satisfied_list(S.learnts(), sc().n_learnts);
        if (sc().remove_satisfied) sc().n_clauses = seq_remove_satisfied_list(S.clauses(), sc().n_clauses);
        if ((double)sc().lc_wasted > (double)sc().lc_size * S.st().garbage_frac) seq_gc(false); /* search.rs:577-581 */
        seq_rebuild_order_heap();
        sc().simp_db_assigns_set = 1;
        sc().simp_db_assigns = S.trail_n;
        sc().simp_db_props = sc().propagations + sc().clauses_literals + sc().learnts_literals;
    }

    /* simplify.rs:303-336 (--asymm), including the reference's preserved skip after a strengthened long clause */
    DEV bool asymm_var(u32 v) {
        if (S.V.assign()[v] != LB_UNDEF) return true;
        occ_lookup(v);
        const u32 nc = olen[v];
        if (nc == 0) return true;
        if (nc > SL.cand_cap) { fail(ST_MEM_SIMP); return false; }
        for (u32 i = 0; i < nc; i++) cls[i] = opool[ostart[v] + i];
        bool bug = false;
        for (u32 i = 0; i < nc; i++) {
            if (bug) { bug = false; continue; }
            const u32 cr = cls[i];
            u32 l = LIT_NONE;
            if (seq_asymmetric_branching(v, cr, l)) {
                if ((arena()[cr] >> H_SHIFT) > 2) bug = true;
                if (!sq_push(cr)) return false;
                if (!strengthen_clause(cr, l)) return false;
            }
            if (failed()) return false;
        }
        return true;
    }

    /* simplify.rs:213-274; false = Err */
    DEV bool eliminate() {
        for (;;) {
            u32 ground = S.nlevels > 1 ? (u32)S.V.lim()[1] : S.trail_n;
            if (!(n_touched != 0 || ground - bwdsub_assigns > 0 || eheap_n > 0)) break;
            if (!enqueue_touched_clauses()) return false;
            if (!backward_subsumption_check()) return false;
            u32 var;
            while (eh_pop(var)) {
                if (is_elim(var) || S.V.assign()[var] != LB_UNDEF) continue;
                if (S.st().use_asymm) { /* simplify.rs:236-247 */
                    const u8 was = S.V.vflags()[var] & VF_FROZEN;
                    S.V.vflags()[var] |= VF_FROZEN;
                    if (!asymm_var(var)) return false;
                    if (!backward_subsumption_check()) return false;
                    S.V.vflags()[var] = (u8)((S.V.vflags()[var] & ~VF_FROZEN) | was);
                }
                if (S.st().use_elim && S.V.assign()[var] == LB_UNDEF && !is_frozen(var)) {
                    if (!eliminate_var(var)) return false;
                    if (!backward_subsumption_check()) return false;
                }
                simp_try_gc();
                if (failed()) return false;
            }
        }
        return true;
    }
    /* ============================================================ warp-cooperative drivers
     * The functions below are executed by ALL lanes with uniform control flow.  The data-parallel
     * parts (candidate subsumption tests, clause classification, resolvent merges) run one item per
     * lane on a snapshot; their effects are then applied by lane 0 strictly in the reference's order,
     * so queue order, heap tie-breaking and GC timing are unchanged.  Lane 0 owns the simplifier's
     * scalar state (the SeqSimp counters of the other lanes are not meaningful). */
    DEV void refresh_views() { /* lane 0 may have garbage-collected / compacted: re-derive the shared views */
        __syncwarp();
        S.arena = S.arena_half(sc().cur_arena);
        opool = base + SL.o_opool[sc().simp_cur_opool & 1];
    }
    DEV u32 merge_len(u32 v, u32 pr, u32 qr, bool &ok) const { /* merge() without the stores */
        const u32 *ar = S.arena;
        u32 np = ar[pr] >> H_SHIFT, nq = ar[qr] >> H_SHIFT;
        const u32 *longer = ar + pr + 1, *shorter = ar + qr + 1;
        u32 nl = np, ns = nq;
        if (np < nq) { longer = ar + qr + 1; nl = nq; shorter = ar + pr + 1; ns = np; }
        u32 n_out = 0;
        ok = true;
        for (u32 i = 0; i < ns; i++) {
            u32 q = shorter[i];
            if ((q >> 1) != v) {
                bool add = true;
                for (u32 j = 0; j < nl; j++) {
                    if ((longer[j] >> 1) == (q >> 1)) {
                        if (longer[j] == (q ^ 1)) { ok = false; return 0; }
                        add = false;
                        break;
                    }
                }
                if (add) n_out++;
            }
        }
        for (u32 j = 0; j < nl; j++) if ((longer[j] >> 1) != v) n_out++;
        return n_out;
    }

    /* simplify.rs:423-523; false = Err (uniform) */
    DEV bool u_backward_subsumption_check() {
        const u32 lane = S.lane;
        const i32 sub_lim = S.st().subsumption_lim;
        u32 *bl = S.V.bl();
        for (;;) {
            u32 have = 0, is_clause = 0, sub_cr = CREF_UNDEF, unit = 0, nc = 0, lstart = 0, err = 0;
            if (lane == 0) {
                for (;;) { /* SubsumptionQueue::pop subsumption_queue.rs:23-42 */
                    if (sq_len() > 0) {
                        u32 cr = sq[sq_head % SL.sq_cap];
                        sq_head++;
                        if (!(arena()[cr] & H_DELETED)) { is_clause = 1; sub_cr = cr; have = 1; break; }
                    } else {
                        u32 ground_end = S.nlevels > 1 ? (u32)S.V.lim()[1] : S.trail_n;
                        if (bwdsub_assigns < ground_end) { unit = S.V.trail()[bwdsub_assigns++]; have = 1; }
                        break;
                    }
                }
                if (have) {
                    u32 lookup_var;
                    if (!is_clause) lookup_var = unit >> 1;
                    else { /* simplify.rs:459-476: occurrence list with the fewest entries, dead ones included */
                        u32 n = arena()[sub_cr] >> H_SHIFT;
                        u32 best = arena()[sub_cr + 1] >> 1;
                        for (u32 i = 1; i < n; i++) {
                            u32 v = arena()[sub_cr + 1 + i] >> 1;
                            if (olen[v] < olen[best]) best = v;
                        }
                        lookup_var = best;
                    }
                    occ_lookup(lookup_var);
                    nc = olen[lookup_var];
                    lstart = ostart[lookup_var];
                    if (nc > SL.cand_cap) { fail(ST_MEM_SIMP); err = 1; }
                }
            }
            have = bcast(have); err = bcast(err);
            if (err) return false;
            if (!have) break;
            is_clause = bcast(is_clause); sub_cr = bcast(sub_cr); unit = bcast(unit); nc = bcast(nc); lstart = bcast(lstart);
            refresh_views();
            for (u32 i = lane; i < nc; i += 32) cand[i] = opool[lstart + i]; /* snapshot, simplify.rs:478 */
            __syncwarp();
            u32 stop = 0;
            for (u32 c0 = 0; c0 < nc && !stop; c0 += 32) {
                const u32 ci = c0 + lane;
                u32 code = SUB_DIFFERENT, l = 0;
                if (ci < nc) {
                    const u32 cj = cand[ci];
                    if (!(is_clause && cj == sub_cr)) {
                        const u32 hj = S.arena[cj];
                        if (!(hj & H_DELETED) && (sub_lim == -1 || (i32)(hj >> H_SHIFT) < sub_lim))
                            code = is_clause ? subsumes(sub_cr, cj, l) : unit_subsumes(unit, cj, l);
                    }
                }
                const u32 m = __ballot_sync(FULL_MASK, code != SUB_DIFFERENT);
                if (!m) continue;
                bl[lane] = code | (l << 2);
                __syncwarp();
                u32 e = 0;
                if (lane == 0) {
                    u32 mm = m;
                    while (mm) {
                        const u32 k = (u32)__ffs(mm) - 1;
                        mm &= mm - 1;
                        if (is_clause && (arena()[sub_cr] & H_DELETED)) { stop = 1; break; }
                        const u32 cjk = cand[c0 + k];
                        if (arena()[cjk] & H_DELETED) continue;
                        const u32 pk = bl[k];
                        if ((pk & 3) == SUB_EXACT) {
                            smudge_clause(cjk);
                            unwatch_lazy(cjk);
                            remove_clause(cjk);
                        } else {
                            if (!sq_push(cjk)) { e = 1; break; }
                            if (!strengthen_clause(cjk, (pk >> 2) ^ 1)) { e = 1; break; }
                        }
                    }
                }
                e = bcast(e); stop = bcast(stop);
                __syncwarp();
                if (e) return false;
            }
        }
        return true;
    }

    /* simplify.rs:338-420; false = Err (uniform) */
    DEV bool u_eliminate_var(u32 v) {
        const u32 lane = S.lane;
        const u32 lt = lanemask_lt(lane);
        u32 nc = 0, lstart = 0, err = 0;
        if (lane == 0) {
            occ_lookup(v);
            nc = olen[v];
            lstart = ostart[v];
            if (nc > SL.cand_cap) { fail(ST_MEM_SIMP); err = 1; }
        }
        err = bcast(err);
        if (err) return false;
        nc = bcast(nc); lstart = bcast(lstart);
        refresh_views();
        for (u32 i = lane; i < nc; i += 32) cls[i] = opool[lstart + i];
        __syncwarp();
        /* split by the sign of v; pos grows from the front of pn, neg from the back (clause order kept) */
        u32 npos = 0, nneg = 0;
        for (u32 c0 = 0; c0 < nc; c0 += 32) {
            const u32 i = c0 + lane;
            u32 cr = 0, sgn = 2;
            if (i < nc) {
                cr = cls[i];
                const u32 n = S.arena[cr] >> H_SHIFT;
                for (u32 k = 0; k < n; k++) {
                    const u32 l = S.arena[cr + 1 + k];
                    if ((l >> 1) == v) { sgn = l & 1; break; }
                }
            }
            const u32 pm = __ballot_sync(FULL_MASK, sgn == 0), nm = __ballot_sync(FULL_MASK, sgn == 1);
            if (sgn == 0) pn[npos + __popc(pm & lt)] = cr;
            if (sgn == 1) pn[SL.cand_cap - 1 - (nneg + __popc(nm & lt))] = cr;
            npos += __popc(pm); nneg += __popc(nm);
        }
        __syncwarp();
        const u64 max_resolvents = S.st().grow + npos + nneg;
        const i32 cl_lim = S.st().clause_lim;
        const u32 npairs = npos * nneg;
        /* pass 1: length / tautology of every resolvent, one pair per lane, limits checked in pair order */
        u32 nres = 0, rtop = 0;
        for (u32 t0 = 0; t0 < npairs; t0 += 32) {
            const u32 t = t0 + lane;
            bool ok = false;
            u32 len = 0;
            if (t < npairs) len = merge_len(v, pn[t / nneg], pn[SL.cand_cap - 1 - (t % nneg)], ok);
            const u32 okm = __ballot_sync(FULL_MASK, ok);
            const u32 rank = __popc(okm & lt);
            const bool viol = ok && ((u64)(nres + rank + 1) > max_resolvents || !(cl_lim == -1 || (i32)len <= cl_lim));
            if (__ballot_sync(FULL_MASK, viol)) return true; /* too many / too long resolvents: keep the variable */
            u32 tot;
            warp_excl_scan(ok ? len : 0, lane, tot);
            nres += __popc(okm);
            rtop += tot;
        }
        if (rtop > SL.rbuf_cap || nres + 2 > SL.roff_cap) { S.fail(ST_MEM_SIMP); return false; }
        /* pass 2: materialise the resolvents in pair order */
        u32 rr = 0, ro = 0;
        for (u32 t0 = 0; t0 < npairs; t0 += 32) {
            const u32 t = t0 + lane;
            bool ok = false;
            u32 len = 0;
            if (t < npairs) len = merge_len(v, pn[t / nneg], pn[SL.cand_cap - 1 - (t % nneg)], ok);
            const u32 okm = __ballot_sync(FULL_MASK, ok);
            u32 tot;
            const u32 off = ro + warp_excl_scan(ok ? len : 0, lane, tot);
            if (ok) {
                u32 n_out = 0;
                merge(v, pn[t / nneg], pn[SL.cand_cap - 1 - (t % nneg)], rbuf + off, n_out);
                const u32 idx = rr + __popc(okm & lt);
                roff[idx] = off;
                roff[idx + 1] = off + n_out; /* the next resolvent rewrites this slot with the same value */
            }
            rr += __popc(okm);
            ro += tot;
        }
        __syncwarp();
        u32 e = 0;
        if (lane == 0) {
            S.V.vflags()[v] |= VF_ELIM;
            if (S.V.vflags()[v] & VF_DEC) { S.V.vflags()[v] &= (u8)~VF_DEC; sc().dec_vars--; } /* set_decision_var(v,false) */
            eliminated_vars++;
            if (npos > nneg) {
                for (u32 b = 0; b < nneg && !e; b++) if (!mk_elim_clause(v, pn[SL.cand_cap - 1 - b])) e = 1;
                if (!e && !mk_elim_unit(2 * v)) e = 1;
            } else {
                for (u32 a = 0; a < npos && !e; a++) if (!mk_elim_clause(v, pn[a])) e = 1;
                if (!e && !mk_elim_unit(2 * v + 1)) e = 1;
            }
            for (u32 i = 0; i < nc && !e; i++) {
                u32 cr = cls[i];
                smudge_clause(cr);
                unwatch_lazy(cr);
                remove_clause(cr);
            }
            for (u32 r = 0; r < nres && !e; r++)
                if (!add_clause(rbuf + roff[r], roff[r + 1] - roff[r])) e = 1;
            if (!e) { olen[v] = 0; ocap[v] = 0; opresent[v] = 0; } /* clear_var */
        }
        e = bcast(e);
        __syncwarp();
        return e == 0;
    }

    /* simplify.rs:213-274; false = Err (uniform) */
    DEV bool u_eliminate() {
        const u32 lane = S.lane;
        for (;;) {
            u32 go = 0, e = 0;
            if (lane == 0) {
                u32 ground = S.nlevels > 1 ? (u32)S.V.lim()[1] : S.trail_n;
                go = (n_touched != 0 || ground - bwdsub_assigns > 0 || eheap_n > 0) ? 1 : 0;
                if (go && !enqueue_touched_clauses()) e = 1;
            }
            go = bcast(go); e = bcast(e);
            if (e) return false;
            if (!go) break;
            if (!u_backward_subsumption_check()) return false;
            for (;;) {
                u32 got = 0, var = 0, do_elim = 0;
                if (lane == 0) {
                    u32 x;
                    while (eh_pop(x)) {
                        if (is_elim(x) || S.V.assign()[x] != LB_UNDEF) continue;
                        got = 1; var = x;
                        do_elim = (S.st().use_elim && !is_frozen(x)) ? 1 : 0;
                        break;
                    }
                }
                got = bcast(got);
                if (!got) break;
                var = bcast(var);
                if (S.st().use_asymm) { /* simplify.rs:236-247: lane 0 branches, the subsumption check stays cooperative */
                    u32 bad = 0, was = 0;
                    if (lane == 0) {
                        was = S.V.vflags()[var] & VF_FROZEN;
                        S.V.vflags()[var] |= VF_FROZEN;
                        if (!asymm_var(var)) bad = 1;
                    }
                    if (bcast(bad)) return false;
                    refresh_views();
                    if (!u_backward_subsumption_check()) return false;
                    if (lane == 0) {
                        S.V.vflags()[var] = (u8)((S.V.vflags()[var] & ~VF_FROZEN) | was);
                        do_elim = (S.st().use_elim && S.V.assign()[var] == LB_UNDEF && !is_frozen(var)) ? 1 : 0;
                    }
                }
                do_elim = bcast(do_elim);
                if (do_elim) {
                    if (!u_eliminate_var(var)) return false;
                    if (!u_backward_subsumption_check()) return false;
                }
                u32 f = 0;
                if (lane == 0) { simp_try_gc(); f = failed() ? 1 : 0; }
                f = bcast(f);
                if (f) return false;
            }
        }
        return true;
    }

    DEV void off() { /* simplify.rs:536-544 */
        sc().remove_satisfied = 1;
        sc().extra_clause_field = 0;
        seq_rebuild_order_heap();
        seq_gc(false);
    }

    /* elim_clauses.rs:43-63 */
    DEV void extend_model(u8 *model) {
        for (u32 i = esizes_n; i-- > 0;) {
            u32 head = i > 0 ? esizes[i - 1] : 0, tail = esizes[i];
            bool sat = false;
            for (u32 k = head; k < tail; k++) {
                u32 l = elits[k];
                u8 m = model[l >> 1];
                if (m != 2 && (m != 0) != ((l & 1) != 0)) { sat = true; break; }
            }
            if (!sat) { u32 l = elits[head]; model[l >> 1] = (l & 1) ? 0 : 1; }
        }
    }
    DEV void flush_traffic() {
        if (COUNT) {
            u64 *t = sc().traffic;
            t[0] += t_vis; t[1] += t_kept; t[2] += t_deref; t[3] += t_tail; t[4] += t_moves; t[5] += t_impl;
            t_vis = t_kept = t_deref = t_tail = t_moves = t_impl = 0;
        }
    }
};

struct SimpPhaseOut { u32 trail_n, qhead, nlevels, n_ingest, n_pre, pre_ok, do_search; i32 verdict; };

/* Ingest + preprocessing of one instance (everything before the search).  Deliberately NOT inlined and
 * handed only plain values, from which it builds its own solver handle: the simplifier keeps a reference
 * to the handle, which forces that object into local memory; when it was the search's own object the
 * search loop lost the shared-memory address-space inference for the var state (measured: 2x slower
 * search, 56 % long-scoreboard stalls instead of 21 %, generic LD instead of LDS in the SASS).  All
 * lasting state lives in the slab and in shared memory. */
template <class VT, bool COUNT>
DEVNI SimpPhaseOut simp_phase(const KParams *Pp, VT V, u8 *slab, const b200sat_settings *stp, u32 lane_, u32 res_index,
                              u32 nvars, u32 cb, u32 ce) {
    WarpSolver<VT, COUNT> S(*Pp, lane_); /* state right after setup() */
    S.V = V; S.slab = slab; S.stp = stp; S.res_index = res_index; S.nvars = nvars;
    S.trail_n = 0; S.qhead = 0; S.nlevels = 1;
    S.arena = S.arena_half(0);
    S.wpool = S.wpool_half(0);
    const KParams &P = S.P;
    const u32 lane = S.lane;
    SimpPhaseOut o;
    o.n_ingest = 0; o.n_pre = 0; o.pre_ok = 1; o.do_search = 0; o.verdict = B200SAT_INTERRUPTED;
    /* every lane binds the region pointers; the simplifier's counters live on lane 0 */
    SeqSimp<WarpSolver<VT, COUNT>, COUNT> Z(S);
    Z.bind();
    u32 ok = 1, bad = 0, simp_alive = 1;
    if (lane == 0) {
        S.sc().extra_clause_field = 1;   /* Simplificator::on simplify.rs:546-549 */
        S.sc().remove_satisfied = 0;
        /* Var creation order matters here (the elimination heap is updated by every occurrence, elim_queue.rs:26-76).
         * A caller of the trait that creates every var first (sat.rs:31-32): var_flags without clause_nvars.  A caller
         * that interleaves new_var with add_clause: clause_nvars[c] vars exist when clause c arrives. */
        if (P.var_flags && !P.clause_nvars) while (Z.n_created < S.nvars) Z.init_var(Z.n_created);
        /* ingest: dimacs.rs:146-167 order -- vars are created lazily, interleaved with the clauses */
        for (u32 c = cb; c < ce && !Z.failed(); c++) {
            u32 b = P.clause_off[c], e = P.clause_off[c + 1];
            if (P.var_flags && P.clause_nvars) { const u32 want = P.clause_nvars[c] < S.nvars ? P.clause_nvars[c] : S.nvars; while (Z.n_created < want) Z.init_var(Z.n_created); }
            for (u32 k = b; k < e; k++) while ((P.lits[k] >> 1) + 1 > Z.n_created) Z.init_var(Z.n_created);
            if (!Z.add_clause(P.lits + b, e - b)) ok = 0; /* SimpSolver::add_clause: no `ok` check before (minisat.rs:146-158) */
        }
        if (P.var_flags) while (Z.n_created < S.nvars && !Z.failed()) Z.init_var(Z.n_created); /* vars created after the last clause */
        o.n_ingest = S.sc().num_clauses;
        o.n_pre = o.n_ingest;
        bad = Z.failed() ? 1 : 0;
        if (!bad && (P.flags & B200SAT_F_PREPROCESS) && ok) { /* CoreSolver::preprocess minisat.rs:50-55 */
            ok = Z.seq_propagate() == CREF_UNDEF ? 1 : 0;
            if (ok) Z.seq_try_simplify();
            bad = Z.failed() ? 1 : 0;
        }
    }
    ok = bcast(ok); bad = bcast(bad);
    if (!bad && (P.flags & B200SAT_F_PREPROCESS)) { /* SimpSolver::preprocess minisat.rs:161-191 */
        u32 res = 0;
        if (ok) {
            res = Z.u_eliminate() ? 1 : 0;
            if (lane == 0) Z.off();
            simp_alive = 0;
        }
        o.pre_ok = res;
        if (lane == 0) { o.n_pre = S.sc().num_clauses; bad = Z.failed() ? 1 : 0; }
        bad = bcast(bad);
    }
    if (!bad) {
        if (!o.pre_ok) o.verdict = B200SAT_UNSAT;
        else if (!simp_alive) o.do_search = 1; /* simp == None: core search, minisat.rs:241-260 (parked; the host launches it unless NO_SOLVE) */
        else if (P.flags & B200SAT_F_NO_SOLVE) o.verdict = B200SAT_INTERRUPTED; /* a live simplifier is not parked */
        else { /* Simplificator::solve_limited simplify.rs:165-211 */
            /* assumptions are frozen for the elimination in front of the search (simplify.rs:173-186) */
            if (lane == 0) for (u32 k = 0; k < P.n_assumps; k++) S.V.vflags()[P.assumps[k] >> 1] |= (u8)VF_FROZEN;
            __syncwarp();
            u32 pr = 0;
            if (lane == 0) { pr = Z.seq_propagate() == CREF_UNDEF ? 1 : 0; if (pr) Z.seq_try_simplify(); }
            pr = bcast(pr);
            if (pr && Z.u_eliminate()) o.do_search = 1; else o.verdict = B200SAT_UNSAT;
        }
    }
    if (lane == 0) {
        S.sc().eliminated_vars = Z.eliminated_vars;
        S.sc().elim_n = Z.esizes_n;
        Z.flush_traffic();
    }
    __syncwarp();
    /* hand the lane-0 scalars to the whole warp */
    o.trail_n = bcast(S.trail_n); o.qhead = bcast(S.qhead); o.nlevels = bcast(S.nlevels);
    o.n_ingest = bcast(o.n_ingest); o.n_pre = bcast(o.n_pre);
    return o;
}

/* SimpSolver ingest phase of one instance (../minisat.rs:123-279 + lib.rs:47-126 flags): var creation and
 * add_clause in DIMACS order, preprocessing (eliminate), and -- when solve_limited is entered with the simplifier
 * still alive -- the elimination pass in front of the search.  Ends like ingest_core: result record written, the
 * instance parked for the search kernel when a search has to follow.  Model extension happens after the search
 * (WarpSolver::finish_search, elim_clauses.rs:43-63) from the eliminated-clause stack left in the slab. */
template <class VT, bool COUNT> DEV bool ingest_simp(WarpSolver<VT, COUNT> &S, u32 inst) {
    const KParams &P = S.P;
    u32 cb, ce;
    bool setup_ok = S.setup(inst, cb, ce);
    u32 n_ingest = 0, n_pre = 0, pre_ok = 1;
    i32 verdict = B200SAT_INTERRUPTED;
    bool runnable = false;
    if (setup_ok && P.L.simp_cap == 0) { S.fail(ST_MEM_SIMP); setup_ok = false; }
    if (setup_ok) {
        const SimpPhaseOut o = simp_phase<VT, COUNT>(&S.P, S.V, S.slab, S.stp, S.lane, S.res_index, S.nvars, cb, ce);
        S.trail_n = o.trail_n; S.qhead = o.qhead; S.nlevels = o.nlevels;
        n_ingest = o.n_ingest; n_pre = o.n_pre; pre_ok = o.pre_ok; verdict = o.verdict;
        S.arena = S.arena_half(S.sc().cur_arena);
        S.wpool = S.wpool_half(S.sc().cur_wpool);
        __syncwarp();
        runnable = !S.failed() && o.do_search;
    }
    return S.park_after_ingest(verdict, n_ingest, n_pre, pre_ok, runnable);
}

} // namespace b200sat
#endif
