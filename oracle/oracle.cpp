/*
 * oracle.cpp -- CPU ORACLE for the batched-CDCL hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A single-threaded C++ restatement of mishun/minisat-rust (reference paths are
 * relative to /root/reference/src/sat).  It exists to (a) check the CUDA path's
 * verdicts, models and trace counters, (b) serve as the CPU baseline in
 * bench.py.  Nothing under minisat-rust_b200/ may call into this file.
 *
 * PARITY: "parity unpinned" for the five trace counters (no Rust toolchain and
 * no `minisat` binary exist in this image; the reference stores no golden
 * vectors -- tests/minisat.rs:70-170 compares against a live binary).  Pinned:
 * verdict by SATLIB file name, model validity, Lit/LBool tables.  See oracle.h.
 *
 * Floating point: build with -ffp-contract=off (no FMA contraction), see
 * SURVEY.md H2; the Makefile does.
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

namespace {

typedef uint32_t Lit;
typedef uint32_t Var;
typedef uint32_t CRef;
static const CRef CREF_UNDEF = 0xFFFFFFFFu;

/* formula.rs:12-18 */
enum : uint8_t { L_FALSE = 0, L_TRUE = 1, L_UNDEF = 2 };

static inline Var  lit_var(Lit l) { return l >> 1; }            /* formula.rs:73 */
static inline bool lit_sign(Lit l) { return (l & 1) != 0; }     /* formula.rs:68 */
static inline Lit  mk_lit(Var v, bool sign) { return (v << 1) | (sign ? 1u : 0u); } /* formula.rs:142 */
static inline uint32_t lit_abstraction(Lit l) { return 1u << ((l >> 1) & 31); }     /* formula.rs:107 */

static const uint32_t GROUND_LEVEL = 1;                          /* assignment.rs:10 */

struct Traffic {
    uint64_t visited = 0, kept = 0, deref = 0, tail = 0, moves = 0, implied = 0, an_words = 0, learnt_words = 0;
};

/* ---------------------------------------------------------------- clause.rs / allocator.rs */
static const uint32_t FLAG_BITS = 4, FLAG_DELETED = 1, FLAG_TOUCHED = 2, FLAG_RELOCED = 8; /* clause.rs:6-9 */
enum : uint32_t { TAG_PLAIN = 0 /* Clause{abstraction:None} */, TAG_ABS = 1 /* Clause{Some} */, TAG_LEARNT = 2 };

struct ClauseAllocator {
    /* clause at word cr: [mark][tag][value][lits...]  (clause.rs:15-19, clause_header.rs:5-8) */
    std::vector<uint32_t> mem;
    size_t lc_size = 0, lc_wasted = 0;   /* LegacyCounter clause.rs:226-251 */
    bool extra_clause_field = false;

    uint32_t len(CRef cr) const { return mem[cr] >> FLAG_BITS; }
    bool deleted(CRef cr) const { return (mem[cr] & FLAG_DELETED) != 0; }
    bool touched(CRef cr) const { return (mem[cr] & FLAG_TOUCHED) != 0; }
    void set_touched(CRef cr, bool b) { if (b) mem[cr] |= FLAG_TOUCHED; else mem[cr] &= ~FLAG_TOUCHED; }
    uint32_t tag(CRef cr) const { return mem[cr + 1]; }
    bool learnt(CRef cr) const { return mem[cr + 1] == TAG_LEARNT; }
    uint32_t &value(CRef cr) { return mem[cr + 2]; }
    float activity(CRef cr) const { float f; memcpy(&f, &mem[cr + 2], 4); return f; }
    void set_activity(CRef cr, float f) { memcpy(&mem[cr + 2], &f, 4); }
    Lit *lits(CRef cr) { return &mem[cr + 3]; }
    const Lit *lits(CRef cr) const { return &mem[cr + 3]; }
    void shrink_by(CRef cr, uint32_t k) { mem[cr] -= (k << FLAG_BITS); }   /* clause.rs:28-31 */

    size_t clause_size(CRef cr) const {                                    /* clause.rs:232-238 */
        return tag(cr) == TAG_PLAIN ? 4 * (1 + (size_t)len(cr)) : 4 * (2 + (size_t)len(cr));
    }
    CRef alloc(const Lit *ls, uint32_t n, uint32_t tg, uint32_t val) {     /* clause.rs:120-134 */
        CRef cr = (CRef)mem.size();
        mem.push_back(n << FLAG_BITS);
        mem.push_back(tg);
        mem.push_back(val);
        mem.insert(mem.end(), ls, ls + n);
        lc_size += clause_size(cr);
        return cr;
    }
    void free_(CRef cr) {                                                  /* clause.rs:157-162 */
        mem[cr] |= FLAG_DELETED;
        lc_wasted += clause_size(cr);
    }
    bool check_garbage(double gf) const {                                  /* clause.rs:165-167 */
        return (double)lc_wasted > (double)lc_size * gf;
    }
};

struct ClauseGC {                                                          /* clause.rs:192-223 */
    ClauseAllocator &src;
    ClauseAllocator dst;
    explicit ClauseGC(ClauseAllocator &s) : src(s) {
        dst.extra_clause_field = s.extra_clause_field;
        dst.mem.reserve(s.mem.size());
    }
    bool relocate(CRef cr, CRef &out) {
        uint32_t mark = src.mem[cr];
        if (mark & FLAG_DELETED) return false;
        if ((mark & FLAG_RELOCED) == 0) {
            /* ClauseAllocator::reloc clause.rs:136-155 */
            uint32_t n = mark >> FLAG_BITS;
            CRef nr = (CRef)dst.mem.size();
            dst.mem.push_back(mark);
            uint32_t tg = src.mem[cr + 1], val = src.mem[cr + 2];
            if (!dst.extra_clause_field && tg != TAG_LEARNT) { tg = TAG_PLAIN; val = 0; }
            dst.mem.push_back(tg);
            dst.mem.push_back(val);
            dst.mem.insert(dst.mem.end(), src.mem.begin() + cr + 3, src.mem.begin() + cr + 3 + n);
            dst.lc_size += dst.clause_size(nr);
            src.mem[cr + 3] = nr;           /* forwarding pointer in prefix[0] */
            src.mem[cr] |= FLAG_RELOCED;
            out = nr;
        } else {
            out = src.mem[cr + 3];
        }
        return true;
    }
    void finish() { std::swap(src.mem, dst.mem); src.lc_size = dst.lc_size; src.lc_wasted = dst.lc_wasted; }
};

/* ---------------------------------------------------------------- assignment.rs */
struct Assignment {
    std::vector<uint8_t> assign;
    std::vector<uint32_t> level;
    std::vector<CRef> reason;
    std::vector<Lit> trail;
    std::vector<size_t> lim{0};
    size_t qhead = 0;

    Var new_var() {                                                        /* assignment.rs:50-66 */
        Var v = (Var)assign.size();
        assign.push_back(L_UNDEF);
        level.push_back(GROUND_LEVEL);
        reason.push_back(CREF_UNDEF);
        return v;
    }
    uint32_t current_level() const { return (uint32_t)lim.size(); }        /* :74-76 */
    bool is_ground_level() const { return lim.size() == 1; }
    size_t number_of_vars() const { return assign.size(); }
    size_t number_of_assigns() const { return trail.size(); }
    size_t number_of_ground_assigns() const { return lim.size() > 1 ? lim[1] : trail.size(); } /* :90-95 */
    void new_decision_level() { lim.push_back(trail.size()); }             /* :99-101 */
    void assign_lit(Lit l, CRef r) {                                       /* :104-113 */
        Var v = lit_var(l);
        if (assign[v] != L_UNDEF) { fprintf(stderr, "oracle: assign_lit on assigned var\n"); abort(); }
        assign[v] = (uint8_t)((~l) & 1);
        reason[v] = r;
        level[v] = (uint32_t)lim.size();
        trail.push_back(l);
    }
    void backtrack_to(uint32_t target_level) {                             /* :116-130 */
        if (target_level < lim.size()) {
            size_t target = lim[target_level];
            for (size_t i = target; i < trail.size(); i++) {
                Var v = lit_var(trail[i]);
                assign[v] = L_UNDEF;
                reason[v] = CREF_UNDEF;
            }
            trail.resize(target);
            lim.resize(target_level);
        }
        qhead = std::min(qhead, trail.size());
    }
    bool is_undef(Var v) const { return assign[v] == L_UNDEF; }
    bool is_true(Lit l) const { return (l & 1) == (uint32_t)(assign[lit_var(l)] ^ 1); }   /* formula.rs:78-80 */
    bool is_false(Lit l) const { return (l & 1) == (uint32_t)assign[lit_var(l)]; }        /* formula.rs:83-85 */
    bool is_reason_for(CRef cr, Lit l) const { return is_true(l) && reason[lit_var(l)] == cr; } /* :225-233 */
    size_t level_begin(uint32_t lv) const { return lim[lv - 1]; }
    size_t level_end(uint32_t lv) const { return lv < lim.size() ? lim[lv] : trail.size(); }
};

/* ---------------------------------------------------------------- index_map.rs:183-309 IdxHeap */
template <class Before> struct IdxHeap {
    std::vector<Var> heap;
    std::vector<int32_t> index;   /* -1 = absent (VecMap::contains_key) */
    Before before;

    size_t len() const { return heap.size(); }
    bool empty() const { return heap.empty(); }
    bool contains(Var k) const { return k < index.size() && index[k] >= 0; }
    void set_index(Var k, size_t i) { if (k >= index.size()) index.resize(k + 1, -1); index[k] = (int32_t)i; }
    void clear() { heap.clear(); std::fill(index.begin(), index.end(), -1); }
    bool insert(Var k) {                                                   /* :218-227 */
        if (contains(k)) return false;
        size_t place = heap.size();
        heap.push_back(k);
        sift_up(place);
        return true;
    }
    bool pop(Var &out) {                                                   /* :230-241 */
        if (heap.empty()) return false;
        out = heap[0];
        heap[0] = heap.back();
        heap.pop_back();
        index[out] = -1;
        if (!heap.empty()) sift_down(0);
        return true;
    }
    bool update(Var k) {                                                   /* :244-255 */
        if (!contains(k)) return false;
        size_t place = (size_t)index[k];
        sift_down(place);
        sift_up(place);
        return true;
    }
    void heapify_from(const std::vector<Var> &from) {                      /* :257-264 */
        std::fill(index.begin(), index.end(), -1);
        heap = from;
        for (size_t i = heap.size(); i-- > 0;) sift_down(i);
    }
    void sift_up(size_t i) {                                               /* :267-280 */
        while (i > 0) {
            size_t p = (i - 1) >> 1;
            if (before(heap[i], heap[p])) {
                set_index(heap[p], i);
                std::swap(heap[i], heap[p]);
                i = p;
            } else break;
        }
        set_index(heap[i], i);
    }
    void sift_down(size_t i) {                                             /* :283-308 */
        for (;;) {
            size_t l = 2 * i + 1;
            if (l >= heap.size()) break;
            size_t r = l + 1;
            size_t c = (r < heap.size() && before(heap[r], heap[l])) ? r : l;
            if (before(heap[c], heap[i])) {
                set_index(heap[c], i);
                std::swap(heap[c], heap[i]);
                i = c;
            } else break;
        }
        set_index(heap[i], i);
    }
};

/* ---------------------------------------------------------------- random.rs */
struct Random {
    double seed;
    double drand() {                                                       /* random.rs:12-17 */
        seed *= 1389796.0;
        int32_t q = (int32_t)(seed / 2147483647.0);
        seed -= (double)q * 2147483647.0;
        return seed / 2147483647.0;
    }
    size_t irand(size_t size) { return (size_t)(drand() * (double)size); } /* :20-22 */
    bool chance(double p) { return drand() < p; }                          /* :24-26 */
};

/* luby.rs:2-20 */
static double luby(double y, uint32_t x) {
    uint32_t size = 1;
    int seq = 0;
    while (size < x + 1) { seq += 1; size = 2 * size + 1; }
    while (size - 1 != x) { size = (size - 1) >> 1; seq -= 1; x = x % size; }
    /* f64::powi: repeated multiplication; exact for the default base 2.0 */
    double r = 1.0, b = y;
    int e = seq;
    while (e > 0) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}
static double powi(double b, int e) { double r = 1.0; while (e > 0) { if (e & 1) r *= b; b *= b; e >>= 1; } return r; }

/* ---------------------------------------------------------------- decision_heuristic.rs */
struct ActBefore {
    const std::vector<double> *act;
    bool operator()(Var a, Var b) const { return (*act)[a] > (*act)[b]; }
};

struct DecisionHeuristic {
    double var_decay, random_var_freq;
    int phase_saving; bool rnd_pol, rnd_init_act;
    double var_inc = 1.0;
    Random rand;
    std::vector<uint8_t> polarity, decision;
    std::vector<int8_t> user_pol;
    std::vector<double> activity;
    IdxHeap<ActBefore> queue;
    size_t dec_vars = 0;
    uint64_t rnd_decisions = 0;

    void init(const oracle_settings &s) {
        var_decay = s.var_decay; random_var_freq = s.random_var_freq; phase_saving = s.phase_saving;
        rnd_pol = s.rnd_pol; rnd_init_act = s.rnd_init_act; rand.seed = s.random_seed;
        queue.before.act = &activity;
    }
    void init_var(Var v, int upol, bool dvar) {                            /* :70-88 */
        activity.push_back(rnd_init_act ? rand.drand() * 0.00001 : 0.0);
        polarity.push_back(1);
        user_pol.push_back((int8_t)upol);
        decision.push_back(0);
        set_decision_var(v, dvar);
    }
    void set_decision_var(Var v, bool b) {                                 /* :90-102 */
        if (b != (decision[v] != 0)) {
            if (b) { dec_vars += 1; queue.insert(v); } else dec_vars -= 1;
            decision[v] = b;
        }
    }
    void save_phase(Lit l, bool top_level) {                               /* :105-116 */
        if (phase_saving == 2 || (phase_saving == 1 && top_level)) polarity[lit_var(l)] = lit_sign(l);
    }
    void try_return_var(Var v) { if (decision[v]) queue.insert(v); }       /* :118-124 */
    void bump_activity(Var v) {                                            /* :126-140 */
        double nw = activity[v] + var_inc;
        if (nw > 1e100) {
            var_inc *= 1e-100;
            for (double &a : activity) a *= 1e-100;
            activity[v] = nw * 1e-100;
        } else activity[v] = nw;
        queue.update(v);
    }
    void decay_activity() { var_inc *= 1.0 / var_decay; }                  /* :142-144 */
    void rebuild_order_heap(const Assignment &as) {                        /* :146-156 */
        std::vector<Var> tmp;
        for (Var v = 0; v < decision.size(); v++)
            if (decision[v] && as.is_undef(v)) tmp.push_back(v);
        queue.heapify_from(tmp);
    }
    bool pick_branch_var(const Assignment &as, Var &out) {                 /* :158-179 */
        if (rand.chance(random_var_freq) && !queue.empty()) {
            Var v = queue.heap[rand.irand(queue.len())];
            if (as.is_undef(v) && decision[v]) { rnd_decisions += 1; out = v; return true; }
        }
        Var v;
        while (queue.pop(v))
            if (as.is_undef(v) && decision[v]) { out = v; return true; }
        return false;
    }
    bool pick_branch_lit(const Assignment &as, Lit &out) {                 /* :181-192 */
        Var v;
        if (!pick_branch_var(as, v)) return false;
        bool sign = user_pol[v] >= 0 ? (user_pol[v] != 0) : (rnd_pol ? rand.chance(0.5) : (polarity[v] != 0));
        out = mk_lit(v, sign);
        return true;
    }
};

/* ---------------------------------------------------------------- watches.rs */
struct Watcher { CRef cref; Lit blocker; };
struct WatchesLine { std::vector<Watcher> ws; bool dirty = false; };

struct Watches {
    std::vector<WatchesLine> lines;
    uint64_t propagations = 0;
    void init_var() { lines.emplace_back(); lines.emplace_back(); }        /* :32-35 */
    void watch_clause(ClauseAllocator &ca, CRef cr) {                      /* :39-44 */
        const Lit *c = ca.lits(cr);
        for (int i = 0; i < 2; i++) lines[c[i] ^ 1].ws.push_back(Watcher{cr, c[i ^ 1]});
    }
    void unwatch_clause_strict(ClauseAllocator &ca, CRef cr) {             /* :46-50 */
        const Lit *c = ca.lits(cr);
        for (int i = 0; i < 2; i++) {
            auto &ws = lines[c[i] ^ 1].ws;
            size_t j = 0;
            for (size_t k = 0; k < ws.size(); k++) if (ws[k].cref != cr) ws[j++] = ws[k];
            ws.resize(j);
        }
    }
    void unwatch_clause_lazy(ClauseAllocator &ca, CRef cr) {               /* :52-56 */
        const Lit *c = ca.lits(cr);
        lines[c[0] ^ 1].dirty = true;
        lines[c[1] ^ 1].dirty = true;
    }
    /* watches.rs:64-152.  Traffic counters follow SURVEY.md 8d. */
    CRef propagate(ClauseAllocator &ca, Assignment &as, Traffic &T) {
        while (as.qhead < as.trail.size()) {
            Lit p = as.trail[as.qhead++];
            propagations += 1;
            Lit not_p = p ^ 1;
            WatchesLine &line = lines[p];
            std::vector<Watcher> &ws = line.ws;
            size_t n = ws.size(), head = 0, tail = 0;
            while (head < n) {
                Watcher pwi = ws[head++];
                if (ca.deleted(pwi.cref)) continue;                         /* :81-84 */
                T.visited++;
                if (as.is_true(pwi.blocker)) { ws[tail++] = pwi; T.kept++; continue; } /* :86-90 */
                T.deref++;
                Lit *c = ca.lits(pwi.cref);
                if (c[0] == not_p) std::swap(c[0], c[1]);                   /* :92-94 */
                Watcher cw{pwi.cref, c[0]};
                if (cw.blocker != pwi.blocker && as.is_true(cw.blocker)) {  /* :98-103 */
                    ws[tail++] = cw; T.kept++; continue;
                }
                uint32_t len = ca.len(pwi.cref);
                bool moved = false;
                for (uint32_t k = 2; k < len; k++) {                        /* :106-117 */
                    T.tail++;
                    if (!as.is_false(c[k])) {
                        lines[c[k] ^ 1].ws.push_back(cw);
                        std::swap(c[1], c[k]);
                        T.moves++;
                        moved = true;
                        break;
                    }
                }
                if (moved) continue;
                ws[tail++] = cw; T.kept++;                                  /* :120-121 */
                if (as.is_false(cw.blocker)) {                              /* :122-140 */
                    if (line.dirty) {
                        while (head < n) { if (!ca.deleted(ws[head].cref)) ws[tail++] = ws[head]; head++; }
                        line.dirty = false;
                    } else {
                        while (head < n) ws[tail++] = ws[head++];
                    }
                    ws.resize(tail);
                    return cw.cref;
                } else {
                    as.assign_lit(cw.blocker, cw.cref);                     /* :142 */
                    T.implied++;
                }
            }
            line.dirty = false;                                             /* :146-147 */
            ws.resize(tail);
        }
        return CREF_UNDEF;
    }
    void gc(ClauseGC &gc) {                                                /* :154-169 */
        for (auto &line : lines) {
            line.dirty = false;
            size_t j = 0;
            for (size_t i = 0; i < line.ws.size(); i++) {
                CRef nr;
                if (gc.relocate(line.ws[i].cref, nr)) { line.ws[j] = line.ws[i]; line.ws[j].cref = nr; j++; }
            }
            line.ws.resize(j);
        }
    }
};

/* ---------------------------------------------------------------- clause_db.rs */
struct DBStats { size_t num_clauses = 0, num_learnts = 0; uint64_t clauses_literals = 0, learnts_literals = 0; };

struct ClauseDB {
    bool remove_satisfied_flag = true;
    double clause_decay = 0.999;
    double cla_inc = 1.0;
    std::vector<CRef> clauses, learnts;
    DBStats stats;

    void st_add(ClauseAllocator &ca, CRef cr) {                            /* :29-41 */
        if (ca.learnt(cr)) { stats.num_learnts += 1; stats.learnts_literals += ca.len(cr); }
        else { stats.num_clauses += 1; stats.clauses_literals += ca.len(cr); }
    }
    void st_del(ClauseAllocator &ca, CRef cr) {                            /* :43-55 */
        if (ca.learnt(cr)) { stats.num_learnts -= 1; stats.learnts_literals -= ca.len(cr); }
        else { stats.num_clauses -= 1; stats.clauses_literals -= ca.len(cr); }
    }
    CRef add_clause(ClauseAllocator &ca, const Lit *ls, uint32_t n) {      /* :78-91 */
        uint32_t tg = TAG_PLAIN, val = 0;
        if (ca.extra_clause_field) { tg = TAG_ABS; for (uint32_t i = 0; i < n; i++) val |= lit_abstraction(ls[i]); }
        CRef cr = ca.alloc(ls, n, tg, val);
        st_add(ca, cr);
        clauses.push_back(cr);
        return cr;
    }
    CRef learn_clause(ClauseAllocator &ca, const Lit *ls, uint32_t n) {    /* :93-100 */
        CRef cr = ca.alloc(ls, n, TAG_LEARNT, 0);
        st_add(ca, cr);
        learnts.push_back(cr);
        bump_activity(ca, cr);
        return cr;
    }
    void remove_clause(ClauseAllocator &ca, CRef cr) { st_del(ca, cr); ca.free_(cr); } /* :102-105 */
    void bump_activity(ClauseAllocator &ca, CRef cr) {                     /* :119-143 */
        if (!ca.learnt(cr)) return;
        double nw = (double)ca.activity(cr) + cla_inc;
        ca.set_activity(cr, (float)nw);
        if (nw > 1e20) {
            cla_inc *= 1e-20;
            for (CRef c : learnts) {
                double scaled = (double)ca.activity(c) * 1e-20;
                ca.set_activity(c, (float)scaled);
            }
        }
    }
    void decay_activity() { cla_inc *= 1.0 / clause_decay; }               /* :145-147 */

    template <class Notify>
    void reduce(ClauseAllocator &ca, const Assignment &as, Notify notify) { /* :156-215 */
        /* slice::sort_by is stable: binary == binary, binary > rest, rest by f32 activity */
        std::stable_sort(learnts.begin(), learnts.end(), [&](CRef rx, CRef ry) {
            bool xb = ca.len(rx) == 2, yb = ca.len(ry) == 2;
            if (xb && yb) return false;
            if (xb) return false;
            if (yb) return true;
            return ca.activity(rx) < ca.activity(ry);
        });
        size_t index_lim = learnts.size() / 2;
        double extra_lim = cla_inc / (double)learnts.size();
        size_t i = 0, j = 0;
        for (size_t k = 0; k < learnts.size(); k++) {
            CRef cr = learnts[k];
            if (ca.deleted(cr)) { i += 1; continue; }
            bool remove = ca.len(cr) > 2 && !as.is_reason_for(cr, ca.lits(cr)[0]) &&
                          (i < index_lim || (double)ca.activity(cr) < extra_lim);
            if (remove) { notify(cr); st_del(ca, cr); }
            i += 1;
            if (remove) ca.free_(cr); else learnts[j++] = cr;
        }
        learnts.resize(j);
    }
    template <class Notify>
    bool retain_clause(ClauseAllocator &ca, const Assignment &as, Notify &notify, CRef cr) { /* :217-236 */
        if (ca.deleted(cr)) return false;
        uint32_t n = ca.len(cr);
        Lit *c = ca.lits(cr);
        bool sat = false;
        for (uint32_t k = 0; k < n; k++) if (as.is_true(c[k])) { sat = true; break; }
        if (sat) { notify(cr); st_del(ca, cr); ca.free_(cr); return false; }
        /* retain_clause :283-303 -- trim false literals at positions >= 2, stats untouched */
        uint32_t l = 2, r = n;
        while (l < r) {
            if (!as.is_false(c[l])) l++;
            else { ca.shrink_by(cr, 1); r--; c[l] = c[r]; }
        }
        return true;
    }
    template <class Notify>
    void remove_satisfied(ClauseAllocator &ca, const Assignment &as, Notify notify) { /* :238-254 */
        size_t j = 0;
        for (size_t k = 0; k < learnts.size(); k++) if (retain_clause(ca, as, notify, learnts[k])) learnts[j++] = learnts[k];
        learnts.resize(j);
        if (remove_satisfied_flag) {
            j = 0;
            for (size_t k = 0; k < clauses.size(); k++) if (retain_clause(ca, as, notify, clauses[k])) clauses[j++] = clauses[k];
            clauses.resize(j);
        }
    }
    void gc(ClauseGC &gc) {                                                /* :256-280 */
        size_t j = 0;
        for (size_t i = 0; i < learnts.size(); i++) { CRef nr; if (gc.relocate(learnts[i], nr)) learnts[j++] = nr; }
        learnts.resize(j);
        j = 0;
        for (size_t i = 0; i < clauses.size(); i++) { CRef nr; if (gc.relocate(clauses[i], nr)) clauses[j++] = nr; }
        clauses.resize(j);
    }
};

/* ---------------------------------------------------------------- conflict.rs */
enum : uint8_t { SEEN_UNDEF = 0, SEEN_SOURCE = 1, SEEN_REMOVABLE = 2, SEEN_FAILED = 3 };
enum ConflictKind { CONFL_GROUND, CONFL_UNIT, CONFL_LEARNED };

struct AnalyzeContext {
    int ccmin_mode = 2;
    std::vector<uint8_t> seen;
    std::vector<Lit> analyze_toclear;
    uint64_t max_literals = 0, tot_literals = 0;
    void init_var() { seen.push_back(SEEN_UNDEF); }

    bool lit_redundant_basic(const ClauseAllocator &ca, const Assignment &as, Lit literal, Traffic &T) { /* :178-192 */
        CRef cr = as.reason[lit_var(literal)];
        if (cr == CREF_UNDEF) return false;
        uint32_t n = ca.len(cr);
        const Lit *c = ca.lits(cr);
        T.an_words += 1 + n;
        for (uint32_t k = 1; k < n; k++)
            if (seen[lit_var(c[k])] == SEEN_UNDEF && as.level[lit_var(c[k])] > GROUND_LEVEL) return false;
        return true;
    }
    bool lit_redundant(const ClauseAllocator &ca, const Assignment &as, Lit literal, Traffic &T) {       /* :195-249 */
        struct Frame { Lit p; CRef cr; uint32_t pos; };
        CRef cr0 = as.reason[lit_var(literal)];
        if (cr0 == CREF_UNDEF) return false;
        std::vector<Frame> stack;
        stack.push_back(Frame{literal, cr0, 1});
        T.an_words += 1 + ca.len(cr0);
        while (!stack.empty()) {
            Frame f = stack.back();
            stack.pop_back();
            if (f.pos < ca.len(f.cr)) {
                Lit l = ca.lits(f.cr)[f.pos];
                stack.push_back(Frame{f.p, f.cr, f.pos + 1});
                Var lv = lit_var(l);
                uint8_t s = seen[lv];
                if (as.level[lv] == GROUND_LEVEL || s == SEEN_SOURCE || s == SEEN_REMOVABLE) continue;
                CRef r = as.reason[lv];
                if (r != CREF_UNDEF && s == SEEN_UNDEF) {
                    stack.push_back(Frame{l, r, 1});
                    T.an_words += 1 + ca.len(r);
                } else {
                    for (const Frame &g : stack) {
                        if (seen[lit_var(g.p)] == SEEN_UNDEF) {
                            seen[lit_var(g.p)] = SEEN_FAILED;
                            analyze_toclear.push_back(g.p);
                        }
                    }
                    return false;
                }
            } else {
                if (seen[lit_var(f.p)] == SEEN_UNDEF) {
                    seen[lit_var(f.p)] = SEEN_REMOVABLE;
                    analyze_toclear.push_back(f.p);
                }
            }
        }
        return true;
    }
};

/* ---------------------------------------------------------------- search.rs */
struct LearningGuard {                                                     /* :72-111 */
    int min_learnts_lim; double size_factor, size_inc; int size_adjust_start_confl; double size_adjust_inc;
    double max_learnts = 0.0, size_adjust_confl = 0.0;
    int size_adjust_cnt = 0;
    void reset(size_t clauses) {
        max_learnts = std::max((double)clauses * size_factor, (double)min_learnts_lim);
        size_adjust_confl = (double)size_adjust_start_confl;
        size_adjust_cnt = size_adjust_start_confl;
    }
    bool bump() {
        size_adjust_cnt -= 1;
        if (size_adjust_cnt == 0) {
            size_adjust_confl *= size_adjust_inc;
            size_adjust_cnt = (int)size_adjust_confl;
            max_learnts *= size_inc;
            return true;
        }
        return false;
    }
};

struct TraceSink { uint64_t *buf = nullptr; uint32_t cap = 0, n = 0; };
struct ProgressRow { double v[8]; };
static std::vector<ProgressRow> *g_progress_rows = nullptr;   /* single-threaded use only (oracle_solve_csr) */
static thread_local TraceSink g_trace;

struct Searcher {
    oracle_settings st;
    ClauseAllocator ca;
    Assignment assigns;
    Watches watches;
    /* SearchCtx :192-198 */
    uint64_t solves = 0, starts = 0, decisions = 0, conflicts = 0;
    ClauseDB db;
    DecisionHeuristic heur;
    AnalyzeContext an;
    bool simp_db_assigns_set = false; size_t simp_db_assigns = 0; uint64_t simp_db_props = 0; /* SimplifyGuard :114-135 */
    Traffic T;
    uint64_t trace_hash = 1469598103934665603ull;
    std::vector<Lit> out_learnt;

    explicit Searcher(const oracle_settings &s) : st(s) {
        db.remove_satisfied_flag = s.remove_satisfied;
        db.clause_decay = s.clause_decay;
        heur.init(s);
        an.ccmin_mode = s.ccmin_mode;
    }
    size_t number_of_vars() const { return assigns.number_of_vars(); }
    size_t number_of_clauses() const { return db.stats.num_clauses; }

    Var new_var(int upol, bool dvar) {                                     /* :336-340 */
        Var v = assigns.new_var();
        watches.init_var();
        heur.init_var(v, upol, dvar);
        an.init_var();
        return v;
    }
    CRef propagate() { return watches.propagate(ca, assigns, T); }
    void attach(CRef cr) { watches.watch_clause(ca, cr); }

    /* search/util.rs:16-42 */
    bool is_implied(const Lit *c, uint32_t n) {
        assigns.new_decision_level();
        for (uint32_t i = 0; i < n; i++) {
            Lit l = c[i];
            if (assigns.is_true(l)) {
                for (size_t k = assigns.level_begin(2); k < assigns.trail.size(); k++) heur.save_phase(assigns.trail[k], true);
                assigns.backtrack_to(GROUND_LEVEL);
                return true;
            } else if (assigns.is_undef(lit_var(l))) {
                assigns.assign_lit(l ^ 1, CREF_UNDEF);
            }
        }
        bool result = propagate() != CREF_UNDEF;
        for (size_t k = assigns.level_begin(2); k < assigns.trail.size(); k++) heur.save_phase(assigns.trail[k], true);
        assigns.backtrack_to(GROUND_LEVEL);
        return result;
    }

    enum AddRes { ADD_UNSAT, ADD_CONSUMED, ADD_ADDED };
    AddRes add_clause(const Lit *clause, uint32_t n, CRef &out_cr) {       /* :342-386 */
        if (st.use_rcheck && is_implied(clause, n)) return ADD_CONSUMED;
        std::vector<Lit> ps(clause, clause + n);
        std::sort(ps.begin(), ps.end());
        ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
        ps.erase(std::remove_if(ps.begin(), ps.end(), [&](Lit l) { return assigns.is_false(l); }), ps.end());
        {
            bool has_prev = false; Lit prev = 0;
            for (Lit l : ps) {
                if (assigns.is_true(l) || (has_prev && prev == (l ^ 1))) return ADD_CONSUMED;
                prev = l; has_prev = true;
            }
        }
        if (ps.empty()) return ADD_UNSAT;
        if (ps.size() == 1) {
            assigns.assign_lit(ps[0], CREF_UNDEF);
            return propagate() == CREF_UNDEF ? ADD_CONSUMED : ADD_UNSAT;
        }
        out_cr = db.add_clause(ca, ps.data(), (uint32_t)ps.size());
        attach(out_cr);
        return ADD_ADDED;
    }

    bool preprocess() {                                                    /* :388-395 */
        if (propagate() == CREF_UNDEF) { try_simplify(); return true; }
        return false;
    }

    /* ---- SearchCtx */
    void ctx_cancel_until(uint32_t level) {                                /* :253-261 */
        uint32_t top = assigns.current_level();
        for (uint32_t lv = top; lv > level; lv--) {
            size_t b = assigns.level_begin(lv), e = assigns.level_end(lv);
            for (size_t k = e; k-- > b;) {
                Lit l = assigns.trail[k];
                heur.save_phase(l, lv == top);
                heur.try_return_var(lit_var(l));
            }
        }
    }
    void cancel_until(uint32_t level) { ctx_cancel_until(level); assigns.backtrack_to(level); } /* :571-574 */

    ConflictKind analyze(CRef confl0, uint32_t &bt_level) {                /* conflict.rs:70-176 */
        if (assigns.is_ground_level()) return CONFL_GROUND;
        out_learnt.clear();
        {
            CRef confl = confl0;
            int path_c = 0;
            size_t index = assigns.trail.size();
            for (;;) {
                db.bump_activity(ca, confl);
                uint32_t n = ca.len(confl);
                const Lit *c = ca.lits(confl);
                T.an_words += 1 + n;
                for (uint32_t k = (confl == confl0 ? 0 : 1); k < n; k++) {
                    Lit q = c[k];
                    Var v = lit_var(q);
                    if (an.seen[v] == SEEN_UNDEF) {
                        uint32_t level = assigns.level[v];
                        if (level > GROUND_LEVEL) {
                            an.seen[v] = SEEN_SOURCE;
                            heur.bump_activity(v);
                            if (level >= assigns.current_level()) path_c += 1;
                            else out_learnt.push_back(q);
                        }
                    }
                }
                do { index -= 1; } while (an.seen[lit_var(assigns.trail[index])] == SEEN_UNDEF);
                Lit pl = assigns.trail[index];
                an.seen[lit_var(pl)] = SEEN_UNDEF;
                path_c -= 1;
                if (path_c <= 0) { out_learnt.insert(out_learnt.begin(), pl ^ 1); break; }
                confl = assigns.reason[lit_var(pl)];
            }
        }
        an.analyze_toclear = out_learnt;
        an.max_literals += out_learnt.size();
        if (an.ccmin_mode == 2) {
            size_t j = 0;
            for (size_t i = 0; i < out_learnt.size(); i++)
                if (!an.lit_redundant(ca, assigns, out_learnt[i], T)) out_learnt[j++] = out_learnt[i];
            out_learnt.resize(j);
        } else if (an.ccmin_mode == 1) {
            size_t j = 0;
            for (size_t i = 0; i < out_learnt.size(); i++)
                if (!an.lit_redundant_basic(ca, assigns, out_learnt[i], T)) out_learnt[j++] = out_learnt[i];
            out_learnt.resize(j);
        }
        an.tot_literals += out_learnt.size();
        for (Lit l : an.analyze_toclear) an.seen[lit_var(l)] = SEEN_UNDEF;
        if (out_learnt.size() == 1) { bt_level = GROUND_LEVEL; return CONFL_UNIT; }
        size_t max_i = 1;
        uint32_t max_level = assigns.level[lit_var(out_learnt[1])];
        for (size_t i = 2; i < out_learnt.size(); i++) {
            uint32_t level = assigns.level[lit_var(out_learnt[i])];
            if (level > max_level) { max_i = i; max_level = level; }
        }
        std::swap(out_learnt[1], out_learnt[max_i]);
        bt_level = max_level;
        return CONFL_LEARNED;
    }

    void trace_conflict(uint32_t bt_level) {
        uint64_t h = trace_hash;
        auto mix = [&](uint32_t x) { for (int b = 0; b < 4; b++) { h ^= (x >> (8 * b)) & 0xff; h *= 1099511628211ull; } };
        mix(bt_level);
        mix((uint32_t)out_learnt.size());
        for (Lit l : out_learnt) mix(l);
        trace_hash = h;
        if (g_trace.buf && g_trace.n < g_trace.cap) {
            uint64_t *r = g_trace.buf + 3 * (size_t)g_trace.n++;
            r[0] = bt_level; r[1] = out_learnt.size(); r[2] = h;
        }
    }

    /* handle_conflict :263-304 ; returns false on ground-level conflict */
    bool handle_conflict(LearningGuard &learnt, CRef confl, uint32_t &level, Lit &lit, CRef &reason) {
        conflicts += 1;
        uint32_t bt;
        ConflictKind k = analyze(confl, bt);
        if (k == CONFL_GROUND) return false;
        trace_conflict(bt);
        ctx_cancel_until(bt);
        level = bt; lit = out_learnt[0];
        if (k == CONFL_UNIT) reason = CREF_UNDEF;
        else {
            reason = db.learn_clause(ca, out_learnt.data(), (uint32_t)out_learnt.size());
            T.learnt_words += 1 + out_learnt.size() + 4;
        }
        heur.decay_activity();
        db.decay_activity();
        if (learnt.bump() && g_progress_rows) {                            /* the info! row of :289-301 (trail not yet backtracked) */
            size_t ground = assigns.lim.size() > 1 ? assigns.lim[1] : assigns.trail.size();
            g_progress_rows->push_back({(double)conflicts, (double)(heur.dec_vars - (int64_t)ground), (double)db.stats.num_clauses,
                                        (double)db.stats.clauses_literals, (double)(uint64_t)learnt.max_learnts, (double)db.stats.num_learnts,
                                        (double)db.stats.learnts_literals / (double)db.stats.num_learnts, progress_estimate() * 100.0});
        }
        return true;
    }

    bool propagate_learn_backtrack(LearningGuard &learnt) {                /* :503-517 */
        CRef confl;
        while ((confl = propagate()) != CREF_UNDEF) {
            uint32_t level; Lit lit; CRef reason;
            if (!handle_conflict(learnt, confl, level, lit, reason)) return false;
            assigns.backtrack_to(level);
            assigns.assign_lit(lit, reason);
            if (reason != CREF_UNDEF) attach(reason);
        }
        return true;
    }

    void try_simplify() {                                                  /* :522-568 */
        if (!assigns.is_ground_level()) return;
        if ((simp_db_assigns_set && simp_db_assigns == assigns.number_of_assigns()) || watches.propagations < simp_db_props) return;
        db.remove_satisfied(ca, assigns, [&](CRef c) { watches.unwatch_clause_lazy(ca, c); });
        try_garbage_collect();
        heur.rebuild_order_heap(assigns);
        simp_db_assigns_set = true;
        simp_db_assigns = assigns.number_of_assigns();
        simp_db_props = watches.propagations + db.stats.clauses_literals + db.stats.learnts_literals;
    }
    void try_garbage_collect() { if (ca.check_garbage(st.garbage_frac)) { ClauseGC g(ca); gc_into(g); g.finish(); } } /* :577-581 */
    void gc_into(ClauseGC &g) {                                            /* :583-587, backtrack.rs:65-70 */
        watches.gc(g);
        for (Lit l : assigns.trail) {                                      /* assignment.rs:236-243 */
            CRef &r = assigns.reason[lit_var(l)];
            if (r != CREF_UNDEF) { CRef nr; if (g.relocate(r, nr)) r = nr; else r = CREF_UNDEF; }
        }
        db.gc(g);
    }

    enum LoopRes { LR_RESTART, LR_UNSAT, LR_SAT, LR_INTERRUPTED };
    int64_t conflict_budget = -1, propagation_budget = -1;                 /* budget.rs:5-34 (no asynchronous interrupt here) */
    std::vector<Lit> assumptions;                                          /* solve_limited's &[Lit] */
    double progress = 0.0;
    bool within_budget() const {                                           /* budget.rs:21-25 */
        return (conflict_budget < 0 || conflicts < (uint64_t)conflict_budget) &&
               (propagation_budget < 0 || watches.propagations < (uint64_t)propagation_budget);
    }
    double progress_estimate() const {                                     /* search/util.rs:5-14 */
        double vars = 1.0 / (double)assigns.number_of_vars();
        double prog = 0.0, factor = vars;
        size_t nl = assigns.lim.size();
        for (size_t level = 1; level <= nl; level++) {                     /* assignment.rs:266-292 all_levels_dir */
            size_t head = assigns.lim[level - 1];
            size_t tail = level < nl ? assigns.lim[level] : assigns.trail.size();
            prog += factor * (double)(tail - head);
            factor *= vars;
        }
        return prog;
    }
    LoopRes search_loop(uint64_t nof_conflicts, LearningGuard &learnt) {   /* :452-501 (no assumptions) */
        starts += 1;
        uint64_t confl_limit = conflicts + nof_conflicts;
        for (;;) {
            if (!propagate_learn_backtrack(learnt)) return LR_UNSAT;
            if (!within_budget()) {                                        /* :473-477 */
                progress = progress_estimate();
                cancel_until(GROUND_LEVEL);
                return LR_INTERRUPTED;
            }
            if (conflicts >= confl_limit) { cancel_until(GROUND_LEVEL); return LR_RESTART; }
            try_simplify();
            if ((double)db.learnts.size() >= learnt.max_learnts + (double)assigns.number_of_assigns()) {
                db.reduce(ca, assigns, [&](CRef c) { watches.unwatch_clause_lazy(ca, c); });
                try_garbage_collect();
            }
            Lit next = 0;
            bool have = false;
            /* SearchCtx::decide :216-232: user assumptions first, one (possibly dummy) decision level each */
            while (assigns.lim.size() - 1 < assumptions.size()) {
                Lit p = assumptions[assigns.lim.size() - 1];
                if (assigns.is_true(p)) assigns.new_decision_level();
                else if (assigns.is_false(p)) {
                    /* analyze_final (conflict.rs:255-279) has no side effect on the counters; its result is dropped by
                     * search_internal ("TODO: implement properly", :431-435) */
                    cancel_until(GROUND_LEVEL);
                    return LR_UNSAT;
                } else { next = p; have = true; break; }
            }
            if (!have) {
                decisions += 1;                                            /* :234-236 */
                if (!heur.pick_branch_lit(assigns, next)) return LR_SAT;
            }
            assigns.new_decision_level();                                  /* backtrack.rs:58-62 */
            assigns.assign_lit(next, CREF_UNDEF);
        }
    }
    /* search_internal :409-442 ; ORACLE_SAT / ORACLE_UNSAT / ORACLE_INTERRUPTED */
    int search() {
        solves += 1;
        LearningGuard learnt{st.min_learnts_lim, st.size_factor, st.size_inc, st.size_adjust_start_confl, st.size_adjust_inc};
        learnt.reset(db.stats.num_clauses);
        uint32_t curr_restarts = 0;
        for (;;) {
            double rest_base = st.luby_restart ? luby(st.restart_inc, curr_restarts) : powi(st.restart_inc, (int)curr_restarts);
            uint64_t conflicts_to_go = (uint64_t)(rest_base * st.restart_first);   /* :38-46 */
            LoopRes r = search_loop(conflicts_to_go, learnt);
            if (r == LR_RESTART) { curr_restarts += 1; continue; }
            return r == LR_SAT ? ORACLE_SAT : r == LR_UNSAT ? ORACLE_UNSAT : ORACLE_INTERRUPTED;
        }
    }
    void get_stats(uint64_t *s) const {                                    /* :590-601, sat.rs:8-18 */
        s[0] = solves; s[1] = starts; s[2] = decisions; s[3] = heur.rnd_decisions; s[4] = conflicts;
        s[5] = watches.propagations; s[6] = an.tot_literals; s[7] = an.max_literals - an.tot_literals;
    }
};

/* ---------------------------------------------------------------- simplify/ submodules */
struct ElimBefore {                                                        /* elim_queue.rs:35-39 */
    const std::vector<int64_t> *n_occ;
    bool operator()(Var a, Var b) const {
        uint64_t ca = (uint64_t)(*n_occ)[a << 1] * (uint64_t)(*n_occ)[(a << 1) | 1];
        uint64_t cb = (uint64_t)(*n_occ)[b << 1] * (uint64_t)(*n_occ)[(b << 1) | 1];
        return ca < cb;
    }
};

struct ElimClauses {                                                       /* elim_clauses.rs */
    std::vector<Lit> literals;
    std::vector<size_t> sizes;
    void mk_elim_unit(Lit x) { literals.push_back(x); sizes.push_back(literals.size()); }   /* :20-23 */
    void mk_elim_clause(Var v, const Lit *c, uint32_t n) {                                  /* :25-41 */
        size_t first = literals.size(), index = 0;
        for (uint32_t i = 0; i < n; i++) if (lit_var(c[i]) == v) { index = i; break; }
        literals.insert(literals.end(), c, c + n);
        sizes.push_back(literals.size());
        std::swap(literals[first], literals[first + index]);
    }
    void extend_model(std::vector<uint8_t> &model) const {                                  /* :43-63 */
        for (size_t i = sizes.size(); i-- > 0;) {
            size_t head = i > 0 ? sizes[i - 1] : 0, tail = sizes[i];
            bool sat = false;
            for (size_t k = head; k < tail; k++) {                          /* formula/util.rs:24-32 */
                Lit l = literals[k];
                uint8_t m = model[lit_var(l)];
                if (m != 2 && (m != 0) != lit_sign(l)) { sat = true; break; }
            }
            if (!sat) { Lit l = literals[head]; model[lit_var(l)] = lit_sign(l) ? 0 : 1; }
        }
    }
};

struct Simplificator {
    oracle_settings st;
    uint64_t asymm_lits = 0, eliminated_vars = 0;
    /* ElimOcc elim_queue.rs:154-203 */
    std::vector<uint8_t> frozen, eliminated;
    std::vector<std::vector<CRef>> occs; std::vector<uint8_t> occ_dirty, occ_present;  /* OccLists :85-151 */
    std::vector<int64_t> n_occ; IdxHeap<ElimBefore> elim_heap;                        /* ElimQueue :13-76 */
    /* Touched simplify.rs:53-112 */
    std::vector<uint8_t> touched; size_t n_touched = 0;
    /* SubsumptionQueue */
    std::deque<CRef> queue; size_t bwdsub_assigns = 0;

    explicit Simplificator(const oracle_settings &s) : st(s) { elim_heap.before.n_occ = &n_occ; }

    void init_var(Var v) {                                                 /* simplify.rs:134-137 */
        frozen.push_back(0); eliminated.push_back(0);
        occs.emplace_back(); occ_dirty.push_back(0); occ_present.push_back(1);
        n_occ.push_back(0); n_occ.push_back(0);
        elim_heap.insert(v);                                               /* elim_queue.rs:26-32 */
        touched.push_back(0);
    }
    bool validate_resolvent_len(size_t len) const { return st.clause_lim == -1 || (int)len <= st.clause_lim; }
    bool validate_subsumption_len(size_t len) const { return st.subsumption_lim == -1 || (int)len < st.subsumption_lim; }

    void update_elim_heap(Var v, const Assignment &as) {                   /* elim_queue.rs:46-58 */
        if (elim_heap.contains(v)) elim_heap.update(v);
        else if (!frozen[v] && !eliminated[v] && as.is_undef(v)) elim_heap.insert(v);
    }
    void bump_lit_occ(Lit l, int64_t d) { n_occ[l] += d; elim_heap.update(lit_var(l)); }    /* :64-70 */
    const std::vector<CRef> &lookup(const ClauseAllocator &ca, Var v) {    /* :118-125 */
        if (occ_dirty[v]) {
            auto &o = occs[v];
            size_t j = 0;
            for (size_t i = 0; i < o.size(); i++) if (!ca.deleted(o[i])) o[j++] = o[i];
            o.resize(j);
            occ_dirty[v] = 0;
        }
        return occs[v];
    }
    void elo_add_clause(CRef cr, const Lit *c, uint32_t n) {               /* :175-180 */
        for (uint32_t i = 0; i < n; i++) { occs[lit_var(c[i])].push_back(cr); bump_lit_occ(c[i], 1); }
    }
    void smudge_clause(const Assignment &as, const Lit *c, uint32_t n) {   /* :182-188 */
        for (uint32_t i = 0; i < n; i++) {
            bump_lit_occ(c[i], -1);
            update_elim_heap(lit_var(c[i]), as);
            occ_dirty[lit_var(c[i])] = 1;
        }
    }
    void remove_lit(const Assignment &as, Lit l, CRef cr) {                /* :190-194 */
        bump_lit_occ(l, -1);
        update_elim_heap(lit_var(l), as);
        auto &o = occs[lit_var(l)];
        o.erase(std::remove(o.begin(), o.end(), cr), o.end());
    }

    bool add_clause(Searcher &S, const Lit *ps, uint32_t n) {              /* simplify.rs:139-163 ; false = Err */
        CRef cr;
        Searcher::AddRes r = S.add_clause(ps, n, cr);
        if (r == Searcher::ADD_UNSAT) return false;
        if (r == Searcher::ADD_CONSUMED) return true;
        queue.push_back(cr);
        uint32_t len = S.ca.len(cr);
        const Lit *c = S.ca.lits(cr);
        elo_add_clause(cr, c, len);
        for (uint32_t i = 0; i < len; i++) { touched[lit_var(c[i])] = 1; n_touched += 1; }  /* :74-79 */
        return true;
    }
    size_t assigns_left(const Assignment &as) const { return as.number_of_ground_assigns() - bwdsub_assigns; }

    void enqueue_touched_clauses(Searcher &S) {                            /* :82-111 */
        if (n_touched == 0) return;
        for (CRef cr : queue) S.ca.set_touched(cr, true);
        for (Var v = 0; v < touched.size(); v++) {
            if (!touched[v] || eliminated[v]) continue;
            const std::vector<CRef> &o = lookup(S.ca, v);
            for (CRef cr : o) {
                if (!S.ca.touched(cr)) { queue.push_back(cr); S.ca.set_touched(cr, true); }
            }
            touched[v] = 0;
        }
        for (CRef cr : queue) S.ca.set_touched(cr, false);
        n_touched = 0;
    }

    bool try_propagate(Searcher &S, Lit p) {                               /* :553-565 */
        if (S.assigns.is_false(p)) return false;
        if (S.assigns.is_undef(lit_var(p))) S.assigns.assign_lit(p, CREF_UNDEF);
        return S.propagate() == CREF_UNDEF;
    }
    bool strengthen_clause(Searcher &S, CRef cr, Lit l) {                  /* :276-301 */
        uint32_t len = S.ca.len(cr);
        if (len == 2) {
            smudge_clause(S.assigns, S.ca.lits(cr), len);
            S.watches.unwatch_clause_lazy(S.ca, cr);
            S.db.remove_clause(S.ca, cr);
            const Lit *c = S.ca.lits(cr);
            Lit unit = (l == c[0]) ? c[1] : c[0];
            return try_propagate(S, unit);
        }
        S.watches.unwatch_clause_strict(S.ca, cr);
        S.db.st_del(S.ca, cr);                                             /* edit_clause clause_db.rs:107-117 */
        {                                                                  /* strengthen :567-591 */
            Lit *c = S.ca.lits(cr);
            uint32_t k = 0;
            while (k < len && c[k] != l) k++;
            if (k == len) { fprintf(stderr, "oracle: strengthen: literal not found\n"); abort(); }
            for (; k + 1 < len; k++) c[k] = c[k + 1];
            S.ca.shrink_by(cr, 1);
            if (!S.ca.learnt(cr)) {
                uint32_t a = 0;
                for (uint32_t i = 0; i + 1 < len; i++) a |= lit_abstraction(c[i]);
                S.ca.mem[cr + 1] = TAG_ABS;
                S.ca.value(cr) = a;
            }
        }
        S.db.st_add(S.ca, cr);
        S.attach(cr);
        remove_lit(S.assigns, l, cr);
        return true;
    }

    enum Sub { SUB_DIFFERENT, SUB_EXACT, SUB_LITSIGN };
    static Sub subsumes(const ClauseAllocator &ca, CRef a, CRef b, Lit &out) {   /* subsumes.rs:10-49 */
        uint32_t na = ca.len(a), nb = ca.len(b);
        if (nb < na) return SUB_DIFFERENT;
        if (ca.tag(a) == TAG_ABS && ca.tag(b) == TAG_ABS)
            if ((ca.mem[a + 2] & ~ca.mem[b + 2]) != 0) return SUB_DIFFERENT;
        Sub ret = SUB_EXACT;
        const Lit *x = ca.lits(a), *y = ca.lits(b);
        for (uint32_t i = 0; i < na; i++) {
            Lit lit = x[i];
            bool found = false;
            for (uint32_t j = 0; j < nb; j++) {
                Lit cur = y[j];
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
    static Sub unit_subsumes(const ClauseAllocator &ca, Lit unit, CRef b, Lit &out) { /* subsumes.rs:51-67 */
        if (ca.tag(b) == TAG_ABS && (lit_abstraction(unit) & ~ca.mem[b + 2]) != 0) return SUB_DIFFERENT;
        uint32_t nb = ca.len(b);
        const Lit *y = ca.lits(b);
        for (uint32_t j = 0; j < nb; j++) {
            if (unit == y[j]) return SUB_EXACT;
            else if (unit == (y[j] ^ 1)) { out = unit; return SUB_LITSIGN; }
        }
        return SUB_DIFFERENT;
    }

    bool backward_subsumption_check(Searcher &S) {                         /* simplify.rs:423-523 */
        for (;;) {
            /* SubsumptionQueue::pop subsumption_queue.rs:23-42 */
            bool is_clause = false; CRef sub_cr = CREF_UNDEF; Lit unit = 0; bool have = false;
            for (;;) {
                if (!queue.empty()) {
                    CRef cr = queue.front(); queue.pop_front();
                    if (!S.ca.deleted(cr)) { is_clause = true; sub_cr = cr; have = true; break; }
                } else {
                    size_t ground_end = S.assigns.level_end(GROUND_LEVEL);
                    if (bwdsub_assigns < ground_end) { unit = S.assigns.trail[bwdsub_assigns++]; have = true; }
                    break;
                }
            }
            if (!have) break;

            Var lookup_var;
            if (!is_clause) lookup_var = lit_var(unit);
            else {                                                          /* :465-475 */
                const Lit *c = S.ca.lits(sub_cr);
                uint32_t n = S.ca.len(sub_cr);
                Var best = lit_var(c[0]);
                for (uint32_t i = 1; i < n; i++)
                    if (occs[lit_var(c[i])].size() < occs[best].size()) best = lit_var(c[i]);
                lookup_var = best;
            }
            std::vector<CRef> cand = lookup(S.ca, lookup_var);             /* snapshot clone :478 */
            for (CRef cj : cand) {
                if (is_clause) {
                    if (S.ca.deleted(sub_cr)) break;
                    if (cj == sub_cr) continue;
                }
                if (S.ca.deleted(cj) || !validate_subsumption_len(S.ca.len(cj))) continue;
                Lit l = 0;
                Sub sub = is_clause ? subsumes(S.ca, sub_cr, cj, l) : unit_subsumes(S.ca, unit, cj, l);
                if (sub == SUB_EXACT) {
                    smudge_clause(S.assigns, S.ca.lits(cj), S.ca.len(cj));
                    S.watches.unwatch_clause_lazy(S.ca, cj);
                    S.db.remove_clause(S.ca, cj);
                } else if (sub == SUB_LITSIGN) {
                    queue.push_back(cj);                                    /* try_push */
                    if (!strengthen_clause(S, cj, l ^ 1)) return false;
                }
            }
        }
        return true;
    }

    /* formula/util.rs:45-77 ; false = tautology */
    static bool merge(Var v, const Lit *ps_, uint32_t np, const Lit *qs_, uint32_t nq, std::vector<Lit> &res) {
        const Lit *longer = ps_, *shorter = qs_; uint32_t nl = np, ns = nq;
        if (np < nq) { longer = qs_; nl = nq; shorter = ps_; ns = np; }
        res.clear();
        for (uint32_t i = 0; i < ns; i++) {
            Lit q = shorter[i];
            if (lit_var(q) != v) {
                bool ok = true;
                for (uint32_t j = 0; j < nl; j++) {
                    if (lit_var(longer[j]) == lit_var(q)) {
                        if (longer[j] == (q ^ 1)) return false;
                        ok = false; break;
                    }
                }
                if (ok) res.push_back(q);
            }
        }
        for (uint32_t j = 0; j < nl; j++) if (lit_var(longer[j]) != v) res.push_back(longer[j]);
        return true;
    }

    /* search/util.rs:44-73 */
    bool asymmetric_branching(Searcher &S, Var v, CRef cr, Lit &out) {
        if (S.ca.deleted(cr)) return false;
        uint32_t n = S.ca.len(cr);
        const Lit *c = S.ca.lits(cr);
        for (uint32_t i = 0; i < n; i++) if (S.assigns.is_true(c[i])) return false;
        S.assigns.new_decision_level();
        Lit vl = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (lit_var(c[i]) == v) vl = c[i];
            else if (S.assigns.is_undef(lit_var(c[i]))) S.assigns.assign_lit(c[i] ^ 1, CREF_UNDEF);
        }
        CRef confl = S.propagate();
        for (size_t k = S.assigns.level_begin(2); k < S.assigns.trail.size(); k++) S.heur.save_phase(S.assigns.trail[k], true);
        S.assigns.backtrack_to(GROUND_LEVEL);
        if (confl == CREF_UNDEF) return false;
        out = vl;
        return true;
    }
    bool asymm_var(Searcher &S, Var v) {                                   /* simplify.rs:303-336 ; false = Err */
        if (!S.assigns.is_undef(v)) return true;
        std::vector<CRef> cls = lookup(S.ca, v);
        if (cls.empty()) return true;
        bool bug = false;
        for (CRef cr : cls) {
            if (bug) { bug = false; continue; }
            Lit l;
            if (asymmetric_branching(S, v, cr, l)) {
                if (S.ca.len(cr) > 2) bug = true;
                asymm_lits += 1;
                queue.push_back(cr);
                if (!strengthen_clause(S, cr, l)) return false;
            }
        }
        return true;
    }

    bool eliminate_var(Searcher &S, ElimClauses &ec, Var v) {              /* :338-420 ; false = Err */
        std::vector<CRef> cls = lookup(S.ca, v);
        std::vector<CRef> pos, neg;
        for (CRef cr : cls) {
            uint32_t n = S.ca.len(cr);
            const Lit *c = S.ca.lits(cr);
            for (uint32_t i = 0; i < n; i++)
                if (lit_var(c[i]) == v) { if (lit_sign(c[i])) neg.push_back(cr); else pos.push_back(cr); break; }
        }
        size_t max_resolvents = (size_t)st.grow + pos.size() + neg.size();
        std::vector<std::vector<Lit>> resolvents;
        std::vector<Lit> tmp;
        for (CRef pr : pos)
            for (CRef nr : neg)
                if (merge(v, S.ca.lits(pr), S.ca.len(pr), S.ca.lits(nr), S.ca.len(nr), tmp)) {
                    size_t len = tmp.size();
                    resolvents.push_back(tmp);
                    if (resolvents.size() > max_resolvents || !validate_resolvent_len(len)) return true;
                }
        eliminated[v] = 1;
        S.heur.set_decision_var(v, false);
        eliminated_vars += 1;
        if (pos.size() > neg.size()) {
            for (CRef cr : neg) ec.mk_elim_clause(v, S.ca.lits(cr), S.ca.len(cr));
            ec.mk_elim_unit(mk_lit(v, false));
        } else {
            for (CRef cr : pos) ec.mk_elim_clause(v, S.ca.lits(cr), S.ca.len(cr));
            ec.mk_elim_unit(mk_lit(v, true));
        }
        for (CRef cr : cls) {
            smudge_clause(S.assigns, S.ca.lits(cr), S.ca.len(cr));
            S.watches.unwatch_clause_lazy(S.ca, cr);
            S.db.remove_clause(S.ca, cr);
        }
        for (auto &r : resolvents) if (!add_clause(S, r.data(), (uint32_t)r.size())) return false;
        occs[v].clear(); occs[v].shrink_to_fit(); occ_present[v] = 0;      /* clear_var */
        return true;
    }

    void try_garbage_collect(Searcher &S) {                                /* :526-532 */
        if (S.ca.check_garbage(st.simp_garbage_frac)) {
            ClauseGC g(S.ca);
            S.gc_into(g);
            for (Var v = 0; v < occs.size(); v++) {                        /* OccLists::gc elim_queue.rs:138-150 */
                if (!occ_present[v]) continue;
                auto &o = occs[v];
                size_t j = 0;
                for (size_t i = 0; i < o.size(); i++) { CRef nr; if (g.relocate(o[i], nr)) o[j++] = nr; }
                o.resize(j);
                occ_dirty[v] = 0;
            }
            {                                                              /* SubsumptionQueue::gc */
                size_t j = 0;
                for (size_t i = 0; i < queue.size(); i++) { CRef nr; if (g.relocate(queue[i], nr)) queue[j++] = nr; }
                queue.resize(j);
            }
            g.finish();
        }
    }

    bool eliminate(Searcher &S, ElimClauses &ec) {                         /* :213-274 ; false = Err */
        while (n_touched != 0 || assigns_left(S.assigns) > 0 || elim_heap.len() > 0) {
            enqueue_touched_clauses(S);
            if (!backward_subsumption_check(S)) return false;
            Var var;
            while (elim_heap.pop(var)) {
                if (eliminated[var] || !S.assigns.is_undef(var)) continue;
                if (st.use_asymm) {
                    bool was_frozen = frozen[var];
                    frozen[var] = 1;
                    if (!asymm_var(S, var)) return false;
                    if (!backward_subsumption_check(S)) return false;
                    frozen[var] = was_frozen;
                }
                if (st.use_elim && S.assigns.is_undef(var) && !frozen[var]) {
                    if (!eliminate_var(S, ec, var)) return false;
                    if (!backward_subsumption_check(S)) return false;
                }
                try_garbage_collect(S);
            }
        }
        return true;
    }
    static void on(Searcher &S) { S.ca.extra_clause_field = true; S.db.remove_satisfied_flag = false; }   /* :546-549 */
    static void off(Searcher &S) {                                         /* :536-544 */
        S.db.remove_satisfied_flag = true;
        S.ca.extra_clause_field = false;
        S.heur.rebuild_order_heap(S.assigns);
        ClauseGC g(S.ca); S.gc_into(g); g.finish();
    }
};

/* ---------------------------------------------------------------- sat/minisat.rs : CoreSolver / SimpSolver */
struct Solver {
    int kind;
    bool ok = true;
    Searcher search;
    ElimClauses elimclauses;
    Simplificator *simp = nullptr;
    bool extend_model;

    Solver(const oracle_settings &s, int kind_) : kind(kind_), search(s), extend_model(s.extend_model) {
        if (kind == ORACLE_SIMP) { Simplificator::on(search); simp = new Simplificator(s); }   /* minisat.rs:263-271 */
    }
    ~Solver() { delete simp; }
    size_t n_vars() const { return search.number_of_vars(); }
    size_t n_clauses() const { return search.number_of_clauses(); }
    Var new_var(int upol, bool dvar) {                                     /* minisat.rs:37-39,132-138 */
        Var v = search.new_var(upol, dvar);
        if (simp) simp->init_var(v);
        return v;
    }
    bool add_clause(const Lit *ps, uint32_t n) {
        if (simp) {                                                        /* minisat.rs:146-158 (no `ok` check) */
            if (!simp->add_clause(search, ps, n)) { ok = false; return false; }
            return true;
        }
        if (ok) { CRef cr; if (search.add_clause(ps, n, cr) == Searcher::ADD_UNSAT) ok = false; }  /* :41-48 */
        return ok;
    }
    bool preprocess() {
        if (ok) ok = search.preprocess();                                  /* :50-55 */
        if (!ok) return false;
        if (kind == ORACLE_CORE || !simp) return true;
        bool result = simp->eliminate(search, elimclauses);                /* :161-191 */
        if (!result) ok = false;
        Simplificator::off(search);
        delete simp; simp = nullptr;
        return result;
    }
    /* returns ORACLE_SAT / ORACLE_UNSAT */
    int solve_limited(std::vector<uint8_t> &model) {
        if (simp) {                                                        /* Simplificator::solve_limited simplify.rs:165-211 */
            if (search.propagate() == CREF_UNDEF) search.try_simplify(); else return ORACLE_UNSAT;
            if (!simp->eliminate(search, elimclauses)) return ORACLE_UNSAT;
        } else if (kind == ORACLE_CORE && !ok) return ORACLE_UNSAT;        /* minisat.rs:80-82 */
        { int r = search.search(); if (r != ORACLE_SAT) return r; }    /* Interrupted hands the solver back (sat.rs:24) */
        model.assign(n_vars(), 2);                                         /* extract_model formula/util.rs:35-41 */
        for (Lit l : search.assigns.trail) model[lit_var(l)] = lit_sign(l) ? 0 : 1;
        if (kind == ORACLE_SIMP && extend_model) elimclauses.extend_model(model);
        return ORACLE_SAT;
    }
};

/* ---------------------------------------------------------------- sat/dimacs.rs:171-365 */
struct DimacsParser {
    const std::string &buf; size_t pos = 0;
    std::string err;
    explicit DimacsParser(const std::string &b) : buf(b) {}
    int cur() const { return pos < buf.size() ? (unsigned char)buf[pos] : -1; }
    static bool is_ws(int c) { return c == ' ' || (c >= 9 && c <= 13); }
    void skip_ws() { while (cur() >= 0 && is_ws(cur())) pos++; }
    void skip_line() { while (cur() >= 0) { if (cur() == '\n') { pos++; break; } pos++; } }
    bool consume(const char *t) { for (; *t; t++) { if (cur() == (unsigned char)*t) pos++; else { err = std::string("failed to consume; expected '") + "p cnf" + "'"; return false; } } return true; }
    bool read_int_body(size_t &v) {
        size_t len = 0; v = 0;
        while (cur() >= '0' && cur() <= '9') { v = v * 10 + (size_t)(cur() - '0'); len++; pos++; }
        if (len == 0) { err = "int expected"; return false; }
        return true;
    }
    bool next_int(int32_t &out) {
        skip_ws();
        int sign = 1;
        if (cur() == '+') pos++; else if (cur() == '-') { pos++; sign = -1; }
        size_t v; if (!read_int_body(v)) return false;
        out = sign * (int32_t)v; return true;
    }
    bool next_uint(size_t &out) { skip_ws(); if (cur() == '+') pos++; return read_int_body(out); }
};

static bool read_file_maybe_gz(const char *path, std::string &out, std::string &err) {
    gzFile f = gzopen(path, "rb");           /* gzopen transparently reads plain files (dimacs.rs:21-31 sniffing) */
    if (!f) { err = std::string("cannot open ") + path; return false; }
    char tmp[1 << 16];
    int n;
    while ((n = gzread(f, tmp, sizeof tmp)) > 0) out.append(tmp, (size_t)n);
    gzclose(f);
    if (n < 0) { err = "read error"; return false; }
    return true;
}

/* SplitMix64 generator of SURVEY 8d */
static inline uint64_t splitmix64(uint64_t &state) {
    state += 0x9E3779B97F4A7C15ull;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

void oracle_default_settings(oracle_settings *s) {
    memset(s, 0, sizeof *s);
    s->var_decay = 0.95; s->clause_decay = 0.999; s->random_seed = 91648253.0; s->random_var_freq = 0.0;
    s->restart_first = 100.0; s->restart_inc = 2.0; s->size_factor = 1.0 / 3.0; s->size_inc = 1.1;
    s->size_adjust_inc = 1.5; s->garbage_frac = 0.20; s->simp_garbage_frac = 0.5; s->grow = 0;
    s->size_adjust_start_confl = 100; s->min_learnts_lim = 0; s->clause_lim = 20; s->subsumption_lim = 1000;
    s->ccmin_mode = 2; s->phase_saving = 2; s->luby_restart = 1; s->rnd_pol = 0; s->rnd_init_act = 0;
    s->use_rcheck = 0; s->use_asymm = 0; s->use_elim = 1; s->extend_model = 1; s->remove_satisfied = 1;
}

void oracle_set_trace(uint64_t *buf, uint32_t cap) { g_trace.buf = buf; g_trace.cap = cap; g_trace.n = 0; }

static std::vector<ProgressRow> g_progress_store;
/* record the search-table rows (search.rs:289-301) of the following oracle_solve_csr calls; read them back afterwards */
void oracle_record_progress(int on) { g_progress_store.clear(); g_progress_rows = on ? &g_progress_store : nullptr; }
uint32_t oracle_get_progress(double *out, uint32_t cap_rows) {
    uint32_t n = (uint32_t)std::min<size_t>(cap_rows, g_progress_store.size());
    for (uint32_t i = 0; i < n; i++) memcpy(out + 8 * (size_t)i, g_progress_store[i].v, sizeof(double) * 8);
    return n;
}
static std::vector<uint8_t> g_var_flags;   /* bit0 upol set, bit1 upol value, bit2 not a decision var */
static std::vector<Lit> g_assumptions;
static std::vector<uint32_t> g_clause_nvars;   /* with g_var_flags: vars that existed when clause c was added */
void oracle_set_clause_nvars(const uint32_t *nv, uint32_t n) { g_clause_nvars.assign(nv, nv + (nv ? n : 0)); }
void oracle_set_var_flags(const uint8_t *flags, uint32_t n) { g_var_flags.assign(flags, flags + (flags ? n : 0)); }
void oracle_set_assumptions(const uint32_t *lits, uint32_t n) { g_assumptions.assign(lits, lits + (lits ? n : 0)); }
static int64_t g_conflict_budget = -1, g_propagation_budget = -1;
static int g_resume_rounds = 0;
void oracle_set_budget(int64_t conflict_budget, int64_t propagation_budget, int resume_rounds) {
    g_conflict_budget = conflict_budget; g_propagation_budget = propagation_budget; g_resume_rounds = resume_rounds;
}

int oracle_solve_csr(const uint32_t *clause_off, const uint32_t *lits, uint32_t n_clauses,
                     const oracle_settings *s, int kind, int flags,
                     oracle_result *out, uint8_t *model_out, uint32_t model_cap) {
    auto t0 = std::chrono::steady_clock::now();
    memset(out, 0, sizeof *out);
    Solver solver(*s, kind);
    std::vector<Lit> tmp;
    if ((flags & ORACLE_F_NO_PRE) && kind == ORACLE_SIMP) solver.preprocess();   /* lib.rs:36-41 */
    /* new_var(upol, dvar) records: either the trait's own call order (sat.rs:31-32: every var before the clauses), or,
     * with a per-clause watermark, new_var interleaved with add_clause the way the caller issued them */
    auto create_flagged = [&](size_t upto) {
        while (solver.n_vars() < upto && solver.n_vars() < g_var_flags.size()) {
            uint8_t f = g_var_flags[solver.n_vars()];
            solver.new_var((f & 1) ? ((f & 2) ? 1 : 0) : -1, !(f & 4));
        }
    };
    const bool interleaved = !g_var_flags.empty() && g_clause_nvars.size() >= n_clauses;
    if (!interleaved) create_flagged(g_var_flags.size());
    for (uint32_t c = 0; c < n_clauses; c++) {                             /* Subst::add_clause dimacs.rs:146-167 */
        uint32_t b = clause_off[c], e = clause_off[c + 1];
        if (interleaved) create_flagged(g_clause_nvars[c]);
        for (uint32_t k = b; k < e; k++)
            while (lit_var(lits[k]) + 1 > solver.n_vars()) solver.new_var(-1, true);
        solver.add_clause(lits + b, e - b);
    }
    create_flagged(g_var_flags.size());
    out->n_vars = (uint32_t)solver.n_vars();
    out->n_clauses_ingest = (uint32_t)solver.n_clauses();
    auto t1 = std::chrono::steady_clock::now();
    bool pre_ok = true;
    if (flags & ORACLE_F_PREPROCESS) pre_ok = solver.preprocess();
    out->pre_ok = pre_ok;
    out->n_clauses_pre = (uint32_t)solver.n_clauses();
    std::vector<uint8_t> model;
    int verdict;
    bool zero_stats = false;
    if (!pre_ok) { verdict = ORACLE_UNSAT; zero_stats = (flags & ORACLE_F_ZERO_STATS_ON_PRE_UNSAT) != 0; }
    else if (flags & ORACLE_F_NO_SOLVE) verdict = ORACLE_INTERRUPTED;
    else {
        solver.search.conflict_budget = g_conflict_budget;
        solver.search.propagation_budget = g_propagation_budget;
        solver.search.assumptions = g_assumptions;
        if (solver.simp) for (Lit a : g_assumptions) solver.simp->frozen[lit_var(a)] = 1;   /* simplify.rs:173-186 */
        verdict = solver.solve_limited(model);
        /* SolveRes::Interrupted(progress, solver): the caller may call solve_limited on it again (sat.rs:24) */
        for (int r = 0; r < g_resume_rounds && verdict == ORACLE_INTERRUPTED; r++) {
            solver.search.conflict_budget = -1; solver.search.propagation_budget = -1;
            verdict = solver.solve_limited(model);
        }
    }
    auto t2 = std::chrono::steady_clock::now();
    (void)t0;
    out->verdict = verdict;
    if (!zero_stats) solver.search.get_stats(out->stats);
    out->eliminated_vars = solver.simp ? (uint32_t)solver.simp->eliminated_vars : 0;
    {
        uint64_t ev = 0;
        for (size_t i = 0; i < solver.elimclauses.sizes.size(); i++) {
            size_t head = i ? solver.elimclauses.sizes[i - 1] : 0;
            if (solver.elimclauses.sizes[i] - head == 1) ev++;
        }
        out->eliminated_vars = (uint32_t)ev;
    }
    const Traffic &T = solver.search.T;
    out->traffic[0] = T.visited; out->traffic[1] = T.kept; out->traffic[2] = T.deref; out->traffic[3] = T.tail;
    out->traffic[4] = T.moves; out->traffic[5] = T.implied; out->traffic[6] = T.an_words; out->traffic[7] = T.learnt_words;
    out->trace_hash = solver.search.trace_hash;
    out->seconds = std::chrono::duration<double>(t2 - t1).count();
    out->progress = verdict == ORACLE_INTERRUPTED && !(flags & ORACLE_F_NO_SOLVE) ? solver.search.progress : 0.0;
    if (verdict == ORACLE_SAT && model_out) {
        uint32_t n = std::min<uint32_t>(model_cap, (uint32_t)model.size());
        memcpy(model_out, model.data(), n);
    }
    return 0;
}

double oracle_solve_batch_csr(const uint32_t *inst_clause_off, const uint32_t *clause_off,
                              const uint32_t *lits, uint32_t n_inst,
                              const oracle_settings *s, int kind, int flags, int threads,
                              oracle_result *out, uint8_t *models, uint32_t model_stride) {
    if (threads < 1) threads = 1;
    std::atomic<uint32_t> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n_inst) break;
            uint32_t cb = inst_clause_off[i], ce = inst_clause_off[i + 1];
            oracle_solve_csr(clause_off + cb, lits, ce - cb, s, kind, flags, &out[i],
                             models ? models + (size_t)i * model_stride : nullptr, model_stride);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int oracle_parse_dimacs(const char *path, int strict, uint32_t **clause_off_out, uint32_t **lits_out,
                        uint32_t *n_clauses, uint32_t *hdr_vars, uint32_t *hdr_clauses,
                        char *err, size_t errcap) {
    std::string buf, e;
    auto fail = [&](const std::string &m) { if (err && errcap) snprintf(err, errcap, "%s", m.c_str()); return 1; };
    if (!read_file_maybe_gz(path, buf, e)) return fail(e);
    DimacsParser p(buf);
    std::vector<uint32_t> off{0}, lits;
    bool parsing = false;
    size_t vars = 0, clauses = 0, found = 0;
    std::vector<uint8_t> seen_var;
    size_t distinct = 0;
    for (;;) {                                                             /* parse_me dimacs.rs:197-251 */
        p.skip_ws();
        int c = p.cur();
        if (!parsing) {
            if (c == 'c') p.skip_line();
            else {
                if (!p.consume("p cnf")) return fail(p.err);
                if (!p.next_uint(vars) || !p.next_uint(clauses)) return fail(p.err);
                parsing = true;
            }
        } else {
            if (c == 'c') p.skip_line();
            else if (c < 0) {
                if (strict) {
                    if (clauses != found) return fail("PARSE ERROR! DIMACS header mismatch: " + std::to_string(clauses) + " clauses declared, " + std::to_string(found) + " found");
                    if (vars < distinct) return fail("PARSE ERROR! DIMACS header mismatch: " + std::to_string(vars) + " vars declared, " + std::to_string(distinct) + " discovered");
                }
                break;
            } else {
                for (;;) {                                                 /* parse_clause :253-265 */
                    int32_t lit;
                    if (!p.next_int(lit)) return fail(p.err);
                    if (lit == 0) { found++; off.push_back((uint32_t)lits.size()); break; }
                    uint32_t v = (uint32_t)(lit < 0 ? -lit : lit);
                    if (v >= seen_var.size()) seen_var.resize(v + 1, 0);
                    if (!seen_var[v]) { seen_var[v] = 1; distinct++; }
                    lits.push_back(((v - 1) << 1) | (lit < 0 ? 1u : 0u));
                }
            }
        }
    }
    *n_clauses = (uint32_t)found; *hdr_vars = (uint32_t)vars; *hdr_clauses = (uint32_t)clauses;
    *clause_off_out = (uint32_t *)malloc(off.size() * 4);
    memcpy(*clause_off_out, off.data(), off.size() * 4);
    *lits_out = (uint32_t *)malloc(std::max<size_t>(1, lits.size()) * 4);
    memcpy(*lits_out, lits.data(), lits.size() * 4);
    return 0;
}

void oracle_free(void *p) { free(p); }

void oracle_gen_ksat(uint64_t seed, uint32_t n, uint32_t m, uint32_t k, uint32_t *lits) {
    uint64_t state = seed;
    for (uint32_t c = 0; c < m; c++) {
        uint32_t *cl = lits + (size_t)c * k;
        for (uint32_t j = 0; j < k;) {
            uint64_t r = splitmix64(state);
            uint32_t v = (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
            bool dup = false;
            for (uint32_t i = 0; i < j; i++) if ((cl[i] >> 1) == v) dup = true;
            if (dup) continue;
            cl[j++] = (v << 1) | (uint32_t)(r & 1);
        }
    }
}

int oracle_check_model(const uint32_t *clause_off, const uint32_t *lits, uint32_t n_clauses,
                       const uint8_t *model, uint32_t n_vars) {
    for (uint32_t c = 0; c < n_clauses; c++) {
        bool sat = false;
        for (uint32_t k = clause_off[c]; k < clause_off[c + 1]; k++) {
            uint32_t v = lits[k] >> 1;
            if (v < n_vars && model[v] != 2 && (model[v] == 1) == ((lits[k] & 1) == 0)) { sat = true; break; }
        }
        if (!sat) return 0;
    }
    return 1;
}

}  // extern "C"
