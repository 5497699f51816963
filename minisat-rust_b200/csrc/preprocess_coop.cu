/*
 * preprocess_coop.cu -- K4 as ONE preprocessor (round 2): backward subsumption, self-subsuming strengthening and
 * bounded variable elimination of one large instance, interleaved until nothing changes, inside ONE persistent
 * cooperative kernel (grid-wide barriers between the passes instead of kernel launches).
 *
 * Reference semantics, not the reference's job order (SURVEY.md 8a S*): `subsumes()` subsumes.rs:10-49, signatures
 * formula/util.rs:5-11, `strengthen` simplify.rs:567-591, `eliminate_var` simplify.rs:338-420 with `merge`
 * formula/util.rs:45-77 and the eliminated-clause stack of elim_clauses.rs:20-41, the interleaving of
 * `Simplificator::eliminate` simplify.rs:213-274 (subsumption queue emptied, then one elimination candidate, then the
 * resolvents are checked for subsumption again ...) as passes: [rebuild occurrence lists] -> [subsume + strengthen
 * rounds to a fixpoint] -> [BVE rounds over independent sets of variables] -> repeat while anything changed.  Every
 * step is implied by the clause set it is computed on (or is a resolution step that keeps satisfiability and is
 * recorded for model extension), so the result is equisatisfiable and every model of it extends to a model of the
 * input (k4_extend_model, elim_clauses.rs:43-63).
 *
 * The scans are multi-block (chunk sums -> one block scans the sums -> chunks add their offset), the kernel is launched
 * once per instance with one block of 256 threads per SM slot (cudaLaunchCooperativeKernel), plus one launch of the
 * compaction kernel: 2 launches per instance instead of round 1's 4 per BVE round.
 */
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;
typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

enum : u32 { C_CHANGED = 0, C_NSEL = 1, C_NELIM = 2, C_OVERFLOW = 3, C_EMPTY = 4, C_NSUB = 5, C_NSTR = 6, C_OUTER = 7, C_BVE_ROUNDS = 8,
             C_SS_ROUNDS = 9, C_NCL = 10, C_NLIT = 11, C_ESTACK = 12, C_KEPT = 13, C_KEPT_LITS = 14, C_WORDS = 32 };

struct K4State {
    u32 *wo, *len, *sig, *wl;           /* clause store: slot c = wl[wo[c] .. wo[c]+len[c]) */
    u8 *removed;
    u32 clause_cap, lit_cap;
    u32 *npos, *nneg, *occ_cnt, *occ_off, *cursor, *occ;   /* per variable; npos..cursor are contiguous (one memset) */
    u8 *eliminated, *selected, *tried;
    u32 *drop_lit;
    u32 *estack; u32 estack_cap;
    u32 *ctl;                           /* C_* words */
    u32 *blk;                           /* scan scratch: one word per block */
    u32 *scan_a, *scan_b;               /* compaction: kept lengths / flags */
    u32 nvars, grow, max_outer;
    int cl_lim;
};

__device__ __forceinline__ u32 warp_incl_scan(u32 v, u32 lane) {
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (u32)o) v += y; }
    return v;
}
/* exclusive scan of in[0..n) into out[0..n], out[n] = total; every thread of the grid calls it */
__device__ void grid_scan(cg::grid_group &g, const u32 *in, u32 *out, u32 n, u32 *blk) {
    __shared__ u32 wsum[32];
    __shared__ u32 carry_s;
    const u32 nb = gridDim.x, bid = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
    const u32 chunk = (n + nb - 1) / nb, lo = bid * chunk, hi = lo + chunk < n ? lo + chunk : n;
    u32 s = 0;
    for (u32 i = lo + t; i < hi; i += blockDim.x) s += in[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) wsum[w] = s;
    __syncthreads();
    if (t == 0) { u32 a = 0; for (u32 k = 0; k < nw; k++) a += wsum[k]; blk[bid] = a; }
    g.sync();
    if (bid == 0 && w == 0) { /* one warp scans the block sums */
        u32 carry = 0;
        for (u32 b0 = 0; b0 < nb; b0 += 32) {
            const u32 v = b0 + lane < nb ? blk[b0 + lane] : 0;
            const u32 inc = warp_incl_scan(v, lane);
            if (b0 + lane < nb) blk[b0 + lane] = carry + inc - v;
            carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) out[n] = carry;
    }
    g.sync();
    if (t == 0) carry_s = blk[bid];
    __syncthreads();
    for (u32 i0 = lo; i0 < hi; i0 += blockDim.x) { /* block-wide scan of the chunk, blockDim.x values per step */
        const u32 i = i0 + t;
        const u32 v = i < hi ? in[i] : 0;
        const u32 inc = warp_incl_scan(v, lane);
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        u32 woff = 0;
        for (u32 k = 0; k < w; k++) woff += wsum[k];
        u32 tot = 0;
        for (u32 k = 0; k < nw; k++) tot += wsum[k];
        const u32 base = carry_s;
        if (i < hi) out[i] = base + woff + inc - v;
        __syncthreads();
        if (t == 0) carry_s = base + tot;
        __syncthreads();
    }
    g.sync();
}

__device__ __forceinline__ bool k4_cand(const K4State &S, u32 v) {
    const u32 p = S.npos[v], n = S.nneg[v];
    return !S.eliminated[v] && !S.tried[v] && p + n > 0 && p + n <= 64 && p * n <= 1024;
}
/* subsumes.rs:27-49: 1 exact, 2 litsign (flip = the literal of D to drop), 0 different */
__device__ __forceinline__ u32 k4_classify(const u32 *wl, u32 cb, u32 cn, u32 db, u32 dn, u32 &flip) {
    u32 ret = 1;
    for (u32 i = 0; i < cn; i++) {
        const u32 l = wl[cb + i];
        bool found = false;
        for (u32 j = 0; j < dn; j++) {
            const u32 d = wl[db + j];
            if (d == l) { found = true; break; }
            if (d == (l ^ 1)) { if (ret == 1) { ret = 2; flip = d; found = true; break; } return 0; }
        }
        if (!found) return 0;
    }
    return ret;
}
/* formula/util.rs:45-77 */
__device__ __forceinline__ bool k4_merge(const K4State &S, u32 v, u32 p, u32 q, u32 *out, u32 &n_out) {
    u32 lp = S.len[p], lq = S.len[q];
    const u32 *longer = S.wl + S.wo[p], *shorter = S.wl + S.wo[q];
    u32 nl = lp, ns = lq;
    if (lp < lq) { longer = S.wl + S.wo[q]; nl = lq; shorter = S.wl + S.wo[p]; ns = lp; }
    n_out = 0;
    for (u32 i = 0; i < ns; i++) {
        const u32 x = shorter[i];
        if ((x >> 1) == v) continue;
        bool add = true;
        for (u32 j = 0; j < nl; j++)
            if ((longer[j] >> 1) == (x >> 1)) { if (longer[j] == (x ^ 1)) return false; add = false; break; }
        if (add) { if (out) out[n_out] = x; n_out++; }
    }
    for (u32 j = 0; j < nl; j++) if ((longer[j] >> 1) != v) { if (out) out[n_out] = longer[j]; n_out++; }
    return true;
}

/* per-variable occurrence lists over the live clauses (histogram -> multi-block scan -> scatter) */
__device__ void k4_rebuild_occ(cg::grid_group &g, const K4State &S) {
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (u32 i = tid; i < 5 * (S.nvars + 2); i += nt) S.npos[i] = 0; /* npos, nneg, occ_cnt, occ_off, cursor */
    g.sync();
    const u32 nc = S.ctl[C_NCL] < S.clause_cap ? S.ctl[C_NCL] : S.clause_cap;
    for (u32 c = tid; c < nc; c += nt) {
        if (S.removed[c]) continue;
        for (u32 k = 0; k < S.len[c]; k++) {
            const u32 l = S.wl[S.wo[c] + k];
            atomicAdd((l & 1) ? &S.nneg[l >> 1] : &S.npos[l >> 1], 1u);
            atomicAdd(&S.occ_cnt[l >> 1], 1u);
        }
    }
    g.sync();
    grid_scan(g, S.occ_cnt, S.occ_off, S.nvars, S.blk);
    for (u32 c = tid; c < nc; c += nt) {
        if (S.removed[c]) continue;
        for (u32 k = 0; k < S.len[c]; k++) {
            const u32 v = S.wl[S.wo[c] + k] >> 1;
            S.occ[S.occ_off[v] + atomicAdd(&S.cursor[v], 1u)] = c;
        }
    }
    g.sync();
}

__global__ void __launch_bounds__(256) k4_preprocess(K4State S) {
    cg::grid_group g = cg::this_grid();
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const u32 lane = threadIdx.x & 31, gw = tid >> 5, nw = nt >> 5;
    const u32 lt = (1u << lane) - 1;
    __shared__ u32 side[8][64];
    u32 *sd = side[(threadIdx.x >> 5) & 7];
    for (u32 outer = 0; outer < S.max_outer; outer++) {
        bool any = false;
        k4_rebuild_occ(g, S);
        /* ---- subsumption + self-subsuming strengthening to a fixpoint; stale occurrence entries only cost a failed test */
        for (u32 r = 0; r < 64; r++) {
            if (tid == 0) S.ctl[C_CHANGED] = 0;
            g.sync();
            const u32 nc = S.ctl[C_NCL] < S.clause_cap ? S.ctl[C_NCL] : S.clause_cap;
            for (u32 c = gw; c < nc; c += nw) { /* warp per clause C */
                if (S.removed[c]) continue;
                const u32 b = S.wo[c], lc = S.len[c];
                if (lc == 0) continue;
                u32 best = S.wl[b] >> 1, bestn = S.occ_off[best + 1] - S.occ_off[best];
                for (u32 k = 1; k < lc; k++) { /* simplify.rs:459-476: the variable with the shortest occurrence list */
                    const u32 v = S.wl[b + k] >> 1, n = S.occ_off[v + 1] - S.occ_off[v];
                    if (n < bestn) { best = v; bestn = n; }
                }
                const u32 sc = S.sig[c];
                for (u32 i = lane; i < bestn; i += 32) {
                    const u32 d = S.occ[S.occ_off[best] + i];
                    if (d == c || S.removed[d] || S.len[d] < lc || (sc & ~S.sig[d]) != 0) continue;
                    u32 flip = 0;
                    const u32 rr = k4_classify(S.wl, b, lc, S.wo[d], S.len[d], flip);
                    if (rr == 1) { if (lc < S.len[d] || c < d) { if (!S.removed[d]) { S.removed[d] = 1; atomicAdd(&S.ctl[C_NSUB], 1u); } S.ctl[C_CHANGED] = 1; } }
                    else if (rr == 2) { if (atomicCAS(&S.drop_lit[d], 0xFFFFFFFFu, flip) == 0xFFFFFFFFu) S.ctl[C_CHANGED] = 1; }
                }
            }
            g.sync();
            for (u32 c = tid; c < nc; c += nt) { /* strengthen: drop the literal, refresh signature / length */
                const u32 dl = S.drop_lit[c];
                if (dl == 0xFFFFFFFFu) continue;
                S.drop_lit[c] = 0xFFFFFFFFu;
                if (S.removed[c]) continue;
                const u32 b = S.wo[c];
                u32 n = S.len[c], j = 0, s = 0;
                for (u32 k = 0; k < n; k++) {
                    const u32 l = S.wl[b + k];
                    if (l == dl) continue;
                    S.wl[b + j++] = l;
                    s |= 1u << ((l >> 1) & 31);
                }
                if (j != n) { S.len[c] = j; S.sig[c] = s; atomicAdd(&S.ctl[C_NSTR], 1u); if (j == 0) S.ctl[C_EMPTY] = 1; }
            }
            g.sync();
            if (tid == 0) S.ctl[C_SS_ROUNDS]++;
            if (!S.ctl[C_CHANGED]) break;
            any = true;
            g.sync();
        }
        if (S.ctl[C_EMPTY] || S.ctl[C_OVERFLOW]) break;
        /* ---- bounded variable elimination: rounds over independent sets (no two selected variables share a clause) */
        u32 elim_at_reset = S.ctl[C_NELIM];
        for (u32 r = 0; r < 4096; r++) {
            if (r > 0 || any) k4_rebuild_occ(g, S); /* the lists of this outer iteration predate its subsumption / strengthening rounds */
            if (tid == 0) S.ctl[C_NSEL] = 0;
            g.sync();
            for (u32 v = tid; v < S.nvars; v += nt) {
                u8 sel = 0;
                if (k4_cand(S, v)) {
                    sel = 1;
                    const u32 cv = S.npos[v] * S.nneg[v];
                    for (u32 i = S.occ_off[v]; i < S.occ_off[v + 1] && sel; i++) {
                        const u32 c = S.occ[i];
                        for (u32 k = 0; k < S.len[c]; k++) {
                            const u32 u = S.wl[S.wo[c] + k] >> 1;
                            if (u == v || !k4_cand(S, u)) continue;
                            const u32 cu = S.npos[u] * S.nneg[u];
                            if (cu < cv || (cu == cv && u < v)) { sel = 0; break; }
                        }
                    }
                }
                S.selected[v] = sel;
                if (sel) atomicAdd(&S.ctl[C_NSEL], 1u);
            }
            g.sync();
            if (S.ctl[C_NSEL] == 0) {
                if (S.ctl[C_NELIM] == elim_at_reset) break;  /* nothing left, and nothing changed since the rejects were recorded */
                elim_at_reset = S.ctl[C_NELIM];
                g.sync();
                for (u32 v = tid; v < S.nvars; v += nt) S.tried[v] = 0;
                g.sync();
                continue;
            }
            for (u32 v = gw; v < S.nvars; v += nw) { /* warp per selected variable, simplify.rs:338-420 */
                if (!S.selected[v]) continue;
                const u32 ob = S.occ_off[v], on = S.occ_off[v + 1] - ob;
                u32 npos = 0, nneg = 0;
                for (u32 i0 = 0; i0 < on; i0 += 32) {
                    const u32 i = i0 + lane;
                    u32 c = 0, sgn = 2;
                    if (i < on) {
                        c = S.occ[ob + i];
                        if (!S.removed[c])
                            for (u32 k = 0; k < S.len[c]; k++) { const u32 l = S.wl[S.wo[c] + k]; if ((l >> 1) == v) { sgn = l & 1; break; } }
                    }
                    const u32 pm = __ballot_sync(0xFFFFFFFFu, sgn == 0), nm = __ballot_sync(0xFFFFFFFFu, sgn == 1);
                    if (sgn == 0) sd[npos + __popc(pm & lt)] = c;
                    if (sgn == 1) sd[63 - (nneg + __popc(nm & lt))] = c;
                    npos += __popc(pm); nneg += __popc(nm);
                }
                __syncwarp();
                const u32 npairs = npos * nneg;
                u32 cnt = 0, sumlen = 0, maxlen = 0;
                for (u32 t0 = 0; t0 < npairs; t0 += 32) {
                    const u32 t = t0 + lane;
                    u32 n_out = 0;
                    bool ok = false;
                    if (t < npairs) ok = k4_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], nullptr, n_out);
                    cnt += __popc(__ballot_sync(0xFFFFFFFFu, ok));
                    u32 l = ok ? n_out : 0;
                    for (int o = 16; o > 0; o >>= 1) { const u32 y = __shfl_xor_sync(0xFFFFFFFFu, l, o); l = l > y ? l : y; }
                    maxlen = maxlen > l ? maxlen : l;
                    u32 s = ok ? n_out : 0;
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                    sumlen += s;
                }
                if (cnt > npos + nneg + S.grow || (S.cl_lim != -1 && (int)maxlen > S.cl_lim)) { /* simplify.rs:368-383 */
                    if (lane == 0) S.tried[v] = 1;
                    __syncwarp();
                    continue;
                }
                const bool save_neg = npos > nneg;
                const u32 ks = save_neg ? nneg : npos;
                u32 rec = 5;
                for (u32 i = 0; i < ks; i++) rec += 1 + S.len[save_neg ? sd[63 - i] : sd[i]];
                u32 cid0 = 0, l0 = 0, e0 = 0, okres = 1;
                if (lane == 0) {
                    cid0 = atomicAdd(&S.ctl[C_NCL], cnt);
                    l0 = atomicAdd(&S.ctl[C_NLIT], sumlen);
                    e0 = atomicAdd(&S.ctl[C_ESTACK], rec);
                    if (cid0 + cnt > S.clause_cap || l0 + sumlen > S.lit_cap || e0 + rec > S.estack_cap) { okres = 0; S.ctl[C_OVERFLOW] = 1; }
                }
                cid0 = __shfl_sync(0xFFFFFFFFu, cid0, 0); l0 = __shfl_sync(0xFFFFFFFFu, l0, 0);
                e0 = __shfl_sync(0xFFFFFFFFu, e0, 0); okres = __shfl_sync(0xFFFFFFFFu, okres, 0);
                if (!okres) { __syncwarp(); continue; } /* the reserved slots stay marked removed */
                u32 rr = 0, ro = 0;
                for (u32 t0 = 0; t0 < npairs; t0 += 32) {
                    const u32 t = t0 + lane;
                    u32 n_out = 0;
                    bool ok = false;
                    if (t < npairs) ok = k4_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], nullptr, n_out);
                    const u32 okm = __ballot_sync(0xFFFFFFFFu, ok);
                    const u32 incl = warp_incl_scan(ok ? n_out : 0, lane);
                    const u32 tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    if (ok) {
                        const u32 slot = cid0 + rr + __popc(okm & lt), off = l0 + ro + incl - n_out;
                        u32 m = 0;
                        k4_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], S.wl + off, m);
                        u32 sg = 0;
                        for (u32 k = 0; k < m; k++) sg |= 1u << ((S.wl[off + k] >> 1) & 31);
                        S.wo[slot] = off; S.len[slot] = m; S.sig[slot] = sg; S.drop_lit[slot] = 0xFFFFFFFFu; S.removed[slot] = 0;
                        if (m == 0) S.ctl[C_EMPTY] = 1;
                    }
                    rr += __popc(okm); ro += tot;
                }
                if (lane == 0) { /* eliminated-clause record: [var, unit lit, round, k, (len, lits...) * k, record words] of the smaller side */
                    u32 *e = S.estack + e0;
                    e[0] = v; e[1] = save_neg ? 2 * v : 2 * v + 1; e[2] = outer * 4096 + r; e[3] = ks;
                    u32 w = 4;
                    for (u32 i = 0; i < ks; i++) {
                        const u32 c = save_neg ? sd[63 - i] : sd[i];
                        const u32 n = S.len[c];
                        e[w++] = n;
                        const u32 first = w;
                        for (u32 k = 0; k < n; k++) { e[w] = S.wl[S.wo[c] + k]; if ((e[w] >> 1) == v) { const u32 tmp = e[first]; e[first] = e[w]; e[w] = tmp; } w++; }
                    }
                    e[w] = w + 1;   /* back link: the stack is walked from its top */
                    S.eliminated[v] = 1;
                    atomicAdd(&S.ctl[C_NELIM], 1u);
                }
                __syncwarp();
                for (u32 i = lane; i < on; i += 32) S.removed[S.occ[ob + i]] = 1;
                __syncwarp();
            }
            g.sync();
            if (tid == 0) S.ctl[C_BVE_ROUNDS]++;
            any = true;
            if (S.ctl[C_OVERFLOW] || S.ctl[C_EMPTY]) break;
        }
        g.sync();
        if (tid == 0) S.ctl[C_OUTER] = outer + 1;
        if (!any || S.ctl[C_OVERFLOW] || S.ctl[C_EMPTY]) break;
    }
    /* ---- compaction offsets (prefix sums of kept lengths / kept flags); the scatter is the second launch */
    g.sync();
    const u32 nc = S.ctl[C_NCL] < S.clause_cap ? S.ctl[C_NCL] : S.clause_cap;
    for (u32 c = tid; c < nc; c += nt) { S.scan_a[c] = S.removed[c] ? 0 : S.len[c]; S.scan_b[c] = S.removed[c] ? 0 : 1; }
    g.sync();
    grid_scan(g, S.scan_a, S.scan_a + S.clause_cap + 2, nc, S.blk);   /* new literal offsets */
    grid_scan(g, S.scan_b, S.scan_b + S.clause_cap + 2, nc, S.blk);   /* new clause indices   */
    if (tid == 0) { S.ctl[C_KEPT] = (S.scan_b + S.clause_cap + 2)[nc]; S.ctl[C_KEPT_LITS] = (S.scan_a + S.clause_cap + 2)[nc]; }
}

__global__ void k4_compact(K4State S, u32 *out_coff, u32 *out_lits) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 nc = S.ctl[C_NCL] < S.clause_cap ? S.ctl[C_NCL] : S.clause_cap;
    if (c >= nc || S.removed[c]) return;
    const u32 o = (S.scan_a + S.clause_cap + 2)[c];
    out_coff[(S.scan_b + S.clause_cap + 2)[c]] = o;
    for (u32 k = 0; k < S.len[c]; k++) out_lits[o + k] = S.wl[S.wo[c] + k];
}

/* elim_clauses.rs:43-63 on the device: walk the records from the top of the stack; model bytes 0 false / 1 true / 2 undefined.
 * Per record the reference's stack holds the saved clauses and then the unit, so walking down the unit comes first: the
 * eliminated variable takes the unit's value whatever the reduced formula's model said about it (it no longer occurs
 * there), and every saved clause that is then unsatisfied flips it to its own literal. */
__global__ void k4_extend_model(const u32 *estack, u32 n_words, u8 *model, u32 n_vars) {
    /* a variable the reduced formula no longer mentions is free: it takes `false` BEFORE the walk (the reference's search
     * assigns every variable), so that a saved clause is judged on the values the final model has */
    for (u32 v = threadIdx.x; v < n_vars; v += blockDim.x) if (model[v] == 2) model[v] = 0;
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (u32 top = n_words; top > 0;) {
        const u32 total = estack[top - 1];
        if (total < 5 || total > top) return;   /* malformed stack: leave the model as it is */
        const u32 p = top - total;
        top = p;
        const u32 unit = estack[p + 1], ks = estack[p + 3];
        { const u32 v = unit >> 1; if (v < n_vars) model[v] = (unit & 1) ? 0 : 1; }
        u32 starts[64];
        u32 q = p + 4;
        for (u32 i = 0; i < ks && i < 64; i++) { starts[i] = q; q += 1 + estack[q]; }
        for (u32 i = ks < 64 ? ks : 64; i-- > 0;) {   /* the saved clauses of this variable, last pushed first */
            const u32 s = starts[i], n = estack[s];
            bool sat = false;
            for (u32 k = 0; k < n; k++) {
                const u32 l = estack[s + 1 + k], v = l >> 1;
                const u8 m = v < n_vars ? model[v] : 2;
                if (m != 2 && (m != 0) != ((l & 1) != 0)) { sat = true; break; }
            }
            if (!sat) { const u32 l = estack[s + 1]; if ((l >> 1) < n_vars) model[l >> 1] = (l & 1) ? 0 : 1; }
        }
    }
}

extern "C" size_t b200sat_k4_words(size_t nc, size_t nvars, size_t nlits) {
    const size_t NC = 2 * nc + 1024, NL = 3 * nlits + 4096;
    return 3 * NC + NL + (NC + 3) / 4 + 6 * (nvars + 2) + NL + 3 * (nvars + 8) / 4 + 4 + NC /* drop */ + NL /* estack */ +
           4 * (NC + 4) /* scans */ + 4096 /* blk */ + C_WORDS + 64;
}

__global__ void k4_init(K4State S, const u32 *coff, const u32 *lits, u32 cb, u32 nc, u32 base) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.clause_cap) return;
    S.drop_lit[c] = 0xFFFFFFFFu;
    if (c >= nc) { S.removed[c] = 1; return; }
    const u32 b = coff[cb + c], e = coff[cb + c + 1];
    u32 s = 0;
    S.wo[c] = b - base;
    for (u32 k = b; k < e; k++) { const u32 l = lits[k]; S.wl[k - base] = l; s |= 1u << ((l >> 1) & 31); }
    S.sig[c] = s; S.len[c] = e - b; S.removed[c] = 0;
}

/* d_counts (host copy): C_* words.  Returns 0, -2 on a CUDA error, -3 when the device cannot launch cooperatively. */
extern "C" int b200sat_k4_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base, u32 *mem,
                                 u32 grow, int cl_lim, u32 max_outer, u32 *d_out_coff, u32 *d_out_lits, u32 **d_estack_out,
                                 u32 *h_counts /* C_WORDS pinned words */, int num_sms, cudaStream_t stream) {
    const u32 NC = 2 * nc + 1024, NL = 3 * nlits + 4096;
    K4State S;
    u32 *p = mem;
    S.wo = p; p += NC; S.len = p; p += NC; S.sig = p; p += NC; S.wl = p; p += NL;
    S.removed = (u8 *)p; p += (NC + 3) / 4;
    S.npos = p; p += nvars + 2; S.nneg = p; p += nvars + 2; S.occ_cnt = p; p += nvars + 2; S.occ_off = p; p += nvars + 2;
    S.cursor = p; p += nvars + 2; p += nvars + 2; S.occ = p; p += NL;
    S.eliminated = (u8 *)p; S.selected = S.eliminated + nvars + 2; S.tried = S.selected + nvars + 2; p += 3 * (nvars + 8) / 4 + 4;
    S.drop_lit = p; p += NC;
    S.estack = p; p += NL;
    S.scan_a = p; p += 2 * (NC + 4); S.scan_b = p; p += 2 * (NC + 4);
    S.blk = p; p += 4096;
    S.ctl = p; p += C_WORDS;
    S.clause_cap = NC; S.lit_cap = NL; S.estack_cap = NL; S.nvars = nvars; S.grow = grow; S.cl_lim = cl_lim; S.max_outer = max_outer;
    int coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return -3;
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k4_preprocess, 256, 0);
    if (bps < 1) return -3;
    if (bps > 4) bps = 4;
    u32 grid = (u32)(num_sms * bps);
    if (grid > 4096) grid = 4096;
    cudaMemsetAsync(S.ctl, 0, C_WORDS * 4, stream);
    cudaMemsetAsync(S.eliminated, 0, 3 * (size_t)(nvars + 2), stream);
    k4_init<<<(NC + 255) / 256, 256, 0, stream>>>(S, d_coff, d_lits, cb, nc, lit_base);
    cudaMemcpyAsync(S.ctl + C_NCL, &nc, 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(S.ctl + C_NLIT, &nlits, 4, cudaMemcpyHostToDevice, stream);
    void *args[] = {(void *)&S};
    if (cudaLaunchCooperativeKernel((const void *)k4_preprocess, dim3(grid), dim3(256), args, 0, stream) != cudaSuccess) return -2;
    k4_compact<<<(NC + 255) / 256, 256, 0, stream>>>(S, d_out_coff, d_out_lits);
    cudaMemcpyAsync(h_counts, S.ctl, C_WORDS * 4, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
    *d_estack_out = S.estack;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int b200sat_k4_extend_model_launch(const u32 *d_estack, u32 n_words, u8 *d_model, u32 n_vars, cudaStream_t stream) {
    k4_extend_model<<<1, 256, 0, stream>>>(d_estack, n_words, d_model, n_vars);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
