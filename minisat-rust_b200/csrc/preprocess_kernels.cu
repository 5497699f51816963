/*
 * preprocess_kernels.cu -- K4: throughput-oriented, multi-CTA preprocessing of ONE large instance.
 *
 * Round-1 scope: backward subsumption over occurrence lists (the `subsumes()` test of the reference,
 * src/sat/minisat/search/simplify/subsumes.rs:10-49, and the signature of src/sat/formula/util.rs:5-11), i.e.
 *   pp_sig_len      thread per clause: 32-bit signature (1 << (var & 31)) and length
 *   pp_occ_count    thread per clause: per-VARIABLE occurrence histogram (OccLists are per var, elim_queue.rs:85-151)
 *   pp_scan         exclusive prefix sum (occurrence-list offsets / compacted clause offsets)
 *   pp_occ_fill     thread per clause: scatter clause ids into the occurrence lists
 *   pp_occ_sort     warp per variable: segmented sort of every occurrence list by clause id (canonical order)
 *   pp_subsume      warp per clause C: candidates = occurrence list of C's least-occurring variable; a lane
 *                   tests one candidate D: size filter, signature filter, then the literal subset test
 *   pp_compact      scatter of the surviving clauses to their prefix-sum offsets (new CSR)
 * A clause D is removed iff some other clause C is a subset of D and (|C| < |D| or C has the smaller index), so
 * the result does not depend on scheduling order.  This is NOT the reference's sequential job order (no
 * strengthening, no BVE here): its parity is the S* definition of SURVEY.md 8a -- the reduced formula is
 * equivalent to the input, every removed clause has a surviving subset clause, verdicts agree -- and the
 * clause counts are reported beside the oracle's.  The trace-exact simplifier is csrc/simp_warp.cuh.
 */
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

__global__ void pp_sig_len(const u32 *coff, const u32 *lits, u32 cb, u32 nc, u32 *sig, u32 *len, u8 *removed) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    u32 b = coff[cb + c], e = coff[cb + c + 1], s = 0;
    for (u32 k = b; k < e; k++) s |= 1u << ((lits[k] >> 1) & 31);
    sig[c] = s;
    len[c] = e - b;
    removed[c] = 0;
}

__global__ void pp_occ_count(const u32 *coff, const u32 *lits, u32 cb, u32 nc, u32 *occ_cnt) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    for (u32 k = coff[cb + c]; k < coff[cb + c + 1]; k++) atomicAdd(&occ_cnt[lits[k] >> 1], 1u);
}

/* single-block exclusive scan of n values (n up to a few million): out[i] = sum(in[0..i)), out[n] = total */
__global__ void pp_scan(const u32 *in, u32 *out, u32 n) {
    __shared__ u32 part[1024];
    __shared__ u32 carry;
    const u32 t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        u32 i = base + t;
        u32 v = i < n ? in[i] : 0;
        part[t] = v;
        __syncthreads();
        for (u32 o = 1; o < 1024; o <<= 1) {
            u32 y = t >= o ? part[t - o] : 0;
            __syncthreads();
            part[t] += y;
            __syncthreads();
        }
        if (i < n) out[i] = carry + part[t] - v;
        __syncthreads();
        if (t == 0) carry += part[1023];
        __syncthreads();
    }
    if (t == 0) out[n] = carry;
}

__global__ void pp_occ_fill(const u32 *coff, const u32 *lits, u32 cb, u32 nc, const u32 *occ_off, u32 *cursor, u32 *occ) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    for (u32 k = coff[cb + c]; k < coff[cb + c + 1]; k++) {
        u32 v = lits[k] >> 1;
        occ[occ_off[v] + atomicAdd(&cursor[v], 1u)] = c;
    }
}

/* segmented sort: one warp per occurrence list, odd-even transposition over the segment */
__global__ void pp_occ_sort(const u32 *occ_off, u32 *occ, u32 nvars) {
    const u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= nvars) return;
    u32 *seg = occ + occ_off[v];
    const u32 n = occ_off[v + 1] - occ_off[v];
    if (n > 2048) return; /* very long lists stay in scatter order (order only affects traversal, not the result) */
    for (u32 pass = 0; pass < n; pass++) {
        for (u32 i = (pass & 1) + 2 * lane; i + 1 < n; i += 64) {
            u32 a = seg[i], b = seg[i + 1];
            if (a > b) { seg[i] = b; seg[i + 1] = a; }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ bool pp_subset(const u32 *lits, u32 cb_, u32 ce_, u32 db, u32 de) {
    for (u32 i = cb_; i < ce_; i++) { /* subsumes.rs:27-47 without the LitSign case */
        u32 l = lits[i];
        bool found = false;
        for (u32 j = db; j < de; j++) if (lits[j] == l) { found = true; break; }
        if (!found) return false;
    }
    return true;
}

__global__ void pp_subsume(const u32 *coff, const u32 *lits, u32 cb, u32 nc, const u32 *sig, const u32 *len, const u32 *occ_off,
                           const u32 *occ, u8 *removed, u32 *n_tests) {
    const u32 c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= nc) return;
    const u32 b = coff[cb + c], e = coff[cb + c + 1];
    if (e == b) return;
    u32 best = lits[b] >> 1, bestn = occ_off[best + 1] - occ_off[best];
    for (u32 k = b + 1; k < e; k++) { /* simplify.rs:459-476: the variable with the shortest occurrence list */
        u32 v = lits[k] >> 1, n = occ_off[v + 1] - occ_off[v];
        if (n < bestn) { best = v; bestn = n; }
    }
    const u32 sc = sig[c], lc = len[c];
    u32 tests = 0;
    for (u32 i = lane; i < bestn; i += 32) {
        const u32 d = occ[occ_off[best] + i];
        if (d == c || len[d] < lc || (sc & ~sig[d]) != 0) continue; /* size + signature filter, subsumes.rs:15-21 */
        tests++;
        if (pp_subset(lits, b, e, coff[cb + d], coff[cb + d + 1]) && (lc < len[d] || c < d)) removed[d] = 1;
    }
    if (tests) atomicAdd(n_tests, tests);
}

__global__ void pp_kept_len(const u32 *len, const u8 *removed, u32 nc, u32 *kept_len, u32 *kept_flag) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    kept_len[c] = removed[c] ? 0 : len[c];
    kept_flag[c] = removed[c] ? 0 : 1;
}

__global__ void pp_compact(const u32 *coff, const u32 *lits, u32 cb, u32 nc, const u8 *removed, const u32 *new_idx, const u32 *new_off,
                           u32 *out_coff, u32 *out_lits) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc || removed[c]) return;
    const u32 base = coff[cb + c], n = coff[cb + c + 1] - base, o = new_off[c];
    out_coff[new_idx[c]] = o;
    for (u32 k = 0; k < n; k++) out_lits[o + k] = lits[base + k];
}

/* host launcher: everything on `stream`; scratch = device buffer of at least pp_scratch_words(nc, nvars, nlits) words */
extern "C" size_t b200sat_pp_scratch_words(size_t nc, size_t nvars, size_t nlits) {
    return 8 * (nc + 2) + 4 * (nvars + 2) + 2 * (nlits + 2) + 64;
}

extern "C" int b200sat_pp_subsume_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 *scratch,
                                         u8 *d_removed, u32 *d_out_coff, u32 *d_out_lits, u32 *d_counts /* [0] kept clauses, [1] kept lits, [2] tests */,
                                         cudaStream_t stream) {
    u32 *sig = scratch, *len = sig + nc + 2, *kept_len = len + nc + 2, *kept_flag = kept_len + nc + 2, *new_off = kept_flag + nc + 2,
        *new_idx = new_off + nc + 2, *occ_cnt = new_idx + nc + 2, *occ_off = occ_cnt + nvars + 2, *cursor = occ_off + nvars + 2,
        *occ = cursor + nvars + 2;
    const u32 T = 256, gc = (nc + T - 1) / T;
    cudaMemsetAsync(occ_cnt, 0, (size_t)(nvars + 2) * 4, stream);
    cudaMemsetAsync(cursor, 0, (size_t)(nvars + 2) * 4, stream);
    cudaMemsetAsync(d_counts, 0, 16, stream);
    pp_sig_len<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, sig, len, d_removed);
    pp_occ_count<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, occ_cnt);
    pp_scan<<<1, 1024, 0, stream>>>(occ_cnt, occ_off, nvars);
    pp_occ_fill<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, occ_off, cursor, occ);
    pp_occ_sort<<<(nvars * 32 + T - 1) / T, T, 0, stream>>>(occ_off, occ, nvars);
    pp_subsume<<<(u32)(((u64)nc * 32 + T - 1) / T), T, 0, stream>>>(d_coff, d_lits, cb, nc, sig, len, occ_off, occ, d_removed, d_counts + 2);
    pp_kept_len<<<gc, T, 0, stream>>>(len, d_removed, nc, kept_len, kept_flag);
    pp_scan<<<1, 1024, 0, stream>>>(kept_len, new_off, nc);
    pp_scan<<<1, 1024, 0, stream>>>(kept_flag, new_idx, nc);
    pp_compact<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, d_removed, new_idx, new_off, d_out_coff, d_out_lits);
    cudaMemcpyAsync(d_counts + 0, new_idx + nc, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 1, new_off + nc, 4, cudaMemcpyDeviceToDevice, stream);
    (void)nlits;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}


/* ------------------------------------------------------------------ subsumption + self-subsuming strengthening to fixpoint
 * (simplify.rs:423-523 semantics, NOT its job order).  Works on a private copy of the instance's literals:
 * clause c = wl[wo[c] .. wo[c] + len[c]).  One round = pp2_scan_pairs (warp per clause C: for every candidate D of
 * C's least-occurring variable classify subsumes(C, D): Exact -> removed[D], LitSign(l) -> D should lose !l,
 * one literal per clause and round through atomicCAS) + pp2_apply (drop the literal, refresh signature/length).
 * Every step is implied by the clause set it was computed on, so the result is equivalent to the input.
 * Occurrence lists are not rebuilt: stale entries only cost a failed subset test. */
__global__ void pp2_copy(const u32 *coff, const u32 *lits, u32 cb, u32 nc, u32 base, u32 *wo, u32 *wl) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    u32 b = coff[cb + c], e = coff[cb + c + 1];
    wo[c] = b - base;
    for (u32 k = b; k < e; k++) wl[k - base] = lits[k];
}

__device__ __forceinline__ u32 pp2_classify(const u32 *wl, u32 cb_, u32 cn, u32 db, u32 dn, u32 &flip) {
    u32 ret = 1; /* 1 exact, 2 litsign, 0 different (subsumes.rs:27-49) */
    for (u32 i = 0; i < cn; i++) {
        const u32 l = wl[cb_ + i];
        bool found = false;
        for (u32 j = 0; j < dn; j++) {
            const u32 d = wl[db + j];
            if (d == l) { found = true; break; }
            if (d == (l ^ 1)) {
                if (ret == 1) { ret = 2; flip = d; found = true; break; }
                return 0;
            }
        }
        if (!found) return 0;
    }
    return ret;
}

__global__ void pp2_scan_pairs(const u32 *wo, const u32 *wl, u32 nc, const u32 *sig, const u32 *len, const u32 *occ_off, const u32 *occ,
                               u8 *removed, u32 *drop_lit, u32 *changed) {
    const u32 c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= nc || removed[c]) return;
    const u32 b = wo[c], lc = len[c];
    if (lc == 0) return;
    u32 best = wl[b] >> 1, bestn = occ_off[best + 1] - occ_off[best];
    for (u32 k = 1; k < lc; k++) {
        u32 v = wl[b + k] >> 1, n = occ_off[v + 1] - occ_off[v];
        if (n < bestn) { best = v; bestn = n; }
    }
    const u32 sc = sig[c];
    for (u32 i = lane; i < bestn; i += 32) {
        const u32 d = occ[occ_off[best] + i];
        if (d == c || removed[d] || len[d] < lc || (sc & ~sig[d]) != 0) continue;
        u32 flip = 0;
        const u32 r = pp2_classify(wl, b, lc, wo[d], len[d], flip);
        if (r == 1) { if (lc < len[d] || c < d) { removed[d] = 1; *changed = 1; } }
        else if (r == 2) { if (atomicCAS(&drop_lit[d], 0xFFFFFFFFu, flip) == 0xFFFFFFFFu) *changed = 1; }
    }
}

__global__ void pp2_apply(const u32 *wo, u32 *wl, u32 nc, u32 *sig, u32 *len, const u8 *removed, u32 *drop_lit, u32 *n_strengthened,
                          u32 *empty_clause) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const u32 dl = drop_lit[c];
    if (dl == 0xFFFFFFFFu) return;
    drop_lit[c] = 0xFFFFFFFFu;
    if (removed[c]) return;
    const u32 b = wo[c];
    u32 n = len[c], j = 0, s = 0;
    for (u32 k = 0; k < n; k++) {
        const u32 l = wl[b + k];
        if (l == dl) continue;
        wl[b + j++] = l; /* order kept, like strengthen (simplify.rs:567-591) */
        s |= 1u << ((l >> 1) & 31);
    }
    if (j != n) { len[c] = j; sig[c] = s; atomicAdd(n_strengthened, 1u); if (j == 0) *empty_clause = 1; }
}

__global__ void pp2_compact(const u32 *wo, const u32 *wl, u32 nc, const u32 *len, const u8 *removed, const u32 *new_idx, const u32 *new_off,
                            u32 *out_coff, u32 *out_lits) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc || removed[c]) return;
    const u32 o = new_off[c];
    out_coff[new_idx[c]] = o;
    for (u32 k = 0; k < len[c]; k++) out_lits[o + k] = wl[wo[c] + k];
}

/* d_counts: [0] kept clauses, [1] kept literals, [2] rounds, [3] strengthened literals, [4] empty clause derived */
extern "C" int b200sat_pp_simplify_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base,
                                          u32 *scratch, u32 *d_work /* nc + nlits + nc + 8 words */, u8 *d_removed, u32 *d_out_coff,
                                          u32 *d_out_lits, u32 *d_counts, u32 *h_flag /* pinned */, u32 max_rounds, cudaStream_t stream) {
    u32 *sig = scratch, *len = sig + nc + 2, *kept_len = len + nc + 2, *kept_flag = kept_len + nc + 2, *new_off = kept_flag + nc + 2,
        *new_idx = new_off + nc + 2, *occ_cnt = new_idx + nc + 2, *occ_off = occ_cnt + nvars + 2, *cursor = occ_off + nvars + 2,
        *occ = cursor + nvars + 2;
    u32 *wo = d_work, *wl = wo + nc + 2, *drop = wl + nlits + 2;
    const u32 T = 256, gc = (nc + T - 1) / T;
    cudaMemsetAsync(occ_cnt, 0, (size_t)(nvars + 2) * 4, stream);
    cudaMemsetAsync(cursor, 0, (size_t)(nvars + 2) * 4, stream);
    cudaMemsetAsync(d_counts, 0, 32, stream);
    cudaMemsetAsync(drop, 0xFF, (size_t)(nc + 2) * 4, stream);
    pp_sig_len<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, sig, len, d_removed);
    pp_occ_count<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, occ_cnt);
    pp_scan<<<1, 1024, 0, stream>>>(occ_cnt, occ_off, nvars);
    pp_occ_fill<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, occ_off, cursor, occ);
    pp_occ_sort<<<(nvars * 32 + T - 1) / T, T, 0, stream>>>(occ_off, occ, nvars);
    pp2_copy<<<gc, T, 0, stream>>>(d_coff, d_lits, cb, nc, lit_base, wo, wl);
    u32 rounds = 0;
    for (; rounds < max_rounds; rounds++) {
        cudaMemsetAsync(d_counts + 5, 0, 4, stream);
        pp2_scan_pairs<<<(u32)(((u64)nc * 32 + T - 1) / T), T, 0, stream>>>(wo, wl, nc, sig, len, occ_off, occ, d_removed, drop, d_counts + 5);
        pp2_apply<<<gc, T, 0, stream>>>(wo, wl, nc, sig, len, d_removed, drop, d_counts + 3, d_counts + 4);
        cudaMemcpyAsync(h_flag, d_counts + 5, 4, cudaMemcpyDeviceToHost, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
        if (!*h_flag) { rounds++; break; }
    }
    pp_kept_len<<<gc, T, 0, stream>>>(len, d_removed, nc, kept_len, kept_flag);
    pp_scan<<<1, 1024, 0, stream>>>(kept_len, new_off, nc);
    pp_scan<<<1, 1024, 0, stream>>>(kept_flag, new_idx, nc);
    pp2_compact<<<gc, T, 0, stream>>>(wo, wl, nc, len, d_removed, new_idx, new_off, d_out_coff, d_out_lits);
    cudaMemcpyAsync(d_counts + 0, new_idx + nc, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 1, new_off + nc, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 2, &rounds, 4, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}


/* ------------------------------------------------------------------ bounded variable elimination, multi-CTA (simplify.rs:338-420
 * semantics, NOT its order).  Rounds of: rebuild per-variable occurrence lists over the live clauses ->
 * bve_select (a candidate variable is selected iff it has the smallest (cost = n_occ+ * n_occ-, id) among the
 * candidates it shares a clause with, so selected variables share no clause) -> bve_eliminate (warp per selected
 * variable: all pos x neg resolvents by formula/util.rs:45-77 `merge`, one pair per lane; eliminated iff the number
 * of non-tautological resolvents <= n_occ + grow and none is longer than cl_lim; resolvents appended at
 * prefix-sum offsets, originals marked removed, the smaller side saved on the eliminated-clause stack with the
 * variable's literal first, then the default unit -- elim_clauses.rs:20-41).  Model extension walks that stack
 * backwards (elim_clauses.rs:43-63). */
struct BveState {
    u32 *wo, *len, *sig, *wl;           /* clause store: slot c = wl[wo[c] .. wo[c]+len[c]) */
    u8 *removed;
    u32 *n_clauses, *n_lits;            /* cursors (device) */
    u32 clause_cap, lit_cap;
    u32 *npos, *nneg, *occ_cnt, *occ_off, *cursor, *occ;
    u8 *eliminated, *selected, *tried;
    u32 *n_selected;
    u32 *estack, *estack_n;
    u32 estack_cap;
    u32 *n_elim, *overflow;
    u32 nvars;
};

__global__ void bve_count(BveState S) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *S.n_clauses || S.removed[c]) return;
    for (u32 k = 0; k < S.len[c]; k++) {
        u32 l = S.wl[S.wo[c] + k];
        atomicAdd((l & 1) ? &S.nneg[l >> 1] : &S.npos[l >> 1], 1u);
        atomicAdd(&S.occ_cnt[l >> 1], 1u);
    }
}
__global__ void bve_fill(BveState S) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *S.n_clauses || S.removed[c]) return;
    for (u32 k = 0; k < S.len[c]; k++) {
        u32 v = S.wl[S.wo[c] + k] >> 1;
        S.occ[S.occ_off[v] + atomicAdd(&S.cursor[v], 1u)] = c;
    }
}
__device__ __forceinline__ bool bve_cand(const BveState &S, u32 v) {
    const u32 p = S.npos[v], n = S.nneg[v];
    return !S.eliminated[v] && !S.tried[v] && p + n > 0 && p + n <= 64 && p * n <= 1024;
}
__global__ void bve_select(BveState S) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S.nvars) return;
    u8 sel = 0;
    if (bve_cand(S, v)) {
        sel = 1;
        const u32 cv = S.npos[v] * S.nneg[v];
        for (u32 i = S.occ_off[v]; i < S.occ_off[v + 1] && sel; i++) {
            const u32 c = S.occ[i];
            for (u32 k = 0; k < S.len[c]; k++) {
                const u32 u = S.wl[S.wo[c] + k] >> 1;
                if (u == v || !bve_cand(S, u)) continue;
                const u32 cu = S.npos[u] * S.nneg[u];
                if (cu < cv || (cu == cv && u < v)) { sel = 0; break; }
            }
        }
    }
    S.selected[v] = sel;
    if (sel) atomicAdd(S.n_selected, 1u);
}

/* merge (formula/util.rs:45-77) of clause slots p and q on variable v: returns false for a tautology; out may be NULL */
__device__ __forceinline__ bool bve_merge(const BveState &S, u32 v, u32 p, u32 q, u32 *out, u32 &n_out) {
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

__global__ void bve_eliminate(BveState S, u32 grow, int cl_lim, u32 round) {
    const u32 v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= S.nvars || !S.selected[v]) return;
    __shared__ u32 side[8][64]; /* per warp: pos ids from the front, neg ids from the back */
    u32 *sd = side[(threadIdx.x >> 5) & 7];
    const u32 lt = (1u << lane) - 1;
    const u32 ob = S.occ_off[v], on = S.occ_off[v + 1] - ob;
    u32 npos = 0, nneg = 0;
    for (u32 i0 = 0; i0 < on; i0 += 32) {
        const u32 i = i0 + lane;
        u32 c = 0, sgn = 2;
        if (i < on) {
            c = S.occ[ob + i];
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
        if (t < npairs) ok = bve_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], nullptr, n_out);
        cnt += __popc(__ballot_sync(0xFFFFFFFFu, ok));
        u32 l = ok ? n_out : 0;
        for (int o = 16; o > 0; o >>= 1) { const u32 y = __shfl_xor_sync(0xFFFFFFFFu, l, o); l = l > y ? l : y; }
        maxlen = maxlen > l ? maxlen : l;
        u32 s = ok ? n_out : 0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        sumlen += s;
    }
    if (cnt > npos + nneg + grow || (cl_lim != -1 && (int)maxlen > cl_lim)) { /* simplify.rs:368-383 */
        if (lane == 0) S.tried[v] = 1; /* not again until the neighbourhood changed (the host resets `tried` after progress) */
        return;
    }
    /* eliminated-clause record: [var, unit lit, round, k, (len, lits...) * k] of the smaller side */
    const bool save_neg = npos > nneg;
    const u32 ks = save_neg ? nneg : npos;
    u32 rec = 4;
    for (u32 i = 0; i < ks; i++) rec += 1 + S.len[save_neg ? sd[63 - i] : sd[i]];
    u32 cid0 = 0, l0 = 0, e0 = 0, okres = 1;
    if (lane == 0) {
        cid0 = atomicAdd(S.n_clauses, cnt);
        l0 = atomicAdd(S.n_lits, sumlen);
        e0 = atomicAdd(S.estack_n, rec);
        if (cid0 + cnt > S.clause_cap || l0 + sumlen > S.lit_cap || e0 + rec > S.estack_cap) { okres = 0; *S.overflow = 1; }
    }
    cid0 = __shfl_sync(0xFFFFFFFFu, cid0, 0); l0 = __shfl_sync(0xFFFFFFFFu, l0, 0);
    e0 = __shfl_sync(0xFFFFFFFFu, e0, 0); okres = __shfl_sync(0xFFFFFFFFu, okres, 0);
    if (!okres) return; /* the reserved slots stay marked removed (initialised so by the host) */
    u32 rr = 0, ro = 0;
    for (u32 t0 = 0; t0 < npairs; t0 += 32) {
        const u32 t = t0 + lane;
        u32 n_out = 0;
        bool ok = false;
        if (t < npairs) ok = bve_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], nullptr, n_out);
        const u32 okm = __ballot_sync(0xFFFFFFFFu, ok);
        u32 incl = ok ? n_out : 0;
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += y; }
        const u32 tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (ok) {
            const u32 slot = cid0 + rr + __popc(okm & lt), off = l0 + ro + incl - n_out;
            u32 m = 0;
            bve_merge(S, v, sd[t / nneg], sd[63 - (t % nneg)], S.wl + off, m);
            u32 sg = 0;
            for (u32 k = 0; k < m; k++) sg |= 1u << ((S.wl[off + k] >> 1) & 31);
            S.wo[slot] = off; S.len[slot] = m; S.sig[slot] = sg; S.removed[slot] = 0;
        }
        rr += __popc(okm); ro += tot;
    }
    if (lane == 0) {
        u32 *e = S.estack + e0;
        e[0] = v; e[1] = save_neg ? 2 * v : 2 * v + 1; e[2] = round; e[3] = ks;
        u32 w = 4;
        for (u32 i = 0; i < ks; i++) {
            const u32 c = save_neg ? sd[63 - i] : sd[i];
            const u32 n = S.len[c];
            e[w++] = n;
            u32 first = w;
            for (u32 k = 0; k < n; k++) { e[w] = S.wl[S.wo[c] + k]; if ((e[w] >> 1) == v) { u32 tmp = e[first]; e[first] = e[w]; e[w] = tmp; } w++; }
        }
        S.eliminated[v] = 1;
        atomicAdd(S.n_elim, 1u);
    }
    __syncwarp();
    for (u32 i = lane; i < on; i += 32) S.removed[S.occ[ob + i]] = 1;
}

extern "C" size_t b200sat_bve_words(size_t nc, size_t nvars, size_t nlits) {
    const size_t NC = 2 * nc + 1024, NL = 3 * nlits + 4096;
    return 3 * NC + NL /* wo,len,sig,wl */ + (NC + 3) / 4 /* removed */ + 5 * (nvars + 2) + NL /* occ */ + 3 * (nvars + 8) / 4 + 4 /* flags */ +
           NL /* estack */ + 3 * (NC + 2) /* compaction scratch */ + 64;
}

/* d_counts: [0] kept clauses [1] kept literals [2] eliminated vars [3] rounds [4] estack words [5] overflow */
extern "C" int b200sat_pp_bve_launch(const u32 *d_coff, const u32 *d_lits, u32 cb, u32 nc, u32 nvars, u32 nlits, u32 lit_base, u32 *mem,
                                     u32 grow, int cl_lim, u32 max_rounds, u32 *d_out_coff, u32 *d_out_lits, u32 **d_estack_out,
                                     u32 *d_counts, u32 *h_pin /* 8 pinned words */, cudaStream_t stream) {
    const u32 NC = 2 * nc + 1024, NL = 3 * nlits + 4096;
    BveState S;
    u32 *p = mem;
    S.wo = p; p += NC; S.len = p; p += NC; S.sig = p; p += NC; S.wl = p; p += NL;
    S.removed = (u8 *)p; p += (NC + 3) / 4;
    S.npos = p; p += nvars + 2; S.nneg = p; p += nvars + 2; S.occ_cnt = p; p += nvars + 2; S.occ_off = p; p += nvars + 2;
    S.cursor = p; p += nvars + 2; S.occ = p; p += NL;
    S.eliminated = (u8 *)p; S.selected = S.eliminated + nvars + 2; S.tried = S.selected + nvars + 2; p += 3 * (nvars + 8) / 4 + 4;
    S.estack = p; p += NL;
    u32 *kept_len = p; p += NC + 2; u32 *kept_flag = p; p += NC + 2; u32 *new_off = p; p += NC + 2;
    S.clause_cap = NC; S.lit_cap = NL; S.estack_cap = NL; S.nvars = nvars;
    S.n_clauses = d_counts + 8; S.n_lits = d_counts + 9; S.estack_n = d_counts + 10; S.n_elim = d_counts + 11; S.overflow = d_counts + 12; S.n_selected = d_counts + 13;
    const u32 T = 256;
    cudaMemsetAsync(d_counts, 0, 64, stream);
    cudaMemsetAsync(S.removed, 1, NC, stream);           /* unused slots count as removed */
    cudaMemsetAsync(S.eliminated, 0, 3 * (size_t)(nvars + 2), stream);
    pp_sig_len<<<(nc + T - 1) / T, T, 0, stream>>>(d_coff, d_lits, cb, nc, S.sig, S.len, S.removed);
    pp2_copy<<<(nc + T - 1) / T, T, 0, stream>>>(d_coff, d_lits, cb, nc, lit_base, S.wo, S.wl);
    cudaMemcpyAsync(S.n_clauses, &nc, 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(S.n_lits, &nlits, 4, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
    u32 rounds = 0, ncur = nc, elim_at_reset = 0;
    for (; rounds < max_rounds; rounds++) {
        cudaMemsetAsync(S.npos, 0, (size_t)5 * (nvars + 2) * 4, stream); /* npos, nneg, occ_cnt, occ_off, cursor */
        cudaMemsetAsync(S.n_selected, 0, 4, stream);
        const u32 g = (ncur + T - 1) / T;
        bve_count<<<g, T, 0, stream>>>(S);
        pp_scan<<<1, 1024, 0, stream>>>(S.occ_cnt, S.occ_off, nvars);
        bve_fill<<<g, T, 0, stream>>>(S);
        bve_select<<<(nvars + T - 1) / T, T, 0, stream>>>(S);
        bve_eliminate<<<(u32)(((u64)nvars * 32 + T - 1) / T), T, 0, stream>>>(S, grow, cl_lim, rounds);
        cudaMemcpyAsync(h_pin + 0, S.n_selected, 4, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(h_pin + 1, S.n_elim, 4, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(h_pin + 2, S.n_clauses, 4, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(h_pin + 3, S.overflow, 4, cudaMemcpyDeviceToHost, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
        ncur = h_pin[2] < NC ? h_pin[2] : NC;
        if (h_pin[3]) { rounds++; break; }
        if (h_pin[0] == 0) {                 /* no candidate left: retry the rejected variables if anything changed since */
            if (h_pin[1] == elim_at_reset) { rounds++; break; }
            elim_at_reset = h_pin[1];
            cudaMemsetAsync(S.tried, 0, nvars + 2, stream);
        }
    }
    const u32 g = (ncur + T - 1) / T;
    pp_kept_len<<<g, T, 0, stream>>>(S.len, S.removed, ncur, kept_len, kept_flag);
    pp_scan<<<1, 1024, 0, stream>>>(kept_len, new_off, ncur);
    pp_scan<<<1, 1024, 0, stream>>>(kept_flag, kept_len, ncur); /* kept_len now holds the new clause index */
    pp2_compact<<<g, T, 0, stream>>>(S.wo, S.wl, ncur, S.len, S.removed, kept_len, new_off, d_out_coff, d_out_lits);
    cudaMemcpyAsync(d_counts + 0, kept_len + ncur, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 1, new_off + ncur, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 2, S.n_elim, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 3, &rounds, 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_counts + 4, S.estack_n, 4, cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(d_counts + 5, S.overflow, 4, cudaMemcpyDeviceToDevice, stream);
    cudaStreamSynchronize(stream);
    *d_estack_out = S.estack;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
