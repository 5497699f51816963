/*
 * cdcl_kernels.cuh -- the persistent warp-per-instance kernels (round 2: phase-specialised).
 *
 *   ingest_smem<NV,COUNT,SIMP> / ingest_glob<COUNT,SIMP>   add_clause* + preprocess of every work item; the state of an
 *                                                          item that needs a search is parked in ITS slab and the item
 *                                                          is pushed on the device work ring
 *   search_smem<NV,COUNT,PF>  / search_glob<COUNT,PF>      pops items from the ring, resumes them from their slab, runs
 *                                                          the CDCL search; an item that used up its conflict slice
 *                                                          while others wait is parked again and re-queued (round-robin)
 *
 * The search kernels carry no ingest / simplifier code and (PF = false) no clause-exchange code, so co-resident warps
 * share one hot loop.  Included by b200sat.cu for the declarations only; each kern_*.cu defines B200SAT_KERNEL_TU_*
 * and gets the templates plus its `b200sat_pick_*` accessor, so the translation units compile in parallel.
 */
#ifndef B200SAT_CDCL_KERNELS_CUH
#define B200SAT_CDCL_KERNELS_CUH
#include "cdcl_types.h"

typedef void (*kernel_fn)(const b200sat::KParams);
/* tier_nv: 128 / 256 / 1024 = shared-memory tier, 0 = global tier.  smem_per_warp: dynamic shared memory one warp
 * needs; save_bytes: bytes of it that make up the parked var state (same for every kernel of a tier). */
kernel_fn b200sat_pick_ingest_c0s0(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_ingest_c1s0(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_ingest_c0s1(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_ingest_c1s1(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_search_c0(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_search_c1(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_search_pf(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_search_rescue(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_replay_c0(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);
kernel_fn b200sat_pick_replay_c1(int tier_nv, size_t *smem_per_warp, size_t *save_bytes);

#ifdef B200SAT_KERNEL_TU_NAME
#include <cuda_runtime.h>
#include <stddef.h>
#include "cdcl_warp.cuh"
#if B200SAT_KERNEL_TU_SIMP
#include "simp_warp.cuh"
#endif

#ifndef B200SAT_SEARCH_BLOCKS
#define B200SAT_SEARCH_BLOCKS 24 /* resident one-warp blocks per SM the search kernels are compiled for: 80 registers per
                                  * thread, no spills.  Measured (profiles/r02_ab_d.log): same throughput as 32 blocks at 64
                                  * registers on a saturated GPU, 2 % faster when the n=250 tier (21 blocks by shared memory) runs */
#endif

namespace b200sat {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

/* The block's shared-memory base, held in a register pair.  Left to itself the compiler rebuilds the shared-window
 * address (S2UR SR_CgaCtaId + UMOV + ULEA) in front of every fixed-offset LDS / STS of the var block: 127 static sites,
 * 3.3 % of the executed instructions of the search kernel and an S2UR latency on every scalar access.  The empty asm
 * hides the constant, the assumption keeps the accesses LDS / STS (measured: -7 % run time, profiles/r02_ab_d.log). */
template <class T> __device__ __forceinline__ T *smem_base(unsigned char *raw) {
    T *p = reinterpret_cast<T *>(raw);
#ifndef B200SAT_EMU
    asm volatile("" : "+l"(p));
    __builtin_assume(__isShared(p));
#endif
    return p;
}

template <class WS> __device__ __forceinline__ void bind_item(WS &S, const KParams &P, u32 w) {
    S.slab = P.slabs + (size_t)(w - P.chunk_base) * P.L.slab_bytes;
    S.item_index = w;
    S.res_index = P.work ? P.work[w] : w;
    S.stp = &P.st;
    if (P.portfolio) { /* item w = replica w of instance 0 with its own diversified settings */
        S.res_index = w;
        S.stp = P.replica_settings + w;
    }
}

#if B200SAT_KERNEL_TU_PHASE == 0
/* ---------------------------------------------------------------- ingest: a static atomic queue over the chunk */
template <bool SIMP, class WS> __device__ __forceinline__ void ingest_loop(WS &S, const KParams &P) {
    for (;;) {
        u32 w = 0;
        if (S.lane == 0) w = atomicAdd(P.queue, 1u);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= P.n_work) break;
        w += P.chunk_base;
        bind_item(S, P, w);
        if constexpr (WS::VT_GLOBAL) S.V.bind(S.slab + P.L.off_vars, P.L.nv_cap);
        const u32 inst = P.portfolio ? 0u : S.res_index;
        bool runnable;
#if B200SAT_KERNEL_TU_SIMP
        if (SIMP) runnable = ingest_simp(S, inst); else
#endif
        runnable = S.ingest_core(inst);
        if (runnable && !(P.flags & B200SAT_F_NO_SOLVE)) S.ring_push(w); else S.count_done();
    }
}
template <int NV, bool COUNT, bool SIMP> __global__ void __launch_bounds__(64, 8) ingest_smem(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSolver<VarsS<NV, SIMP>, COUNT> S(P, lane);
    S.V.s = smem_base<SmemVars<NV, SIMP>>(smem_raw) + warp;
    ingest_loop<SIMP>(S, P);
}
template <bool COUNT, bool SIMP> __global__ void __launch_bounds__(64, 8) ingest_glob(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSolver<VarsG, COUNT> S(P, lane);
    S.V.s = smem_base<SmemSmall>(smem_raw) + warp;
    ingest_loop<SIMP>(S, P);
}
#elif B200SAT_KERNEL_TU_PHASE == 2
/* ---------------------------------------------------------------- K2 bcp_replay: every parked item re-executes the
 * recorded prefix of its search with BCP as the only real work (measurement kernel for the BCP-only roofline) */
template <class WS> __device__ __forceinline__ void replay_loop(WS &S, const KParams &P) {
    for (;;) {
        u32 w = 0;
        if (S.lane == 0) w = atomicAdd(P.queue, 1u);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= P.n_work) break;
        w += P.chunk_base;
        bind_item(S, P, w);
        if constexpr (WS::VT_GLOBAL) S.V.bind(S.slab + P.L.off_vars, P.L.nv_cap);
        b200sat_result *r = P.results + S.res_index;
        if (r->verdict != B200SAT_INTERRUPTED || r->status != B200SAT_OK) continue; /* decided at ingest: nothing was recorded */
        S.load_state();
        const bool ok = S.bcp_replay(S.res_index);
        if (S.lane == 0) {
            r->status = ok ? B200SAT_OK : -(i32)(100 + ST_INTERNAL);
            r->stats[5] = S.sc().propagations;
            for (int i = 0; i < 8; i++) r->traffic[i] = S.sc().traffic[i];
        }
        __syncwarp();
    }
}
template <int NV, bool COUNT> __global__ void __launch_bounds__(32, 32) replay_smem(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsS<NV, false>, COUNT, false> S(P, threadIdx.x);
    S.V.s = smem_base<SmemVars<NV, false>>(smem_raw);
    replay_loop(S, P);
}
template <bool COUNT> __global__ void __launch_bounds__(32, 16) replay_glob(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsG, COUNT, false> S(P, threadIdx.x);
    S.V.s = smem_base<SmemSmall>(smem_raw);
    replay_loop(S, P);
}
#else
/* ---------------------------------------------------------------- search: pops the work ring until every item is done */
template <class WS> __device__ __forceinline__ void search_loop_ring(WS &S, const KParams &P) {
    bool first = true;
#ifndef B200SAT_EMU
    if (S.lane == 0 && blockIdx.x == 0) *(volatile unsigned long long *)(P.qctl + QC_T_START) = global_timer_ns();
#endif
    for (;;) {
        u32 w = 0xFFFFFFFFu;
        if (S.lane == 0) {
            /* the first item of every warp is handed out statically (ring slot = block index, published by the
             * ingest kernel): thousands of warps popping at once would otherwise serialise on the head word */
            if (first && blockIdx.x < WS::ld_vol(P.qctl + QC_HEAD0)) w = (u32)*(volatile unsigned long long *)(P.ring + (blockIdx.x & P.ring_mask));
            else for (;;) {
                const u32 h = WS::ld_vol(P.qctl + QC_HEAD), t = WS::ld_vol(P.qctl + QC_TAIL);
                if ((i32)(t - h) > 0) {
                    if (atomicCAS(P.qctl + QC_HEAD, h, h + 1) != h) continue;
                    volatile unsigned long long *slot = P.ring + (h & P.ring_mask);
                    unsigned long long e;
                    do { e = *slot; } while ((u32)(e >> 32) != h + 1);
                    w = (u32)e;
                    break;
                }
                /* Empty ring: this warp is done.  Items enter the ring only when a running instance yields, and it
                 * yields only while the ring is non-empty, so an empty ring stays empty -- except for a yield already
                 * under way, and the yielding warp pops again right after its push, so no item is ever stranded.
                 * (Waiting here instead was measured: thousands of idle warps polling the ring took most of the
                 * issue slots of the SMs that still had work.) */
                break;
            }
            __threadfence(); /* acquire: also drops this SM's stale L1 lines of the slab (CCTL.IVALL) */
        }
        first = false;
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w == 0xFFFFFFFFu) {
#ifndef B200SAT_EMU
            if (S.lane == 0) { /* tail accounting: when the ring first ran dry, when the last warp left */
                const unsigned long long t = global_timer_ns();
                atomicCAS((unsigned long long *)(P.qctl + QC_T_DRAIN), 0ull, t);
                atomicMax((unsigned long long *)(P.qctl + QC_T_END), t);
            }
#endif
            break;
        }
        bind_item(S, P, w);
        if constexpr (WS::VT_GLOBAL) S.V.bind(S.slab + P.L.off_vars, P.L.nv_cap);
        S.load_state();
        const u32 r = S.run_search();
        if (r == WS::SR_SUSPEND) {
            S.save_state();
            if (S.lane == 0) atomicAdd(P.qctl + QC_SUSPENDS, 1u);
            S.ring_push(w);
            continue;
        }
        S.finish_search(r, P.kind == B200SAT_SIMP);
        S.count_done();
    }
}
#if B200SAT_KERNEL_TU_RESCUE
/* ---------------------------------------------------------------- search with straggler rescue (free-running mode):
 * phase 1 = search_loop_ring with every popped instance registered as "main replica" of its warp; phase 2 = the warp
 * has nothing left to pop and helps instances that are still running (KParams::rescue, WarpSolver::rescue_*). */
template <class WS> __device__ __forceinline__ void search_loop_rescue(WS &S, const KParams &P) {
    const u32 my_warp = blockIdx.x, n_warps = gridDim.x;
    bool first = true, have_slab = false;
    if (S.lane == 0 && blockIdx.x == 0) *(volatile unsigned long long *)(P.qctl + QC_T_START) = global_timer_ns();
    for (;;) {
        u32 w = 0xFFFFFFFFu;
        if (S.lane == 0) {
            if (first && blockIdx.x < WS::ld_vol(P.qctl + QC_HEAD0)) w = (u32)*(volatile unsigned long long *)(P.ring + (blockIdx.x & P.ring_mask));
            else for (;;) {
                const u32 h = WS::ld_vol(P.qctl + QC_HEAD), t = WS::ld_vol(P.qctl + QC_TAIL);
                if ((i32)(t - h) > 0) {
                    if (atomicCAS(P.qctl + QC_HEAD, h, h + 1) != h) continue;
                    volatile unsigned long long *slot = P.ring + (h & P.ring_mask);
                    unsigned long long e;
                    do { e = *slot; } while ((u32)(e >> 32) != h + 1);
                    w = (u32)e;
                }
                break;
            }
            __threadfence();
        }
        first = false;
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w == 0xFFFFFFFFu) break;
        bind_item(S, P, w);
        if constexpr (WS::VT_GLOBAL) S.V.bind(S.slab + P.L.off_vars, P.L.nv_cap);
        S.load_state();
        have_slab = true;
        S.rescue_begin_main(my_warp, w);
        const u32 r = S.run_search();
        if (r == WS::SR_SUSPEND) {
            S.rescue_end_main(my_warp);
            S.save_state();
            if (S.lane == 0) atomicAdd(P.qctl + QC_SUSPENDS, 1u);
            S.ring_push(w);
            continue;
        }
        bool mine = !(r == WS::SR_ABORT && S.sc().status == ST_CANCELLED); /* a helper decided it first */
        if (mine) mine = S.rescue_claim();
        if (mine) S.finish_search(r, P.kind == B200SAT_SIMP);
        S.rescue_end_main(my_warp);
    }
    if (S.lane == 0) {
        const unsigned long long t = global_timer_ns();
        atomicCAS((unsigned long long *)(P.qctl + QC_T_DRAIN), 0ull, t);
    }
    if (have_slab) {
        for (u32 round = 0;; round++) {
            const u32 mw = S.rescue_pick(n_warps, my_warp, round);
            if (mw == 0xFFFFFFFFu) break;
            if (!S.rescue_join(mw)) continue;
            const u32 r = S.run_search();
            if ((r == WS::SR_SAT || r == WS::SR_UNSAT) && S.rescue_claim()) {
                S.finish_search(r, P.kind == B200SAT_SIMP);
                if (S.lane == 0) atomicAdd(P.qctl + QC_RESCUED, 1u);
            }
            __syncwarp();
        }
    }
    if (S.lane == 0) atomicMax((unsigned long long *)(P.qctl + QC_T_END), global_timer_ns());
}
template <int NV> __global__ void __launch_bounds__(32, B200SAT_SEARCH_BLOCKS) rescue_smem(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsS<NV, false>, false, true> S(P, threadIdx.x);
    S.V.s = smem_base<SmemVars<NV, false>>(smem_raw);
    search_loop_rescue(S, P);
}
__global__ void __launch_bounds__(32, 16) rescue_glob(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsG, false, true> S(P, threadIdx.x);
    S.V.s = smem_base<SmemSmall>(smem_raw);
    search_loop_rescue(S, P);
}
#else
/* ONE warp per block: the shared-memory base and the lane id are then block constants (no per-warp offset to
 * rematerialise in the hot loop); 32 blocks per SM = 32 resident warps at 64 registers per thread */
template <int NV, bool COUNT, bool PF> __global__ void __launch_bounds__(32, B200SAT_SEARCH_BLOCKS) search_smem(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsS<NV, false>, COUNT, PF> S(P, threadIdx.x);
    S.V.s = smem_base<SmemVars<NV, false>>(smem_raw);
    search_loop_ring(S, P);
}
template <bool COUNT, bool PF> __global__ void __launch_bounds__(32, 16) search_glob(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSolver<VarsG, COUNT, PF> S(P, threadIdx.x);
    S.V.s = smem_base<SmemSmall>(smem_raw);
    search_loop_ring(S, P);
}
#endif
#endif

} // namespace b200sat

kernel_fn B200SAT_KERNEL_TU_NAME(int tier_nv, size_t *spw, size_t *save) {
    using namespace b200sat;
    constexpr bool C = B200SAT_KERNEL_TU_COUNT != 0;
#if B200SAT_KERNEL_TU_PHASE == 0
    constexpr bool Sp = B200SAT_KERNEL_TU_SIMP != 0;
    switch (tier_nv) {
    case 128: *spw = sizeof(SmemVars<128, Sp>); *save = SmemSave<128>::BYTES; return ingest_smem<128, C, Sp>;
    case 256: *spw = sizeof(SmemVars<256, Sp>); *save = SmemSave<256>::BYTES; return ingest_smem<256, C, Sp>;
    case 1024: *spw = sizeof(SmemVars<1024, Sp>); *save = SmemSave<1024>::BYTES; return ingest_smem<1024, C, Sp>;
    default: *spw = sizeof(SmemSmall); *save = sizeof(SmemSmall); return ingest_glob<C, Sp>;
    }
#elif B200SAT_KERNEL_TU_PHASE == 2
    switch (tier_nv) {
    case 128: *spw = sizeof(SmemVars<128, false>); *save = SmemSave<128>::BYTES; return replay_smem<128, C>;
    case 256: *spw = sizeof(SmemVars<256, false>); *save = SmemSave<256>::BYTES; return replay_smem<256, C>;
    case 1024: *spw = sizeof(SmemVars<1024, false>); *save = SmemSave<1024>::BYTES; return replay_smem<1024, C>;
    default: *spw = sizeof(SmemSmall); *save = sizeof(SmemSmall); return replay_glob<C>;
    }
#elif B200SAT_KERNEL_TU_RESCUE
    switch (tier_nv) {
    case 128: *spw = sizeof(SmemVars<128, false>); *save = SmemSave<128>::BYTES; return rescue_smem<128>;
    case 256: *spw = sizeof(SmemVars<256, false>); *save = SmemSave<256>::BYTES; return rescue_smem<256>;
    case 1024: *spw = sizeof(SmemVars<1024, false>); *save = SmemSave<1024>::BYTES; return rescue_smem<1024>;
    default: *spw = sizeof(SmemSmall); *save = sizeof(SmemSmall); return rescue_glob;
    }
#else
    constexpr bool Pf = B200SAT_KERNEL_TU_PF != 0;
    switch (tier_nv) {
    case 128: *spw = sizeof(SmemVars<128, false>); *save = SmemSave<128>::BYTES; return search_smem<128, C, Pf>;
    case 256: *spw = sizeof(SmemVars<256, false>); *save = SmemSave<256>::BYTES; return search_smem<256, C, Pf>;
    case 1024: *spw = sizeof(SmemVars<1024, false>); *save = SmemSave<1024>::BYTES; return search_smem<1024, C, Pf>;
    default: *spw = sizeof(SmemSmall); *save = sizeof(SmemSmall); return search_glob<C, Pf>;
    }
#endif
}
#endif
#endif
