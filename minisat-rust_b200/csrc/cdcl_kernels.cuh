/*
 * cdcl_kernels.cuh -- the persistent warp-per-instance kernels.  Included by b200sat.cu for
 * the declarations only; each kern_*.cu defines B200SAT_KERNEL_TU_COUNT / _SIMP and gets the
 * kernel templates plus one `b200sat_pick_*` accessor for its COUNT x SIMP combination, so the
 * four combinations compile in parallel.
 */
#ifndef B200SAT_CDCL_KERNELS_CUH
#define B200SAT_CDCL_KERNELS_CUH
#include "cdcl_types.h"

typedef void (*kernel_fn)(const b200sat::KParams);
kernel_fn b200sat_pick_c0s0(int tier_nv, size_t *smem_per_warp);
kernel_fn b200sat_pick_c1s0(int tier_nv, size_t *smem_per_warp);
kernel_fn b200sat_pick_c0s1(int tier_nv, size_t *smem_per_warp);
kernel_fn b200sat_pick_c1s1(int tier_nv, size_t *smem_per_warp);

#ifdef B200SAT_KERNEL_TU_NAME
#include <cuda_runtime.h>
#include "cdcl_warp.cuh"
#if B200SAT_KERNEL_TU_SIMP
#include "simp_warp.cuh"
#endif

#ifndef B200SAT_MINBLOCKS
#define B200SAT_MINBLOCKS 16
#endif

namespace b200sat {
template <bool SIMP, class VT, bool COUNT> __device__ __forceinline__ void warp_work_loop(WarpSolver<VT, COUNT> &S, const KParams &P) {
    for (;;) {
        u32 w = 0;
        if (S.lane == 0) w = atomicAdd(P.queue, 1u);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= P.n_work) break;
        u32 inst = P.work ? P.work[w] : w;
        S.res_index = inst;
        if (P.portfolio) { /* replica w of instance 0 with its own diversified settings */
            S.res_index = w;
            S.stp = P.replica_settings + w;
            inst = 0;
        }
#if B200SAT_KERNEL_TU_SIMP
        if (SIMP) { solve_simp(S, inst); continue; }
#endif
        S.solve_core(inst);
    }
}

/* <= 64 threads per block and >= 16 blocks per SM: 64 registers per thread, 32 resident warps per SM */
template <int NV, bool COUNT, bool SIMP> __global__ void __launch_bounds__(64, B200SAT_MINBLOCKS) cdcl_warp_smem(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5;
    WarpSolver<VarsS<NV, SIMP>, COUNT> S(P, lane);
    S.V.s = reinterpret_cast<SmemVars<NV, SIMP> *>(smem_raw) + warp;
    S.slab = P.slabs + (size_t)(blockIdx.x * wpb + warp) * P.L.slab_bytes;
    warp_work_loop<SIMP>(S, P);
}

template <bool COUNT, bool SIMP> __global__ void __launch_bounds__(64, 8) cdcl_warp_glob(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 wpb = blockDim.x >> 5;
    WarpSolver<VarsG, COUNT> S(P, lane);
    S.V.s = reinterpret_cast<SmemSmall *>(smem_raw) + warp;
    S.slab = P.slabs + (size_t)(blockIdx.x * wpb + warp) * P.L.slab_bytes;
    S.V.bind(S.slab + P.L.off_vars, P.L.nv_cap);
    warp_work_loop<SIMP>(S, P);
}


} // namespace b200sat

kernel_fn B200SAT_KERNEL_TU_NAME(int tier_nv, size_t *spw) {
    using namespace b200sat;
    constexpr bool C = B200SAT_KERNEL_TU_COUNT != 0, Sp = B200SAT_KERNEL_TU_SIMP != 0;
    switch (tier_nv) {
    case 128: *spw = sizeof(SmemVars<128, Sp>); return cdcl_warp_smem<128, C, Sp>;
    case 256: *spw = sizeof(SmemVars<256, Sp>); return cdcl_warp_smem<256, C, Sp>;
    case 1024: *spw = sizeof(SmemVars<1024, Sp>); return cdcl_warp_smem<1024, C, Sp>;
    default: *spw = sizeof(SmemSmall); return cdcl_warp_glob<C, Sp>;
    }
}
#endif
#endif
