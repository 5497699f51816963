/* cdcl_warp kernels: COUNT=0 SIMP=0 */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_c0s0
#define B200SAT_KERNEL_TU_COUNT 0
#define B200SAT_KERNEL_TU_SIMP 0
#include "cdcl_kernels.cuh"
