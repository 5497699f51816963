/* cdcl_warp kernels: COUNT=1 SIMP=0 */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_c1s0
#define B200SAT_KERNEL_TU_COUNT 1
#define B200SAT_KERNEL_TU_SIMP 0
#include "cdcl_kernels.cuh"
