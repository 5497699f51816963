/* cdcl_warp kernels: COUNT=1 SIMP=1 */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_c1s1
#define B200SAT_KERNEL_TU_COUNT 1
#define B200SAT_KERNEL_TU_SIMP 1
#include "cdcl_kernels.cuh"
