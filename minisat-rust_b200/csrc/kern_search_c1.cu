/* search kernels: batch mode, traffic counters on */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_search_c1
#define B200SAT_KERNEL_TU_COUNT 1
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 1
#define B200SAT_KERNEL_TU_PF 0
#define B200SAT_KERNEL_TU_RESCUE 0
#include "cdcl_kernels.cuh"
