/* search kernels: batch mode, counters off (the hot kernel) */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_search_c0
#define B200SAT_KERNEL_TU_COUNT 0
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 1
#define B200SAT_KERNEL_TU_PF 0
#define B200SAT_KERNEL_TU_RESCUE 0
#include "cdcl_kernels.cuh"
