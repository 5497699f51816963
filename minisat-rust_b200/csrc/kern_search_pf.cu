/* search kernels: portfolio mode (clause exchange + first-finisher cancel) */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_search_pf
#define B200SAT_KERNEL_TU_COUNT 0
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 1
#define B200SAT_KERNEL_TU_PF 1
#define B200SAT_KERNEL_TU_RESCUE 0
#include "cdcl_kernels.cuh"
