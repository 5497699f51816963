/* search kernels with straggler rescue (free-running mode): idle warps help instances that are still running */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_search_rescue
#define B200SAT_KERNEL_TU_COUNT 0
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 1
#define B200SAT_KERNEL_TU_PF 1
#define B200SAT_KERNEL_TU_RESCUE 1
#include "cdcl_kernels.cuh"
