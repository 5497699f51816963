/* K2 bcp_replay kernels (measurement: BCP-only roofline), traffic counters off */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_replay_c0
#define B200SAT_KERNEL_TU_COUNT 0
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 2
#define B200SAT_KERNEL_TU_PF 0
#define B200SAT_KERNEL_TU_RESCUE 0
#include "cdcl_kernels.cuh"
