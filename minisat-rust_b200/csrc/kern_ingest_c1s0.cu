/* ingest kernels: CoreSolver, traffic counters on */
#define B200SAT_KERNEL_TU_NAME b200sat_pick_ingest_c1s0
#define B200SAT_KERNEL_TU_COUNT 1
#define B200SAT_KERNEL_TU_SIMP 0
#define B200SAT_KERNEL_TU_PHASE 0
#define B200SAT_KERNEL_TU_PF 0
#define B200SAT_KERNEL_TU_RESCUE 0
#include "cdcl_kernels.cuh"
