/*
 * simt_emu.h -- TEST-ONLY single-warp SIMT emulator.
 *
 * Lets the device source (minisat-rust_b200/csrc/cdcl_warp.cuh) be compiled
 * with g++ and run on the CPU so that trace parity against the oracle can be
 * debugged in a container without a GPU.  It is test infrastructure: nothing
 * in the product links or loads it, and results produced through it are never
 * reported as GPU results.
 *
 * Model: the 32 lanes of one warp are 32 ucontext coroutines.  A lane runs
 * until it reaches a warp collective (__ballot_sync, __shfl*_sync, __syncwarp);
 * when all 32 lanes have arrived at the SAME collective the results are
 * computed and the lanes are resumed.  Between collectives the lanes are run
 * one after another in a pseudo-random order that changes at every step, so
 * code that forgets a __syncwarp() between a write by one lane and a read by
 * another sees stale or too-new data for some seeds.  Lanes arriving at
 * different collectives abort the run (divergence bug).
 */
#ifndef SIMT_EMU_H
#define SIMT_EMU_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define DEV inline __attribute__((always_inline))
#define DEVNI __attribute__((noinline))

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

namespace simt {

enum Kind { K_NONE = 0, K_BALLOT, K_SHFL, K_SHFL_XOR, K_SHFL_UP, K_SYNC, K_MATCH, K_EXIT };

struct Warp {
    ucontext_t sched;
    ucontext_t lanes[32];
    char *stacks[32];
    int kind[32];
    uint64_t arg[32];
    int arg2[32];
    uint64_t res[32];
    bool done[32];
    int cur;
    uint64_t rng;
    uint64_t n_collectives;
    void (*entry)(void *, int);
    void *user;
};

extern Warp *g_warp;

static inline void arrive(int kind, uint64_t a, int a2) {
    Warp *w = g_warp;
    int l = w->cur;
    w->kind[l] = kind;
    w->arg[l] = a;
    w->arg2[l] = a2;
    swapcontext(&w->lanes[l], &w->sched);
}
static inline uint64_t result() { return g_warp->res[g_warp->cur]; }

void run_warp(void (*entry)(void *, int), void *user, uint64_t seed);

} // namespace simt

static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    if (mask != 0xFFFFFFFFu) { fprintf(stderr, "simt_emu: partial mask\n"); abort(); }
    simt::arrive(simt::K_BALLOT, pred ? 1 : 0, 0);
    return (uint32_t)simt::result();
}
template <class T> static inline T __shfl_sync(uint32_t mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl size");
    uint64_t a = 0; memcpy(&a, &v, sizeof(T));
    simt::arrive(simt::K_SHFL, a, src & 31);
    uint64_t r = simt::result(); T o; memcpy(&o, &r, sizeof(T)); return o;
}
template <class T> static inline T __shfl_xor_sync(uint32_t mask, T v, int x) {
    uint64_t a = 0; memcpy(&a, &v, sizeof(T));
    simt::arrive(simt::K_SHFL_XOR, a, x);
    uint64_t r = simt::result(); T o; memcpy(&o, &r, sizeof(T)); return o;
}
template <class T> static inline T __shfl_up_sync(uint32_t mask, T v, int d) {
    uint64_t a = 0; memcpy(&a, &v, sizeof(T));
    simt::arrive(simt::K_SHFL_UP, a, d);
    uint64_t r = simt::result(); T o; memcpy(&o, &r, sizeof(T)); return o;
}
static inline uint32_t __match_any_sync(uint32_t mask, uint32_t v) {
    simt::arrive(simt::K_MATCH, v, 0);
    return (uint32_t)simt::result();
}
static inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu) { simt::arrive(simt::K_SYNC, 0, 0); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline float __uint_as_float(uint32_t x) { float f; memcpy(&f, &x, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t x; memcpy(&x, &f, 4); return x; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) {}
static inline void __threadfence_system() {}
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }

#endif
