// Minimal host emulation of the CUDA execution model for the kernel headers under
// hypatia.jl_b200/csrc/*_kernels.cuh (test infrastructure; CPU-only test tier).
//
// One pthread per CUDA thread of a block, blocks run one after the other; __syncthreads is a
// pthread barrier over the block, warp shuffles exchange through a per-block buffer guarded by
// per-warp barriers; __shared__ variables become function-level statics (blocks are sequential, so
// a static is "per block") collected in one ELF section, which emu::launch fills with NaN bytes before
// every block - shared memory does not survive a block on the device either.  Only what the kernel
// headers use is provided.
#pragma once
#include <pthread.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct double2 {
    double x, y;
};

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace emu {
struct BlockState {
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<double> shf;
    int nthreads = 0;
};
extern BlockState* g_block;
extern void* g_dyn_smem;
}  // namespace emu

extern thread_local dim3 threadIdx;
extern thread_local dim3 blockIdx;
extern dim3 blockDim;
extern dim3 gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#ifdef HYP_EMU_PLAIN_SHARED
// AddressSanitizer build (tools/emu_asan.sh): plain statics, so that every __shared__ array keeps its red zones
// (globals in a named section are not instrumented); the per-block NaN fill then covers dynamic shared memory only
#define __shared__ static
#else
#define __shared__ static __attribute__((section("emu_shared")))
#endif
#define HYP_DYN_SMEM(type, name) type* name = (type*)emu::g_dyn_smem
#define INFINITY_EMU INFINITY

template <class T>
static inline T __ldg(const T* p) {
    return *p;
}

static inline void __syncthreads() { pthread_barrier_wait(&emu::g_block->bar); }

static inline void __syncwarp() { pthread_barrier_wait(&emu::g_block->warp_bar[threadIdx.x >> 5]); }

static inline double emu_shfl(double v, int src_lane_abs) {
    emu::BlockState* b = emu::g_block;
    const int tid = (int)threadIdx.x;
    const int w = tid >> 5;
    b->shf[tid] = v;
    pthread_barrier_wait(&b->warp_bar[w]);
    double out = (src_lane_abs >= 0 && src_lane_abs < b->nthreads) ? b->shf[src_lane_abs] : v;
    pthread_barrier_wait(&b->warp_bar[w]);
    return out;
}
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    return emu_shfl(v, (int)(threadIdx.x ^ (unsigned)o));
}
static inline double __shfl_sync(unsigned, double v, int lane) {
    return emu_shfl(v, (int)((threadIdx.x & ~31u) + (unsigned)lane));
}

static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline int atomicAdd(int* addr, int val) { return __atomic_fetch_add(addr, val, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* addr, unsigned long long val) {
    unsigned long long old = __atomic_load_n(addr, __ATOMIC_SEQ_CST);
    while (old < val && !__atomic_compare_exchange_n(addr, &old, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
static inline double __ldcg(const double* p) { return *p; }
// conversions / byte permute of the digit slicing (ozaki_slice_kernels.cuh)
static inline long long __double2ll_rn(double x) { return llrint(x); }   // default rounding mode: to nearest even
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, sizeof d);
    return d;
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// blocks run one after the other and the callers below are single-threaded at the call site (threadIdx.x == 0)
static inline int atomicCAS(int* addr, int compare, int val) {
    int old = __atomic_load_n(addr, __ATOMIC_SEQ_CST);
    __atomic_compare_exchange_n(addr, &compare, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return old;
}

namespace emu {
// run `body` for every thread of every block of the grid (1-D blocks, up to 2-D grids)
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);
}
