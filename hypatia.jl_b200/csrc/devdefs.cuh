// Definitions shared by the kernel headers (*_kernels.cuh).  The kernel headers contain only
// __global__ / __device__ code and include nothing else, so that tests/emu/ can compile them for
// the host (HYP_EMU: one pthread per CUDA thread, barriers for __syncthreads) and check the device
// code against the CPU oracle in the CPU-only test tier.  The emulation is test infrastructure; the
// library itself is built by nvcc for sm_100a only and has no CPU path.
#pragma once
#ifdef HYP_EMU
#include "../../tests/emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#define HYP_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

// Several threads of a block raising the same shared flag between two barriers (all store the same
// value; it is read after the next __syncthreads).  A plain store on the device; the emulation
// spells it as a relaxed atomic so that ThreadSanitizer (tools/emu_tsan.sh) does not report it.
#ifdef HYP_EMU
#define HYP_RAISE_FLAG(flag) __atomic_store_n(&(flag), 1, __ATOMIC_RELAXED)
#else
#define HYP_RAISE_FLAG(flag) (flag) = 1
#endif

#ifndef HYP_EPS
#define HYP_EPS 2.220446049250313e-16
#endif
#define HYP_RTEPS 1.4901161193847656e-08   // sqrt(eps)

// separable spectral functions of EpiPerSepSpectral (sepspectralfun.jl:17-116); = HYP_SSF_* of the ABI
#define SSF_INV 0
#define SSF_NEGLOG 1
#define SSF_NEGENTROPY 2
#define SSF_POWER12 3

namespace hypdev {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; `sm` holds one double per warp; every thread gets the total
__device__ __forceinline__ double block_sum(double v, double* sm) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = (int)((blockDim.x + 31) >> 5);
    for (int i = 0; i < nw; i++) t += sm[i];
    __syncthreads();
    return t;
}

// row / column of entry idx of the svec order (columns of the upper triangle, a <= b)
__device__ __forceinline__ void svec_rc(int64_t idx, int& a, int& b) {
    int bb = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((int64_t)(bb + 1) * (bb + 2) / 2 <= idx) bb++;
    while ((int64_t)bb * (bb + 1) / 2 > idx) bb--;
    b = bb;
    a = (int)(idx - (int64_t)bb * (bb + 1) / 2);
}

}  // namespace hypdev
