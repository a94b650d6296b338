// O(q) / O(n) vector kernels: axpby, linear combinations, deterministic dot.
// All HBM-bound; grid sizes are multiples of the SM count with grid-stride loops.
#include "common.cuh"

namespace {

// out = a * x + sign * dscal[0] * z  (scalar read from device memory: no host round trip)
__global__ void axpy_dev_kernel(int64_t len, double* __restrict__ out, double a,
                                const double* __restrict__ x, const double* __restrict__ dscal,
                                double sign, const double* __restrict__ z) {
    const double t = sign * dscal[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + t * z[i];
}

__global__ void zero_outside_kernel(double* __restrict__ v, int64_t q, int64_t lo, int64_t hi) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x)
        if (i < lo || i >= hi) v[i] = 0.0;
}

__global__ void axpby_kernel(int64_t len, double a, const double* __restrict__ x, double b,
                             double* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        double yv = (b == 0.0) ? 0.0 : b * y[i];
        y[i] = a * x[i] + yv;
    }
}

__global__ void lincomb3_kernel(int64_t len, double* __restrict__ out, double a,
                                const double* __restrict__ x, double b, const double* __restrict__ y,
                                double c, const double* __restrict__ z) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = a * x[i];
        if (y) v += b * y[i];
        if (z) v += c * z[i];
        out[i] = v;
    }
}

__global__ void fill_kernel(int64_t len, double* __restrict__ dst, double val) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = val;
}

// Stage 1: every block reduces a fixed contiguous slice in a fixed order -> partial[block].
__global__ void dot_stage1_kernel(int64_t len, const double* __restrict__ x,
                                  const double* __restrict__ y, double* __restrict__ partial) {
    __shared__ double sm[256];
    int64_t per = (len + gridDim.x - 1) / gridDim.x;
    int64_t lo = blockIdx.x * per, hi = min(len, lo + per);
    double acc = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += x[i] * y[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void dot_stage2_kernel(int nparts, const double* __restrict__ partial,
                                  double* __restrict__ out, int accumulate) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + sm[0] : sm[0];
}

inline int grid_for(hyp_ctx* ctx, int64_t len, int threads) {
    int64_t blocks = (len + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

void hyp_axpy_dev(hyp_ctx* ctx, int64_t len, double* out, double a, const double* x,
                  const double* dscal, double sign, const double* z) {
    if (len <= 0) return;
    axpy_dev_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(len, out, a, x, dscal, sign, z);
    ctx->launches++;
}

void hyp_zero_outside(hyp_ctx* ctx, double* v) {
    if (!hyp_row_sharded(ctx) || ctx->q == 0) return;
    zero_outside_kernel<<<grid_for(ctx, ctx->q, 256), 256, 0, ctx->stream>>>(v, ctx->q, ctx->row_lo,
                                                                             ctx->row_hi);
    ctx->launches++;
}

void hyp_axpby(hyp_ctx* ctx, int64_t len, double a, const double* x, double b, double* y) {
    if (len <= 0) return;
    axpby_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(len, a, x, b, y);
    ctx->launches++;
}

void hyp_lincomb3(hyp_ctx* ctx, int64_t len, double* out, double a, const double* x, double b,
                  const double* y, double c, const double* z) {
    if (len <= 0) return;
    lincomb3_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(len, out, a, x, b, y, c, z);
    ctx->launches++;
}

void hyp_copy(hyp_ctx* ctx, int64_t len, double* dst, const double* src) {
    if (len <= 0 || dst == src) return;
    CUDA_TRY(cudaMemcpyAsync(dst, src, len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
}

void hyp_fill(hyp_ctx* ctx, int64_t len, double* dst, double val) {
    if (len <= 0) return;
    fill_kernel<<<grid_for(ctx, len, 256), 256, 0, ctx->stream>>>(len, dst, val);
    ctx->launches++;
}

void hyp_dot(hyp_ctx* ctx, int64_t len, const double* x, const double* y, double* d_out,
             bool accumulate) {
    if (len <= 0) {
        if (!accumulate) hyp_fill(ctx, 1, d_out, 0.0);
        return;
    }
    int nparts = (int)std::min<int64_t>(ctx->sm_count * 2, (len + 1023) / 1024);
    if (nparts < 1) nparts = 1;
    dot_stage1_kernel<<<nparts, 256, 0, ctx->stream>>>(len, x, y, ctx->d_partial);
    dot_stage2_kernel<<<1, 256, 0, ctx->stream>>>(nparts, ctx->d_partial, d_out, accumulate ? 1 : 0);
    ctx->launches += 2;
}

void hyp_time_begin(hyp_ctx* ctx, int slot) {
    if (!ctx->timing_enabled) return;
    TimingSlot& t = ctx->timing[slot];
    if (!t.e0) {
        cudaEventCreate(&t.e0);
        cudaEventCreate(&t.e1);
    }
    cudaEventRecord(t.e0, ctx->stream);
    t.open = true;
}

void hyp_time_end(hyp_ctx* ctx, int slot) {
    if (!ctx->timing_enabled) return;
    TimingSlot& t = ctx->timing[slot];
    if (!t.open) return;
    cudaEventRecord(t.e1, ctx->stream);
    cudaEventSynchronize(t.e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    t.total_ms += ms;
    t.count++;
    t.open = false;
}
