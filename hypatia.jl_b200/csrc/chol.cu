// Dense Cholesky factorisation and triangular solves of the Schur complement (K3 / K5 of
// SURVEY.md section 2.3).
//
// reference call sites: posdef_fact!(A) = cholesky!(Symmetric(A, :U), check=false)
// (src/linearalgebra/dense.jl:191-192, LAPACK dpotrf 'U') called from update_lhs_fact
// (qrchol.jl:249-250); ldiv!(x, fact, rhs) (qrchol.jl:68, LAPACK dpotrs).
//
// potrf: right-looking blocked upper Cholesky, block size 128.
//   per block column k:  (1) one-CTA panel kernel: factor the 128 x 128 diagonal block in
//   registers (1024 threads x 4 x 4 cyclic sub-blocks, one barrier per pivot) and invert the
//   triangular factor in shared memory;  (2) U12 = U11^-T A12 as a TN GEMM with the inverted
//   block (syrk.cu, TMA + DMMA);  (3) trailing update A22 -= U12' U12 on the upper tiles with the
//   same TMA + DMMA kernel as the Schur SYRK.  Bound: tensor (FP64 DMMA); m^3/3 flops.
//   The inverted diagonal blocks are kept: the triangular solves use them.
// trsv: one persistent kernel per triangular solve; CTAs take block columns in dependency order
//   from a ticket counter and publish finished 128-blocks of the solution through release /
//   acquire flags, so a solve is one launch instead of 2 * m / 128.  Bound: HBM (reads the
//   triangle once: 4 m^2 bytes) - in practice latency of the block dependency chain.
#include "common.cuh"

namespace {

constexpr int NB = 128;
constexpr int LDU = NB + 1;   // padded leading dimension of the shared-memory block

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Factor (FACTOR) and invert one upper-triangular diagonal block of at most 128 x 128, by one CTA
// of 1024 threads.  Ab: the block (upper part is read; U is written back, optionally with zeros
// below the diagonal); Db receives U^-1 (dn x dn entries, leading dim ldd; rows / columns past nb
// are identity padding).  Returns the 1-based index of the first non-positive pivot or 0.
template <bool FACTOR>
__device__ __forceinline__ int panel_body(double* __restrict__ Ab, int64_t lda, int nb,
                                          double* __restrict__ Db, int ldd, int dn, bool zero_lower,
                                          double* sU, double (*rowbuf)[NB]) {
    const int tid = threadIdx.x;
    const int tj = tid & 31;                  // row residue
    const int ti = tid >> 5;                  // col residue
    int bad = 0;

    double v[4][4];   // v[a][b] = element (row tj + 32 a, col ti + 32 b)
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int r = tj + 32 * a, c = ti + 32 * b;
            double x = 0.0;
            if (r < nb && c < nb) {
                if (r <= c) x = Ab[r + (int64_t)c * lda];
            } else if (r == c) {
                x = 1.0;   // identity padding of a ragged block
            }
            v[a][b] = x;
        }

    if (FACTOR) {
        for (int j = 0; j < NB; j++) {
            const int aj = j >> 5, rj = j & 31;
            if (tj == rj) {
#pragma unroll
                for (int a = 0; a < 4; a++)
                    if (a == aj) {
#pragma unroll
                        for (int b = 0; b < 4; b++) rowbuf[j & 1][ti + 32 * b] = v[a][b];
                    }
            }
            __syncthreads();
            const double* rb = rowbuf[j & 1];
            double d = rb[j];
            double sd, rinv;
            if (d > 0.0) {
                sd = sqrt(d);
                rinv = 1.0 / sd;
            } else {
                if (!bad) bad = j + 1;
                sd = 0.0;
                rinv = 0.0;
            }
            double uc[4];
#pragma unroll
            for (int b = 0; b < 4; b++) uc[b] = rb[ti + 32 * b] * rinv;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                int r = tj + 32 * a;
                if (r > j) {
                    double ur = rb[r] * rinv;
#pragma unroll
                    for (int b = 0; b < 4; b++) v[a][b] -= ur * uc[b];
                } else if (r == j) {
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        int c = ti + 32 * b;
                        v[a][b] = (c == j) ? sd : uc[b];
                    }
                }
            }
        }
        // write U back (upper part of the valid block); coalesced along rows
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int r = tj + 32 * a, c = ti + 32 * b;
                if (r < nb && c < nb) {
                    if (r <= c) Ab[r + (int64_t)c * lda] = v[a][b];
                    else if (zero_lower) Ab[r + (int64_t)c * lda] = 0.0;
                }
            }
    }

    // ---- stage U in shared memory (zeros below the diagonal) and invert in place ----
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int r = tj + 32 * a, c = ti + 32 * b;
            sU[r + c * LDU] = (r <= c) ? v[a][b] : 0.0;
        }
    __syncthreads();

    // unblocked upper triangular inverse (dtrti2): column by column,
    // X[0:j, j] = -X[j, j] * X[0:j, 0:j] * U[0:j, j]; 8 threads per row split the k range.
    const int ri = tid >> 3, sub = tid & 7;
    for (int j = 0; j < NB; j++) {
        double ujj = sU[j + j * LDU];
        double ajj = (ujj != 0.0) ? 1.0 / ujj : 0.0;
        double acc = 0.0;
        if (ri < j) {
            for (int k = ri + sub; k < j; k += 8) acc += sU[ri + k * LDU] * sU[k + j * LDU];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        __syncthreads();   // all reads of column j done
        if (sub == 0) {
            if (ri < j) sU[ri + j * LDU] = -ajj * acc;
            else if (ri == j) sU[j + j * LDU] = ajj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < dn * dn; idx += 1024) {
        int r = idx % dn, c = idx / dn;
        Db[r + (int64_t)c * ldd] = sU[r + c * LDU];
    }
    return bad;
}

//   FACTOR:  grid = 1; diagonal block blk0 of the m x m matrix A; a bad pivot sets *info.
//   !FACTOR: grid = number of diagonal blocks of the m x m triangular matrix A; inverts only.
template <bool FACTOR>
__global__ void __launch_bounds__(1024, 1)
panel_kernel(double* __restrict__ A, int64_t lda, int64_t m, int64_t blk0, double* __restrict__ dinv,
             int* __restrict__ info) {
    extern __shared__ double sU[];            // NB x NB col-major, ld = LDU
    __shared__ double rowbuf[2][NB];
    const int64_t blk = FACTOR ? blk0 : (int64_t)blockIdx.x;
    const int64_t k0 = blk * NB;
    const int nb = (int)min((int64_t)NB, m - k0);
    int bad = panel_body<FACTOR>(A + k0 + k0 * lda, lda, nb, dinv + blk * (int64_t)NB * NB, NB, NB, false,
                                 sU, rowbuf);
    if (FACTOR && bad && threadIdx.x == 0) atomicCAS(info, 0, (int)(k0 + bad));
}

// Batched variant for the matrix cones: CTA c factors the side x side matrix at U + moff[c]
// (leading dim = side rounded up to even) in place and writes its inverse to Ui + moff[c].
// A failed factorisation clears flag[kidx[c]].  Cones with side > 128 are skipped (blocked path).
__global__ void __launch_bounds__(1024, 1)
chol_batched_kernel(int ncones, const int* __restrict__ sides, const int64_t* __restrict__ moff,
                    const int* __restrict__ kidx, double* __restrict__ U, double* __restrict__ Ui,
                    uint8_t* __restrict__ flag) {
    extern __shared__ double sU[];
    __shared__ double rowbuf[2][NB];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int side = sides[c];
    if (side > NB) return;
    const int lde = (side + 1) & ~1;
    int bad = panel_body<true>(U + moff[c], lde, side, Ui + moff[c], lde, side, true, sU, rowbuf);
    if (bad && threadIdx.x == 0) flag[kidx[c]] = 0;
}

// ---- triangular solve with the blocked factor -------------------------------------------
// flags[0] = ticket counter (zeroed by the host before the launch), flags[1 + k] = epoch when
// block k of the solution is final.
template <bool TRANS>
__global__ void __launch_bounds__(256)
trsv_kernel(const double* __restrict__ F, int64_t ldf, int64_t m, const double* __restrict__ dinv,
            double* x, int* flags, int nblk, int epoch) {
    __shared__ double sv[NB];
    __shared__ double sacc[2][NB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) s_ticket = atomicAdd(&flags[0], 1);
        __syncthreads();
        const int t = s_ticket;
        __syncthreads();
        if (t >= nblk) return;
        const int k = TRANS ? t : nblk - 1 - t;
        const int64_t c0 = (int64_t)k * NB;
        const double* Dk = dinv + (int64_t)k * NB * NB;

        if (TRANS) {
            // y_k = Dinv_k' (b_k - sum_{j<k} U[j-block, k-block]' y_j); warp w owns 16 columns
            double pacc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) pacc[i] = 0.0;
            for (int j = 0; j < k; j++) {
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                if (tid < NB) sv[tid] = __ldcg(x + (int64_t)j * NB + tid);
                __syncthreads();
                const double* Ut = F + (int64_t)j * NB + (c0 + warp * 16) * ldf;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (c0 + warp * 16 + i < m) {
                        const double* col = Ut + (int64_t)i * ldf;
                        pacc[i] += col[lane] * sv[lane] + col[lane + 32] * sv[lane + 32] +
                                   col[lane + 64] * sv[lane + 64] + col[lane + 96] * sv[lane + 96];
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                double a = pacc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                pacc[i] = a;
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int64_t c = c0 + warp * 16 + i;
                    sv[warp * 16 + i] = (c < m) ? (x[c] - pacc[i]) : 0.0;
                }
            }
            __syncthreads();
            // y[c] = sum_{r <= c} Dinv[r, c] v[r]
            double res[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const double* col = Dk + (warp * 16 + i) * NB;
                double a = col[lane] * sv[lane] + col[lane + 32] * sv[lane + 32] +
                           col[lane + 64] * sv[lane + 64] + col[lane + 96] * sv[lane + 96];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                res[i] = a;
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int64_t c = c0 + warp * 16 + i;
                    if (c < m) x[c] = res[i];
                }
            }
        } else {
            // x_k = Dinv_k (y_k - sum_{j>k} U[k-block, j-block] x_j); thread owns a row, two
            // halves of the CTA split the 128 columns of a tile
            const int r = tid & (NB - 1), half = tid >> 7;
            const int64_t grow = c0 + r;
            double acc = 0.0;
            for (int j = nblk - 1; j > k; j--) {
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                if (tid < NB) {
                    int64_t c = (int64_t)j * NB + tid;
                    sv[tid] = (c < m) ? __ldcg(x + c) : 0.0;
                }
                __syncthreads();
                if (grow < m) {
                    const int64_t cb = (int64_t)j * NB + half * 64;
                    const int ncv = (int)max((int64_t)0, min((int64_t)64, m - cb));
                    const double* Ut = F + grow + cb * ldf;
                    const double* svh = sv + half * 64;
                    if (ncv == 64) {
#pragma unroll 16
                        for (int c = 0; c < 64; c++) acc += Ut[(int64_t)c * ldf] * svh[c];
                    } else {
                        for (int c = 0; c < ncv; c++) acc += Ut[(int64_t)c * ldf] * svh[c];
                    }
                }
                __syncthreads();
            }
            sacc[half][r] = acc;
            __syncthreads();
            if (tid < NB) sv[tid] = (c0 + tid < m) ? (x[c0 + tid] - sacc[0][tid] - sacc[1][tid]) : 0.0;
            __syncthreads();
            // x[r] = sum_{c >= r} Dinv[r, c] v[c]
            double a = 0.0;
            {
                const double* row = Dk + r + (int64_t)(half * 64) * NB;
                const double* svh = sv + half * 64;
#pragma unroll 16
                for (int c = 0; c < 64; c++) a += row[c * NB] * svh[c];
            }
            sacc[half][r] = a;
            __syncthreads();
            if (tid < NB && c0 + tid < m) x[c0 + tid] = sacc[0][tid] + sacc[1][tid];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&flags[1 + k], epoch);
    }
}

bool g_panel_attr_set = false;
void set_panel_attr() {
    if (g_panel_attr_set) return;
    CUDA_TRY(cudaFuncSetAttribute(panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    CUDA_TRY(cudaFuncSetAttribute(panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    CUDA_TRY(cudaFuncSetAttribute(chol_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    g_panel_attr_set = true;
}

}  // namespace

void hyp_potrf_upper(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    if (m <= 0) return;
    TimeScope ts(ctx, T_POTRF);
    set_panel_attr();
    CUDA_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
    int nblk = ceil_div(m, NB);
    for (int k = 0; k < nblk; k++) {
        int64_t k0 = (int64_t)k * NB;
        int64_t nb = std::min<int64_t>(NB, m - k0);
        panel_kernel<true><<<1, 1024, NB * LDU * 8, ctx->stream>>>(A, lda, m, k, d_dinv, d_info);
        ctx->launches++;
        int64_t rest = m - k0 - nb;
        if (rest > 0) {
            double* A12 = A + k0 + (k0 + nb) * lda;
            double* A22 = A + (k0 + nb) + (k0 + nb) * lda;
            const double* Dk = d_dinv + (int64_t)k * NB * NB;
            // U12 = U11^-T A12  (in place: every output tile depends on its own columns only)
            hyp_gemm_tn(ctx, Dk, NB, A12, lda, nb, nb, rest, A12, lda, 1.0, 0.0);
            // A22 -= U12' U12 (upper tiles)
            hyp_atb_upper(ctx, A12, lda, A12, lda, nb, rest, A22, lda, -1.0, 1.0);
        }
    }
    CUDA_TRY(cudaGetLastError());
}

void hyp_trtri_diag(hyp_ctx* ctx, const double* U, int64_t ldu, int64_t m, double* d_dinv) {
    if (m <= 0) return;
    set_panel_attr();
    int nblk = ceil_div(m, NB);
    panel_kernel<false><<<nblk, 1024, NB * LDU * 8, ctx->stream>>>(const_cast<double*>(U), ldu, m, 0,
                                                                d_dinv, nullptr);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_trsv_upper(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, const double* d_dinv,
                    double* x, bool trans) {
    if (m <= 0) return;
    TimeScope ts(ctx, T_TRSV);
    int nblk = ceil_div(m, NB);
    CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int epoch = ++ctx->trsv_epoch;
    int grid = std::min(nblk, 2 * ctx->sm_count);
    if (trans)
        trsv_kernel<true><<<grid, 256, 0, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch);
    else
        trsv_kernel<false><<<grid, 256, 0, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_chol_batched(hyp_ctx* ctx, int ncones, const int* d_sides, const int64_t* d_moff,
                      const int* d_kidx, double* U, double* Ui, uint8_t* d_flag) {
    if (ncones <= 0) return;
    set_panel_attr();
    chol_batched_kernel<<<ncones, 1024, NB * LDU * 8, ctx->stream>>>(ncones, d_sides, d_moff, d_kidx, U, Ui,
                                                                    d_flag);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}
