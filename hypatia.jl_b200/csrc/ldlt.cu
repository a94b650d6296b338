// Rook-pivoted symmetric-indefinite factorisation A = L D L' (Bunch-Kaufman, 1x1 and 2x2
// pivots) and solve: the fallback of the reference's posdef_fact_copy! chain (K4 of SURVEY.md
// section 2.3).
//
// reference: symm_fact!(A) = bunchkaufman!(A, true, check=false) (src/linearalgebra/dense.jl:
// 164-165, LAPACK dsytrf_rook) reached from posdef_fact_copy! (dense.jl:194-215) when the Cholesky
// factorisation of the Schur complement fails; increase_diag! (dense.jl:106-113).
//
// This is an exceptional path (the Schur complement of an interior iterate is positive definite),
// so the design favours being device-resident and simple over speed: an unblocked right-looking
// factorisation on FULL symmetric storage with all pivoting decisions taken on the device.  The
// host enqueues, without ever synchronising, one five-kernel sequence per column (pivot search by
// one CTA, two symmetric interchanges, pivot-column scaling, rank-1 / rank-2 trailing update); a
// device-side cursor makes the sequences left over after 2x2 pivots no-ops.  Pivot selection is the
// rook search of LAPACK's dsytf2_rook (alpha = (1 + sqrt 17) / 8).  Bound: HBM (the trailing update
// streams the trailing matrix once per column: ~16 m^3 / 3 bytes in total).
#include "common.cuh"
#include "ldlt_kernels.cuh"

using namespace hypdev;   // the kernels and the ST_* cursor layout

void hyp_increase_diag(hyp_ctx* ctx, double* A, int64_t lda, int64_t m) {
    if (m <= 0) return;
    increase_diag_kernel<<<std::max(1, std::min(ceil_div(m, 256), ctx->sm_count)), 256, 0, ctx->stream>>>(A, lda, m);
    ctx->launches++;
}

// A: upper triangle of the symmetric matrix on entry (lower part arbitrary); on exit L (strictly
// lower) and D (diagonal / sub-diagonal of 2x2 blocks).  d_ipiv: 3 m ints.  d_info[0] = 0 or the
// 1-based index of the first exactly-singular pivot column.  Needs ctx->d_ldl_work (4 m doubles).
void hyp_ldlt_factor(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, int* d_ipiv, int* d_info) {
    if (m <= 0) return;
    int* st = d_info + 1;   // ST_NUM ints of cursor state live behind the info word
    init_state_kernel<<<1, 32, 0, ctx->stream>>>(st);
    dim3 sg(std::max(1, std::min(ceil_div(m, 256), 8)), (unsigned)std::min<int64_t>(m, 65535));
    symmetrize_kernel<<<sg, 256, 0, ctx->stream>>>(A, lda, m);
    ctx->launches += 2;
    const int vg = std::max(1, std::min(ceil_div(m, 256), ctx->sm_count));
    for (int64_t seq = 0; seq < m; seq++) {
        // grid sizes follow the largest trailing block this sequence can see (columns >= seq / 2 ... )
        int64_t rem = m - seq / 2;   // the cursor is at least seq / 2 ... at most seq
        pivot_kernel<<<1, 1024, 0, ctx->stream>>>(A, lda, m, st, d_ipiv);
        swap_kernel<<<vg, 256, 0, ctx->stream>>>(A, lda, m, st, 0);
        swap_kernel<<<vg, 256, 0, ctx->stream>>>(A, lda, m, st, 1);
        colprep_kernel<<<vg, 256, 0, ctx->stream>>>(A, lda, m, st, ctx->d_ldl_work);
        dim3 ug(std::max(1, std::min(ceil_div(rem, 256), 16)),
                (unsigned)std::max<int64_t>(1, std::min<int64_t>(rem, 4 * ctx->sm_count)));
        update_kernel<<<ug, 256, 0, ctx->stream>>>(A, lda, m, st, ctx->d_ldl_work);
        ctx->launches += 5;
    }
    finish_info_kernel<<<1, 32, 0, ctx->stream>>>(st, d_info);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_ldlt_solve(hyp_ctx* ctx, const double* A, int64_t lda, int64_t m, const int* d_ipiv, double* x) {
    if (m <= 0) return;
    ldlt_solve_kernel<<<1, 1024, 0, ctx->stream>>>(A, lda, m, d_ipiv, x);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}
