// Batched cone barrier oracles (K7-K9 of SURVEY.md section 2.3): cone table, vector cones
// (Nonnegative, EpiNormEucl) and the type dispatch.  Matrix cones live in cones_mat.cu.
//
// reference: src/Cones/Cones.jl:27-310 (generic API, check_numerics, get_proxsqr),
// nonnegative.jl:42-145, epinormeucl.jl:44-228.  The reference loops `for k in eachindex(cones)`
// on the host (e.g. qrchol.jl:214-246, search.jl:112-135); here one launch per cone TYPE covers
// every cone of that type, driven by a device cone table (offset, dim) - "warp per
// (cone, column)" for second-order cones, "thread per element" for the nonnegative orthant.
// All kernels are HBM-bound streaming kernels: algorithmic bytes = 16 * rows * ncols per product.
#include "common.cuh"
#include "cones_mat.cuh"
#include "cones_vec3_kernels.cuh"
#include "cones_vec_kernels.cuh"

static_assert(VK_HESS == HYP_PROD_HESS && VK_INV_HESS == HYP_PROD_INV_HESS && VK_SQRT_HESS == HYP_PROD_SQRT_HESS &&
                  VK_INV_SQRT_HESS == HYP_PROD_INV_SQRT_HESS && VK_NONNEGATIVE == HYP_CONE_NONNEGATIVE,
              "cones_vec_kernels.cuh constants must match the ABI");

namespace {

inline int grid_for(hyp_ctx* ctx, int64_t len, int threads, int mult = 8) {
    int64_t blocks = (len + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sm_count * mult;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    if (v.empty()) return d;
    CUDA_TRY(cudaMalloc(&d, v.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

// columns are dealt round-robin to gridDim.y CTAs per chunk: about four columns per CTA amortise the staging of the points
inline int soc_chunk_gy(int64_t ncols) {
    return (int)std::min<int64_t>(std::min<int64_t>(ncols, 65535), std::max<int64_t>(64, ncols / 4));
}

template <int MODE>
void launch_vec_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols,
                     int64_t ld_prod, int64_t ld_arr, int64_t row_shift) {
    int gy = (int)std::min<int64_t>(ncols, 65535);
    if (g.type == HYP_CONE_NONNEGATIVE) {
        int gx = grid_for(ctx, g.rows, 256, ncols > 1 ? 1 : 8);
        hypdev::nn_prod_kernel<MODE><<<dim3(gx, gy), 256, 0, ctx->stream>>>(
            g.rows, g.d_rows, ctx->d_point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
    } else if (ncols >= 16 && g.n_chunks > 0 && g.chunks_cover_all) {
        hypdev::soc_prod_chunk_kernel<MODE><<<dim3(g.n_chunks, soc_chunk_gy(ncols)), 256, g.chunk_smem, ctx->stream>>>(
            g.d_crow0, g.d_crows, g.d_ccone0, g.d_ccount, g.d_off, g.d_dim, g.d_scal, ctx->d_point, arr,
            ld_arr, prod, ld_prod, ncols, row_shift);
    } else {
        int gx = ceil_div(g.count, 8);
        hypdev::soc_prod_kernel<MODE><<<dim3(gx, gy), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_scal, ctx->d_point, arr, ld_arr, prod, ld_prod, ncols,
            row_shift);
    }
    ctx->launches++;
}

// EpiPerSquare / HypoPerLog / EpiNormInf: one warp per (cone, column); mode 4 / 5 = the block modes
void launch_v3_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                    int64_t ld_arr, int mode, int64_t row_shift) {
    if (g.type != HYP_CONE_EPIPERSQUARE && (mode == HYP_PROD_SQRT_HESS || mode == HYP_PROD_INV_SQRT_HESS))
        throw HypError{"sqrt_hess_prod is not defined for HypoPerLog / EpiNormInf"};
    dim3 grid(ceil_div(g.count, 8), (unsigned)std::min<int64_t>(ncols, 65535));
    hypdev::v3_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.type, mode, g.count, g.d_off, g.d_dim, g.d_dual, g.d_hkind,
                                                         g.d_hparam, g.d_scal,
                                                         ctx->d_point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
    ctx->launches++;
}

}  // namespace

void hyp_cones_build_groups(hyp_ctx* ctx) {
    hyp_cones_free_groups(ctx);
    for (int type = 0; type < HYP_NUM_CONE_TYPES; type++) {
        ConeGroup g;
        g.type = type;
        std::vector<int> rows, rowcone;
        for (int k = ctx->cone_lo; k < ctx->cone_hi; k++) {
            if (ctx->h_cone_type[k] != type) continue;
            int dim = (int)ctx->h_cone_dim[k];
            g.h_off.push_back(ctx->h_cone_off[k]);
            g.h_dim.push_back(dim);
            g.h_kidx.push_back(k);
            g.h_dual.push_back(ctx->h_cone_dual[k]);
            int side = 0;
            if (cone_is_matrix(type)) {
                int64_t len = dim - cone_mat_lead(type);
                side = (int)((std::sqrt(1.0 + 8.0 * (double)len) - 1.0) / 2.0 + 0.5);
                while ((int64_t)side * (side + 1) / 2 < len) side++;
                while ((int64_t)side * (side + 1) / 2 > len) side--;
                if ((int64_t)side * (side + 1) / 2 != len) throw HypError{"matrix cone dimension is not triangular"};
            }
            g.h_side.push_back(side);
            g.h_moff.push_back(g.mat_total);
            g.mat_total += round_up((int64_t)side * side, 2);
            g.max_dim = std::max(g.max_dim, dim);
            g.max_side = std::max(g.max_side, side);
            g.rows += dim;
            if (type == HYP_CONE_NONNEGATIVE)
                for (int i = 0; i < dim; i++) {
                    rows.push_back((int)(ctx->h_cone_off[k] + i));
                    rowcone.push_back(k);
                }
        }
        g.count = (int)g.h_off.size();
        if (g.count == 0) continue;
        g.d_off = upload(g.h_off);
        g.d_dim = upload(g.h_dim);
        g.d_kidx = upload(g.h_kidx);
        g.d_dual = upload(g.h_dual);
        g.d_side = upload(g.h_side);
        g.d_moff = upload(g.h_moff);
        CUDA_TRY(cudaMalloc(&g.d_scal, (size_t)g.count * 8 * sizeof(double)));
        CUDA_TRY(cudaMemset(g.d_scal, 0, (size_t)g.count * 8 * sizeof(double)));
        CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
        if (type == HYP_CONE_NONNEGATIVE) {
            g.d_rows = upload(rows);
            g.d_rowcone = upload(rowcone);
        }
        if (type == HYP_CONE_EPINORMEUCL) {
            // chunks of consecutive cones with at most CHUNK_ROWS rows (shared-memory staging)
            const int CHUNK_ROWS = hypdev::SOC_CHUNK_ROWS, CHUNK_CONES = 256;
            std::vector<int64_t> crow0;
            std::vector<int> crows, ccone0, ccount;
            g.chunks_cover_all = true;
            int i = 0;
            while (i < g.count) {
                if (g.h_dim[i] > CHUNK_ROWS) { g.chunks_cover_all = false; break; }
                int64_t r0 = g.h_off[i];
                int rows_c = 0, n_c = 0;
                while (i < g.count && n_c < CHUNK_CONES && rows_c + g.h_dim[i] <= CHUNK_ROWS &&
                       g.h_off[i] == r0 + rows_c) {
                    rows_c += g.h_dim[i];
                    n_c++;
                    i++;
                }
                crow0.push_back(r0);
                crows.push_back(rows_c);
                ccone0.push_back(i - n_c);
                ccount.push_back(n_c);
            }
            if (g.chunks_cover_all) {
                g.n_chunks = (int)crow0.size();
                g.chunk_smem = 2 * CHUNK_ROWS * (int)sizeof(double);   // the staged column and the staged points
                g.d_crow0 = upload(crow0);
                g.d_crows = upload(crows);
                g.d_ccone0 = upload(ccone0);
                g.d_ccount = upload(ccount);
            }
        }
        if (type == HYP_CONE_EPIPERSEPSPECTRAL_MAT) hyp_spec_alloc_group(ctx, g);
        else if (cone_is_matrix(type)) hyp_mat_alloc_group(ctx, g);
        if (cone_is_genfact(type)) hyp_gpow_alloc_group(ctx, g);
        if (type == HYP_CONE_EPIPERSEPSPECTRAL_VEC) {
            for (int kk : g.h_kidx) {
                g.h_hkind.push_back(ctx->h_cone_hkind[kk]);
                g.h_hparam.push_back(ctx->h_cone_hparam[kk]);
            }
            g.d_hkind = upload(g.h_hkind);
            g.d_hparam = upload(g.h_hparam);
        }
        ctx->groups.push_back(g);
    }
    CUDA_TRY(cudaDeviceSynchronize());   // uploads above ran on the legacy stream
}

void hyp_cones_free_groups(hyp_ctx* ctx) {
    for (auto& g : ctx->groups) {
        cudaFree(g.d_off);
        cudaFree(g.d_dim);
        cudaFree(g.d_kidx);
        cudaFree(g.d_dual);
        cudaFree(g.d_side);
        cudaFree(g.d_moff);
        cudaFree(g.d_scal);
        cudaFree(g.d_rows);
        cudaFree(g.d_rowcone);
        cudaFree(g.d_crow0);
        cudaFree(g.d_crows);
        cudaFree(g.d_ccone0);
        cudaFree(g.d_ccount);
        cudaFree(g.d_W);
        cudaFree(g.d_U);
        cudaFree(g.d_Ut);
        cudaFree(g.d_Ui);
        cudaFree(g.d_Uit);
        cudaFree(g.d_Wi);
        cudaFree(g.d_hkind);
        cudaFree(g.d_hparam);
        cudaFree(g.d_voff);
        cudaFree(g.d_voff7);
        cudaFree(g.d_vecs);
    }
    ctx->groups.clear();
}

// feas / grad / per-cone factorisations for the points in ctx->d_point / d_dual
void hyp_cones_update_state(hyp_ctx* ctx) {
    TimeScope ts(ctx, T_CONE_STATE);
    if (ctx->K == 0) return;
    CUDA_TRY(cudaMemsetAsync(ctx->d_feas, 1, ctx->K, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_dual_feas, 1, ctx->K, ctx->stream));
    for (auto& g : ctx->groups) {
        if (g.type == HYP_CONE_NONNEGATIVE) {
            hypdev::nn_state_kernel<<<grid_for(ctx, g.rows, 256), 256, 0, ctx->stream>>>(
                g.rows, g.d_rows, g.d_rowcone, ctx->d_point, ctx->d_dual, ctx->d_grad, ctx->d_feas,
                ctx->d_dual_feas);
            ctx->launches++;
        } else if (g.type == HYP_CONE_EPINORMEUCL) {
            hypdev::soc_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
                g.count, g.d_off, g.d_dim, g.d_kidx, ctx->d_point, ctx->d_dual, ctx->d_grad, g.d_scal,
                ctx->d_feas, ctx->d_dual_feas);
            ctx->launches++;
        } else if (cone_is_genfact(g.type)) {
            hyp_gpow_update_state(ctx, g);
        } else if (cone_is_vec3(g.type)) {
            hypdev::v3_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
                g.type, g.count, g.d_off, g.d_dim, g.d_kidx, g.d_hkind, g.d_hparam, ctx->d_point, ctx->d_dual, ctx->d_grad,
                g.d_scal, ctx->d_feas, ctx->d_dual_feas);
            ctx->launches++;
        } else if (g.type == HYP_CONE_EPIPERSEPSPECTRAL_MAT) {
            hyp_spec_update_state(ctx, g);
        } else {
            hyp_mat_update_state(ctx, g);
        }
    }
    CUDA_TRY(cudaGetLastError());
    if (hyp_row_sharded(ctx)) {
        hyp_allreduce_min_u8(ctx, ctx->d_feas, ctx->K);
        hyp_allreduce_min_u8(ctx, ctx->d_dual_feas, ctx->K);
        hyp_replicate_q(ctx, ctx->d_grad);
    }
}

static int resolve_mode(const ConeGroup& g, int k_in_group, int mode) {
    (void)k_in_group;
    return mode;
}

void hyp_cones_prod(hyp_ctx* ctx, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                    int64_t ld_arr, int mode, int64_t row_shift) {
    if (ncols <= 0) return;
    for (auto& g : ctx->groups) {
        int m = resolve_mode(g, 0, mode);
        if (g.type <= HYP_CONE_EPINORMEUCL) {
            if (m == HYP_PROD_BLOCK) m = HYP_PROD_HESS;   // no dual-barrier variants of these cones
            if (m == HYP_PROD_BLOCK_INV) m = HYP_PROD_INV_HESS;
            switch (m) {
                case HYP_PROD_HESS:
                    launch_vec_prod<HYP_PROD_HESS>(ctx, g, prod, arr, ncols, ld_prod, ld_arr, row_shift);
                    break;
                case HYP_PROD_INV_HESS:
                    launch_vec_prod<HYP_PROD_INV_HESS>(ctx, g, prod, arr, ncols, ld_prod, ld_arr, row_shift);
                    break;
                case HYP_PROD_SQRT_HESS:
                    launch_vec_prod<HYP_PROD_SQRT_HESS>(ctx, g, prod, arr, ncols, ld_prod, ld_arr, row_shift);
                    break;
                case HYP_PROD_INV_SQRT_HESS:
                    launch_vec_prod<HYP_PROD_INV_SQRT_HESS>(ctx, g, prod, arr, ncols, ld_prod, ld_arr, row_shift);
                    break;
                default:
                    throw HypError{"hyp_cones_prod: bad mode"};
            }
        } else if (cone_is_genfact(g.type)) {
            hyp_gpow_prod(ctx, g, prod, arr, ncols, ld_prod, ld_arr, m, row_shift);
        } else if (cone_is_vec3(g.type)) {
            launch_v3_prod(ctx, g, prod, arr, ncols, ld_prod, ld_arr, m, row_shift);
        } else if (g.type == HYP_CONE_EPIPERSEPSPECTRAL_MAT) {
            hyp_spec_prod(ctx, g, prod, arr, ncols, ld_prod, ld_arr, m, row_shift);
        } else {
            hyp_mat_prod(ctx, g, prod, arr, ncols, ld_prod, ld_arr, m, row_shift);
        }
    }
    CUDA_TRY(cudaGetLastError());
}

// Schur pre-pass (qrchol.jl:219-246): HG_k = H_k^{1/2} GQ2_k for cones with closed-form square
// roots, HG_k = H_k GQ2_k (block_hess_prod!) for the log-det family.
void hyp_cones_schur_prepass(hyp_ctx* ctx) {
    TimeScope ts(ctx, T_SQRT_PREPASS);
    const double* GQ2 = ctx->d_GQ + ctx->p * ctx->ldg;
    for (auto& g : ctx->groups) {
        if (g.type <= HYP_CONE_EPINORMEUCL) {
            launch_vec_prod<HYP_PROD_SQRT_HESS>(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg,
                                                ctx->row_lo);
        } else if (cone_is_genfact(g.type)) {
            hyp_gpow_prod(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg, HYP_PROD_BLOCK, ctx->row_lo);
        } else if (cone_is_vec3(g.type)) {
            launch_v3_prod(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg,
                           g.type == HYP_CONE_EPIPERSQUARE ? HYP_PROD_SQRT_HESS : HYP_PROD_BLOCK, ctx->row_lo);
        } else if (g.type == HYP_CONE_POSSEMIDEFTRI) {
            hyp_mat_prod(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg, HYP_PROD_SQRT_HESS,
                         ctx->row_lo);
        } else if (g.type == HYP_CONE_EPIPERSEPSPECTRAL_MAT) {
            hyp_spec_prod(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg, HYP_PROD_BLOCK, ctx->row_lo);
        } else {
            hyp_mat_prod(ctx, g, ctx->d_HG, GQ2, ctx->nmp, ctx->ldg, ctx->ldg, HYP_PROD_BLOCK,
                         ctx->row_lo);
        }
    }
    CUDA_TRY(cudaGetLastError());
}

// Fused pre-pass of the digit-sliced Schur SYRK for models whose local cones are all EpiNormEucl (BASELINE config 3):
// the rows of H^{1/2} G are never written - pass 1 recomputes them chunk by chunk in shared memory and keeps the
// column maxima, pass 2 recomputes them and stores the radix-256 digit slices directly (qrchol.jl:219-234 + the
// slicing of ozaki.cu in two reads of G and one write of the digits: 11.5 GB instead of 19.5 GB on C3).
int hyp_cones_prepass_sliced(hyp_ctx* ctx, int8_t* digits, int64_t ldd, int64_t slice_stride, int* expo, double* dscale) {
    // returns 0 = not applicable (caller runs the plain pre-pass and the slicer), 1 = digit slices written (H^{1/2} G never
    // stored; opt-in), 2 = H^{1/2} G stored and the column exponents / scales ready (default: the slicer skips its
    // column-maximum pass over the 4 GB product)
    // full fusion is opt-in (HYP_FUSED_PREPASS=1): measured on C3 (profiles/r02_bench_c3_pair64_{fused,nofuse}.json) the two
    // recomputing passes cost 7.6 ms against 3.0 ms (pre-pass) + 3.4 ms (colmax + slice256) of the three-kernel sequence -
    // the second-order-cone product kernel is bound by its shared-memory sweeps, not by HBM, so computing it twice loses
    static int fused = -1, with_max = -1;
    if (fused < 0) {
        fused = getenv("HYP_FUSED_PREPASS") ? 1 : 0;
        with_max = getenv("HYP_NO_PREPASS_COLMAX") ? 0 : 1;
    }
    if ((!fused && !with_max) || hyp_ozaki_radix() != 256 || ctx->groups.size() != 1 || ctx->nmp < 16) return 0;
    ConeGroup& g = ctx->groups[0];
    if (g.type != HYP_CONE_EPINORMEUCL || g.n_chunks <= 0 || !g.chunks_cover_all || g.rows != ctx->qloc) return 0;
    const double* GQ2 = ctx->d_GQ + ctx->p * ctx->ldg;
    const int64_t ncols = ctx->nmp;
    const int gy = soc_chunk_gy(ncols);
    if (!ctx->d_colbits) CUDA_TRY(cudaMalloc((void**)&ctx->d_colbits, (size_t)std::max<int64_t>(ncols, 1) * 8));
    if (!fused) {
        TimeScope ts(ctx, T_SQRT_PREPASS);
        CUDA_TRY(cudaMemsetAsync(ctx->d_colbits, 0, (size_t)ncols * 8, ctx->stream));
        hypdev::soc_prod_chunk_kernel<HYP_PROD_SQRT_HESS, 3><<<dim3(g.n_chunks, gy), 256, g.chunk_smem, ctx->stream>>>(
            g.d_crow0, g.d_crows, g.d_ccone0, g.d_ccount, g.d_off, g.d_dim, g.d_scal, ctx->d_point, GQ2, ctx->ldg, ctx->d_HG,
            ctx->ldg, ncols, ctx->row_lo, ctx->d_colbits);
        hypdev::expo_from_bits_kernel<<<std::max(1, std::min(ceil_div(ncols, 256), ctx->sm_count)), 256, 0, ctx->stream>>>(
            ncols, ctx->d_colbits, expo, dscale, 1);
        ctx->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return 2;
    }
    // chunk starts (local row numbers) must be multiples of 8: the digit words are 8 rows wide
    if (g.chunk_align8 < 0) {
        g.chunk_align8 = 1;
        int64_t expect = ctx->row_lo;
        for (size_t i = 0, c = 0; i < g.h_off.size(); c++) {
            // re-walk the chunk construction of hyp_cones_build_groups
            int64_t r0 = g.h_off[i];
            int rows_c = 0, n_c = 0;
            while (i < g.h_off.size() && n_c < 256 && rows_c + g.h_dim[i] <= hypdev::SOC_CHUNK_ROWS && g.h_off[i] == r0 + rows_c) {
                rows_c += g.h_dim[i];
                n_c++;
                i++;
            }
            if ((r0 - ctx->row_lo) % 8 != 0 || r0 != expect) g.chunk_align8 = 0;
            expect = r0 + rows_c;
        }
    }
    if (!g.chunk_align8) return 0;
    TimeScope ts(ctx, T_SQRT_PREPASS);
    CUDA_TRY(cudaMemsetAsync(ctx->d_colbits, 0, (size_t)ncols * 8, ctx->stream));
    hypdev::soc_prod_chunk_kernel<HYP_PROD_SQRT_HESS, 1><<<dim3(g.n_chunks, gy), 256, g.chunk_smem, ctx->stream>>>(
        g.d_crow0, g.d_crows, g.d_ccone0, g.d_ccount, g.d_off, g.d_dim, g.d_scal, ctx->d_point, GQ2, ctx->ldg, nullptr, 0,
        ncols, ctx->row_lo, ctx->d_colbits);
    hypdev::expo_from_bits_kernel<<<std::max(1, std::min(ceil_div(ncols, 256), ctx->sm_count)), 256, 0, ctx->stream>>>(
        ncols, ctx->d_colbits, expo, dscale, 1);
    hypdev::soc_prod_chunk_kernel<HYP_PROD_SQRT_HESS, 2><<<dim3(g.n_chunks, gy), 256, g.chunk_smem, ctx->stream>>>(
        g.d_crow0, g.d_crows, g.d_ccone0, g.d_ccount, g.d_off, g.d_dim, g.d_scal, ctx->d_point, GQ2, ctx->ldg, nullptr, 0,
        ncols, ctx->row_lo, nullptr, expo, digits, ldd, slice_stride, 7);
    ctx->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return 1;
}

void hyp_cones_dder3_dev(hyp_ctx* ctx, double* out, const double* dir) {
    TimeScope ts(ctx, T_CONE_PROD);
    for (auto& g : ctx->groups) {
        if (g.type == HYP_CONE_NONNEGATIVE) {
            hypdev::nn_dder3_kernel<<<grid_for(ctx, g.rows, 256), 256, 0, ctx->stream>>>(
                g.rows, g.d_rows, ctx->d_point, dir, out);
            ctx->launches++;
        } else if (g.type == HYP_CONE_EPINORMEUCL) {
            hypdev::soc_dder3_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
                g.count, g.d_off, g.d_dim, g.d_scal, ctx->d_point, dir, out);
            ctx->launches++;
        } else if (cone_is_genfact(g.type)) {
            hyp_gpow_dder3(ctx, g, out, dir);
        } else if (cone_is_vec3(g.type)) {
            hypdev::v3_dder3_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
                g.type, g.count, g.d_off, g.d_dim, g.d_hkind, g.d_hparam, g.d_scal, ctx->d_point, dir, out);
            ctx->launches++;
        } else if (g.type == HYP_CONE_EPIPERSEPSPECTRAL_MAT) {
            hyp_spec_dder3(ctx, g, out, dir);
        } else {
            hyp_mat_dder3(ctx, g, out, dir);
        }
    }
    CUDA_TRY(cudaGetLastError());
}

// check_numerics + get_proxsqr for all local cones -> ctx->d_proxsqr / d_num_ok (global K)
void hyp_cones_prox_dev(hyp_ctx* ctx, double irtmu, int use_max) {
    TimeScope ts(ctx, T_CONE_PROD);
    if (ctx->K == 0) return;
    int nloc = ctx->cone_hi - ctx->cone_lo;
    CUDA_TRY(cudaMemsetAsync(ctx->d_proxsqr, 0, ctx->K * sizeof(double), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_num_ok, 1, ctx->K, ctx->stream));
    if (nloc > 0) {
        // v1 = irtmu * dual + grad ; v2 = Hinv v1 ; v3 = Hinv grad
        hyp_lincomb3(ctx, ctx->q, ctx->d_vq1, irtmu, ctx->d_dual, 1.0, ctx->d_grad, 0.0, nullptr);
        hyp_cones_prod(ctx, ctx->d_vq2, ctx->d_vq1, 1, ctx->q, ctx->q, HYP_PROD_INV_HESS, 0);
        hyp_cones_prod(ctx, ctx->d_vq3, ctx->d_grad, 1, ctx->q, ctx->q, HYP_PROD_INV_HESS, 0);
        hypdev::cone_prox_kernel<<<nloc, 128, 0, ctx->stream>>>(
            ctx->cone_lo, ctx->d_cone_type, ctx->d_cone_off, ctx->d_cone_dim, ctx->d_cone_nu,
            ctx->d_point, ctx->d_dual, ctx->d_grad, ctx->d_vq1, ctx->d_vq2, ctx->d_vq3, irtmu, use_max,
            ctx->d_proxsqr, ctx->d_num_ok);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    if (hyp_row_sharded(ctx)) {
        hyp_allreduce_sum(ctx, ctx->d_proxsqr, ctx->K);
        hyp_allreduce_min_u8(ctx, ctx->d_num_ok, ctx->K);
    }
}
