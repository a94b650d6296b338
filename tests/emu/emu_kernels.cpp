// Host-emulated launches of the kernel headers (test infrastructure, CPU-only test tier): the very
// same __global__ functions nvcc compiles for sm_100a, run with one pthread per CUDA thread.
#define HYP_EMU 1
#include "cuda_emu.h"
#include "../../hypatia.jl_b200/csrc/eig_kernels.cuh"

extern "C" {

// eigen-decomposition of nmat matrices through syevj_batched_kernel; smem = 1: shared-memory
// variant (emulated dynamic smem), 0: global-scratch variant
int emu_syevj(int nmat, int max_side, const int* sides, const int64_t* in_off, const double* Ain, double* Vout,
              const int64_t* lam_off, double* lam, const double* divv, const int64_t* div_off, int div_idx,
              int smem, int threads) {
    const bool wantv = Vout != nullptr;
    const int64_t wd = hypdev::syevj_work_doubles(max_side, true);
    std::vector<double> gwork((size_t)wd * nmat + 8);
    auto run = [&](auto kern, size_t dyn) {
        emu::launch(dim3(nmat), dim3(threads), dyn, [&] {
            kern(nmat, sides, in_off, Ain, Vout, lam_off, lam, divv, div_off, div_idx, gwork.data(), wd);
        });
    };
    if (smem) {
        if (wantv) run(hypdev::syevj_batched_kernel<true, true>, wd * 8);
        else run(hypdev::syevj_batched_kernel<true, false>, wd * 8);
    } else {
        if (wantv) run(hypdev::syevj_batched_kernel<false, true>, 0);
        else run(hypdev::syevj_batched_kernel<false, false>, 0);
    }
    return 0;
}

}  // extern "C"

#include "../../hypatia.jl_b200/csrc/cones_mat_kernels.cuh"
#include "../../hypatia.jl_b200/csrc/cones_spec_kernels.cuh"

extern "C" {

int emu_unpack_state(int ncones, const int64_t* off, const int* sides, const int64_t* moff, int lead,
                     const double* vec, double* A, double* B, int gy) {
    emu::launch(dim3(ncones, gy), dim3(64), 0,
                [&] { hypdev::unpack_state_kernel(ncones, off, sides, moff, lead, vec, A, B); });
    return 0;
}

int emu_unpack_cols(int d, int lde, int64_t len, const double* arr, int64_t ld_arr, int64_t cc, double* Mall,
                    int gx) {
    emu::launch(dim3(gx, (unsigned)cc), dim3(64), 0,
                [&] { hypdev::unpack_cols_kernel(d, lde, len, arr, ld_arr, cc, Mall); });
    return 0;
}

int emu_pack_cols(int d, int lde, int64_t len, const double* Yall, int64_t cc, const double* alpha,
                  const double* beta, const double* vecB, double* prod, int64_t ld_prod, int gx) {
    emu::launch(dim3(gx, (unsigned)cc), dim3(64), 0,
                [&] { hypdev::pack_cols_kernel(d, lde, len, Yall, cc, alpha, beta, vecB, prod, ld_prod); });
    return 0;
}

int emu_spec_post(int ncones, const int64_t* off, const int* sides, const int64_t* moff, const int64_t* voff,
                  const int* kidx, const int* hkind, const double* hparam, const double* point, const double* V,
                  double* Vt, double* theta, double* Dh, double* vecs, double* scal, double* grad, uint8_t* feas,
                  int threads) {
    emu::launch(dim3(ncones), dim3(threads), 0, [&] {
        hypdev::spec_post_kernel(ncones, off, sides, moff, voff, kidx, hkind, hparam, point, V, Vt, theta, Dh, vecs,
                                 scal, grad, feas);
    });
    return 0;
}

int emu_spec_dualfeas(int ncones, const int64_t* off, const int* sides, const int64_t* lam_off, const int* kidx,
                      const int* hkind, const double* hparam, const double* dual, const double* lamd,
                      const uint8_t* chol_ok, uint8_t* dual_feas) {
    emu::launch(dim3(ncones), dim3(64), 0, [&] {
        hypdev::spec_dualfeas_kernel(ncones, off, sides, lam_off, kidx, hkind, hparam, dual, lamd, chol_ok,
                                     dual_feas);
    });
    return 0;
}

int emu_spec_mid(int inverse, int d, int lde, const double* sc, const double* vecs, const double* theta,
                 const double* Dh, double* Mall, const double* arr, int64_t ld_arr, double* pr, int64_t ld_prod,
                 int64_t cc, int grid, int threads) {
    emu::launch(dim3(grid), dim3(threads), 0, [&] {
        hypdev::spec_mid_kernel(inverse, d, lde, sc, vecs, theta, Dh, Mall, arr, ld_arr, pr, ld_prod, cc);
    });
    return 0;
}

int emu_spec_dder3(int d, int lde, const double* sc, const double* vecs, const double* Dh, const double* R,
                   double* X, double* OUT, const double* dir, double* out, int threads) {
    emu::launch(dim3(1), dim3(threads), 0,
                [&] { hypdev::spec_dder3_kernel(d, lde, sc, vecs, Dh, R, X, OUT, dir, out); });
    return 0;
}

}  // extern "C"

#include "../../hypatia.jl_b200/csrc/cones_vec3_kernels.cuh"

extern "C" {

int emu_v3_state(int type, int ncones, const int64_t* off, const int* dim, const int* kidx, const int* hkind,
                 const double* hparam, const double* point, const double* dual, double* grad, double* scal,
                 uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::v3_state_kernel(type, ncones, off, dim, kidx, hkind, hparam, point, dual, grad, scal, feas, dual_feas);
    });
    return 0;
}

int emu_v3_prod(int type, int mode, int ncones, const int64_t* off, const int* dim, const int* dualf,
                const int* hkind, const double* hparam, const double* scal, const double* point, const double* arr, int64_t ld_arr, double* prod,
                int64_t ld_prod, int64_t ncols, int64_t row_shift, int gy) {
    emu::launch(dim3((ncones + 1) / 2, gy), dim3(64), 0, [&] {
        hypdev::v3_prod_kernel(type, mode, ncones, off, dim, dualf, hkind, hparam, scal, point, arr, ld_arr, prod,
                               ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_v3_dder3(int type, int ncones, const int64_t* off, const int* dim, const int* hkind, const double* hparam,
                 const double* scal, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::v3_dder3_kernel(type, ncones, off, dim, hkind, hparam, scal, point, dir, out); });
    return 0;
}

}  // extern "C"

extern "C" {

int emu_mat_small_prod(int type, int mode, int ncones, int max_side, const int64_t* off, const int* sides,
                       const int64_t* moff, const int* dualf, const double* W, const double* Wi, const double* Ui,
                       const double* Ut, const double* scal, const double* point, const double* wivec,
                       const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t row_shift,
                       int ncols, int threads) {
    const size_t smem = (size_t)2 * max_side * (max_side | 1) * sizeof(double);
    emu::launch(dim3(ncones, ncols), dim3(threads), smem, [&] {
        hypdev::mat_small_prod_kernel(type, mode, ncones, off, sides, moff, dualf, W, Wi, Ui, Ut, scal, point, wivec,
                                      arr, ld_arr, prod, ld_prod, row_shift);
    });
    return 0;
}

}  // extern "C"

extern "C" {

int emu_spec_small_prod(int mode, int ncones, int max_side, const int64_t* off, const int* sides,
                        const int64_t* moff, const int64_t* voff, const int* dualf, const double* V, const double* Vt,
                        const double* theta, const double* Dh, const double* vecs, const double* scal,
                        const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t row_shift, int ncols,
                        int threads) {
    const size_t smem = (size_t)2 * max_side * (max_side | 1) * sizeof(double);
    emu::launch(dim3(ncones, ncols), dim3(threads), smem, [&] {
        hypdev::spec_small_prod_kernel(mode, ncones, off, sides, moff, voff, dualf, V, Vt, theta, Dh, vecs, scal, arr,
                                       ld_arr, prod, ld_prod, row_shift);
    });
    return 0;
}

}  // extern "C"

#include "../../hypatia.jl_b200/csrc/gemv_kernels.cuh"

extern "C" {

// w = alphaN * M x + betaN * w ; y = alphaT * M' z + betaT * y through the fused one-pass kernel
int emu_gemv_nt(int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x, const double* z,
                int nchunks, double alphaN, double betaN, double* w, double alphaT, double betaT, double* y) {
    const int rb = (int)((rows + 255) / 256);
    const int64_t cpc = (ncols + nchunks - 1) / nchunks;
    nchunks = (int)((ncols + cpc - 1) / cpc);
    std::vector<double> pN((size_t)nchunks * rows + 2), pT((size_t)rb * 4 * ncols + 2);
    emu::launch(dim3(rb, nchunks), dim3(128), 0,
                [&] { hypdev::gemv_nt_kernel(rows, ncols, M, ld, x, z, cpc, pN.data(), pT.data()); });
    emu::launch(dim3(2), dim3(64), 0,
                [&] { hypdev::gemv_n_reduce2_kernel(rows, nchunks, pN.data(), alphaN, betaN, w); });
    emu::launch(dim3(2), dim3(64), 0,
                [&] { hypdev::gemv_t_reduce_kernel(ncols, rb * 4, pT.data(), alphaT, betaT, y); });
    return 0;
}

}  // extern "C"

extern "C" {

int emu_mat_small_dder3(int type, int ncones, int max_side, const int64_t* off, const int* sides, const int64_t* moff,
                        const double* Ui, const double* Uit, const double* scal, const double* dir, double* out,
                        int threads) {
    const size_t smem = (size_t)2 * max_side * (max_side | 1) * sizeof(double);
    emu::launch(dim3(ncones), dim3(threads), smem, [&] {
        hypdev::mat_small_dder3_kernel(type, ncones, off, sides, moff, Ui, Uit, scal, dir, out);
    });
    return 0;
}

}  // extern "C"

#include "../../hypatia.jl_b200/csrc/cones_gpow_kernels.cuh"

extern "C" {

int emu_gpow_state(int ncones, const int64_t* off, const int* dim, const int* mu, const int64_t* aoff,
                   const double* alpha, const int* kidx, const int64_t* moff, const double* point, const double* dual,
                   double* grad, double* scal, double* H, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::gpow_state_kernel(ncones, off, dim, mu, aoff, alpha, kidx, moff, point, dual, grad, scal, H, feas,
                                  dual_feas);
    });
    return 0;
}

int emu_gpow_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int* mu, const int64_t* aoff,
                  const double* alpha, const int* dualf, const double* scal, const double* point, const double* arr,
                  int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::gpow_prod_kernel(ncones, want_dual, off, dim, mu, aoff, alpha, dualf, scal, point, arr, ld_arr, prod,
                                 ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_gen_invhess_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int64_t* moff,
                         const int* dualf, const double* Ui, const double* arr, int64_t ld_arr, double* prod,
                         int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::gen_invhess_prod_kernel(ncones, want_dual, off, dim, moff, dualf, Ui, arr, ld_arr, prod, ld_prod, ncols,
                                        row_shift);
    });
    return 0;
}

int emu_gpow_dder3(int ncones, const int64_t* off, const int* dim, const int* mu, const int64_t* aoff,
                   const double* alpha, const double* scal, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::gpow_dder3_kernel(ncones, off, dim, mu, aoff, alpha, scal, point, dir, out); });
    return 0;
}

}  // extern "C"

extern "C" {

int emu_etr_state(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const int* kidx,
                  const int64_t* moff, const double* point, double* grad, double* scal, double* H, uint8_t* feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::etr_state_kernel(ncones, off, dim, voff, vecs, kidx, moff, point, grad, scal, H, feas);
    });
    return 0;
}

int emu_etr_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int64_t* voff, const double* vecs,
                 const int* dualf, const double* scal, const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod,
                 int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::etr_prod_kernel(ncones, want_dual, off, dim, voff, vecs, dualf, scal, arr, ld_arr, prod, ld_prod, ncols,
                                row_shift);
    });
    return 0;
}

int emu_etr_dder3(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const double* scal,
                  const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::etr_dder3_kernel(ncones, off, dim, voff, vecs, scal, dir, out); });
    return 0;
}

int emu_sps_state(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const int* kidx,
                  const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::sps_state_kernel(ncones, off, dim, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_sps_dder3(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const double* dir,
                  double* out) {
    emu::launch(dim3(ncones), dim3(256), 0, [&] { hypdev::sps_dder3_kernel(ncones, off, dim, voff, vecs, dir, out); });
    return 0;
}

int emu_wone_state(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const int* kidx, const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::wone_state_kernel(ncones, off, dim, Rs, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_wone_dder3(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const double* dir, double* out) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::wone_dder3_kernel(ncones, off, dim, Rs, voff, vecs, dir, out); });
    return 0;
}

int emu_weuc_state(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const int* kidx, const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::weuc_state_kernel(ncones, off, dim, Rs, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_weuc_dder3(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const double* dir, double* out) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::weuc_dder3_kernel(ncones, off, dim, Rs, voff, vecs, dir, out); });
    return 0;
}

int emu_wpsd_state(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const int* kidx, const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::wpsd_state_kernel(ncones, off, dim, Rs, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_wpsd_dder3(int ncones, const int64_t* off, const int* dim, const int* Rs, const int64_t* voff, double* vecs,
                   const double* dir, double* out) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::wpsd_dder3_kernel(ncones, off, dim, Rs, voff, vecs, dir, out); });
    return 0;
}

int emu_mep_state(int ncones, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff, double* vecs,
                  const int* kidx, const int64_t* moff, const double* point, const double* dual, double* grad,
                  double* scal, double* H, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::mep_state_kernel(ncones, off, dim, d1s, voff, vecs, kidx, moff, point, dual, grad, scal, H, feas,
                                 dual_feas);
    });
    return 0;
}

int emu_mep_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff,
                 const double* vecs, const int* dualf, const double* scal, const double* point, const double* arr,
                 int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::mep_prod_kernel(ncones, want_dual, off, dim, d1s, voff, vecs, dualf, scal, point, arr, ld_arr, prod,
                                ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_mep_dder3(int ncones, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff, double* vecs,
                  const double* scal, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::mep_dder3_kernel(ncones, off, dim, d1s, voff, vecs, scal, point, dir, out); });
    return 0;
}

int emu_dnn_state(int ncones, const int64_t* off, const int* dim, const int* sides, const int64_t* voff, double* vecs,
                  const int* kidx, const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::dnn_state_kernel(ncones, off, dim, sides, voff, vecs, kidx, moff, point, grad, H, feas);
    });
    return 0;
}

int emu_dnn_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int* sides, const int64_t* voff,
                 const double* vecs, const int* dualf, const double* point, const double* arr, int64_t ld_arr,
                 double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::dnn_prod_kernel(ncones, want_dual, off, dim, sides, voff, vecs, dualf, point, arr, ld_arr, prod, ld_prod,
                                ncols, row_shift);
    });
    return 0;
}

int emu_dnn_dder3(int ncones, const int64_t* off, const int* dim, const int* sides, const int64_t* voff,
                  const double* vecs, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::dnn_dder3_kernel(ncones, off, dim, sides, voff, vecs, point, dir, out); });
    return 0;
}

int emu_lmi_state(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const int* kidx,
                  const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::lmi_state_kernel(ncones, off, dim, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_lmi_dder3(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const double* dir,
                  double* out) {
    emu::launch(dim3(ncones), dim3(256), 0, [&] { hypdev::lmi_dder3_kernel(ncones, off, dim, voff, vecs, dir, out); });
    return 0;
}

int emu_wsos_state(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const int* kidx,
                   const int64_t* moff, const double* point, double* grad, double* H, uint8_t* feas) {
    emu::launch(dim3(ncones), dim3(256), 0,
                [&] { hypdev::wsos_state_kernel(ncones, off, dim, voff, vecs, kidx, moff, point, grad, H, feas); });
    return 0;
}

int emu_wsos_dder3(int ncones, const int64_t* off, const int* dim, const int64_t* voff, double* vecs, const double* dir,
                   double* out) {
    emu::launch(dim3(ncones), dim3(256), 0, [&] { hypdev::wsos_dder3_kernel(ncones, off, dim, voff, vecs, dir, out); });
    return 0;
}

int emu_gen_hess_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int64_t* moff,
                      const int* dualf, const double* H, const double* arr, int64_t ld_arr, double* prod,
                      int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::gen_hess_prod_kernel(ncones, want_dual, off, dim, moff, dualf, H, arr, ld_arr, prod, ld_prod, ncols,
                                     row_shift);
    });
    return 0;
}

int emu_ens_state(int ncones, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff, double* vecs,
                  const int* kidx, const int64_t* moff, const double* point, const double* dual, double* grad,
                  double* scal, double* H, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::ens_state_kernel(ncones, off, dim, d1s, voff, vecs, kidx, moff, point, dual, grad, scal, H, feas,
                                 dual_feas);
    });
    return 0;
}

int emu_ens_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff,
                 const double* vecs, const int* dualf, const double* scal, const double* point, const double* arr,
                 int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::ens_prod_kernel(ncones, want_dual, off, dim, d1s, voff, vecs, dualf, scal, point, arr, ld_arr, prod,
                                ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_ens_dder3(int ncones, const int64_t* off, const int* dim, const int* d1s, const int64_t* voff,
                  const double* vecs, const double* scal, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::ens_dder3_kernel(ncones, off, dim, d1s, voff, vecs, scal, point, dir, out); });
    return 0;
}

int emu_hpm_state(int ncones, const int64_t* off, const int* dim, const int64_t* aoff, const double* alpha,
                  const int* kidx, const int64_t* moff, const double* point, const double* dual, double* grad,
                  double* scal, double* H, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::hpm_state_kernel(ncones, off, dim, aoff, alpha, kidx, moff, point, dual, grad, scal, H, feas, dual_feas);
    });
    return 0;
}

int emu_hpm_prod(int ncones, int want_dual, const int64_t* off, const int* dim, const int64_t* aoff,
                 const double* alpha, const int* dualf, const double* scal, const double* point, const double* arr,
                 int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        hypdev::hpm_prod_kernel(ncones, want_dual, off, dim, aoff, alpha, dualf, scal, point, arr, ld_arr, prod,
                                ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_hpm_dder3(int ncones, const int64_t* off, const int* dim, const int64_t* aoff, const double* alpha,
                  const double* scal, const double* point, const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::hpm_dder3_kernel(ncones, off, dim, aoff, alpha, scal, point, dir, out); });
    return 0;
}

}  // extern "C"

// ---- Nonnegative / EpiNormEucl and the proximity reductions (csrc/cones_vec_kernels.cuh) ----
#include "../../hypatia.jl_b200/csrc/cones_vec_kernels.cuh"

extern "C" {

int emu_nn_state(int64_t nrows, const int* rows, const int* rowcone, const double* point, const double* dual,
                 double* grad, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3(2), dim3(64), 0,
                [&] { hypdev::nn_state_kernel(nrows, rows, rowcone, point, dual, grad, feas, dual_feas); });
    return 0;
}

int emu_nn_prod(int mode, int64_t nrows, const int* rows, const double* point, const double* arr, int64_t ld_arr,
                double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3(2, 2), dim3(64), 0, [&] {
        if (mode == 0) hypdev::nn_prod_kernel<0>(nrows, rows, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 1) hypdev::nn_prod_kernel<1>(nrows, rows, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 2) hypdev::nn_prod_kernel<2>(nrows, rows, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else hypdev::nn_prod_kernel<3>(nrows, rows, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_nn_dder3(int64_t nrows, const int* rows, const double* point, const double* dir, double* out) {
    emu::launch(dim3(2), dim3(64), 0, [&] { hypdev::nn_dder3_kernel(nrows, rows, point, dir, out); });
    return 0;
}

int emu_soc_state(int ncones, const int64_t* off, const int* dim, const int* kidx, const double* point,
                  const double* dual, double* grad, double* scal, uint8_t* feas, uint8_t* dual_feas) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0, [&] {
        hypdev::soc_state_kernel(ncones, off, dim, kidx, point, dual, grad, scal, feas, dual_feas);
    });
    return 0;
}

int emu_soc_prod(int mode, int ncones, const int64_t* off, const int* dim, const double* scal, const double* point,
                 const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    emu::launch(dim3((ncones + 1) / 2, 2), dim3(64), 0, [&] {
        if (mode == 0) hypdev::soc_prod_kernel<0>(ncones, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 1) hypdev::soc_prod_kernel<1>(ncones, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 2) hypdev::soc_prod_kernel<2>(ncones, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else hypdev::soc_prod_kernel<3>(ncones, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_soc_prod_chunk(int mode, int nchunks, int smem_bytes, const int64_t* crow0, const int* crows, const int* ccone0,
                       const int* ccount, const int64_t* off, const int* dim, const double* scal, const double* point,
                       const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols,
                       int64_t row_shift) {
    emu::launch(dim3(nchunks, 2), dim3(64), smem_bytes, [&] {
        if (mode == 0)
            hypdev::soc_prod_chunk_kernel<0>(crow0, crows, ccone0, ccount, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 1)
            hypdev::soc_prod_chunk_kernel<1>(crow0, crows, ccone0, ccount, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (mode == 2)
            hypdev::soc_prod_chunk_kernel<2>(crow0, crows, ccone0, ccount, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else
            hypdev::soc_prod_chunk_kernel<3>(crow0, crows, ccone0, ccount, off, dim, scal, point, arr, ld_arr, prod, ld_prod, ncols, row_shift);
    });
    return 0;
}

int emu_soc_dder3(int ncones, const int64_t* off, const int* dim, const double* scal, const double* point,
                  const double* dir, double* out) {
    emu::launch(dim3((ncones + 1) / 2), dim3(64), 0,
                [&] { hypdev::soc_dder3_kernel(ncones, off, dim, scal, point, dir, out); });
    return 0;
}

int emu_cone_prox(int ncones, const int* ctype, const int64_t* coff, const int64_t* cdim, const double* cnu,
                  const double* point, const double* dual, const double* grad, const double* v1, const double* v2,
                  const double* v3, double irtmu, int use_max, double* proxsqr, uint8_t* num_ok) {
    emu::launch(dim3(ncones), dim3(128), 0, [&] {
        hypdev::cone_prox_kernel(0, ctype, coff, cdim, cnu, point, dual, grad, v1, v2, v3, irtmu, use_max, proxsqr,
                                 num_ok);
    });
    return 0;
}

}  // extern "C"

// ---- digit slicing of the FP64-accurate SYRK (csrc/ozaki_slice_kernels.cuh) ----
#include "../../hypatia.jl_b200/csrc/ozaki_slice_kernels.cuh"

extern "C" {

// expo / dscale and the digit slices D[s][k + j * ldd] of the K x ncols matrix A; radix = 128 or 256
int emu_ozaki_slice(int radix, int64_t K, int64_t ncols, const double* A, int64_t lda, int nslices, int* expo,
                    double* dscale, int8_t* D, int64_t ldd, int64_t slice_stride) {
    emu::launch(dim3((unsigned)ncols), dim3(64), 0,
                [&] { hypdev::colmax_kernel(K, ncols, A, lda, expo, dscale, radix == 256 ? 1 : 0); });
    emu::launch(dim3(2, 2), dim3(64), 0, [&] {
        if (radix == 256) hypdev::slice256_kernel(K, ncols, A, lda, expo, nslices, D, ldd, slice_stride);
        else hypdev::slice_kernel(K, ncols, A, lda, expo, nslices, D, ldd, slice_stride);
    });
    return 0;
}

}  // extern "C"

// ---- the single-product GEMV kernels (csrc/gemv_kernels.cuh), launched as hyp_gemv_t / hyp_gemv_n do (gemv.cu) ----
extern "C" {

// y = alpha * M' x + beta * y; kind 0: CTA per column (vectorised loads), 1: CTA per column (scalar), 2: warp per column
int emu_gemv_t(int kind, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x, double alpha,
               double beta, double* y) {
    if (kind == 2) {
        emu::launch(dim3(2), dim3(256), 0, [&] { hypdev::gemv_t_warp_kernel(rows, ncols, M, ld, x, alpha, beta, y); });
    } else {
        emu::launch(dim3(3), dim3(256), 0, [&] {
            if (kind == 0) hypdev::gemv_t_cta_kernel<true>(rows, ncols, M, ld, x, alpha, beta, y);
            else hypdev::gemv_t_cta_kernel<false>(rows, ncols, M, ld, x, alpha, beta, y);
        });
    }
    return 0;
}

// y = alpha * M x + beta * y through the column-chunked partial sums
int emu_gemv_n(int vec, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x, int nchunks,
               double alpha, double beta, double* y) {
    const int rb = (int)((rows + 255) / 256);
    const int64_t cpc = (ncols + nchunks - 1) / nchunks;
    nchunks = (int)((ncols + cpc - 1) / cpc);
    std::vector<double> part((size_t)nchunks * rows + 2);
    emu::launch(dim3(rb, nchunks), dim3(128), 0, [&] {
        if (vec) hypdev::gemv_n_kernel<true>(rows, ncols, M, ld, x, cpc, part.data());
        else hypdev::gemv_n_kernel<false>(rows, ncols, M, ld, x, cpc, part.data());
    });
    emu::launch(dim3(2), dim3(256), 0,
                [&] { hypdev::gemv_n_reduce_kernel(rows, nchunks, part.data(), alpha, beta, y); });
    return 0;
}

}  // extern "C"

// ---- factor-and-invert kernel of the blocked Cholesky and its batched variant (csrc/chol_kernels.cuh) ----
#include "../../hypatia.jl_b200/csrc/chol_kernels.cuh"
#include "../../hypatia.jl_b200/csrc/trsv_tasks.h"

extern "C" {

// diagonal block blk0 of the m x m matrix A (upper triangle): A -> U in place, dinv block <- U^-1; info as dpotrf
int emu_panel_factor(double* A, int64_t lda, int64_t m, int64_t blk0, double* dinv, int* info) {
    emu::launch(dim3(1), dim3(hypdev::PT), (size_t)hypdev::NB * hypdev::LDU * 8,
                [&] { hypdev::panel_kernel<true>(A, lda, m, blk0, dinv, info); });
    return 0;
}

// inverts every diagonal block of the upper triangular m x m matrix A (hyp_trtri_diag)
int emu_panel_invert(double* A, int64_t lda, int64_t m, double* dinv) {
    const int nblk = (int)((m + hypdev::NB - 1) / hypdev::NB);
    emu::launch(dim3(nblk), dim3(hypdev::PT), (size_t)hypdev::NB * hypdev::LDU * 8,
                [&] { hypdev::panel_kernel<false>(A, lda, m, 0, dinv, nullptr); });
    return 0;
}

int emu_chol_batched(int ncones, const int* sides, const int64_t* moff, const int* kidx, double* U, double* Ui,
                     uint8_t* flag) {
    emu::launch(dim3(ncones), dim3(hypdev::PT), (size_t)hypdev::NB * hypdev::LDU * 8,
                [&] { hypdev::chol_batched_kernel(ncones, sides, moff, kidx, U, Ui, flag); });
    return 0;
}

}  // extern "C"

// ---- the whole state update of a group of PosSemidefTri / HypoPerLogdetTri / HypoRootdetTri cones with sides <= 128,
// as hyp_mat_update_state (cones_mat.cu) launches it: unpack -> batched Cholesky + inverse -> mat_post, then the dual
// point: unpack -> batched Cholesky -> mat_dualfeas ----
extern "C" {

int emu_mat_state(int type, int ncones, const int64_t* off, const int* sides, const int64_t* moff, const int* kidx,
                  const double* point, const double* dual, double* W, double* U, double* Ui, double* Ut, double* Uit,
                  double* Wi, double* U2, double* Ui2, double* scal, double* grad, double* wivec, uint8_t* feas,
                  uint8_t* dual_feas) {
    const int lead = type == 2 ? 0 : type == 3 ? 2 : 1;
    const size_t csm = (size_t)hypdev::NB * hypdev::LDU * 8;
    emu::launch(dim3(ncones, 2), dim3(64), 0,
                [&] { hypdev::unpack_state_kernel(ncones, off, sides, moff, lead, point, W, U); });
    emu::launch(dim3(ncones), dim3(hypdev::PT), csm,
                [&] { hypdev::chol_batched_kernel(ncones, sides, moff, kidx, U, Ui, feas); });
    emu::launch(dim3(ncones), dim3(256), 0, [&] {
        hypdev::mat_post_kernel(type, ncones, off, sides, moff, kidx, point, U, Ui, Ut, Uit, Wi, scal, grad, wivec, feas);
    });
    emu::launch(dim3(ncones, 2), dim3(64), 0,
                [&] { hypdev::unpack_state_kernel(ncones, off, sides, moff, lead, dual, U2, nullptr); });
    emu::launch(dim3(ncones), dim3(hypdev::PT), csm,
                [&] { hypdev::chol_batched_kernel(ncones, sides, moff, kidx, U2, Ui2, dual_feas); });
    if (type != 2)
        emu::launch(dim3(ncones), dim3(128), 0,
                    [&] { hypdev::mat_dualfeas_kernel(type, ncones, off, sides, moff, kidx, dual, U2, dual_feas); });
    return 0;
}

}  // extern "C"

// ---- rook-pivoted LDL' factorisation and solve (csrc/ldlt_kernels.cuh); the launch loop mirrors hyp_ldlt_factor /
// hyp_ldlt_solve of ldlt.cu (one five-kernel sequence per column, device-side cursor in st) ----
#include "../../hypatia.jl_b200/csrc/ldlt_kernels.cuh"

extern "C" {

// A: upper triangle on entry -> L, D in place; ipiv: 3 m ints; returns info (0 or the first exactly singular column)
int emu_ldlt_factor(double* A, int64_t lda, int64_t m, int* ipiv) {
    std::vector<int> st(hypdev::ST_NUM + 1, 0);
    std::vector<double> work((size_t)4 * m + 4, 0.0);
    int info = 0;
    emu::launch(dim3(1), dim3(32), 0, [&] { hypdev::init_state_kernel(st.data()); });
    emu::launch(dim3(1, (unsigned)m), dim3(64), 0, [&] { hypdev::symmetrize_kernel(A, lda, m); });
    for (int64_t seq = 0; seq < m; seq++) {
        emu::launch(dim3(1), dim3(1024), 0, [&] { hypdev::pivot_kernel(A, lda, m, st.data(), ipiv); });
        emu::launch(dim3(2), dim3(64), 0, [&] { hypdev::swap_kernel(A, lda, m, st.data(), 0); });
        emu::launch(dim3(2), dim3(64), 0, [&] { hypdev::swap_kernel(A, lda, m, st.data(), 1); });
        emu::launch(dim3(2), dim3(64), 0, [&] { hypdev::colprep_kernel(A, lda, m, st.data(), work.data()); });
        emu::launch(dim3(2, 3), dim3(256), 0, [&] { hypdev::update_kernel(A, lda, m, st.data(), work.data()); });
    }
    emu::launch(dim3(1), dim3(32), 0, [&] { hypdev::finish_info_kernel(st.data(), &info); });
    return info;
}

int emu_ldlt_solve(const double* A, int64_t lda, int64_t m, const int* ipiv, double* x) {
    emu::launch(dim3(1), dim3(1024), 0, [&] { hypdev::ldlt_solve_kernel(A, lda, m, ipiv, x); });
    return 0;
}

int emu_increase_diag(double* A, int64_t lda, int64_t m) {
    emu::launch(dim3(2), dim3(64), 0, [&] { hypdev::increase_diag_kernel(A, lda, m); });
    return 0;
}

}  // extern "C"

// ---- triangular solves with the blocked factor (trsv_kernel of csrc/chol_kernels.cuh, launched as hyp_trsv_upper) ----
extern "C" {

// x <- U^-T x (trans = 1) or U^-1 x (trans = 0) with the inverted diagonal blocks `dinv` of the upper factor F
int emu_trsv_upper(const double* F, int64_t ldf, int64_t m, const double* dinv, double* x, int trans) {
    const int nblk = (int)((m + hypdev::NB - 1) / hypdev::NB);
    std::vector<int> flags((size_t)nblk + 2, 0);
    const int epoch = 7;
    emu::launch(dim3(nblk < 3 ? nblk : 3), dim3(256), (size_t)hypdev::NB * hypdev::NB * 8, [&] {
        if (trans) hypdev::trsv_kernel<true>(F, ldf, m, dinv, x, flags.data(), nblk, epoch);
        else hypdev::trsv_kernel<false>(F, ldf, m, dinv, x, flags.data(), nblk, epoch);
    });
    return 0;
}

// the packet variant (trsv_pkt_kernel, the default of hyp_trsv_upper / hyp_trsv_upper2): nrhs = 1 or 2 right-hand sides at
// stride xstride; the packet buffer is zeroed once and reused by both sweeps with growing epochs, as in the library
int emu_trsv_upper_pkt(const double* F, int64_t ldf, int64_t m, const double* dinv, double* x, int64_t xstride, int nrhs,
                       int trans, unsigned long long* pkt, int epoch) {
    const int nblk = (int)((m + hypdev::NB - 1) / hypdev::NB);
    std::vector<int> flags((size_t)nblk + 2, 0);
    emu::launch(dim3(nblk < 3 ? nblk : 3), dim3(256), (size_t)hypdev::NB * hypdev::NB * 8, [&] {
        if (nrhs == 2) {
            if (trans) hypdev::trsv_pkt_kernel<true, 2>(F, ldf, m, dinv, x, flags.data(), pkt, nblk, epoch, xstride);
            else hypdev::trsv_pkt_kernel<false, 2>(F, ldf, m, dinv, x, flags.data(), pkt, nblk, epoch, xstride);
        } else {
            if (trans) hypdev::trsv_pkt_kernel<true>(F, ldf, m, dinv, x, flags.data(), pkt, nblk, epoch);
            else hypdev::trsv_pkt_kernel<false>(F, ldf, m, dinv, x, flags.data(), pkt, nblk, epoch);
        }
    });
    return 0;
}

// the segmented variant (trsv_seg_kernel, the default of hyp_trsv_upper) with block columns cut into runs of `seg` tiles
int emu_trsv_upper_seg(const double* F, int64_t ldf, int64_t m, const double* dinv, double* x, int trans, int seg) {
    const int nblk = (int)((m + hypdev::NB - 1) / hypdev::NB);
    int maxseg = 1;
    std::vector<hypdev::TrsvTask> tasks = hypdev::trsv_build_tasks(nblk, trans != 0, seg, &maxseg);
    std::vector<int> flags((size_t)2 * nblk + 2, 0);
    std::vector<double> part((size_t)nblk * maxseg * 2 * hypdev::NB, std::nan(""));
    const int epoch = 5;
    emu::launch(dim3(3), dim3(256), (size_t)hypdev::NB * hypdev::NB * 8, [&] {
        if (trans)
            hypdev::trsv_seg_kernel<true>(F, ldf, m, dinv, x, flags.data(), tasks.data(), (int)tasks.size(), nblk, epoch,
                                          part.data(), maxseg);
        else
            hypdev::trsv_seg_kernel<false>(F, ldf, m, dinv, x, flags.data(), tasks.data(), (int)tasks.size(), nblk, epoch,
                                           part.data(), maxseg);
    });
    return 0;
}

}  // extern "C"
