// K4: boundary conditions -- right-hand side, Dirichlet mask, and in-place matrix modification.
// Replaces the Dirichlet / "Neumann" loops of FemProblem.assemble (/root/reference/elements.py:185-197).
#include "fe_common.cuh"

namespace fe {

__global__ void k_mark_dirichlet(int64_t n_dir, const int64_t* __restrict__ dofs, double* __restrict__ rhs,
                                 uint8_t* __restrict__ mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dir) return;
    const int64_t d = dofs[i];
    mask[d] = 1;
    rhs[d] = 0.0;  // elements.py:193
}

__global__ void k_set_load(int64_t n_neu, const int64_t* __restrict__ dofs, double value, double* __restrict__ rhs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_neu) rhs[dofs[i]] = value;  // elements.py:195-197, runs after the Dirichlet loop
}

// One thread per row.  Dirichlet row: zeros + unit diagonal (elements.py:187-192).  In symmetric mode a
// free row also gives up its Dirichlet-column entries to the right-hand side, in column order.
template <typename IdxT, bool SYM>
__global__ void __launch_bounds__(256) k_apply_bc(int64_t nrows, const IdxT* __restrict__ row_ptr,
                                                  const int32_t* __restrict__ col_idx, double* __restrict__ vals,
                                                  double* __restrict__ rhs, const uint8_t* __restrict__ mask) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t j0 = row_ptr[r], j1 = row_ptr[r + 1];
    if (mask[r]) {
        for (int64_t j = j0; j < j1; ++j) vals[j] = (col_idx[j] == r) ? 1.0 : 0.0;
        return;
    }
    if (!SYM) return;
    double lift = 0.0;
    bool any = false;
    for (int64_t j = j0; j < j1; ++j) {
        const int c = col_idx[j];
        if (mask[c]) {
            lift = fma(vals[j], rhs[c], lift);
            vals[j] = 0.0;
            any = true;
        }
    }
    if (any) rhs[r] -= lift;
}

}  // namespace fe

extern "C" {

int fe_bc_rhs(int64_t nrows, const int64_t* dir_dofs, int64_t n_dir, const int64_t* neu_dofs, int64_t n_neu,
              double neu_value, double* rhs, uint8_t* dirmask, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows >= 0 && n_dir >= 0 && n_neu >= 0 && rhs && dirmask, "null pointer or negative size");
    FE_REQUIRE((n_dir == 0 || dir_dofs) && (n_neu == 0 || neu_dofs), "dof list is null");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(rhs, 0, (size_t)nrows * sizeof(double), st));
    FE_CUDA_TRY(cudaMemsetAsync(dirmask, 0, (size_t)nrows, st));
    if (n_dir > 0) {
        k_mark_dirichlet<<<grid_for(n_dir, 256), 256, 0, st>>>(n_dir, dir_dofs, rhs, dirmask);
        FE_CHECK_LAUNCH("k_mark_dirichlet");
    }
    if (n_neu > 0) {
        k_set_load<<<grid_for(n_neu, 256), 256, 0, st>>>(n_neu, neu_dofs, neu_value, rhs);
        FE_CHECK_LAUNCH("k_set_load");
    }
    return FE_OK;
}

int fe_apply_bc(int mode, int64_t nrows, int idx64, const void* row_ptr, const int32_t* col_idx, double* vals,
                double* rhs, const uint8_t* dirmask, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows >= 0 && row_ptr && col_idx && vals && dirmask, "null pointer or negative size");
    FE_REQUIRE(mode == FE_BC_ROWS || mode == FE_BC_SYMMETRIC, "unknown mode");
    FE_REQUIRE(mode == FE_BC_ROWS || rhs, "symmetric mode needs rhs");
    if (nrows == 0) return FE_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned g = grid_for(nrows, 256);
    if (idx64) {
        const int64_t* rp = static_cast<const int64_t*>(row_ptr);
        if (mode == FE_BC_SYMMETRIC) k_apply_bc<int64_t, true><<<g, 256, 0, st>>>(nrows, rp, col_idx, vals, rhs, dirmask);
        else k_apply_bc<int64_t, false><<<g, 256, 0, st>>>(nrows, rp, col_idx, vals, rhs, dirmask);
    } else {
        const int32_t* rp = static_cast<const int32_t*>(row_ptr);
        if (mode == FE_BC_SYMMETRIC) k_apply_bc<int32_t, true><<<g, 256, 0, st>>>(nrows, rp, col_idx, vals, rhs, dirmask);
        else k_apply_bc<int32_t, false><<<g, 256, 0, st>>>(nrows, rp, col_idx, vals, rhs, dirmask);
    }
    FE_CHECK_LAUNCH("k_apply_bc");
    return FE_OK;
}

}  // extern "C"
