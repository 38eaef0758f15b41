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

// "Corrected" Neumann load (the reference writes a nodal 1.0 per dof regardless of the
// mesh size, elements.py:195-197 FIXME): consistent load of a constant traction (tx, ty) on boundary edges,
// rhs[2n+c] += t_c * int N_n |dx/dxi| dxi.  NEN = 2: straight 2-node edge (a, b); NEN = 3: quadratic edge
// (a, b, mid), 3-point Gauss-Legendre.  A node receives at most two contributions, so the atomic adds
// commute and the result does not depend on scheduling.  Dirichlet dofs (dirmask) are left at their value.
template <int NEN>
__global__ void k_edge_load(int64_t nedge, const int32_t* __restrict__ edges, const double2* __restrict__ coords,
                            double tx, double ty, const uint8_t* __restrict__ dirmask, double* __restrict__ rhs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nedge) return;
    int n[NEN];
    double2 p[NEN];
    double w[NEN];
#pragma unroll
    for (int k = 0; k < NEN; ++k) {
        n[k] = edges[e * NEN + k];
        p[k] = coords[n[k]];
        w[k] = 0.0;
    }
    if (NEN == 2) {
        const double len = sqrt((p[1].x - p[0].x) * (p[1].x - p[0].x) + (p[1].y - p[0].y) * (p[1].y - p[0].y));
        w[0] = w[1] = 0.5 * len;
    } else {
        const double gp[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
        const double gw[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const double xi = gp[g];
            const double N0 = 0.5 * xi * (xi - 1.0), N1 = 0.5 * xi * (xi + 1.0), N2 = 1.0 - xi * xi;
            const double d0 = xi - 0.5, d1 = xi + 0.5, d2 = -2.0 * xi;
            const double jx = d0 * p[0].x + d1 * p[1].x + d2 * p[NEN - 1].x;
            const double jy = d0 * p[0].y + d1 * p[1].y + d2 * p[NEN - 1].y;
            const double ds = sqrt(jx * jx + jy * jy) * gw[g];
            w[0] = fma(N0, ds, w[0]);
            w[1] = fma(N1, ds, w[1]);
            w[NEN - 1] = fma(N2, ds, w[NEN - 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < NEN; ++k) {
        const int64_t d = 2 * (int64_t)n[k];
        if (!dirmask || !dirmask[d]) atomicAdd(rhs + d, tx * w[k]);
        if (!dirmask || !dirmask[d + 1]) atomicAdd(rhs + d + 1, ty * w[k]);
    }
}

}  // namespace fe

extern "C" {

int fe_edge_load(int nodes_per_edge, int64_t nedge, const int32_t* edges, const double* coords, double tx, double ty,
                 const uint8_t* dirmask, double* rhs, void* stream) {
    using namespace fe;
    FE_REQUIRE(nedge >= 0, "negative size");
    FE_REQUIRE(nodes_per_edge == 2 || nodes_per_edge == 3, "edges have 2 (linear) or 3 (quadratic: a, b, mid) nodes");
    if (nedge == 0) return FE_OK;
    FE_REQUIRE(edges && coords && rhs, "null pointer");
    cudaStream_t st = as_stream(stream);
    const double2* xy = reinterpret_cast<const double2*>(coords);
    if (nodes_per_edge == 2) k_edge_load<2><<<grid_for(nedge, 128), 128, 0, st>>>(nedge, edges, xy, tx, ty, dirmask, rhs);
    else k_edge_load<3><<<grid_for(nedge, 128), 128, 0, st>>>(nedge, edges, xy, tx, ty, dirmask, rhs);
    FE_CHECK_LAUNCH("k_edge_load");
    return FE_OK;
}

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
