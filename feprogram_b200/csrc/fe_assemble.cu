// K1 (stand-alone element matrices) and K3 (deterministic numeric assembly).
#include "fe_element.cuh"

namespace fe {

// ------------------------------------------------------------------------------------------------
// K1: dense Ke per element, one thread per (element, local node a) writing rows 2a and 2a+1.
// Parity/debug entry (FemProblem.getKe); the assembly path below never stores Ke.
// ------------------------------------------------------------------------------------------------
template <int ET, bool MASS, bool SCALAR = false>
__global__ void __launch_bounds__(128) k_element_matrices(int64_t nelem, const int32_t* __restrict__ conn,
                                                          const double2* __restrict__ coords,
                                                          double* __restrict__ Ke, int weighted) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE, M = SCALAR ? N : 2 * N;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nelem * N) return;
    const int64_t e = idx / N;
    const int a = (int)(idx - e * N);
    int t[N];
    double x[N], y[N], L[N], P[N], Q[N];
    load_element<ET>(conn, coords, e, t, x, y);
    element_row_pair<ET, MASS>(x, y, a, weighted != 0, L, P, Q);
    if constexpr (SCALAR) {         // FE_OP_POISSON: Ke[a][b] = L = Ke2[2a][2b]
        double* r = Ke + e * (M * M) + a * M;
#pragma unroll
        for (int b = 0; b < N; ++b) r[b] = L[b];
        return;
    }
    double* r0 = Ke + e * (M * M) + (2 * a) * M;
    double* r1 = r0 + M;
#pragma unroll
    for (int b = 0; b < N; ++b) {
        r0[2 * b] = L[b];
        r0[2 * b + 1] = P[b];
        r1[2 * b] = Q[b];
        r1[2 * b + 1] = L[b];
    }
}

// ------------------------------------------------------------------------------------------------
// K3 (row-tile form): one CTA owns `R` consecutive nodes = one contiguous range of CSR values.
// One thread per node walks that node's incident elements in ascending element id (einc order),
// evaluates the node's row pair of each element matrix in registers and accumulates into the
// node's private row block in shared memory -- a fixed summation order without atomics, so the
// result is bit-identical from run to run.  The tile is then written once, fully coalesced.
// Shared layout mirrors the global layout: value (row 2n+c, p, d) at 4*(nptr[n]-base) + 2*c*deg + 2p + d
// (SCALAR = FE_OP_POISSON, one dof per node: value (row n, p) at nptr[n]-base + p, the L entries only).
// ------------------------------------------------------------------------------------------------
template <int ET, bool MASS, bool SCALAR = false>
__global__ void __launch_bounds__(128) k_assemble_rows(int64_t nnode, int R, const int32_t* __restrict__ eptr,
                                                       const int32_t* __restrict__ einc,
                                                       const uint8_t* __restrict__ epos,
                                                       const int32_t* __restrict__ nptr,
                                                       const int32_t* __restrict__ conn,
                                                       const double2* __restrict__ coords,
                                                       double* __restrict__ vals, int weighted) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    extern __shared__ double acc[];
    const int64_t n0 = (int64_t)blockIdx.x * R;
    const int64_t n1 = (n0 + R < nnode) ? n0 + R : nnode;
    constexpr int kPerNz = SCALAR ? 1 : 4;
    const int base = nptr[n0];
    const int len = kPerNz * (nptr[n1] - base);
    for (int i = threadIdx.x; i < len; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();

    const int64_t n = n0 + threadIdx.x;
    if ((int)threadIdx.x < R && n < n1) {
        const int p0 = nptr[n];
        const int deg = nptr[n + 1] - p0;
        double* row0 = acc + kPerNz * (p0 - base);
        double* row1 = row0 + 2 * deg;
        const int k1 = eptr[n + 1];
        for (int k = eptr[n]; k < k1; ++k) {
            const int item = __ldg(einc + k);
            const int64_t e = item >> 4;
            const int a = item & 15;
            int t[N];
            double x[N], y[N], L[N], P[N], Q[N];
            load_element<ET>(conn, coords, e, t, x, y);
            element_row_pair<ET, MASS>(x, y, a, weighted != 0, L, P, Q);
            const uint8_t* ps = epos + (int64_t)k * N;
#pragma unroll
            for (int b = 0; b < N; ++b) {
                if constexpr (SCALAR) {
                    row0[(int)__ldg(ps + b)] += L[b];
                } else {
                    const int pp = 2 * (int)__ldg(ps + b);
                    row0[pp] += L[b];
                    row0[pp + 1] += P[b];
                    row1[pp] += Q[b];
                    row1[pp + 1] += L[b];
                }
            }
        }
    }
    __syncthreads();
    double* out = vals + kPerNz * (int64_t)base;
    for (int i = threadIdx.x; i < len; i += blockDim.x) out[i] = acc[i];
}

// 1-based node tags of any integer width -> int32 0-based connectivity (elements.py:172 `elem-1`).
template <typename T>
__global__ void k_tags_to_conn(int64_t n, const T* __restrict__ tags, int32_t* __restrict__ conn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) conn[i] = (int32_t)(tags[i] - 1);
}

template <int ET, bool MASS, bool SCALAR = false>
static int launch_assemble_rows(int64_t nnode, const int32_t* conn, const double* coords, const int32_t* eptr,
                                const int32_t* einc, const uint8_t* epos, const int32_t* nptr, int32_t max_deg,
                                double* vals, int weighted, cudaStream_t stream) {
    // rows per CTA: as many as fit a 48 KB tile at the worst-case row-block size, warp multiple; meshes with nodes of
    // very high valence (more than 48 neighbours) get one warp of rows on a tile of up to 200 KB
    const size_t per_node = (size_t)max_deg * (SCALAR ? 1 : 4) * sizeof(double);
    int R = (int)((48 * 1024) / per_node);
    R = R >= 128 ? 128 : (R / 32) * 32;
    if (R < 32 && 32 * per_node <= 200 * 1024) R = 32;
    if (R < 32) {
        set_error("fe_assemble: max_deg %d exceeds the row-tile capacity", (int)max_deg);
        return FE_ERR_CAPACITY;
    }
    const size_t smem = (size_t)R * per_node;
    FE_CUDA_TRY(cudaFuncSetAttribute(k_assemble_rows<ET, MASS, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_rows<ET, MASS, SCALAR><<<grid_for(nnode, R), R, smem, stream>>>(nnode, R, eptr, einc, epos, nptr, conn,
                                                                 reinterpret_cast<const double2*>(coords), vals,
                                                                 weighted);
    FE_CHECK_LAUNCH("k_assemble_rows");
    return FE_OK;
}

}  // namespace fe

extern "C" {

int fe_tags_to_conn(int tag_bytes, int64_t n, const void* tags, int32_t* conn, void* stream) {
    using namespace fe;
    FE_REQUIRE(n >= 0, "negative size");
    FE_REQUIRE(tag_bytes == 4 || tag_bytes == 8, "tags must be 4- or 8-byte integers");
    if (n == 0) return FE_OK;
    FE_REQUIRE(tags && conn, "null pointer");
    cudaStream_t st = as_stream(stream);
    if (tag_bytes == 4)
        k_tags_to_conn<int32_t><<<grid_for(n, 256), 256, 0, st>>>(n, static_cast<const int32_t*>(tags), conn);
    else
        k_tags_to_conn<int64_t><<<grid_for(n, 256), 256, 0, st>>>(n, static_cast<const int64_t*>(tags), conn);
    FE_CHECK_LAUNCH("k_tags_to_conn");
    return FE_OK;
}

int fe_element_matrices(int elem_type, int op, int64_t nelem, const int32_t* conn, const double* coords, double* Ke,
                        void* stream) {
    using namespace fe;
    FE_REQUIRE(nelem >= 0, "negative size");
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED && op != FE_OP_MASS && op != FE_OP_POISSON) {
        set_error("fe_element_matrices: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (nelem == 0) return FE_OK;
    FE_REQUIRE(conn && coords && Ke, "null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const double2* xy = reinterpret_cast<const double2*>(coords);
    const int w = op == FE_OP_WEIGHTED;
#define FE_KE_CASE(ET, N)                                                                                       \
    case ET:                                                                                                    \
        if (op == FE_OP_MASS)                                                                                   \
            k_element_matrices<ET, true><<<grid_for(nelem * N, 128), 128, 0, st>>>(nelem, conn, xy, Ke, 1);     \
        else if (op == FE_OP_POISSON)                                                                           \
            k_element_matrices<ET, false, true><<<grid_for(nelem * N, 128), 128, 0, st>>>(nelem, conn, xy, Ke, 0); \
        else                                                                                                    \
            k_element_matrices<ET, false><<<grid_for(nelem * N, 128), 128, 0, st>>>(nelem, conn, xy, Ke, w);    \
        break;
    switch (elem_type) {
        FE_KE_CASE(FE_QUAD4, 4)
        FE_KE_CASE(FE_QUAD9, 9)
        FE_KE_CASE(FE_TRI3, 3)
#undef FE_KE_CASE
        default:
            set_error("fe_element_matrices: Invalid elemType %d", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
    FE_CHECK_LAUNCH("k_element_matrices");
    return FE_OK;
}

int fe_assemble(int elem_type, int op, int64_t nelem, int64_t nnode, const int32_t* conn, const double* coords,
                const int32_t* eptr, const int32_t* einc, const uint8_t* epos, const int32_t* nptr, int32_t max_deg,
                double* vals, void* stream) {
    using namespace fe;
    FE_REQUIRE(nelem >= 0 && nnode >= 0 && conn && coords && eptr && einc && epos && nptr && vals,
               "null pointer or negative size");
    FE_REQUIRE(max_deg > 0, "max_deg must be positive");
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED && op != FE_OP_MASS && op != FE_OP_POISSON) {
        set_error("fe_assemble: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (nnode == 0) return FE_OK;
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const int w = op == FE_OP_WEIGHTED;
#define FE_ROWS_CASE(ET)                                                                                          \
    case ET:                                                                                                      \
        if (op == FE_OP_POISSON)                                                                                  \
            return launch_assemble_rows<ET, false, true>(nnode, conn, coords, eptr, einc, epos, nptr, max_deg, vals, 0, st); \
        return op == FE_OP_MASS                                                                                   \
                   ? launch_assemble_rows<ET, true>(nnode, conn, coords, eptr, einc, epos, nptr, max_deg, vals, 1, st) \
                   : launch_assemble_rows<ET, false>(nnode, conn, coords, eptr, einc, epos, nptr, max_deg, vals, w, st);
    switch (elem_type) {
        FE_ROWS_CASE(FE_QUAD4)
        FE_ROWS_CASE(FE_QUAD9)
        FE_ROWS_CASE(FE_TRI3)
#undef FE_ROWS_CASE
        default:
            set_error("fe_assemble: Invalid elemType %d", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
