// K5 (CSR SpMV) and K6 (fused Jacobi-PCG vector kernels).
//
// Replaces spsolve(K, brhs) (/root/reference/elements.py:241-243) by Jacobi-preconditioned CG on the
// SPD system the reference's row-only Dirichlet treatment is equivalent to:
// x_D = g_D, K_II x_I = b_I - K_ID g_D.  The matrix is never modified: the search direction is kept
// zero on Dirichlet dofs and the SpMV zeroes Dirichlet rows, which is the same operator.
//
// Per iteration: 3 kernels, 2 deterministic two-stage reductions, all scalars stay on the device.
//   A: Ap = A p (masked), partial dots p.Ap                       [fe SpMV traffic]
//   B: alpha = rz/pAp; x += alpha p; r -= alpha Ap; partial r.z, r.r with z = Minv r   [7 vector streams]
//   C: beta = rz'/rz; p = Minv r + beta p; convergence flag                            [4 vector streams]
#include "fe_common.cuh"

namespace fe {

constexpr int kVecThreads = 256;
constexpr int kMaxPartials = 4096;  // upper bound on persistent grid sizes
constexpr int kScalarDoubles = 16;

// scalar block layout (doubles)
// Flags and scalars are double-buffered between kernels so that no kernel reads a location one of its
// own blocks writes:  DONE_C / RZ_CUR are written by kernel C and read by A and B;  DONE_B / RZ_PREV
// are written by kernel B and read by C.
// S_ERR: 0 = fine, kErrPeerTimeout = a peer-memory wait timed out, kErrBreakdown = CG breakdown (NaN residual or
// p.Ap <= 0: the matrix is not SPD on the free dofs).  info[3] reports 1 only for a genuine ||r|| <= tol stop.
constexpr int kErrPeerTimeout = 1, kErrBreakdown = 2;
enum { S_RZ_CUR = 0, S_RZ_PREV = 1, S_RR = 2, S_RR0 = 3, S_TOL2 = 4, S_ITERS = 5, S_DONE_C = 6, S_PAP = 7, S_DONE_B = 8, S_ERR = 9 };

// ---- deterministic block reduction (fixed shuffle tree, then warp 0 over the warp sums) ----------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

// Result valid in thread 0.  `sm` needs blockDim/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sm reuse
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

// Every block sums the same `n` partials in the same order => identical value in all blocks.
// Result broadcast to all threads.
__device__ __forceinline__ double sum_partials(const double* __restrict__ part, int n, int stride, double* sm) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[(size_t)i * stride];
    v = block_sum(v, sm);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = v;
    __syncthreads();
    return bc;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory all-reduce of the PCG scalars (multi-GPU, fe_dist_pcg_p2p).  Every rank owns a slot table
// double[2][kMaxRanks][2] and a version array int64[kMaxRanks] in peer-mapped memory.  Reduction number
// `seq` (0, 1, 2 ... over the life of the buffers): a rank stores its local sum into slot [seq & 1][rank] of
// EVERY rank (itself included) and then release-stores seq + 1 into that rank's version[rank]; a consumer
// waits until all `world` versions reach seq + 1 and adds the slots in rank order, so every rank (and
// every CTA) gets the same bits.  Slot reuse at seq + 2 is safe: nobody can publish seq + 2 before it has
// consumed seq + 1, which needs every rank's seq + 1, which each rank publishes after consuming seq.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
struct PeerRed {
    double* const* slots;        // device array [world] of peer pointers
    long long* const* versions;  // device array [world] of peer pointers
    int rank, world;             // world == 0: not used
};

__device__ __forceinline__ long long ld_acquire_sys_s64(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_s64(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// all threads of the CTA call this; v[0..nv) <- the global sums of reduction `seq`; false on time-out
__device__ __forceinline__ bool peer_reduced(const PeerRed& pr, long long seq, int nv, double* v) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if ((int)threadIdx.x < pr.world) {
        const long long* ver = pr.versions[pr.rank] + threadIdx.x;
        if (ld_acquire_sys_s64(ver) < seq + 1) {           // exponential back-off, wall-clock bound (2 s)
            unsigned long long t0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            unsigned ns = 32;
            while (ld_acquire_sys_s64(ver) < seq + 1) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > 2000000000ull) {
                    timed_out = 1;
                    break;
                }
                __nanosleep(ns);
                if (ns < 2048) ns <<= 1;
            }
        }
    }
    __syncthreads();
    const double* s = pr.slots[pr.rank] + (seq & 1) * (2 * kMaxRanks);
    for (int k = 0; k < nv; ++k) {
        double a = 0.0;
        for (int q = 0; q < pr.world; ++q) a += __ldcg(s + 2 * q + k);   // L2: the slots are written by peers
        v[k] = a;
    }
    return timed_out == 0;
}

// ------------------------------------------------------------------------------------------------
// K5: CSR SpMV, LPR lanes per row, persistent grid-stride over row groups.  Lanes of a row read
// consecutive (val, col) entries, adjacent rows are adjacent in memory => coalesced streams; x is
// gathered through L1/L2.  Fixed in-row reduction tree => deterministic.
// DOT: also accumulates sum_r x[r]*y[r] per block into partials[blockIdx.x] (x is the SpMV input).
// ------------------------------------------------------------------------------------------------
template <typename IdxT, int LPR, bool DOT>
__global__ void __launch_bounds__(kVecThreads) k_spmv(int64_t nrows, const IdxT* __restrict__ row_ptr,
                                                      const int32_t* __restrict__ col_idx,
                                                      const double* __restrict__ vals, const double* __restrict__ x,
                                                      double* __restrict__ y, const uint8_t* __restrict__ rowmask,
                                                      double* __restrict__ partials,
                                                      const double* __restrict__ scal) {
    __shared__ double sm[kVecThreads / 32];
    if (DOT && scal[S_DONE_C] != 0.0) return;
    constexpr int RPB = kVecThreads / LPR;
    const int lane = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    double dot = 0.0;
    for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < nrows; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + grp;
        double sum = 0.0;
        if (row < nrows) {
            const int64_t j0 = row_ptr[row], j1 = row_ptr[row + 1];
            for (int64_t j = j0 + lane; j < j1; j += LPR) sum = fma(__ldg(vals + j), __ldg(x + __ldg(col_idx + j)), sum);
        }
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d, LPR);
        if (lane == 0 && row < nrows) {
            if (rowmask && rowmask[row]) sum = 0.0;
            y[row] = sum;
            if (DOT) dot = fma(__ldg(x + row), sum, dot);
        }
    }
    if (DOT) {
        dot = block_sum(dot, sm);
        if (threadIdx.x == 0) partials[blockIdx.x] = dot;
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256) k_csr_diagonal(int64_t nrows, const IdxT* __restrict__ row_ptr,
                                                      const int32_t* __restrict__ col_idx,
                                                      const double* __restrict__ vals, double* __restrict__ diag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double d = 0.0;
    for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j)
        if (col_idx[j] == r) d = vals[j];
    diag[r] = d;
}

// ------------------------------------------------------------------------------------------------
// K5 (node-block form): the same matrix read through the node-level pattern it was assembled from
// (nptr/nadj): one column index per 2x2 block instead of four, x gathered as double2.  Values stay in the
// CSR layout (row 2n then row 2n+1 of node n), so nothing is converted.  24 % fewer bytes than the CSR
// walk (cfg2: 6.04 GB vs 7.92 GB); used by the PCG solve when the pattern is available.
// ------------------------------------------------------------------------------------------------
template <int L, bool DOT>
__global__ void __launch_bounds__(kVecThreads) k_spmv_nodal(int64_t nnode, const int32_t* __restrict__ nptr,
                                                            const int32_t* __restrict__ nadj,
                                                            const double* __restrict__ vals,
                                                            const double* __restrict__ x, double* __restrict__ y,
                                                            const uint8_t* __restrict__ rowmask,
                                                            double* __restrict__ partials,
                                                            const double* __restrict__ scal) {
    __shared__ double sm[kVecThreads / 32];
    if (DOT && scal[S_DONE_C] != 0.0) return;
    constexpr int GPB = kVecThreads / L;
    const int lane = threadIdx.x % L;
    const int grp = threadIdx.x / L;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    double dot = 0.0;
    for (int64_t n0 = (int64_t)blockIdx.x * GPB; n0 < nnode; n0 += (int64_t)gridDim.x * GPB) {
        const int64_t n = n0 + grp;
        double a0 = 0.0, a1 = 0.0;
        if (n < nnode) {
            const int p0 = __ldg(nptr + n), deg = __ldg(nptr + n + 1) - p0;
            const double2* r0 = reinterpret_cast<const double2*>(vals + 4 * (int64_t)p0);
            const double2* r1 = r0 + deg;
            for (int p = lane; p < deg; p += L) {
                const double2 xv = __ldg(x2 + __ldg(nadj + p0 + p));
                const double2 u = __ldg(r0 + p), v = __ldg(r1 + p);
                a0 = fma(u.x, xv.x, fma(u.y, xv.y, a0));
                a1 = fma(v.x, xv.x, fma(v.y, xv.y, a1));
            }
        }
#pragma unroll
        for (int d = L / 2; d > 0; d >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, d, L);
            a1 += __shfl_down_sync(0xffffffffu, a1, d, L);
        }
        if (lane == 0 && n < nnode) {
            if (rowmask) {
                if (rowmask[2 * n]) a0 = 0.0;
                if (rowmask[2 * n + 1]) a1 = 0.0;
            }
            reinterpret_cast<double2*>(y)[n] = make_double2(a0, a1);
            if (DOT) {
                const double2 xn = __ldg(x2 + n);
                dot = fma(xn.x, a0, fma(xn.y, a1, dot));
            }
        }
    }
    if (DOT) {
        dot = block_sum(dot, sm);
        if (threadIdx.x == 0) partials[blockIdx.x] = dot;
    }
}

// ---- PCG vector kernels ---------------------------------------------------------------------------
// init 1: x = g on Dirichlet dofs, 0 elsewhere;  minv = 1/diag (0 on Dirichlet dofs and empty rows)
__global__ void __launch_bounds__(kVecThreads) k_pcg_init1(int64_t n, const double* __restrict__ rhs,
                                                           const uint8_t* __restrict__ mask, double* __restrict__ x,
                                                           double* __restrict__ minv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool d = mask && mask[i];
        x[i] = d ? rhs[i] : 0.0;
        const double dg = minv[i];  // holds diag on entry
        minv[i] = (d || dg == 0.0) ? 0.0 : 1.0 / dg;
    }
}

// init 2: r = rhs - A x0 on free dofs (Ap holds A x0 with Dirichlet rows zeroed), p = Minv r
__global__ void __launch_bounds__(kVecThreads) k_pcg_init2(int64_t n, const double* __restrict__ rhs,
                                                           const uint8_t* __restrict__ mask,
                                                           const double* __restrict__ Ax0,
                                                           const double* __restrict__ minv, double* __restrict__ r,
                                                           double* __restrict__ p, double* __restrict__ partB) {
    __shared__ double sm[kVecThreads / 32];
    double rz = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = (mask && mask[i]) ? 0.0 : rhs[i] - Ax0[i];
        const double zi = minv[i] * ri;
        r[i] = ri;
        p[i] = zi;
        rz = fma(ri, zi, rz);
        rr = fma(ri, ri, rr);
    }
    rz = block_sum(rz, sm);
    rr = block_sum(rr, sm);
    if (threadIdx.x == 0) {
        partB[2 * blockIdx.x] = rz;
        partB[2 * blockIdx.x + 1] = rr;
    }
}

// `g` (optional): globally reduced scalars of a multi-GPU solve; when given they replace the local sums.
__global__ void __launch_bounds__(kVecThreads) k_pcg_init3(int nb, const double* __restrict__ partB, double rtol,
                                                           double atol, double* __restrict__ scal,
                                                           const double* __restrict__ g, PeerRed pr, long long seq) {
    __shared__ double sm[kVecThreads / 32];
    double gv[2];
    bool ok = true;
    if (pr.world) ok = peer_reduced(pr, seq, 2, gv);
    const double rz = pr.world ? gv[0] : g ? g[0] : sum_partials(partB, nb, 2, sm);
    const double rr = pr.world ? gv[1] : g ? g[1] : sum_partials(partB + 1, nb, 2, sm);
    if (threadIdx.x == 0) {
        scal[S_RZ_CUR] = rz;
        scal[S_RZ_PREV] = rz;
        scal[S_RR] = rr;
        scal[S_RR0] = rr;
        const double t = fmax(rtol * rtol * rr, atol * atol);
        scal[S_TOL2] = t;
        scal[S_ITERS] = 0.0;
        scal[S_DONE_C] = (rr <= t) ? 1.0 : 0.0;
        scal[S_DONE_B] = 0.0;
        scal[S_PAP] = 0.0;
        scal[S_ERR] = ok ? 0.0 : (double)kErrPeerTimeout;
    }
}

// B: x += alpha p; r -= alpha Ap; partial (r.z, r.r)
__global__ void __launch_bounds__(kVecThreads) k_pcg_update(int64_t n, int nbA, const double* __restrict__ partA,
                                                            const double* __restrict__ p,
                                                            const double* __restrict__ Ap,
                                                            const double* __restrict__ minv, double* __restrict__ x,
                                                            double* __restrict__ r, double* __restrict__ partB,
                                                            double* __restrict__ scal, const double* __restrict__ g,
                                                            PeerRed pr, long long seq) {
    __shared__ double sm[kVecThreads / 32];
    const bool done = scal[S_DONE_C] != 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE_B] = done ? 1.0 : 0.0;
    if (done) return;
    double gv[1];
    if (pr.world && !peer_reduced(pr, seq, 1, gv) && threadIdx.x == 0) scal[S_ERR] = (double)kErrPeerTimeout;
    const double pAp = pr.world ? gv[0] : g ? g[0] : sum_partials(partA, nbA, 1, sm);
    const double rz_cur = scal[S_RZ_CUR];
    const double alpha = rz_cur / pAp;
    double rz = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, Ap[i], r[i]);
        r[i] = ri;
        const double zi = minv[i] * ri;
        rz = fma(ri, zi, rz);
        rr = fma(ri, ri, rr);
    }
    rz = block_sum(rz, sm);
    rr = block_sum(rr, sm);
    if (threadIdx.x == 0) {
        partB[2 * blockIdx.x] = rz;
        partB[2 * blockIdx.x + 1] = rr;
        if (blockIdx.x == 0) {
            scal[S_RZ_PREV] = rz_cur;  // nobody reads RZ_PREV during this kernel
            scal[S_PAP] = pAp;
            // p.Ap <= 0 with a non-zero direction: the operator is not positive definite on the free dofs
            if (!(pAp > 0.0) && rz_cur != 0.0) scal[S_ERR] = fmax(scal[S_ERR], (double)kErrBreakdown);
        }
    }
}

// C: beta = rz'/rz; p = Minv r + beta p; block 0 publishes rz', r.r, iteration count, done flag
__global__ void __launch_bounds__(kVecThreads) k_pcg_direction(int64_t n, int nbB, const double* __restrict__ partB,
                                                               const double* __restrict__ r,
                                                               const double* __restrict__ minv,
                                                               const double* p_in, double* p_out,
                                                               double* __restrict__ scal,
                                                               const double* __restrict__ g, PeerRed pr,
                                                               long long seq) {
    // p_out may alias p_in (single GPU) or be the other half of a double-buffered direction vector (P2P halo)
    __shared__ double sm[kVecThreads / 32];
    if (scal[S_DONE_B] != 0.0) return;
    double gv[2];
    if (pr.world && !peer_reduced(pr, seq, 2, gv) && threadIdx.x == 0) scal[S_ERR] = (double)kErrPeerTimeout;
    const double rz_new = pr.world ? gv[0] : g ? g[0] : sum_partials(partB, nbB, 2, sm);
    const double rr = pr.world ? gv[1] : g ? g[1] : sum_partials(partB + 1, nbB, 2, sm);
    const double beta = rz_new / scal[S_RZ_PREV];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p_out[i] = fma(beta, p_in[i], minv[i] * r[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[S_RZ_CUR] = rz_new;
        scal[S_RR] = rr;
        scal[S_ITERS] += 1.0;
        if (rr <= scal[S_TOL2]) {
            scal[S_DONE_C] = 1.0;
        } else if (!(rr == rr) || !(rr < 1.0e300)) {   // breakdown (NaN / overflow): stop, and say so -- not "converged"
            scal[S_DONE_C] = 1.0;
            scal[S_ERR] = fmax(scal[S_ERR], (double)kErrBreakdown);
        }
    }
}

// info[0..3] = iterations, ||r||/||r0||, ||r0||, stop reason: 1 converged, 0 iteration limit, -2 breakdown
static void fill_info(const double* hs, double* info) {
    if (!info) return;
    info[0] = hs[S_ITERS];
    info[1] = hs[S_RR0] > 0.0 ? sqrt(hs[S_RR] / hs[S_RR0]) : 0.0;
    info[2] = sqrt(hs[S_RR0]);
    info[3] = hs[S_ERR] != 0.0 ? -hs[S_ERR] : ((hs[S_DONE_C] != 0.0 && hs[S_RR] <= hs[S_TOL2]) ? 1.0 : 0.0);
}

static int persistent_grid(const void* func, int threads, int* out) {
    int dev = 0, sms = 0, occ = 0;
    FE_CUDA_TRY(cudaGetDevice(&dev));
    FE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, 0));
    if (occ < 1) occ = 1;
    int g = sms * occ;
    if (g > kMaxPartials) g = kMaxPartials;
    *out = g;
    return FE_OK;
}

static int pick_lpr(int64_t nrows, int64_t nnz) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    // few lanes per row keep several independent loads in flight per thread (measured on B200:
    // 18 nnz/row: 4 lanes 6.26 TB/s, 8 lanes 4.95, 2 lanes 5.16; 32 nnz/row: 4 lanes 4.64 TB/s, 8 lanes 4.14,
    // 16 lanes 2.70; 14 nnz/row: 4 lanes 5.51 TB/s, 2 lanes 5.14)
    if (avg < 8.0) return 2;
    if (avg < 64.0) return 4;
    if (avg < 128.0) return 8;
    if (avg < 256.0) return 16;
    return 32;
}

template <typename IdxT, bool DOT>
static int launch_spmv(int lpr, int64_t nrows, const IdxT* row_ptr, const int32_t* col_idx, const double* vals,
                       const double* x, double* y, const uint8_t* mask, double* partials, const double* scal,
                       int* grid_out, cudaStream_t st) {
#define FE_SPMV_CASE(L)                                                                                          \
    case L: {                                                                                                    \
        int g = 0;                                                                                               \
        int rc = persistent_grid((const void*)k_spmv<IdxT, L, DOT>, kVecThreads, &g);                            \
        if (rc != FE_OK) return rc;                                                                              \
        const int64_t need = (nrows + (kVecThreads / L) - 1) / (kVecThreads / L);                                \
        if (!DOT && need < g) g = (int)(need > 0 ? need : 1);                                                    \
        k_spmv<IdxT, L, DOT><<<g, kVecThreads, 0, st>>>(nrows, row_ptr, col_idx, vals, x, y, mask, partials, scal); \
        if (grid_out) *grid_out = g;                                                                             \
        break;                                                                                                   \
    }
    switch (lpr) {
        FE_SPMV_CASE(2)
        FE_SPMV_CASE(4)
        FE_SPMV_CASE(8)
        FE_SPMV_CASE(16)
        FE_SPMV_CASE(32)
        default:
            set_error("spmv: bad lanes-per-row %d", lpr);
            return FE_ERR_INVALID;
    }
#undef FE_SPMV_CASE
    FE_CHECK_LAUNCH("k_spmv");
    return FE_OK;
}

template <bool DOT>
static int launch_spmv_nodal(int64_t nnode, const int32_t* nptr, const int32_t* nadj, const double* vals,
                             const double* x, double* y, const uint8_t* mask, double* partials, const double* scal,
                             int* grid_out, cudaStream_t st) {
    constexpr int lanes = 4;   // measured on B200, 9 blocks per node row: 4 lanes 664 it/s, 2 lanes 638, 1 lane 564
#define FE_NODAL_CASE(LN)                                                                                      \
    case LN: {                                                                                                  \
        int g = 0;                                                                                              \
        int rc = persistent_grid((const void*)k_spmv_nodal<LN, DOT>, kVecThreads, &g);                          \
        if (rc != FE_OK) return rc;                                                                             \
        const int64_t need = (nnode + (kVecThreads / LN) - 1) / (kVecThreads / LN);                             \
        if (!DOT && need < g) g = (int)(need > 0 ? need : 1);                                                   \
        k_spmv_nodal<LN, DOT><<<g, kVecThreads, 0, st>>>(nnode, nptr, nadj, vals, x, y, mask, partials, scal);  \
        if (grid_out) *grid_out = g;                                                                            \
        break;                                                                                                  \
    }
    switch (lanes) {
        FE_NODAL_CASE(4)
    }
#undef FE_NODAL_CASE
    FE_CHECK_LAUNCH("k_spmv_nodal");
    return FE_OK;
}

// Matrix handle used by the solvers: CSR always, node-level pattern when the caller has it (ndof = 2).
template <typename IdxT> struct Mat {
    int64_t nrows;
    const IdxT* row_ptr;
    const int32_t* col_idx;
    const double* vals;
    const int32_t* nptr;   // optional
    const int32_t* nadj;
    int lpr;
};

template <typename IdxT, bool DOT>
static int mat_spmv(const Mat<IdxT>& A, const double* x, double* y, const uint8_t* mask, double* partials,
                    const double* scal, int* grid_out, cudaStream_t st) {
    if (A.nptr && A.nadj && (A.nrows % 2) == 0)
        return launch_spmv_nodal<DOT>(A.nrows / 2, A.nptr, A.nadj, A.vals, x, y, mask, partials, scal, grid_out, st);
    return launch_spmv<IdxT, DOT>(A.lpr, A.nrows, A.row_ptr, A.col_idx, A.vals, x, y, mask, partials, scal, grid_out,
                                  st);
}

struct PcgWork {
    double *r, *p, *Ap, *minv, *scal, *partA, *partB;
};

static PcgWork carve(void* work, int64_t nrows) {
    PcgWork w;
    double* d = static_cast<double*>(work);
    const int64_t n = (nrows + 31) / 32 * 32;
    w.r = d;
    w.p = d + n;
    w.Ap = d + 2 * n;
    w.minv = d + 3 * n;
    w.scal = d + 4 * n;
    w.partA = w.scal + kScalarDoubles;
    w.partB = w.partA + kMaxPartials;
    return w;
}

constexpr int64_t kGraphMaxNnz = (int64_t)1 << 27;      // above this an iteration is long enough to hide its launches
template <typename IdxT>
static int pcg_impl(int64_t nrows, const IdxT* row_ptr, const int32_t* col_idx, const double* vals, const double* rhs,
                    const uint8_t* mask, double* x, double rtol, double atol, int64_t maxit, int check_every,
                    void* work, double* info, const int32_t* nptr, const int32_t* nadj, int64_t nnz, cudaStream_t st) {
    PcgWork w = carve(work, nrows);
    const int lpr = pick_lpr(nrows, nnz);
    const Mat<IdxT> A{nrows, row_ptr, col_idx, vals, nptr, nadj, lpr};
    int gv = 0, rc;
    if ((rc = persistent_grid((const void*)k_pcg_update, kVecThreads, &gv)) != FE_OK) return rc;
    const int64_t need = (nrows + kVecThreads - 1) / kVecThreads;
    if (need < gv) gv = (int)(need > 0 ? need : 1);

    // setup: minv <- diag; x0; r0 = b - A x0; p0
    k_csr_diagonal<IdxT><<<grid_for(nrows, 256), 256, 0, st>>>(nrows, row_ptr, col_idx, vals, w.minv);
    FE_CHECK_LAUNCH("k_csr_diagonal");
    k_pcg_init1<<<gv, kVecThreads, 0, st>>>(nrows, rhs, mask, x, w.minv);
    FE_CHECK_LAUNCH("k_pcg_init1");
    if ((rc = mat_spmv<IdxT, false>(A, x, w.Ap, mask, nullptr, nullptr, nullptr, st)) != FE_OK) return rc;
    k_pcg_init2<<<gv, kVecThreads, 0, st>>>(nrows, rhs, mask, w.Ap, w.minv, w.r, w.p, w.partB);
    FE_CHECK_LAUNCH("k_pcg_init2");
    k_pcg_init3<<<1, kVecThreads, 0, st>>>(gv, w.partB, rtol, atol, w.scal, nullptr, PeerRed{}, 0);
    FE_CHECK_LAUNCH("k_pcg_init3");

    auto iterate = [&](cudaStream_t s) -> int {          // one PCG iteration = three launches, same arguments every time
        int ga = 0;
        const int r = mat_spmv<IdxT, true>(A, w.p, w.Ap, mask, w.partA, w.scal, &ga, s);
        if (r != FE_OK) return r;
        k_pcg_update<<<gv, kVecThreads, 0, s>>>(nrows, ga, w.partA, w.p, w.Ap, w.minv, x, w.r, w.partB, w.scal, nullptr,
                                                PeerRed{}, 0);
        k_pcg_direction<<<gv, kVecThreads, 0, s>>>(nrows, gv, w.partB, w.r, w.minv, w.p, w.p, w.scal, nullptr, PeerRed{},
                                                   0);
        return FE_OK;
    };
    // Small systems are bound by launch latency (three launches of a few microseconds each per iteration): a batch of
    // check_every iterations is captured once into a CUDA graph (on a private stream: the caller's may be the legacy
    // default stream, which cannot capture) and replayed on the caller's stream.  Large systems launch directly.
    struct Replay {
        cudaStream_t cs = nullptr;
        cudaGraphExec_t exec = nullptr;
        ~Replay() {
            if (exec) cudaGraphExecDestroy(exec);
            if (cs) cudaStreamDestroy(cs);
        }
    } replay;
    if (nnz <= kGraphMaxNnz && check_every >= 4 && maxit >= check_every) {
        if (cudaStreamCreateWithFlags(&replay.cs, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamBeginCapture(replay.cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rcc = FE_OK;
            for (int k = 0; k < check_every && rcc == FE_OK; ++k) rcc = iterate(replay.cs);
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(replay.cs, &graph);
            if (rcc == FE_OK && e == cudaSuccess && graph &&
                cudaGraphInstantiate(&replay.exec, graph, 0) != cudaSuccess)
                replay.exec = nullptr;
            if (graph) cudaGraphDestroy(graph);
        }
        (void)cudaGetLastError();        // a failed capture only means direct launches
    }

    double hs[kScalarDoubles];
    int64_t launched = 0;
    for (;;) {
        FE_CUDA_TRY(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
        FE_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs[S_DONE_C] != 0.0 || hs[S_ERR] != 0.0 || launched >= maxit) break;
        int64_t batch = check_every;
        if (launched + batch > maxit) batch = maxit - launched;
        if (replay.exec && batch == check_every) {
            FE_CUDA_TRY(cudaGraphLaunch(replay.exec, st));
        } else {
            for (int64_t k = 0; k < batch; ++k)
                if ((rc = iterate(st)) != FE_OK) return rc;
        }
        FE_CHECK_LAUNCH("pcg iteration");
        launched += batch;
    }
    fill_info(hs, info);
    return FE_OK;
}

}  // namespace fe

extern "C" {

int fe_spmv(int64_t nrows, int64_t nnz, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
            const double* x, double* y, const uint8_t* rowmask, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows >= 0 && nnz >= 0 && row_ptr && col_idx && vals && x && y, "null pointer or negative size");
    if (nrows == 0) return FE_OK;
    cudaStream_t st = as_stream(stream);
    const int lpr = pick_lpr(nrows, nnz);   // lanes per row from the average row length (nnz is a host argument: no read-back)
    if (idx64)
        return launch_spmv<int64_t, false>(lpr, nrows, static_cast<const int64_t*>(row_ptr), col_idx, vals, x, y,
                                           rowmask, nullptr, nullptr, nullptr, st);
    return launch_spmv<int32_t, false>(lpr, nrows, static_cast<const int32_t*>(row_ptr), col_idx, vals, x, y, rowmask,
                                       nullptr, nullptr, nullptr, st);
}

int fe_spmv_nodal(int64_t nnode, const int32_t* nptr, const int32_t* nadj, const double* vals, const double* x,
                  double* y, const uint8_t* rowmask, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0, "negative size");
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(nptr && nadj && vals && x && y, "null pointer");
    return launch_spmv_nodal<false>(nnode, nptr, nadj, vals, x, y, rowmask, nullptr, nullptr, nullptr,
                                    as_stream(stream));
}

int fe_csr_diagonal(int64_t nrows, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
                    double* diag, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows >= 0 && row_ptr && col_idx && vals && diag, "null pointer or negative size");
    if (nrows == 0) return FE_OK;
    cudaStream_t st = as_stream(stream);
    if (idx64)
        k_csr_diagonal<int64_t><<<grid_for(nrows, 256), 256, 0, st>>>(nrows, static_cast<const int64_t*>(row_ptr),
                                                                      col_idx, vals, diag);
    else
        k_csr_diagonal<int32_t><<<grid_for(nrows, 256), 256, 0, st>>>(nrows, static_cast<const int32_t*>(row_ptr),
                                                                      col_idx, vals, diag);
    FE_CHECK_LAUNCH("k_csr_diagonal");
    return FE_OK;
}

size_t fe_pcg_workspace_bytes(int64_t nrows) {
    const int64_t n = (nrows + 31) / 32 * 32;
    return (size_t)(4 * n + fe::kScalarDoubles + 3 * fe::kMaxPartials) * sizeof(double);
}

int fe_pcg(int64_t nrows, int64_t nnz, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
           const double* rhs, const uint8_t* dirmask, double* x, double rtol, double atol, int64_t maxit,
           int32_t check_every, void* work, double* info, const int32_t* nptr, const int32_t* nadj, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows >= 0 && nnz >= 0 && row_ptr && col_idx && vals && rhs && x && work, "null pointer or negative size");
    FE_REQUIRE(maxit >= 0 && check_every >= 1, "maxit must be >= 0 and check_every >= 1");
    if (nrows == 0) {
        if (info) info[0] = info[1] = info[2] = 0.0, info[3] = 1.0;
        return FE_OK;
    }
    cudaStream_t st = as_stream(stream);
    if (idx64)
        return pcg_impl<int64_t>(nrows, static_cast<const int64_t*>(row_ptr), col_idx, vals, rhs, dirmask, x, rtol,
                                 atol, maxit, check_every, work, info, nptr, nadj, nnz, st);
    return pcg_impl<int32_t>(nrows, static_cast<const int32_t*>(row_ptr), col_idx, vals, rhs, dirmask, x, rtol, atol,
                             maxit, check_every, work, info, nptr, nadj, nnz, st);
}

}  // extern "C"

#include "fe_dist.cuh"
