// K7: multi-GPU glue for the PCG solve (included at the end of fe_solver.cu so it can reuse its kernels).
// One host process per GPU; rank r owns a block of rows (owned dofs first in its local numbering, ghost
// dofs after them, grouped by owning rank).  Per iteration: one halo exchange of the search direction
// (pack kernel + grouped ncclSend/ncclRecv straight into the ghost segment), then two scalar all-reduces
// (p.Ap and {r.z, r.r}) -- every rank applies the same global scalars, so all ranks take identical steps.
#pragma once
#ifdef FE_WITH_NCCL
#include <nccl.h>

namespace fe {

#define FE_NCCL_TRY(expr)                                                                   \
    do {                                                                                    \
        ncclResult_t _r = (expr);                                                           \
        if (_r != ncclSuccess) {                                                            \
            set_error("NCCL error %d (%s) in %s", (int)_r, ncclGetErrorString(_r), #expr);  \
            return FE_ERR_NCCL;                                                             \
        }                                                                                   \
    } while (0)

__global__ void k_pack(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ src,
                       double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// out[k] = sum_i part[i*stride + k], k < nvals; one block, fixed order
__global__ void __launch_bounds__(kVecThreads) k_reduce_partials(int nb, const double* __restrict__ part, int stride,
                                                                 int nvals, double* __restrict__ out) {
    __shared__ double sm[kVecThreads / 32];
    for (int k = 0; k < nvals; ++k) {
        const double v = sum_partials(part + k, nb, stride, sm);
        if (threadIdx.x == 0) out[k] = v;
    }
}

struct Halo {
    ncclComm_t comm;
    int n_nbr;
    const int32_t* nbr_rank;   // host
    const int64_t* send_off;   // host, n_nbr + 1, offsets into send_idx / sendbuf
    const int32_t* send_idx;   // device, owned dof indices to send, neighbour-major
    double* sendbuf;           // device
    const int64_t* recv_off;   // host, n_nbr + 1, offsets into the vector (ghost segments)
};

static int halo_exchange(const Halo& h, double* vec, cudaStream_t st) {
    const int64_t nsend = h.send_off[h.n_nbr];
    if (nsend > 0) {
        k_pack<<<grid_for(nsend, 256), 256, 0, st>>>(nsend, h.send_idx, vec, h.sendbuf);
        FE_CHECK_LAUNCH("k_pack");
    }
    FE_NCCL_TRY(ncclGroupStart());
    for (int k = 0; k < h.n_nbr; ++k) {
        const int64_t ns = h.send_off[k + 1] - h.send_off[k], nr = h.recv_off[k + 1] - h.recv_off[k];
        if (ns > 0) FE_NCCL_TRY(ncclSend(h.sendbuf + h.send_off[k], (size_t)ns, ncclDouble, h.nbr_rank[k], h.comm, st));
        if (nr > 0) FE_NCCL_TRY(ncclRecv(vec + h.recv_off[k], (size_t)nr, ncclDouble, h.nbr_rank[k], h.comm, st));
    }
    FE_NCCL_TRY(ncclGroupEnd());
    return FE_OK;
}

template <typename IdxT>
static int dist_pcg_impl(int64_t nown, int64_t nloc, const IdxT* row_ptr, const int32_t* col_idx, const double* vals,
                         const double* rhs, const uint8_t* mask, double* x, double rtol, double atol, int64_t maxit,
                         int check_every, void* work, double* info, const Halo& h, const int32_t* nptr,
                         const int32_t* nadj, int64_t nnz_owned, cudaStream_t st) {
    PcgWork w = carve(work, nloc);
    double* g = w.partB + 2 * kMaxPartials;  // 8 doubles of globally reduced scalars
    const int lpr = pick_lpr(nown, nnz_owned);
    const Mat<IdxT> A{nown, row_ptr, col_idx, vals, nptr, nadj, lpr};
    int gv = 0, rc;
    if ((rc = persistent_grid((const void*)k_pcg_update, kVecThreads, &gv)) != FE_OK) return rc;
    const int64_t need = (nown + kVecThreads - 1) / kVecThreads;
    if (need < gv) gv = (int)(need > 0 ? need : 1);

    if (nown > 0) {
        k_csr_diagonal<IdxT><<<grid_for(nown, 256), 256, 0, st>>>(nown, row_ptr, col_idx, vals, w.minv);
        FE_CHECK_LAUNCH("k_csr_diagonal");
    }
    k_pcg_init1<<<gv, kVecThreads, 0, st>>>(nown, rhs, mask, x, w.minv);
    FE_CHECK_LAUNCH("k_pcg_init1");
    if ((rc = halo_exchange(h, x, st)) != FE_OK) return rc;   // Dirichlet values of ghost dofs
    if ((rc = mat_spmv<IdxT, false>(A, x, w.Ap, mask, nullptr, nullptr, nullptr, st)) != FE_OK) return rc;
    k_pcg_init2<<<gv, kVecThreads, 0, st>>>(nown, rhs, mask, w.Ap, w.minv, w.r, w.p, w.partB);
    FE_CHECK_LAUNCH("k_pcg_init2");
    k_reduce_partials<<<1, kVecThreads, 0, st>>>(gv, w.partB, 2, 2, g);
    FE_NCCL_TRY(ncclAllReduce(g, g, 2, ncclDouble, ncclSum, h.comm, st));
    k_pcg_init3<<<1, kVecThreads, 0, st>>>(gv, w.partB, rtol, atol, w.scal, g, PeerRed{}, 0);
    FE_CHECK_LAUNCH("k_pcg_init3");

    double hs[kScalarDoubles];
    int64_t launched = 0;
    for (;;) {
        FE_CUDA_TRY(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
        FE_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs[S_DONE_C] != 0.0 || hs[S_ERR] != 0.0 || launched >= maxit) break;   // identical on all ranks (same global scalars)
        int64_t batch = check_every;
        if (launched + batch > maxit) batch = maxit - launched;
        for (int64_t k = 0; k < batch; ++k) {
            int ga = 0;
            if ((rc = halo_exchange(h, w.p, st)) != FE_OK) return rc;
            if ((rc = mat_spmv<IdxT, true>(A, w.p, w.Ap, mask, w.partA, w.scal, &ga, st)) != FE_OK) return rc;
            k_reduce_partials<<<1, kVecThreads, 0, st>>>(ga, w.partA, 1, 1, g + 2);
            FE_NCCL_TRY(ncclAllReduce(g + 2, g + 2, 1, ncclDouble, ncclSum, h.comm, st));
            k_pcg_update<<<gv, kVecThreads, 0, st>>>(nown, ga, w.partA, w.p, w.Ap, w.minv, x, w.r, w.partB, w.scal,
                                                     g + 2, PeerRed{}, 0);
            k_reduce_partials<<<1, kVecThreads, 0, st>>>(gv, w.partB, 2, 2, g + 4);
            FE_NCCL_TRY(ncclAllReduce(g + 4, g + 4, 2, ncclDouble, ncclSum, h.comm, st));
            k_pcg_direction<<<gv, kVecThreads, 0, st>>>(nown, gv, w.partB, w.r, w.minv, w.p, w.p, w.scal, g + 4,
                                                        PeerRed{}, 0);
        }
        FE_CHECK_LAUNCH("dist pcg iteration");
        launched += batch;
    }
    if ((rc = halo_exchange(h, x, st)) != FE_OK) return rc;   // ghost part of the solution
    fill_info(hs, info);
    return FE_OK;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory halo (NVLink / NVSwitch): the SpMV reads the ghost entries of the search direction straight
// out of the owning GPU's memory -- no pack kernel, no send/recv, no ghost copy.  The direction vector is
// double-buffered in symmetric memory (p_k lives in half k & 1); after writing p_{k+1} a rank stores the
// version number into each neighbour's flag array, and only lanes that touch a ghost column wait for the
// owner's flag, so interior rows overlap the neighbours' tail.  WAR safety: p_{k+2} overwrites p_k only after
// the all-reduce of iteration k+1, which every neighbour joins after finishing its reads of p_k.
// ------------------------------------------------------------------------------------------------
struct PeerHalo {
    double* const* peer_p;        // device array [world]: base of every rank's double-buffered direction vector
    long long* const* peer_flags; // device array [world]: every rank's flag array (int64 [world])
    const int32_t* ghost_rank;    // device [ghost nodes]: owning rank
    const int32_t* ghost_ridx;    // device [ghost nodes]: node index inside the owner
    int64_t pstride;              // doubles between the two halves of a direction buffer (same on all ranks)
    int rank, world;
    int n_nbr;
    const int32_t* nbr_dev;       // device [n_nbr]: neighbour ranks (for the signal kernel)
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Waits until *flag >= need (a peer publishes it with st.release.sys).  Exponential __nanosleep back-off, a
// wall-clock bound (kPeerTimeoutNs) instead of a spin count, and an early exit once any waiter of this solve has
// already recorded an error -- a dead or absent peer costs one time-out, not one per waiter.
constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;   // 2 s
__device__ __forceinline__ bool peer_wait(const long long* flag, long long need, const double* scal) {
    if (ld_acquire_sys_s64(flag) >= need) return true;
    const unsigned long long t0 = global_ns();
    unsigned ns = 32;
    for (;;) {
        if (ld_acquire_sys_s64(flag) >= need) return true;
        if (scal && *reinterpret_cast<const volatile double*>(scal + S_ERR) != 0.0) return false;
        if (global_ns() - t0 > kPeerTimeoutNs) return false;
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
    }
}

__device__ __forceinline__ double2 ld_peer_f64x2(const double2* p) {
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int L, bool DOT>
__global__ void __launch_bounds__(kVecThreads) k_spmv_nodal_p2p(
    int64_t nnode, int64_t nown_nodes, const int32_t* __restrict__ nptr, const int32_t* __restrict__ nadj,
    const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
    const uint8_t* __restrict__ rowmask, double* __restrict__ partials, double* __restrict__ scal, PeerHalo h,
    int parity, long long need) {
    __shared__ double sm[kVecThreads / 32];
    if (DOT && scal[S_DONE_C] != 0.0) return;
    constexpr int GPB = kVecThreads / L;
    const int lane = threadIdx.x % L;
    const int grp = threadIdx.x / L;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const long long* myflags = h.peer_flags[h.rank];
    double dot = 0.0;
    for (int64_t n0 = (int64_t)blockIdx.x * GPB; n0 < nnode; n0 += (int64_t)gridDim.x * GPB) {
        const int64_t n = n0 + grp;
        double a0 = 0.0, a1 = 0.0;
        if (n < nnode) {
            const int p0 = __ldg(nptr + n), deg = __ldg(nptr + n + 1) - p0;
            const double2* r0 = reinterpret_cast<const double2*>(vals + 4 * (int64_t)p0);
            const double2* r1 = r0 + deg;
            for (int p = lane; p < deg; p += L) {
                const int m = __ldg(nadj + p0 + p);
                double2 xv;
                if (m < nown_nodes) {
                    xv = x2[m];
                } else {                                  // ghost column: the owner's copy over NVLink
                    const int g = m - (int)nown_nodes;
                    const int q = __ldg(h.ghost_rank + g);
                    if (!peer_wait(myflags + q, need, scal)) scal[S_ERR] = (double)kErrPeerTimeout;   // host reports it
                    const double2* src = reinterpret_cast<const double2*>(h.peer_p[q] + (int64_t)parity * h.pstride);
                    xv = ld_peer_f64x2(src + __ldg(h.ghost_ridx + g));
                }
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
                const double2 xn = x2[n];
                dot = fma(xn.x, a0, fma(xn.y, a1, dot));
            }
        }
    }
    if (DOT) {
        dot = block_sum(dot, sm);
        if (threadIdx.x == 0) partials[blockIdx.x] = dot;
    }
}

// publishes "direction vector version `version` is complete" to every neighbour (runs after the kernel that
// wrote it, so the data is already in this GPU's L2, which is where peer loads are served from)
__global__ void k_signal_neighbours(PeerHalo h, long long version) {
    const int k = threadIdx.x;
    if (k >= h.n_nbr) return;
    st_release_sys_s64(h.peer_flags[h.nbr_dev[k]] + h.rank, version);
}

// local sums of `nvals` scalars -> slot [seq & 1][rank] of every rank, then the version (see PeerRed)
__global__ void __launch_bounds__(kVecThreads) k_reduce_push(int nb, const double* __restrict__ part, int stride,
                                                             int nvals, const double* __restrict__ scal, PeerRed pr,
                                                             long long seq) {
    __shared__ double sm[kVecThreads / 32];
    if (scal && scal[S_DONE_C] != 0.0) return;   // converged: the consumers return early too (no slot reuse race)
    double v[2] = {0.0, 0.0};
    for (int k = 0; k < nvals; ++k) v[k] = sum_partials(part + k, nb, stride, sm);
    const int q = threadIdx.x;
    if (q < pr.world) {
        double* slot = pr.slots[q] + (seq & 1) * (2 * kMaxRanks) + 2 * pr.rank;
        for (int k = 0; k < nvals; ++k) slot[k] = v[k];
        st_release_sys_s64(pr.versions[q] + pr.rank, seq + 1);
    }
}

template <typename IdxT>
static int dist_pcg_p2p_impl(int64_t nown, int64_t nloc, const IdxT* row_ptr, const int32_t* col_idx,
                             const double* vals, const double* rhs, const uint8_t* mask, double* x, double rtol,
                             double atol, int64_t maxit, int check_every, void* work, double* info, const Halo& h,
                             const int32_t* nptr, const int32_t* nadj, const PeerHalo& ph, double* pbuf,
                             long long epoch, const PeerRed& pr, long long seq, cudaStream_t st) {
    PcgWork w = carve(work, nloc);
    double* g = w.partB + 2 * kMaxPartials;
    int gv = 0, rc;
    if ((rc = persistent_grid((const void*)k_pcg_update, kVecThreads, &gv)) != FE_OK) return rc;
    const int64_t need = (nown + kVecThreads - 1) / kVecThreads;
    if (need < gv) gv = (int)(need > 0 ? need : 1);
    int gs = 0;
    if ((rc = persistent_grid((const void*)k_spmv_nodal_p2p<4, true>, kVecThreads, &gs)) != FE_OK) return rc;
    const Mat<IdxT> A{nown, row_ptr, col_idx, vals, nptr, nadj, 4};
    double* pcur = pbuf;                       // half 0 holds p_0
    double* pnext = pbuf + ph.pstride;
    // stream-ordered barrier over all ranks: nobody may still be reading this rank's direction buffer or
    // reduction slots on behalf of the previous solve when this one starts overwriting them
    FE_CUDA_TRY(cudaMemsetAsync(g, 0, sizeof(double), st));
    FE_NCCL_TRY(ncclAllReduce(g, g, 1, ncclDouble, ncclSum, h.comm, st));

    if (nown > 0) {
        k_csr_diagonal<IdxT><<<grid_for(nown, 256), 256, 0, st>>>(nown, row_ptr, col_idx, vals, w.minv);
        FE_CHECK_LAUNCH("k_csr_diagonal");
    }
    k_pcg_init1<<<gv, kVecThreads, 0, st>>>(nown, rhs, mask, x, w.minv);
    FE_CHECK_LAUNCH("k_pcg_init1");
    if ((rc = halo_exchange(h, x, st)) != FE_OK) return rc;   // Dirichlet values of ghost dofs (once, NCCL)
    if ((rc = mat_spmv<IdxT, false>(A, x, w.Ap, mask, nullptr, nullptr, nullptr, st)) != FE_OK) return rc;
    k_pcg_init2<<<gv, kVecThreads, 0, st>>>(nown, rhs, mask, w.Ap, w.minv, w.r, pcur, w.partB);
    FE_CHECK_LAUNCH("k_pcg_init2");
    k_reduce_push<<<1, kVecThreads, 0, st>>>(gv, w.partB, 2, 2, nullptr, pr, seq);
    k_pcg_init3<<<1, kVecThreads, 0, st>>>(gv, w.partB, rtol, atol, w.scal, g, pr, seq);
    FE_CHECK_LAUNCH("k_pcg_init3");
    ++seq;
    if (ph.n_nbr > 0) k_signal_neighbours<<<1, 32, 0, st>>>(ph, epoch + 1);   // p_0 is version epoch + 1

    double hs[kScalarDoubles];
    int64_t launched = 0;
    for (;;) {
        FE_CUDA_TRY(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
        FE_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs[S_ERR] == (double)kErrPeerTimeout) {
            set_error("fe_dist_pcg_p2p: timed out waiting for a neighbour's direction vector or reduction slot");
            return FE_ERR_NCCL;
        }
        if (hs[S_DONE_C] != 0.0 || hs[S_ERR] != 0.0 || launched >= maxit) break;
        int64_t batch = check_every;
        if (launched + batch > maxit) batch = maxit - launched;
        for (int64_t k = 0; k < batch; ++k) {
            const int64_t it = launched + k;
            const int parity = (int)(it & 1);
            k_spmv_nodal_p2p<4, true><<<gs, kVecThreads, 0, st>>>(nown / 2, nown / 2, nptr, nadj, vals, pcur, w.Ap, mask,
                                                                 w.partA, w.scal, ph, parity, epoch + 1 + it);
            k_reduce_push<<<1, kVecThreads, 0, st>>>(gs, w.partA, 1, 1, w.scal, pr, seq);
            k_pcg_update<<<gv, kVecThreads, 0, st>>>(nown, gs, w.partA, pcur, w.Ap, w.minv, x, w.r, w.partB, w.scal,
                                                     g + 2, pr, seq);
            k_reduce_push<<<1, kVecThreads, 0, st>>>(gv, w.partB, 2, 2, w.scal, pr, seq + 1);
            k_pcg_direction<<<gv, kVecThreads, 0, st>>>(nown, gv, w.partB, w.r, w.minv, pcur, pnext, w.scal, g + 4, pr,
                                                        seq + 1);
            seq += 2;
            if (ph.n_nbr > 0) k_signal_neighbours<<<1, 32, 0, st>>>(ph, epoch + 2 + it);   // p_{it+1}
            double* t = pcur;
            pcur = pnext;
            pnext = t;
        }
        FE_CHECK_LAUNCH("dist pcg (p2p) iteration");
        launched += batch;
    }
    if ((rc = halo_exchange(h, x, st)) != FE_OK) return rc;
    fill_info(hs, info);
    return FE_OK;
}

}  // namespace fe
#endif  // FE_WITH_NCCL

extern "C" {

int fe_dist_available(void) {
#ifdef FE_WITH_NCCL
    return 1;
#else
    return 0;
#endif
}

int fe_dist_unique_id(uint8_t* out128) {
    using namespace fe;
    FE_REQUIRE(out128, "null pointer");
#ifdef FE_WITH_NCCL
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    FE_NCCL_TRY(ncclGetUniqueId(&id));
    memcpy(out128, &id, 128);
    return FE_OK;
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

int fe_dist_comm_init(const uint8_t* id128, int rank, int nranks, void** comm_out) {
    using namespace fe;
    FE_REQUIRE(id128 && comm_out && rank >= 0 && rank < nranks, "bad argument");
#ifdef FE_WITH_NCCL
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c;
    FE_NCCL_TRY(ncclCommInitRank(&c, nranks, id, rank));
    *comm_out = c;
    return FE_OK;
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

int fe_dist_comm_destroy(void* comm) {
    using namespace fe;
#ifdef FE_WITH_NCCL
    if (comm) FE_NCCL_TRY(ncclCommDestroy(static_cast<ncclComm_t>(comm)));
    return FE_OK;
#else
    return FE_OK;
#endif
}

int fe_dist_halo_exchange(void* comm, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_off,
                          const int32_t* send_idx, double* sendbuf, const int64_t* recv_off, double* vec,
                          void* stream) {
    using namespace fe;
    FE_REQUIRE(comm && n_nbr >= 0 && vec, "bad argument");
#ifdef FE_WITH_NCCL
    if (n_nbr == 0) return FE_OK;
    FE_REQUIRE(nbr_rank && send_off && recv_off, "null pointer");
    Halo h{static_cast<ncclComm_t>(comm), n_nbr, nbr_rank, send_off, send_idx, sendbuf, recv_off};
    return halo_exchange(h, vec, as_stream(stream));
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

int fe_dist_pcg_p2p(int64_t nrows_owned, int64_t nrows_local, int idx64, const void* row_ptr, const int32_t* col_idx,
                    const double* vals, const double* rhs, const uint8_t* dirmask, double* x, double rtol, double atol,
                    int64_t maxit, int32_t check_every, void* work, double* info, void* comm, int32_t n_nbr,
                    const int32_t* nbr_rank, const int64_t* send_off, const int32_t* send_idx, double* sendbuf,
                    const int64_t* recv_off, const int32_t* nptr, const int32_t* nadj, int32_t rank, int32_t world,
                    const void* peer_p, const void* peer_flags, const int32_t* ghost_rank, const int32_t* ghost_ridx,
                    const int32_t* nbr_dev, int64_t pstride, double* pbuf, int64_t epoch, const void* peer_slots,
                    const void* peer_versions, int64_t seq, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows_owned >= 0 && nrows_local >= nrows_owned && (nrows_owned % 2) == 0 && row_ptr && col_idx && vals &&
                   rhs && x && work && comm && nptr && nadj && peer_p && peer_flags && pbuf && peer_slots &&
                   peer_versions && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && seq >= 0,
               "null pointer, bad size or more ranks than the peer tables hold (16)");
    FE_REQUIRE(maxit >= 0 && check_every >= 1 && n_nbr >= 0 && n_nbr <= 32 && pstride >= nrows_local,
               "bad iteration / neighbour / stride arguments");
    FE_REQUIRE(n_nbr == 0 || (nbr_rank && send_off && recv_off && ghost_rank && ghost_ridx && nbr_dev),
               "null halo description");
#ifdef FE_WITH_NCCL
    static const int64_t zero2[2] = {0, 0};
    Halo h{static_cast<ncclComm_t>(comm), n_nbr, nbr_rank, n_nbr ? send_off : zero2, send_idx, sendbuf,
           n_nbr ? recv_off : zero2};
    PeerHalo ph{static_cast<double* const*>(peer_p), static_cast<long long* const*>(peer_flags), ghost_rank, ghost_ridx,
                pstride, rank, world, n_nbr, nbr_dev};
    PeerRed pr{static_cast<double* const*>(peer_slots), static_cast<long long* const*>(peer_versions), rank, world};
    cudaStream_t st = as_stream(stream);
    if (idx64)
        return dist_pcg_p2p_impl<int64_t>(nrows_owned, nrows_local, static_cast<const int64_t*>(row_ptr), col_idx, vals,
                                          rhs, dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj,
                                          ph, pbuf, epoch, pr, seq, st);
    return dist_pcg_p2p_impl<int32_t>(nrows_owned, nrows_local, static_cast<const int32_t*>(row_ptr), col_idx, vals, rhs,
                                      dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj, ph, pbuf,
                                      epoch, pr, seq, st);
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

size_t fe_dist_pcg_workspace_bytes(int64_t nrows_local) { return fe_pcg_workspace_bytes(nrows_local) + 64; }

int fe_dist_pcg(int64_t nrows_owned, int64_t nrows_local, int64_t nnz_owned, int idx64, const void* row_ptr,
                const int32_t* col_idx, const double* vals, const double* rhs, const uint8_t* dirmask, double* x,
                double rtol, double atol, int64_t maxit, int32_t check_every, void* work, double* info, void* comm,
                int32_t n_nbr,
                const int32_t* nbr_rank, const int64_t* send_off, const int32_t* send_idx, double* sendbuf,
                const int64_t* recv_off, const int32_t* nptr, const int32_t* nadj, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows_owned >= 0 && nrows_local >= nrows_owned && nnz_owned >= 0 && row_ptr && col_idx && vals && rhs &&
                   x && work && comm,
               "null pointer or bad size");
    FE_REQUIRE(maxit >= 0 && check_every >= 1 && n_nbr >= 0, "maxit must be >= 0 and check_every >= 1");
    FE_REQUIRE(n_nbr == 0 || (nbr_rank && send_off && recv_off), "null halo description");
#ifdef FE_WITH_NCCL
    static const int64_t zero2[2] = {0, 0};
    Halo h{static_cast<ncclComm_t>(comm), n_nbr, nbr_rank, n_nbr ? send_off : zero2, send_idx, sendbuf,
           n_nbr ? recv_off : zero2};
    cudaStream_t st = as_stream(stream);
    if (idx64)
        return dist_pcg_impl<int64_t>(nrows_owned, nrows_local, static_cast<const int64_t*>(row_ptr), col_idx, vals,
                                      rhs, dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj,
                                      nnz_owned, st);
    return dist_pcg_impl<int32_t>(nrows_owned, nrows_local, static_cast<const int32_t*>(row_ptr), col_idx, vals, rhs,
                                  dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj, nnz_owned, st);
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

}  // extern "C"
