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
                         const int32_t* nadj, cudaStream_t st) {
    PcgWork w = carve(work, nloc);
    double* g = w.partB + 2 * kMaxPartials;  // 8 doubles of globally reduced scalars
    IdxT last;
    FE_CUDA_TRY(cudaMemcpyAsync(&last, row_ptr + nown, sizeof(IdxT), cudaMemcpyDeviceToHost, st));
    FE_CUDA_TRY(cudaStreamSynchronize(st));
    const int lpr = pick_lpr(nown, (int64_t)last);
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
    k_pcg_init3<<<1, kVecThreads, 0, st>>>(gv, w.partB, rtol, atol, w.scal, g);
    FE_CHECK_LAUNCH("k_pcg_init3");

    double hs[kScalarDoubles];
    int64_t launched = 0;
    for (;;) {
        FE_CUDA_TRY(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
        FE_CUDA_TRY(cudaStreamSynchronize(st));
        if (hs[S_DONE_C] != 0.0 || launched >= maxit) break;   // identical on all ranks (same global scalars)
        int64_t batch = check_every;
        if (launched + batch > maxit) batch = maxit - launched;
        for (int64_t k = 0; k < batch; ++k) {
            int ga = 0;
            if ((rc = halo_exchange(h, w.p, st)) != FE_OK) return rc;
            if ((rc = mat_spmv<IdxT, true>(A, w.p, w.Ap, mask, w.partA, w.scal, &ga, st)) != FE_OK) return rc;
            k_reduce_partials<<<1, kVecThreads, 0, st>>>(ga, w.partA, 1, 1, g + 2);
            FE_NCCL_TRY(ncclAllReduce(g + 2, g + 2, 1, ncclDouble, ncclSum, h.comm, st));
            k_pcg_update<<<gv, kVecThreads, 0, st>>>(nown, ga, w.partA, w.p, w.Ap, w.minv, x, w.r, w.partB, w.scal,
                                                     g + 2);
            k_reduce_partials<<<1, kVecThreads, 0, st>>>(gv, w.partB, 2, 2, g + 4);
            FE_NCCL_TRY(ncclAllReduce(g + 4, g + 4, 2, ncclDouble, ncclSum, h.comm, st));
            k_pcg_direction<<<gv, kVecThreads, 0, st>>>(nown, gv, w.partB, w.r, w.minv, w.p, w.scal, g + 4);
        }
        FE_CHECK_LAUNCH("dist pcg iteration");
        launched += batch;
    }
    if ((rc = halo_exchange(h, x, st)) != FE_OK) return rc;   // ghost part of the solution
    if (info) {
        info[0] = hs[S_ITERS];
        info[1] = hs[S_RR0] > 0.0 ? sqrt(hs[S_RR] / hs[S_RR0]) : 0.0;
        info[2] = sqrt(hs[S_RR0]);
        info[3] = hs[S_DONE_C];
    }
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

size_t fe_dist_pcg_workspace_bytes(int64_t nrows_local) { return fe_pcg_workspace_bytes(nrows_local) + 64; }

int fe_dist_pcg(int64_t nrows_owned, int64_t nrows_local, int idx64, const void* row_ptr, const int32_t* col_idx,
                const double* vals, const double* rhs, const uint8_t* dirmask, double* x, double rtol, double atol,
                int64_t maxit, int32_t check_every, void* work, double* info, void* comm, int32_t n_nbr,
                const int32_t* nbr_rank, const int64_t* send_off, const int32_t* send_idx, double* sendbuf,
                const int64_t* recv_off, const int32_t* nptr, const int32_t* nadj, void* stream) {
    using namespace fe;
    FE_REQUIRE(nrows_owned >= 0 && nrows_local >= nrows_owned && row_ptr && col_idx && vals && rhs && x && work && comm,
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
                                      rhs, dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj, st);
    return dist_pcg_impl<int32_t>(nrows_owned, nrows_local, static_cast<const int32_t*>(row_ptr), col_idx, vals, rhs,
                                  dirmask, x, rtol, atol, maxit, check_every, work, info, h, nptr, nadj, st);
#else
    set_error("built without NCCL");
    return FE_ERR_NCCL;
#endif
}

}  // extern "C"
