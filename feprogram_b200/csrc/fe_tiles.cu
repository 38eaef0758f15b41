// K3 (tiled form) and its one-time plan.
//
// A tile is TE consecutive elements of a spatially compact element order (Morton order of the element
// centroids, computed once per mesh).  Every node is owned by the tile of its first incident element in
// that order, so each CSR row block belongs to exactly one tile.  The numeric kernel runs one CTA per
// tile:
//   phase 1a  one thread per own element: symmetric-compressed element matrix -> shared memory
//   phase 1b  one thread per halo item (an incident element of an owned node that lives in another
//             tile): only the owned node's row pair -> shared memory
//   phase 2   one thread per (owned node, neighbour position): sums the node's contributions in
//             ascending element id straight out of shared memory and writes the 2x2 CSR block once.
// No atomics, fixed summation order => bit-identical reruns; element matrices never touch HBM.
// Per element ~350 fp64 ops (vs ~1000 for the row-tile kernel) so the kernel is HBM- not fp64-bound.
#include "fe_element.cuh"

namespace fe {

constexpr int kTileThreads = 256;
constexpr int kTileElems = 128;                // TE: own elements per tile (threads 0..TE-1 in phase 1a)
constexpr int kHaloBytes = 16 * 1024 + 512;    // shared-memory budget of the halo row pairs
constexpr unsigned kHaloFlag = 0x8000u;

template <int N> struct TileLayout {
    static constexpr int kSlabStride = SymLayout<N>::kSize | 1;  // odd stride: conflict-free 64-bit rows
    static constexpr int kHaloStride = (3 * N) | 1;
    static constexpr int kHaloCap = kHaloBytes / (kHaloStride * 8);
    static constexpr size_t kSmem = (size_t)(kTileElems * kSlabStride + kHaloCap * kHaloStride) * sizeof(double);
};

// ------------------------------------------------------------------------------------------------
// plan kernels
// ------------------------------------------------------------------------------------------------
// 30-bit Morton key of the element centroid inside the mesh bounding box (15 bits per axis).
__device__ __forceinline__ unsigned spread15(unsigned v) {
    v &= 0x7fffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void k_morton_keys(int64_t nelem, int nnpe, const int32_t* __restrict__ conn,
                              const double2* __restrict__ coords, double x0, double y0, double sx, double sy,
                              int32_t* __restrict__ keys) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nelem) return;
    double cx = 0.0, cy = 0.0;
    for (int a = 0; a < nnpe; ++a) {
        const double2 p = coords[conn[e * nnpe + a]];
        cx += p.x;
        cy += p.y;
    }
    cx /= nnpe;
    cy /= nnpe;
    int ix = (int)((cx - x0) * sx), iy = (int)((cy - y0) * sy);
    ix = ix < 0 ? 0 : (ix > 32767 ? 32767 : ix);
    iy = iy < 0 ? 0 : (iy > 32767 ? 32767 : iy);
    keys[e] = (int32_t)(spread15((unsigned)ix) | (spread15((unsigned)iy) << 1));
}

// einv[eperm[i]] = i
__global__ void k_invert_perm(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

__device__ __forceinline__ int tile_of(const int32_t* __restrict__ einv, int e, int te) {
    return (einv ? einv[e] : e) / te;
}

// owner tile of each node (-1: no incident element), owned-node count per tile, halo items per node
__global__ void k_tile_owner(int64_t nnode, int te, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
                             const int32_t* __restrict__ einv, int32_t* __restrict__ owner,
                             int32_t* __restrict__ tile_ncount, int32_t* __restrict__ node_hcount) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int k0 = eptr[n], k1 = eptr[n + 1];
    int own = 0x7fffffff;
    for (int k = k0; k < k1; ++k) {
        const int t = tile_of(einv, einc[k] >> 4, te);
        own = t < own ? t : own;
    }
    if (k0 == k1) {
        owner[n] = -1;
        node_hcount[n] = 0;
        return;
    }
    int h = 0;
    for (int k = k0; k < k1; ++k) h += tile_of(einv, einc[k] >> 4, te) != own;
    owner[n] = own;
    node_hcount[n] = h;
    atomicAdd(tile_ncount + own, 1);
}

__global__ void k_tile_bucket(int64_t nnode, const int32_t* __restrict__ owner, const int32_t* __restrict__ tile_nptr,
                              int32_t* __restrict__ cursor, int32_t* __restrict__ tile_nodes) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int t = owner[n];
    if (t < 0) return;
    tile_nodes[tile_nptr[t] + atomicAdd(cursor + t, 1)] = (int32_t)n;
}

// One thread per tile: sort the tile's node list ascending (deterministic order after the atomic bucket
// fill), then assign each node the base index of its halo items inside the tile.
__global__ void k_tile_finish(int64_t ntiles, const int32_t* __restrict__ tile_nptr, int32_t* __restrict__ tile_nodes,
                              const int32_t* __restrict__ node_hcount, int32_t* __restrict__ node_hoff,
                              int32_t* __restrict__ tile_hcount, int32_t* __restrict__ stats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) {
        tile_hcount[t] = 0;
        return;
    }
    const int lo = tile_nptr[t], hi = tile_nptr[t + 1];
    for (int i = lo + 1; i < hi; ++i) {
        const int v = tile_nodes[i];
        int j = i;
        while (j > lo && tile_nodes[j - 1] > v) {
            tile_nodes[j] = tile_nodes[j - 1];
            --j;
        }
        tile_nodes[j] = v;
    }
    int h = 0;
    for (int i = lo; i < hi; ++i) {
        const int n = tile_nodes[i];
        node_hoff[n] = h;
        h += node_hcount[n];
    }
    tile_hcount[t] = h;
    atomicMax(stats + 0, h);
    atomicMax(stats + 1, hi - lo);
}

// per incidence entry: 16-bit locator.  own element: (a << 11) | local element index;
// halo: 0x8000 | halo index inside the tile, and hitems[tile_hptr[t] + index] = incidence entry.
__global__ void k_tile_items(int64_t nnode, int te, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
                             const int32_t* __restrict__ einv, const int32_t* __restrict__ owner,
                             const int32_t* __restrict__ node_hoff, const int32_t* __restrict__ tile_hptr,
                             uint16_t* __restrict__ iloc, int32_t* __restrict__ hitems) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int own = owner[n];
    if (own < 0) return;
    int h = node_hoff[n];
    const int hbase = tile_hptr[own];
    for (int k = eptr[n]; k < eptr[n + 1]; ++k) {
        const int item = einc[k];
        const int e = item >> 4, a = item & 15;
        const int pos = einv ? einv[e] : e;
        if (pos / te == own) {
            iloc[k] = (uint16_t)((a << 11) | (pos - own * te));
        } else {
            iloc[k] = (uint16_t)(kHaloFlag | (unsigned)h);
            hitems[hbase + h] = k;
            ++h;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(kTileThreads, 2) k_assemble_tiled(
    int64_t nelem, const int32_t* __restrict__ conn, const double2* __restrict__ coords,
    const int32_t* __restrict__ eperm, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
    const uint8_t* __restrict__ epos, const int32_t* __restrict__ nptr, const int32_t* __restrict__ tile_nptr,
    const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ tile_hptr, const int32_t* __restrict__ hitems,
    const uint16_t* __restrict__ iloc, double* __restrict__ vals, int weighted, int lpn_log2) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    using S = SymLayout<N>;
    using T = TileLayout<N>;
    extern __shared__ double smem[];
    double* slab = smem;
    double* halo = smem + kTileElems * T::kSlabStride;

    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t e0 = t * kTileElems;

    if (tid < kTileElems) {
        // ---- phase 1a: own elements --------------------------------------------------------------
        if (e0 + tid < nelem) {
            const int64_t e = eperm ? (int64_t)__ldg(eperm + e0 + tid) : e0 + tid;
            int tg[N];
            double x[N], y[N], ke[S::kSize];
            load_element<ET>(conn, coords, e, tg, x, y);
            element_sym<ET>(x, y, weighted != 0, ke);
            double* dst = slab + tid * T::kSlabStride;
#pragma unroll
            for (int i = 0; i < S::kSize; ++i) dst[i] = ke[i];
        }
    } else {
        // ---- phase 1b: halo items (row pairs of elements owned by other tiles) -------------------
        const int h0 = __ldg(tile_hptr + t);
        const int nh = __ldg(tile_hptr + t + 1) - h0;
        for (int h = tid - kTileElems; h < nh; h += kTileThreads - kTileElems) {
            const int k = __ldg(hitems + h0 + h);
            const int item = __ldg(einc + k);
            int tg[N];
            double x[N], y[N], L[N], P[N], Q[N];
            load_element<ET>(conn, coords, item >> 4, tg, x, y);
            element_row_pair<ET>(x, y, item & 15, weighted != 0, L, P, Q);
            double* dst = halo + h * T::kHaloStride;
#pragma unroll
            for (int b = 0; b < N; ++b) {
                dst[3 * b] = L[b];
                dst[3 * b + 1] = P[b];
                dst[3 * b + 2] = Q[b];
            }
        }
    }
    __syncthreads();

    // ---- phase 2: one lane per (owned node, neighbour position) ------------------------------------
    const int n0 = __ldg(tile_nptr + t);
    const int nn = __ldg(tile_nptr + t + 1) - n0;
    const int lpn = 1 << lpn_log2;
    const int groups = kTileThreads >> lpn_log2;
    const int lane = tid & (lpn - 1);
    for (int ln = tid >> lpn_log2; ln < nn; ln += groups) {
        const int n = __ldg(tile_nodes + n0 + ln);
        const int p0 = __ldg(nptr + n);
        const int deg = __ldg(nptr + n + 1) - p0;
        const int k0 = __ldg(eptr + n), k1 = __ldg(eptr + n + 1);
        double* out = vals + 4 * (int64_t)p0;
        for (int p = lane; p < deg; p += lpn) {
            double L = 0.0, P = 0.0, Q = 0.0;
            for (int k = k0; k < k1; ++k) {
                int b = -1;
                if constexpr (N == 4) {
                    const unsigned w = __ldg(reinterpret_cast<const unsigned*>(epos) + k);
                    const unsigned m = __vcmpeq4(w, (unsigned)p * 0x01010101u);
                    if (m) b = (__ffs(m) - 1) >> 3;
                } else {
                    const uint8_t* ps = epos + (int64_t)k * N;
#pragma unroll
                    for (int bb = 0; bb < N; ++bb) b = (__ldg(ps + bb) == p) ? bb : b;
                }
                if (b < 0) continue;
                const unsigned code = __ldg(iloc + k);
                double l, pp, q;
                if (code & kHaloFlag) {
                    const double* src = halo + (code & 0x7fffu) * T::kHaloStride + 3 * b;
                    l = src[0];
                    pp = src[1];
                    q = src[2];
                } else {
                    sym_block<N>(slab + (code & 0x7ffu) * T::kSlabStride, (int)(code >> 11), b, l, pp, q);
                }
                L += l;
                P += pp;
                Q += q;
            }
            __stcs(reinterpret_cast<double2*>(out + 2 * p), make_double2(L, P));
            __stcs(reinterpret_cast<double2*>(out + 2 * deg + 2 * p), make_double2(Q, L));
        }
    }
}

template <int ET>
static int launch_tiled(int64_t nelem, const int32_t* conn, const double* coords, const int32_t* eperm,
                        const int32_t* eptr, const int32_t* einc, const uint8_t* epos, const int32_t* nptr,
                        const int32_t* tile_nptr, const int32_t* tile_nodes, const int32_t* tile_hptr,
                        const int32_t* hitems, const uint16_t* iloc, int32_t max_deg, double* vals, int weighted,
                        cudaStream_t st) {
    using T = TileLayout<Elem<ET>::NNPE>;
    int lpn_log2 = 0;
    while ((1 << lpn_log2) < max_deg && lpn_log2 < 5) ++lpn_log2;
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    FE_CUDA_TRY(cudaFuncSetAttribute(k_assemble_tiled<ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::kSmem));
    k_assemble_tiled<ET><<<(unsigned)ntiles, kTileThreads, T::kSmem, st>>>(
        nelem, conn, reinterpret_cast<const double2*>(coords), eperm, eptr, einc, epos, nptr, tile_nptr, tile_nodes,
        tile_hptr, hitems, iloc, vals, weighted, lpn_log2);
    FE_CHECK_LAUNCH("k_assemble_tiled");
    return FE_OK;
}

}  // namespace fe

extern "C" {

int fe_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scan_ws, void* stream) {
    using namespace fe;
    FE_REQUIRE(n >= 0, "negative size");
    if (n == 0) return FE_OK;
    FE_REQUIRE(in && out && scan_ws, "null pointer");
    return exclusive_scan_i32(in, out, n, scan_ws, as_stream(stream));
}

void fe_tile_params(int elem_type, int32_t out[4]) {
    using namespace fe;
    out[0] = kTileElems;
    out[1] = elem_type == FE_QUAD4 ? TileLayout<4>::kHaloCap
             : elem_type == FE_TRI3 ? TileLayout<3>::kHaloCap : 0;   // 0: tiled kernel not available
    out[2] = 0x7ff + 1;  // local element index range of the 16-bit locator
    out[3] = kTileThreads;
}

int fe_morton_keys(int elem_type, int64_t nelem, const int32_t* conn, const double* coords, double xmin, double ymin,
                   double xmax, double ymax, int32_t* keys, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_morton_keys: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nelem >= 0, "negative size");
    if (nelem == 0) return FE_OK;
    FE_REQUIRE(conn && coords && keys, "null pointer");
    const double sx = xmax > xmin ? 32768.0 / (xmax - xmin) : 0.0;
    const double sy = ymax > ymin ? 32768.0 / (ymax - ymin) : 0.0;
    k_morton_keys<<<grid_for(nelem, 256), 256, 0, as_stream(stream)>>>(
        nelem, nnpe, conn, reinterpret_cast<const double2*>(coords), xmin, ymin, sx, sy, keys);
    FE_CHECK_LAUNCH("k_morton_keys");
    return FE_OK;
}

int fe_invert_permutation(int64_t n, const int32_t* perm, int32_t* inv, void* stream) {
    using namespace fe;
    FE_REQUIRE(n >= 0, "negative size");
    if (n == 0) return FE_OK;
    FE_REQUIRE(perm && inv, "null pointer");
    k_invert_perm<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, perm, inv);
    FE_CHECK_LAUNCH("k_invert_perm");
    return FE_OK;
}

int fe_tile_owner(int64_t nnode, int64_t nelem, const int32_t* eptr, const int32_t* einc, const int32_t* einv,
                  int32_t* owner, int32_t* tile_ncount, int32_t* node_hcount, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && nelem >= 0, "negative size");
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    FE_REQUIRE(tile_ncount, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(tile_ncount, 0, (size_t)(ntiles + 1) * sizeof(int32_t), st));
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(eptr && einc && owner && node_hcount, "null pointer");
    k_tile_owner<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, kTileElems, eptr, einc, einv, owner, tile_ncount,
                                                      node_hcount);
    FE_CHECK_LAUNCH("k_tile_owner");
    return FE_OK;
}

int fe_tile_nodes(int64_t nnode, int64_t nelem, const int32_t* owner, const int32_t* tile_nptr,
                  const int32_t* node_hcount, int32_t* cursor, int32_t* tile_nodes, int32_t* node_hoff,
                  int32_t* tile_hcount, int32_t* stats, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && nelem >= 0, "negative size");
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    FE_REQUIRE(tile_nptr && cursor && tile_hcount && stats, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)(ntiles > 0 ? ntiles : 1) * sizeof(int32_t), st));
    FE_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * sizeof(int32_t), st));
    if (nnode > 0) {
        FE_REQUIRE(owner && tile_nodes && node_hcount && node_hoff, "null pointer");
        k_tile_bucket<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, owner, tile_nptr, cursor, tile_nodes);
        FE_CHECK_LAUNCH("k_tile_bucket");
    }
    k_tile_finish<<<grid_for(ntiles + 1, 128), 128, 0, st>>>(ntiles, tile_nptr, tile_nodes, node_hcount, node_hoff,
                                                            tile_hcount, stats);
    FE_CHECK_LAUNCH("k_tile_finish");
    return FE_OK;
}

int fe_tile_items(int64_t nnode, const int32_t* eptr, const int32_t* einc, const int32_t* einv, const int32_t* owner,
                  const int32_t* node_hoff, const int32_t* tile_hptr, uint16_t* iloc, int32_t* hitems, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0, "negative size");
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(eptr && einc && owner && node_hoff && tile_hptr && iloc && hitems, "null pointer");
    k_tile_items<<<grid_for(nnode, 256), 256, 0, as_stream(stream)>>>(nnode, kTileElems, eptr, einc, einv, owner,
                                                                     node_hoff, tile_hptr, iloc, hitems);
    FE_CHECK_LAUNCH("k_tile_items");
    return FE_OK;
}

int fe_assemble_tiled(int elem_type, int op, int64_t nelem, int64_t nnode, const int32_t* conn, const double* coords,
                      const int32_t* eperm, const int32_t* eptr, const int32_t* einc, const uint8_t* epos,
                      const int32_t* nptr, const int32_t* tile_nptr, const int32_t* tile_nodes,
                      const int32_t* tile_hptr, const int32_t* hitems, const uint16_t* iloc, int32_t max_deg,
                      double* vals, void* stream) {
    using namespace fe;
    FE_REQUIRE(nelem >= 0 && nnode >= 0, "negative size");
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED) {
        set_error("fe_assemble_tiled: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (nelem == 0 || nnode == 0) return FE_OK;
    FE_REQUIRE(conn && coords && eptr && einc && epos && nptr && tile_nptr && tile_nodes && tile_hptr && hitems &&
                   iloc && vals,
               "null pointer");
    FE_REQUIRE(max_deg > 0, "max_deg must be positive");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const int w = op == FE_OP_WEIGHTED;
    switch (elem_type) {
        case FE_QUAD4:
            return launch_tiled<FE_QUAD4>(nelem, conn, coords, eperm, eptr, einc, epos, nptr, tile_nptr, tile_nodes,
                                          tile_hptr, hitems, iloc, max_deg, vals, w, st);
        case FE_TRI3:
            return launch_tiled<FE_TRI3>(nelem, conn, coords, eperm, eptr, einc, epos, nptr, tile_nptr, tile_nodes,
                                         tile_hptr, hitems, iloc, max_deg, vals, w, st);
        default:
            set_error("fe_assemble_tiled: element type %d has no tiled kernel (use fe_assemble)", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
