// K3 (tiled form) and its one-time plan.
//
// A tile is TE consecutive elements of a spatially compact element order (Morton order of the element
// centroids, computed once per mesh).  Every node is owned by the tile of its first incident element in
// that order, so each CSR row block belongs to exactly one tile.  The numeric kernel runs one CTA per
// tile:
//   phase 0   cp.async the tile's node descriptors and item locators into shared memory
//   phase 1a  one thread per own element: symmetric (upper-triangular, block) element matrix -> smem
//   phase 1b  one thread per halo item (an incident element of an owned node that lives in another
//             tile): only the owned node's row pair -> smem
//   phase 2   one thread per work item = (owned node, neighbour position): the plan lists, for every
//             work item, which (incident element, local node) pairs contribute, in ascending element id;
//             the thread sums them straight out of shared memory and writes the 2x2 CSR block once.
// No atomics, fixed summation order => bit-identical reruns; element matrices never touch HBM.
// Per element ~350 fp64 ops (vs ~1000 for the row-tile kernel) so the kernel is HBM- not fp64-bound.
//
// Plan storage (tile-major, contiguous per tile => every kernel read is a coalesced stream):
//   tile_ptr[4][ntiles+1]   prefix sums: owned nodes / items / halo items / work items per tile
//   tnode[]   int2 per owned node  {nptr[n], deg | cnt << 8 | item_offset << 16}
//   tloc[]    uint16 per item      own: a << 11 | local element;  halo: 0x8000 | a << 11 | halo index
//   thalo[]   int32 per halo item  element << 4 | a
//   twork[]   uint32 per work item src0 | src1 << 8 | local node << 16 | p << 25, src = item << 4 | b,
//             0xFF = none, src0 == 0xFE marks the node's own (diagonal) block: sum over all its items.
#include "fe_element.cuh"

namespace fe {

constexpr int kTileThreads = 256;
constexpr int kTileElems = 128;      // TE: own elements per tile (threads 0..TE-1 in phase 1a)
constexpr int kTileMaxNodes = 512;   // owned nodes per tile (9-bit local node index)
constexpr int kTileMaxItems = 2048;  // incidence items of owned nodes per tile
constexpr int kHaloBytes = 16 * 1024 + 512;
constexpr unsigned kHaloFlag = 0x8000u;
constexpr unsigned kSrcNone = 0xFFu, kSrcSelf = 0xFEu;

template <int N> struct TileLayout {
    static constexpr int kBlocks = N * (N + 1) / 2;            // upper-triangular node pairs incl. diagonal
    static constexpr int kSlabStride = (3 * kBlocks) | 1;      // odd stride: conflict-free 64-bit rows
    static constexpr int kHaloStride = (3 * N) | 1;
    static constexpr int kHaloCap = kHaloBytes / (kHaloStride * 8);
    static constexpr size_t kSmem = (size_t)(kTileElems * kSlabStride + kHaloCap * kHaloStride) * sizeof(double) +
                                    (size_t)kTileMaxNodes * sizeof(int2) + (size_t)kTileMaxItems * sizeof(uint16_t);
    // block index of the pair lo <= hi in the upper triangle (row-major)
    __host__ __device__ static constexpr int tri(int lo, int hi) { return lo * N - lo * (lo - 1) / 2 + (hi - lo); }
};

// ------------------------------------------------------------------------------------------------
// plan kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread15(unsigned v) {
    v &= 0x7fffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// 30-bit Morton key of the element centroid inside the mesh bounding box (15 bits per axis).
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

__global__ void k_invert_perm(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

__device__ __forceinline__ int tile_of(const int32_t* __restrict__ einv, int e) {
    return (einv ? einv[e] : e) / kTileElems;
}

// owner tile of each node (-1: no incident element); per-tile owned-node count
__global__ void k_tile_owner(int64_t nnode, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
                             const int32_t* __restrict__ einv, int32_t* __restrict__ owner,
                             int32_t* __restrict__ tile_ncount) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int k0 = eptr[n], k1 = eptr[n + 1];
    int own = 0x7fffffff;
    for (int k = k0; k < k1; ++k) {
        const int t = tile_of(einv, einc[k] >> 4);
        own = t < own ? t : own;
    }
    if (k0 == k1) {
        owner[n] = -1;
        return;
    }
    owner[n] = own;
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
// fill) and lay out the tile's items / halo items / work items: per-node offsets and per-tile counts.
// noff[3*i + {0,1,2}] = item / halo / work offset of tile-node slot i inside its tile.
__global__ void k_tile_layout(int64_t ntiles, const int32_t* __restrict__ tile_nptr, int32_t* __restrict__ tile_nodes,
                              const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
                              const int32_t* __restrict__ einv, const int32_t* __restrict__ nptr,
                              int32_t* __restrict__ noff, int32_t* __restrict__ tile_icount,
                              int32_t* __restrict__ tile_hcount, int32_t* __restrict__ tile_wcount,
                              int32_t* __restrict__ stats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) {
        tile_icount[t] = tile_hcount[t] = tile_wcount[t] = 0;
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
    int ni = 0, nh = 0, nw = 0, maxcnt = 0;
    for (int i = lo; i < hi; ++i) {
        const int n = tile_nodes[i];
        noff[3 * (int64_t)i] = ni;
        noff[3 * (int64_t)i + 1] = nh;
        noff[3 * (int64_t)i + 2] = nw;
        const int k0 = eptr[n], k1 = eptr[n + 1];
        ni += k1 - k0;
        maxcnt = (k1 - k0) > maxcnt ? (k1 - k0) : maxcnt;
        for (int k = k0; k < k1; ++k) nh += tile_of(einv, einc[k] >> 4) != (int)t;
        nw += nptr[n + 1] - nptr[n];
    }
    tile_icount[t] = (ni + 1) & ~1;  // even: item locators are staged with 4-byte cp.async
    tile_hcount[t] = nh;
    tile_wcount[t] = nw;
    atomicMax(stats + 0, hi - lo);
    atomicMax(stats + 1, ni);
    atomicMax(stats + 2, nh);
    atomicMax(stats + 3, maxcnt);
}

// One thread per tile-node slot: node descriptor, item locators, halo item list, work records.
__global__ void __launch_bounds__(128) k_tile_records(
    int64_t nslots, int nnpe, const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ owner,
    const int32_t* __restrict__ tile_nptr, const int32_t* __restrict__ tile_iptr,
    const int32_t* __restrict__ tile_hptr, const int32_t* __restrict__ tile_wptr, const int32_t* __restrict__ noff,
    const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc, const uint8_t* __restrict__ epos,
    const int32_t* __restrict__ einv, const int32_t* __restrict__ nptr, const int32_t* __restrict__ nadj,
    int2* __restrict__ tnode, uint16_t* __restrict__ tloc, int32_t* __restrict__ thalo, uint32_t* __restrict__ twork,
    int32_t* __restrict__ stats) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const int n = tile_nodes[i];
    const int t = owner[n];
    const int ln = (int)(i - tile_nptr[t]);
    const int ioff = noff[3 * i], hoff = noff[3 * i + 1], woff = noff[3 * i + 2];
    const int k0 = eptr[n], cnt = eptr[n + 1] - k0;
    const int p0 = nptr[n], deg = nptr[n + 1] - p0;
    tnode[i] = make_int2(p0, deg | (cnt << 8) | (ioff << 16));
    if (cnt > 14 || deg > 127 || ioff > 0xFFFF) atomicMax(stats + 4, 1);  // not representable: use fe_assemble

    unsigned char src[kMaxAdj][2];
    for (int p = 0; p < deg; ++p) src[p][0] = src[p][1] = (unsigned char)kSrcNone;
    // position of the node itself inside its own adjacency list
    int self = 0;
    for (int p = 0; p < deg; ++p) self = (nadj[p0 + p] == n) ? p : self;
    int h = hoff;
    for (int j = 0; j < cnt; ++j) {
        const int item = einc[k0 + j];
        const int e = item >> 4, a = item & 15;
        const int pos = einv ? einv[e] : e;
        unsigned loc;
        if (pos / kTileElems == t) {
            loc = (unsigned)(a << 11) | (unsigned)(pos - t * kTileElems);
        } else {
            const int hl = h;  // halo index inside the tile
            loc = kHaloFlag | (unsigned)(a << 11) | (unsigned)hl;
            thalo[tile_hptr[t] + hl] = item;
            if (hl > 0x7FF) atomicMax(stats + 4, 1);
            ++h;
        }
        tloc[tile_iptr[t] + ioff + j] = (uint16_t)loc;
        if (j < 15) {
            for (int b = 0; b < nnpe; ++b) {
                const int p = epos[(int64_t)(k0 + j) * nnpe + b];
                if (p == self) continue;
                const unsigned char s = (unsigned char)((j << 4) | b);
                if (src[p][0] == kSrcNone) src[p][0] = s;
                else if (src[p][1] == kSrcNone) src[p][1] = s;
                else atomicMax(stats + 4, 1);  // an edge shared by more than two elements
            }
        }
    }
    for (int p = 0; p < deg; ++p) {
        const unsigned s0 = (p == self) ? kSrcSelf : src[p][0];
        const unsigned s1 = (p == self) ? kSrcNone : src[p][1];
        twork[tile_wptr[t] + woff + p] = s0 | (s1 << 8) | ((unsigned)ln << 16) | ((unsigned)p << 25);
    }
}

// ------------------------------------------------------------------------------------------------
// numeric kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Full upper-triangular block element matrix: block (lo,hi) = [L, P, Q] at 3*tri(lo,hi); diagonal blocks
// store Q = P.  Same pair arithmetic as element_row_pair (fe_element.cuh).
template <int ET>
__device__ __forceinline__ void element_tri(const double (&x)[Elem<ET>::NNPE], const double (&y)[Elem<ET>::NNPE],
                                            bool weighted, double (&out)[3 * TileLayout<Elem<ET>::NNPE>::kBlocks]) {
    using E = Elem<ET>;
    using T = TileLayout<E::NNPE>;
#pragma unroll
    for (int i = 0; i < 3 * T::kBlocks; ++i) out[i] = 0.0;
#pragma unroll
    for (int g = 0; g < E::NG; ++g) {
        double ax[E::NNPE], ay[E::NNPE], sx[E::NNPE], sy[E::NNPE];
        gauss_point<ET>(g, x, y, weighted, ax, ay, sx, sy);
#pragma unroll
        for (int a = 0; a < E::NNPE; ++a) {
#pragma unroll
            for (int b = a; b < E::NNPE; ++b) {
                const int o = 3 * T::tri(a, b);
                out[o] = fma(sx[a], ax[b], fma(sy[a], ay[b], out[o]));
                out[o + 1] = fma(sy[a], ax[b], out[o + 1]);
                if (b > a) out[o + 2] = fma(sx[a], ay[b], out[o + 2]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < E::NNPE; ++a) out[3 * T::tri(a, a) + 2] = out[3 * T::tri(a, a) + 1];
}

template <int ET>
__global__ void __launch_bounds__(kTileThreads, 2) k_assemble_tiled(
    int64_t nelem, const int32_t* __restrict__ conn, const double2* __restrict__ coords,
    const int32_t* __restrict__ eperm, const int32_t* __restrict__ tile_nptr, const int32_t* __restrict__ tile_iptr,
    const int32_t* __restrict__ tile_hptr, const int32_t* __restrict__ tile_wptr, const int2* __restrict__ tnode,
    const uint16_t* __restrict__ tloc, const int32_t* __restrict__ thalo, const uint32_t* __restrict__ twork,
    double* __restrict__ vals, int weighted) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    using T = TileLayout<N>;
    extern __shared__ double smem[];
    double* slab = smem;
    double* halo = smem + kTileElems * T::kSlabStride;
    int2* s_node = reinterpret_cast<int2*>(halo + T::kHaloCap * T::kHaloStride);
    uint16_t* s_loc = reinterpret_cast<uint16_t*>(s_node + kTileMaxNodes);

    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t e0 = t * kTileElems;

    // ---- phase 0: asynchronous staging of the tile's descriptors -----------------------------------
    const int n0 = __ldg(tile_nptr + t), nn = __ldg(tile_nptr + t + 1) - n0;
    const int i0 = __ldg(tile_iptr + t), ni = __ldg(tile_iptr + t + 1) - i0;  // even, i0 even
    const int w0 = __ldg(tile_wptr + t), nw = __ldg(tile_wptr + t + 1) - w0;
    for (int i = tid; i < nn; i += kTileThreads) cp_async_8(s_node + i, tnode + n0 + i);
    for (int i = tid; 2 * i < ni; i += kTileThreads) cp_async_4(s_loc + 2 * i, tloc + i0 + 2 * i);
    unsigned rec = (tid < nw) ? __ldg(twork + w0 + tid) : 0u;  // first work record, prefetched

    if (tid < kTileElems) {
        // ---- phase 1a: own elements --------------------------------------------------------------
        if (e0 + tid < nelem) {
            const int64_t e = eperm ? (int64_t)__ldg(eperm + e0 + tid) : e0 + tid;
            int tg[N];
            double x[N], y[N], ke[3 * T::kBlocks];
            load_element<ET>(conn, coords, e, tg, x, y);
            element_tri<ET>(x, y, weighted != 0, ke);
            double* dst = slab + tid * T::kSlabStride;
#pragma unroll
            for (int i = 0; i < 3 * T::kBlocks; ++i) dst[i] = ke[i];
        }
    } else {
        // ---- phase 1b: halo items (row pairs of elements owned by other tiles) -------------------
        const int h0 = __ldg(tile_hptr + t);
        const int nh = __ldg(tile_hptr + t + 1) - h0;
        for (int h = tid - kTileElems; h < nh; h += kTileThreads - kTileElems) {
            const int item = __ldg(thalo + h0 + h);
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
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 2: one thread per work item ----------------------------------------------------------
    for (int q = tid; q < nw; q += kTileThreads) {
        if (q != tid) rec = __ldg(twork + w0 + q);
        const unsigned s0 = rec & 0xFFu, s1 = (rec >> 8) & 0xFFu;
        const int ln = (rec >> 16) & 0x1FF, p = rec >> 25;
        const int2 nd = s_node[ln];
        const int deg = nd.y & 0xFF, cnt = (nd.y >> 8) & 0xFF, ioff = (unsigned)nd.y >> 16;
        double L = 0.0, P = 0.0, Q = 0.0;
        // contribution of item j, block (a_item, b): own -> slab (transposed when a > b), halo -> row pair
        auto add = [&](int j, int b, bool diag) {
            const unsigned loc = s_loc[ioff + j];
            const int a = (loc >> 11) & 15;
            if (diag) b = a;
            const double* src;
            bool swap = false;
            if (loc & kHaloFlag) {
                src = halo + (loc & 0x7FFu) * T::kHaloStride + 3 * b;
            } else {
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                swap = a > b;
                src = slab + (loc & 0x7FFu) * T::kSlabStride + 3 * (lo * N - ((lo * (lo - 1)) >> 1) + (hi - lo));
            }
            const double l = src[0], u = src[1], v = src[2];
            L += l;
            P += swap ? v : u;
            Q += swap ? u : v;
        };
        if (s0 == kSrcSelf) {
            for (int j = 0; j < cnt; ++j) add(j, 0, true);
        } else {
            if (s0 != kSrcNone) add(s0 >> 4, s0 & 15, false);
            if (s1 != kSrcNone) add(s1 >> 4, s1 & 15, false);
        }
        double* out = vals + 4 * (int64_t)nd.x;
        __stcs(reinterpret_cast<double2*>(out + 2 * p), make_double2(L, P));
        __stcs(reinterpret_cast<double2*>(out + 2 * deg + 2 * p), make_double2(Q, L));
    }
}

template <int ET>
static int launch_tiled(int64_t nelem, const int32_t* conn, const double* coords, const int32_t* eperm,
                        const int32_t* tile_ptr, const int2* tnode, const uint16_t* tloc, const int32_t* thalo,
                        const uint32_t* twork, double* vals, int weighted, cudaStream_t st) {
    using T = TileLayout<Elem<ET>::NNPE>;
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    const int64_t s = ntiles + 1;
    FE_CUDA_TRY(cudaFuncSetAttribute(k_assemble_tiled<ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::kSmem));
    k_assemble_tiled<ET><<<(unsigned)ntiles, kTileThreads, T::kSmem, st>>>(
        nelem, conn, reinterpret_cast<const double2*>(coords), eperm, tile_ptr, tile_ptr + s, tile_ptr + 2 * s,
        tile_ptr + 3 * s, tnode, tloc, thalo, twork, vals, weighted);
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
    out[2] = kTileMaxNodes;
    out[3] = kTileMaxItems;
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
                  int32_t* owner, int32_t* tile_ncount, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && nelem >= 0, "negative size");
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    FE_REQUIRE(tile_ncount, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(tile_ncount, 0, (size_t)(ntiles + 1) * sizeof(int32_t), st));
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(eptr && einc && owner, "null pointer");
    k_tile_owner<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, eptr, einc, einv, owner, tile_ncount);
    FE_CHECK_LAUNCH("k_tile_owner");
    return FE_OK;
}

int fe_tile_layout(int64_t nnode, int64_t nelem, const int32_t* owner, const int32_t* tile_nptr, const int32_t* eptr,
                   const int32_t* einc, const int32_t* einv, const int32_t* nptr, int32_t* cursor,
                   int32_t* tile_nodes, int32_t* noff, int32_t* tile_counts, int32_t* stats, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && nelem >= 0, "negative size");
    const int64_t ntiles = (nelem + kTileElems - 1) / kTileElems;
    FE_REQUIRE(tile_nptr && cursor && tile_counts && stats, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)(ntiles > 0 ? ntiles : 1) * sizeof(int32_t), st));
    FE_CUDA_TRY(cudaMemsetAsync(stats, 0, 8 * sizeof(int32_t), st));
    if (nnode > 0) {
        FE_REQUIRE(owner && tile_nodes && noff && eptr && einc && nptr, "null pointer");
        k_tile_bucket<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, owner, tile_nptr, cursor, tile_nodes);
        FE_CHECK_LAUNCH("k_tile_bucket");
    }
    const int64_t s = ntiles + 1;
    k_tile_layout<<<grid_for(ntiles + 1, 128), 128, 0, st>>>(ntiles, tile_nptr, tile_nodes, eptr, einc, einv, nptr,
                                                            noff, tile_counts, tile_counts + s, tile_counts + 2 * s,
                                                            stats);
    FE_CHECK_LAUNCH("k_tile_layout");
    return FE_OK;
}

int fe_tile_records(int elem_type, int64_t nslots, const int32_t* tile_nodes, const int32_t* owner,
                    const int32_t* tile_ptr, int64_t ntiles, const int32_t* noff, const int32_t* eptr,
                    const int32_t* einc, const uint8_t* epos, const int32_t* einv, const int32_t* nptr,
                    const int32_t* nadj, void* tnode, uint16_t* tloc, int32_t* thalo, uint32_t* twork, int32_t* stats,
                    void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_tile_records: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nslots >= 0 && ntiles >= 0, "negative size");
    if (nslots == 0) return FE_OK;
    FE_REQUIRE(tile_nodes && owner && tile_ptr && noff && eptr && einc && epos && nptr && nadj && tnode && tloc &&
                   thalo && twork && stats,
               "null pointer");
    const int64_t s = ntiles + 1;
    k_tile_records<<<grid_for(nslots, 128), 128, 0, as_stream(stream)>>>(
        nslots, nnpe, tile_nodes, owner, tile_ptr, tile_ptr + s, tile_ptr + 2 * s, tile_ptr + 3 * s, noff, eptr, einc,
        epos, einv, nptr, nadj, static_cast<int2*>(tnode), tloc, thalo, twork, stats);
    FE_CHECK_LAUNCH("k_tile_records");
    return FE_OK;
}

int fe_assemble_tiled(int elem_type, int op, int64_t nelem, const int32_t* conn, const double* coords,
                      const int32_t* eperm, const int32_t* tile_ptr, const void* tnode, const uint16_t* tloc,
                      const int32_t* thalo, const uint32_t* twork, double* vals, void* stream) {
    using namespace fe;
    FE_REQUIRE(nelem >= 0, "negative size");
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED) {
        set_error("fe_assemble_tiled: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (nelem == 0) return FE_OK;
    FE_REQUIRE(conn && coords && tile_ptr && tnode && tloc && thalo && twork && vals, "null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const int w = op == FE_OP_WEIGHTED;
    const int2* nd = static_cast<const int2*>(tnode);
    switch (elem_type) {
        case FE_QUAD4:
            return launch_tiled<FE_QUAD4>(nelem, conn, coords, eperm, tile_ptr, nd, tloc, thalo, twork, vals, w, st);
        case FE_TRI3:
            return launch_tiled<FE_TRI3>(nelem, conn, coords, eperm, tile_ptr, nd, tloc, thalo, twork, vals, w, st);
        default:
            set_error("fe_assemble_tiled: element type %d has no tiled kernel (use fe_assemble)", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
