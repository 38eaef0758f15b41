// K3 (tiled form) and its one-time plan.
//
// A tile is TE consecutive elements of a spatially compact element order (Morton order of the element
// centroids, computed once per mesh).  Every node is owned by the tile of its first incident element in
// that order, so each CSR row block belongs to exactly one tile.  Elements of other tiles that touch an
// owned node are the tile's halo elements (deduplicated, <= HE per tile) and are simply recomputed.
// The numeric kernel runs one CTA of TE + HE threads per tile:
//   phase 0   cp.async the tile's node descriptors and item locators into shared memory
//   phase 1   one thread per own / halo element: upper-triangular block element matrix -> smem slab
//   phase 2a  one thread per work item = (owned node, neighbour position != self): the plan gives the
//             (at most two) slab blocks that contribute, in ascending element id; sum, write the 2x2
//             CSR block once (both rows, 16-byte streaming stores)
//   phase 2b  one thread per owned node: its diagonal block = sum over all its incident elements
// No atomics, fixed summation order => bit-identical reruns; element matrices never touch HBM.
//
// Plan storage (tile-major, contiguous per tile => every kernel read is a coalesced stream):
//   tile_ptr[4][ntiles+1]  prefix sums per tile: owned nodes / items (even) / halo elements / work items (x32)
//   tnode[]   int2 per owned node   {nptr[n], deg | self_pos << 7 | cnt << 14 | item_offset << 18}
//   tloc[]    uint16 per item       slab row << 4 | a          (items of a node: ascending element id)
//   thalo[]   int32, HE per tile    halo element ids (ascending)
//   twork[]   uint32 per work item  src0 | src1 << 14 | first_of_node << 28 | is_self << 29,
//             src = (offset of the block in the slab, in doubles) << 1 | transposed;  an absent source points at the
//             slab's zero block (ready-made offsets: the numeric kernel adds them to the slab base, nothing to decode)
//   tchunk[]  uint32 per 32 work items  (local node of the chunk's first item, +1, -1 if that item starts a
//             node) << 7 | p of that item: lanes rebuild (node, p) with one ballot
#include "fe_tile_config.cuh"

namespace fe {

template <int N> struct TileLayout {
    static constexpr int kBlocks = N * (N + 1) / 2;            // upper-triangular node pairs incl. diagonal
#ifndef FE_SLAB_PAD
#define FE_SLAB_PAD 0
#endif
    static constexpr int kSlabStride = ((3 * kBlocks) | 1) + FE_SLAB_PAD;   // odd stride: conflict-free 64-bit rows
    static constexpr int kTE = tile_elems(N);
    static constexpr int kHalo = halo_cap(N);
    static constexpr int kParts = tile_parts(N);
    static constexpr int kRowThreads = kTE + halo_threads(N);  // threads per part: one per slab row
    static constexpr int kThreads = kRowThreads * kParts;
    static constexpr int kRows = kTE + kHalo;                  // rows per tile of tconn / the slab
    static constexpr int kZeroRow = kTE + kHalo;               // slab row of zeros (target of absent sources)
    static constexpr int kMaxNodes = max_nodes(N), kMaxItems = max_items(N), kMaxWork = max_work(N);
    static constexpr int kCtas = N == 9 ? FE_Q9_CTAS : kParts > 1 ? 2 : kTileCtasPerSm;   // persistent CTAs per SM
    static constexpr int kSrcBits = 14;                        // record source field: slab offset << 1 | transposed
    static constexpr unsigned kSrcMask = (1u << kSrcBits) - 1u;
    static constexpr int kSlabDoubles = (((kZeroRow + 1) * kSlabStride) + 1) & ~1;   // keeps what follows 16-byte aligned
    static_assert((kZeroRow + 1) * kSlabStride < (1 << (kSrcBits - 1)), "slab offset does not fit a source field");
    // shared copy of the rows' node coordinates (multi-thread-per-element types only)
    static constexpr size_t kXyBytes = N == 9 ? (size_t)kRows * N * 16 : 0;
    // per row: (j00, j01, j10, j11, scale) of every Gauss point (multi-thread-per-element types only)
    static constexpr int kJacStride = 5 * N;       // Quad9: N Gauss points; odd stride
    static constexpr size_t kJacBytes = N == 9 ? (size_t)kRows * kJacStride * 8 : 0;
    // one set of staged plan records (16-byte aligned size); the kernel double-buffers it
    static constexpr size_t kRecBytes = (size_t)kMaxWork * 4 + (size_t)kMaxNodes * 8 + (size_t)(kMaxWork / 32) * 4 +
                                        (size_t)kMaxItems * 2;
    static constexpr size_t kSmem = (size_t)kSlabDoubles * sizeof(double) + 2 * kRecBytes + kXyBytes + kJacBytes;
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
                              const double2* __restrict__ coords, double x0, double y0, double inv_cx,
                              double inv_cy, int32_t* __restrict__ keys) {
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
    int ix = (int)((cx - x0) * inv_cx), iy = (int)((cy - y0) * inv_cy);
    ix = ix < 0 ? 0 : (ix > 32767 ? 32767 : ix);
    iy = iy < 0 ? 0 : (iy > 32767 ? 32767 : iy);
    keys[e] = (int32_t)(spread15((unsigned)ix) | (spread15((unsigned)iy) << 1));
}

__global__ void k_invert_perm(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

// eloc[e] = tile << 8 | row: the tile an element belongs to and its row inside that tile
__device__ __forceinline__ int tile_of(const int32_t* __restrict__ eloc, int e) {
    return eloc[e] >> 8;
}

// owner tile of each node (-1: no incident element); per-tile owned-node count
__global__ void k_tile_owner(int64_t nnode, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
                             const int32_t* __restrict__ einv, int32_t* __restrict__ owner,
                             int32_t* __restrict__ tile_ncount) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int k0 = eptr[n], k1 = eptr[n + 1];
    int own = 0x7fffffff;
    for (int kb = k0; kb < k1; kb += 4) {         // four incidences at a time: independent loads, two round trips per batch
        int e[4], t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = kb + i < k1 ? einc[kb + i] >> 4 : -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = e[i] >= 0 ? tile_of(einv, e[i]) : 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 4; ++i) own = t[i] < own ? t[i] : own;
    }
    if (k0 == k1) {
        owner[n] = -1;
        return;
    }
    owner[n] = own;
    (void)warp_agg_inc(tile_ncount, own);       // neighbouring nodes mostly share their owner: one atomic per group
}

// Four consecutive nodes per thread: the owner -> (atomic slot, tile start) -> store chains of the four are
// independent, so four of them are in flight per thread (the kernel is bound by that latency, not by bandwidth).
__global__ void k_tile_bucket(int64_t nnode, const int32_t* __restrict__ owner, const int32_t* __restrict__ tile_nptr,
                              int32_t* __restrict__ cursor, int32_t* __restrict__ tile_nodes) {
    const int64_t n0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (n0 >= nnode) return;
    int t[4], base[4], slot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = n0 + j < nnode ? owner[n0 + j] : -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        base[j] = slot[j] = 0;
        if (t[j] >= 0) base[j] = tile_nptr[t[j]];
        slot[j] = warp_agg_inc(cursor, t[j]);        // (every lane calls it: the j loop is uniform)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (t[j] >= 0) tile_nodes[base[j] + slot[j]] = (int32_t)(n0 + j);
}

// One thread per tile: sort the tile's node list ascending (deterministic order after the atomic bucket
// fill); lay out items and work items (noff[2i], noff[2i+1] = item / work offset of tile-node slot i
// inside its tile); collect the sorted, distinct halo elements.
// One warp per tile: (1) sort the tile's owned nodes by id (rank sort in shared memory), (2) per-node item /
// work offsets by warp scans, (3) the halo list = sorted set of the foreign elements incident to owned nodes
// (candidates collected in shared memory, rank-sorted, first occurrences compacted).
constexpr int kLayoutWarps = 4, kLayoutNodes = 256, kLayoutCand = 512;     // (candidates: foreign incidences of a tile, <= 4 x halo capacity)
__global__ void __launch_bounds__(32 * kLayoutWarps) k_tile_layout(
    int64_t ntiles, const int32_t* __restrict__ tile_nptr, int32_t* __restrict__ tile_nodes,
    const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc, const int32_t* __restrict__ einv,
    const int32_t* __restrict__ nptr, int32_t* __restrict__ noff, int32_t* __restrict__ thalo,
    int32_t* __restrict__ tile_icount, int32_t* __restrict__ tile_hcount, int32_t* __restrict__ tile_wcount,
    int32_t* __restrict__ tile_blob16, int32_t* __restrict__ tile_runs, int32_t* __restrict__ stats, int he) {
    __shared__ int s_nodes[kLayoutWarps][kLayoutNodes];
    __shared__ int s_cand[kLayoutWarps][kLayoutCand];
    __shared__ int s_sorted[kLayoutWarps][kLayoutCand];
    __shared__ int s_ncand[kLayoutWarps];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * kLayoutWarps + w;
    if (t > ntiles) return;
    if (t == ntiles) {
        if (lane == 0) tile_icount[t] = tile_hcount[t] = tile_wcount[t] = tile_blob16[t] = tile_runs[t] = 0;
        return;
    }
    const int lo = tile_nptr[t], hi = tile_nptr[t + 1], nn = hi - lo;
    if (nn > kLayoutNodes) {        // beyond every tile capacity: report it, the host discards this plan
        if (lane == 0) {
            tile_icount[t] = tile_hcount[t] = tile_wcount[t] = tile_blob16[t] = tile_runs[t] = 0;
            atomicMax(stats + 0, nn);
        }
        return;
    }
    int* nodes = s_nodes[w];
    int* cand = s_cand[w];
    int* sorted = s_sorted[w];
    if (lane == 0) s_ncand[w] = 0;
    {   // sort the node ids: bitonic network over 256 keys held 8 per lane (key index = 32 * r + lane), padded with
        // INT_MAX -- exchanges at distance >= 32 are register to register, the others one shuffle per key
        constexpr int R = kLayoutNodes / 32;
        int v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = 32 * r + lane < nn ? tile_nodes[lo + 32 * r + lane] : 0x7fffffff;
        warp_bitonic_sort<R>(v, lane);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (32 * r + lane < nn) nodes[32 * r + lane] = v[r];
    }
    __syncwarp();
    int ni = 0, nw = 0, maxcnt = 0, maxdeg = 0, nruns = 0;
    for (int i0 = 0; i0 < nn; i0 += 32) {
        const int i = i0 + lane;
        int ci = 0, cw = 0, n = 0, k0 = 0;
        bool run_start = false;
        if (i < nn) {
            n = nodes[i];
            tile_nodes[lo + i] = n;
            k0 = eptr[n];
            ci = eptr[n + 1] - k0;
            cw = nptr[n + 1] - nptr[n];
            run_start = i == 0 || nodes[i - 1] != n - 1;     // consecutive ids = one contiguous range of CSR rows
        }
        nruns += __popc(__ballot_sync(0xffffffffu, run_start));
        maxdeg = cw > maxdeg ? cw : maxdeg;
        int si = ci, sw = cw;               // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ui = __shfl_up_sync(0xffffffffu, si, d), uw = __shfl_up_sync(0xffffffffu, sw, d);
            if (lane >= d) {
                si += ui;
                sw += uw;
            }
        }
        if (i < nn) {
            noff[2 * (int64_t)(lo + i)] = ni + si - ci;
            noff[2 * (int64_t)(lo + i) + 1] = nw + sw - cw;
            maxcnt = ci > maxcnt ? ci : maxcnt;
            for (int kb = k0; kb < k0 + ci; kb += 4) {      // four incidences at a time: independent load chains
                int e[4], te4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = kb + q < k0 + ci ? einc[kb + q] >> 4 : -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) te4[q] = e[q] >= 0 ? tile_of(einv, e[q]) : (int)t;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (te4[q] == (int)t) continue;
                    const int slot = atomicAdd(&s_ncand[w], 1);
                    if (slot < kLayoutCand) cand[slot] = e[q];
                }
            }
        }
        ni += __shfl_sync(0xffffffffu, si, 31);
        nw += __shfl_sync(0xffffffffu, sw, 31);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int o = __shfl_xor_sync(0xffffffffu, maxcnt, d);
        maxcnt = o > maxcnt ? o : maxcnt;
        const int g = __shfl_xor_sync(0xffffffffu, maxdeg, d);
        maxdeg = g > maxdeg ? g : maxdeg;
    }
    __syncwarp();
    int nc = s_ncand[w];
    bool overflow = nc > kLayoutCand;
    if (overflow) nc = kLayoutCand;
    for (int i = lane; i < nc; i += 32) {       // stable rank sort (duplicates keep their order)
        const int v = cand[i];
        int r = 0;
        for (int j = 0; j < nc; ++j) {
            const int u = cand[j];
            r += (u < v) || (u == v && j < i);
        }
        sorted[r] = v;
    }
    __syncwarp();
    int nh = 0;                                  // compact first occurrences, ascending
    for (int i0 = 0; i0 < nc; i0 += 32) {
        const int i = i0 + lane;
        const bool first = i < nc && (i == 0 || sorted[i] != sorted[i - 1]);
        const unsigned m = __ballot_sync(0xffffffffu, first);
        if (first) {
            const int pos = nh + __popc(m & ((1u << lane) - 1u));
            if (pos < he) thalo[t * he + pos] = sorted[i];
        }
        nh += __popc(m);
    }
    if (nh > he) {
        overflow = true;
        nh = he;
    }
    if (lane == 0) {
        tile_icount[t] = (ni + 1) & ~1;    // even: item locators are staged with 4-byte cp.async
        tile_hcount[t] = nh;
        tile_wcount[t] = (nw + 31) & ~31;  // chunks of 32 work items never straddle tiles
        // warp-specialised kernel (fe_tiles_ws.cu): record-blob size and runs of contiguous CSR rows of this tile
        const int nnp = (nn + 31) & ~31;
        const int blob = (16 + 8 * nnp + 4 * maxdeg * nnp + 2 * ((ni + 1) & ~1) + 15) & ~15;
        tile_blob16[t] = blob >> 4;
        tile_runs[t] = nruns;
        // running maxima: seven same-address atomics per tile would serialise in L2 (a million of them on cfg2), so a
        // tile only issues the ones that can still raise the value it reads (a stale read just costs a spare atomic)
        auto raise = [&](int k, int v) {
            if (v > *reinterpret_cast<volatile int32_t*>(stats + k)) atomicMax(stats + k, v);
        };
        raise(6, blob);
        raise(7, 32 * nw);
        raise(0, nn);
        raise(1, ni);
        raise(2, overflow ? he + 1 : nh);
        raise(3, maxcnt);
        raise(5, nw);
    }
}

// Tile-ordered copy of the connectivity: row r of tile t = own element r (r < own count) or halo element
// r - TE; unused rows are -1.
__global__ void k_tile_conn(int64_t nrows, int nnpe, const int32_t* __restrict__ conn,
                            const int32_t* __restrict__ eperm, const int32_t* __restrict__ tile_estart,
                            const int32_t* __restrict__ thalo, const int32_t* __restrict__ tile_hptr,
                            int32_t* __restrict__ tconn, int te, int he) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int rows = te + he;
    const int64_t t = i / rows;
    const int r = (int)(i - t * rows);
    int64_t e = -1;
    if (r < te) {
        if (r < tile_estart[t + 1] - tile_estart[t]) e = eperm[tile_estart[t] + r];
    } else if (r - te < tile_hptr[t + 1] - tile_hptr[t]) {
        e = thalo[t * he + (r - te)];
    }
    for (int a = 0; a < nnpe; ++a) tconn[i * nnpe + a] = e >= 0 ? conn[e * nnpe + a] : -1;
}

// One thread per tile-node slot: node descriptor, item locators, work records, chunk headers.
__global__ void __launch_bounds__(128) k_tile_records(
    int64_t nslots, int nnpe, const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ owner,
    const int32_t* __restrict__ tile_nptr, const int32_t* __restrict__ tile_iptr,
    const int32_t* __restrict__ tile_hptr, const int32_t* __restrict__ tile_wptr, const int32_t* __restrict__ noff,
    const int32_t* __restrict__ thalo, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
    const uint8_t* __restrict__ epos, const int32_t* __restrict__ einv, const int32_t* __restrict__ nptr,
    const int32_t* __restrict__ nadj, int2* __restrict__ tnode, uint16_t* __restrict__ tloc,
    uint32_t* __restrict__ twork, uint32_t* __restrict__ tchunk, int32_t* __restrict__ stats, int te, int he,
    int ss) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const int n = tile_nodes[i];
    const int t = owner[n];
    const int ln = (int)(i - tile_nptr[t]);
    const int ioff = noff[2 * i], woff = noff[2 * i + 1];
    const int k0 = eptr[n], cnt = eptr[n + 1] - k0;
    const int p0 = nptr[n], deg = nptr[n + 1] - p0;
    const int nh = tile_hptr[t + 1] - tile_hptr[t];
    const int32_t* hl = thalo + (int64_t)t * he;
    // position of n in its own neighbour list = its slot as seen from any incident element (owned nodes have one)
    const int self = cnt > 0 ? (int)epos[(int64_t)k0 * nnpe + (einc[k0] & 15)] : 0;
    (void)nadj;
    if (cnt > 15 || deg > 127 || deg > kMaxAdj || ioff > 0x3FFF) {  // not representable: caller uses fe_assemble
        atomicMax(stats + 4, 1);
        return;
    }
    tnode[i] = make_int2(p0, deg | (self << 7) | (cnt << 14) | (ioff << 18));

    const unsigned kZeroSrc = (unsigned)((te + he) * ss) << 1;
    unsigned short src[kMaxAdj][2];
    for (int p = 0; p < deg; ++p) src[p][0] = src[p][1] = (unsigned short)kZeroSrc;
    for (int j = 0; j < cnt; ++j) {
        const int item = einc[k0 + j];
        const int e = item >> 4, a = item & 15;
        const int pos = einv[e];
        int row;
        if ((pos >> 8) == t) {
            row = pos & 255;
        } else {  // position inside the tile's sorted halo list
            int lo = 0, hi = nh - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (hl[mid] < e) lo = mid + 1;
                else hi = mid;
            }
            row = te + lo;
        }
        tloc[tile_iptr[t] + ioff + j] = (uint16_t)((row << 4) | a);
        for (int b = 0; b < nnpe; ++b) {
            const int p = epos[(int64_t)(k0 + j) * nnpe + b];
            if (p == self) continue;
            const int blo = a < b ? a : b, bhi = a < b ? b : a;
            const int tri = blo * nnpe - blo * (blo - 1) / 2 + (bhi - blo);
            const unsigned short s = (unsigned short)(((row * ss + 3 * tri) << 1) | (a > b ? 1 : 0));
            if (src[p][0] == kZeroSrc) src[p][0] = s;
            else if (src[p][1] == kZeroSrc) src[p][1] = s;
            else atomicMax(stats + 4, 1);  // a node pair shared by more than two elements
        }
    }
    const int64_t wbase = (int64_t)tile_wptr[t] + woff;
    for (int p = 0; p < deg; ++p) {
        twork[wbase + p] = (unsigned)src[p][0] | ((unsigned)src[p][1] << 14) | (p == 0 ? kWorkFirst : 0u) |
                           (p == self ? kWorkSelf : 0u);
        if (((woff + p) & 31) == 0) tchunk[(wbase + p) >> 5] = ((unsigned)(ln + (p == 0 ? 0 : 1)) << 7) | (unsigned)p;
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
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// One of NPARTS threads of an element: the upper-triangular blocks whose row node `a` this part owns (part 0:
// the first and the last local node, part 1: the others -- 5 + 5 blocks for Quad4), written straight to the
// element's slab row.  Same pair arithmetic as element_tri; the geometry is evaluated by every part.
template <int ET, bool MASS, int PART>
__device__ __forceinline__ void element_tri_part(const double (&x)[Elem<ET>::NNPE], const double (&y)[Elem<ET>::NNPE],
                                                 bool weighted, double* __restrict__ dst) {
    using E = Elem<ET>;
    using T = TileLayout<E::NNPE>;
    constexpr int N = E::NNPE;
    double out[3 * T::kBlocks];
#pragma unroll
    for (int i = 0; i < 3 * T::kBlocks; ++i) out[i] = 0.0;
    ElemGeom<ET> G;
    geom_init<ET, MASS>(G, x, y, weighted);
#pragma unroll
    for (int g = 0; g < E::NG; ++g) {
        double ax[N], ay[N], sx[N], sy[N];
        gauss_point<ET, MASS>(g, G, weighted, ax, ay, sx, sy);
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (((a == 0 || a == N - 1) ? 0 : 1) != PART) continue;
#pragma unroll
            for (int b = a; b < N; ++b) {
                const int o = 3 * T::tri(a, b);
                out[o] = fma(sx[a], ax[b], fma(sy[a], ay[b], out[o]));
                out[o + 1] = fma(sy[a], ax[b], out[o + 1]);
                if (b > a) out[o + 2] = fma(sx[a], ay[b], out[o + 2]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
        if (((a == 0 || a == N - 1) ? 0 : 1) != PART) continue;
        out[3 * T::tri(a, a) + 2] = out[3 * T::tri(a, a) + 1];
#pragma unroll
        for (int b = a; b < N; ++b) {
            const int o = 3 * T::tri(a, b);
            dst[o] = out[o];
            dst[o + 1] = out[o + 1];
            dst[o + 2] = out[o + 2];
        }
    }
}

// Quad9: Jacobian J_g = dH_g X and the scale s_g (1/det, times the Gauss weight when weighted; det * weight for
// the mass operator) of Gauss point g -- same expressions as gauss_point<ET> (fe_element.cuh), computed ONCE per
// element (the five threads of an element share the nine Gauss points) instead of by every thread.
template <bool MASS>
__device__ __forceinline__ void q9_jacobian(const double2* __restrict__ xy, int g, bool weighted, double* __restrict__ out) {
    double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const double2 p = xy[b];
        const double hr = c_q9.dH[g][0][b], hs = c_q9.dH[g][1][b];
        j00 = fma(hr, p.x, j00);
        j01 = fma(hr, p.y, j01);
        j10 = fma(hs, p.x, j10);
        j11 = fma(hs, p.y, j11);
    }
    const double det = fma(j00, j11, -__dmul_rn(j01, j10));
    double s = MASS ? __dmul_rn(det, c_q9.w[g]) : 1.0 / det;
    if (!MASS && weighted) s = __dmul_rn(s, c_q9.w[g]);
    out[0] = j00;
    out[1] = j01;
    out[2] = j10;
    out[3] = j11;
    out[4] = s;
}

// Quad9: one of five threads of an element evaluates the block rows A1 (and A2) of the upper triangle --
// {0}, {1,8}, {2,7}, {3,6}, {4,5}: nine blocks = 27 accumulators each -- reading the element's node
// coordinates from shared memory.  Same pair arithmetic (and therefore the same bits) as element_row_pair.
#ifndef FE_Q9_GP_UNROLL
#define FE_Q9_GP_UNROLL 3
#endif
constexpr int kQ9GpUnroll = FE_Q9_GP_UNROLL;      // Gauss points unrolled per trip of the block-row loop
template <int A1, int A2, bool MASS>
__device__ __forceinline__ void q9_block_rows(const double* __restrict__ Jrow, double* __restrict__ dst) {
    constexpr int N = 9;
    using T = TileLayout<N>;
    constexpr int N1 = N - A1, N2 = A2 >= 0 ? N - A2 : 1;
    double acc1[3 * N1], acc2[3 * N2];
#pragma unroll
    for (int i = 0; i < 3 * N1; ++i) acc1[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3 * N2; ++i) acc2[i] = 0.0;
#pragma unroll kQ9GpUnroll
    for (int g = 0; g < 9; ++g) {
        // Jacobian and scale of this Gauss point: evaluated once per element by q9_jacobians (shared memory)
        const double j00 = Jrow[5 * g], j01 = Jrow[5 * g + 1], j10 = Jrow[5 * g + 2], j11 = Jrow[5 * g + 3];
        const double s = Jrow[5 * g + 4];
        double sx1 = 0.0, sy1 = 0.0, sx2 = 0.0, sy2 = 0.0;
#pragma unroll
        for (int b = A1; b < N; ++b) {
            const double hr = c_q9.dH[g][0][b], hs = c_q9.dH[g][1][b];
            const double ax = MASS ? c_q9.H[g][b] : fma(j11, hr, -__dmul_rn(j01, hs));
            const double ay = MASS ? 0.0 : fma(j00, hs, -__dmul_rn(j10, hr));
            if (b == A1) {
                sx1 = __dmul_rn(ax, s);
                sy1 = __dmul_rn(ay, s);
            }
            {
                const int o = 3 * (b - A1);
                acc1[o] = fma(sx1, ax, fma(sy1, ay, acc1[o]));
                acc1[o + 1] = fma(sy1, ax, acc1[o + 1]);
                if (b > A1) acc1[o + 2] = fma(sx1, ay, acc1[o + 2]);
            }
            if (A2 >= 0 && b >= A2) {
                if (b == A2) {
                    sx2 = __dmul_rn(ax, s);
                    sy2 = __dmul_rn(ay, s);
                }
                const int o = 3 * (b - A2);
                acc2[o] = fma(sx2, ax, fma(sy2, ay, acc2[o]));
                acc2[o + 1] = fma(sy2, ax, acc2[o + 1]);
                if (b > A2) acc2[o + 2] = fma(sx2, ay, acc2[o + 2]);
            }
        }
    }
    acc1[2] = acc1[1];   // diagonal block: Q == P
#pragma unroll
    for (int b = A1; b < N; ++b) {
        double* d = dst + 3 * T::tri(A1, b);
        d[0] = acc1[3 * (b - A1)];
        d[1] = acc1[3 * (b - A1) + 1];
        d[2] = acc1[3 * (b - A1) + 2];
    }
    if (A2 >= 0) {
        acc2[2] = acc2[1];
#pragma unroll
        for (int b = A2 >= 0 ? A2 : N; b < N; ++b) {
            double* d = dst + 3 * T::tri(A2 >= 0 ? A2 : 0, b);
            d[0] = acc2[3 * (b - (A2 >= 0 ? A2 : 0))];
            d[1] = acc2[3 * (b - (A2 >= 0 ? A2 : 0)) + 1];
            d[2] = acc2[3 * (b - (A2 >= 0 ? A2 : 0)) + 2];
        }
    }
}

// Persistent kernel: CTA c processes tiles c, c + G, c + 2G, ...  While tile t is computed, the plan
// records of tile t + G stream into the other shared-memory buffer (cp.async) and its connectivity row
// and coordinates are prefetched into registers, so no global-memory latency is exposed per tile.
// PF selects where the prefetches of the following tiles are issued (bit 0: connectivity row, bit 1: tile
// descriptor): 0 = at the top of the iteration (row two tiles ahead; they land during phase 1), 1 = right
// after barrier A of the previous tile (in flight during phase 2).  Chosen per element type / tile regularity
// from measurements (cfg2 Quad4: 0 -> 1.70 ms, 1 -> 1.80 ms; perturbed Quad4 in pack mode: 1.90 / 1.78 ms;
// Tri3, whose phase 1 is too short to hide a load: 3 -> 1.39 ms against 1.75 ms).
// SCALAR = FE_OP_POISSON: one dof per node, value (row n, neighbour position p) at nptr[n] + p = the L entry of the
// node pair's 2 x 2 block (K[0::2, 0::2] of the reference operator).  Same plan, same slab layout; phase 1 evaluates
// and stores only the L entries (Quad4 / Tri3), phase 2 reads only those.
template <int ET, bool MASS, int PF, bool SCALAR = false>
__global__ void __launch_bounds__(TileLayout<Elem<ET>::NNPE>::kThreads, TileLayout<Elem<ET>::NNPE>::kCtas) k_assemble_tiled(
    int64_t ntiles, const int32_t* __restrict__ tconn, const double2* __restrict__ coords,
    const int4* __restrict__ tdesc, const int2* __restrict__ tnode, const uint16_t* __restrict__ tloc,
    const uint32_t* __restrict__ twork, const uint32_t* __restrict__ tchunk, double* __restrict__ vals, int weighted) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    using T = TileLayout<N>;
    constexpr int SS = T::kSlabStride;
    constexpr int kTileThreads = T::kThreads, kZeroRow = T::kZeroRow, kTileElems = T::kTE;
    extern __shared__ __align__(16) double slab[];
    unsigned char* rec_base = reinterpret_cast<unsigned char*>(slab + T::kSlabDoubles);

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const unsigned le_mask = 0xFFFFFFFFu >> (31u - lane);
    if (tid < 3) slab[kZeroRow * SS + tid] = 0.0;

    // staged record views of buffer b
    auto s_work = [&](int b) { return reinterpret_cast<uint32_t*>(rec_base + b * T::kRecBytes); };
    auto s_node = [&](int b) { return reinterpret_cast<int2*>(s_work(b) + T::kMaxWork); };
    auto s_chunk = [&](int b) { return reinterpret_cast<uint32_t*>(s_node(b) + T::kMaxNodes); };
    auto s_loc = [&](int b) { return reinterpret_cast<uint16_t*>(s_chunk(b) + T::kMaxWork / 32); };
    // tdesc[2t] = {n0, nn, i0, ni}, tdesc[2t+1] = {w0, nw, -, -}
    auto stage = [&](int b, const int4 d0, const int4 d1) {
        const int n0 = d0.x, nn = d0.y, i0 = d0.z, ni = d0.w, w0 = d1.x, nw = d1.y;
        for (int i = tid; 4 * i < nw; i += kTileThreads) cp_async_16(s_work(b) + 4 * i, twork + w0 + 4 * i);
        for (int i = tid; 32 * i < nw; i += kTileThreads) cp_async_4(s_chunk(b) + i, tchunk + (w0 >> 5) + i);
        for (int i = tid; i < nn; i += kTileThreads) cp_async_8(s_node(b) + i, tnode + n0 + i);
        for (int i = tid; 2 * i < ni; i += kTileThreads) cp_async_4(s_loc(b) + 2 * i, tloc + i0 + 2 * i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Prefetching loads are volatile asm so that the compiler keeps them where they are written (well
    // ahead of their first use, on the far side of the barriers) instead of sinking them to the use.
    // register path (Quad4 / Tri3): thread -> (slab row, part); one part per element unless FE_Q4_PARTS = 2
    const int part = tid / T::kRowThreads, rowid = tid - part * T::kRowThreads;
    auto load_row = [&](int64_t t, int (&tg)[N]) {
        const int32_t* row = tconn + (t * T::kRows + rowid) * N;
        if constexpr (N == 4) {
            asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(tg[0]), "=r"(tg[1]), "=r"(tg[2]), "=r"(tg[3]) : "l"(row));
        } else {
#pragma unroll
            for (int a = 0; a < N; ++a) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(tg[a]) : "l"(row + a));
        }
    };
    auto load_xy = [&](const int (&tg)[N], double (&x)[N], double (&y)[N]) {
        if (tg[0] < 0) return;
#pragma unroll
        for (int a = 0; a < N; ++a)
            asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(x[a]), "=d"(y[a]) : "l"(coords + tg[a]));
    };

    // ---- prologue: first tile of this CTA --------------------------------------------------------------
    constexpr bool kSingle = N != 9;                  // register path: rows and coordinates prefetched per thread
    constexpr bool kDescAhead = (PF & 2) != 0, kRowBehind = (PF & 1) != 0;
    constexpr int NR = kSingle ? N : 1;               // register prefetch buffers exist only in that case
    int64_t t = blockIdx.x;
    if (t >= ntiles) return;
    auto load_desc = [&](int64_t tt, int4& a, int4& b) {   // pinned like the other prefetches
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(tdesc + 2 * tt));
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(tdesc + 2 * tt + 1));
    };
    int4 d0, d1, e0, e1;
    load_desc(t, d0, d1);
    e0 = d0;
    e1 = d1;
    if (t + gridDim.x < ntiles) load_desc(t + gridDim.x, e0, e1);
    int tg[N], tgn[N];
    double x[N], y[N];
    (void)NR;
    if constexpr (kSingle) {
        load_row(t, tg);
        tgn[0] = -1;
        if (t + gridDim.x < ntiles) load_row(t + gridDim.x, tgn);
    }
    stage(0, d0, d1);
    if constexpr (kSingle) load_xy(tg, x, y);
    int buf = 0;
    double2* s_xy = reinterpret_cast<double2*>(rec_base + 2 * T::kRecBytes);   // Quad9 only
    double* s_jac = reinterpret_cast<double*>(rec_base + 2 * T::kRecBytes + T::kXyBytes);   // Quad9 only
    // Quad9: the rows' node coordinates are gathered into shared memory by cp.async ONE TILE AHEAD (issued behind
    // barrier A, in flight during phase 2; s_xy is read only by the Jacobian step, which is over by then), from node
    // ids each thread loaded one tile earlier still -- so neither of the two dependent DRAM latencies of the gather
    // sits on a tile.  A thread carries two of the kRows * 9 ids and the "row is live" flag of its slab row.
    constexpr int kXyEntries = kSingle ? 0 : T::kRows * N;
    static_assert(kSingle || 2 * kTileThreads >= T::kRows * N, "two coordinate entries per thread cover a tile");
    int nid0 = -1, nid1 = -1, live_n = -1;          // ids of the NEXT tile to gather / first node of this thread's row there
    bool live = false;
    auto load_ids = [&](int64_t tt, int& a, int& b, int& first) {
        a = b = first = -1;
        if (tt >= ntiles) return;
        const int32_t* rows = tconn + tt * T::kRows * N;
        if (tid < kXyEntries) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(a) : "l"(rows + tid));
        if (tid + kTileThreads < kXyEntries) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(b) : "l"(rows + tid + kTileThreads));
        if (rowid < T::kRows) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(first) : "l"(rows + rowid * N));
    };
    auto request_xy = [&](int a, int b) {
        if (a >= 0) cp_async_16(s_xy + tid, coords + a);
        if (b >= 0) cp_async_16(s_xy + tid + kTileThreads, coords + b);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if constexpr (!kSingle) {
        int first;
        load_ids(t, nid0, nid1, first);
        live = first >= 0;
        request_xy(nid0, nid1);
        load_ids(t + gridDim.x, nid0, nid1, live_n);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }

    for (; t < ntiles; t += gridDim.x) {
        const int64_t tn = t + gridDim.x, tn2 = tn + gridDim.x;
        const bool more = tn < ntiles;
        // Prefetch placement (per element type, measured): either at the top of the iteration (descriptor of
        // the next tile, connectivity row two tiles ahead; they land during phase 1) or right after barrier A
        // of the previous tile (everything in flight during phase 2, nothing issued here).
        int tgn2[N];
        tgn2[0] = -1;
        if constexpr (!kDescAhead) {
            e0 = d0;
            e1 = d1;
            if (more) {      // plain loads: the compiler may schedule them (measurably better than pinning here)
                e0 = __ldg(tdesc + 2 * tn);
                e1 = __ldg(tdesc + 2 * tn + 1);
            }
        }
        if constexpr (kSingle && !kRowBehind) {
            if (more && tn2 < ntiles) load_row(tn2, tgn2);
        }
        // ---- phase 1: own and halo elements -> slab --------------------------------------------------
        if constexpr (kSingle) {
            const bool active = tg[0] >= 0;
            if (active) {
                double* dst = slab + rowid * SS;
                if constexpr (T::kParts == 1) {
                    double ke[3 * T::kBlocks];
                    element_tri<ET, MASS, SCALAR>(x, y, weighted != 0, ke);
#pragma unroll
                    for (int i = 0; i < 3 * T::kBlocks; ++i)
                        if (!SCALAR || i % 3 == 0) dst[i] = ke[i];
                } else if (part == 0) {          // warp-uniform: kRowThreads is a multiple of 32
                    element_tri_part<ET, MASS, 0>(x, y, weighted != 0, dst);
                } else {
                    element_tri_part<ET, MASS, 1>(x, y, weighted != 0, dst);
                }
            }
        } else {
            // Quad9: the rows' node coordinates are in s_xy (gathered one tile ahead); five threads per row share the
            // nine Jacobians, then evaluate their block rows
            const int row = rowid;
            if (live) {
                for (int g = part; g < 9; g += T::kParts)
                    q9_jacobian<MASS>(s_xy + row * N, g, weighted != 0, s_jac + row * T::kJacStride + 5 * g);
            }
            __syncthreads();
            if (live) {
                const double* jr = s_jac + row * T::kJacStride;
                double* dst = slab + row * SS;
                switch (part) {          // warp-uniform: kRowThreads is a multiple of 32
                    case 0: q9_block_rows<0, -1, MASS>(jr, dst); break;
                    case 1: q9_block_rows<1, 8, MASS>(jr, dst); break;
                    case 2: q9_block_rows<2, 7, MASS>(jr, dst); break;
                    case 3: q9_block_rows<3, 6, MASS>(jr, dst); break;
                    default: q9_block_rows<4, 5, MASS>(jr, dst); break;
                }
            }
        }
        if constexpr (kSingle && T::kRows > T::kRowThreads) {
            // overflow halo rows (irregular tiles only): second pass of the halo warp(s), not prefetched
            const int nh = d1.z;
            const int row = rowid + (T::kRowThreads - kTileElems);     // rows kRowThreads.. for threads TE..
            if (nh > T::kRowThreads - kTileElems && rowid >= kTileElems && row < kTileElems + nh) {
                const int32_t* rp = tconn + (t * T::kRows + row) * N;
                int tq[N];
                double xq[N], yq[N], ke[3 * T::kBlocks];
#pragma unroll
                for (int a = 0; a < N; ++a) tq[a] = __ldg(rp + a);
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    const double2 pq = __ldg(coords + tq[a]);
                    xq[a] = pq.x;
                    yq[a] = pq.y;
                }
                double* dst = slab + row * SS;
                if constexpr (T::kParts == 1) {
                    element_tri<ET, MASS, SCALAR>(xq, yq, weighted != 0, ke);
#pragma unroll
                    for (int i = 0; i < 3 * T::kBlocks; ++i) dst[i] = ke[i];
                } else if (part == 0) {
                    element_tri_part<ET, MASS, 0>(xq, yq, weighted != 0, dst);
                } else {
                    element_tri_part<ET, MASS, 1>(xq, yq, weighted != 0, dst);
                }
            }
        }
        // records of the next tile start streaming into the other buffer
        if (more) stage(buf ^ 1, e0, e1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this tile's records have landed
        __syncthreads();
        // prefetches, in flight during phase 2: descriptor of the tile after next, connectivity row of the tile
        // after next (straight into the register set the next tile just vacated), coordinates of the next tile
        int4 f0 = e0, f1 = e1;
        if constexpr (!kSingle) {
            if (more) request_xy(nid0, nid1);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            live = live_n >= 0;
            load_ids(tn2, nid0, nid1, live_n);
        }
        if constexpr (kSingle) {
            if (more) {
#pragma unroll
                for (int a = 0; a < N; ++a) tg[a] = tgn[a];
                if constexpr (kRowBehind) {
                    load_xy(tg, x, y);                     // first: these are needed soonest
                    tgn[0] = -1;
                    if (tn2 < ntiles) load_row(tn2, tgn);   // straight into the set the next tile just vacated
                } else {
#pragma unroll
                    for (int a = 0; a < N; ++a) tgn[a] = tgn2[a];
                    load_xy(tg, x, y);
                }
            }
        }
        if constexpr (kDescAhead) {
            if (tn2 < ntiles) load_desc(tn2, f0, f1);
        }
        {
            const int nn = d0.y, nw = d1.y;
            const uint32_t* work = s_work(buf);
            const uint32_t* chunk = s_chunk(buf);
            const int2* node = s_node(buf);
            const uint16_t* locs = s_loc(buf);
            // ---- phase 2a: off-diagonal work items ---------------------------------------------------
            for (int q = tid; q < nw; q += kTileThreads) {
                const unsigned rec = work[q];
                const unsigned ch = chunk[q >> 5];
                const unsigned before = __ballot_sync(0xFFFFFFFFu, (rec & kWorkFirst) != 0) & le_mask;
                const int ln = (int)(ch >> 7) - 1 + __popc(before);
                const int p = before ? (int)lane - (31 - __clz(before)) : (int)(ch & 127u) + (int)lane;
                if (rec == 0u || (rec & kWorkSelf)) continue;   // padding or diagonal block
                const int2 nd = node[ln];
                const int deg = nd.y & 127;
                const unsigned s0 = rec & T::kSrcMask, s1 = (rec >> T::kSrcBits) & T::kSrcMask;
                const double* b0 = slab + (s0 >> 1);
                const double* b1 = slab + (s1 >> 1);                 // the zero block if absent
                const double L = b0[0] + b1[0];
                if constexpr (SCALAR) {
                    (void)deg;
                    __stcs(vals + (int64_t)nd.x + p, L);
                } else {
                    const int f0 = s0 & 1, f1 = s1 & 1;            // transposed block: P and Q trade places
                    const double P = b0[1 + f0] + b1[1 + f1];
                    const double Q = b0[2 - f0] + b1[2 - f1];
                    double* out = vals + 4 * (int64_t)nd.x + 2 * p;
                    __stcs(reinterpret_cast<double2*>(out), make_double2(L, P));
                    __stcs(reinterpret_cast<double2*>(out + 2 * deg), make_double2(Q, L));
                }
            }
            // ---- phase 2b: diagonal blocks, one thread per owned node ---------------------------------
            for (int ln = tid; ln < nn; ln += kTileThreads) {
                const int2 nd = node[ln];
                const int deg = nd.y & 127, self = (nd.y >> 7) & 127, cnt = (nd.y >> 14) & 15;
                const int ioff = (unsigned)nd.y >> 18;
                double L = 0.0, P = 0.0;
                for (int j = 0; j < cnt; ++j) {
                    const unsigned loc = locs[ioff + j];
                    const int a = loc & 15;
                    const double* b = slab + (loc >> 4) * SS + 3 * (a * N - ((a * (a - 1)) >> 1));
                    L += b[0];
                    if constexpr (!SCALAR) P += b[1];
                }
                if constexpr (SCALAR) {
                    (void)deg;
                    __stcs(vals + (int64_t)nd.x + self, L);
                } else {
                    double* out = vals + 4 * (int64_t)nd.x + 2 * self;
                    __stcs(reinterpret_cast<double2*>(out), make_double2(L, P));
                    __stcs(reinterpret_cast<double2*>(out + 2 * deg), make_double2(P, L));
                }
            }
        }
        if constexpr (!kSingle) asm volatile("cp.async.wait_group 0;" ::: "memory");   // next tile's coordinates (and records)
        __syncthreads();   // slab and this record buffer are free again
        d0 = e0;
        d1 = e1;
        if constexpr (kDescAhead) {
            e0 = f0;
            e1 = f1;
        }
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int ET, bool MASS, int PF, bool SCALAR = false>
static int launch_tiled(int64_t ntiles, const int32_t* tconn, const double* coords, const int32_t* tdesc,
                        const int2* tnode, const uint16_t* tloc, const uint32_t* twork, const uint32_t* tchunk,
                        double* vals, int weighted, cudaStream_t st) {
    using T = TileLayout<Elem<ET>::NNPE>;
    int dev = 0, sms = 0;
    FE_CUDA_TRY(cudaGetDevice(&dev));
    FE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t grid = (int64_t)sms * T::kCtas;
    if (grid > ntiles) grid = ntiles;
    FE_CUDA_TRY(cudaFuncSetAttribute(k_assemble_tiled<ET, MASS, PF, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)T::kSmem));
    k_assemble_tiled<ET, MASS, PF, SCALAR><<<(unsigned)grid, T::kThreads, T::kSmem, st>>>(
        ntiles, tconn, reinterpret_cast<const double2*>(coords), reinterpret_cast<const int4*>(tdesc), tnode, tloc,
        twork, tchunk, vals, weighted);
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

void fe_tile_params(int elem_type, int32_t out[6]) {
    using namespace fe;
    const int nn = nnpe_of(elem_type);
    out[0] = tile_elems(nn);
    out[1] = nn ? halo_cap(nn) : 0;   // 0: no tiled kernel
    out[2] = max_nodes(nn);
    out[3] = max_items(nn);
    out[4] = max_work(nn);
    out[5] = tile_elems(nn) + halo_cap(nn);   // rows per tile of the tile-ordered connectivity
}

int fe_morton_keys(int elem_type, int64_t nelem, const int32_t* conn, const double* coords, double xmin, double ymin,
                   double cell_x, double cell_y, int32_t* keys, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_morton_keys: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nelem >= 0, "negative size");
    if (nelem == 0) return FE_OK;
    FE_REQUIRE(conn && coords && keys, "null pointer");
    FE_REQUIRE(cell_x > 0.0 && cell_y > 0.0, "cell size must be positive");
    k_morton_keys<<<grid_for(nelem, 256), 256, 0, as_stream(stream)>>>(
        nelem, nnpe, conn, reinterpret_cast<const double2*>(coords), xmin, ymin, 1.0 / cell_x,
        1.0 / cell_y, keys);
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

int fe_tile_owner(int64_t nnode, int64_t ntiles, const int32_t* eptr, const int32_t* einc, const int32_t* einv,
                  int32_t* owner, int32_t* tile_ncount, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && ntiles >= 0, "negative size");
    FE_REQUIRE(tile_ncount && einv, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(tile_ncount, 0, (size_t)(ntiles + 1) * sizeof(int32_t), st));
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(eptr && einc && owner, "null pointer");
    k_tile_owner<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, eptr, einc, einv, owner, tile_ncount);
    FE_CHECK_LAUNCH("k_tile_owner");
    return FE_OK;
}

int fe_tile_layout(int elem_type, int64_t nnode, int64_t ntiles, const int32_t* owner, const int32_t* tile_nptr, const int32_t* eptr,
                   const int32_t* einc, const int32_t* einv, const int32_t* nptr, int32_t* cursor,
                   int32_t* tile_nodes, int32_t* noff, int32_t* thalo, int32_t* tile_counts, int32_t* stats,
                   void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0 && ntiles >= 0, "negative size");
    FE_REQUIRE(tile_nptr && cursor && tile_counts && stats && thalo && einv, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)(ntiles > 0 ? ntiles : 1) * sizeof(int32_t), st));
    FE_CUDA_TRY(cudaMemsetAsync(stats, 0, 8 * sizeof(int32_t), st));
    if (nnode > 0) {
        FE_REQUIRE(owner && tile_nodes && noff && eptr && einc && nptr, "null pointer");
        k_tile_bucket<<<grid_for((nnode + 3) / 4, 256), 256, 0, st>>>(nnode, owner, tile_nptr, cursor, tile_nodes);
        FE_CHECK_LAUNCH("k_tile_bucket");
    }
    const int64_t s = ntiles + 1;
    k_tile_layout<<<grid_for(ntiles + 1, kLayoutWarps), 32 * kLayoutWarps, 0, st>>>(ntiles, tile_nptr, tile_nodes, eptr, einc, einv, nptr, noff,
                                                          thalo, tile_counts, tile_counts + s, tile_counts + 2 * s,
                                                          tile_counts + 3 * s, tile_counts + 4 * s, stats,
                                                          halo_cap(nnpe_of(elem_type)));
    FE_CHECK_LAUNCH("k_tile_layout");
    return FE_OK;
}

int fe_tile_records(int elem_type, int64_t nslots, const int32_t* tile_nodes, const int32_t* owner,
                    const int32_t* tile_ptr, int64_t ntiles, const int32_t* noff, const int32_t* thalo,
                    const int32_t* eptr, const int32_t* einc, const uint8_t* epos, const int32_t* einv,
                    const int32_t* nptr, const int32_t* nadj, void* tnode, uint16_t* tloc, uint32_t* twork,
                    uint32_t* tchunk, int32_t* stats, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_tile_records: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nslots >= 0 && ntiles >= 0, "negative size");
    if (nslots == 0) return FE_OK;
    FE_REQUIRE(tile_nodes && owner && tile_ptr && noff && thalo && eptr && einc && epos && nptr && nadj && tnode &&
                   tloc && twork && tchunk && stats,
               "null pointer");
    const int64_t s = ntiles + 1;
    k_tile_records<<<grid_for(nslots, 128), 128, 0, as_stream(stream)>>>(
        nslots, nnpe, tile_nodes, owner, tile_ptr, tile_ptr + s, tile_ptr + 2 * s, tile_ptr + 3 * s, noff, thalo, eptr,
        einc, epos, einv, nptr, nadj, static_cast<int2*>(tnode), tloc, twork, tchunk, stats, tile_elems(nnpe),
        halo_cap(nnpe), nnpe == 9 ? TileLayout<9>::kSlabStride : nnpe == 4 ? TileLayout<4>::kSlabStride : TileLayout<3>::kSlabStride);
    FE_CHECK_LAUNCH("k_tile_records");
    return FE_OK;
}

int fe_tile_conn(int elem_type, int64_t ntiles, const int32_t* conn, const int32_t* eperm, const int32_t* tile_estart,
                 const int32_t* thalo, const int32_t* tile_ptr, int32_t* tconn, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_tile_conn: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(ntiles >= 0, "negative size");
    if (ntiles == 0) return FE_OK;
    FE_REQUIRE(conn && eperm && tile_estart && thalo && tile_ptr && tconn, "null pointer");
    const int he = halo_cap(nnpe);
    const int te = tile_elems(nnpe);
    const int64_t nrows = ntiles * (te + he);
    k_tile_conn<<<grid_for(nrows, 256), 256, 0, as_stream(stream)>>>(nrows, nnpe, conn, eperm, tile_estart, thalo,
                                                                    tile_ptr + 2 * (ntiles + 1), tconn, te, he);
    FE_CHECK_LAUNCH("k_tile_conn");
    return FE_OK;
}

int fe_assemble_tiled(int elem_type, int op, int64_t ntiles, const int32_t* tconn, const double* coords,
                      const int32_t* tdesc, const void* tnode, const uint16_t* tloc, const uint32_t* twork,
                      const uint32_t* tchunk, double* vals, void* stream) {
    using namespace fe;
    FE_REQUIRE(ntiles >= 0, "negative size");
    const bool irregular = (op & FE_HINT_IRREGULAR_TILES) != 0;   // scheduling hint only: same results
    op &= ~FE_HINT_IRREGULAR_TILES;
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED && op != FE_OP_MASS && op != FE_OP_POISSON) {
        set_error("fe_assemble_tiled: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (ntiles == 0) return FE_OK;
    FE_REQUIRE(tconn && coords && tdesc && tnode && tloc && twork && tchunk && vals, "null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const int w = op == FE_OP_WEIGHTED;
    const int2* nd = static_cast<const int2*>(tnode);
#define FE_TILED_ARGS ntiles, tconn, coords, tdesc, nd, tloc, twork, tchunk, vals
    if (op == FE_OP_POISSON) {      // one dof per node: vals[nptr[n] + p]
        switch (elem_type) {
            case FE_QUAD4: return launch_tiled<FE_QUAD4, false, 0, true>(FE_TILED_ARGS, 0, st);
            case FE_TRI3: return launch_tiled<FE_TRI3, false, 3, true>(FE_TILED_ARGS, 0, st);
            case FE_QUAD9: return launch_tiled<FE_QUAD9, false, 2, true>(FE_TILED_ARGS, 0, st);
            default: break;
        }
    }
    switch (elem_type) {
        case FE_QUAD4:
            if (op == FE_OP_MASS) return launch_tiled<FE_QUAD4, true, 0>(FE_TILED_ARGS, 1, st);
            return irregular ? launch_tiled<FE_QUAD4, false, 1>(FE_TILED_ARGS, w, st)
                             : launch_tiled<FE_QUAD4, false, 0>(FE_TILED_ARGS, w, st);
        case FE_TRI3:
            return op == FE_OP_MASS ? launch_tiled<FE_TRI3, true, 3>(FE_TILED_ARGS, 1, st)
                                    : launch_tiled<FE_TRI3, false, 3>(FE_TILED_ARGS, w, st);
        case FE_QUAD9:
            return op == FE_OP_MASS ? launch_tiled<FE_QUAD9, true, 2>(FE_TILED_ARGS, 1, st)
                                    : launch_tiled<FE_QUAD9, false, 2>(FE_TILED_ARGS, w, st);
#undef FE_TILED_ARGS
        default:
            set_error("fe_assemble_tiled: element type %d has no tiled kernel (use fe_assemble)", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
