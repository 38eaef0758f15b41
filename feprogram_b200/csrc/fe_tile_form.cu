// Tile formation on the device (one-time plan step 1 of fe_assemble_tiled / fe_assemble_ws): groups the elements
// into tiles of <= TE elements of one aligned Morton block WITHOUT a global sort -- a counting sort by block
// (histogram, scan, scatter) followed by a per-block sort in shared memory.  Result (identical to sorting the
// elements by (Morton key, id), cutting every block into runs of TE, and ordering each tile by element id):
//   eperm[nelem]        elements in tile order
//   eloc[nelem]         tile << 8 | row of every element
//   tile_estart[nt+1]   offsets of the tiles in eperm
// Integer atomics only bucket; every bucket is sorted afterwards, so the outputs are deterministic.
#include "fe_tile_config.cuh"

namespace fe {

constexpr int kFormCap = 1024;     // elements of one Morton block the per-block sort handles

// Bounding box of the node coordinates: block partials, then one block over the partials.  out = {xmin, ymin, xmax, ymax}.
constexpr int kBoxBlocks = 592;
__global__ void __launch_bounds__(256) k_bbox(int64_t n, const double2* __restrict__ xy, int nparts,
                                              const double4* __restrict__ parts_in, double4* __restrict__ out) {
    __shared__ double4 sm[8];
    double4 b = make_double4(1.0e300, 1.0e300, -1.0e300, -1.0e300);
    if (parts_in) {
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
            const double4 p = parts_in[i];
            b.x = fmin(b.x, p.x); b.y = fmin(b.y, p.y); b.z = fmax(b.z, p.z); b.w = fmax(b.w, p.w);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const double2 p = __ldg(xy + i);
            b.x = fmin(b.x, p.x); b.y = fmin(b.y, p.y); b.z = fmax(b.z, p.x); b.w = fmax(b.w, p.y);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        b.x = fmin(b.x, __shfl_xor_sync(0xffffffffu, b.x, d));
        b.y = fmin(b.y, __shfl_xor_sync(0xffffffffu, b.y, d));
        b.z = fmax(b.z, __shfl_xor_sync(0xffffffffu, b.z, d));
        b.w = fmax(b.w, __shfl_xor_sync(0xffffffffu, b.w, d));
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            b.x = fmin(b.x, sm[w].x); b.y = fmin(b.y, sm[w].y); b.z = fmax(b.z, sm[w].z); b.w = fmax(b.w, sm[w].w);
        }
        out[blockIdx.x] = b;
    }
}

// Lattice numbering: every Quad4 element is {b, b+1, b+W+1, b+W} for one W (what a row-major structured generator --
// and gmsh's transfinite meshes -- produce).  Then (b mod W, b div W) are the element's lattice coordinates and the
// Morton keys need no node coordinates: the whole tile plan can be built while the coordinates are still uploading.
// Generalised form (the local mesh of a rank of a row-block partition: owned node rows first, ghost rows after them):
// every element is {b, b+1, c+1, c} with b and c in the same column of two node rows of W consecutive ids each, in
// any row order; the element's row coordinate is then the index of its UPPER node row (c div W), which is unique per
// element row.  W is read from the middle element, the form (strict / generalised) from the first one, so every
// thread takes the same decision.
// status[0] |= 1 when some element does not fit the pattern (the caller then uses fe_morton_keys); status[1] = W.
__device__ __forceinline__ unsigned spread15f(unsigned v) {
    v &= 0x7fffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
__global__ void k_lattice_keys(int64_t nelem, const int4* __restrict__ conn, int32_t* __restrict__ keys,
                               int32_t* __restrict__ status) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int4 c0 = __ldg(conn), cm = __ldg(conn + nelem / 2);
    const int W = cm.w - cm.x;
    const bool strict = c0.w - c0.x == W;          // rows in lattice order (a single-GPU structured mesh)
    if (e == 0) status[1] = W;
    if (e >= nelem) return;
    const int4 c = __ldg(conn + e);
    bool ok = W >= 2 && c.y == c.x + 1 && c.z == c.w + 1 && c.x >= 0 && c.w >= 0;
    int ix = 0, iy = 0;
    if (ok) {
        const int rx = c.x / W, rw = c.w / W;
        ix = c.x - rx * W;
        ok = ix == c.w - rw * W && rx != rw && (!strict || rw == rx + 1);
        iy = strict ? rx : rw;
        ok = ok && ix < W - 1 && ix < 32768 && iy < 32768;
    }
    if (!ok) {
        atomicOr(status, 1);
        return;
    }
    keys[e] = (int32_t)(spread15f((unsigned)ix) | (spread15f((unsigned)iy) << 1));
}

// (four elements per thread in the two bucketing kernels: independent atomics in flight)
__global__ void k_block_hist(int64_t nelem, const int32_t* __restrict__ keys, int shift, int32_t* __restrict__ hist) {
    const int64_t e0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll
    for (int q = 0; q < 4; ++q) (void)warp_agg_inc(hist, e0 + q < nelem ? keys[e0 + q] >> shift : -1);
}

// tpb[b] = tiles of block b (and tpb[nb] = 0: slot of the grand total after the scan); status |= 1 on oversize blocks
__global__ void k_block_tiles(int64_t nb, int32_t* __restrict__ hist, int te, int32_t* __restrict__ tpb,
                              int32_t* __restrict__ status) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    if (b == nb) {
        hist[b] = 0;
        tpb[b] = 0;
        return;
    }
    const int c = hist[b];
    tpb[b] = (c + te - 1) / te;
    if (c > kFormCap) atomicMax(status, 1);
}

__global__ void k_block_scatter(int64_t nelem, const int32_t* __restrict__ keys, int shift,
                                const int32_t* __restrict__ ebase, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ tmp) {
    const int64_t e0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    int b[4], base[4], slot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) b[q] = e0 + q < nelem ? keys[e0 + q] >> shift : -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        base[q] = b[q] >= 0 ? ebase[b[q]] : 0;
        slot[q] = warp_agg_inc(cursor, b[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (b[q] >= 0) tmp[base[q] + slot[q]] = (int32_t)(e0 + q);
}

// One warp per non-empty block.  <= TE elements: sort by element id.  More: rank-sort by (key, id), cut into
// runs of TE, then order every run by element id.  Two launches: k_block_sort_small takes the blocks of at most
// min(TE, 256) elements (all of them on regular meshes) with a register bitonic sort and no shared memory, so that
// the kernel runs at full occupancy; k_block_sort takes the others with rank sorts in shared memory.
constexpr int kFormWarps = 4;
__global__ void __launch_bounds__(256) k_block_sort_small(int64_t nb, const int32_t* __restrict__ ebase,
                                                          const int32_t* __restrict__ tbase,
                                                          const int32_t* __restrict__ tmp, int te,
                                                          int32_t* __restrict__ eperm, int32_t* __restrict__ eloc,
                                                          int32_t* __restrict__ tile_estart) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b > nb) return;
    if (b == nb) {
        if (lane == 0) tile_estart[tbase[nb]] = ebase[nb];
        return;
    }
    const int lo = ebase[b], cnt = ebase[b + 1] - lo;
    if (cnt == 0 || cnt > te || cnt > 256) return;
    const int t0 = tbase[b];
    int v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = 32 * r + lane < cnt ? tmp[lo + 32 * r + lane] : 0x7fffffff;
    warp_bitonic_sort<8>(v, lane);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = 32 * r + lane;
        if (i < cnt) {
            eperm[lo + i] = v[r];
            eloc[v[r]] = (t0 << 8) | i;
        }
    }
    if (lane == 0) tile_estart[t0] = lo;
}

__global__ void __launch_bounds__(32 * kFormWarps) k_block_sort(int64_t nb, const int32_t* __restrict__ ebase,
                                                               const int32_t* __restrict__ tbase,
                                                               const int32_t* __restrict__ tmp,
                                                               const int32_t* __restrict__ keys, int te,
                                                               int32_t* __restrict__ eperm, int32_t* __restrict__ eloc,
                                                               int32_t* __restrict__ tile_estart) {
    __shared__ int s_a[kFormWarps][kFormCap];
    __shared__ int s_b[kFormWarps][kFormCap];
    __shared__ int s_k[kFormWarps][kFormCap];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent warps stride over the blocks 32 at a time: every lane looks at one block's size, the warp then takes
    // the oversize ones in turn (on regular meshes there are none and the kernel ends after a couple of coalesced loads).
    for (int64_t b0 = 32 * ((int64_t)blockIdx.x * kFormWarps + w); b0 < nb; b0 += 32 * (int64_t)gridDim.x * kFormWarps) {
      int c = 0;
      if (b0 + lane < nb) c = ebase[b0 + lane + 1] - ebase[b0 + lane];
      unsigned todo = __ballot_sync(0xffffffffu, c > 0 && c <= kFormCap && !(c <= te && c <= 256));   // else: k_block_sort_small's
      while (todo) {
        const int64_t b = b0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int lo = ebase[b], cnt = ebase[b + 1] - lo;
        const int t0 = tbase[b];
        int* a = s_a[w];
        int* o = s_b[w];
        for (int i = lane; i < cnt; i += 32) a[i] = tmp[lo + i];
        __syncwarp();
        if (cnt <= te) {
            for (int i = lane; i < cnt; i += 32) {        // ids are distinct: rank = number of smaller ids
                const int v = a[i];
                int r = 0;
                for (int j = 0; j < cnt; ++j) r += a[j] < v;
                eperm[lo + r] = v;
                eloc[v] = (t0 << 8) | r;
            }
            if (lane == 0) tile_estart[t0] = lo;
            __syncwarp();
            continue;
        }
        int* kk = s_k[w];
        for (int i = lane; i < cnt; i += 32) kk[i] = keys[a[i]];
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {            // order by (key, id)
            const int v = a[i], kv = kk[i];
            int r = 0;
            for (int j = 0; j < cnt; ++j) r += (kk[j] < kv) || (kk[j] == kv && a[j] < v);
            o[r] = v;
        }
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {            // inside each run of TE: order by id
            const int v = o[i];
            const int r0 = (i / te) * te, r1 = r0 + te < cnt ? r0 + te : cnt;
            int r = 0;
            for (int j = r0; j < r1; ++j) r += o[j] < v;
            eperm[lo + r0 + r] = v;
            eloc[v] = ((t0 + i / te) << 8) | r;
        }
        for (int q = lane; q * te < cnt; q += 32) tile_estart[t0 + q] = lo + q * te;
        __syncwarp();                                     // the shared arrays are reused by this warp's next block
      }
    }
}

}  // namespace fe

extern "C" {

// out[0..3] <- {xmin, ymin, xmax, ymax} of coords[nnode][2] (DEVICE doubles, 32-byte aligned); scratch: 592 * 4 doubles.
int fe_bbox(int64_t nnode, const double* coords, double* scratch, double* out, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 1 && coords && scratch && out, "null pointer or empty mesh");
    cudaStream_t st = as_stream(stream);
    int g = (int)((nnode + 255) / 256);
    if (g > kBoxBlocks) g = kBoxBlocks;
    k_bbox<<<g, 256, 0, st>>>(nnode, reinterpret_cast<const double2*>(coords), 0, nullptr, reinterpret_cast<double4*>(scratch));
    FE_CHECK_LAUNCH("k_bbox");
    k_bbox<<<1, 256, 0, st>>>(0, nullptr, g, reinterpret_cast<const double4*>(scratch), reinterpret_cast<double4*>(out));
    FE_CHECK_LAUNCH("k_bbox");
    return FE_OK;
}

// Morton keys from a lattice numbering (Quad4 only); see k_lattice_keys.  status: 2 int32 (zeroed here).
int fe_lattice_keys(int elem_type, int64_t nelem, const int32_t* conn, int32_t* keys, int32_t* status, void* stream) {
    using namespace fe;
    FE_REQUIRE(status, "null pointer");
    cudaStream_t st = as_stream(stream);
    FE_CUDA_TRY(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
    if (elem_type != FE_QUAD4 || nelem <= 0) {      // no lattice form for this element type: say "does not fit"
        FE_CUDA_TRY(cudaMemsetAsync(status, 1, 1, st));
        return FE_OK;
    }
    FE_REQUIRE(conn && keys, "null pointer");
    k_lattice_keys<<<grid_for(nelem, 256), 256, 0, st>>>(nelem, reinterpret_cast<const int4*>(conn), keys, status);
    FE_CHECK_LAUNCH("k_lattice_keys");
    return FE_OK;
}

// Step A (before the host reads ntiles): histogram, tiles per block, both scans, scatter.
//   keys[nelem]; nb = number of Morton blocks (keys >> log2(TE) < nb); hist[nb+1], tpb[nb+1], ebase[nb+1], tbase[nb+1],
//   cursor[nb], tmp[nelem], status[1] int32 scratch / outputs; scan_ws: fe_scan_workspace_bytes(nb + 1).
// Afterwards tbase[nb] = ntiles, status[0] != 0 = a block holds more than 1024 elements (use another tiling).
int fe_tile_form_count(int elem_type, int64_t nelem, int64_t nb, const int32_t* keys, int32_t* hist, int32_t* tpb,
                       int32_t* ebase, int32_t* tbase, int32_t* cursor, int32_t* tmp, int32_t* status, void* scan_ws,
                       void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_tile_form_count: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nelem >= 0 && nb >= 1 && nb < ((int64_t)1 << 30), "bad size");
    FE_REQUIRE(keys && hist && tpb && ebase && tbase && cursor && tmp && status && scan_ws, "null pointer");
    cudaStream_t st = as_stream(stream);
    const int te = tile_elems(nnpe);
    int shift = 0;
    while ((1 << shift) < te) ++shift;
    FE_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)(nb + 1) * sizeof(int32_t), st));
    FE_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)nb * sizeof(int32_t), st));
    FE_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (nelem > 0) {
        k_block_hist<<<grid_for((nelem + 3) / 4, 256), 256, 0, st>>>(nelem, keys, shift, hist);
        FE_CHECK_LAUNCH("k_block_hist");
    }
    k_block_tiles<<<grid_for(nb + 1, 256), 256, 0, st>>>(nb, hist, te, tpb, status);
    FE_CHECK_LAUNCH("k_block_tiles");
    int rc = exclusive_scan_i32(hist, ebase, nb + 1, scan_ws, st);
    if (rc != FE_OK) return rc;
    rc = exclusive_scan_i32(tpb, tbase, nb + 1, scan_ws, st);
    if (rc != FE_OK) return rc;
    if (nelem > 0) {
        k_block_scatter<<<grid_for((nelem + 3) / 4, 256), 256, 0, st>>>(nelem, keys, shift, ebase, cursor, tmp);
        FE_CHECK_LAUNCH("k_block_scatter");
    }
    return FE_OK;
}

// Step B: per-block sort -> eperm, eloc, tile_estart[ntiles + 1].
int fe_tile_form_sort(int elem_type, int64_t nb, const int32_t* ebase, const int32_t* tbase, const int32_t* tmp,
                      const int32_t* keys, int32_t* eperm, int32_t* eloc, int32_t* tile_estart, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_tile_form_sort: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nb >= 1 && ebase && tbase && tmp && keys && eperm && eloc && tile_estart, "null pointer or bad size");
    k_block_sort_small<<<grid_for(nb + 1, 8), 256, 0, as_stream(stream)>>>(nb, ebase, tbase, tmp, tile_elems(nnpe), eperm,
                                                                       eloc, tile_estart);
    FE_CHECK_LAUNCH("k_block_sort_small");
    int64_t grid_large = (nb + 32 * kFormWarps - 1) / (32 * kFormWarps);
    if (grid_large > 4 * 148) grid_large = 4 * 148;
    k_block_sort<<<(unsigned)grid_large, 32 * kFormWarps, 0, as_stream(stream)>>>(
        nb, ebase, tbase, tmp, keys, tile_elems(nnpe), eperm, eloc, tile_estart);
    FE_CHECK_LAUNCH("k_block_sort");
    return FE_OK;
}

}  // extern "C"
