// Tile configuration shared by the tile-plan kernels and the two numeric tiled kernels (fe_tiles.cu: barrier-phased
// kernel, used for Quad9; fe_tiles_ws.cu: warp-specialised kernel, used for Quad4 / Tri3).
#pragma once

#include "fe_element.cuh"

namespace fe {

#ifndef FE_TILE_ELEMS
#define FE_TILE_ELEMS 128
#define FE_HALO_ELEMS 32
#define FE_TILE_CTAS 3
#endif
#ifndef FE_Q4_PARTS
#define FE_Q4_PARTS 1       // experiment: 2 = two threads per Quad4 element (five node-pair blocks each), 2 CTAs per SM
#endif
#ifndef FE_Q9_CTAS
#define FE_Q9_CTAS 2
#endif
// ---- per-element-type tile configuration ------------------------------------------------------------
// TE own elements per tile.  Quad4 / Tri3: 128 (one thread per element).  Quad9: 32 elements, FIVE threads
// per element (its 45 node-pair blocks = 135 accumulators do not fit one thread).
__host__ __device__ constexpr int tile_elems(int nnpe) { return nnpe == 9 ? 32 : FE_TILE_ELEMS; }
// halo element threads per CTA.  Quad4 threads need ~124 registers, so its CTA is kept at 5 warps
// (128 + 32 threads, 15 warps per SM => 128 registers available); Tri3 is register-light.
__host__ __device__ constexpr int halo_threads(int nnpe) { return nnpe == 4 ? FE_HALO_ELEMS : nnpe == 9 ? 32 : 64; }
// Slab rows for halo elements.  Quad4 tiles may carry up to 16 more halo elements than the CTA has halo
// threads (irregular tiles of perturbed / unstructured meshes); those rows are evaluated in a short second
// pass by the halo warp, so regular tiles pay nothing.
__host__ __device__ constexpr int halo_cap(int nnpe) { return nnpe == 4 ? FE_HALO_ELEMS + 16 : nnpe == 9 ? 16 : 64; }
__host__ __device__ constexpr int tile_parts(int nnpe) { return nnpe == 9 ? 5 : nnpe == 4 ? FE_Q4_PARTS : 1; }   // threads per element
__host__ __device__ constexpr int max_nodes(int nnpe) { return nnpe == 9 ? 192 : 2 * tile_elems(nnpe); }
__host__ __device__ constexpr int max_items(int nnpe) { return nnpe == 9 ? 640 : 8 * tile_elems(nnpe); }
__host__ __device__ constexpr int max_work(int nnpe) { return nnpe == 9 ? 3072 : 12 * tile_elems(nnpe); }   // x32
constexpr int kTileCtasPerSm = FE_TILE_CTAS;     // persistent grid = SMs x this
constexpr unsigned kWorkFirst = 1u << 28, kWorkSelf = 1u << 29;    // flag bits of a work record (above its two 14-bit sources)

// Warp-aggregated counter bump: lanes of the (converged) calling set that hit the same counter do ONE atomic between them.
// Returns the value this lane would have got from its own atomicAdd(counter + idx, 1) in lane order.  idx < 0: no bump.
__device__ __forceinline__ int warp_agg_inc(int32_t* __restrict__ counter, int idx) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, idx);
    if (idx < 0) return 0;
    const int lane = (int)(threadIdx.x & 31u);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter + idx, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

// Ascending sort of 32 * R keys held R per lane (key index = 32 * r + lane; pad with INT_MAX): bitonic network, exchanges
// at distance >= 32 are register to register, the others one shuffle per key.  All 32 lanes must call it.
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(int (&v)[R], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int q = r ^ (j >> 5);
                    if (q > r) {                                  // pair (r, q): ascending iff bit k of the index is clear
                        const bool up = ((32 * r) & k) == 0;
                        const int a = v[r], b = v[q];
                        const int mn = a < b ? a : b, mx = a < b ? b : a;
                        v[r] = up ? mn : mx;
                        v[q] = up ? mx : mn;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int idx = 32 * r + lane;
                    const int o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool up = (idx & k) == 0, lower = (lane & j) == 0;
                    const int mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
                    v[r] = (up == lower) ? mn : mx;
                }
            }
        }
    }
}

// block index of the pair lo <= hi in the upper triangle (row-major) of an N-node element
template <int N> __host__ __device__ constexpr int tri_index(int lo, int hi) { return lo * N - lo * (lo - 1) / 2 + (hi - lo); }

// Full upper-triangular block element matrix: block (lo,hi) = [L, P, Q] at 3*tri_index(lo,hi); diagonal blocks
// store Q = P.  Same pair arithmetic as element_row_pair (fe_element.cuh), so every kernel produces the same bits.
// SCALAR (FE_OP_POISSON): only the L entries are evaluated (same expressions, hence the same bits as the L
// entries of the full blocks); P and Q stay zero.
template <int ET, bool MASS, bool SCALAR = false>
__device__ __forceinline__ void element_tri(const double (&x)[Elem<ET>::NNPE], const double (&y)[Elem<ET>::NNPE],
                                            bool weighted,
                                            double (&out)[3 * (Elem<ET>::NNPE * (Elem<ET>::NNPE + 1) / 2)]) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    constexpr int kBlocks = N * (N + 1) / 2;
#pragma unroll
    for (int i = 0; i < 3 * kBlocks; ++i) out[i] = 0.0;
    ElemGeom<ET> G;
    geom_init<ET, MASS>(G, x, y, weighted);
#pragma unroll
    for (int g = 0; g < E::NG; ++g) {
        double ax[N], ay[N], sx[N], sy[N];
        gauss_point<ET, MASS>(g, G, weighted, ax, ay, sx, sy);
#pragma unroll
        for (int a = 0; a < N; ++a) {
#pragma unroll
            for (int b = a; b < N; ++b) {
                const int o = 3 * tri_index<N>(a, b);
                out[o] = fma(sx[a], ax[b], fma(sy[a], ay[b], out[o]));
                if constexpr (!SCALAR) {
                    out[o + 1] = fma(sy[a], ax[b], out[o + 1]);
                    if (b > a) out[o + 2] = fma(sx[a], ay[b], out[o + 2]);
                }
            }
        }
    }
    if constexpr (!SCALAR) {
#pragma unroll
        for (int a = 0; a < N; ++a) out[3 * tri_index<N>(a, a) + 2] = out[3 * tri_index<N>(a, a) + 1];
    }
}

}  // namespace fe
