// Per-element arithmetic of the hot path (device functions), tables in __constant__ memory.
//
// Restates Elem.funB + FemProblem.getKe (/root/reference/elements.py:15-23, 153-163):
//   J = dH_g X, det = |J|, Dh = J^-1 dH_g, Ke += B^T B det with
//   B rows [Dx on u | Dy on v | Dy on u + Dx on v | 0]  =>  per node pair (a,b)
//     Ke[2a  ,2b  ] = Ke[2a+1,2b+1] = sum_g (Dx_a Dx_b + Dy_a Dy_b) det     ("L")
//     Ke[2a  ,2b+1] = sum_g Dy_a Dx_b det                                    ("P")
//     Ke[2a+1,2b  ] = sum_g Dx_a Dy_b det                                    ("Q")
// evaluated through the adjugate A = det * J^-1 dH so that each Gauss point needs one division:
//   (Dx_a Dx_b + ..) det = (Ax_a Ax_b + ..) / det.
// Mass extension (FE_OP_MASS; the reference builds the shape-value tables H and the weights gpWei but never
// uses them, elements.py:94-113): Me[2a+c, 2b+c] = sum_g gpWei_g H_g[a] H_g[b] det_g, zero across components.
// It runs through the same (L, P, Q) pair arithmetic with "ax" = H_g, "sx" = H_g * gpWei_g * det_g and
// ay = sy = 0, so every assembly kernel handles it unchanged and K and M share one CSR pattern.
// fp64 throughout, no tensor cores (blocks are 8x8 / 18x18 and the kernel is HBM-bound).
#pragma once

#include "fe_common.cuh"

namespace fe {

// One private copy per translation unit (no relocatable device code needed).
static __constant__ TablesQ4 c_q4;
static __constant__ TablesQ9 c_q9;
static __constant__ TablesT3 c_t3;

// Uploads this translation unit's tables once per device.
static int ensure_tables_tu(cudaStream_t stream) {
    static bool done[64] = {false};
    int dev = 0;
    FE_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (done[dev]) return FE_OK;
    TablesQ4 q4;
    TablesQ9 q9;
    TablesT3 t3;
    fe_tables(FE_QUAD4, &q4.dH[0][0][0], &q4.H[0][0], q4.w, nullptr, nullptr);
    fe_tables(FE_QUAD9, &q9.dH[0][0][0], &q9.H[0][0], q9.w, nullptr, nullptr);
    fe_tables(FE_TRI3, &t3.dH[0][0][0], &t3.H[0][0], t3.w, nullptr, nullptr);
    FE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_q4, &q4, sizeof(q4), 0, cudaMemcpyHostToDevice, stream));
    FE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_q9, &q9, sizeof(q9), 0, cudaMemcpyHostToDevice, stream));
    FE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_t3, &t3, sizeof(t3), 0, cudaMemcpyHostToDevice, stream));
    FE_CUDA_TRY(cudaStreamSynchronize(stream));  // host buffers are on this stack frame
    done[dev] = true;
    return FE_OK;
}

template <int ET> struct Elem;
template <> struct Elem<FE_QUAD4> {
    static constexpr int NNPE = 4, NG = 4;
    __device__ __forceinline__ static double dHr(int g, int a) { return c_q4.dH[g][0][a]; }
    __device__ __forceinline__ static double dHs(int g, int a) { return c_q4.dH[g][1][a]; }
    __device__ __forceinline__ static double w(int g) { return c_q4.w[g]; }
    __device__ __forceinline__ static double H(int g, int a) { return c_q4.H[g][a]; }
};
template <> struct Elem<FE_QUAD9> {
    static constexpr int NNPE = 9, NG = 9;
    __device__ __forceinline__ static double dHr(int g, int a) { return c_q9.dH[g][0][a]; }
    __device__ __forceinline__ static double dHs(int g, int a) { return c_q9.dH[g][1][a]; }
    __device__ __forceinline__ static double w(int g) { return c_q9.w[g]; }
    __device__ __forceinline__ static double H(int g, int a) { return c_q9.H[g][a]; }
};
template <> struct Elem<FE_TRI3> {
    static constexpr int NNPE = 3, NG = 1;
    __device__ __forceinline__ static double dHr(int g, int a) { return c_t3.dH[g][0][a]; }
    __device__ __forceinline__ static double dHs(int g, int a) { return c_t3.dH[g][1][a]; }
    __device__ __forceinline__ static double w(int g) { return c_t3.w[g]; }
    __device__ __forceinline__ static double H(int g, int a) { return c_t3.H[g][a]; }
};

// Gathers the element's node ids and coordinates (Elem.getpos, elements.py:34-38).
template <int ET>
__device__ __forceinline__ void load_element(const int32_t* __restrict__ conn, const double2* __restrict__ coords,
                                             int64_t e, int (&t)[Elem<ET>::NNPE], double (&x)[Elem<ET>::NNPE],
                                             double (&y)[Elem<ET>::NNPE]) {
    constexpr int N = Elem<ET>::NNPE;
    if constexpr (N == 4) {
        const int4 c = __ldg(reinterpret_cast<const int4*>(conn) + e);
        t[0] = c.x; t[1] = c.y; t[2] = c.z; t[3] = c.w;
    } else {
#pragma unroll
        for (int a = 0; a < N; ++a) t[a] = __ldg(conn + e * N + a);
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double2 p = __ldg(coords + t[a]);
        x[a] = p.x;
        y[a] = p.y;
    }
}

// Per-element geometry state shared by all Gauss points.
//
// Generic elements keep the node coordinates; every Gauss point evaluates J = dH_g X from the tables and
// its own reciprocal determinant.
//
// Quad4 uses the structure of the bilinear tables: with p = (1+t)/4, m = (1-t)/4 (t = the Gauss abscissa;
// both values are read from the reference-exact table, so the truncated abscissae are honoured)
//     dH/dr = [ps, -ps, -ms, ms],   dH/ds = [pr, mr, -mr, -pr]
// so J only needs four coordinate differences per axis and two table values, the adjugate gradients are
// sums/differences of eight products, and the four reciprocal determinants come from ONE division:
//     1/det_g = (prod_{h != g} det_h) / (det_0 det_1 det_2 det_3).
// Same mathematics as the table-driven path (differences are a few ulp); roughly 35 % fewer fp64
// instructions and three divisions less per element.
template <int ET> struct ElemGeom {
    double x[Elem<ET>::NNPE], y[Elem<ET>::NNPE];
};
template <> struct ElemGeom<FE_QUAD4> {
    double dx01, dx32, dx03, dx12, dy01, dy32, dy03, dy12;  // coordinate differences
    double s[4];                                            // weight_g / det_g
};

template <int ET, bool MASS = false>
__device__ __forceinline__ void geom_init(ElemGeom<ET>& G, const double (&x)[Elem<ET>::NNPE],
                                          const double (&y)[Elem<ET>::NNPE], bool weighted) {
#pragma unroll
    for (int a = 0; a < Elem<ET>::NNPE; ++a) {
        G.x[a] = x[a];
        G.y[a] = y[a];
    }
}

// (p, m) table values of Gauss point g along s (is = g & 1) and r (ir = g >> 1)
__device__ __forceinline__ void q4_table(int g, double& ps, double& ms, double& pr, double& mr) {
    const double al = c_q4.dH[3][0][0];  // (1 + gamma) / 4
    const double be = c_q4.dH[0][0][0];  // (1 - gamma) / 4
    ps = (g & 1) ? al : be;
    ms = (g & 1) ? be : al;
    pr = (g >> 1) ? al : be;
    mr = (g >> 1) ? be : al;
}

__device__ __forceinline__ void q4_jacobian(int g, const ElemGeom<FE_QUAD4>& G, double& j00, double& j01, double& j10,
                                            double& j11) {
    double ps, ms, pr, mr;
    q4_table(g, ps, ms, pr, mr);
    j00 = fma(ps, G.dx01, __dmul_rn(ms, G.dx32));
    j01 = fma(ps, G.dy01, __dmul_rn(ms, G.dy32));
    j10 = fma(pr, G.dx03, __dmul_rn(mr, G.dx12));
    j11 = fma(pr, G.dy03, __dmul_rn(mr, G.dy12));
}

template <bool MASS>
__device__ __forceinline__ void geom_init_q4(ElemGeom<FE_QUAD4>& G, const double (&x)[4], const double (&y)[4],
                                             bool weighted) {
    G.dx01 = __dsub_rn(x[0], x[1]); G.dx32 = __dsub_rn(x[3], x[2]);
    G.dx03 = __dsub_rn(x[0], x[3]); G.dx12 = __dsub_rn(x[1], x[2]);
    G.dy01 = __dsub_rn(y[0], y[1]); G.dy32 = __dsub_rn(y[3], y[2]);
    G.dy03 = __dsub_rn(y[0], y[3]); G.dy12 = __dsub_rn(y[1], y[2]);
    double d[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        double j00, j01, j10, j11;
        q4_jacobian(g, G, j00, j01, j10, j11);
        d[g] = fma(j00, j11, -__dmul_rn(j01, j10));
    }
    if constexpr (MASS) {   // the mass operator scales by det itself
#pragma unroll
        for (int g = 0; g < 4; ++g) G.s[g] = __dmul_rn(d[g], c_q4.w[g]);
        return;
    }
    const double d01 = __dmul_rn(d[0], d[1]), d23 = __dmul_rn(d[2], d[3]);
    const double inv = 1.0 / __dmul_rn(d01, d23);
    const double i01 = __dmul_rn(inv, d23), i23 = __dmul_rn(inv, d01);   // 1/(d0 d1), 1/(d2 d3)
    G.s[0] = __dmul_rn(i01, d[1]);
    G.s[1] = __dmul_rn(i01, d[0]);
    G.s[2] = __dmul_rn(i23, d[3]);
    G.s[3] = __dmul_rn(i23, d[2]);
    if (weighted) {
#pragma unroll
        for (int g = 0; g < 4; ++g) G.s[g] = __dmul_rn(G.s[g], c_q4.w[g]);
    }
}

template <> __device__ __forceinline__ void geom_init<FE_QUAD4, false>(ElemGeom<FE_QUAD4>& G, const double (&x)[4],
                                                                       const double (&y)[4], bool weighted) {
    geom_init_q4<false>(G, x, y, weighted);
}
template <> __device__ __forceinline__ void geom_init<FE_QUAD4, true>(ElemGeom<FE_QUAD4>& G, const double (&x)[4],
                                                                      const double (&y)[4], bool weighted) {
    geom_init_q4<true>(G, x, y, weighted);
}

// Adjugate gradients at Gauss point g:  ax_b = J11 dHr_b - J01 dHs_b, ay_b = J00 dHs_b - J10 dHr_b and
// their scaled copies sx_b = ax_b / det, sy_b = ay_b / det (times the Gauss weight when weighted).
template <int ET, bool MASS = false>
__device__ __forceinline__ void gauss_point(int g, const ElemGeom<ET>& G, bool weighted,
                                            double (&ax)[Elem<ET>::NNPE], double (&ay)[Elem<ET>::NNPE],
                                            double (&sx)[Elem<ET>::NNPE], double (&sy)[Elem<ET>::NNPE]) {
    using E = Elem<ET>;
    double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) {
        j00 = fma(E::dHr(g, b), G.x[b], j00);
        j01 = fma(E::dHr(g, b), G.y[b], j01);
        j10 = fma(E::dHs(g, b), G.x[b], j10);
        j11 = fma(E::dHs(g, b), G.y[b], j11);
    }
    // every product / sum below is written with explicit roundings so that all kernels that evaluate an
    // element (tiled, row-tile, multi-thread Quad9) produce the same bits regardless of FMA contraction
    const double det = fma(j00, j11, -__dmul_rn(j01, j10));
    if constexpr (MASS) {
        const double m = __dmul_rn(det, E::w(g));
#pragma unroll
        for (int b = 0; b < E::NNPE; ++b) {
            ax[b] = E::H(g, b);
            ay[b] = 0.0;
            sx[b] = __dmul_rn(ax[b], m);
            sy[b] = 0.0;
        }
        return;
    }
    double s = 1.0 / det;
    if (weighted) s = __dmul_rn(s, E::w(g));
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) {
        ax[b] = fma(j11, E::dHr(g, b), -__dmul_rn(j01, E::dHs(g, b)));
        ay[b] = fma(j00, E::dHs(g, b), -__dmul_rn(j10, E::dHr(g, b)));
        sx[b] = __dmul_rn(ax[b], s);
        sy[b] = __dmul_rn(ay[b], s);
    }
}

template <>
__device__ __forceinline__ void gauss_point<FE_QUAD4, true>(int g, const ElemGeom<FE_QUAD4>& G, bool weighted,
                                                            double (&ax)[4], double (&ay)[4], double (&sx)[4],
                                                            double (&sy)[4]) {
    const double m = G.s[g];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        ax[b] = c_q4.H[g][b];
        ay[b] = 0.0;
        sx[b] = __dmul_rn(ax[b], m);
        sy[b] = 0.0;
    }
}

template <>
__device__ __forceinline__ void gauss_point<FE_QUAD4, false>(int g, const ElemGeom<FE_QUAD4>& G, bool weighted,
                                                             double (&ax)[4], double (&ay)[4], double (&sx)[4],
                                                             double (&sy)[4]) {
    double ps, ms, pr, mr, j00, j01, j10, j11;
    q4_table(g, ps, ms, pr, mr);
    q4_jacobian(g, G, j00, j01, j10, j11);
    const double u1 = __dmul_rn(j11, ps), u2 = __dmul_rn(j11, ms), v1 = __dmul_rn(j01, pr), v2 = __dmul_rn(j01, mr);
    const double w1 = __dmul_rn(j00, pr), w2 = __dmul_rn(j00, mr), z1 = __dmul_rn(j10, ps), z2 = __dmul_rn(j10, ms);
    ax[0] = __dsub_rn(u1, v1); ax[1] = -__dadd_rn(u1, v2); ax[2] = __dsub_rn(v2, u2); ax[3] = __dadd_rn(u2, v1);
    ay[0] = __dsub_rn(w1, z1); ay[1] = __dadd_rn(w2, z1);  ay[2] = __dsub_rn(z2, w2); ay[3] = -__dadd_rn(w1, z2);
    const double s = G.s[g];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        sx[b] = __dmul_rn(ax[b], s);
        sy[b] = __dmul_rn(ay[b], s);
    }
}

// Canonical pair arithmetic.  For a node pair lo <= hi and Gauss point g
//     L(lo,hi) += sx_lo ax_hi + sy_lo ay_hi,   P(lo,hi) += sy_lo ax_hi,   Q(lo,hi) += sx_lo ay_hi
// and the mirrored block is its transpose: L(hi,lo) = L(lo,hi), P(hi,lo) = Q(lo,hi), Q(hi,lo) = P(lo,hi).
// Every code path (full symmetric element, single row pair) evaluates a pair with exactly these
// operations in this order, so a matrix entry has the same bits no matter which path produced it and
// the assembled K is bit-wise symmetric.

// Symmetric-compressed element matrix: diagonal pairs (a,a) store [L, P] at 2a (Q == P there);
// off-diagonal pairs a < b store [L, P, Q] at 2N + 3*pair_index(a,b).
template <int N> struct SymLayout {
    static constexpr int kSize = 2 * N + 3 * (N * (N - 1) / 2);
    __host__ __device__ static constexpr int pair_index(int a, int b) { return a * (2 * N - a - 1) / 2 + (b - a - 1); }
    __host__ __device__ static constexpr int offdiag(int a, int b) { return 2 * N + 3 * pair_index(a, b); }
};

template <int ET, bool MASS = false>
__device__ __forceinline__ void element_sym(const double (&x)[Elem<ET>::NNPE], const double (&y)[Elem<ET>::NNPE],
                                            bool weighted, double (&out)[SymLayout<Elem<ET>::NNPE>::kSize]) {
    using E = Elem<ET>;
    using S = SymLayout<E::NNPE>;
#pragma unroll
    for (int i = 0; i < S::kSize; ++i) out[i] = 0.0;
    ElemGeom<ET> G;
    geom_init<ET, MASS>(G, x, y, weighted);
#pragma unroll
    for (int g = 0; g < E::NG; ++g) {
        double ax[E::NNPE], ay[E::NNPE], sx[E::NNPE], sy[E::NNPE];
        gauss_point<ET, MASS>(g, G, weighted, ax, ay, sx, sy);
#pragma unroll
        for (int a = 0; a < E::NNPE; ++a) {
            out[2 * a] = fma(sx[a], ax[a], fma(sy[a], ay[a], out[2 * a]));
            out[2 * a + 1] = fma(sy[a], ax[a], out[2 * a + 1]);
#pragma unroll
            for (int b = a + 1; b < E::NNPE; ++b) {
                const int o = S::offdiag(a, b);
                out[o] = fma(sx[a], ax[b], fma(sy[a], ay[b], out[o]));
                out[o + 1] = fma(sy[a], ax[b], out[o + 1]);
                out[o + 2] = fma(sx[a], ay[b], out[o + 2]);
            }
        }
    }
}

// (L, P, Q) of block (a, b) out of a symmetric-compressed element matrix (run-time a, b).
template <int N>
__device__ __forceinline__ void sym_block(const double* __restrict__ s, int a, int b, double& L, double& P, double& Q) {
    using S = SymLayout<N>;
    if (a == b) {
        L = s[2 * a];
        P = s[2 * a + 1];
        Q = P;
    } else if (a < b) {
        const int o = S::offdiag(a, b);
        L = s[o];
        P = s[o + 1];
        Q = s[o + 2];
    } else {
        const int o = S::offdiag(b, a);
        L = s[o];
        P = s[o + 2];
        Q = s[o + 1];
    }
}

// Row pair of local node a:  L[b], P[b], Q[b] for all b.  `a` is a run-time value; operands are picked
// with selects so everything stays in registers.  Same pair arithmetic as element_sym.
template <int ET, bool MASS = false>
__device__ __forceinline__ void element_row_pair(const double (&x)[Elem<ET>::NNPE],
                                                 const double (&y)[Elem<ET>::NNPE], int a, bool weighted,
                                                 double (&L)[Elem<ET>::NNPE], double (&P)[Elem<ET>::NNPE],
                                                 double (&Q)[Elem<ET>::NNPE]) {
    using E = Elem<ET>;
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) L[b] = P[b] = Q[b] = 0.0;
    ElemGeom<ET> G;
    geom_init<ET, MASS>(G, x, y, weighted);
#pragma unroll
    for (int g = 0; g < E::NG; ++g) {
        double ax[E::NNPE], ay[E::NNPE], sx[E::NNPE], sy[E::NNPE];
        gauss_point<ET, MASS>(g, G, weighted, ax, ay, sx, sy);
        double axa = 0.0, aya = 0.0, sxa = 0.0, sya = 0.0;
#pragma unroll
        for (int b = 0; b < E::NNPE; ++b) {
            axa = (b == a) ? ax[b] : axa;
            aya = (b == a) ? ay[b] : aya;
            sxa = (b == a) ? sx[b] : sxa;
            sya = (b == a) ? sy[b] : sya;
        }
#pragma unroll
        for (int b = 0; b < E::NNPE; ++b) {
            const bool lo = a <= b;  // the lower-numbered local node carries the 1/det scaling
            const double ux = lo ? sxa : sx[b], uy = lo ? sya : sy[b];   // scaled (lo) operands
            const double vx = lo ? ax[b] : axa, vy = lo ? ay[b] : aya;   // plain (hi) operands
            L[b] = fma(ux, vx, fma(uy, vy, L[b]));
            // lo == a: P = sy_a ax_b, Q = sx_a ay_b.   lo == b: P(a,b) = Q(b,a) = sx_b ay_a, Q(a,b) = P(b,a) = sy_b ax_a
            P[b] = fma(lo ? uy : ux, lo ? vx : vy, P[b]);
            Q[b] = fma(lo ? ux : uy, lo ? vy : vx, Q[b]);
        }
    }
    // diagonal pair: Q(a,a) == P(a,a) mathematically; use the one stored value (as element_sym does)
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) Q[b] = (b == a) ? P[b] : Q[b];
}

}  // namespace fe
