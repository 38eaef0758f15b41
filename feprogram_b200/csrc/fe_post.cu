// Post-processing on the solved field: the velocity-gradient operator the reference defines but never calls
// (Elem.getHgrad, /root/reference/elements.py:25-32) evaluated at every Gauss point of every element, and
// the nodal strain triple the reference's writer reserves a (commented-out) "Tensiones" array for
// (/root/reference/python_to_xml.py:19-22).  Both re-use the assembly's Dh = J^-1 dH (elements.py:15-23).
#include "fe_element.cuh"

namespace fe {

// J^-1 dH at Gauss point g of an element: Dx[a], Dy[a]  (Elem.funB / getHgrad: Dh = inv(J) * dHrs)
template <int ET>
__device__ __forceinline__ void gradients_at(int g, const double (&x)[Elem<ET>::NNPE], const double (&y)[Elem<ET>::NNPE],
                                             double (&Dx)[Elem<ET>::NNPE], double (&Dy)[Elem<ET>::NNPE]) {
    using E = Elem<ET>;
    double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) {
        j00 = fma(E::dHr(g, b), x[b], j00);
        j01 = fma(E::dHr(g, b), y[b], j01);
        j10 = fma(E::dHs(g, b), x[b], j10);
        j11 = fma(E::dHs(g, b), y[b], j11);
    }
    const double inv = 1.0 / (j00 * j11 - j01 * j10);
#pragma unroll
    for (int b = 0; b < E::NNPE; ++b) {
        Dx[b] = (j11 * E::dHr(g, b) - j01 * E::dHs(g, b)) * inv;
        Dy[b] = (j00 * E::dHs(g, b) - j10 * E::dHr(g, b)) * inv;
    }
}

// grad[e][g][0..3] = H_grad_g u_e = (du/dx, du/dy, dv/dx, dv/dy); one thread per (element, Gauss point)
template <int ET>
__global__ void __launch_bounds__(128) k_gauss_gradients(int64_t nelem, const int32_t* __restrict__ conn,
                                                         const double2* __restrict__ coords,
                                                         const double2* __restrict__ U, double* __restrict__ grad) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE, NG = E::NG;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nelem * NG) return;
    const int64_t e = idx / NG;
    const int g = (int)(idx - e * NG);
    int t[N];
    double x[N], y[N], Dx[N], Dy[N];
    load_element<ET>(conn, coords, e, t, x, y);
    gradients_at<ET>(g, x, y, Dx, Dy);
    double ux = 0.0, uy = 0.0, vx = 0.0, vy = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double2 u = __ldg(U + t[a]);
        ux = fma(Dx[a], u.x, ux);
        uy = fma(Dy[a], u.x, uy);
        vx = fma(Dx[a], u.y, vx);
        vy = fma(Dy[a], u.y, vy);
    }
    double2* o = reinterpret_cast<double2*>(grad + 4 * idx);
    o[0] = make_double2(ux, uy);
    o[1] = make_double2(vx, vy);
}

// Nodal strain: for every node the mean over its incident elements (ascending element id -- fixed order, no
// atomics) of the element's Gauss-point mean of B u_e = (du/dx, dv/dy, du/dy + dv/dx)  (Elem.funB rows 0-2).
template <int ET>
__global__ void __launch_bounds__(128) k_nodal_strain(int64_t nnode, const int32_t* __restrict__ eptr,
                                                      const int32_t* __restrict__ einc,
                                                      const int32_t* __restrict__ conn,
                                                      const double2* __restrict__ coords,
                                                      const double2* __restrict__ U, double* __restrict__ out) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE, NG = E::NG;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int i0 = eptr[n], i1 = eptr[n + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = i0; i < i1; ++i) {
        const int64_t e = einc[i] >> 4;   // einc packs element * 16 + local node
        int t[N];
        double x[N], y[N], ue[N], ve[N];
        load_element<ET>(conn, coords, e, t, x, y);
#pragma unroll
        for (int a = 0; a < N; ++a) {
            const double2 u = __ldg(U + t[a]);
            ue[a] = u.x;
            ve[a] = u.y;
        }
        double e0 = 0.0, e1 = 0.0, e2 = 0.0;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            double Dx[N], Dy[N];
            gradients_at<ET>(g, x, y, Dx, Dy);
#pragma unroll
            for (int a = 0; a < N; ++a) {
                e0 = fma(Dx[a], ue[a], e0);
                e1 = fma(Dy[a], ve[a], e1);
                e2 = fma(Dy[a], ue[a], fma(Dx[a], ve[a], e2));
            }
        }
        s0 += e0 / NG;
        s1 += e1 / NG;
        s2 += e2 / NG;
    }
    const double c = i1 > i0 ? 1.0 / (double)(i1 - i0) : 0.0;
    out[3 * n] = s0 * c;
    out[3 * n + 1] = s1 * c;
    out[3 * n + 2] = s2 * c;
}

}  // namespace fe

extern "C" {

int fe_gauss_gradients(int elem_type, int64_t nelem, const int32_t* conn, const double* coords, const double* U,
                       double* grad, void* stream) {
    using namespace fe;
    FE_REQUIRE(nelem >= 0, "negative size");
    if (elem_type != FE_QUAD4 && elem_type != FE_QUAD9 && elem_type != FE_TRI3) {
        set_error("fe_gauss_gradients: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    if (nelem == 0) return FE_OK;
    FE_REQUIRE(conn && coords && U && grad, "null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const double2* xy = reinterpret_cast<const double2*>(coords);
    const double2* u = reinterpret_cast<const double2*>(U);
    switch (elem_type) {
        case FE_QUAD4: k_gauss_gradients<FE_QUAD4><<<grid_for(nelem * 4, 128), 128, 0, st>>>(nelem, conn, xy, u, grad); break;
        case FE_QUAD9: k_gauss_gradients<FE_QUAD9><<<grid_for(nelem * 9, 128), 128, 0, st>>>(nelem, conn, xy, u, grad); break;
        default: k_gauss_gradients<FE_TRI3><<<grid_for(nelem, 128), 128, 0, st>>>(nelem, conn, xy, u, grad); break;
    }
    FE_CHECK_LAUNCH("k_gauss_gradients");
    return FE_OK;
}

int fe_nodal_strain(int elem_type, int64_t nnode, const int32_t* eptr, const int32_t* einc, const int32_t* conn,
                    const double* coords, const double* U, double* strain, void* stream) {
    using namespace fe;
    FE_REQUIRE(nnode >= 0, "negative size");
    if (elem_type != FE_QUAD4 && elem_type != FE_QUAD9 && elem_type != FE_TRI3) {
        set_error("fe_nodal_strain: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    if (nnode == 0) return FE_OK;
    FE_REQUIRE(eptr && einc && conn && coords && U && strain, "null pointer");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const double2* xy = reinterpret_cast<const double2*>(coords);
    const double2* u = reinterpret_cast<const double2*>(U);
    switch (elem_type) {
        case FE_QUAD4: k_nodal_strain<FE_QUAD4><<<grid_for(nnode, 128), 128, 0, st>>>(nnode, eptr, einc, conn, xy, u, strain); break;
        case FE_QUAD9: k_nodal_strain<FE_QUAD9><<<grid_for(nnode, 128), 128, 0, st>>>(nnode, eptr, einc, conn, xy, u, strain); break;
        default: k_nodal_strain<FE_TRI3><<<grid_for(nnode, 128), 128, 0, st>>>(nnode, eptr, einc, conn, xy, u, strain); break;
    }
    FE_CHECK_LAUNCH("k_nodal_strain");
    return FE_OK;
}

}  // extern "C"
