// Error plumbing, identification, host-side table evaluation and the device-wide exclusive scan.
#include <stdarg.h>

#include "fe_common.cuh"

namespace fe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return FE_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Small device -> host read-back by a KERNEL writing into pinned (device-mapped) host memory.  A cudaMemcpy of a
// few bytes queues on the device->host copy engine behind whatever large download is in flight (the 268 MB
// right-hand side of cfg2 holds it for 5 ms); a store from a kernel does not.
// ------------------------------------------------------------------------------------------------
__global__ void k_peek(const uint32_t* __restrict__ src, volatile uint32_t* dst, int nwords) {
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan, int32.  Three phases per level: tile sums -> scan of tile sums (recursive) ->
// tile-local scan + offset.  Deterministic; one-time symbolic-phase utility, not a hot kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < (kScanThreads / 32)) smem_warp[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = wi;                         // block total
    }
    __syncthreads();
    total = smem_warp[32];
    return smem_warp[warp] + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                                 int32_t* __restrict__ sums) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += in[base + i];
    int total;
    block_exclusive_scan(s, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_apply(const int32_t* in, int32_t* out, int64_t n,
                                                                  const int32_t* __restrict__ offsets) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int total;
    int run = block_exclusive_scan(s, sm, total) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

static inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

size_t scan_workspace_bytes(int64_t n) {
    size_t total = 0;
    int64_t m = scan_tiles(n);
    while (m > 1) {
        total += (size_t)((m + 63) / 64 * 64) * sizeof(int32_t);
        m = scan_tiles(m);
    }
    return total + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, cudaStream_t stream) {
    if (n <= 0) return FE_OK;
    const int64_t tiles = scan_tiles(n);
    if (tiles == 1) {
        k_scan_tile_apply<<<1, kScanThreads, 0, stream>>>(in, out, n, nullptr);
        FE_CHECK_LAUNCH("k_scan_tile_apply");
        return FE_OK;
    }
    int32_t* sums = static_cast<int32_t*>(ws);
    k_scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, n, sums);
    FE_CHECK_LAUNCH("k_scan_tile_sums");
    int32_t* next_ws = sums + (tiles + 63) / 64 * 64;
    int rc = exclusive_scan_i32(sums, sums, tiles, next_ws, stream);
    if (rc != FE_OK) return rc;
    k_scan_tile_apply<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, n, sums);
    FE_CHECK_LAUNCH("k_scan_tile_apply");
    return FE_OK;
}

}  // namespace fe

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int fe_peek(const void* dev_src, void* pinned_host_dst, int64_t nbytes, void* stream) {
    using namespace fe;
    FE_REQUIRE(dev_src && pinned_host_dst && nbytes >= 0 && (nbytes & 3) == 0 && nbytes <= (1 << 20),
               "null pointer, or size not a multiple of 4 / above 1 MiB");
    if (nbytes == 0) return FE_OK;
    k_peek<<<1, 256, 0, as_stream(stream)>>>(static_cast<const uint32_t*>(dev_src),
                                             static_cast<volatile uint32_t*>(pinned_host_dst), (int)(nbytes >> 2));
    FE_CHECK_LAUNCH("k_peek");
    return FE_OK;
}

void fe_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return;
    strncpy(buf, fe::g_err, len - 1);
    buf[len - 1] = '\0';
}

const char* fe_version(void) { return "feprogram_b200 0.1.0 sm_100a"; }

void fe_limits(int32_t out[4]) {
    out[0] = fe::kMaxAdj;
    out[1] = fe::kMaxInc;
    out[2] = fe::kMaxNnpe;
    out[3] = fe::kMaxGauss;
}

size_t fe_scan_workspace_bytes(int64_t n) { return fe::scan_workspace_bytes(n); }

// ------------------------------------------------------------------------------------------------
// T1-T4.  Expression order mirrors tools/interpFunctions.py so the doubles are bit-identical to the
// reference's tables (checked in tests/test_capi_cpu.py against the golden fixtures).
// ------------------------------------------------------------------------------------------------
static void shape4(double r, double s, double* h, double* hr, double* hs) {
    const double rp = 1 + r, rm = 1 - r, sp = 1 + s, sm = 1 - s;
    if (h) {  // interpFunctions.py:3-15
        h[0] = 0.25 * rp * sp;
        h[1] = 0.25 * rm * sp;
        h[2] = 0.25 * rm * sm;
        h[3] = 0.25 * rp * sm;
    }
    if (hr) {  // interpFunctions.py:39-53
        const double a[4] = {1 + s, -1 - s, -1 + s, 1 - s};
        const double b[4] = {1 + r, 1 - r, -1 + r, -1 - r};
        for (int i = 0; i < 4; ++i) {
            hr[i] = 0.25 * a[i];
            hs[i] = 0.25 * b[i];
        }
    }
}

static void shape9(double r, double s, double* h, double* hr, double* hs) {
    const double rq = 1 - r * r, sq = 1 - s * s;
    const double rp = 1 + r, rm = 1 - r, sp = 1 + s, sm = 1 - s;
    if (h) {  // interpFunctions.py:17-37
        const double h9 = rq * sq;
        const double h8 = 0.5 * sq * rp - 0.5 * h9;
        const double h7 = 0.5 * rq * sm - 0.5 * h9;
        const double h6 = 0.5 * sq * rm - 0.5 * h9;
        const double h5 = 0.5 * rq * sp - 0.5 * h9;
        h[8] = h9; h[7] = h8; h[6] = h7; h[5] = h6; h[4] = h5;
        h[3] = 0.25 * rp * sm - 0.5 * h7 - 0.5 * h8 - 0.25 * h9;
        h[2] = 0.25 * rm * sm - 0.5 * h6 - 0.5 * h7 - 0.25 * h9;
        h[1] = 0.25 * rm * sp - 0.5 * h5 - 0.5 * h6 - 0.25 * h9;
        h[0] = 0.25 * rp * sp - 0.5 * h5 - 0.5 * h8 - 0.25 * h9;
    }
    if (hr) {  // interpFunctions.py:55-84
        const double a4[4] = {1 + s, -1 - s, -1 + s, 1 - s};
        const double a8[4] = {-2 * r * (1 + s), -1 * (1 - s * s), -2 * r * (1 - s), (1 - s * s)};
        const double a9 = -2 * r * (1 - s * s);
        const double b4[4] = {1 + r, 1 - r, -1 + r, -1 - r};
        const double b8[4] = {(1 - r * r), -2 * s * (1 - r), -1 * (1 - r * r), -2 * s * (1 + r)};
        const double b9 = -2 * s * (1 - r * r);
        for (int i = 0; i < 4; ++i) {
            hr[i + 4] = 0.5 * a8[i] - 0.5 * a9;
            hs[i + 4] = 0.5 * b8[i] - 0.5 * b9;
        }
        hr[8] = a9;
        hs[8] = b9;
        for (int i = 1; i < 4; ++i) {
            hr[i] = 0.25 * a4[i] - 0.5 * hr[i + 4] - 0.5 * hr[i + 3] - 0.25 * a9;
            hs[i] = 0.25 * b4[i] - 0.5 * hs[i + 4] - 0.5 * hs[i + 3] - 0.25 * b9;
        }
        hr[0] = 0.25 * a4[0] - 0.5 * hr[4] - 0.5 * hr[7] - 0.25 * a9;
        hs[0] = 0.25 * b4[0] - 0.5 * hs[4] - 0.5 * hs[7] - 0.25 * b9;
    }
}

int fe_tables(int elem_type, double* dH, double* H, double* w, int32_t* nnpe_out, int32_t* ng2_out) {
    const int nnpe = fe::nnpe_of(elem_type);
    const int ng2 = fe::ngauss_of(elem_type);
    if (!nnpe) {
        fe::set_error("fe_tables: Invalid elemType %d, use FE_QUAD4, FE_QUAD9 or FE_TRI3", elem_type);
        return FE_ERR_UNSUPPORTED;  // elements.py:92 raises on unknown type
    }
    if (nnpe_out) *nnpe_out = nnpe;
    if (ng2_out) *ng2_out = ng2;
    if (elem_type == FE_TRI3) {  // extension: one centroid point, area weight 1/2
        if (dH) {
            const double t[6] = {-1.0, 1.0, 0.0, -1.0, 0.0, 1.0};
            memcpy(dH, t, sizeof(t));
        }
        if (H) H[0] = H[1] = H[2] = 1.0 / 3;
        if (w) w[0] = 0.5;
        return FE_OK;
    }
    // elements.py:80-81, 86-87: truncated Gauss points, weights computed but never applied by getKe
    const double gp4[2] = {-0.5773, 0.5773}, w4[2] = {1, 1};
    const double gp9[3] = {-0.774, 0, 0.774}, w9[3] = {0.55555, 0.88888, 0.55555};
    const int ng = elem_type == FE_QUAD4 ? 2 : 3;
    const double* gp = elem_type == FE_QUAD4 ? gp4 : gp9;
    const double* gw = elem_type == FE_QUAD4 ? w4 : w9;
    for (int ir = 0; ir < ng; ++ir)
        for (int is = 0; is < ng; ++is) {  // r outer, s inner: g = ir*ng + is (elements.py:131-139)
            const int g = ir * ng + is;
            double hr[9], hs[9], h[9];
            if (elem_type == FE_QUAD4) shape4(gp[ir], gp[is], h, hr, hs);
            else shape9(gp[ir], gp[is], h, hr, hs);
            for (int a = 0; a < nnpe; ++a) {
                if (dH) {
                    dH[(g * 2 + 0) * nnpe + a] = hr[a];
                    dH[(g * 2 + 1) * nnpe + a] = hs[a];
                }
                if (H) H[g * nnpe + a] = h[a];
            }
            if (w) w[g] = gw[ir] * gw[is];
        }
    return FE_OK;
}

}  // extern "C"
