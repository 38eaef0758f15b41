// Shared definitions for the feprogram_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/feprogram_b200.h"

namespace fe {

// ---- capacities (reported by fe_limits) -------------------------------------------------------
constexpr int kMaxAdj = 96;    // distinct neighbour nodes per node (Quad9 corner: 25)
constexpr int kMaxInc = 32;    // incident elements per node
constexpr int kMaxNnpe = 9;
constexpr int kMaxGauss = 9;

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define FE_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) return fe::cuda_fail(_e, #expr); \
    } while (0)

#define FE_CHECK_LAUNCH(name)                                            \
    do {                                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return fe::cuda_fail(_e, "launch " name); \
    } while (0)

#define FE_REQUIRE(cond, msg)                        \
    do {                                             \
        if (!(cond)) {                               \
            fe::set_error("%s: %s", __func__, msg);  \
            return FE_ERR_INVALID;                   \
        }                                            \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int nnpe_of(int elem_type) {
    return (elem_type == FE_QUAD4 || elem_type == FE_QUAD9 || elem_type == FE_TRI3) ? elem_type : 0;
}
inline int ngauss_of(int elem_type) {
    return elem_type == FE_QUAD4 ? 4 : elem_type == FE_QUAD9 ? 9 : elem_type == FE_TRI3 ? 1 : 0;
}

// ---- element tables in constant memory --------------------------------------------------------
// dH[g][0][a] = dN_a/dr, dH[g][1][a] = dN_a/ds at Gauss point g; w[g] = tensor weight.
// One copy per element type so that fully unrolled kernels address them as c[bank][imm].
struct TablesQ4 { double dH[4][2][4]; double w[4]; double H[4][4]; };
struct TablesQ9 { double dH[9][2][9]; double w[9]; double H[9][9]; };
struct TablesT3 { double dH[1][2][3]; double w[1]; double H[1][3]; };

// Uploads the tables for the current device once (idempotent, cheap afterwards).
int ensure_tables(cudaStream_t stream);

// Grid sizing: ceil-div with 64-bit safety.
inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- exclusive scan (int32) -------------------------------------------------------------------
size_t scan_workspace_bytes(int64_t n);
// out[i] = sum_{j<i} in[j] for i in [0, n); in and out may alias.
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, cudaStream_t stream);

}  // namespace fe
