// K2: one-time symbolic phase -- node/element incidence, node adjacency (the CSR pattern at node
// granularity), the element -> CSR-slot map, and the scipy-compatible dof-level CSR pattern.
//
// Replaces getRowData (/root/reference/elements.py:143-151) and the structural half of
// coo_matrix(...).tolil() (elements.py:183).  Instead of one global sort over the
// nelem*(ndof*nnpe)^2 element->DOF pairs (1.07 G keys for a 4096^2 Quad4 mesh) the pairs are
// bucketed by node with a counting pass and sorted/uniqued per node -- a segmented sort + unique
// whose segments are tiny (<= 36 candidates), at node granularity (ndof^2 fewer keys).
// Integer atomics are used only to bucket; every bucket is sorted afterwards, so all outputs are
// deterministic.
#include "fe_common.cuh"

namespace fe {

__global__ void k_zero_i32(int32_t* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

__global__ void k_count_incidence(int64_t nslots, const int32_t* __restrict__ conn, int32_t* __restrict__ deg) {
    const int64_t i0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);     // four slots per thread
    if (i0 >= nslots) return;
    if (i0 + 3 < nslots && (reinterpret_cast<uintptr_t>(conn) & 15) == 0) {
        const int4 c = *reinterpret_cast<const int4*>(conn + i0);
        atomicAdd(deg + c.x, 1);
        atomicAdd(deg + c.y, 1);
        atomicAdd(deg + c.z, 1);
        atomicAdd(deg + c.w, 1);
    } else {
        for (int64_t i = i0; i < nslots && i < i0 + 4; ++i) atomicAdd(deg + conn[i], 1);
    }
}

__global__ void k_fill_incidence(int64_t nslots, int nnpe, const int32_t* __restrict__ conn,
                                 const int32_t* __restrict__ eptr, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ einc) {
    // four consecutive slots per thread: four independent (atomic slot, row start) -> store chains in flight
    const int64_t i0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= nslots) return;
    int n[4], slot[4], base[4];
    if (i0 + 3 < nslots && (reinterpret_cast<uintptr_t>(conn) & 15) == 0) {      // aligned groups of four: one load
        const int4 c = *reinterpret_cast<const int4*>(conn + i0);
        n[0] = c.x, n[1] = c.y, n[2] = c.z, n[3] = c.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) n[q] = i0 + q < nslots ? conn[i0 + q] : -1;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        slot[q] = base[q] = 0;
        if (n[q] >= 0) {
            slot[q] = atomicAdd(cursor + n[q], 1);
            base[q] = eptr[n[q]];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (n[q] < 0) continue;
        const int64_t i = i0 + q;
        const int e = (int)(i / nnpe);
        const int a = (int)(i - (int64_t)e * nnpe);
        einc[base[q] + slot[q]] = e * 16 + a;
    }
}

// In-place sort of each node's (short) incidence segment: ascending element id.  Segments of up to eight entries (every
// node of a 2-D mesh of decent quality) are loaded with independent loads, sorted in registers by a fixed network of
// compare-exchanges and written back -- one memory round trip instead of one per insertion step; longer ones take
// the insertion sort.
__global__ void k_sort_incidence(int64_t nnode, const int32_t* __restrict__ eptr, int32_t* __restrict__ einc) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int lo = eptr[n], hi = eptr[n + 1];
    const int cnt = hi - lo;
    if (cnt <= 1) return;
    if (cnt <= 8) {
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = i < cnt ? einc[lo + i] : 0x7fffffff;
        auto cx = [&](int& a, int& b) {
            const int x = a < b ? a : b, y = a < b ? b : a;
            a = x;
            b = y;
        };
        // 19-comparator network for eight keys (Batcher's odd-even merge sort)
        cx(v[0], v[1]); cx(v[2], v[3]); cx(v[4], v[5]); cx(v[6], v[7]);
        cx(v[0], v[2]); cx(v[1], v[3]); cx(v[4], v[6]); cx(v[5], v[7]);
        cx(v[1], v[2]); cx(v[5], v[6]);
        cx(v[0], v[4]); cx(v[1], v[5]); cx(v[2], v[6]); cx(v[3], v[7]);
        cx(v[2], v[4]); cx(v[3], v[5]);
        cx(v[1], v[2]); cx(v[3], v[4]); cx(v[5], v[6]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < cnt) einc[lo + i] = v[i];
        return;
    }
    for (int i = lo + 1; i < hi; ++i) {
        const int v = einc[i];
        int j = i;
        while (j > lo && einc[j - 1] > v) {
            einc[j] = einc[j - 1];
            --j;
        }
        einc[j] = v;
    }
}

// Sorted, distinct neighbour list of node n in thread-local storage.  Returns the count, or -1 when
// more than kMaxAdj distinct neighbours exist.
__device__ __forceinline__ int build_adjacency(int nnpe, const int32_t* __restrict__ conn,
                                               const int32_t* __restrict__ einc, int k0, int k1, int* list) {
    int cnt = 0;
    for (int k = k0; k < k1; ++k) {
        const int64_t e = einc[k] >> 4;
        for (int b = 0; b < nnpe; ++b) {
            const int v = conn[e * nnpe + b];
            int j = cnt;
            while (j > 0 && list[j - 1] > v) --j;
            if (j > 0 && list[j - 1] == v) continue;
            if (cnt == kMaxAdj) return -1;
            for (int i = cnt; i > j; --i) list[i] = list[i - 1];
            list[j] = v;
            ++cnt;
        }
    }
    return cnt;
}

// Fast form of build_adjacency: the list lives in shared memory, entry i of thread t at s[i * kAdjStride + t], at most
// kFastAdj entries.  The odd stride makes both views conflict free: a thread walking its own list (fixed t per lane,
// same i) and a warp reading the lists of its 32 nodes transposed (k_adjacency_fill's coalesced write-out).  Returns
// the count, or -1 when the node has more distinct neighbours than that (the caller then takes the general
// thread-local path above; structured and typical unstructured 2-D meshes never do: 9 / 25 / 7 neighbours for
// Quad4 / Quad9 / Tri3 lattices).
constexpr int kFastAdj = 32, kAdjStride = 129;
template <int NNPE>
__device__ __forceinline__ int build_adjacency_smem(const int32_t* __restrict__ conn, const int32_t* __restrict__ einc,
                                                    int k0, int k1, int* s) {
    int cnt = 0;
    constexpr int B = NNPE == 9 ? 2 : 4;          // incident elements fetched together: B independent load chains
    for (int kb = k0; kb < k1; kb += B) {
        int64_t e[B];
        int row[B][NNPE];
#pragma unroll
        for (int q = 0; q < B; ++q) e[q] = kb + q < k1 ? (int64_t)(einc[kb + q] >> 4) : -1;
#pragma unroll
        for (int q = 0; q < B; ++q) {
            if (e[q] < 0) continue;
            if constexpr (NNPE == 4) {
                const int4 r = __ldg(reinterpret_cast<const int4*>(conn) + e[q]);
                row[q][0] = r.x, row[q][1] = r.y, row[q][2] = r.z, row[q][3] = r.w;
            } else {
#pragma unroll
                for (int b = 0; b < NNPE; ++b) row[q][b] = __ldg(conn + e[q] * NNPE + b);
            }
        }
#pragma unroll
        for (int q = 0; q < B; ++q) {
            if (e[q] < 0) continue;
#pragma unroll
            for (int b = 0; b < NNPE; ++b) {
                const int v = row[q][b];
                int j = cnt;
                while (j > 0 && s[(j - 1) * kAdjStride] > v) --j;
                if (j > 0 && s[(j - 1) * kAdjStride] == v) continue;
                if (cnt == kFastAdj) return -1;
                for (int i = cnt; i > j; --i) s[i * kAdjStride] = s[(i - 1) * kAdjStride];
                s[j * kAdjStride] = v;
                ++cnt;
            }
        }
    }
    return cnt;
}

// Quad4 node with at most four incident elements (every regular node): its 16 candidate neighbours are sorted as packed
// keys (node id << 4 | candidate index: ids below 2^27) by a fixed 63-comparator network (Batcher's odd-even merge
// sort) in registers, duplicates are dropped by comparing neighbours -- no insertion into a list in memory, no
// searches: the rank of a key among the distinct ids is also the position the slot map (epos) needs for its candidate.
// keys[4 q + b] = node b of the q-th incident element; absent elements give 0xFFFFFFFF.  Returns the distinct count;
// rank[i] = position of sorted key i in the node's list, first[i] = key i starts a new id.
constexpr int64_t kPackedIdLimit = (int64_t)1 << 27;
__device__ __forceinline__ void quad4_sorted_candidates(const int32_t* __restrict__ conn, const int32_t* __restrict__ einc,
                                                        int k0, int ninc, unsigned (&key)[16]) {
    int e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) e[q] = q < ninc ? einc[k0 + q] >> 4 : -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int4 r = make_int4(-1, -1, -1, -1);
        if (e[q] >= 0) r = __ldg(reinterpret_cast<const int4*>(conn) + e[q]);
        key[4 * q + 0] = e[q] >= 0 ? ((unsigned)r.x << 4) | (unsigned)(4 * q + 0) : 0xFFFFFFFFu;
        key[4 * q + 1] = e[q] >= 0 ? ((unsigned)r.y << 4) | (unsigned)(4 * q + 1) : 0xFFFFFFFFu;
        key[4 * q + 2] = e[q] >= 0 ? ((unsigned)r.z << 4) | (unsigned)(4 * q + 2) : 0xFFFFFFFFu;
        key[4 * q + 3] = e[q] >= 0 ? ((unsigned)r.w << 4) | (unsigned)(4 * q + 3) : 0xFFFFFFFFu;
    }
    // Batcher's odd-even merge sort for 16 keys: 63 compare-exchanges on registers
#define CX(i, j)                                   \
    {                                              \
        const unsigned a_ = key[i], b_ = key[j];   \
        key[i] = a_ < b_ ? a_ : b_;                \
        key[j] = a_ < b_ ? b_ : a_;                \
    }
    CX(0, 1); CX(2, 3); CX(4, 5); CX(6, 7); CX(8, 9); CX(10, 11); CX(12, 13); CX(14, 15);
    CX(0, 2); CX(1, 3); CX(4, 6); CX(5, 7); CX(8, 10); CX(9, 11); CX(12, 14); CX(13, 15);
    CX(1, 2); CX(5, 6); CX(9, 10); CX(13, 14); CX(0, 4); CX(1, 5); CX(2, 6); CX(3, 7);
    CX(8, 12); CX(9, 13); CX(10, 14); CX(11, 15); CX(2, 4); CX(3, 5); CX(10, 12); CX(11, 13);
    CX(1, 2); CX(3, 4); CX(5, 6); CX(9, 10); CX(11, 12); CX(13, 14); CX(0, 8); CX(1, 9);
    CX(2, 10); CX(3, 11); CX(4, 12); CX(5, 13); CX(6, 14); CX(7, 15); CX(4, 8); CX(5, 9);
    CX(6, 10); CX(7, 11); CX(2, 4); CX(3, 5); CX(6, 8); CX(7, 9); CX(10, 12); CX(11, 13);
    CX(1, 2); CX(3, 4); CX(5, 6); CX(7, 8); CX(9, 10); CX(11, 12); CX(13, 14);
#undef CX
}

template <int NNPE>
__global__ void __launch_bounds__(128) k_adjacency_count(int64_t nnode, const int32_t* __restrict__ conn,
                                                         const int32_t* __restrict__ eptr,
                                                         const int32_t* __restrict__ einc, int32_t* __restrict__ ndeg,
                                                         int32_t* __restrict__ status) {
    __shared__ int s_list[kFastAdj * kAdjStride];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > nnode) return;
    if (n == nnode) {  // slot for the grand total after the scan
        ndeg[n] = 0;
        return;
    }
    const int k0 = eptr[n], k1 = eptr[n + 1];
    if constexpr (NNPE == 4) {
        if (k1 - k0 <= 4 && nnode < kPackedIdLimit) {        // regular node: sorting network in registers
            unsigned key[16];
            quad4_sorted_candidates(conn, einc, k0, k1 - k0, key);
            int c = key[0] != 0xFFFFFFFFu;
#pragma unroll
            for (int i = 1; i < 16; ++i) c += key[i] != 0xFFFFFFFFu && (key[i] >> 4) != (key[i - 1] >> 4);
            ndeg[n] = c;
            return;
        }
    }
    int cnt = build_adjacency_smem<NNPE>(conn, einc, k0, k1, s_list + threadIdx.x);
    if (cnt < 0) {                                   // rare: more than kFastAdj neighbours
        int list[kMaxAdj];
        cnt = build_adjacency(NNPE, conn, einc, k0, k1, list);
    }
    if (cnt < 0) {
        atomicMax(status, 1);
        cnt = 0;
    }
    ndeg[n] = cnt;
}

// Lane o = the last lane whose (ascending) start offset is <= q: the owner of flat position q of a warp's range.
__device__ __forceinline__ int owner_lane(int start, int q) {
    int o = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int c = o + step;
        if (__shfl_sync(0xffffffffu, start, c) <= q) o = c;
    }
    return o;
}

// nadj (the sorted neighbour list of every node) and epos (for every incidence (node, element) the position of the
// element's nodes in that node's list).  One thread per node builds the list in shared memory; the lists of a warp's
// 32 consecutive nodes form ONE contiguous range of nadj, and their incidences one contiguous range of epos, so both
// are written out flat by the whole warp (coalesced: a thread writing its own row would touch one 32-byte sector per
// 4-byte store).  A node that exceeds a staging capacity writes its own rows directly.
constexpr int kPosWords = 16;            // staged epos words per node: 16 incidences (Quad4 / Tri3), 5 (Quad9)
template <int NNPE>
__global__ void __launch_bounds__(128) k_adjacency_fill(int64_t nnode, const int32_t* __restrict__ conn,
                                                        const int32_t* __restrict__ eptr,
                                                        const int32_t* __restrict__ einc,
                                                        const int32_t* __restrict__ nptr, int32_t* __restrict__ nadj,
                                                        uint8_t* __restrict__ epos) {
    constexpr int W = (NNPE + 3) / 4;    // words of packed positions per incidence
    __shared__ int s_list[kFastAdj * kAdjStride];
    __shared__ unsigned s_pos[kPosWords * kAdjStride];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = n < nnode;
    const int64_t nc = valid ? n : nnode;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    const int k0 = eptr[nc], k1 = valid ? eptr[nc + 1] : k0;
    const int p0 = nptr[nc];
    int* s = s_list + threadIdx.x;
    int cnt = 0;
    bool pos_done = false;                           // positions already staged (the register fast path)
    bool fast = false;
    if constexpr (NNPE == 4) fast = valid && k1 - k0 <= 4 && nnode < kPackedIdLimit;
    if (fast) {
        if constexpr (NNPE == 4) {
            unsigned key[16];
            quad4_sorted_candidates(conn, einc, k0, k1 - k0, key);
            // candidate o = 4 q + b (node b of incidence q) sits at rank r of the list: byte b of the incidence's staged
            // position word -- stored straight into that word (s_pos[q * kAdjStride + t]) as a byte
            uint8_t* pos_bytes = reinterpret_cast<uint8_t*>(s_pos + threadIdx.x);
            int r = -1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bool live = key[i] != 0xFFFFFFFFu;
                const bool first = live && (i == 0 || (key[i] >> 4) != (key[i - 1] >> 4));
                r += first ? 1 : 0;
                if (first) s[r * kAdjStride] = (int)(key[i] >> 4);
                const unsigned o = key[i] & 15u;
                if (live) pos_bytes[(o >> 2) * (4u * kAdjStride) + (o & 3u)] = (uint8_t)r;
            }
            cnt = r + 1;
            pos_done = true;
        }
    } else if (valid) {
        cnt = build_adjacency_smem<NNPE>(conn, einc, k0, k1, s);
    }
    int staged = cnt;                                // entries of nadj this lane leaves to the flat write-out
    if (cnt < 0) {                                   // rare: the general thread-local path, rows written directly
        staged = 0;
        int list[kMaxAdj];
        cnt = build_adjacency(NNPE, conn, einc, k0, k1, list);
        if (cnt >= 0) {
            for (int i = 0; i < cnt; ++i) nadj[p0 + i] = list[i];
            for (int k = k0; k < k1; ++k) {
                const int64_t e = einc[k] >> 4;
                for (int b = 0; b < NNPE; ++b) {
                    const int v = conn[e * NNPE + b];
                    int lo = 0, hi = cnt - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (list[mid] < v) lo = mid + 1;
                        else hi = mid;
                    }
                    epos[(int64_t)k * NNPE + b] = (uint8_t)lo;
                }
            }
        }
        cnt = 0;                                     // nothing more to do for this node below
    }
    __syncwarp();
    {   // nadj: flat positions [p0 of lane 0, end of lane 31)
        const int qb = __shfl_sync(0xffffffffu, p0, 0), qe = __shfl_sync(0xffffffffu, p0 + (staged > 0 ? staged : 0), 31);
        const int qe_all = __shfl_sync(0xffffffffu, valid ? nptr[nc + 1] : p0, 31);
        (void)qe;
        const int c0 = __shfl_sync(0xffffffffu, staged, 0);
        if (c0 > 0 && __all_sync(0xffffffffu, staged == c0)) {
            // all 32 nodes have the same number of neighbours (the interior of a lattice): owner and entry by division
            for (int q0 = 0; q0 < 32 * c0; q0 += 32) {
                const int q = q0 + lane, o = q / c0;
                nadj[qb + q] = s_list[(q - o * c0) * kAdjStride + wbase + o];
            }
        } else {
            for (int q0 = qb; q0 < qe_all; q0 += 32) {
                const int q = q0 + lane;
                const int o = owner_lane(p0, q);
                const int i = q - __shfl_sync(0xffffffffu, p0, o);
                const int c = __shfl_sync(0xffffffffu, staged, o);
                if (q < qe_all && i < c) nadj[q] = s_list[i * kAdjStride + wbase + o];
            }
        }
    }
    // positions: packed bytes per incidence, staged (word w of incidence j at s_pos[(j * W + w) * kAdjStride + t])
    const int ninc = k1 - k0;
    const bool pos_staged = cnt > 0 && ninc * W <= kPosWords;
    for (int k = k0; k < k1 && cnt > 0 && !pos_done; ++k) {
        const int64_t e = einc[k] >> 4;
        unsigned packed[W];
#pragma unroll
        for (int w = 0; w < W; ++w) packed[w] = 0u;
#pragma unroll
        for (int b = 0; b < NNPE; ++b) {
            const int v = __ldg(conn + e * NNPE + b);
            int lo = 0, hi = cnt - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s[mid * kAdjStride] < v) lo = mid + 1;
                else hi = mid;
            }
            packed[b >> 2] |= (unsigned)lo << (8 * (b & 3));
        }
        if (pos_staged) {
#pragma unroll
            for (int w = 0; w < W; ++w) s_pos[((k - k0) * W + w) * kAdjStride + threadIdx.x] = packed[w];
        } else {
            uint8_t* dst = epos + (int64_t)k * NNPE;
            if constexpr (NNPE == 4) {
                *reinterpret_cast<unsigned*>(dst) = packed[0];             // 4-byte aligned: k * 4
            } else {
#pragma unroll
                for (int b = 0; b < NNPE; ++b) dst[b] = (uint8_t)(packed[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
    __syncwarp();
    {   // epos: flat units [k0 of lane 0, k1 of lane 31) x (one word per incidence for Quad4, else NNPE bytes)
        constexpr int U = NNPE == 4 ? 1 : NNPE;
        const int kb = __shfl_sync(0xffffffffu, k0, 0), ke = __shfl_sync(0xffffffffu, k1, 31);
        const int flag = pos_staged ? 1 : 0;
        const int n0 = __shfl_sync(0xffffffffu, ninc, 0);
        if (NNPE == 4 && n0 > 0 && __all_sync(0xffffffffu, ninc == n0 && pos_staged)) {
            // every node of the warp has n0 staged incidences: word u of the warp belongs to lane u / n0
            for (int u0 = 0; u0 < 32 * n0; u0 += 32) {
                const int u = u0 + lane, o = u / n0;
                reinterpret_cast<unsigned*>(epos)[(long long)kb + u] = s_pos[(u - o * n0) * kAdjStride + wbase + o];
            }
        } else
        for (long long u0 = (long long)kb * U; u0 < (long long)ke * U; u0 += 32) {
            const long long u = u0 + lane;
            const int k = (int)(u / U), b = (int)(u - (long long)k * U);
            const int o = owner_lane(k0, k);
            const int j = k - __shfl_sync(0xffffffffu, k0, o);
            const int ok = __shfl_sync(0xffffffffu, flag, o);
            if (u < (long long)ke * U && ok) {
                if constexpr (NNPE == 4) {
                    reinterpret_cast<unsigned*>(epos)[u] = s_pos[j * kAdjStride + wbase + o];
                } else {
                    const unsigned w = s_pos[(j * W + (b >> 2)) * kAdjStride + wbase + o];
                    epos[u] = (uint8_t)(w >> (8 * (b & 3)));
                }
            }
        }
    }
}

// One CTA expands kExpandNodes consecutive nodes: their rows form one contiguous range of col_idx, written
// flat (coalesced, 16-byte stores for ndof = 2); the owning node of an entry is found by binary search in the
// CTA's slice of nptr held in shared memory.  A node's ndof rows are ndof copies of the same column sequence.
constexpr int kExpandNodes = 512;
template <typename IdxT>
__global__ void __launch_bounds__(256) k_csr_expand(int64_t nnode, int ndof, const int32_t* __restrict__ nptr,
                                                    const int32_t* __restrict__ nadj, IdxT* __restrict__ row_ptr,
                                                    int32_t* __restrict__ col_idx) {
    __shared__ int sp[kExpandNodes + 1];
    const int64_t n0 = (int64_t)blockIdx.x * kExpandNodes;
    const int cnt = (int)((nnode - n0) < kExpandNodes ? (nnode - n0) : kExpandNodes);
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) sp[i] = nptr[n0 + i];
    __syncthreads();
    const int nd2 = ndof * ndof;
    for (int i = threadIdx.x; i < cnt * ndof; i += blockDim.x) {
        const int ln = i / ndof, c = i - ln * ndof;
        row_ptr[n0 * ndof + i] = (IdxT)((int64_t)nd2 * sp[ln] + (int64_t)c * ndof * (sp[ln + 1] - sp[ln]));
    }
    if (n0 + cnt == nnode && threadIdx.x == 0) row_ptr[nnode * ndof] = (IdxT)((int64_t)nd2 * sp[cnt]);
    const int q0 = sp[0], q1 = sp[cnt];
    auto node_of = [&](int q) {   // largest i with sp[i] <= q
        int lo = 0, hi = cnt;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sp[mid] <= q) lo = mid; else hi = mid;
        }
        return lo;
    };
    if (ndof == 2) {
        int4* out = reinterpret_cast<int4*>(col_idx);   // cudaMalloc'd base, segments start at multiples of 4
        for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
            const int i = node_of(q);
            const int p0 = sp[i], deg2 = 2 * (sp[i + 1] - p0);
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int kk = 4 * (q - p0) + j;
                if (kk >= deg2) kk -= deg2;
                v[j] = 2 * __ldg(nadj + p0 + (kk >> 1)) + (kk & 1);
            }
            out[q] = make_int4(v[0], v[1], v[2], v[3]);
        }
    } else {
        const int64_t g0 = (int64_t)nd2 * q0, g1 = (int64_t)nd2 * q1;
        for (int64_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
            const int i = node_of((int)(g / nd2));
            const int p0 = sp[i], deg = sp[i + 1] - p0;
            const int kk = (int)((g - (int64_t)nd2 * p0) % (ndof * deg));
            col_idx[g] = ndof * __ldg(nadj + p0 + kk / ndof) + kk % ndof;
        }
    }
}

}  // namespace fe

extern "C" {

int fe_incidence_count(int elem_type, int64_t nelem, int64_t nnode, const int32_t* conn, int32_t* eptr, void* scan_ws,
                       void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_incidence_count: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nelem >= 0 && nnode >= 0 && conn && eptr && scan_ws, "null pointer or negative size");
    FE_REQUIRE(nelem < ((int64_t)1 << 27) && nelem * nnpe * nnpe < ((int64_t)1 << 31),
               "element count exceeds the packed (element*16+a) / int32 slot range");
    cudaStream_t st = as_stream(stream);
    k_zero_i32<<<grid_for(nnode + 1, 256), 256, 0, st>>>(eptr, nnode + 1);
    FE_CHECK_LAUNCH("k_zero_i32");
    if (nelem > 0) {
        k_count_incidence<<<grid_for((nelem * nnpe + 3) / 4, 256), 256, 0, st>>>(nelem * nnpe, conn, eptr);
        FE_CHECK_LAUNCH("k_count_incidence");
    }
    return exclusive_scan_i32(eptr, eptr, nnode + 1, scan_ws, st);
}

int fe_incidence_fill(int elem_type, int64_t nelem, int64_t nnode, const int32_t* conn, const int32_t* eptr,
                      int32_t* cursor, int32_t* einc, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_incidence_fill: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nelem >= 0 && nnode >= 0 && conn && eptr && cursor && einc, "null pointer or negative size");
    if (nelem == 0 || nnode == 0) return FE_OK;
    cudaStream_t st = as_stream(stream);
    k_zero_i32<<<grid_for(nnode, 256), 256, 0, st>>>(cursor, nnode);
    FE_CHECK_LAUNCH("k_zero_i32");
    k_fill_incidence<<<grid_for((nelem * nnpe + 3) / 4, 256), 256, 0, st>>>(nelem * nnpe, nnpe, conn, eptr, cursor, einc);
    FE_CHECK_LAUNCH("k_fill_incidence");
    k_sort_incidence<<<grid_for(nnode, 256), 256, 0, st>>>(nnode, eptr, einc);
    FE_CHECK_LAUNCH("k_sort_incidence");
    return FE_OK;
}

int fe_adjacency_count(int elem_type, int64_t nnode, const int32_t* conn, const int32_t* eptr, const int32_t* einc,
                       int32_t* nptr, int32_t* status, void* scan_ws, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_adjacency_count: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nnode >= 0 && conn && eptr && einc && nptr && status && scan_ws, "null pointer or negative size");
    cudaStream_t st = as_stream(stream);
    k_zero_i32<<<1, 32, 0, st>>>(status, 1);
    FE_CHECK_LAUNCH("k_zero_i32");
    const unsigned grid = grid_for(nnode + 1, 128);
    switch (nnpe) {
        case 4: k_adjacency_count<4><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, status); break;
        case 9: k_adjacency_count<9><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, status); break;
        default: k_adjacency_count<3><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, status); break;
    }
    FE_CHECK_LAUNCH("k_adjacency_count");
    return exclusive_scan_i32(nptr, nptr, nnode + 1, scan_ws, st);
}

int fe_adjacency_fill(int elem_type, int64_t nnode, const int32_t* conn, const int32_t* eptr, const int32_t* einc,
                      const int32_t* nptr, int32_t* nadj, uint8_t* epos, void* stream) {
    using namespace fe;
    const int nnpe = nnpe_of(elem_type);
    if (!nnpe) {
        set_error("fe_adjacency_fill: Invalid elemType %d", elem_type);
        return FE_ERR_UNSUPPORTED;
    }
    FE_REQUIRE(nnode >= 0 && conn && eptr && einc && nptr && nadj && epos, "null pointer or negative size");
    if (nnode == 0) return FE_OK;
    const unsigned grid = grid_for(nnode, 128);
    cudaStream_t st = as_stream(stream);
    switch (nnpe) {
        case 4: k_adjacency_fill<4><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, nadj, epos); break;
        case 9: k_adjacency_fill<9><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, nadj, epos); break;
        default: k_adjacency_fill<3><<<grid, 128, 0, st>>>(nnode, conn, eptr, einc, nptr, nadj, epos); break;
    }
    FE_CHECK_LAUNCH("k_adjacency_fill");
    return FE_OK;
}

int fe_csr_expand(int ndof, int64_t nnode, const int32_t* nptr, const int32_t* nadj, int idx64, void* row_ptr,
                  int32_t* col_idx, void* stream) {
    using namespace fe;
    FE_REQUIRE(ndof >= 1 && ndof <= 3 && nnode >= 0 && nptr && nadj && row_ptr && col_idx,
               "null pointer, negative size or ndof out of range");
    FE_REQUIRE(ndof != 2 || (reinterpret_cast<uintptr_t>(col_idx) & 15) == 0, "col_idx must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const unsigned grid = nnode > 0 ? grid_for(nnode, kExpandNodes) : 1u;
    if (idx64)
        k_csr_expand<int64_t><<<grid, 256, 0, st>>>(nnode, ndof, nptr, nadj,
                                                                        static_cast<int64_t*>(row_ptr), col_idx);
    else
        k_csr_expand<int32_t><<<grid, 256, 0, st>>>(nnode, ndof, nptr, nadj,
                                                                        static_cast<int32_t*>(row_ptr), col_idx);
    FE_CHECK_LAUNCH("k_csr_expand");
    return FE_OK;
}

}  // extern "C"
