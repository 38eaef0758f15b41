// K3, warp-specialised form: the numeric assembly kernel for Quad4 / Tri3 meshes, and the two plan kernels that
// lay out its per-tile records.
//
// Same tiling, ownership and arithmetic as fe_tiles.cu (a tile = <= TE elements of one Morton block; every node,
// hence every CSR row block, is owned by exactly one tile; element matrices are evaluated once per tile into a
// shared-memory slab; every CSR value is the sum of the contributing slab blocks in ascending element id and is
// written exactly once) -- so the bits are those of fe_assemble / fe_assemble_tiled.  What differs is how a tile
// moves through the SM.  One persistent CTA per SM, three roles that never meet at a CTA-wide barrier:
//
//   producer warps   tasks = (tile, chunk of 32 slab rows), dealt round-robin over the producer warps ACROSS tiles:
//                    load the connectivity row (prefetched one task ahead) and the node coordinates, evaluate the
//                    upper-triangular block element matrix in fp64 registers, wait until the slab slot is free,
//                    store the row, arrive on slab_full[slot].
//   consumer warps   one thread per (owned node, half of its neighbour list): wait for the tile's records
//                    (rec_full) and slab (slab_full); for every neighbour position p sum the <= 2 contributing
//                    slab blocks (lanes = consecutive nodes: same block, consecutive odd-strided rows => no bank
//                    conflicts on regular meshes), the diagonal block over all incident elements, and write the
//                    2x2 block into the staging buffer AT ITS FINAL CSR OFFSET inside the tile's row range;
//                    release slab and records, fence.proxy.async, arrive on staged_full[slot].
//   DMA warp         cp.async.bulk (TMA, 1-D) of the next tiles' record blobs into the record ring
//                    (mbarrier complete_tx), and -- once a staging buffer is full -- one cp.async.bulk
//                    shared -> global per contiguous run of owned CSR rows (a tile of a row-major lattice: 8 runs
//                    of 4.6 KB instead of 2300 16-byte stores), then wait_group.read and stage_free[slot].
//
// Rings: 2 slab slots, 2 record slots, 2 staging slots; six mbarrier pairs; phases advance with the tile index.
//
// Per-tile plan (built once per mesh by fe_tile_layout + fe_tile_ws_build from the same owner / layout /
// halo arrays as the fe_tiles.cu plan):
//   tconn      int32 [ntiles][TE + HC][nnpe]   connectivity in tile order (own rows, then halo rows; -1 = unused)
//   tdesc      int32 x 4 per tile              {blob offset / 16, blob bytes, first run, run count}
//   blob       16-byte aligned, tile-local indices only (identical for all interior tiles of a regular mesh):
//                header   uint32 x 4           {owned nodes nn, nn rounded up to 32 (nnp), max degree, byte offset of locs}
//                node     uint32 x 2 [nnp]     {staging offset / 16 | deg << 16 | self << 23,  cnt | item offset << 4}
//                rec      uint32 [maxdeg][nnp] neighbour position p of node ln at [p * nnp + ln]:
//                                              src0 | src1 << 14, src = (slab offset of the block, in doubles) << 1 |
//                                              transposed; an absent source (and p = self) points at the zero block
//                locs     uint16 [items]       incidence items of the diagonal blocks: slab row << 4 | local node
//   truns      16 bytes per run                {CSR value offset / 2 (int64), staging offset / 16, length / 16}
#include "fe_tile_config.cuh"

namespace fe {

template <int N> struct WsLayout {
    static constexpr int kBlocks = N * (N + 1) / 2;
    static constexpr int kSS = (3 * kBlocks) | 1;                 // slab row stride in doubles (odd: conflict-free rows)
    static constexpr int kTE = tile_elems(N);
    static constexpr int kRows = kTE + halo_cap(N);               // slab rows = rows per tile of tconn
    static constexpr int kChunks = (kRows + 31) / 32;             // producer tasks per tile
    static constexpr int kZeroRow = kRows;
    // Task (tile k, c) evaluates slab rows of chunk (c + k) % kChunks: with as many producer warps as chunks a warp
    // would otherwise own the same chunk of every tile, and the last chunks of a tile (its halo rows) are partly or
    // wholly empty -- on a Quad4 lattice one warp in six would idle and two of the busy ones share a scheduler.
    __host__ __device__ static constexpr int rotated_chunk(int c, int k) { return (c + k) % kChunks; }
    static constexpr int kSlabDoubles = (((kRows + 1) * kSS) + 1) & ~1;
#ifndef FE_WS_PW
#define FE_WS_PW 6
#define FE_WS_CW 8
#endif
#ifndef FE_WS_DW
#define FE_WS_DW 2
#endif
    static constexpr int kPW = FE_WS_PW, kCW = FE_WS_CW, kDW = FE_WS_DW;   // producer / consumer / DMA warps
    // Register budget.  The element arithmetic needs ~125 registers, the consumers ~96, the DMA warps ~40: with one
    // allocation for all, 16 warps of 128 registers fill the file.  FE_WS_REGSPLIT launches warp GROUPS of four
    // (producers, consumers, DMA) with kRegLaunch registers each and lets every group set its own count
    // (setmaxnreg), which fits 8 + 8 + 4 warps.
#ifndef FE_WS_REGSPLIT
#define FE_WS_REGSPLIT 0
#endif
    static constexpr bool kRegSplit = FE_WS_REGSPLIT != 0;
    static constexpr int kDWPad = kRegSplit ? ((kDW + 3) & ~3) : kDW;             // DMA warps incl. idle ones of their group
    static constexpr int kThreads = 32 * (kPW + kCW + kDWPad);
    static constexpr int kRegP = 128, kRegC = 88, kRegD = 40;
#ifndef FE_WS_BATCH
#define FE_WS_BATCH 5
#endif
#ifndef FE_WS_PARTS
#define FE_WS_PARTS 2
#endif
#ifndef FE_WS_REC_SLOTS
#define FE_WS_REC_SLOTS 2
#endif
    static constexpr int kBatch = FE_WS_BATCH;                    // neighbour positions a consumer thread keeps in flight
    static constexpr int kParts = FE_WS_PARTS;                    // slices of a node's neighbour list (one consumer thread each)
#ifndef FE_WS_UNITS
#define FE_WS_UNITS 1
#endif
    static constexpr int kUnits = FE_WS_UNITS;                    // work units a consumer warp processes with interleaved loads
    static constexpr int kRecSlots = FE_WS_REC_SLOTS;             // record ring: blobs are requested this many tiles ahead
    static constexpr int kXYDepth = 2;                            // coordinate gathers in flight per producer warp (tasks)
    // record source field: (slab offset in doubles) << 1 | transposed; two of them per record
    static constexpr int kSrcBits = 14;
    static constexpr unsigned kSrcMask = (1u << kSrcBits) - 1u;
    static constexpr unsigned kZeroSrc = (unsigned)(kRows * kSS) << 1;            // the slab's zero block
    static constexpr unsigned kZeroRec = kZeroSrc | (kZeroSrc << kSrcBits);
    static constexpr int kBarBytes = 256;
#ifndef FE_WS_SLAB_SLOTS
#define FE_WS_SLAB_SLOTS 2
#endif
#ifndef FE_WS_STAGE_SLOTS
#define FE_WS_STAGE_SLOTS 2
#endif
    static constexpr int kSlabSlots = FE_WS_SLAB_SLOTS;           // slab ring (producers -> consumers), 2..4
    static constexpr int kStageSlots = FE_WS_STAGE_SLOTS;         // staging ring (consumers -> DMA warps), 2..4
#ifndef FE_WS_BLOB_CAP
#define FE_WS_BLOB_CAP (10 * 1024)
#endif
#ifndef FE_WS_ROW_RING
#define FE_WS_ROW_RING 4
#endif
    static constexpr int kBlobCap = FE_WS_BLOB_CAP;               // bytes of one record slot
    static constexpr int kRowRing = FE_WS_ROW_RING;               // connectivity rows staged per producer lane (3 or 4)
#ifndef FE_WS_STAGE_CAP
#define FE_WS_STAGE_CAP (42 * 1024)
#endif
    static constexpr int kStageCap = FE_WS_STAGE_CAP;             // bytes of one staging slot (CSR values of a tile)
    static constexpr int kCoordBytes = N * 32 * 16;               // per producer warp and task in flight: node coordinates
    static constexpr size_t kSlabBytes = (size_t)kSlabDoubles * sizeof(double);
    static constexpr int kProdBytes = kXYDepth * kCoordBytes + kRowRing * 32 * 16;   // + ring of connectivity rows per lane
    static constexpr size_t kSmem = kBarBytes + kSlabSlots * kSlabBytes + (size_t)kRecSlots * kBlobCap +
                                    kStageSlots * (size_t)kStageCap +
                                    (size_t)kPW * kProdBytes;
    static_assert((kRows + 1) * kSS < (1 << (kSrcBits - 1)), "slab offset does not fit a source field");
    static_assert(kDW == 1 || kDW == 2, "one DMA warp, or two that take alternate tiles (= one staging slot each)");
    static_assert(kRecSlots >= 2 && kRecSlots <= 4, "record ring: 2..4 slots");
    static_assert(kRowRing == 3 || kRowRing == 4, "rows of tasks i+2, i+3, i+4 are live at once");
    // A parity wait tells "the preceding phase" from "the current one", nothing older.  A producer warp waits on the
    // slot of a tile it still owes a chunk to, so that slot cannot run more than one phase ahead of it; the bound below
    // only keeps a warp's consecutive tasks within one turn of the slab ring.
    static_assert(kPW <= kSlabSlots * kChunks, "too many producer warps for the slab ring");
    static_assert(kRegSplit ? (kPW % 4 == 0 && kCW % 4 == 0 && 32 * (kPW * kRegP + kCW * kRegC + kDWPad * kRegD) <= 65536)
                            : (kPW + kCW + kDW <= 16),
                  "register file: 512 threads x 128 registers, or warp groups with their own counts");
    static_assert(kSmem <= 232448, "shared memory of one CTA");
};

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
#ifndef FE_WS_L2_HINTS
#define FE_WS_L2_HINTS 0      // developer builds: 1 = output stores evict-first in L2, 2 = record-blob loads too
#endif
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
#if FE_WS_L2_HINTS >= 2
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar), "l"(l2_evict_first())
                 : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
#if FE_WS_L2_HINTS >= 1
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_src),
                 "r"(bytes), "l"(l2_evict_first())
                 : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
                 : "memory");
#endif
}

#ifdef FE_WS_TRACE      // developer build: per-warp event time stamps (SM clock) of the first kTraceTiles tiles of CTA 0
constexpr int kTraceTiles = 512, kTraceEvents = 8;
__device__ long long g_ws_trace[16][kTraceTiles][kTraceEvents];
#define FE_TRACE(k, ev)                                                                          \
    do {                                                                                          \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (k) < kTraceTiles) g_ws_trace[threadIdx.x >> 5][(k)][(ev)] = clock64(); \
    } while (0)
#else
#define FE_TRACE(k, ev) ((void)0)
#endif

// =================================================================================================
// numeric kernel
// =================================================================================================
// SCALAR = FE_OP_POISSON (one dof per node, vals[nptr[n] + p] = the L entry of the node pair's block): same plan, same
// rings; producers evaluate and store only the L entries, consumers stage one 8-byte value per record at
// (staging offset / 2), and the DMA warps write the runs with coalesced 8-byte stores instead of bulk copies (the runs
// of 8-byte values are not 16-byte aligned, and the output is a quarter of the two-dof one).
template <int ET, bool MASS, bool SCALAR = false>
__global__ void __launch_bounds__(WsLayout<Elem<ET>::NNPE>::kThreads, 1) k_assemble_ws(
    int ntiles, const int32_t* __restrict__ tconn, const double2* __restrict__ coords, const int4* __restrict__ tdesc,
    const unsigned char* __restrict__ blobs, const int4* __restrict__ truns, double* __restrict__ vals, int weighted) {
    using E = Elem<ET>;
    constexpr int N = E::NNPE;
    using T = WsLayout<N>;
    constexpr int SS = T::kSS;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    double* slab0 = reinterpret_cast<double*>(smem + T::kBarBytes);
    unsigned char* rec0 = smem + T::kBarBytes + T::kSlabSlots * T::kSlabBytes;
    unsigned char* stage0 = rec0 + T::kRecSlots * T::kBlobCap;
    unsigned char* coord0 = stage0 + T::kStageSlots * T::kStageCap;
    // barrier ids (x2 slots each; the record ring has kRecSlots <= 4)
    enum { B_SLAB_FULL = 0, B_SLAB_EMPTY = 4, B_STAGED = 8, B_STAGE_FREE = 12, B_REC_FULL = 16, B_REC_EMPTY = 20 };
    static_assert(8 * (B_REC_EMPTY + 4) <= T::kBarBytes, "barrier block");
    constexpr int DS = T::kSlabSlots, DG = T::kStageSlots;
    static_assert(DS >= 2 && DS <= 4 && DG >= 2 && DG <= 4, "ring depths");
    constexpr int R = T::kRecSlots;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int id, int slot) { return bar0 + 8u * (uint32_t)(id + slot); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const int K = (ntiles - (int)blockIdx.x + G - 1) / G;       // tiles of this CTA: blockIdx.x + k * G
    if (tid == 0) {
        for (int s = 0; s < DS; ++s) {
            mbar_init(bar(B_SLAB_FULL, s), T::kChunks);
            mbar_init(bar(B_SLAB_EMPTY, s), T::kCW);
        }
        for (int s = 0; s < DG; ++s) {
            mbar_init(bar(B_STAGED, s), T::kCW);
            mbar_init(bar(B_STAGE_FREE, s), 1);
        }
        for (int s = 0; s < R; ++s) {
            mbar_init(bar(B_REC_FULL, s), 1);
            mbar_init(bar(B_REC_EMPTY, s), T::kCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3 * DS) slab0[(tid / 3) * T::kSlabDoubles + T::kZeroRow * SS + (tid % 3)] = 0.0;   // zero block of every slot
    __syncthreads();

    if constexpr (T::kRegSplit) {           // warp-group-uniform: every warp of a group of four runs the same branch
        if (warp < T::kPW) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T::kRegP));
        else if (warp < T::kPW + T::kCW) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T::kRegC));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T::kRegD));
        if (warp >= T::kPW + T::kCW + T::kDW) return;      // padding warps of the DMA group
    }
    if (warp < T::kPW) {
        // ================================ producers ================================
        // Task j of this CTA = (tile j / kChunks, chunk j % kChunks); K * kChunks < 2^31 (K <= ntiles / gridDim.x).
        const int total = K * T::kChunks;
        auto row_addr = [&](int j) -> const int32_t* {
            if (j >= total) return nullptr;
            const int k = j / T::kChunks, c = j - k * T::kChunks;
            const int row = T::rotated_chunk(c, k) * 32 + lane;
            if (row >= T::kRows) return nullptr;
            const long long t = (long long)blockIdx.x + (long long)k * G;
            return tconn + (t * T::kRows + row) * N;
        };
        // Software pipeline per warp.  DRAM latency under the output stream is about two task times, and nothing an
        // iteration loads may be read by the same iteration (that would make every task one DRAM latency long -- also
        // through a register-to-register rotation of prefetched rows).  So everything travels with cp.async through
        // this warp's staging blocks, one commit group per task: while task i is evaluated, the connectivity rows of
        // tasks i+3, i+4 and the node coordinates of tasks i+1, i+2 are in flight; wait_group 1 at the top of task i
        // guarantees the group of task i-2 (= coordinates of task i, row of task i+2).
        unsigned char* wbuf = coord0 + warp * T::kProdBytes;
        double2* cbuf = reinterpret_cast<double2*>(wbuf);                          // [2][a][lane]
        int4* rbuf = reinterpret_cast<int4*>(wbuf + 2 * T::kCoordBytes);           // [4][lane]
        auto request_row = [&](int j, int ring) {      // ring: slot index in [0, kRowRing)
            const int32_t* p = row_addr(j);
            int4* dst = rbuf + ring * 32 + lane;
#if defined(FE_WS_WHATIF) && (FE_WS_WHATIF & 8)     // developer timing build: producers load nothing
            p = nullptr;
#endif
            if (!p) {
                *dst = make_int4(-1, -1, -1, -1);
            } else if constexpr (N == 4) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(p) : "memory");
            } else {
#pragma unroll
                for (int a = 0; a < N; ++a)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst) + 4u * a), "l"(p + a) : "memory");
            }
        };
        auto request_xy = [&](const int4 t, int buf) {
            const int tg[4] = {t.x, t.y, t.z, t.w};
            if (t.x >= 0) {
#pragma unroll
                for (int a = 0; a < N; ++a)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(cbuf + (buf * N + a) * 32 + lane)),
                                 "l"(coords + tg[a])
                                 : "memory");
            }
        };
        auto direct_row = [&](int j) {          // prologue only: the first two rows are needed at once
            int4 t = make_int4(-1, -1, -1, -1);
            const int32_t* p = row_addr(j);
#if defined(FE_WS_WHATIF) && (FE_WS_WHATIF & 8)
            p = nullptr;
#endif
            if (p) {
                t.x = __ldg(p), t.y = __ldg(p + 1), t.z = __ldg(p + 2);
                if constexpr (N == 4) t.w = __ldg(p + 3);
            }
            return t;
        };
        bool act0, act1;                    // tasks i, i+1 have an element in this lane
        int j = warp;
        {
            const int4 r0 = direct_row(j), r1 = direct_row(j + T::kPW);
            request_row(j + 2 * T::kPW, 2 % T::kRowRing);
            request_row(j + 3 * T::kPW, 3 % T::kRowRing);
            request_xy(r0, 0);
            asm volatile("cp.async.commit_group;" ::: "memory");
            request_xy(r1, 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            act0 = r0.x >= 0;
            act1 = r1.x >= 0;
        }
        int k = 0, c = warp;                // tile / chunk of task j (kPW <= 2 * kChunks)
        while (c >= T::kChunks) c -= T::kChunks, ++k;
        int ring_rd = 2 % T::kRowRing, ring_wr = 4 % T::kRowRing;   // ring slots of the rows of tasks i+2 (read) and i+4 (written)
        for (int it = 0; j < total; j += T::kPW, ++it) {
            const int row = T::rotated_chunk(c, k) * 32 + lane;
            const int slot = k % DS, buf = it & 1;
            FE_TRACE(k, 0);
            request_row(j + 4 * T::kPW, ring_wr);
            const bool active = act0;
            asm volatile("cp.async.wait_group 1;" ::: "memory");        // coordinates of this task, row of task i+2
            FE_TRACE(k, 1);
            double x[N], y[N];
            if (active) {
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    const double2 p = cbuf[(buf * N + a) * 32 + lane];
                    x[a] = p.x;
                    y[a] = p.y;
                }
            }
            const int4 ahead = rbuf[ring_rd * 32 + lane];
            ring_rd = ring_rd + 1 == T::kRowRing ? 0 : ring_rd + 1;
            ring_wr = ring_wr + 1 == T::kRowRing ? 0 : ring_wr + 1;
            request_xy(ahead, buf);                                     // task i+2: in flight for two task times
            asm volatile("cp.async.commit_group;" ::: "memory");
            double ke[3 * T::kBlocks];
            if (active) {
#if defined(FE_WS_WHATIF) && (FE_WS_WHATIF & 1)     // developer timing build: no element arithmetic (wrong results)
#pragma unroll
                for (int i = 0; i < 3 * T::kBlocks; ++i) ke[i] = x[i & (N - 1)];
#else
                element_tri<ET, MASS, SCALAR>(x, y, weighted != 0, ke);
#endif
            }
            FE_TRACE(k, 2);
            mbar_wait(bar(B_SLAB_EMPTY, slot), ((k / DS) & 1) ^ 1);     // consumers are done with tile k - DS
            FE_TRACE(k, 3);
            if (active) {
                double* dst = slab0 + slot * T::kSlabDoubles + row * SS;
#pragma unroll
                for (int i = 0; i < 3 * T::kBlocks; ++i)
                    if (!SCALAR || i % 3 == 0) dst[i] = ke[i];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_SLAB_FULL, slot));
            FE_TRACE(k, 4);
            act0 = act1;
            act1 = ahead.x >= 0;
            c += T::kPW;
            while (c >= T::kChunks) c -= T::kChunks, ++k;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp < T::kPW + T::kCW) {
        // ================================ consumers ================================
        const int cw = warp - T::kPW;
        int rslot = 0;
        uint32_t rpar = 0;
        for (int k = 0; k < K; ++k) {
            const int slot = k % DS, gslot = k % DG;
            const uint32_t par = (k / DS) & 1, gpar = (k / DG) & 1;
            const unsigned char* rec = rec0 + rslot * T::kBlobCap;
            const double* slab = slab0 + slot * T::kSlabDoubles;
            double* stage = reinterpret_cast<double*>(stage0 + gslot * T::kStageCap);
            FE_TRACE(k, 0);
            mbar_wait(bar(B_REC_FULL, rslot), rpar);
            FE_TRACE(k, 1);
            const uint4 hdr = *reinterpret_cast<const uint4*>(rec);
            const int nn = (int)hdr.x, nnp = (int)hdr.y;
            const uint2* node = reinterpret_cast<const uint2*>(rec + 16);
            const uint32_t* recs = reinterpret_cast<const uint32_t*>(rec + 16 + 8 * nnp);
            const uint16_t* locs = reinterpret_cast<const uint16_t*>(rec + hdr.w);
            // Work units = (group of 32 owned nodes, one of kParts slices of the neighbour lists), dealt round-robin over
            // the consumer warps.  Halves: a Quad4 lattice tile has 5 groups = 10 units on 8 warps.  Thirds would fill
            // both rounds (15 units) but were measured slower (1.39 against 1.30 ms): every unit re-reads the node
            // descriptors and the kernel is bound by the instructions all warps issue, not by its longest warp.
            constexpr int parts = T::kParts;
            const int groups = nnp >> 5;
            mbar_wait(bar(B_SLAB_FULL, slot), par);
            FE_TRACE(k, 2);
            // The staging slot is still being read by the bulk stores of tile k - 2 for most of this tile's time.  So a
            // thread first sums everything it owns into registers (slab reads + adds: the long part) and only then waits
            // for the slot: the slot is busy for the few store instructions, not for the whole consumer pass.
            bool slot_free = false;
            // A warp's units are processed kUnits at a time with their loads interleaved (records of all, then slab
            // blocks of all, then the stores).  Measured on cfg2: 1 unit 1.31 ms, 2 units 1.52 ms, 3 thirds 1.65 ms -- a
            // consumer pass is not bound by the latency of its dependent chain, and the wider code costs registers.
            constexpr int NU = T::kUnits;
            const int nunits = groups * parts;
            for (int u0 = cw; u0 < nunits; u0 += NU * T::kCW) {
                int ln[NU], deg[NU], self[NU], p0[NU], p1[NU];
                uint2 nd[NU];
                bool diag[NU];
                double* row0[NU];
                double Ld[NU], Pd[NU];
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    const int u = u0 + j * T::kCW;
                    const int g = u / parts, part = u - g * parts;
                    ln[j] = g * 32 + lane;
#if defined(FE_WS_WHATIF) && (FE_WS_WHATIF & 2)     // developer timing build: consumers only hand the buffers on
                    const bool live = false;
#else
                    const bool live = u < nunits && ln[j] < nn;
#endif
                    nd[j] = live ? node[ln[j]] : make_uint2(0u, 0u);
                    deg[j] = (nd[j].x >> 16) & 127, self[j] = (nd[j].x >> 23) & 127;
                    p0[j] = (deg[j] * part) / parts, p1[j] = (deg[j] * (part + 1)) / parts;
                    row0[j] = stage + (SCALAR ? (nd[j].x & 0xFFFFu) >> 1 : 2 * (nd[j].x & 0xFFFFu));
                    diag[j] = live && self[j] >= p0[j] && self[j] < p1[j];
                    Ld[j] = Pd[j] = 0.0;
                }
                // A node's row 2n+1 lies 16 * deg bytes behind its row 2n: for odd degrees (9 on a Quad4 lattice, 7 on a
                // Tri3 lattice) the two rows fall into different 16-byte bank groups, so letting every other group of
                // four lanes write row 2n+1 first makes each quarter-warp store hit eight distinct groups.
                const bool flip = !SCALAR && ((lane >> 2) & 1) != 0;
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    if (diag[j]) {                                      // diagonal block: all incident elements
                        const int cnt = nd[j].y & 15, ioff = nd[j].y >> 4;
                        for (int q = 0; q < cnt; ++q) {
                            const unsigned loc = locs[ioff + q];
                            const int a = loc & 15;
                            const double* b = slab + (loc >> 4) * SS + 3 * (a * N - ((a * (a - 1)) >> 1));
                            Ld[j] += b[0];
                            if constexpr (!SCALAR) Pd[j] += b[1];
                        }
                    }
                }
                // kBatch neighbour positions in flight per thread and unit: all records, then all slab loads, then the stores
                for (int it = 0;; ++it) {
                    bool more = !slot_free;
#pragma unroll
                    for (int j = 0; j < NU; ++j) more = more || p0[j] + it * T::kBatch < p1[j];
                    if (!more) break;
                    uint32_t r[NU][T::kBatch];
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        const int pb = p0[j] + it * T::kBatch;
#pragma unroll
                        for (int q = 0; q < T::kBatch; ++q) r[j][q] = pb + q < p1[j] ? recs[(pb + q) * nnp + ln[j]] : T::kZeroRec;
                    }
                    double L[NU][T::kBatch], P[NU][T::kBatch], Q[NU][T::kBatch];
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
#pragma unroll
                        for (int q = 0; q < T::kBatch; ++q) {
                            const uint32_t s0 = r[j][q] & T::kSrcMask, s1 = r[j][q] >> T::kSrcBits;
                            const double* b0 = slab + (s0 >> 1);
                            const double* b1 = slab + (s1 >> 1);            // the zero block if absent
                            const int f0 = s0 & 1, f1 = s1 & 1;             // transposed block: P and Q trade places
                            L[j][q] = b0[0] + b1[0];
                            if constexpr (!SCALAR) {
                                P[j][q] = b0[1 + f0] + b1[1 + f1];
                                Q[j][q] = b0[2 - f0] + b1[2 - f1];
                            } else {
                                P[j][q] = Q[j][q] = 0.0;
                            }
                        }
                    }
                    if (!slot_free) {                                   // warp-uniform: first store of this warp into the slot
                        FE_TRACE(k, 3);
                        mbar_wait(bar(B_STAGE_FREE, gslot), gpar ^ 1);  // bulk stores of tile k - DG have read the buffer
                        FE_TRACE(k, 4);
                        slot_free = true;
                    }
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        const int pb = p0[j] + it * T::kBatch;
                        double* row1 = row0[j] + 2 * deg[j];
                        double* rowA = flip ? row1 : row0[j];
                        double* rowB = flip ? row0[j] : row1;
#pragma unroll
                        for (int q = 0; q < T::kBatch; ++q) {
                            const int p = pb + q;
                            if (p < p1[j] && p != self[j]) {
                                if constexpr (SCALAR) {
                                    row0[j][p] = L[j][q];
                                } else {
                                    *reinterpret_cast<double2*>(rowA + 2 * p) =
                                        flip ? make_double2(Q[j][q], L[j][q]) : make_double2(L[j][q], P[j][q]);
                                    *reinterpret_cast<double2*>(rowB + 2 * p) =
                                        flip ? make_double2(L[j][q], P[j][q]) : make_double2(Q[j][q], L[j][q]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    if (diag[j]) {
                        if constexpr (SCALAR) {
                            row0[j][self[j]] = Ld[j];
                        } else {
                            double* row1 = row0[j] + 2 * deg[j];
                            double* rowA = flip ? row1 : row0[j];
                            double* rowB = flip ? row0[j] : row1;
                            *reinterpret_cast<double2*>(rowA + 2 * self[j]) = flip ? make_double2(Pd[j], Ld[j]) : make_double2(Ld[j], Pd[j]);
                            *reinterpret_cast<double2*>(rowB + 2 * self[j]) = flip ? make_double2(Ld[j], Pd[j]) : make_double2(Pd[j], Ld[j]);
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(B_SLAB_EMPTY, slot));
                mbar_arrive(bar(B_REC_EMPTY, rslot));
                mbar_arrive(bar(B_STAGED, gslot));
            }
            FE_TRACE(k, 5);
            if (++rslot == R) rslot = 0, rpar ^= 1u;
        }
    } else {
        // ================================ DMA warps ================================
        auto tile_of = [&](int k) { return (long long)blockIdx.x + (long long)k * G; };
        auto desc_of = [&](int k) { return k < K ? __ldg(tdesc + tile_of(k)) : make_int4(0, 0, 0, 0); };
        auto issue_rec = [&](int k, const int4 d) {     // lane 0 only
            const int rs = k % R;
            mbar_arrive_expect_tx(bar(B_REC_FULL, rs), (uint32_t)d.y);
            bulk_load(smem_u32(rec0 + rs * T::kBlobCap), blobs + 16ll * (long long)(unsigned)d.x, (uint32_t)d.y,
                      bar(B_REC_FULL, rs));
        };
        // One DMA warp walks all tiles; two DMA warps take the even / odd tiles (= one staging slot each), so the
        // bulk stores of tile k + 1 are issued while those of tile k are still draining.  What limits the output stream
        // is how many staged bytes are in flight (measured, stores alone: 1.39 ms with one tile, 1.02 ms with two).
        // Record blobs are requested kRecSlots tiles ahead: the slot of tile k is refilled with tile k + kRecSlots as
        // soon as the consumers have released it.
        const int dw = warp - (T::kPW + T::kCW);
        for (int k = dw; k < R && k < K; k += T::kDW) {
            const int4 d = desc_of(k);
            if (lane == 0) issue_rec(k, d);
        }
        int4 d_cur = desc_of(dw), d_ahead = desc_of(dw + R);
        for (int k = dw; k < K; k += T::kDW) {
            const int slot = k % DG;
            const uint32_t par = (k / DG) & 1;
            const int4 dn_cur = desc_of(k + T::kDW), dn_ahead = desc_of(k + T::kDW + R);   // next iteration's descriptors
            const int r0 = d_cur.z, nr = d_cur.w;
            int4 run = make_int4(0, 0, 0, 0);
            if (lane < nr) run = __ldg(truns + r0 + lane);               // in flight while the consumers finish the tile
            FE_TRACE(k, 0);
            mbar_wait(bar(B_STAGED, slot), par);
            FE_TRACE(k, 1);
            if (lane == 0 && k + R < K) {
                mbar_wait(bar(B_REC_EMPTY, k % R), (uint32_t)((k / R) & 1));   // (already complete: consumers arrive on it first)
                issue_rec(k + R, d_ahead);
            }
            if constexpr (SCALAR) {
                // one-dof values: the warp walks the runs and writes each with coalesced 8-byte streaming stores
                const double* sst = reinterpret_cast<const double*>(stage0 + slot * T::kStageCap);
                for (int i = 0; i < nr; ++i) {
                    int4 r;                              // runs 0..31 were fetched ahead of the wait, one per lane
                    if (i < 32) {
                        r.x = __shfl_sync(0xffffffffu, run.x, i), r.y = __shfl_sync(0xffffffffu, run.y, i);
                        r.z = __shfl_sync(0xffffffffu, run.z, i), r.w = __shfl_sync(0xffffffffu, run.w, i);
                    } else {
                        r = __ldg(truns + r0 + i);
                    }
                    const long long g0 = (long long)(((unsigned long long)(unsigned)r.y << 32) | (unsigned)r.x) >> 1;   // nptr[first]
                    const double* src = sst + (r.z >> 1);
                    const int len = r.w >> 1;
                    for (int e = lane; e < len; e += 32) __stcs(vals + g0 + e, src[e]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_STAGE_FREE, slot));
                d_cur = dn_cur;
                d_ahead = dn_ahead;
                continue;
            }
            const uint32_t sbase = smem_u32(stage0 + slot * T::kStageCap);
            // one bulk store per run: fewer, larger copies drain faster (~35 ns per copy + 34 ps per byte measured)
            for (int i = lane; i < nr; i += 32) {
                if (i >= 32) run = __ldg(truns + r0 + i);
                const long long off2 = (long long)(((unsigned long long)(unsigned)run.y << 32) | (unsigned)run.x);
#if !(defined(FE_WS_WHATIF) && (FE_WS_WHATIF & 4))  // developer timing build: no output stores
                bulk_store(vals + 2 * off2, sbase + 16u * (uint32_t)run.z, 16u * (uint32_t)run.w);
#else
                (void)off2;
                (void)sbase;
#endif
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            FE_TRACE(k, 2);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            FE_TRACE(k, 3);
            if (lane == 0) mbar_arrive(bar(B_STAGE_FREE, slot));
            d_cur = dn_cur;
            d_ahead = dn_ahead;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int ET, bool MASS, bool SCALAR = false>
static int launch_ws(int64_t ntiles, const int32_t* tconn, const double* coords, const int32_t* tdesc, const void* blobs,
                     const void* truns, double* vals, int weighted, cudaStream_t st) {
    using T = WsLayout<Elem<ET>::NNPE>;
    int dev = 0, sms = 0;
    FE_CUDA_TRY(cudaGetDevice(&dev));
    FE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t grid = sms;
    if (grid > ntiles) grid = ntiles;
    FE_CUDA_TRY(cudaFuncSetAttribute(k_assemble_ws<ET, MASS, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::kSmem));
    k_assemble_ws<ET, MASS, SCALAR><<<(unsigned)grid, T::kThreads, T::kSmem, st>>>(
        (int)ntiles, tconn, reinterpret_cast<const double2*>(coords), reinterpret_cast<const int4*>(tdesc),
        static_cast<const unsigned char*>(blobs), static_cast<const int4*>(truns), vals, weighted);
    FE_CHECK_LAUNCH("k_assemble_ws");
    return FE_OK;
}

// =================================================================================================
// plan kernels
// =================================================================================================
// One warp per tile: blob header, the padding entries [nn, nnp) of the node / record arrays, the run list, tdesc.
__global__ void __launch_bounds__(128) k_tile_ws_runs(int64_t ntiles, const int32_t* __restrict__ tile_nptr,
                                                      const int32_t* __restrict__ tile_iptr,
                                                      const int32_t* __restrict__ tile_nodes,
                                                      const int32_t* __restrict__ nptr, const int32_t* __restrict__ noff,
                                                      const int32_t* __restrict__ blob_off,
                                                      const int32_t* __restrict__ run_off,
                                                      unsigned char* __restrict__ blobs, int4* __restrict__ truns,
                                                      int4* __restrict__ tdesc, int zero_src) {
    __shared__ int s_nodes[4][256];      // the tile's (sorted) owned nodes: the run walk below reads them many times
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    int* sn = s_nodes[threadIdx.x >> 5];
    const int lo = tile_nptr[t], nn = tile_nptr[t + 1] - lo;
    const int ni = tile_iptr[t + 1] - tile_iptr[t];
    const int nnp = (nn + 31) & ~31;
    const bool cached = nn <= 256;       // (larger tiles exceed every capacity: the host discards the plan)
    int maxdeg = 0;
    for (int i = lane; i < nn; i += 32) {
        const int n = tile_nodes[lo + i];
        if (cached) sn[i] = n;
        const int deg = nptr[n + 1] - nptr[n];
        maxdeg = deg > maxdeg ? deg : maxdeg;
    }
    __syncwarp();
    auto node_at = [&](int i) { return cached ? sn[i] : tile_nodes[lo + i]; };
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int m = __shfl_xor_sync(0xffffffffu, maxdeg, d);
        maxdeg = m > maxdeg ? m : maxdeg;
    }
    unsigned char* blob = blobs + 16ll * blob_off[t];
    const int bytes = 16 * (blob_off[t + 1] - blob_off[t]);
    const int loc_byte = 16 + 8 * nnp + 4 * maxdeg * nnp;
    if (lane == 0) {
        *reinterpret_cast<uint4*>(blob) = make_uint4((unsigned)nn, (unsigned)nnp, (unsigned)maxdeg, (unsigned)loc_byte);
        tdesc[t] = make_int4(blob_off[t], bytes, run_off[t], run_off[t + 1] - run_off[t]);
    }
    uint2* node = reinterpret_cast<uint2*>(blob + 16);
    uint32_t* recs = reinterpret_cast<uint32_t*>(blob + 16 + 8 * nnp);
    const uint32_t zrec = (uint32_t)zero_src | ((uint32_t)zero_src << 14);
    for (int ln = nn + lane; ln < nnp; ln += 32) {
        node[ln] = make_uint2(0u, 0u);
        for (int p = 0; p < maxdeg; ++p) recs[p * nnp + ln] = zrec;
    }
    for (int b = loc_byte + 2 * ni + lane; b < bytes; b += 32) blob[b] = 0;      // tail padding
    if (lane == 0 && ni > 0) reinterpret_cast<uint16_t*>(blob + loc_byte)[ni - 1] = 0;   // item counts are padded to even
    // runs: a run starts at slot i when the previous owned node is not n - 1
    int base = run_off[t];
    for (int i0 = 0; i0 < nn; i0 += 32) {
        const int i = i0 + lane;
        int n = -2, start = 0;
        if (i < nn) {
            n = node_at(i);
            start = (i == 0 || node_at(i - 1) != n - 1) ? 1 : 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, start != 0);
        if (start) {
            // length: up to the next run start (or the end of the tile)
            int j = i + 1;
            while (j < nn && node_at(j) == node_at(j - 1) + 1) ++j;
            const int last = node_at(j - 1);
            const long long v2 = 2ll * nptr[n];                                   // value offset / 2 (4 values per node-level nz)
            const int len16 = 2 * (nptr[last + 1] - nptr[n]);                     // 32 bytes per node-level non-zero
            const int st16 = 2 * noff[2 * (int64_t)(lo + i) + 1];                 // staging offset / 16 = 2 * work offset
            truns[base + __popc(m & ((1u << lane) - 1u))] =
                make_int4((int)(v2 & 0xffffffffll), (int)(v2 >> 32), st16, len16);
        }
        base += __popc(m);
    }
}

// One thread per tile-node slot: node descriptor, records (one per neighbour position), item locators.
constexpr int kWsMaxDeg = 32;           // neighbours per node the plan of the warp-specialised kernel represents
__global__ void __launch_bounds__(128) k_tile_ws_records(
    int64_t nslots, int nnpe, const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ owner,
    const int32_t* __restrict__ tile_nptr, const int32_t* __restrict__ tile_hptr, const int32_t* __restrict__ noff,
    const int32_t* __restrict__ thalo, const int32_t* __restrict__ eptr, const int32_t* __restrict__ einc,
    const uint8_t* __restrict__ epos, const int32_t* __restrict__ einv, const int32_t* __restrict__ nptr,
    const int32_t* __restrict__ nadj, const int32_t* __restrict__ blob_off, unsigned char* __restrict__ blobs,
    int32_t* __restrict__ stats, int te, int he, int ss) {
    __shared__ uint32_t s_rec[kWsMaxDeg * 128];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const int n = tile_nodes[i];
    const int t = owner[n];
    const int lo = tile_nptr[t], nn = tile_nptr[t + 1] - lo;
    const int nnp = (nn + 31) & ~31;
    const int ln = (int)(i - lo);
    const int ioff = noff[2 * i], woff = noff[2 * i + 1];
    const int k0 = eptr[n], cnt = eptr[n + 1] - k0;
    const int p0 = nptr[n], deg = nptr[n + 1] - p0;
    const int nh = tile_hptr[t + 1] - tile_hptr[t];
    const int32_t* hl = thalo + (int64_t)t * he;
    unsigned char* blob = blobs + 16ll * blob_off[t];
    const uint4 hdr = *reinterpret_cast<const uint4*>(blob);        // written by k_tile_ws_runs (earlier launch)
    const int maxdeg = (int)hdr.z;
    // position of n in its own neighbour list = its slot as seen from any incident element (owned nodes have one)
    const int self = cnt > 0 ? (int)epos[(int64_t)k0 * nnpe + (einc[k0] & 15)] : 0;
    (void)nadj;
    const int soff = 2 * woff;                                                    // staging offset / 16 inside the tile's slot
    if (cnt > 15 || deg > kWsMaxDeg || soff > 0xFFFF || ioff >= (1 << 28)) {   // not representable
        atomicMax(stats + 4, 1);
        return;
    }
    uint2* node = reinterpret_cast<uint2*>(blob + 16);
    uint32_t* recs = reinterpret_cast<uint32_t*>(blob + 16 + 8 * nnp);
    uint16_t* locs = reinterpret_cast<uint16_t*>(blob + hdr.w);
    node[ln] = make_uint2((unsigned)soff | ((unsigned)deg << 16) | ((unsigned)self << 23),
                          (unsigned)cnt | ((unsigned)ioff << 4));
    const unsigned kZeroSrc = (unsigned)((te + he) * ss) << 1;
    const unsigned kZeroRec = kZeroSrc | (kZeroSrc << 14);
    // the node's records are collected in shared memory (entry p of thread t at s_rec[p * 128 + t]: conflict free;
    // a thread-local array would live in local memory)
    uint32_t* src = s_rec + threadIdx.x;
    for (int p = 0; p < deg; ++p) src[p * 128] = kZeroRec;
    // Four incidences at a time: their element ids, (tile, row) words, position bytes and halo-list searches are
    // independent, so each step has four loads in flight instead of one (the kernel is bound by that latency).
    for (int jb = 0; jb < cnt; jb += 4) {
        int item[4], pos[4], lo4[4], hi4[4];
        unsigned pb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) item[i] = jb + i < cnt ? einc[k0 + jb + i] : -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pos[i] = item[i] >= 0 ? einv[item[i] >> 4] : (t << 8);
            pb[i] = 0u;
            if (item[i] >= 0) {
                const uint8_t* ep = epos + (int64_t)(k0 + jb + i) * nnpe;
                if (nnpe == 4) {
                    pb[i] = *reinterpret_cast<const unsigned*>(ep);            // 4-byte aligned: 4 * incidence
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (b < nnpe) pb[i] |= (unsigned)ep[b] << (8 * b);
                }
            }
        }
        bool searching = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // foreign element: its position inside the tile's sorted halo list
            lo4[i] = 0;
            hi4[i] = (pos[i] >> 8) == t ? 0 : nh - 1;
            searching = searching || hi4[i] > 0;
        }
        while (searching) {
            searching = false;
            int v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = lo4[i] < hi4[i] ? hl[(lo4[i] + hi4[i]) >> 1] : 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (lo4[i] < hi4[i]) {
                    const int mid = (lo4[i] + hi4[i]) >> 1;
                    if (v[i] < (item[i] >> 4)) lo4[i] = mid + 1;
                    else hi4[i] = mid;
                    searching = searching || lo4[i] < hi4[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (item[i] < 0) continue;
            const int a = item[i] & 15;
            const int row = (pos[i] >> 8) == t ? (pos[i] & 255) : te + lo4[i];
            locs[ioff + jb + i] = (uint16_t)((row << 4) | a);
            for (int b = 0; b < nnpe; ++b) {
                const int p = (int)((pb[i] >> (8 * b)) & 255u);
                if (p == self) continue;
                const int blo = a < b ? a : b, bhi = a < b ? b : a;
                const int tri = blo * nnpe - blo * (blo - 1) / 2 + (bhi - blo);
                const unsigned sv = (unsigned)(((row * ss + 3 * tri) << 1) | (a > b ? 1 : 0));
                const uint32_t cur = src[p * 128];
                if ((cur & 0x3FFFu) == kZeroSrc) src[p * 128] = (cur & ~0x3FFFu) | sv;
                else if ((cur >> 14) == kZeroSrc) src[p * 128] = (cur & 0x3FFFu) | (sv << 14);
                else atomicMax(stats + 4, 1);  // a node pair shared by more than two elements
            }
        }
    }
    for (int p = 0; p < maxdeg; ++p) recs[p * nnp + ln] = p < deg ? src[p * 128] : kZeroRec;
}

}  // namespace fe

extern "C" {

#ifdef FE_WS_TRACE
int fe_ws_trace_read(void* host, size_t bytes) {          // developer build only
    using namespace fe;
    FE_REQUIRE(bytes <= sizeof(g_ws_trace), "too many bytes");
    FE_CUDA_TRY(cudaMemcpyFromSymbol(host, g_ws_trace, bytes));
    return FE_OK;
}
#endif

void fe_tile_ws_params(int elem_type, int32_t out[4]) {
    using namespace fe;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (elem_type == FE_QUAD4) {
        out[0] = WsLayout<4>::kBlobCap, out[1] = WsLayout<4>::kStageCap, out[2] = kWsMaxDeg, out[3] = 256;
    } else if (elem_type == FE_TRI3) {
        out[0] = WsLayout<3>::kBlobCap, out[1] = WsLayout<3>::kStageCap, out[2] = kWsMaxDeg, out[3] = 256;
    }
}

int fe_tile_ws_build(int elem_type, int64_t nslots, int64_t ntiles, const int32_t* tile_nodes, const int32_t* owner,
                     const int32_t* tile_ptr, const int32_t* noff, const int32_t* thalo, const int32_t* eptr,
                     const int32_t* einc, const uint8_t* epos, const int32_t* eloc, const int32_t* nptr,
                     const int32_t* nadj, const int32_t* blob_off, const int32_t* run_off, void* blobs, void* truns,
                     int32_t* tdesc, int32_t* stats, void* stream) {
    using namespace fe;
    FE_REQUIRE(elem_type == FE_QUAD4 || elem_type == FE_TRI3, "the warp-specialised kernel handles Quad4 / Tri3");
    FE_REQUIRE(nslots >= 0 && ntiles >= 0, "negative size");
    if (ntiles == 0) return FE_OK;
    FE_REQUIRE(tile_nodes && owner && tile_ptr && noff && thalo && eptr && einc && epos && eloc && nptr && nadj &&
                   blob_off && run_off && blobs && truns && tdesc && stats,
               "null pointer");
    const int nnpe = nnpe_of(elem_type);
    cudaStream_t st = as_stream(stream);
    const int64_t s = ntiles + 1;
    const int te = tile_elems(nnpe), he = halo_cap(nnpe);
    const int ss = nnpe == 4 ? WsLayout<4>::kSS : WsLayout<3>::kSS;
    k_tile_ws_runs<<<grid_for(ntiles, 4), 128, 0, st>>>(ntiles, tile_ptr, tile_ptr + s, tile_nodes, nptr, noff, blob_off,
                                                        run_off, static_cast<unsigned char*>(blobs),
                                                        static_cast<int4*>(truns), reinterpret_cast<int4*>(tdesc),
                                                        ((te + he) * ss) << 1);
    FE_CHECK_LAUNCH("k_tile_ws_runs");
    if (nslots > 0) {
        k_tile_ws_records<<<grid_for(nslots, 128), 128, 0, st>>>(
            nslots, nnpe, tile_nodes, owner, tile_ptr, tile_ptr + 2 * s, noff, thalo, eptr, einc, epos, eloc, nptr, nadj,
            blob_off, static_cast<unsigned char*>(blobs), stats, te, he, ss);
        FE_CHECK_LAUNCH("k_tile_ws_records");
    }
    return FE_OK;
}

int fe_assemble_ws(int elem_type, int op, int64_t ntiles, const int32_t* tconn, const double* coords, const int32_t* tdesc,
                   const void* blobs, const void* truns, double* vals, void* stream) {
    using namespace fe;
    FE_REQUIRE(ntiles >= 0 && ntiles < ((int64_t)1 << 31), "bad tile count");
    if (op != FE_OP_REFERENCE && op != FE_OP_WEIGHTED && op != FE_OP_MASS && op != FE_OP_POISSON) {
        set_error("fe_assemble_ws: unsupported operator %d", op);
        return FE_ERR_UNSUPPORTED;
    }
    if (ntiles == 0) return FE_OK;
    FE_REQUIRE(tconn && coords && tdesc && blobs && truns && vals, "null pointer");
    FE_REQUIRE((reinterpret_cast<uintptr_t>(vals) & 15) == 0 && (reinterpret_cast<uintptr_t>(blobs) & 15) == 0,
               "vals and blobs must be 16-byte aligned (bulk copies)");
    cudaStream_t st = as_stream(stream);
    int rc = ensure_tables_tu(st);
    if (rc != FE_OK) return rc;
    const int w = op == FE_OP_WEIGHTED;
    if (op == FE_OP_POISSON) {      // one dof per node: vals[nptr[n] + p]
        if (elem_type == FE_QUAD4) return launch_ws<FE_QUAD4, false, true>(ntiles, tconn, coords, tdesc, blobs, truns, vals, 0, st);
        if (elem_type == FE_TRI3) return launch_ws<FE_TRI3, false, true>(ntiles, tconn, coords, tdesc, blobs, truns, vals, 0, st);
    }
    switch (elem_type) {
        case FE_QUAD4:
            return op == FE_OP_MASS ? launch_ws<FE_QUAD4, true>(ntiles, tconn, coords, tdesc, blobs, truns, vals, 1, st)
                                    : launch_ws<FE_QUAD4, false>(ntiles, tconn, coords, tdesc, blobs, truns, vals, w, st);
        case FE_TRI3:
            return op == FE_OP_MASS ? launch_ws<FE_TRI3, true>(ntiles, tconn, coords, tdesc, blobs, truns, vals, 1, st)
                                    : launch_ws<FE_TRI3, false>(ntiles, tconn, coords, tdesc, blobs, truns, vals, w, st);
        default:
            set_error("fe_assemble_ws: element type %d is not handled by this kernel (use fe_assemble_tiled)", elem_type);
            return FE_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
