/*
 * feprogram_b200 -- C ABI of the B200-native FEM hot path
 * (element matrices -> deterministic CSR assembly -> Jacobi-PCG).
 *
 * The reference (ibejarano/feprogram) is pure Python and has no FFI layer: its boundary for this
 * path is the FemProblem/Elem class API driven by main.py:17-40.  This header is the native
 * layer placed directly beneath that API; every entry point cites the reference lines it
 * replaces.  feprogram_b200/elements.py (same class/method names as the reference's
 * elements.py) calls these through feprogram_b200/ops.py (PyTorch custom ops over ctypes).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No C++ types, no exceptions cross the boundary.
 *   - Every function returns FE_OK (0) or a negative FE_ERR_* code; fe_last_error() gives the
 *     thread-local message of the last failure.
 *   - All array arguments are CALLER-OWNED DEVICE pointers (contiguous; fp64 / int32 / int64 /
 *     uint8 as declared) unless the parameter is documented "host".  The library allocates no
 *     device memory: scratch arrays are explicit caller-provided arguments.
 *   - Every device function takes a stream (a cudaStream_t passed as void*) and is stream-ordered:
 *     no hidden synchronisation, no hidden host<->device copies except where documented
 *     (fe_pcg polls one 64-byte status block every `check_every` iterations).
 *   - Thread-safe for distinct streams.  Multi-GPU = one host process per GPU; communicators
 *     are passed in as ncclComm_t (void*).
 *   - Connectivity on the device is int32, 0-BASED, row-major (nelem x nnpe).  The host facade
 *     converts the reference's 1-based tags (elements.py:172 `elem-1`).  Coordinates are fp64
 *     (nnode x 2), interleaved x,y (elements.py:34-38).
 *   - DOF numbering is the reference's: dof = ndof*node + c (elements.py:148), ndof = 2.
 *
 * Storage produced by the symbolic phase (all index arrays in the reference's numbering)
 *   eptr[nnode+1] int32, einc[nelem*nnpe] int32 : node -> incident (element,local node) pairs,
 *        einc = element*16 + a, ascending element id inside each node (the summation order).
 *   nptr[nnode+1] int32, nadj[nnzN] int32       : node adjacency (sorted, self included) --
 *        the CSR pattern at node granularity; the dof-level pattern is its ndof x ndof blow-up.
 *   epos[nelem*nnpe*nnpe] uint8                 : for incidence entry k and local node b, the
 *        position of node conn[e][b] inside nadj[nptr[n]..): the element -> CSR-slot map.
 *   row_ptr[ndof*nnode+1] int32|int64, col_idx[nnz] int32 : scipy-compatible CSR pattern,
 *        bit-identical to FemProblem.K.tocsr() of the reference before BCs (elements.py:183).
 *   vals[nnz] fp64 : CSR values; value of (row ndof*n+c, col ndof*nadj[nptr[n]+p]+d) lives at
 *        ndof*ndof*nptr[n] + c*ndof*deg(n) + ndof*p + d.
 */
#ifndef FEPROGRAM_B200_H
#define FEPROGRAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------------------- */
#define FE_OK 0
#define FE_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unknown enum)      */
#define FE_ERR_UNSUPPORTED (-2) /* element type / operator / ndof combination not implemented    */
#define FE_ERR_CUDA (-3)        /* a CUDA runtime call or kernel launch failed                    */
#define FE_ERR_CAPACITY (-4)    /* a documented per-node capacity was exceeded (see fe_limits)    */
#define FE_ERR_NCCL (-5)        /* an NCCL call failed                                            */

/* ---- element types (elements.py:55: Quad4 if len(conectivity[0])==4 else Quad9) --------- */
#define FE_QUAD4 4
#define FE_QUAD9 9
#define FE_TRI3 3 /* extension, not in the reference */

/* ---- operators -------------------------------------------------------------------------- */
#define FE_OP_REFERENCE 0 /* Ke = sum_g B^T B detJ, 2 dof/node, no weights (elements.py:153-163) */
#define FE_OP_WEIGHTED 1  /* same, each Gauss term times gpWei[g] (elements.py:137; extension)   */
#define FE_HINT_IRREGULAR_TILES 256 /* may be or-ed into `op` of fe_assemble_tiled: the plan's tiles are irregular
                                       (packed sub-blocks); selects a kernel schedule, never changes a result */
#define FE_OP_MASS 2      /* Me[2a+c,2b+c] = sum_g gpWei[g] H_g[a] H_g[b] detJ_g from the reference's own H and
                             gpWei tables (elements.py:94-113, built but unused there; extension); entries
                             across the two components are stored zeros, so K and M share one pattern */
#define FE_OP_POISSON 3   /* scalar operator, ONE dof per node: Ks[a,b] = sum_g (Dx_a Dx_b + Dy_a Dy_b) detJ_g, i.e.
                             exactly the even/even sub-block K[0::2, 0::2] of the reference operator
                             (elements.py:15-23, 153-163; BASELINE's "2D Poisson" wording; extension).  Its CSR
                             pattern is the node-level pattern itself (row_ptr = nptr, col_idx = nadj) and
                             vals[nptr[n] + p] is the value of (row n, column nadj[nptr[n] + p]).  Accepted by
                             fe_element_matrices (Ke is nnpe x nnpe), fe_assemble, fe_assemble_tiled and
                             fe_assemble_ws; the
                             values are bit-identical to the corresponding entries of FE_OP_REFERENCE. */

/* Copies the thread-local message of the last failing call into buf (NUL-terminated). */
void fe_last_error(char* buf, size_t len);

/* Library / build identification: "feprogram_b200 <version> sm_100a". */
const char* fe_version(void);

/* Compile-time capacities: out[0] = max distinct neighbours per node, out[1] = max incident
 * elements per node, out[2] = max nnpe, out[3] = max Gauss points. */
void fe_limits(int32_t out[4]);

/* Small read-back (sizes, status words) without the copy engine: a kernel stores nbytes (multiple of 4, <= 1 MiB)
 * from device memory into PINNED host memory (cudaHostAlloc / cudaMallocHost: device-accessible under unified
 * addressing).  Stream-ordered; the caller synchronises the stream (or an event) before reading the host copy.
 * Unlike a cudaMemcpy it does not queue behind a large download that is in flight on another stream. */
int fe_peek(const void* dev_src, void* pinned_host_dst, int64_t nbytes, void* stream);

/* ---- T1-T4: tables (host) ------------------------------------------------------------------
 * Replaces funH4/funH9/funHrs4/funHrs9 (tools/interpFunctions.py:3-84) evaluated on the tensor
 * Gauss grid by quadH/quadHrs (elements.py:94-140), same expression order => bit-identical.
 * HOST pointers: dH[ng2*2*nnpe] (g-major, then r/s row, then node), H[ng2*nnpe], w[ng2].
 * Any of the outputs may be NULL.  *nnpe_out / *ng2_out may be NULL. */
int fe_tables(int elem_type, double* dH, double* H, double* w, int32_t* nnpe_out, int32_t* ng2_out);

/* ---- input conversion ---------------------------------------------------------------------------
 * conn[i] <- tags[i] - 1: the reference's 1-based gmsh tags (int32, or int64/uint64 as gmsh returns
 * them; tag_bytes = 4 or 8) to the int32 0-based connectivity every other entry point takes
 * (elements.py:172 `self.getKe(elem-1)`). */
int fe_tags_to_conn(int tag_bytes, int64_t n, const void* tags, int32_t* conn, void* stream);

/* ---- E1-E3: stand-alone element matrices -----------------------------------------------------
 * Replaces Elem.getpos / Elem.funB / FemProblem.getKe (elements.py:34-38, 15-23, 153-163) for a
 * batch: Ke[e] is the dense row-major (ndof*nnpe)^2 block of element e.  Debug / parity entry;
 * the assembly path never materialises Ke in HBM. */
int fe_element_matrices(int elem_type, int op, int64_t nelem, const int32_t* conn, const double* coords,
                        double* Ke, void* stream);

/* ---- A1/A3 symbolic phase (one-time per mesh) ---------------------------------------------
 * Replaces getRowData (elements.py:143-151) and the pattern half of
 * coo_matrix(...).tolil() (elements.py:183).  Steps, in order; the caller allocates between
 * steps from the counts the previous step produced (read back eptr[nnode] / nptr[nnode]):
 *
 *  1. fe_incidence_count : eptr[0..nnode] <- exclusive scan of per-node element counts.
 *                          scratch: scan_ws (fe_scan_workspace_bytes(nnode+1) bytes).
 *  2. fe_incidence_fill  : einc <- (element*16 + a), ascending inside each node.
 *                          scratch: cursor[nnode] int32.
 *  3. fe_adjacency_count : nptr[0..nnode] <- exclusive scan of distinct-neighbour counts;
 *                          status[0] != 0 afterwards means FE_ERR_CAPACITY on some node.
 *  4. fe_adjacency_fill  : nadj, epos.
 *  5. fe_csr_expand      : row_ptr (int32 if idx64 == 0 else int64), col_idx.
 */
size_t fe_scan_workspace_bytes(int64_t n);
int fe_incidence_count(int elem_type, int64_t nelem, int64_t nnode, const int32_t* conn, int32_t* eptr,
                       void* scan_ws, void* stream);
int fe_incidence_fill(int elem_type, int64_t nelem, int64_t nnode, const int32_t* conn, const int32_t* eptr,
                      int32_t* cursor, int32_t* einc, void* stream);
int fe_adjacency_count(int elem_type, int64_t nnode, const int32_t* conn, const int32_t* eptr,
                       const int32_t* einc, int32_t* nptr, int32_t* status, void* scan_ws, void* stream);
int fe_adjacency_fill(int elem_type, int64_t nnode, const int32_t* conn, const int32_t* eptr,
                      const int32_t* einc, const int32_t* nptr, int32_t* nadj, uint8_t* epos, void* stream);
int fe_csr_expand(int ndof, int64_t nnode, const int32_t* nptr, const int32_t* nadj, int idx64, void* row_ptr,
                  int32_t* col_idx, void* stream);

/* ---- A2/A3 numeric assembly -------------------------------------------------------------------
 * Replaces the element loop + duplicate-summing COO->CSR conversion of FemProblem.assemble
 * (elements.py:165-183): evaluates every element matrix on the fly (never stored) and reduces
 * the contributions of each CSR slot in ascending element order -- no atomics, bit-identical
 * from run to run.  Writes every entry of vals exactly once.
 * max_deg = max_n (nptr[n+1]-nptr[n]) (sizes the shared-memory row tile; caller computed). */
int fe_assemble(int elem_type, int op, int64_t nelem, int64_t nnode, const int32_t* conn, const double* coords,
                const int32_t* eptr, const int32_t* einc, const uint8_t* epos, const int32_t* nptr,
                int32_t max_deg, double* vals, void* stream);

/* ---- A2/A3 numeric assembly, tiled form (the fast path) ------------------------------------------
 * Same contract and bit-identical results as fe_assemble, ~3x fewer fp64 operations: a tile is
 * TE = out[0] of fe_tile_params consecutive elements of a spatially compact element order; each
 * element matrix is evaluated once per tile into shared memory (upper-triangular block form), every
 * node (CSR row block) is owned by the tile of its first incident element, and elements of
 * neighbouring tiles that touch an owned node ("halo items") contribute just that node's row pair.
 *
 * One-time plan, in order (s = ntiles + 1):
 *  1. fe_morton_keys        : 30-bit Morton key of each element centroid on a grid of cell_x x cell_y
 *                             cells anchored at (xmin, ymin) (a cell ~ one element).  The caller sorts
 *                             elements by key (stable) -> eperm and cuts the sorted list into tiles:
 *                             a tile = the elements of one aligned Z-block of 2^7 cells, split further
 *                             when it holds more than TE elements.  That yields ntiles,
 *                             tile_estart[s] (offsets into eperm) and eloc[nelem] = tile << 8 | row.
 *                             (Any other tiling with <= TE elements per tile is equally valid, e.g.
 *                             runs of the input order.)
 *  2. fe_tile_owner         : owner[nnode] (smallest tile id among the node's elements, -1 if none),
 *                             tile_ncount[s] (owned nodes per tile, last entry 0).
 *  3. fe_exclusive_scan_i32 : tile_ptr[0..s) <- scan(tile_ncount)
 *  4. fe_tile_layout        : tile_nodes (ascending node id inside each tile), noff[2*nslots]
 *                             (item / work offset of each tile-node slot inside its tile),
 *                             thalo[ntiles*out[1]] (sorted distinct halo elements of each tile),
 *                             tile_counts[5*s] (per tile: items (even), halo elements, work items
 *                             (x32), record-blob size / 16 and runs of consecutive owned node ids --
 *                             the last two for fe_assemble_ws), stats[8]: [0] max owned nodes, [1] max
 *                             items, [2] max halo elements of a tile, [3] max incident elements of a
 *                             node, [5] max work items, [6] max blob bytes, [7] max bytes of CSR values
 *                             of a tile.  scratch: cursor[ntiles].
 *  5. fe_exclusive_scan_i32 x3 (x5) : tile_ptr[s..2s), [2s..3s), [3s..4s) (and blob_off, run_off) <- scans
 *                             of the rows of tile_counts
 *  6. fe_tile_records       : tnode (int2 per slot), tloc (uint16 per item), twork (uint32 per work
 *                             item, ZERO-INITIALISED by the caller), tchunk (uint32 per 32 work
 *                             items); stats[4] != 0 afterwards means the mesh is not representable
 *                             (use fe_assemble).
 * Steps 2-5 (and the tile-ordered connectivity, fe_tile_conn) need the incidence and the node DEGREES (nptr) only,
 * not nadj / epos: a caller may run them concurrently with fe_adjacency_fill (the facade does, on a side stream).
 * The plan is usable iff stats[0] <= out[2], stats[1] < out[3], stats[2] <= out[1], stats[5] <= out[4]
 * of fe_tile_params and stats[4] == 0.  Record formats are documented in feprogram_b200/csrc/fe_tiles.cu. */
/* Bounding box of the mesh (anchors the Morton grid): out[0..3] <- {xmin, ymin, xmax, ymax} (DEVICE, 32-byte aligned);
 * scratch: 592 * 4 doubles (32-byte aligned). */
int fe_bbox(int64_t nnode, const double* coords, double* scratch, double* out, void* stream);
/* Morton keys WITHOUT coordinates, for lattice numberings: when every Quad4 element is {b, b+1, b+W+1, b+W} for one W
 * (row-major structured meshes, whatever their geometry), (b mod W, b div W) are the element's lattice coordinates.
 * Also accepted: {b, b+1, c+1, c} with b, c in the same column of two node rows of W consecutive ids in any row order
 * (the local mesh of a rank of a row-block partition: owned node rows first, ghost rows after them); the row
 * coordinate is then c div W.
 * status[0] != 0 afterwards: the mesh does not fit (use fe_morton_keys); status[1] = W.  The tile plan built from
 * these keys needs no coordinates, so the facade builds it while the coordinates are still uploading. */
int fe_lattice_keys(int elem_type, int64_t nelem, const int32_t* conn, int32_t* keys, int32_t* status, void* stream);
/* Step 1 without a global sort (the default path of the Python facade): fe_tile_form_count buckets the elements
 * by Morton block (keys >> log2(TE); nb blocks) -- histogram, tiles per block, scans, scatter -- after which
 * tbase[nb] = ntiles and status[0] != 0 flags a block of more than 1024 elements (use another tiling);
 * fe_tile_form_sort orders every block (a register bitonic network for blocks of at most one tile, rank sorts in
 * shared memory for larger ones) and writes eperm, eloc and tile_estart[ntiles + 1].
 * Same result as sorting by (key, element id), cutting blocks into runs of TE and ordering each tile by id.
 * Scratch: hist, tpb, ebase, tbase [nb + 1], cursor [nb], tmp [nelem], status [1] int32; scan_ws for nb + 1. */
int fe_tile_form_count(int elem_type, int64_t nelem, int64_t nb, const int32_t* keys, int32_t* hist, int32_t* tpb,
                       int32_t* ebase, int32_t* tbase, int32_t* cursor, int32_t* tmp, int32_t* status, void* scan_ws,
                       void* stream);
int fe_tile_form_sort(int elem_type, int64_t nb, const int32_t* ebase, const int32_t* tbase, const int32_t* tmp,
                      const int32_t* keys, int32_t* eperm, int32_t* eloc, int32_t* tile_estart, void* stream);
void fe_tile_params(int elem_type, int32_t out[6]);
int fe_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scan_ws, void* stream);
int fe_morton_keys(int elem_type, int64_t nelem, const int32_t* conn, const double* coords, double xmin,
                   double ymin, double cell_x, double cell_y, int32_t* keys, void* stream);
int fe_invert_permutation(int64_t n, const int32_t* perm, int32_t* inv, void* stream);
int fe_tile_owner(int64_t nnode, int64_t ntiles, const int32_t* eptr, const int32_t* einc, const int32_t* eloc,
                  int32_t* owner, int32_t* tile_ncount, void* stream);
int fe_tile_layout(int elem_type, int64_t nnode, int64_t ntiles, const int32_t* owner, const int32_t* tile_nptr,
                   const int32_t* eptr, const int32_t* einc, const int32_t* eloc, const int32_t* nptr,
                   int32_t* cursor, int32_t* tile_nodes, int32_t* noff, int32_t* thalo, int32_t* tile_counts,
                   int32_t* stats, void* stream);
int fe_tile_records(int elem_type, int64_t nslots, const int32_t* tile_nodes, const int32_t* owner,
                    const int32_t* tile_ptr, int64_t ntiles, const int32_t* noff, const int32_t* thalo,
                    const int32_t* eptr, const int32_t* einc, const uint8_t* epos, const int32_t* eloc,
                    const int32_t* nptr, const int32_t* nadj, void* tnode, uint16_t* tloc, uint32_t* twork,
                    uint32_t* tchunk, int32_t* stats, void* stream);
/*  7. fe_tile_conn : tconn[ntiles * out[5] * nnpe] <- connectivity in tile order (own elements then halo
 *                     elements of each tile, -1 rows unused), so the numeric kernel streams it. */
int fe_tile_conn(int elem_type, int64_t ntiles, const int32_t* conn, const int32_t* eperm,
                 const int32_t* tile_estart, const int32_t* thalo, const int32_t* tile_ptr, int32_t* tconn,
                 void* stream);
/* tdesc[8*t..] = {node offset, node count, item offset, item count, work offset, work count, halo element
 * count, 0} of tile t (from tile_ptr rows 0, 1, 3, 2), 32 bytes per tile, 16-byte aligned. */
int fe_assemble_tiled(int elem_type, int op, int64_t ntiles, const int32_t* tconn, const double* coords,
                      const int32_t* tdesc, const void* tnode, const uint16_t* tloc, const uint32_t* twork,
                      const uint32_t* tchunk, double* vals, void* stream);

/* ---- A2/A3 numeric assembly, warp-specialised tiled form (Quad4 / Tri3; the benchmarked kernel) ----------
 * Same contract, tiling and bit-identical results as fe_assemble_tiled; one persistent CTA per SM whose producer
 * warps (element matrices -> shared-memory slab), consumer warps (slab -> CSR blocks staged in their final layout)
 * and DMA warp (1-D TMA bulk copies: record blobs in, contiguous CSR row ranges out) hand tiles to each other
 * through mbarriers.  Its plan reuses steps 1-5 and 7 of the fe_assemble_tiled plan (owner, layout, halo lists,
 * tile_ptr rows, tconn; blob_off / run_off = the scans of rows 3 and 4 of fe_tile_layout's tile_counts) and
 * replaces step 6 by
 *  6'. fe_tile_ws_build : blobs (16 * blob_off[ntiles] bytes), truns (16 bytes per run), tdesc (4 int32 per tile);
 *                         stats[4] != 0 afterwards means the mesh is not representable (use fe_assemble).
 * The plan is usable iff fe_tile_layout's stats[6] <= out[0], stats[7] <= out[1], stats[0] <= out[3] of
 * fe_tile_ws_params (all 0 for element types this kernel does not handle) and stats[4] == 0 after the build.
 * Blob / run formats are documented in feprogram_b200/csrc/fe_tiles_ws.cu.  vals and blobs must be 16-byte aligned.
 * A node with more than out[2] distinct neighbours makes the plan unrepresentable (stats[4]).  op = FE_OP_POISSON
 * uses the SAME plan (the plan holds no dof count) and writes one value per node-level non-zero. */
void fe_tile_ws_params(int elem_type, int32_t out[4]);
int fe_tile_ws_build(int elem_type, int64_t nslots, int64_t ntiles, const int32_t* tile_nodes, const int32_t* owner,
                     const int32_t* tile_ptr, const int32_t* noff, const int32_t* thalo, const int32_t* eptr,
                     const int32_t* einc, const uint8_t* epos, const int32_t* eloc, const int32_t* nptr,
                     const int32_t* nadj, const int32_t* blob_off, const int32_t* run_off, void* blobs, void* truns,
                     int32_t* tdesc, int32_t* stats, void* stream);
int fe_assemble_ws(int elem_type, int op, int64_t ntiles, const int32_t* tconn, const double* coords,
                   const int32_t* tdesc, const void* blobs, const void* truns, double* vals, void* stream);

/* ---- B1-B3: boundary conditions ---------------------------------------------------------------
 * fe_bc_rhs replaces the brhs writes of elements.py:185-197: rhs <- 0; rhs[dir] <- 0;
 * rhs[neu] <- neu_value (Neumann written last, so it wins on shared dofs); dirmask[d] <- 1 on
 * Dirichlet dofs, 0 elsewhere.  dofs are int64 device arrays.
 *
 * fe_apply_bc modifies vals in place.  mode FE_BC_ROWS is the reference's row-only treatment
 * (elements.py:185-193): Dirichlet rows -> 0 with unit diagonal, columns untouched, rhs untouched.
 * mode FE_BC_SYMMETRIC additionally eliminates Dirichlet columns, moving them to the right-hand
 * side: rhs[j] -= K[j,d]*rhs[d] (the SPD system equivalent to the reference's).
 * row_ptr is int32 (idx64 == 0) or int64. */
#define FE_BC_ROWS 0
#define FE_BC_SYMMETRIC 1
int fe_bc_rhs(int64_t nrows, const int64_t* dir_dofs, int64_t n_dir, const int64_t* neu_dofs, int64_t n_neu,
              double neu_value, double* rhs, uint8_t* dirmask, void* stream);
int fe_apply_bc(int mode, int64_t nrows, int idx64, const void* row_ptr, const int32_t* col_idx, double* vals,
                double* rhs, const uint8_t* dirmask, void* stream);
/* "Corrected" load (extension): consistent nodal load of a constant traction (tx, ty)
 * on boundary edges instead of the reference's mesh-independent nodal 1.0 (elements.py:195-197 FIXME):
 * rhs[2n+c] += t_c * int N_n ds.  edges: DEVICE int32 [nedge][nodes_per_edge], 0-based, (a, b) or
 * (a, b, mid).  Dirichlet dofs (dirmask != 0; dirmask may be NULL) are not touched.  Call after fe_bc_rhs. */
int fe_edge_load(int nodes_per_edge, int64_t nedge, const int32_t* edges, const double* coords, double tx, double ty,
                 const uint8_t* dirmask, double* rhs, void* stream);

/* ---- S: solve -----------------------------------------------------------------------------------
 * Replaces spsolve(K, brhs) (elements.py:241-243) by Jacobi-preconditioned CG on the SPD system
 * the reference's Dirichlet treatment is equivalent to.
 *
 * fe_spmv: y <- A x with A in CSR.  If rowmask != NULL rows with rowmask[r] != 0 give y[r] = 0.
 *   nnz (= row_ptr[nrows], a HOST value the caller knows from the symbolic phase) only selects the
 *   lanes-per-row schedule; passing it keeps the call free of any device read-back or synchronisation.
 *
 * fe_csr_diagonal: diag[r] <- A[r,r] (0 if not stored).
 *
 * fe_pcg: solves the Dirichlet-constrained system in place on the UNMODIFIED (pre-BC) matrix:
 *   x_D = rhs_D, and CG runs on the remaining dofs with the Dirichlet columns moved to the
 *   right-hand side (dirmask may be NULL: plain SPD solve).  x is the output (initial guess
 *   ignored: starts from 0 / the Dirichlet lift).  Converged when ||r||_2 <= rtol*||r0||_2 or
 *   ||r||_2 <= atol.  work: 4*nrows doubles + FE_PCG_SCALAR_DOUBLES doubles + 2*fe_pcg_blocks(nrows)
 *   doubles of scratch, see fe_pcg_workspace_bytes.  info (HOST, 4 doubles): iterations, final
 *   ||r||/||r0||, ||r0||, stop reason: 1 = converged (||r|| <= tolerance), 0 = iteration limit reached,
 *   -2 = CG breakdown (NaN residual or p.Ap <= 0: the matrix is not SPD on the free dofs; x is not a
 *   solution), -1 = a peer-memory wait timed out (fe_dist_pcg_p2p only; that call also returns an error).
 *   check_every: iterations between host polls (>=1): the one documented synchronisation of this header.
 *   Systems of up to 2^27 non-zeros (launch-bound iterations) replay each batch of check_every iterations from a
 *   CUDA graph captured once per call on a private stream of the library; the batches still run on `stream`.
 *   nptr/nadj (optional, may be NULL): node-level pattern of a 2-dof matrix -> node-block SpMV.
 */
int fe_spmv(int64_t nrows, int64_t nnz, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
            const double* x, double* y, const uint8_t* rowmask, void* stream);
/* Same product read through the node-level pattern of a 2-dof matrix in this library's value layout (one
 * column index per 2x2 block): 24 % fewer bytes.  fe_pcg / fe_dist_pcg use it when nptr/nadj are given. */
int fe_spmv_nodal(int64_t nnode, const int32_t* nptr, const int32_t* nadj, const double* vals, const double* x,
                  double* y, const uint8_t* rowmask, void* stream);
int fe_csr_diagonal(int64_t nrows, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
                    double* diag, void* stream);
size_t fe_pcg_workspace_bytes(int64_t nrows);
int fe_pcg(int64_t nrows, int64_t nnz, int idx64, const void* row_ptr, const int32_t* col_idx, const double* vals,
           const double* rhs, const uint8_t* dirmask, double* x, double rtol, double atol, int64_t maxit,
           int32_t check_every, void* work, double* info, const int32_t* nptr, const int32_t* nadj, void* stream);

/* ---- post-processing of the solved field -----------------------------------------
 * fe_gauss_gradients: grad[e][g][0..3] = H_grad_g u_e = (du/dx, du/dy, dv/dx, dv/dy) at every Gauss point,
 *   the operator Elem.getHgrad builds (elements.py:25-32; defined there, never called).  U is the solution
 *   as solveProblem returns it (2 * nnode doubles, node-interleaved); grad has nelem * ng2 * 4 doubles.
 * fe_nodal_strain: strain[n][0..2] = mean over the elements incident to node n (eptr/einc from
 *   fe_incidence_*: element * 16 + local node, ascending element id) of the element's Gauss-point mean of B u_e =
 *   (du/dx, dv/dy, du/dy + dv/dx) (Elem.funB rows 0-2, elements.py:15-23): the 3-component nodal array the
 *   reference's writer reserves as "Tensiones" (python_to_xml.py:19-22, commented out there). */
int fe_gauss_gradients(int elem_type, int64_t nelem, const int32_t* conn, const double* coords, const double* U,
                       double* grad, void* stream);
int fe_nodal_strain(int elem_type, int64_t nnode, const int32_t* eptr, const int32_t* einc, const int32_t* conn,
                    const double* coords, const double* U, double* strain, void* stream);

/* ---- multi-GPU (one host process per GPU, NCCL over NVLink) ------------------------------------------
 * The reference has no distributed code; this is the north_star's scale-out of the solve.  Elements are
 * partitioned into contiguous blocks; rank r owns the nodes whose first incident element it holds and
 * assembles exactly those rows from its own elements plus one layer of recomputed ghost elements (no
 * communication during assembly; fe_assemble* on the local mesh).  Local dof numbering: owned dofs
 * first, then ghost dofs grouped by owning rank (ascending rank, ascending global id inside a rank), so a
 * neighbour's contribution is received straight into a contiguous segment of the vector.
 *
 * fe_dist_available      : 1 when the library was built with NCCL.
 * fe_dist_unique_id      : HOST, 128-byte ncclUniqueId (rank 0 creates it, the caller broadcasts it).
 * fe_dist_comm_init      : HOST, collective; *comm_out is an ncclComm_t.
 * fe_dist_halo_exchange  : vec[ghost segments] <- owners' values.  nbr_rank[n_nbr], send_off[n_nbr+1],
 *                          recv_off[n_nbr+1] are HOST arrays; send_idx (owned dof indices, neighbour-major)
 *                          and sendbuf[send_off[n_nbr]] are DEVICE arrays; recv_off are offsets into vec.
 * fe_dist_pcg            : fe_pcg over all ranks: rows [0, nrows_owned) of the local CSR matrix (columns
 *                          are local dof ids < nrows_local), one halo exchange + two scalar all-reduces
 *                          per iteration; every rank returns the same iteration count and residual.
 *                          x, rhs, dirmask have nrows_local entries; on return x is valid on owned AND
 *                          ghost dofs.  work: fe_dist_pcg_workspace_bytes(nrows_local).  nnz_owned =
 *                          row_ptr[nrows_owned] as a HOST value (schedule selection only, no read-back). */
int fe_dist_available(void);
int fe_dist_unique_id(uint8_t* out128);
int fe_dist_comm_init(const uint8_t* id128, int rank, int nranks, void** comm_out);
int fe_dist_comm_destroy(void* comm);
int fe_dist_halo_exchange(void* comm, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_off,
                          const int32_t* send_idx, double* sendbuf, const int64_t* recv_off, double* vec,
                          void* stream);
size_t fe_dist_pcg_workspace_bytes(int64_t nrows_local);
int fe_dist_pcg(int64_t nrows_owned, int64_t nrows_local, int64_t nnz_owned, int idx64, const void* row_ptr,
                const int32_t* col_idx, const double* vals, const double* rhs, const uint8_t* dirmask, double* x,
                double rtol, double atol, int64_t maxit, int32_t check_every, void* work, double* info, void* comm,
                int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_off, const int32_t* send_idx,
                double* sendbuf, const int64_t* recv_off, const int32_t* nptr, const int32_t* nadj, void* stream);
/* fe_dist_pcg with the per-iteration halo taken out of NCCL: the SpMV loads the ghost entries of the search
 * direction directly from the owning GPU's memory over NVLink (peer pointers from CUDA IPC / torch symmetric
 * memory), gated by per-neighbour version flags.  Same arithmetic, same iterates as fe_dist_pcg (bitwise).
 *   peer_p     DEVICE array [world] of double*: every rank's direction buffer `pbuf` (2 * pstride doubles,
 *              pstride >= nrows_local on every rank); pbuf is this rank's own entry.
 *   peer_flags DEVICE array [world] of int64*: every rank's flag array (int64[world], zero-initialised once).
 *   ghost_rank / ghost_ridx  DEVICE int32 [nnode_local - nnode_owned]: owner rank and node index inside the
 *              owner for every ghost node;  nbr_dev DEVICE copy of nbr_rank.
 *   epoch      version base; the call uses versions (epoch, epoch + maxit + 1], so the caller advances it
 *              by maxit + 2 between calls (flags only ever grow).
 *   peer_slots / peer_versions  DEVICE arrays [world] of peer pointers to every rank's reduction table
 *              (double[64]) and version array (int64[16], zero-initialised once): the two scalar all-reduces
 *              of an iteration are stores into every peer's table followed by a release-store of the version;
 *              consumers add the `world` slots in rank order (same bits everywhere).  seq: reduction counter,
 *              the call uses [seq, seq + 2 * maxit + 1); the caller advances it between calls.  world <= 16.
 * Only the first / last exchange of x stay on NCCL. */
int fe_dist_pcg_p2p(int64_t nrows_owned, int64_t nrows_local, int idx64, const void* row_ptr, const int32_t* col_idx,
                    const double* vals, const double* rhs, const uint8_t* dirmask, double* x, double rtol, double atol,
                    int64_t maxit, int32_t check_every, void* work, double* info, void* comm, int32_t n_nbr,
                    const int32_t* nbr_rank, const int64_t* send_off, const int32_t* send_idx, double* sendbuf,
                    const int64_t* recv_off, const int32_t* nptr, const int32_t* nadj, int32_t rank, int32_t world,
                    const void* peer_p, const void* peer_flags, const int32_t* ghost_rank, const int32_t* ghost_ridx,
                    const int32_t* nbr_dev, int64_t pstride, double* pbuf, int64_t epoch, const void* peer_slots,
                    const void* peer_versions, int64_t seq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEPROGRAM_B200_H */
