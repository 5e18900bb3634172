/*
 * feddb200.h -- C ABI of the B200-native finite-element assembly engine (libfeddb200.so).
 *
 * Drop-in boundary for FEDDLib's element-wise matrix assembly hot path.  Every entry point
 * cites the reference interface it replaces (paths under /root/reference/feddlib).  The
 * reference's host side (class FE<SC,LO,GO,NO>) keeps its C++ signatures; a glue header
 * (feddlib_b200/csrc/host/FE_b200.hpp) extracts flat arrays from
 * `Domain`, calls the functions below and wraps the returned CSR arrays in a Tpetra
 * CrsMatrix on the same row/column maps.
 *
 * Conventions
 *  - return 0 on success, <0 on error; feddb200_last_error() gives the text (thread-local).
 *    This mirrors the reference's TEUCHOS_TEST_FOR_EXCEPTION(std::logic_error|runtime_error)
 *    convention (core/FE/FE_def.hpp:610,676,1691,2746,6950): FEDDB200_ELOGIC for argument /
 *    "not implemented for this FE type" errors, FEDDB200_ERUNTIME for CUDA/runtime failures.
 *  - plain pointers and sizes only.  Pointers are HOST pointers unless the name ends in _d
 *    (device pointer on the context's device).  The library never frees caller memory.
 *  - there is no CPU fallback: every call fails loudly if no CUDA device is usable.
 *  - local ids are int32 (reference LO = int, core/General/DefaultTypeDefs.hpp:7), global ids
 *    int64 (GO = long long), scalars double (SC = double).
 */
#ifndef FEDDB200_H
#define FEDDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEDDB200_OK        0
#define FEDDB200_ELOGIC   (-1)   /* std::logic_error in the reference   */
#define FEDDB200_ERUNTIME (-2)   /* std::runtime_error in the reference */

typedef struct feddb200_ctx  feddb200_ctx;   /* one per GPU / rank                                  */
typedef struct feddb200_mesh feddb200_mesh;  /* uploaded connectivity + repeated points              */
typedef struct feddb200_pat  feddb200_pat;   /* node-level CSR pattern + scatter/gather maps         */

/* scatter modes (north_star: coloured-deterministic with an atomic variant alongside; the
 * row-gather mode is the B200-first default: every CSR value is written exactly once) */
#define FEDDB200_SCATTER_ATOMIC   0
#define FEDDB200_SCATTER_COLOURED 1
#define FEDDB200_SCATTER_GATHER   2

/* dof layouts of a matrix on a node-level pattern (node-wise numbering dim*g+d,
 * core/LinearAlgebra/Map_def.hpp:95-108) */
#define FEDDB200_BLOCK_SCALAR 0  /* 1 dof per row node, 1 per col node   (assemblyLaplace)            */
#define FEDDB200_BLOCK_DIAG   1  /* dim x dim, only (d,d) stored         (LaplaceVecField, Advection) */
#define FEDDB200_BLOCK_FULL   2  /* row_dofs x col_dofs full blocks      (LinElas, AdvectionInU, B, BT) */

const char *feddb200_last_error(void);
int  feddb200_device_count(void);

/* ---- context ------------------------------------------------------------------------- */
int  feddb200_create(feddb200_ctx **ctx, int device);
void feddb200_destroy(feddb200_ctx *ctx);
/* run all work of this context on an existing CUDA stream (cudaStream_t as void*; NULL = the
 * legacy default stream, as in the CUDA runtime).  A new context runs on its own non-blocking
 * stream; feddb200_use_own_stream switches back to it. */
int  feddb200_set_stream(feddb200_ctx *ctx, void *cuda_stream);
int  feddb200_use_own_stream(feddb200_ctx *ctx);
int  feddb200_set_scatter_mode(feddb200_ctx *ctx, int mode);
int  feddb200_get_scatter_mode(const feddb200_ctx *ctx);
/* Row phases of the gather mode, for overlapping the ghost-row exchange with the assembly of the owned rows:
 * ROWS_GHOST = geometry pre-pass + the rows owned by other ranks (their values, contiguous behind the owned
 * ones, can be shipped as soon as this call's work is done); ROWS_OWNED = the owned rows, using the geometry
 * of the preceding ROWS_GHOST call on the same pattern; ROWS_ALL (default) = both.  The atomic and coloured
 * modes do everything in the ROWS_GHOST call. */
#define FEDDB200_ROWS_ALL   0
#define FEDDB200_ROWS_GHOST 1
#define FEDDB200_ROWS_OWNED 2
/* the ROWS_GHOST call in two parts, so that the ghost rows can run on another stream NEXT TO the owned rows (only the
 * geometry pre-pass is a true dependency of both): ROWS_GEOM = the geometry pre-pass alone, ROWS_GHOST_ONLY = the rows
 * owned by other ranks, using the geometry of a preceding ROWS_GEOM call (gather mode; the element-wise modes treat
 * ROWS_GEOM as a no-op and ROWS_GHOST_ONLY like ROWS_GHOST) */
#define FEDDB200_ROWS_GEOM       3
#define FEDDB200_ROWS_GHOST_ONLY 4
int  feddb200_set_row_phase(feddb200_ctx *ctx, int phase);
int  feddb200_synchronize(feddb200_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t feddb200_launch_count(const feddb200_ctx *ctx);

/* ---- device memory helpers (so a pure C/C++ host can keep values resident) ------------- */
int  feddb200_dev_alloc(feddb200_ctx *ctx, void **ptr_d, int64_t bytes);
int  feddb200_dev_free(feddb200_ctx *ctx, void *ptr_d);
int  feddb200_copy_h2d(feddb200_ctx *ctx, void *dst_d, const void *src, int64_t bytes);
int  feddb200_copy_d2h(feddb200_ctx *ctx, void *dst, const void *src_d, int64_t bytes);
/* page-locked host memory for the buffers the host-pointer entry points fill (the CSR values that the seat step hands to
 * Tpetra: core/LinearAlgebra/Matrix_def.hpp:46-51 allocates them inside Tpetra in the reference): a device-to-host copy
 * into pageable memory runs at a fraction of the PCIe rate.  ctx may be NULL for feddb200_host_free. */
/* Binds the CALLING thread to the CPUs of the NUMA node the context's GPU hangs off (PCI bus id -> /sys/bus/pci/devices/.../
 * numa_node -> cpulist), so that page-locked buffers allocated afterwards are node-local and the device-to-host copies of the
 * CSR values do not cross the socket interconnect (8 ranks on a two-socket box otherwise share one socket's memory and links).
 * *numa_node = the node, or -1 when the platform does not tell (nothing is changed then).  Optional; call it once per rank before
 * feddb200_host_alloc / the first host-pointer assembly. */
int  feddb200_bind_host_numa(feddb200_ctx *ctx, int *numa_node);
int  feddb200_host_alloc(feddb200_ctx *ctx, void **ptr, int64_t bytes);
int  feddb200_host_free(feddb200_ctx *ctx, void *ptr);

/* ---- mesh ------------------------------------------------------------------------------
 * Replaces the per-call reads of Domain::getElementsC / getPointsRepeated
 * (core/FE/FE_def.hpp:617-621, core/FE/Domain_def.hpp:421-508) with a one-time upload; the
 * natural call site is FE::addFE (core/FE/FE_def.hpp:64-72).
 *   dim   2|3;  nloc 3,6 (P1/P2 triangle) or 4,10 (P1/P2 tetrahedron)
 *   conn  [ne][nloc]  local (repeated-map) node ids, reference local ordering
 *   coords[nn_rep][dim] repeated points, AoS like vec2D_dbl_Type */
int  feddb200_mesh_upload(feddb200_ctx *ctx, feddb200_mesh **mesh, int dim, int nloc, int64_t ne,
                          const int32_t *conn, int64_t nn_rep, const double *coords);
/* replace the point coordinates (moving meshes / per-step e2e upload) */
int  feddb200_mesh_update_coords(feddb200_ctx *ctx, feddb200_mesh *mesh, const double *coords);
void feddb200_mesh_free(feddb200_mesh *mesh);

/* ---- pattern ---------------------------------------------------------------------------
 * One-time sparsity builder.  Replaces what Matrix(map,numEntries) + insertGlobalValues +
 * fillComplete build dynamically inside Tpetra (core/LinearAlgebra/Matrix_def.hpp:46-51,
 * 88-92, 192-199): the CSR graph over (row node, col node) pairs of every element, rows in
 * row-map order, columns ascending by column-map local index.
 *   row_mesh/col_mesh   share the element index (same ne); identical for square operators,
 *                       pressure/velocity for B and B^T (core/FE/FE_def.hpp:1982-2016)
 *   row_lid[nn_rep(row_mesh)]  row index of each repeated node: 0..n_owned-1 = position in the
 *                       unique (row) map, n_owned..n_rows-1 = ghost rows (nodes owned by another
 *                       rank; their values are shipped to the owner, the Tpetra Export/ADD
 *                       step), -1 = node produces no row.  NULL = identity (one rank: the
 *                       unique map lists the repeated nodes in order, Map_def.hpp:201-206).
 *   col_lid[nn_rep(col_mesh)]  column-map local index of each repeated node; NULL = identity.
 *   extra_row/extra_col [n_extra]  additional (row, col) node entries that other ranks
 *                       contribute to owned rows (received ghost-row structure); may be NULL. */
int  feddb200_pattern_build(feddb200_ctx *ctx, feddb200_pat **pat,
                            const feddb200_mesh *row_mesh, const feddb200_mesh *col_mesh,
                            int64_t n_rows, int64_t n_owned_rows, const int32_t *row_lid,
                            int64_t n_cols, const int32_t *col_lid,
                            int64_t n_extra, const int32_t *extra_row, const int32_t *extra_col);
void feddb200_pat_free(feddb200_pat *pat);

/* sizes: node rows (owned+ghost), owned node rows, node cols, node nnz (all rows), node nnz of
 * owned rows, longest node row, number of element colours */
int  feddb200_pattern_info(const feddb200_pat *pat, int64_t *n_rows, int64_t *n_owned_rows, int64_t *n_cols,
                           int64_t *nnz_nodes, int64_t *nnz_owned_nodes, int32_t *max_row_len, int32_t *n_colours);
/* node-level CSR to host: rowptr[n_rows+1], colind[nnz_nodes] */
int  feddb200_pattern_get_nodes(feddb200_ctx *ctx, const feddb200_pat *pat, int64_t *rowptr, int32_t *colind);
/* number of values of the dof-level matrix (owned + ghost rows) for a layout */
int64_t feddb200_pattern_nnz(const feddb200_pat *pat, int row_dofs, int col_dofs, int block_mode);
/* dof-level CSR (what the Tpetra::CrsMatrix is built from): rowptr[row_dofs*n_rows+1],
 * colind[nnz] (column-map local dof indices, col_dofs*col_node+d) */
int  feddb200_pattern_expand(feddb200_ctx *ctx, const feddb200_pat *pat, int row_dofs, int col_dofs, int block_mode,
                             int64_t *rowptr, int32_t *colind);

/* ---- assembly (values only; the pattern is fixed) ---------------------------------------
 * `values_d` is a device array of feddb200_pattern_nnz(...) doubles in dof-level CSR order;
 * it is fully overwritten.  Each function computes exactly what the cited reference routine
 * inserts, before any scaling the problem classes apply afterwards. */

/* FE::assemblyLaplace (core/FE/FE_def.hpp:604-667) [vec_field=0, BLOCK_SCALAR] and
 * FE::assemblyLaplaceVecField (:670-734) [vec_field=1, BLOCK_DIAG with dim dofs] */
int  feddb200_assemble_laplace_d(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, double *values_d);
/* FE::assemblyMass (core/FE/FE_def.hpp:454-521): fieldType "Scalar" [vec_field=0, BLOCK_SCALAR] / "Vector"
 * [vec_field=1, BLOCK_DIAG with dim dofs] -- SURVEY.md 8(f) rank 2 */
int  feddb200_assemble_mass_d(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, double *values_d);
/* FE::assemblyStress (core/FE/FE_def.hpp:2407-2735): the symmetric-gradient viscous block
 *   K^{ab}_ij = |det B| sum_k func(x_k) w_k (delta_ab grad phi_i . grad phi_j + d_b phi_i d_a phi_j)
 * on a BLOCK_FULL pattern (Stokes_def.hpp:69-70, NavierStokes_def.hpp:142-143 call it with func = 1).  The coefficient
 * callback stays on the host: stress_quadrature gives the reference points of the rule (deg = Grad x Grad), the glue
 * evaluates func at  x_k = B q_k + p_1  for every element (:2479-2484, :2599-2608) and passes either one constant
 * (coef = NULL: all three scatter modes, row-gather kernels) or the array coef[ne][nq] (element-row kernels with the
 * quadrature loop; gather mode is served by the coloured mode). */
int  feddb200_stress_quadrature(int dim, int nloc, int *nq, double *ref_points /*[nq][dim] or NULL*/, double *weights /*[nq] or NULL*/);
int  feddb200_assemble_stress_d(feddb200_ctx *ctx, const feddb200_pat *pat, double coef_const, const double *coef_d, double *values_d);
int  feddb200_assemble_stress(feddb200_ctx *ctx, const feddb200_pat *pat, double coef_const, const double *coef, int64_t n_coef,
                              double *values);
/* FE::assemblyBDStabilization (core/FE/FE_def.hpp:2151-2220): the pressure stabilisation of the
 * P1-P1 Stokes / Navier-Stokes drivers (Stokes_def.hpp:97-104), C_ij = |det B| (sum_q w phi_i phi_j - size * scale) on a
 * P1 x P1 BLOCK_SCALAR pattern; any other FE type is the reference's logic_error "Only implemented for P1". */
int  feddb200_assemble_bdstab_d(feddb200_ctx *ctx, const feddb200_pat *pat, double *values_d);
/* FE::assemblyLinElasXDim (core/FE/FE_def.hpp:2739-3040) [BLOCK_FULL, dim x dim] */
int  feddb200_assemble_linelas_d(feddb200_ctx *ctx, const feddb200_pat *pat, double lambda, double mu, double *values_d);
/* FE::assemblyAdvectionVecField (core/FE/FE_def.hpp:1685-1836) [BLOCK_DIAG]; u_rep_d is the
 * repeated, node-wise interleaved velocity u[dim*LID+d] (:1785,1800) */
int  feddb200_assemble_advection_d(feddb200_ctx *ctx, const feddb200_pat *pat, const double *u_rep_d, double *values_d);
/* FE::assemblyAdvectionInUVecField (core/FE/FE_def.hpp:1839-1929) [BLOCK_FULL, dim x dim] */
int  feddb200_assemble_advection_in_u_d(feddb200_ctx *ctx, const feddb200_pat *pat, const double *u_rep_d, double *values_d);
/* FE::assemblyDivAndDivT / assemblyDivAndDivTFast (core/FE/FE_def.hpp:1932-2057, 2061-2148):
 * patB has pressure rows x velocity cols [BLOCK_FULL 1 x dim], patBT velocity rows x pressure
 * cols [BLOCK_FULL dim x 1]; positive sign, callers scale by -1 (NavierStokes_def.hpp:217-218).
 * Either output may be NULL. */
int  feddb200_assemble_div_divT_d(feddb200_ctx *ctx, const feddb200_pat *patB, const feddb200_pat *patBT,
                                  double *valuesB_d, double *valuesBT_d);
/* Fused (0,0) block of the Navier-Stokes system, the three calls of
 * NavierStokes::reAssemble + assembleConstantMatrices (problems/specific/NavierStokes_def.hpp:
 * 140-152, 297-313) in one pass on the union (BLOCK_FULL) pattern:
 *   values = rho*nu*LaplaceVecField + rho*N(u) [+ rho*W(u) if newton] */
int  feddb200_assemble_ns_jacobian_d(feddb200_ctx *ctx, const feddb200_pat *pat, double rho, double nu,
                                     const double *u_rep_d, int newton, double *values_d);

/* host-pointer forms of the same calls (what the glue uses when the matrix lives on the
 * host): H2D of u, D2H of the values inside the call */
int  feddb200_assemble_laplace(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, double *values);
int  feddb200_assemble_mass(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, double *values);
int  feddb200_assemble_bdstab(feddb200_ctx *ctx, const feddb200_pat *pat, double *values);
int  feddb200_assemble_linelas(feddb200_ctx *ctx, const feddb200_pat *pat, double lambda, double mu, double *values);
int  feddb200_assemble_advection(feddb200_ctx *ctx, const feddb200_pat *pat, const double *u_rep, double *values);
int  feddb200_assemble_advection_in_u(feddb200_ctx *ctx, const feddb200_pat *pat, const double *u_rep, double *values);
int  feddb200_assemble_div_divT(feddb200_ctx *ctx, const feddb200_pat *patB, const feddb200_pat *patBT,
                                double *valuesB, double *valuesBT);
int  feddb200_assemble_ns_jacobian(feddb200_ctx *ctx, const feddb200_pat *pat, double rho, double nu,
                                   const double *u_rep, int newton, double *values);

/* ---- ghost-row exchange helpers (the Tpetra Export/ADD step of fillComplete) -------------
 * After an assembly the values of ghost rows sit behind the owned rows in `values_d`
 * (offset feddb200_pattern_nnz of the owned part).  The host ships them to the owners
 * (NCCL) and calls unpack_add on the receiving side:  values_d[slot[k]] += recv_d[k]. */
int64_t feddb200_pattern_nnz_owned(const feddb200_pat *pat, int row_dofs, int col_dofs, int block_mode);
int  feddb200_unpack_add_d(feddb200_ctx *ctx, double *values_d, const double *recv_d, const int64_t *slot_d, int64_t n);
/* Fused variant of the same step over NVLink peer memory (one process per GPU, CUDA IPC): every rank allocates its
 * receive buffer with ipc_alloc, publishes the 64-byte handle, and opens the buffers of its peers.  With ghost targets
 * set, the row-gather kernels of the Laplace / elasticity operators write the ghost rows -- the values at offsets
 * [seg_begin[o], seg_begin[o+1]) of the values array, o = 0..nseg-1 (destination ranks in rank order, nseg <= 8) --
 * directly to seg_ptr_d[o] + (offset - seg_begin[o]) as they are computed (TMA bulk stores / coalesced stores to the
 * peer mapping) instead of behind the owned rows; no send buffer, no collective moves data.  The owner adds the
 * received values with unpack_add after a cross-rank barrier.  nseg = 0 switches the targets off.  The other
 * operators and the atomic / coloured modes keep the NCCL path (they ignore the targets and write behind the owned rows). */
int  feddb200_ipc_alloc(feddb200_ctx *ctx, int64_t bytes, void **ptr_d, unsigned char *handle64);
int  feddb200_ipc_open(feddb200_ctx *ctx, const unsigned char *handle64, void **ptr_d);
int  feddb200_ipc_close(feddb200_ctx *ctx, void *ptr_d);
int  feddb200_set_ghost_targets(feddb200_ctx *ctx, int nseg, const int64_t *seg_begin, void *const *seg_ptr_d);
/* FE::assemblyRHS (core/FE/FE_def.hpp:4694-4766) for the constant source the reference
 * supports ("for now just const", :4730): value_func[dim] = the host callback's result (evaluated once by the
 * caller, as :4735 does), deg_func = its declared polynomial degree (last entry of funcParameter, :4716).  Output: the
 * load vector in pattern-row order (one rank: repeated-node order; owned rows first, ghost rows behind, to be shipped
 * like ghost matrix rows = the reference's exportFromVector(..., "Add"), Problem_def.hpp:213), dofs = 1 ("Scalar",
 * vec_field = 0) or dim ("Vector"), node-wise interleaved.  The vector is overwritten. */
int  feddb200_assemble_rhs_d(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, int deg_func,
                             const double *value_func, double *rhs_d);
int  feddb200_assemble_rhs(feddb200_ctx *ctx, const feddb200_pat *pat, int vec_field, int deg_func,
                           const double *value_func, double *rhs);

/* BCBuilder::setDirichletBC -> setLocalRowOne / setLocalRowZero (core/General/BCBuilder_def.hpp:618-709; SURVEY.md
 * 8(f) rank 1) on resident values: node_mask_d[I] (one byte per OWNED row node, device) has bit a set when dof a of
 * node I carries a Dirichlet condition ("Dirichlet" = all dofs, "Dirichlet_X" = bit 0, "Dirichlet_X_Z" = bits 0|2, ...).
 * Every selected dof row is zeroed; on a diagonal block (diagonal_block != 0) its diagonal entry becomes 1. */
int  feddb200_set_dirichlet_rows_d(feddb200_ctx *ctx, const feddb200_pat *pat, int row_dofs, int col_dofs, int block_mode,
                                   const uint8_t *node_mask_d, int diagonal_block, double *values_d);
/* BCBuilder::setRHS, Dirichlet part (core/General/BCBuilder_def.hpp:93-166) on a resident right-hand side: rhs_d[dofs * I + a] =
 * bc_values_d[dofs * I + a] for every owned node I and dof a with bit a of node_mask_d[I] set.  bc_values holds the boundary
 * function evaluated by the host glue at the node's coordinates (the reference calls the user's boost::function per node, :134);
 * entries of dofs without a condition are ignored. */
int  feddb200_set_dirichlet_rhs_d(feddb200_ctx *ctx, int64_t n_nodes, int dofs, const uint8_t *node_mask_d, const double *bc_values_d,
                                  double *rhs_d);
/* Matrix::scale (core/LinearAlgebra/Matrix_def.hpp:257) on resident values */
int  feddb200_scale_d(feddb200_ctx *ctx, double *values_d, int64_t n, double alpha);

/* ---- CSR algebra on resident matrices ---------------------------------------------------------
 * Device CSR arrays (int64 rowptr, int32 column indices ascending per row, f64 values), e.g. from pattern_expand +
 * assemble_*_d.  Two phases each: symbolic fills rowptr of the result and returns its nnz, the caller allocates, numeric
 * fills column indices and values.  Explicit zeros are kept.
 *
 * Matrix::addMatrix (core/LinearAlgebra/Matrix_def.hpp:281-287 -> Xpetra TwoMatrixAdd; callers NavierStokes_def.hpp:303-304,
 * 312-313):  C = alpha*A + beta*B on the union of the two patterns (same row and column index spaces). */
int  feddb200_csr_add_symbolic_d(feddb200_ctx *ctx, int64_t n_rows, const int64_t *rowptrA_d, const int32_t *colindA_d,
                                 const int64_t *rowptrB_d, const int32_t *colindB_d, int64_t *rowptrC_d, int64_t *nnzC);
int  feddb200_csr_add_numeric_d(feddb200_ctx *ctx, int64_t n_rows, double alpha, const int64_t *rowptrA_d,
                                const int32_t *colindA_d, const double *valuesA_d, double beta, const int64_t *rowptrB_d,
                                const int32_t *colindB_d, const double *valuesB_d, const int64_t *rowptrC_d,
                                int32_t *colindC_d, double *valuesC_d);
/* BlockMatrix::merge / mergeBlockNew (core/LinearAlgebra/BlockMatrix_def.hpp:119-289): the monolithic matrix of an
 * nb x nb block system (nb <= 4).  Block (i, j) = rowptr_d[i*nb + j] (NULL: absent block), its rows offset by the row
 * counts of the block rows above, its column indices by the column counts n_cols[0..j-1] of the block columns to its
 * left (determineLocalOffsets, :170-209); the pointer arrays are HOST arrays of device pointers. */
int  feddb200_block_merge_symbolic_d(feddb200_ctx *ctx, int nb, const int64_t *n_rows, const int32_t *n_cols,
                                     const int64_t *const *rowptr_d, int64_t *rowptrM_d, int64_t *nnzM);
int  feddb200_block_merge_numeric_d(feddb200_ctx *ctx, int nb, const int64_t *n_rows, const int32_t *n_cols,
                                    const int64_t *const *rowptr_d, const int32_t *const *colind_d,
                                    const double *const *values_d, const int64_t *rowptrM_d, int32_t *colindM_d,
                                    double *valuesM_d);

#ifdef __cplusplus
}
#endif
#endif /* FEDDB200_H */
