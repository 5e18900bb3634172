/* feddb200_halo.h -- multi-rank host plan of the assembly engine (C ABI, host code only).
 *
 * What it replaces in the reference: with more than one rank every rank assembles its own elements; entries of rows
 * owned by another rank are shipped to the owner and ADDed inside Matrix::fillComplete
 * (feddlib/core/LinearAlgebra/Matrix_def.hpp:192-199 -> Tpetra globalAssemble), the rows live on the unique map built
 * from the repeated map (feddlib/core/LinearAlgebra/Map_def.hpp:184-263) and the column map is "owned GIDs in
 * domain-map order, then remote GIDs grouped by owning rank, ascending GID" (SURVEY.md Appendix C).
 *
 * The plan is built once per (mesh partition, node pattern) from exactly what a Domain gives: the global ids of the
 * repeated nodes and the rank that owns each of them.  Communication goes through two caller-supplied callbacks, so an
 * MPI build (MPI_Alltoallv on the Teuchos communicator), a torch.distributed build and the in-process test communicator
 * plug in without this library linking any of them.  The node pattern of the rank's elements is produced by a third
 * callback (feddb200_pattern_build on the device in production, a host builder in the CPU tests).
 */
#ifndef FEDDB200_HALO_H
#define FEDDB200_HALO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* communicator: all-to-all-v of 64-bit words (int64 payloads; doubles travel bit-wise).  The callee returns the received
 * words of rank s in [recv_offsets[s], recv_offsets[s+1]) of a buffer it owns until the next call on the same comm. */
typedef struct feddb200_comm {
    void *user;
    int rank, size;
    int (*alltoallv64)(void *user, const int64_t *send, const int64_t *send_counts /*[size]*/, const int64_t **recv,
                       int64_t *recv_counts /*[size]*/);
} feddb200_comm;

/* node pattern of this rank's elements for a given row / column numbering (the semantics of feddb200_pattern_build):
 * rows row_lid[node] (n_rows, the first n_owned owned), columns col_lid[node] (NULL: repeated ids, n_cols = nodes) plus
 * the extra entries (row, col); returns rowptr[n_rows+1] / colind[nnz], valid until the next call with the same user -- the
 * arrays of the LAST call (the final pattern) are referenced by the plan, not copied, and must outlive it. */
typedef int (*feddb200_node_pattern_fn)(void *user, const int32_t *row_lid, int64_t n_rows, int64_t n_owned, const int32_t *col_lid,
                                        int64_t n_cols, const int32_t *extra_row, const int32_t *extra_col, int64_t n_extra,
                                        const int64_t **rowptr, const int32_t **colind);

typedef struct feddb200_halo feddb200_halo;

/* gid_rep[nn]: global id of every repeated node (Domain::getMapRepeated); owner[nn]: rank owning it (the unique map:
 * Map::buildUniqueMap, Map_def.hpp:184-212).  Collective over the communicator. */
int  feddb200_halo_create(feddb200_halo **plan, const feddb200_comm *comm, int64_t nn, const int64_t *gid_rep, const int32_t *owner,
                          feddb200_node_pattern_fn pattern_fn, void *pattern_user);
void feddb200_halo_free(feddb200_halo *plan);

/* sizes: n_owned, n_ghost, n_rows, n_colmap (Tpetra column map), n_cols (column map + ghost-only columns), n_extra,
 * nnz_owned_nodes, nnz_nodes, n_recv (received node entries) */
int  feddb200_halo_sizes(const feddb200_halo *plan, int64_t *n_owned, int64_t *n_ghost, int64_t *n_rows, int64_t *n_colmap,
                         int64_t *n_cols, int64_t *n_extra, int64_t *nnz_owned_nodes, int64_t *nnz_nodes, int64_t *n_recv);
/* arrays (owned by the plan): which = 0 row_lid[nn] i32, 1 col_lid[nn] i32, 2 extra_row i32, 3 extra_col i32,
 * 4 colmap_gids[n_colmap] i64, 5 unique_gids[n_owned] i64, 6 ghost_row_gids[n_ghost] i64, 7 ghost_row_owner[n_ghost] i64,
 * 8 rowptr[n_rows+1] i64, 9 colind[nnz_nodes] i32, 10 send_counts_nodes[size] i64, 11 recv_counts_nodes[size] i64,
 * 12 recv_row i64, 13 recv_pos i64, 14 recv_len_sender i64, 15 recv_q i64 (12-15: [n_recv]),
 * 16 import_send_rows i64 (owned rows the other ranks hold as ghosts, grouped by requesting rank, in their ghost order),
 * 17 import_send_counts[size] i64, 18 rep_of_row[n_rows] i64 (repeated node of every pattern row) */
const void *feddb200_halo_array(const feddb200_halo *plan, int which, int64_t *count);

/* value slots (into this rank's dof-level CSR values) of every received value, in arrival order; block_mode as in
 * feddb200.h (0 scalar, 1 block-diagonal, 2 full blocks).  slots must hold factor * n_recv entries. */
int  feddb200_halo_recv_slots(const feddb200_halo *plan, int row_dofs, int col_dofs, int block_mode, int64_t *slots);
/* values per rank sent / received for one matrix of that layout */
int  feddb200_halo_split_sizes(const feddb200_halo *plan, int row_dofs, int col_dofs, int block_mode, int64_t *send_counts, int64_t *recv_counts);

/* Host-side globalAssemble of one matrix: values[nnz] holds this rank's assembled values (owned rows first, ghost rows
 * behind, as the assembly entry points return them); the ghost part is shipped to the owners through the communicator
 * and added into the owned part (fixed order: sender by sender, arrival order -> deterministic sums). */
int  feddb200_halo_export_add(const feddb200_halo *plan, const feddb200_comm *comm, int row_dofs, int col_dofs, int block_mode,
                              int64_t nnz_owned_values, double *values);

/* Import of a node-wise interleaved vector from the unique to the repeated map (the reference:
 * MultiVector::importFromVector, feddlib/core/LinearAlgebra/MultiVector_def.hpp:258-294, called on the velocity before the
 * advection assemblies, problems/specific/NavierStokes_def.hpp:294): u_unique[dofs * n_owned] in unique-map order ->
 * u_rep[dofs * nn] in repeated order.  Collective. */
int  feddb200_halo_import_vector(const feddb200_halo *plan, const feddb200_comm *comm, int dofs, const double *u_unique, double *u_rep);

#ifdef __cplusplus
}
#endif
#endif
