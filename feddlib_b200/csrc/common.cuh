// Internal definitions shared by the translation units of libfeddb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/feddb200.h"

namespace fb {

void set_error(const std::string &msg);

#define FB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            fb::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ +   \
                          ":" + std::to_string(__LINE__) + ")");                                   \
            return FEDDB200_ERUNTIME;                                                              \
        }                                                                                          \
    } while (0)

#define FB_LOGIC(cond, msg)                                                                        \
    do {                                                                                           \
        if (cond) {                                                                                \
            fb::set_error(msg);                                                                    \
            return FEDDB200_ELOGIC;                                                                \
        }                                                                                          \
    } while (0)

// operators of the hot path
enum Op { OP_LAP = 0, OP_ELAS = 1, OP_ADV = 2, OP_ADVU = 3, OP_B = 4, OP_BT = 5, OP_NSJ = 6, OP_MASS = 7, OP_COUNT = 8 };

// Reference-element tables of one operator (restated from the FE type and quadrature degree,
// see tables.cu): weights, gradients of the "velocity" space, values of the "value" space.
constexpr int MAXQ = 15;
constexpr int MAXN = 10;
struct OpTables {
    int nq, nv, np, dim;
    double w[MAXQ];
    double dphi[MAXQ * MAXN * 3]; // [q][i][c], stride nv*dim
    double phi[MAXQ * MAXN];      // [q][i],   stride np
    double lam[MAXQ * 4];         // barycentric coordinates of the quadrature points
};

struct Bucket {            // rows of one type with node-row length <= lcap, processed by one gather launch
    int type;              // 0 vertex-node rows, 1 ring-ordered edge-node rows (3D P2), 2 edge-node rows (generic)
    int ghost;             // 1: rows owned by another rank (index >= n_owned); their values are shipped after the assembly
    int lcap;
    int64_t start, count;  // range in row_perm
    // block-task kernel (star_kernels.cuh: k_task): tiles of this bucket in the pattern's tile array; count 0: not used
    int64_t tile_start = 0, tile_count = 0;
    int npt = 0;           // row nodes per tile at most
    int task_max_tets = 0, task_max_passes = 0;   // per tile, over the bucket's tiles
    // fan kernel (star_kernels.cuh: k_fan): the bucket's padded records start at fan_off (in records) of the pattern's
    // fan-record array, fan_W per row; fan_W 0: not used
    int64_t fan_off = 0;
    int fan_W = 0, fan_npw = 0;
    int in_star = 0;       // the bucket's rows are tiles of the star kernel's address-ordered tile list (k_star)
};

} // namespace fb

struct feddb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    // side streams + events: independent bucket launches of one assembly run concurrently (fork after the
    // geometry pre-pass, join before the call returns to the caller's stream)
    static constexpr int kSide = 3;
    cudaStream_t side[kSide] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {nullptr, nullptr, nullptr};
    int mode = FEDDB200_SCATTER_GATHER;
    int row_phase = FEDDB200_ROWS_ALL;
    int64_t launches = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    fb::OpTables *tab_d = nullptr; // device copies of the operator tables, one slot per operator
    int tab_key[fb::OP_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
    void *scratch_d[2] = {nullptr, nullptr}; // grow-only device scratch of the host-pointer entry points
    // ghost-row targets of the next gather assemblies (feddb200_set_ghost_targets); n_ghost_seg == 0: none
    int n_ghost_seg = 0;
    int64_t ghost_seg_begin[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double *ghost_seg_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[2] = {0, 0};
    // kernels whose shared-memory attribute has been raised / whose occupancy has been queried on this context (per call
    // these driver queries cost more than a small assembly)
    struct OccEntry { const void *f; int nt; size_t smem; int per_sm; };
    std::vector<OccEntry> occ_cache;
};

struct feddb200_mesh {
    feddb200_ctx *ctx = nullptr;
    int device = 0;              // device ordinal (copied: the mesh may be freed after its context)
    int dim = 0, nloc = 0;
    int64_t ne = 0, nn = 0;
    int32_t *conn_d = nullptr;
    double *coords_d = nullptr;
    std::vector<int32_t> conn_h; // kept for the host-side greedy element colouring
};

struct feddb200_pat {
    feddb200_ctx *ctx = nullptr;
    int device = 0;              // device ordinal (copied: the pattern may be freed after its context)
    const feddb200_mesh *rm = nullptr, *cm = nullptr;
    int64_t n_rows = 0, n_owned = 0, n_cols = 0, nnz = 0, nnz_owned = 0;
    int max_len = 0;
    int pos_stride = 0;
    int64_t *rowptr_d = nullptr;  // [n_rows+1]  node-level
    int32_t *colind_d = nullptr;  // [nnz]
    int32_t *row_lid_d = nullptr; // [rm->nn] or null (identity)
    uint16_t *pos_d = nullptr;    // [ne*nr][pos_stride] position of col node j in the row of node i
    // row -> (element, local index) incidences, rows ascending, elements ascending within a row
    int64_t n_inc = 0;
    int64_t *inc_ptr_d = nullptr; // [n_rows+1]
    int32_t *inc_d = nullptr;     // [n_inc]  (e << 4) | i
    int32_t *row_perm_d = nullptr; // rows ordered by bucket (built lazily with the gather maps)
    uint32_t *rec_d = nullptr;     // [n_inc][rec_words] per-incidence gather records (positions | element | permutation)
    void *rowinfo_d = nullptr;     // [n_rows] RowInfo records in bucket order
    uint32_t *ahead_d = nullptr;   // [n_inc] 3D P2: element of the incidence kRsAhead places further along the same row (k_gather_s)
    int rec_words = 0;
    void *task_tiles_d = nullptr;  // [n_tiles] TaskTile: tiles of the block-task kernel (3D P2 vertex-node rows)
    uint64_t *tasks_d = nullptr;   // task programs of the tiles (tasks.cuh)
    void *star_tiles_d[2] = {nullptr, nullptr};   // address-ordered tile lists of k_star: owned rows, ghost rows
    int64_t star_n[2] = {0, 0};
    uint32_t *fanrec_d = nullptr;  // padded per-bucket records of the fan kernel (3D P2 ring-ordered edge-node rows)
    void *tiletet_d = nullptr;     // [n_tiles] tile blocks (400 bytes: header, row nodes, incident elements; star_kernels.cuh)
    double *geom_d = nullptr;      // [ne][GS] per-element geometry cache, recomputed by every assembly
    // boundary-sector fragments of the row-gather path (kernels.cuh: "fragment protocol"): two 32-byte slots per node row
    // (head, tail); null if some node row is too short for the protocol
    double *frag_d = nullptr;
    // points path of the scalar Laplace rows (3D P2, kernels.cuh: geo_from_points): canonical vertex ids per incidence (built
    // at the first Laplace assembly of the pattern) and the padded coordinates (rewritten by every such assembly)
    void *vtx_d = nullptr;         // uint4 [n_inc]
    double *coords4_d = nullptr;   // [rm->nn][4]
    double *sloc_d = nullptr;      // [ne][nloc][nloc] scalar local matrices of N(u) / the fused Navier-Stokes block (k_sloc)
    double *sloc_tab_d = nullptr;  // coefficient tensors of k_sloc (uploaded once per pattern)
    double *dt_d = nullptr;        // [ne][dim][dim][4] |det| * grad u at the element's vertices
    bool gather_ready = false;
    std::vector<fb::Bucket> buckets;
    std::vector<int> bucket_order;   // launch order of the buckets (largest first), fixed with the gather maps
    // element colouring (lazy)
    int n_colours = 0;
    std::vector<int64_t> colour_ptr;
    int32_t *colour_perm_d = nullptr;
    std::vector<int64_t> rowptr_h;
};

namespace fb {
int build_tables(OpTables &t, int op, int dim, int nloc_v, int nloc_p);
int rhs_coefficients(int dim, int nloc, int deg_func, double *c);
int ensure_colouring(feddb200_pat *pat);
} // namespace fb
