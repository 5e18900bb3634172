// Star kernels of the row-gather path (3D P2): the lanes of a warp work on the elements around ONE row node (or a few)
// at the same time, instead of one thread walking them one after the other (kernels.cuh: k_ring, k_gather_s).
//
//  * k_task -- vertex-node rows.  Block-task kernel: the warp stages, per incident element, the row gradient and the ten
//    column vectors  q_jc = sum_t R_jc^t G_sv(jc,t)  in shared memory (one lane per element), then every lane runs a TASK of
//    the tile's program (tasks.cuh): it sums  X += (|det| G_0) (x) q_jc  over up to four (element, column) pairs of ONE node
//    block in registers and stores the finished 3x3 block  mu (tr X I + X^T) + lambda X  into the node's shared-memory row.
//    No read-modify-write, no atomics, fixed summation order; the row leaves by one TMA bulk store per node.
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "tasks.cuh"

namespace fb {

// tile block of the block-task kernel (400 bytes = 25 chunks of 16 bytes, one contiguous record per tile):
//   chunk 0      header: n_nodes | n_tets << 8 | n_passes << 16, first task (in units of 32 task words), 0, 0
//   chunks 1-8   row nodes: rowptr of the node (int64), row length (u32), flags (bit 0: holes)
//   chunks 9-24  incident elements, two per chunk: (element, canonical permutation)
constexpr int kTileBlkB = 400, kTileBlkChunks = 25;

// one-time: task programs.  pass 1 (count): passes per tile; pass 2 (emit): task words + the tile blocks
struct TaskBuildArgs {
    const RowInfo *rowinfo;      // bucket-ordered row records
    const uint32_t *rec;         // incidence records (8 words)
    TaskTile *tiles;             // q0, n_nodes filled by the host; n_tets, n_passes by pass 1; task_off by the host before pass 2
    int64_t n_tiles;
    uint64_t *tasks;             // pass 2
    uint32_t *tileblk;           // pass 2: [tile][100 words]
    int *status;                 // set to 1 if a tile violates the format limits
};

__global__ void k_task_build(const TaskBuildArgs A, int emit)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < A.n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
        TaskTile &T = A.tiles[t];
        int len[kTaskMaxNodes], ninc[kTaskMaxNodes];
        int64_t k0[kTaskMaxNodes];
        int ntet = 0;
        for (int i = 0; i < T.n_nodes; i++) {
            const RowInfo &R = A.rowinfo[T.q0 + i];
            len[i] = R.len; ninc[i] = R.ninc; k0[i] = R.k0;
            ntet += R.ninc;
        }
        const int np = schedule_tile(T.n_nodes, len, ninc, k0, A.rec, 8, 10, emit ? A.tasks + (int64_t)T.task_off * 32 : nullptr);
        if (np < 0 || np > 255) { *A.status = 1; continue; }
        if (!emit) { T.n_tets = (uint8_t)ntet; T.n_passes = (uint8_t)np; continue; }
        uint32_t *B = A.tileblk + t * (kTileBlkB / 4);
        for (int x = 0; x < kTileBlkB / 4; x++) B[x] = 0;
        B[0] = (uint32_t)T.n_nodes | (uint32_t)ntet << 8 | (uint32_t)np << 16;
        B[1] = T.task_off;
        int m = 0;
        for (int i = 0; i < T.n_nodes; i++) {
            const RowInfo &R = A.rowinfo[T.q0 + i];
            B[4 + 4 * i] = (uint32_t)((uint64_t)R.base & 0xffffffffu);
            B[5 + 4 * i] = (uint32_t)((uint64_t)R.base >> 32);
            B[6 + 4 * i] = (uint32_t)R.len;
            B[7 + 4 * i] = ((R.pad & 16) ? 1u : 0u) | ((R.pad & 32) ? 2u : 0u) | (uint32_t)((uint64_t)R.pad >> 32) << 2;   // holes | tail plain | row
            for (int j = 0; j < ninc[i]; j++, m++) {
                const uint32_t *w = A.rec + (k0[i] + j) * 8;
                B[36 + 2 * m] = w[5];
                B[37 + 2 * m] = w[6] & 0xffu;
            }
        }
    }
}

struct TaskArgs {
    GatherArgs G;                // geom, values, c0/c1 (lambda, mu), R, pitch, vec_dim, ghost segments
    const uint4 *tileblk;        // tile blocks of this launch
    int64_t n_tiles;
    const uint64_t *tasks;
    int npt;                     // row nodes per tile at most (shared-memory rows per warp)
    int max_tets, max_passes;    // per tile, over the tiles of this launch (sizes of the staging areas)
};

constexpr int kTaskVec = 11;     // staged vectors per element: |det| G_0, q_0 .. q_9

// bytes of shared memory per warp
__host__ __device__ inline size_t task_warp_bytes(int npt, int pitch, int max_tets, int max_passes)
{
    return (size_t)npt * pitch * 8 + (size_t)max_tets * kTaskVec * 24 + 2 * (size_t)kTileBlkChunks * 16 + (size_t)max_tets * 128 +
           2 * (size_t)max_passes * 256;
}

// OPG 0: Laplace (one value per node block; vec_dim: replicated on the block diagonal), OPG 1: elasticity (3x3 blocks).
// Inputs are streamed by cp.async: the tile block two tiles ahead, the geometry lines (canonical vertex order) and the task
// words one tile ahead.
#ifndef FB_TASK_MINBLOCKS
#define FB_TASK_MINBLOCKS 6   // shared memory allows six blocks per SM for the L = 65 rows: registers up to 170 are free
#endif
template <int OPG>
__global__ void __launch_bounds__(64, FB_TASK_MINBLOCKS) k_task(const TaskArgs A)
{
    constexpr int DIM = 3, NVTX = 4, NL = 10;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int pitch = A.G.pitch, MT = A.max_tets;                  // pitch: doubles per node row area (even)
    char *const wbase = reinterpret_cast<char *>(smem) + (size_t)wib * task_warp_bytes(A.npt, pitch, MT, A.max_passes);
    double *const rows = reinterpret_cast<double *>(wbase);       // every area is a multiple of 16 bytes
    double2 *const XY = reinterpret_cast<double2 *>(rows + (size_t)A.npt * pitch);
    double *const Z = reinterpret_cast<double *>(XY + MT * kTaskVec);
    const uint4 *const blk = reinterpret_cast<const uint4 *>(Z + MT * kTaskVec);              // [2][25]
    const uint4 *const geo = blk + 2 * kTileBlkChunks;                                         // [8][MT] chunk-major
    const uint64_t *const tsk = reinterpret_cast<const uint64_t *>(geo + 8 * MT);              // [2][max_passes][32]
    const uint32_t blk_s = (uint32_t)__cvta_generic_to_shared(blk), geo_s = (uint32_t)__cvta_generic_to_shared(geo),
                   tsk_s = (uint32_t)__cvta_generic_to_shared(tsk);
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;
    const int64_t step = (int64_t)gridDim.x * wpb;
    int64_t tile = (int64_t)blockIdx.x * wpb + wib;
    if (tile >= A.n_tiles) return;

    auto request_blk = [&](int64_t t, int b) {
        if (lane < kTileBlkChunks) cp_async16_s(blk_s + (b * kTileBlkChunks + lane) * 16, A.tileblk + t * kTileBlkChunks + lane);
    };
    auto request_geo_tasks = [&](int b) {   // of the tile whose block is in buffer b
        const uint4 h = blk[b * kTileBlkChunks];
        const int nt = (h.x >> 8) & 0xff, np = (h.x >> 16) & 0xff;
        if (lane < nt) {
            const uint2 et = reinterpret_cast<const uint2 *>(blk + b * kTileBlkChunks + 9)[lane];
            const char *g = reinterpret_cast<const char *>(A.G.geom) + (int64_t)et.x * 128;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const char *gv = g + 32 * ((et.y >> (2 * v)) & 3);
                cp_async16_s(geo_s + ((2 * v) * MT + lane) * 16, gv);
                cp_async16_s(geo_s + ((2 * v + 1) * MT + lane) * 16, gv + 16);
            }
        }
        const char *tp = reinterpret_cast<const char *>(A.tasks + (int64_t)h.y * 32);
        for (int x = lane; x < np * 16; x += 32) cp_async16_s(tsk_s + (b * A.max_passes * 16 + x) * 16, tp + x * 16);
    };
    request_blk(tile, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    request_geo_tasks(0);
    if (tile + step < A.n_tiles) request_blk(tile + step, 1);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    int buf = 0;

    for (;;) {
        const int64_t tile_n = tile + step;
        const bool more = tile_n < A.n_tiles;
        const uint4 hdr = blk[buf * kTileBlkChunks];
        const int n_nodes = hdr.x & 0xff, n_tets = (hdr.x >> 8) & 0xff, n_passes = (hdr.x >> 16) & 0xff;
        // row nodes of the tile (lane = node slot)
        int64_t base = 0;
        int L = 0;
        uint32_t nflags = 0;
        if (lane < n_nodes) {
            const uint4 nd = blk[buf * kTileBlkChunks + 1 + lane];
            base = (int64_t)((uint64_t)nd.x | (uint64_t)nd.y << 32);
            L = (int)nd.z;
            nflags = nd.w;
        }
        // stage: lane m = element slot m.  vector 0: |det| G_0 (row function: canonical vertex 0), vector 1 + jc:
        // q_jc = R_jc^0 G_sv(jc,0) + R_jc^1 G_sv(jc,1)
        if (lane < n_tets) {
            double G[4][3], adet = 0.0;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 a = geo[(2 * v) * MT + lane], b = geo[(2 * v + 1) * MT + lane];
                G[v][0] = __hiloint2double((int)a.y, (int)a.x);
                G[v][1] = __hiloint2double((int)a.w, (int)a.z);
                G[v][2] = __hiloint2double((int)b.y, (int)b.x);
                if (v == 0) adet = __hiloint2double((int)b.w, (int)b.z);
            }
            XY[lane * kTaskVec] = make_double2(adet * G[0][0], adet * G[0][1]);
            Z[lane * kTaskVec] = adet * G[0][2];
#pragma unroll
            for (int jc = 0; jc < NL; jc++) {
                const double r0 = A.G.R.r[0][jc][0][0];
                double q[3];
#pragma unroll
                for (int d = 0; d < 3; d++) q[d] = r0 * G[canon_sv<DIM>(jc, 0)][d];
                if (jc >= NVTX) {
                    const double r1 = A.G.R.r[0][jc][0][1];
#pragma unroll
                    for (int d = 0; d < 3; d++) q[d] = fma(r1, G[canon_sv<DIM>(jc, 1)][d], q[d]);
                }
                XY[lane * kTaskVec + 1 + jc] = make_double2(q[0], q[1]);
                Z[lane * kTaskVec + 1 + jc] = q[2];
            }
        }
        // the geometry buffer is consumed: request the next tile's geometry and task words (its block arrived a tile ago)
        // and the block of the tile after it; they land while this tile is processed
        __syncwarp();
        if (more) request_geo_tasks(buf ^ 1);
        if (tile_n + step < A.n_tiles) request_blk(tile_n + step, buf);
        cp_async_commit();

        // node row of slot s: `pitch` doubles apart, shifted by one double where that gives the row the 16-byte phase of
        // its destination (the TMA bulk store needs both sides 16-byte aligned)
        const int n = NB * L;                                         // values per dof row (lane = node slot)
        const int64_t off_node = (int64_t)TPR * NB * nrep * base;
        double *const outp = lane < n_nodes ? out_ptr(A.G, off_node) : nullptr;
        const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
        const int my_rowoff = lane * pitch + head;                    // lane = node slot; pitch is even

#pragma unroll 1
        for (int pass = 0; pass < n_passes; pass++) {
            const uint64_t task = tsk[(buf * A.max_passes + pass) * 32 + lane];
            const int np = (int)((task >> 36) & 7);
            const int nsteps = __reduce_max_sync(FULL, np);
            double X[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) X[i] = 0.0;
#ifndef FB_TASK_BATCHED_LOADS   // one pair per loop iteration (default)
#pragma unroll 1
            for (int s = 0; s < nsteps; s++) {
                if (s < np) {
                    const uint32_t pr = (uint32_t)(task >> (9 * s)) & 0x1ffu;
                    const int m = pr & 31, jc = pr >> 5;
                    const double2 gxy = XY[m * kTaskVec], qxy = XY[m * kTaskVec + 1 + jc];
                    const double gz = Z[m * kTaskVec], qz = Z[m * kTaskVec + 1 + jc];
                    if constexpr (OPG == 1) {
                        X[0] = fma(gxy.x, qxy.x, X[0]); X[1] = fma(gxy.x, qxy.y, X[1]); X[2] = fma(gxy.x, qz, X[2]);
                        X[3] = fma(gxy.y, qxy.x, X[3]); X[4] = fma(gxy.y, qxy.y, X[4]); X[5] = fma(gxy.y, qz, X[5]);
                        X[6] = fma(gz, qxy.x, X[6]);    X[7] = fma(gz, qxy.y, X[7]);    X[8] = fma(gz, qz, X[8]);
                    } else {
                        X[0] = fma(gxy.x, qxy.x, X[0]); X[0] = fma(gxy.y, qxy.y, X[0]); X[0] = fma(gz, qz, X[0]);
                    }
                }
            }
#else
            // FB_TASK_BATCHED_LOADS: all operand loads of the pass first (up to kTaskQ pairs, warp-uniform step count), then the
            // multiply-adds -- one shared-memory latency per pass instead of one per pair.  Measured equal (2.469 vs 2.467 ms
            // per assembly), so the simpler loop stays.
            double2 gxy[kTaskQ], qxy[kTaskQ];
            double gz[kTaskQ], qz[kTaskQ];
#pragma unroll
            for (int s = 0; s < kTaskQ; s++) {
                gxy[s] = qxy[s] = make_double2(0.0, 0.0);
                gz[s] = qz[s] = 0.0;
                if (s < nsteps && s < np) {
                    const uint32_t pr = (uint32_t)(task >> (9 * s)) & 0x1ffu;
                    const int m = pr & 31, jc = pr >> 5;
                    gxy[s] = XY[m * kTaskVec]; qxy[s] = XY[m * kTaskVec + 1 + jc];
                    gz[s] = Z[m * kTaskVec]; qz[s] = Z[m * kTaskVec + 1 + jc];
                }
            }
#pragma unroll
            for (int s = 0; s < kTaskQ; s++) {
                if (s < nsteps) {
                    if constexpr (OPG == 1) {
                        X[0] = fma(gxy[s].x, qxy[s].x, X[0]); X[1] = fma(gxy[s].x, qxy[s].y, X[1]); X[2] = fma(gxy[s].x, qz[s], X[2]);
                        X[3] = fma(gxy[s].y, qxy[s].x, X[3]); X[4] = fma(gxy[s].y, qxy[s].y, X[4]); X[5] = fma(gxy[s].y, qz[s], X[5]);
                        X[6] = fma(gz[s], qxy[s].x, X[6]);    X[7] = fma(gz[s], qxy[s].y, X[7]);    X[8] = fma(gz[s], qz[s], X[8]);
                    } else {
                        X[0] = fma(gxy[s].x, qxy[s].x, X[0]); X[0] = fma(gxy[s].y, qxy[s].y, X[0]); X[0] = fma(gz[s], qz[s], X[0]);
                    }
                }
            }
#endif
            // positions with more than four pairs: the partial sums of the group's lanes, fixed order; only the levels the
            // largest group of the pass needs (a ring-of-6 column split in two needs one, the row node's own column three)
            const int rem = (int)((task >> 50) & 7);
            const int maxrem = __reduce_max_sync(FULL, rem);
#pragma unroll
            for (int d = 1; d <= 4; d <<= 1) {
                if (d <= maxrem) {
#pragma unroll
                    for (int i = 0; i < NV; i++) {
                        const double t = __shfl_down_sync(FULL, X[i], d);
                        if (d <= rem) X[i] += t;
                    }
                }
            }
            const int slot = (int)((task >> 47) & 7), pos = (int)((task >> 39) & 255);
            const int rowoff = __shfl_sync(FULL, my_rowoff, slot), ns = __shfl_sync(FULL, n, slot);
            if ((task >> 53) & 1) {
                double *p = rows + rowoff + NB * pos;
                if constexpr (OPG == 1) {
                    // K^{ab} = mu (delta_ab tr X + X[b][a]) + lambda X[a][b]
                    const double mtr = mu * (X[0] + X[4] + X[8]);
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int b = 0; b < 3; b++) p[a * ns + b] = fma(mu, X[3 * b + a], fma(lam, X[3 * a + b], a == b ? mtr : 0.0));
                } else p[0] = X[0];
            }
        }
        __syncwarp();

        // write-out: ONE TMA bulk store per node (its TPR dof rows are one contiguous run); odd head / tail doubles and
        // replicated rows of the other phase by plain stores
        if (lane < n_nodes && n > 0 && A.G.frag != nullptr && A.G.nseg == 0) {
            // fragment protocol (kernels.cuh): whole sectors by one bulk store, head / tail doubles into the row's slots
            bulk_fence();
            const int total = TPR * n;
            const double *nodep = rows + my_rowoff;
            const RunSplit sp = run_split(outp, total);
            const int body_n = total - sp.hc - sp.tc;
            if (body_n > 0) bulk_store(outp + sp.hc, nodep + sp.hc, body_n * 8);
            if (sp.hc) frag_head(A.G.frag, nflags >> 2, nodep, sp.hc);
            if (sp.tc) {
                if (nflags & 2u) { for (int x = total - sp.tc; x < total; x++) outp[x] = nodep[x]; }
                else frag_tail(A.G.frag, nflags >> 2, nodep + total - sp.tc, sp.tc);
            }
        } else
        if (lane < n_nodes && n > 0) {
            bulk_fence();
            const int total = TPR * n;
            const double *nodep = rows + my_rowoff;
#pragma unroll 1
            for (int d = 0; d < nrep; d++) {
                double *out = outp + (int64_t)d * total;
                const int h = (int)((reinterpret_cast<uintptr_t>(out) >> 3) & 1);
                if (h == head) {
                    const int body_n = (total - h) & ~1;
                    if (h) out[0] = nodep[0];
                    if (body_n > 0) bulk_store(out + h, nodep + h, body_n * 8);
                    if (h + body_n < total) out[total - 1] = nodep[total - 1];
                } else {
                    for (int x = 0; x < total; x++) out[x] = nodep[x];
                }
            }
        }
        bulk_commit_wait_read();   // the bulk stores read this warp's shared memory: wait before it is reused
        cp_async_wait<0>();
        __syncwarp();
        if (!more) break;
        tile = tile_n;
        buf ^= 1;
    }
}

} // namespace fb

namespace fb {

// =========================================================================================
//  * k_fan -- ring-ordered edge-node rows.  One LANE per incident element (W lanes per row node, 32 / W nodes per warp):
//    with canonical vertices (v0, v1, w_in, w_out) a lane evaluates the ten 3x3 column blocks of its element in
//    registers, column by column:  X = (|det| G_0) (x) q_0 + (|det| G_1) (x) q_1,  q_s = sum_t R_jc^{s t} G_sv(jc,t),
//    block = mu (tr X I + X^T) + lambda X.  The blocks of the out-face (v0, v1, w_out) go to the lane of the next
//    element of the ring by shuffles and are added to its in-face blocks, which are then final and stored ONCE into the
//    node's shared-memory row (closed rings: the last lane feeds the first); the three columns every element contributes
//    to (v0, v1, the row node itself) are summed through a shared-memory scratch (aliased with the rows, fixed order).
//    No read-modify-write, no atomics; one TMA bulk store per node.
//
//    Inputs are STREAMED: a lane's 32-byte fan record (positions, permutation, wiring, element, row base and length) and
//    its element's geometry line enter shared memory by cp.async (LDGSTS) -- the record two tiles ahead, the geometry
//    one tile ahead, in canonical vertex order -- so a warp never waits for global memory after its first tile and no
//    register is held across the latency.
// =========================================================================================
struct FanArgs {
    GatherArgs G;            // geom, values, c0/c1, R, pitch, vec_dim, ghost segments
    const uint4 *fanrec;     // [tile][32 lanes][2 x 16 bytes]: see k_fan_records
    int64_t ntiles;
    int W, npw;              // lanes per row node, row nodes per warp
};

// fan record of a lane (32 bytes):
//   bytes 0-9 positions of the ten canonical column nodes | 10 canonical permutation | 11 wiring: bits 0-4 lane whose
//   out-face blocks are this lane's carry, bit 5 carry present, bit 6 the own out-face blocks are final (the chain ends on
//   a boundary face), bit 7 the lane has an element | 12-15 element | 16-23 rowptr of the row node | 24-25 row length |
//   26 elements of the node | 27 flags: bit 0 the row has positions without a local contribution, bit 1 the node exists
__global__ void k_fan_records(const RowInfo *__restrict__ info, int64_t start, int64_t count, const uint32_t *__restrict__ rec,
                              int W, int npw, int64_t ntiles, uint32_t *__restrict__ fanrec)
{
    const int64_t nslots = ntiles * npw;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nslots; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = r / npw;
        const int slot = (int)(r - tile * npw), lane0 = slot * W;
        uint32_t *out = fanrec + (tile * 32 + lane0) * 8;
        if (r >= count) {
            for (int j = 0; j < W; j++)
                for (int x = 0; x < 8; x++) out[j * 8 + x] = x == 3 ? 0xffffffffu : 0u;
            continue;
        }
        const RowInfo &R = info[start + r];
        uint32_t wire[32];
        for (int j = 0; j < W; j++) wire[j] = 0;
        int chain_start = 0;
        for (int j = 0; j < R.ninc && j < W; j++) {
            const uint32_t mode = (rec[(R.k0 + j) * 8 + 6] >> 8) & 3;
            wire[j] |= 128u;
            if (mode == 0) { if (j + 1 < W) wire[j + 1] |= 32u | (uint32_t)(lane0 + j); }
            else if (mode == 1) { wire[j] |= 64u; chain_start = j + 1; }
            else { wire[chain_start] |= 32u | (uint32_t)(lane0 + j); chain_start = j + 1; }
        }
        const uint32_t flags = ((R.pad & 16) ? 1u : 0u) | 2u;
        for (int j = 0; j < W; j++) {
            uint32_t *o = out + j * 8;
            uint32_t w0 = 0, w1 = 0, w2 = 0, e = 0xffffffffu;
            if (j < R.ninc) {
                const uint32_t *w = rec + (R.k0 + j) * 8;
                uint32_t pos[10];
                for (int jc = 0; jc < 10; jc++) pos[jc] = (w[jc >> 1] >> (16 * (jc & 1))) & 0xffu;
                w0 = pos[0] | pos[1] << 8 | pos[2] << 16 | pos[3] << 24;
                w1 = pos[4] | pos[5] << 8 | pos[6] << 16 | pos[7] << 24;
                w2 = pos[8] | pos[9] << 8 | (w[6] & 0xffu) << 16;
                e = w[5];
            }
            o[0] = w0; o[1] = w1; o[2] = w2 | wire[j] << 24; o[3] = e;
            o[4] = (uint32_t)((uint64_t)R.base & 0xffffffffu); o[5] = (uint32_t)((uint64_t)R.base >> 32);
            o[6] = (uint32_t)R.len | (uint32_t)R.ninc << 16 | flags << 24; o[7] = 0;
        }
    }
}

// block of canonical column JC seen from an edge-node row (row support: canonical vertices 0, 1)
template <int OPG, int JC>
__device__ __forceinline__ void fan_block(const CanonR &R, const double (&G)[4][3], const double (&gs)[2][3], double mu, double lam,
                                          double (&out)[OPG == 1 ? 9 : 1])
{
    constexpr int DIM = 3, NVTX = 4;
    double q[2][3];
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int d = 0; d < 3; d++) q[s][d] = R.r[1][JC][s][0] * G[canon_sv<DIM>(JC, 0)][d];
        if (JC >= NVTX) {
#pragma unroll
            for (int d = 0; d < 3; d++) q[s][d] = fma(R.r[1][JC][s][1], G[canon_sv<DIM>(JC, 1)][d], q[s][d]);
        }
    }
    if constexpr (OPG == 1) {
        double X[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) X[a][b] = fma(gs[1][a], q[1][b], gs[0][a] * q[0][b]);
        const double mtr = mu * (X[0][0] + X[1][1] + X[2][2]);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) out[3 * a + b] = fma(mu, X[b][a], fma(lam, X[a][b], a == b ? mtr : 0.0));
    } else {
        double x = 0.0;
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int d = 0; d < 3; d++) x = fma(gs[s][d], q[s][d], x);
        out[0] = x;
    }
}

constexpr int kFanStageB = 2 * 32 * 32 + 32 * 128;   // bytes per warp: two record buffers + one geometry buffer

// WT > 0: lanes per node known at compile time (WT == A.W); WT == 0: run-time W
template <int OPG, int WT>
__global__ void __launch_bounds__(64, 8) k_fan(const FanArgs A)
{
    constexpr int DIM = 3, NVTX = 4;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr int SP = 33;                                  // scratch pitch (doubles per value row)
    constexpr int MAXR = OPG == 1 ? (WT > 0 ? (27 + WT - 1) / WT : 7) : 1;   // heavy values per lane (W >= 4)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int W = WT > 0 ? WT : A.W, npw = WT > 0 ? 32 / WT : A.npw, pitch = A.G.pitch;
    size_t warp_doubles = (size_t)npw * pitch;
    if (warp_doubles < (size_t)3 * NV * SP + 1) warp_doubles = (size_t)3 * NV * SP + 1;
    warp_doubles = (warp_doubles + 1) & ~(size_t)1;
    double *const rows = smem + (size_t)wib * (warp_doubles + kFanStageB / 8);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(rows + warp_doubles);   // 16-byte aligned
    // staging layout (16-byte chunks, lane-minor: conflict-free): record buffer b chunk c at ((2 b + c) 32 + lane),
    // geometry chunk c (canonical vertex c / 2, half c % 2) at (4 + c) 32 + lane
    const uint4 *const stage = reinterpret_cast<const uint4 *>(rows + warp_doubles);
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;
    const int slot = lane / W;
    const bool lane_used = slot < npw;
    const int idx = lane - slot * W;
    const int64_t step = (int64_t)gridDim.x * wpb;
    int64_t tile = (int64_t)blockIdx.x * wpb + wib;
    if (tile >= A.ntiles) return;

    auto request_rec = [&](int64_t t, int b) {
        const uint4 *src = A.fanrec + (t * 32 + lane) * 2;
        cp_async16_s(stage_s + ((2 * b + 0) * 32 + lane) * 16, src);
        cp_async16_s(stage_s + ((2 * b + 1) * 32 + lane) * 16, src + 1);
    };
    auto request_geo = [&](int b) {   // geometry line of the element in record buffer b, canonical vertex order
        const uint4 r0 = stage[(2 * b + 0) * 32 + lane];
        if (r0.w != 0xffffffffu) {
            const char *g = reinterpret_cast<const char *>(A.G.geom) + (int64_t)r0.w * 128;
            const uint32_t perm = (r0.z >> 16) & 0xffu;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const char *gv = g + 32 * ((perm >> (2 * v)) & 3);
                cp_async16_s(stage_s + ((4 + 2 * v) * 32 + lane) * 16, gv);
                cp_async16_s(stage_s + ((5 + 2 * v) * 32 + lane) * 16, gv + 16);
            }
        }
    };
    // prologue: record of the first tile, then its geometry and the record of the second tile
    request_rec(tile, 0);
    cp_async_commit();
    cp_async_wait<0>();
    request_geo(0);
    if (tile + step < A.ntiles) request_rec(tile + step, 1);
    cp_async_commit();
    cp_async_wait<0>();
    int buf = 0;

    for (;;) {
        const int64_t tile_n = tile + step;
        const bool more = tile_n < A.ntiles;
        // this tile's inputs: shared memory -> registers
        const uint4 r0 = stage[(2 * buf + 0) * 32 + lane], r1 = stage[(2 * buf + 1) * 32 + lane];
        const bool live = lane_used && r0.w != 0xffffffffu;
        double G[4][3], adet = 0.0;
        {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const uint4 a = stage[(4 + 2 * v) * 32 + lane], b = stage[(5 + 2 * v) * 32 + lane];
                G[v][0] = live ? __hiloint2double((int)a.y, (int)a.x) : 0.0;
                G[v][1] = live ? __hiloint2double((int)a.w, (int)a.z) : 0.0;
                G[v][2] = live ? __hiloint2double((int)b.y, (int)b.x) : 0.0;
                if (v == 0) adet = live ? __hiloint2double((int)b.w, (int)b.z) : 0.0;
            }
        }
        // requests: the record two tiles ahead into the buffer just read, the geometry of the next tile
        if (tile_n + step < A.ntiles) request_rec(tile_n + step, buf);
        if (more) request_geo(buf ^ 1);
        cp_async_commit();

        const bool node_ok = lane_used && (r1.z >> 24 & 2u) != 0;
        const int64_t base = (int64_t)((uint64_t)r1.x | (uint64_t)r1.y << 32);
        const int L = node_ok ? (int)(r1.z & 0xffffu) : 0;
        const int ninc = node_ok ? (int)((r1.z >> 16) & 0xffu) : 0;
        const bool holes = node_ok && (r1.z >> 24 & 1u) != 0;
        const int n = NB * L;
        const int64_t off_node = (int64_t)TPR * NB * nrep * base;
        double *const outp = node_ok ? out_ptr(A.G, off_node) : nullptr;
        const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
        double *const nodep = rows + (size_t)(lane_used ? slot : 0) * pitch + head;   // pitch is even
        const uint32_t pw[3] = {r0.x, r0.y, r0.z};
        auto posof = [&](int jc) { return (int)((pw[jc >> 2] >> (8 * (jc & 3))) & 0xffu) * NB; };

        // row-side vectors |det| G_0, |det| G_1
        double gs[2][3];
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int d = 0; d < 3; d++) gs[s][d] = adet * G[s][d];
        auto store_block = [&](int p, const double (&v)[NV]) {   // p = NB * position
            if constexpr (OPG == 1) {
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++) nodep[a * n + p + b] = v[3 * a + b];
            } else nodep[p] = v[0];
        };

        // ---- columns v0, v1 and the row node itself (canonical 0, 1, 4): every element of the star contributes ----
        {
            double hb[NV];
            fan_block<OPG, 0>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
            for (int v = 0; v < NV; v++) rows[(0 * NV + v) * SP + lane] = hb[v];
            fan_block<OPG, 1>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
            for (int v = 0; v < NV; v++) rows[(1 * NV + v) * SP + lane] = hb[v];
            fan_block<OPG, 4>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
            for (int v = 0; v < NV; v++) rows[(2 * NV + v) * SP + lane] = hb[v];
        }
        __syncwarp();
        double hv[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; r++) {
            const int v = idx + r * W;
            double s = 0.0;
            if (node_ok && v < 3 * NV) {
                const double *src = rows + v * SP + slot * W;
                if constexpr (WT > 0) {   // dead lanes wrote zeros
#pragma unroll
                    for (int m = 0; m < WT; m++) s += src[m];
                } else {
                    for (int m = 0; m < ninc; m++) s += src[m];
                }
            }
            hv[r] = s;
        }
        // positions of the three heavy columns: the node's first lane has a record whenever the node has elements
        const int lane0 = (lane_used ? slot : 0) * W;
        const uint32_t pw0 = __shfl_sync(FULL, pw[0], lane0), pw1 = __shfl_sync(FULL, pw[1], lane0);
        const int ph0 = (int)(pw0 & 0xffu) * NB, ph1 = (int)((pw0 >> 8) & 0xffu) * NB, ph4 = (int)(pw1 & 0xffu) * NB;
        __syncwarp();   // the scratch has been read: the rows may be written now
        if (__any_sync(FULL, holes)) { // rare: rows with positions no local element contributes to
            for (int x = lane; x < npw * pitch; x += 32) rows[x] = 0.0;
            __syncwarp();
        }
        if (node_ok && ninc > 0) {
#pragma unroll
            for (int r = 0; r < MAXR; r++) {
                const int v = idx + r * W;
                if (v < 3 * NV) {
                    const int ci = v / NV, ab = v - ci * NV;
                    const int p = ci == 0 ? ph0 : (ci == 1 ? ph1 : ph4);
                    if constexpr (OPG == 1) nodep[(ab / 3) * n + p + (ab % 3)] = hv[r];
                    else nodep[p] = hv[r];
                }
            }
        }

        // ---- out-face blocks (3, 7, 8) travel to the next element of the ring, whose in-face blocks (2, 6, 5) become final ----
        const uint32_t wire = r0.z >> 24;
        const int src = (int)(wire & 31u);
        const bool has_src = live && (wire & 32u) != 0, store_out = live && (wire & 64u) != 0;
#define FB_FAN_FACE(JO, JI)                                                                         \
        {                                                                                           \
            double bo[NV], bi[NV], cy[NV];                                                          \
            fan_block<OPG, JO>(A.G.R, G, gs, mu, lam, bo);                                           \
            _Pragma("unroll") for (int v = 0; v < NV; v++) cy[v] = __shfl_sync(FULL, bo[v], src);   \
            fan_block<OPG, JI>(A.G.R, G, gs, mu, lam, bi);                                           \
            if (live) {                                                                             \
                _Pragma("unroll") for (int v = 0; v < NV; v++) bi[v] += has_src ? cy[v] : 0.0;      \
                store_block(posof(JI), bi);                                                         \
                if (store_out) store_block(posof(JO), bo);                                          \
            }                                                                                       \
        }
        FB_FAN_FACE(3, 2)
        FB_FAN_FACE(7, 6)
        FB_FAN_FACE(8, 5)
#undef FB_FAN_FACE
        {
            double b9[NV];
            fan_block<OPG, 9>(A.G.R, G, gs, mu, lam, b9);
            if (live) store_block(posof(9), b9);
        }

        // ---- write-out: ONE TMA bulk store per node ----
        __syncwarp();
        if (node_ok && idx == 0 && n > 0) {
            bulk_fence();
            const int total = TPR * n;
#pragma unroll 1
            for (int d = 0; d < nrep; d++) {
                double *out = outp + (int64_t)d * total;
                const int h = (int)((reinterpret_cast<uintptr_t>(out) >> 3) & 1);
                if (h == head) {
                    const int body_n = (total - h) & ~1;
                    if (h) out[0] = nodep[0];
                    if (body_n > 0) bulk_store(out + h, nodep + h, body_n * 8);
                    if (h + body_n < total) out[total - 1] = nodep[total - 1];
                } else {
                    for (int x = 0; x < total; x++) out[x] = nodep[x];
                }
            }
        }
        bulk_commit_wait_read();
        cp_async_wait<0>();     // the next tile's geometry and the record after it have landed (requested a tile ago)
        __syncwarp();
        if (!more) break;
        tile = tile_n;
        buf ^= 1;
    }
}

} // namespace fb

namespace fb {

// =========================================================================================
//  * k_star -- ONE persistent kernel for all star tiles of an assembly (fan tiles of the edge-node rows, task tiles of the
//    vertex-node rows), walked in the ADDRESS ORDER of their CSR rows.  The type-bucketed launches write ~2 KB row runs
//    scattered over the whole values array (row types interleave node by node in the CSR): measured on B200
//    (tools/microbench_write.cu) 2 KB runs reach 3.2-3.6 TB/s in scattered order against 4.1-5.6 TB/s in address order
//    and 6.9 TB/s for a flat fill -- the row kernels were sitting on the scattered-write ceiling.  Here the warps of the
//    grid work at any moment on one contiguous window of the matrix (all row types), so DRAM pages fill while they are
//    open and the geometry lines of the window's elements are reused from L2 by every row type.
// =========================================================================================
struct StarArgs {
    GatherArgs G;            // geom, values, c0/c1, R, vec_dim, ghost segments
    const uint4 *tiles;      // [n_tiles] x = kind | W << 8 | npt << 16, y = index of the tile in its array, z = row capacity (nodes) of its bucket
    int64_t n_tiles;
    const uint4 *fanrec;     // fan tiles: [index][32 lanes][2 x 16 bytes]   (k_fan_records)
    const uint4 *tileblk;    // task tiles: [index][25 x 16 bytes]           (k_task_build)
    const uint64_t *tasks;
    int warp_bytes;          // shared memory per warp
};

// the warp copies the rows of its tile's nodes to their place in the values array.  Lane s < n_nodes holds the node's
// destination, the offset of its row in the warp's shared memory and its length (all dof rows, contiguous); 16-byte
// stores where source and destination have the same 16-byte phase (arranged by the callers), 8-byte stores otherwise
__device__ __forceinline__ void star_write_out(const double *rows, int n_nodes, double *outp, int rowoff, int total, int nrep, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll 1
    for (int s = 0; s < n_nodes; s++) {
        double *const op = reinterpret_cast<double *>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(outp), s));
        const int ro = __shfl_sync(FULL, rowoff, s), tot = __shfl_sync(FULL, total, s);
        if (tot == 0) continue;
        const double *src = rows + ro;
        const int hs = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
#pragma unroll 1
        for (int d = 0; d < nrep; d++) {
            double *out = op + (int64_t)d * tot;
            const int h = (int)((reinterpret_cast<uintptr_t>(out) >> 3) & 1);
            if (h == hs) {
                const int body = (tot - h) & ~1;
                if (lane == 0 && h) out[0] = src[0];
                if (lane == 1 && h + body < tot) out[tot - 1] = src[tot - 1];
                const double2 *s2 = reinterpret_cast<const double2 *>(src + h);
                double2 *o2 = reinterpret_cast<double2 *>(out + h);
                for (int x = lane; x < (body >> 1); x += 32) o2[x] = s2[x];
            } else {
                for (int x = lane; x < tot; x += 32) out[x] = src[x];
            }
        }
    }
}

// one fan tile: 32 / W ring-ordered edge-node rows, one lane per incident element (see k_fan)
template <int OPG, int WT>
__device__ __forceinline__ void star_fan_tile(const StarArgs &A, const uint4 *__restrict__ fr, double *rows, int pitch, int W_rt)
{
    constexpr int DIM = 3, NVTX = 4;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr int SP = 33;
    constexpr int MAXR = OPG == 1 ? (WT > 0 ? (27 + WT - 1) / WT : 7) : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int W = WT > 0 ? WT : W_rt, npw = 32 / W;
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;
    const int slot = lane / W;
    const bool lane_used = slot < npw;
    const int idx = lane - slot * W;

    const uint4 r0 = __ldg(fr + 2 * lane), r1 = __ldg(fr + 2 * lane + 1);
    const bool live = lane_used && r0.w != 0xffffffffu;
    double G[4][3], adet = 0.0;
#pragma unroll
    for (int v = 0; v < 4; v++) G[v][0] = G[v][1] = G[v][2] = 0.0;
    if (live) {
        const double *g = A.G.geom + (int64_t)r0.w * 16;
        const uint32_t perm = (r0.z >> 16) & 0xffu;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            double t[4];
            ld_v4g(g + 4 * ((perm >> (2 * v)) & 3), t);
            G[v][0] = t[0]; G[v][1] = t[1]; G[v][2] = t[2];
            if (v == 0) adet = t[3];
        }
    }
    const bool node_ok = lane_used && (r1.z >> 24 & 2u) != 0;
    const int64_t base = (int64_t)((uint64_t)r1.x | (uint64_t)r1.y << 32);
    const int L = node_ok ? (int)(r1.z & 0xffffu) : 0;
    const int ninc = node_ok ? (int)((r1.z >> 16) & 0xffu) : 0;
    const bool holes = node_ok && (r1.z >> 24 & 1u) != 0;
    const int n = NB * L;
    const int64_t off_node = (int64_t)TPR * NB * nrep * base;
    double *const outp = node_ok ? out_ptr(A.G, off_node) : nullptr;
    const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
    const int rowoff = (lane_used ? slot : 0) * pitch + head;          // pitch is even, rows is 16-byte aligned
    double *const nodep = rows + rowoff;
    const uint32_t pw[3] = {r0.x, r0.y, r0.z};
    auto posof = [&](int jc) { return (int)((pw[jc >> 2] >> (8 * (jc & 3))) & 0xffu) * NB; };

    double gs[2][3];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int d = 0; d < 3; d++) gs[s][d] = adet * G[s][d];
    auto store_block = [&](int p, const double (&v)[NV]) {
        if constexpr (OPG == 1) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) nodep[a * n + p + b] = v[3 * a + b];
        } else nodep[p] = v[0];
    };
    {
        double hb[NV];
        fan_block<OPG, 0>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
        for (int v = 0; v < NV; v++) rows[(0 * NV + v) * SP + lane] = hb[v];
        fan_block<OPG, 1>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
        for (int v = 0; v < NV; v++) rows[(1 * NV + v) * SP + lane] = hb[v];
        fan_block<OPG, 4>(A.G.R, G, gs, mu, lam, hb);
#pragma unroll
        for (int v = 0; v < NV; v++) rows[(2 * NV + v) * SP + lane] = hb[v];
    }
    __syncwarp();
    double hv[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; r++) {
        const int v = idx + r * W;
        double s = 0.0;
        if (node_ok && v < 3 * NV) {
            const double *src = rows + v * SP + slot * W;
            if constexpr (WT > 0) {
#pragma unroll
                for (int m = 0; m < WT; m++) s += src[m];
            } else {
                for (int m = 0; m < ninc; m++) s += src[m];
            }
        }
        hv[r] = s;
    }
    const int lane0 = (lane_used ? slot : 0) * W;
    const uint32_t pw0 = __shfl_sync(FULL, pw[0], lane0), pw1 = __shfl_sync(FULL, pw[1], lane0);
    const int ph0 = (int)(pw0 & 0xffu) * NB, ph1 = (int)((pw0 >> 8) & 0xffu) * NB, ph4 = (int)(pw1 & 0xffu) * NB;
    __syncwarp();
    if (__any_sync(FULL, holes)) {
        for (int x = lane; x < npw * pitch; x += 32) rows[x] = 0.0;
        __syncwarp();
    }
    if (node_ok && ninc > 0) {
#pragma unroll
        for (int r = 0; r < MAXR; r++) {
            const int v = idx + r * W;
            if (v < 3 * NV) {
                const int ci = v / NV, ab = v - ci * NV;
                const int p = ci == 0 ? ph0 : (ci == 1 ? ph1 : ph4);
                if constexpr (OPG == 1) nodep[(ab / 3) * n + p + (ab % 3)] = hv[r];
                else nodep[p] = hv[r];
            }
        }
    }
    const uint32_t wire = r0.z >> 24;
    const int src = (int)(wire & 31u);
    const bool has_src = live && (wire & 32u) != 0, store_out = live && (wire & 64u) != 0;
#define FB_FAN_FACE(JO, JI)                                                                         \
    {                                                                                               \
        double bo[NV], bi[NV], cy[NV];                                                              \
        fan_block<OPG, JO>(A.G.R, G, gs, mu, lam, bo);                                               \
        _Pragma("unroll") for (int v = 0; v < NV; v++) cy[v] = __shfl_sync(FULL, bo[v], src);       \
        fan_block<OPG, JI>(A.G.R, G, gs, mu, lam, bi);                                               \
        if (live) {                                                                                 \
            _Pragma("unroll") for (int v = 0; v < NV; v++) bi[v] += has_src ? cy[v] : 0.0;          \
            store_block(posof(JI), bi);                                                             \
            if (store_out) store_block(posof(JO), bo);                                              \
        }                                                                                           \
    }
    FB_FAN_FACE(3, 2)
    FB_FAN_FACE(7, 6)
    FB_FAN_FACE(8, 5)
#undef FB_FAN_FACE
    {
        double b9[NV];
        fan_block<OPG, 9>(A.G.R, G, gs, mu, lam, b9);
        if (live) store_block(posof(9), b9);
    }
    __syncwarp();
    // lane s (< npw) publishes node s: its values live in lane s * W
    {
        const int sl = lane < npw ? lane * W : 0;
        double *const op = reinterpret_cast<double *>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(outp), sl));
        const int ro = __shfl_sync(FULL, rowoff, sl), tot = __shfl_sync(FULL, TPR * n, sl);
        star_write_out(rows, npw, op, ro, lane < npw ? tot : 0, nrep, lane);
    }
    __syncwarp();
}

// one task tile: a few vertex-node rows, lanes = node blocks (see k_task)
template <int OPG>
__device__ __forceinline__ void star_task_tile(const StarArgs &A, const uint4 *__restrict__ blk, double *rows, int pitch, int npt)
{
    constexpr int DIM = 3, NVTX = 4, NL = 10;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double2 *const XY = reinterpret_cast<double2 *>(rows + (size_t)npt * pitch);
    double *const Z = reinterpret_cast<double *>(XY + kTaskMaxTets * kTaskVec);
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;

    const uint4 hdr = __ldg(blk);
    const int n_nodes = hdr.x & 0xff, n_tets = (hdr.x >> 8) & 0xff, n_passes = (hdr.x >> 16) & 0xff;
    const uint2 et = __ldg(reinterpret_cast<const uint2 *>(blk + 9) + lane);
    int64_t base = 0;
    int L = 0;
    if (lane < n_nodes) {
        const uint4 nd = __ldg(blk + 1 + lane);
        base = (int64_t)((uint64_t)nd.x | (uint64_t)nd.y << 32);
        L = (int)nd.z;
    }
    const uint64_t *tp = A.tasks + (int64_t)hdr.y * 32 + lane;
    uint64_t tw0 = 0, tw1 = 0, tw2 = 0;   // the first three passes' task words travel with the geometry
    if (n_passes > 0) tw0 = __ldg(reinterpret_cast<const unsigned long long *>(tp));
    if (n_passes > 1) tw1 = __ldg(reinterpret_cast<const unsigned long long *>(tp + 32));
    if (n_passes > 2) tw2 = __ldg(reinterpret_cast<const unsigned long long *>(tp + 64));
    if (lane < n_tets) {
        double G[4][4];
        const double *g = A.G.geom + (int64_t)et.x * 16;
#pragma unroll
        for (int v = 0; v < 4; v++) ld_v4g(g + 4 * ((et.y >> (2 * v)) & 3), G[v]);
        const double adet = G[0][3];
        XY[lane * kTaskVec] = make_double2(adet * G[0][0], adet * G[0][1]);
        Z[lane * kTaskVec] = adet * G[0][2];
#pragma unroll
        for (int jc = 0; jc < NL; jc++) {
            const double r0 = A.G.R.r[0][jc][0][0];
            double q[3];
#pragma unroll
            for (int d = 0; d < 3; d++) q[d] = r0 * G[canon_sv<DIM>(jc, 0)][d];
            if (jc >= NVTX) {
                const double r1 = A.G.R.r[0][jc][0][1];
#pragma unroll
                for (int d = 0; d < 3; d++) q[d] = fma(r1, G[canon_sv<DIM>(jc, 1)][d], q[d]);
            }
            XY[lane * kTaskVec + 1 + jc] = make_double2(q[0], q[1]);
            Z[lane * kTaskVec + 1 + jc] = q[2];
        }
    }
    __syncwarp();
    const int n = NB * L;
    const int64_t off_node = (int64_t)TPR * NB * nrep * base;
    double *const outp = lane < n_nodes ? out_ptr(A.G, off_node) : nullptr;
    const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
    const int my_rowoff = lane * pitch + head;

#pragma unroll 1
    for (int pass = 0; pass < n_passes; pass++) {
        const uint64_t task = pass == 0 ? tw0 : (pass == 1 ? tw1 : (pass == 2 ? tw2 : __ldg(reinterpret_cast<const unsigned long long *>(tp + 32 * pass))));
        const int np = (int)((task >> 36) & 7);
        const int nsteps = __reduce_max_sync(FULL, np);
        double X[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) X[i] = 0.0;
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            if (s < np) {
                const uint32_t pr = (uint32_t)(task >> (9 * s)) & 0x1ffu;
                const int m = pr & 31, jc = pr >> 5;
                const double2 gxy = XY[m * kTaskVec], qxy = XY[m * kTaskVec + 1 + jc];
                const double gz = Z[m * kTaskVec], qz = Z[m * kTaskVec + 1 + jc];
                if constexpr (OPG == 1) {
                    X[0] = fma(gxy.x, qxy.x, X[0]); X[1] = fma(gxy.x, qxy.y, X[1]); X[2] = fma(gxy.x, qz, X[2]);
                    X[3] = fma(gxy.y, qxy.x, X[3]); X[4] = fma(gxy.y, qxy.y, X[4]); X[5] = fma(gxy.y, qz, X[5]);
                    X[6] = fma(gz, qxy.x, X[6]);    X[7] = fma(gz, qxy.y, X[7]);    X[8] = fma(gz, qz, X[8]);
                } else {
                    X[0] = fma(gxy.x, qxy.x, X[0]); X[0] = fma(gxy.y, qxy.y, X[0]); X[0] = fma(gz, qz, X[0]);
                }
            }
        }
        const int rem = (int)((task >> 50) & 7);
        if (__any_sync(FULL, rem != 0)) {
#pragma unroll
            for (int d = 1; d <= 4; d <<= 1) {
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    const double t = __shfl_down_sync(FULL, X[i], d);
                    if (d <= rem) X[i] += t;
                }
            }
        }
        const int slot = (int)((task >> 47) & 7), pos = (int)((task >> 39) & 255);
        const int rowoff = __shfl_sync(FULL, my_rowoff, slot), ns = __shfl_sync(FULL, n, slot);
        if ((task >> 53) & 1) {
            double *p = rows + rowoff + NB * pos;
            if constexpr (OPG == 1) {
                const double mtr = mu * (X[0] + X[4] + X[8]);
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++) p[a * ns + b] = fma(mu, X[3 * b + a], fma(lam, X[3 * a + b], a == b ? mtr : 0.0));
            } else p[0] = X[0];
        }
    }
    __syncwarp();
    star_write_out(rows, n_nodes, outp, my_rowoff, lane < n_nodes ? TPR * n : 0, nrep, lane);
    __syncwarp();
}

template <int OPG>
__global__ void __launch_bounds__(64, 8) k_star(const StarArgs A)
{
    extern __shared__ double smem[];
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *const rows = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + (size_t)wib * A.warp_bytes);
    const int64_t step = (int64_t)gridDim.x * wpb;
    int64_t i = (int64_t)blockIdx.x * wpb + wib;
    if (i >= A.n_tiles) return;
    uint4 e = __ldg(A.tiles + i);
    for (;;) {
        const int64_t i_n = i + step;
        uint4 e_n = make_uint4(0u, 0u, 0u, 0u);
        if (i_n < A.n_tiles) e_n = __ldg(A.tiles + i_n);
        constexpr int BS = OPG == 1 ? 9 : 1;                       // values per node block
        const int kind = e.x & 0xff, w = (e.x >> 8) & 0xff, npt = (e.x >> 16) & 0xff;
        const int pitch = (BS * (int)e.z + 2) & ~1;              // e.z: row capacity of the tile's bucket (nodes); even pitch
        if (kind == 0) star_task_tile<OPG>(A, A.tileblk + (int64_t)e.y * kTileBlkChunks, rows, pitch, npt);
        else if (w == 6) star_fan_tile<OPG, 6>(A, A.fanrec + (int64_t)e.y * 64, rows, pitch, w);
        else if (w == 4) star_fan_tile<OPG, 4>(A, A.fanrec + (int64_t)e.y * 64, rows, pitch, w);
        else star_fan_tile<OPG, 0>(A, A.fanrec + (int64_t)e.y * 64, rows, pitch, w);
        if (i_n >= A.n_tiles) break;
        i = i_n;
        e = e_n;
    }
}

} // namespace fb
