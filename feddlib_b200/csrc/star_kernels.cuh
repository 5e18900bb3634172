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

// one-time: task programs.  pass 1 (count): passes per tile; pass 2 (emit): task words + the tile's element table
struct TaskBuildArgs {
    const RowInfo *rowinfo;      // bucket-ordered row records
    const uint32_t *rec;         // incidence records (8 words)
    TaskTile *tiles;             // q0, n_nodes filled by the host; n_tets, n_passes by pass 1; task_off by the host before pass 2
    int64_t n_tiles;
    uint64_t *tasks;             // pass 2
    uint2 *tiletet;              // pass 2: [tile][32] (element, canonical permutation)
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
        int m = 0;
        for (int i = 0; i < T.n_nodes; i++)
            for (int j = 0; j < ninc[i]; j++, m++) {
                const uint32_t *w = A.rec + (k0[i] + j) * 8;
                A.tiletet[t * kTaskMaxTets + m] = make_uint2(w[5], w[6]);
            }
        for (; m < kTaskMaxTets; m++) A.tiletet[t * kTaskMaxTets + m] = make_uint2(0u, 0u);
    }
}

struct TaskArgs {
    GatherArgs G;                // rowinfo, geom, values, c0/c1 (lambda, mu), R, pitch, vec_dim, ghost segments
    const TaskTile *tiles;       // tiles of this launch
    int64_t n_tiles;
    const uint64_t *tasks;
    const uint2 *tiletet;
    int npt;                     // row nodes per tile at most (shared-memory rows per warp)
};

constexpr int kTaskVec = 11;                                   // staged vectors per element: |det| G_0, q_0 .. q_9
constexpr int kTaskStageB = kTaskMaxTets * kTaskVec * 24;      // bytes per warp: XY (double2) + Z (double) planes

// OPG 0: Laplace (one value per node block; vec_dim: replicated on the block diagonal), OPG 1: elasticity (3x3 blocks)
template <int OPG>
__global__ void __launch_bounds__(64, 8) k_task(const TaskArgs A)
{
    constexpr int DIM = 3, NVTX = 4, NL = 10;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int pitch = A.G.pitch;                                   // doubles per node row area (even)
    const size_t warp_doubles = (size_t)A.npt * pitch + kTaskStageB / 8;
    double *const rows = smem + (size_t)wib * warp_doubles;        // 16-byte aligned: pitch and the stage size are even
    double2 *const XY = reinterpret_cast<double2 *>(rows + (size_t)A.npt * pitch);
    double *const Z = reinterpret_cast<double *>(XY + kTaskMaxTets * kTaskVec);
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;
    const int64_t step = (int64_t)gridDim.x * wpb;
    int64_t tile = (int64_t)blockIdx.x * wpb + wib;
    if (tile >= A.n_tiles) return;

    auto load_geo_lane = [&](uint2 et, double (&G)[4][4]) {
        const double *g = A.G.geom + (int64_t)et.x * 16;
#pragma unroll
        for (int v = 0; v < 4; v++) ld_v4g(g + 4 * ((et.y >> (2 * v)) & 3), G[v]);
    };
    // prologue: header, element table and geometry of the first tile
    uint4 hdr = __ldg(reinterpret_cast<const uint4 *>(A.tiles + tile));
    uint2 et = __ldg(A.tiletet + tile * kTaskMaxTets + lane);
    double G[4][4];
    if (lane < (int)((hdr.y >> 8) & 0xff)) load_geo_lane(et, G);

    for (;;) {
        const int n_nodes = hdr.y & 0xff, n_tets = (hdr.y >> 8) & 0xff, n_passes = (hdr.y >> 16) & 0xff;
        const int64_t tile_n = tile + step;
        const bool more = tile_n < A.n_tiles;
        uint4 hdr_n = make_uint4(0u, 0u, 0u, 0u);
        uint2 et_n = make_uint2(0u, 0u);
        if (more) {
            hdr_n = __ldg(reinterpret_cast<const uint4 *>(A.tiles + tile_n));
            et_n = __ldg(A.tiletet + tile_n * kTaskMaxTets + lane);
        }
        // row records of the tile's nodes (lane = node slot)
        int64_t base = 0;
        int L = 0;
        if (lane < n_nodes) {
            double raw[4];
            ld_v4(reinterpret_cast<const double *>(A.G.rowinfo + hdr.x + lane), raw);
            base = __double_as_longlong(raw[0]);
            L = (int)(__double_as_longlong(raw[2]) & 0xffffffff);
        }
        // stage: lane m = element slot m.  vector 0: |det| G_0 (row function: canonical vertex 0), vector 1 + jc:
        // q_jc = R_jc^0 G_sv(jc,0) + R_jc^1 G_sv(jc,1)
        if (lane < n_tets) {
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
                    for (int d = 0; d < 3; d++) q[d] += r1 * G[canon_sv<DIM>(jc, 1)][d];
                }
                XY[lane * kTaskVec + 1 + jc] = make_double2(q[0], q[1]);
                Z[lane * kTaskVec + 1 + jc] = q[2];
            }
        }
        // the geometry registers are free: request the next tile's lines now, they land while this tile is processed
        if (more && lane < (int)((hdr_n.y >> 8) & 0xff)) load_geo_lane(et_n, G);
        __syncwarp();

        // node row of slot s: `pitch` doubles apart, shifted by one double where that gives the row the 16-byte phase of
        // its destination (the TMA bulk store needs both sides 16-byte aligned)
        const int n = NB * L;                                         // values per dof row (lane = node slot)
        const int64_t off_node = (int64_t)TPR * NB * nrep * base;
        double *const outp = lane < n_nodes ? out_ptr(A.G, off_node) : nullptr;
        const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
        const int my_rowoff = lane * pitch + ((lane * pitch + head) & 1);   // lane = node slot

        const uint64_t *tp = A.tasks + (int64_t)hdr.z * 32 + lane;
#pragma unroll 1
        for (int pass = 0; pass < n_passes; pass++, tp += 32) {
            const uint64_t task = __ldg(reinterpret_cast<const unsigned long long *>(tp));
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
                        X[0] += gxy.x * qxy.x; X[1] += gxy.x * qxy.y; X[2] += gxy.x * qz;
                        X[3] += gxy.y * qxy.x; X[4] += gxy.y * qxy.y; X[5] += gxy.y * qz;
                        X[6] += gz * qxy.x;    X[7] += gz * qxy.y;    X[8] += gz * qz;
                    } else {
                        X[0] += gxy.x * qxy.x; X[0] += gxy.y * qxy.y; X[0] += gz * qz;
                    }
                }
            }
            // positions with more than four pairs: the partial sums of the group's lanes, fixed order
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
                    // K^{ab} = mu (delta_ab tr X + X[b][a]) + lambda X[a][b]
                    const double mtr = mu * (X[0] + X[4] + X[8]);
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int b = 0; b < 3; b++) p[a * ns + b] = mu * X[3 * b + a] + lam * X[3 * a + b] + (a == b ? mtr : 0.0);
                } else p[0] = X[0];
            }
        }

        // write-out: ONE TMA bulk store per node (its TPR dof rows are one contiguous run); odd head / tail doubles and
        // replicated rows of the other phase by plain stores
        __syncwarp();
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
        __syncwarp();
        if (!more) break;
        tile = tile_n;
        hdr = hdr_n;
        et = et_n;
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
// =========================================================================================
struct FanArgs {
    GatherArgs G;            // rowinfo, start/count (bucket rows), geom, values, c0/c1, R, pitch, vec_dim, ghost segments
    const uint32_t *fanrec;  // [(row - start) * W + idx][8]: the bucket's records, W per row (w[5] = 0xffffffff: none)
    int W, npw;              // lanes per row node, row nodes per warp
};

// one-time: the padded per-bucket record array of k_fan.  Word 7 (the natural-index word, unused by these operators)
// carries the lane wiring: bits 0-4 lane whose out-face blocks are this lane's carry, bit 5 carry present, bit 6 the own
// out-face blocks are final (the chain ends on a boundary face).
__global__ void k_fan_records(const RowInfo *__restrict__ info, int64_t start, int64_t count, const uint32_t *__restrict__ rec,
                              int W, int npw, uint32_t *__restrict__ fanrec)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < count; r += (int64_t)gridDim.x * blockDim.x) {
        const RowInfo &R = info[start + r];
        const int lane0 = (int)(r % npw) * W;
        uint32_t wire[32];
        for (int j = 0; j < W && j < 32; j++) wire[j] = 0;
        int chain_start = 0;
        for (int j = 0; j < R.ninc && j < W; j++) {
            const uint32_t mode = (rec[(R.k0 + j) * 8 + 6] >> 8) & 3;
            if (mode == 0) { if (j + 1 < W) wire[j + 1] |= 32u | (uint32_t)(lane0 + j); }
            else if (mode == 1) { wire[j] |= 64u; chain_start = j + 1; }
            else { wire[chain_start] |= 32u | (uint32_t)(lane0 + j); chain_start = j + 1; }
        }
        for (int j = 0; j < W; j++) {
            uint32_t *o = fanrec + (r * W + j) * 8;
            if (j < R.ninc) {
                const uint32_t *w = rec + (R.k0 + j) * 8;
                for (int x = 0; x < 7; x++) o[x] = w[x];
                o[7] = wire[j];
            } else {
                for (int x = 0; x < 8; x++) o[x] = x == 5 ? 0xffffffffu : 0u;
            }
        }
    }
}

// block of canonical column JC seen from an edge-node row (row support: canonical vertices 0, 1)
template <int OPG, int JC>
__device__ __forceinline__ void fan_block(const CanonR &R, const double (&G)[4][4], const double (&gs)[2][3], double mu, double lam,
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
            for (int d = 0; d < 3; d++) q[s][d] += R.r[1][JC][s][1] * G[canon_sv<DIM>(JC, 1)][d];
        }
    }
    if constexpr (OPG == 1) {
        double X[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) X[a][b] = gs[0][a] * q[0][b] + gs[1][a] * q[1][b];
        const double mtr = mu * (X[0][0] + X[1][1] + X[2][2]);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) out[3 * a + b] = mu * X[b][a] + lam * X[a][b] + (a == b ? mtr : 0.0);
    } else {
        double x = 0.0;
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int d = 0; d < 3; d++) x += gs[s][d] * q[s][d];
        out[0] = x;
    }
}

template <int OPG>
__global__ void __launch_bounds__(64, 8) k_fan(const FanArgs A)
{
    constexpr int DIM = 3, NVTX = 4;
    constexpr int NB = OPG == 1 ? DIM : 1, TPR = OPG == 1 ? DIM : 1, NV = OPG == 1 ? 9 : 1;
    constexpr int SP = 33;                                  // scratch pitch (doubles per value row)
    constexpr int MAXR = OPG == 1 ? 7 : 1;                  // heavy values per lane: ceil(3 * NV / W), W >= 4
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int W = A.W, npw = A.npw, pitch = A.G.pitch;
    size_t warp_doubles = (size_t)npw * pitch;
    if (warp_doubles < (size_t)3 * NV * SP + 1) warp_doubles = (size_t)3 * NV * SP + 1;
    warp_doubles = (warp_doubles + 1) & ~(size_t)1;
    double *const rows = smem + (size_t)wib * warp_doubles;
    const double mu = A.G.c1, lam = A.G.c0;
    const int nrep = (OPG == 0 && A.G.vec_dim != 0) ? A.G.vec_dim : 1;
    const int slot = lane / W;
    const bool lane_used = slot < npw;
    const int idx = lane - slot * W;
    const int64_t ntiles = (A.G.count + npw - 1) / npw;
    const int64_t step = (int64_t)gridDim.x * wpb;
    int64_t tile = (int64_t)blockIdx.x * wpb + wib;
    if (tile >= ntiles) return;

    auto load_rec = [&](int64_t t, uint32_t (&w)[8], double (&raw)[4]) {
        const int64_t node = t * npw + slot;
#pragma unroll
        for (int x = 0; x < 8; x++) w[x] = 0u;
        w[5] = 0xffffffffu;
        raw[0] = raw[1] = raw[2] = raw[3] = 0.0;
        if (lane_used && node < A.G.count) {
            ld_v8u(A.fanrec + (node * W + idx) * 8, w);
            ld_v4(reinterpret_cast<const double *>(A.G.rowinfo + A.G.start + node), raw);
        }
    };
    auto load_geo_lane = [&](const uint32_t (&w)[8], double (&G)[4][4]) {
        if (w[5] != 0xffffffffu) {
            const double *g = A.G.geom + (int64_t)w[5] * 16;
#pragma unroll
            for (int v = 0; v < 4; v++) ld_v4g(g + 4 * ((w[6] >> (2 * v)) & 3), G[v]);
        } else {
#pragma unroll
            for (int v = 0; v < 4; v++)
#pragma unroll
                for (int d = 0; d < 4; d++) G[v][d] = 0.0;
        }
    };
    uint32_t w[8];
    double raw[4], G[4][4];
    load_rec(tile, w, raw);
    load_geo_lane(w, G);

    for (;;) {
        const int64_t tile_n = tile + step;
        const bool more = tile_n < ntiles;
        uint32_t wn[8];
        double rawn[4];
        if (more) load_rec(tile_n, wn, rawn);

        const bool live = w[5] != 0xffffffffu;
        const bool node_ok = lane_used && tile * npw + slot < A.G.count;
        const int64_t base = __double_as_longlong(raw[0]);
        const int L = node_ok ? (int)(__double_as_longlong(raw[2]) & 0xffffffff) : 0;
        const int ninc = node_ok ? (int)(__double_as_longlong(raw[2]) >> 32) : 0;
        const bool holes = node_ok && (__double_as_longlong(raw[3]) & 16) != 0;
        const int n = NB * L;
        const int64_t off_node = (int64_t)TPR * NB * nrep * base;
        double *const outp = node_ok ? out_ptr(A.G, off_node) : nullptr;
        const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
        double *const nodep = rows + (size_t)(lane_used ? slot : 0) * pitch + head;   // pitch is even
        auto posof = [&](int jc) { return (int)((w[jc >> 1] >> (16 * (jc & 1))) & 0xffffu) * NB; };

        // row-side vectors |det| G_0, |det| G_1
        const double adet = G[0][3];
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
                for (int m = 0; m < ninc; m++) s += src[m];
            }
            hv[r] = s;
        }
        // positions of the three heavy columns: the node's first lane has a record whenever the node has elements
        const int lane0 = (lane_used ? slot : 0) * W;
        const int ph0 = __shfl_sync(FULL, posof(0), lane0), ph1 = __shfl_sync(FULL, posof(1), lane0), ph4 = __shfl_sync(FULL, posof(4), lane0);
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
        const int src = (int)(w[7] & 31u);
        const bool has_src = live && (w[7] & 32u) != 0, store_out = live && (w[7] & 64u) != 0;
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
        // the geometry registers are free: the next tile's lines
        if (more) load_geo_lane(wn, G);

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
        __syncwarp();
        if (!more) break;
        tile = tile_n;
#pragma unroll
        for (int x = 0; x < 8; x++) w[x] = wn[x];
#pragma unroll
        for (int x = 0; x < 4; x++) raw[x] = rawn[x];
    }
}

} // namespace fb
