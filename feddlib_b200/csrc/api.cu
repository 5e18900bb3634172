// C-ABI assembly entry points of libfeddb200.so: operator dispatch, scatter-mode selection and
// kernel launches.  See include/feddb200.h for the reference interfaces each call replaces.
#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "kernels.cuh"
#include "star_kernels.cuh"

using namespace fb;

namespace {

int elem_index(int dim, int nloc)
{
    if (dim == 2) return nloc == 3 ? 0 : (nloc == 6 ? 1 : -1);
    if (dim == 3) return nloc == 4 ? 2 : (nloc == 10 ? 3 : -1);
    return -1;
}

// raises the dynamic shared-memory limit of a kernel and asks for its residency once per (kernel, block size, shared memory)
// and context; per_sm may be null (attribute only)
template <class K>
int kernel_cfg(feddb200_ctx *c, K kernel, int nt, size_t smem, size_t budget, int *per_sm)
{
    const void *f = reinterpret_cast<const void *>(kernel);
    const size_t key_smem = per_sm ? smem : (size_t)-1;
    for (const feddb200_ctx::OccEntry &e : c->occ_cache)
        if (e.f == f && e.nt == nt && e.smem == key_smem) { if (per_sm) *per_sm = e.per_sm; return FEDDB200_OK; }
    cudaFuncAttributes fa;
    FB_CUDA(cudaFuncGetAttributes(&fa, kernel));   // static shared memory counts against the opt-in limit
    const size_t dyn_max = std::min(budget, (size_t)c->smem_optin - std::min((size_t)c->smem_optin, fa.sharedSizeBytes));
    FB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max));
    int v = 1;
    if (per_sm) { FB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, nt, smem)); *per_sm = v; }
    c->occ_cache.push_back({f, nt, key_smem, v});
    return FEDDB200_OK;
}

// tuning knob from the environment, read once per call site
#define FB_ENV_INT(NAME, DEFAULT) ([] { static const int v = [] { const char *f = getenv(NAME); return f ? atoi(f) : (DEFAULT); }(); return v; }())

// device copy of the operator tables, cached per context slot
int get_tables(feddb200_ctx *c, int op, int dim, int nv, int np, const OpTables **out_d, OpTables *out_h = nullptr)
{
    OpTables t;
    if (build_tables(t, op, dim, nv, np) != 0) {
        set_error("no quadrature/basis tables for this combination of dimension and FE types");
        return FEDDB200_ELOGIC;
    }
    const int key = 1 + dim * 10000 + nv * 100 + np;
    OpTables *slot = c->tab_d + op;
    if (c->tab_key[op] != key) {
        FB_CUDA(cudaMemcpyAsync(slot, &t, sizeof(OpTables), cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(cudaStreamSynchronize(c->stream));
        c->tab_key[op] = key;
    }
    *out_d = slot;
    if (out_h) *out_h = t;
    return FEDDB200_OK;
}

template <int OP, int DIM, int NR, int NC>
int launch_elem_t(feddb200_ctx *c, ElemArgs &A, bool atomic)
{
    constexpr int RD = RowDofs<OP, DIM>::value;
    const int64_t nthreads = A.n_items * NR * RD;
    if (nthreads == 0) return FEDDB200_OK;
    const int64_t blocks = (nthreads + 127) / 128;
    FB_LOGIC(blocks >= (int64_t(1) << 31), "launch too large");
    if (atomic) k_elem<OP, DIM, NR, NC, true><<<(unsigned)blocks, 128, 0, c->stream>>>(A);
    else k_elem<OP, DIM, NR, NC, false><<<(unsigned)blocks, 128, 0, c->stream>>>(A);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

template <int OP>
int launch_elem(feddb200_ctx *c, int dim, int nr, int nc, ElemArgs &A, bool atomic)
{
    if constexpr (OP == OP_B) {
        if (dim == 2 && nr == 1 && nc == 3) return launch_elem_t<OP, 2, 1, 3>(c, A, atomic);   // P0 pressure: rows on the element map
        if (dim == 2 && nr == 1 && nc == 6) return launch_elem_t<OP, 2, 1, 6>(c, A, atomic);
        if (dim == 2 && nr == 3 && nc == 3) return launch_elem_t<OP, 2, 3, 3>(c, A, atomic);
        if (dim == 2 && nr == 3 && nc == 6) return launch_elem_t<OP, 2, 3, 6>(c, A, atomic);
        if (dim == 3 && nr == 4 && nc == 4) return launch_elem_t<OP, 3, 4, 4>(c, A, atomic);
        if (dim == 3 && nr == 4 && nc == 10) return launch_elem_t<OP, 3, 4, 10>(c, A, atomic);
    } else if constexpr (OP == OP_BT) {
        if (dim == 2 && nr == 3 && nc == 1) return launch_elem_t<OP, 2, 3, 1>(c, A, atomic);   // P0 pressure: columns on the element map
        if (dim == 2 && nr == 6 && nc == 1) return launch_elem_t<OP, 2, 6, 1>(c, A, atomic);
        if (dim == 2 && nr == 3 && nc == 3) return launch_elem_t<OP, 2, 3, 3>(c, A, atomic);
        if (dim == 2 && nr == 6 && nc == 3) return launch_elem_t<OP, 2, 6, 3>(c, A, atomic);
        if (dim == 3 && nr == 4 && nc == 4) return launch_elem_t<OP, 3, 4, 4>(c, A, atomic);
        if (dim == 3 && nr == 10 && nc == 4) return launch_elem_t<OP, 3, 10, 4>(c, A, atomic);
    } else {
        if (nr == nc) {
            if (dim == 2 && nr == 3) return launch_elem_t<OP, 2, 3, 3>(c, A, atomic);
            if (dim == 2 && nr == 6) return launch_elem_t<OP, 2, 6, 6>(c, A, atomic);
            if (dim == 3 && nr == 4) return launch_elem_t<OP, 3, 4, 4>(c, A, atomic);
            if (dim == 3 && nr == 10) return launch_elem_t<OP, 3, 10, 10>(c, A, atomic);
        }
    }
    set_error("this operator is not implemented for the given combination of FE types");
    return FEDDB200_ELOGIC;
}

int launch_elem_op(feddb200_ctx *c, int op, int dim, int nr, int nc, ElemArgs &A, bool atomic)
{
    switch (op) {
    case OP_LAP:  return launch_elem<OP_LAP>(c, dim, nr, nc, A, atomic);
    case OP_ELAS: return launch_elem<OP_ELAS>(c, dim, nr, nc, A, atomic);
    case OP_ADV:  return launch_elem<OP_ADV>(c, dim, nr, nc, A, atomic);
    case OP_MASS: return launch_elem<OP_MASS>(c, dim, nr, nc, A, atomic);
    case OP_ADVU: return launch_elem<OP_ADVU>(c, dim, nr, nc, A, atomic);
    case OP_NSJ:  return launch_elem<OP_NSJ>(c, dim, nr, nc, A, atomic);
    case OP_B:    return launch_elem<OP_B>(c, dim, nr, nc, A, atomic);
    case OP_BT:   return launch_elem<OP_BT>(c, dim, nr, nc, A, atomic);
    }
    set_error("unknown operator");
    return FEDDB200_ELOGIC;
}

// ---------------------------------------------------------------------------------------
// gather path preparation (lazy, once per pattern): canonical position map, row types,
// (type, length) buckets
// ---------------------------------------------------------------------------------------
// the star kernel (k_star: all fan and task tiles of an assembly in one address-ordered pass) is an alternative path of 3D P2,
// selected with FEDDB200_STAR=1.  Measured on B200 (config 3): 3.4 ms against 2.49 ms for one launch per bucket -- the three
// tile bodies do not fit the instruction caches together (ncu: 3.3 no-instruction stall cycles per issue), which costs more
// than the address-ordered writes gain.
bool star_enabled()
{
    static const bool on = [] { const char *f = getenv("FEDDB200_STAR"); return f && atoi(f) != 0; }();
    return on;
}

// task programs of the block-task kernel (k_task): tiles of consecutive bucket rows, scheduled on the device
int build_task_programs(feddb200_pat *p, const std::vector<RowInfo> &info)
{
    feddb200_ctx *c = p->ctx;
    static const bool enabled = [] { const char *f = getenv("FEDDB200_TASK"); return !f || atoi(f) != 0; }(); // tuning aid
    if (!enabled || !(p->rm->dim == 3 && p->rm->nloc == 10 && p->cm->nloc == 10)) return FEDDB200_OK;
    std::vector<TaskTile> tiles;
    for (Bucket &b : p->buckets) {
        b.tile_start = 0; b.tile_count = 0; b.npt = 0;
        if (b.type != 0 || b.lcap > kTaskMaxLen) continue;
        int max_ninc = 0;
        for (int64_t q = b.start; q < b.start + b.count; q++) max_ninc = std::max(max_ninc, info[q].ninc);
        if (max_ninc > kTaskMaxTets) continue;
        // row nodes per tile: as many as the element slots allow, but the tile's shared-memory rows (elasticity: 9 lcap + 2
        // doubles per node) stay within the ~5 KB of the interior vertex rows, so small boundary buckets do not set the
        // shared memory per warp (and with it the residency) of the star kernel
        b.npt = std::max(1, std::min({kTaskMaxNodes, kTaskMaxTets / std::max(1, max_ninc), 5120 / (72 * b.lcap + 16)}));
        b.tile_start = (int64_t)tiles.size();
        for (int64_t q = b.start; q < b.start + b.count;) {
            TaskTile t;
            std::memset(&t, 0, sizeof(t));
            t.q0 = (uint32_t)q;
            int n = 0, tets = 0;
            while (q < b.start + b.count && n < b.npt && tets + info[q].ninc <= kTaskMaxTets) { tets += info[q].ninc; n++; q++; }
            t.n_nodes = (uint8_t)n;
            tiles.push_back(t);
        }
        b.tile_count = (int64_t)tiles.size() - b.tile_start;
    }
    const int64_t n_tiles = (int64_t)tiles.size();
    if (n_tiles == 0) return FEDDB200_OK;
    TaskBuildArgs A;
    int *status_d = nullptr;
    FB_CUDA(cudaMalloc(&p->task_tiles_d, sizeof(TaskTile) * n_tiles));
    FB_CUDA(cudaMalloc(&status_d, sizeof(int)));
    FB_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int), c->stream));
    FB_CUDA(cudaMemcpyAsync(p->task_tiles_d, tiles.data(), sizeof(TaskTile) * n_tiles, cudaMemcpyHostToDevice, c->stream));
    A.rowinfo = (const RowInfo *)p->rowinfo_d; A.rec = p->rec_d; A.tiles = (TaskTile *)p->task_tiles_d; A.n_tiles = n_tiles;
    A.tasks = nullptr; A.tileblk = nullptr; A.status = status_d;
    const unsigned grid = (unsigned)std::min<int64_t>((n_tiles + 63) / 64, 148 * 32);
    k_task_build<<<grid, 64, 0, c->stream>>>(A, 0);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    int status = 0;
    FB_CUDA(cudaMemcpyAsync(tiles.data(), p->task_tiles_d, sizeof(TaskTile) * n_tiles, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaMemcpyAsync(&status, status_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    int64_t passes = 0;
    for (TaskTile &t : tiles) { t.task_off = (uint32_t)passes; passes += t.n_passes; }
    for (Bucket &b : p->buckets) {
        b.task_max_tets = 0; b.task_max_passes = 0;
        for (int64_t t = b.tile_start; t < b.tile_start + b.tile_count; t++) {
            b.task_max_tets = std::max<int>(b.task_max_tets, tiles[t].n_tets);
            b.task_max_passes = std::max<int>(b.task_max_passes, tiles[t].n_passes);
        }
    }
    if (status != 0 || passes >= (int64_t(1) << 32)) { // a tile outside the format limits: the buckets keep their other kernels
        for (Bucket &b : p->buckets) b.tile_count = 0;
        cudaFree(status_d);
        return FEDDB200_OK;
    }
    FB_CUDA(cudaMalloc(&p->tasks_d, sizeof(uint64_t) * 32 * std::max<int64_t>(passes, 1)));
    FB_CUDA(cudaMalloc(&p->tiletet_d, (size_t)kTileBlkB * n_tiles));
    FB_CUDA(cudaMemcpyAsync(p->task_tiles_d, tiles.data(), sizeof(TaskTile) * n_tiles, cudaMemcpyHostToDevice, c->stream));
    A.tasks = p->tasks_d; A.tileblk = (uint32_t *)p->tiletet_d;
    k_task_build<<<grid, 64, 0, c->stream>>>(A, 1);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaMemcpyAsync(&status, status_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(status_d);
    FB_LOGIC(status != 0, "task programs: inconsistent schedule");
    return FEDDB200_OK;
}

// padded per-bucket records of the fan kernel (k_fan): ring-ordered edge-node rows of 3D P2
int build_fan_records(feddb200_pat *p, const std::vector<RowInfo> &info)
{
    feddb200_ctx *c = p->ctx;
    // off by default: measured equal to / slower than k_ring on B200; FEDDB200_FAN=1 selects it
    static const bool enabled = [] { const char *f = getenv("FEDDB200_FAN"); return f && atoi(f) != 0; }() || star_enabled();
    if (!enabled || !(p->rm->dim == 3 && p->rm->nloc == 10 && p->cm->nloc == 10)) return FEDDB200_OK;
    int64_t total = 0;   // in tiles
    for (Bucket &b : p->buckets) {
        b.fan_off = 0; b.fan_W = 0; b.fan_npw = 0;
        if (b.type != 1 || b.lcap > 255) continue;
        int max_ninc = 0;
        for (int64_t q = b.start; q < b.start + b.count; q++) max_ninc = std::max(max_ninc, info[q].ninc);
        if (max_ninc > 32) continue;
        b.fan_W = std::max(4, max_ninc);
        b.fan_npw = 32 / b.fan_W;
        b.fan_off = total;
        total += (b.count + b.fan_npw - 1) / b.fan_npw;
    }
    if (total == 0) return FEDDB200_OK;
    FB_CUDA(cudaMalloc(&p->fanrec_d, sizeof(uint32_t) * 8 * 32 * total));
    FB_CUDA(cudaMemsetAsync(p->fanrec_d, 0, sizeof(uint32_t) * 8 * 32 * total, c->stream));
    for (const Bucket &b : p->buckets) {
        if (b.fan_W == 0) continue;
        const int64_t ntiles = (b.count + b.fan_npw - 1) / b.fan_npw;
        k_fan_records<<<(unsigned)std::min<int64_t>((ntiles * b.fan_npw + 127) / 128, 148 * 16), 128, 0, c->stream>>>(
            (const RowInfo *)p->rowinfo_d, b.start, b.count, p->rec_d, b.fan_W, b.fan_npw, ntiles, p->fanrec_d + b.fan_off * 32 * 8);
        c->launches++;
        FB_CUDA(cudaGetLastError());
    }
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FEDDB200_OK;
}

// address-ordered tile lists of the star kernel: every fan tile and task tile of the pattern, sorted by the position of its
// first row in the values array; owned and ghost rows separately (the ghost rows of a multi-GPU assembly run first)
int build_star_tiles(feddb200_pat *p, const std::vector<RowInfo> &info)
{
    feddb200_ctx *c = p->ctx;
    if (!star_enabled() || !(p->rm->dim == 3 && p->rm->nloc == 10 && p->cm->nloc == 10)) return FEDDB200_OK;
    struct Item { int64_t key; uint32_t x, y, z; };
    std::vector<Item> items[2];
    for (Bucket &b : p->buckets) {
        b.in_star = 0;
        if (b.type == 0 && b.tile_count > 0 && p->tasks_d) {
            std::vector<TaskTile> tiles((size_t)b.tile_count);
            FB_CUDA(cudaMemcpy(tiles.data(), (const TaskTile *)p->task_tiles_d + b.tile_start, sizeof(TaskTile) * b.tile_count, cudaMemcpyDeviceToHost));
            for (int64_t t = 0; t < b.tile_count; t++)
                items[b.ghost].push_back({info[tiles[t].q0].base, 0u | (uint32_t)b.npt << 16, (uint32_t)(b.tile_start + t), (uint32_t)b.lcap});
            b.in_star = 1;
        } else if (b.type == 1 && b.fan_W > 0 && p->fanrec_d) {
            const int64_t nt = (b.count + b.fan_npw - 1) / b.fan_npw;
            for (int64_t t = 0; t < nt; t++)
                items[b.ghost].push_back({info[b.start + t * b.fan_npw].base, 1u | (uint32_t)b.fan_W << 8, (uint32_t)(b.fan_off + t), (uint32_t)b.lcap});
            b.in_star = 1;
        }
    }
    for (int g = 0; g < 2; g++) {
        p->star_n[g] = (int64_t)items[g].size();
        if (items[g].empty()) continue;
        std::sort(items[g].begin(), items[g].end(), [](const Item &a, const Item &b) { return a.key < b.key; });
        std::vector<uint32_t> flat(items[g].size() * 4);
        for (size_t i = 0; i < items[g].size(); i++) { flat[4 * i] = items[g][i].x; flat[4 * i + 1] = items[g][i].y; flat[4 * i + 2] = items[g][i].z; flat[4 * i + 3] = 0; }
        FB_CUDA(cudaMalloc(&p->star_tiles_d[g], flat.size() * 4));
        FB_CUDA(cudaMemcpy(p->star_tiles_d[g], flat.data(), flat.size() * 4, cudaMemcpyHostToDevice));
    }
    (void)c;
    return FEDDB200_OK;
}

int ensure_gather(feddb200_pat *p)
{
    if (p->gather_ready) return FEDDB200_OK;
    {   // a previous attempt may have failed half-way (out of memory): drop what it left before building again
        auto drop = [](auto *&q) { if (q) { cudaFree(q); q = nullptr; } };
        drop(p->rec_d); drop(p->rowinfo_d); drop(p->row_perm_d); drop(p->ahead_d); drop(p->task_tiles_d); drop(p->tasks_d);
        drop(p->tiletet_d); drop(p->fanrec_d); drop(p->star_tiles_d[0]); drop(p->star_tiles_d[1]); drop(p->geom_d); drop(p->frag_d);
    }
    feddb200_ctx *c = p->ctx;
    const int dim = p->rm->dim, nl = p->rm->nloc;
    const int64_t n_rows = p->n_rows;
    const int nlc = p->cm->nloc;
    p->rec_words = nlc <= 4 ? 4 : 8;
    FB_LOGIC(p->max_len >= 0xffff, "gather path supports node rows of up to 65534 entries");
    FB_CUDA(cudaMalloc(&p->rec_d, sizeof(uint32_t) * std::max<int64_t>(p->n_inc * p->rec_words, 1)));
    int8_t *rtype_d = nullptr;
    uint64_t *sig_d = nullptr;
    FB_CUDA(cudaMalloc(&rtype_d, std::max<int64_t>(n_rows, 1)));
    FB_CUDA(cudaMalloc(&sig_d, sizeof(uint64_t) * std::max<int64_t>(n_rows, 1)));
    const int use_ring = getenv("FEDDB200_NO_RING") == nullptr; // tuning aid: generic kernel for every edge row
    if (n_rows > 0) {
        const int grid = (int)std::min<int64_t>((n_rows + 127) / 128, 148 * 64);
#define FB_MKREC(D, NR, NC) k_make_records<D, NR, NC><<<grid, 128, 0, c->stream>>>(n_rows, p->inc_ptr_d, p->inc_d, p->rowptr_d, p->pos_d, p->pos_stride, p->rm->conn_d, use_ring, p->rec_d, rtype_d, sig_d)
        const int combo = dim * 10000 + nl * 100 + nlc;
        switch (combo) {
        case 20303: FB_MKREC(2, 3, 3); break;
        case 20606: FB_MKREC(2, 6, 6); break;
        case 30404: FB_MKREC(3, 4, 4); break;
        case 31010: FB_MKREC(3, 10, 10); break;
        case 20306: FB_MKREC(2, 3, 6); break;   // B:   P1 pressure rows x P2 velocity columns
        case 20603: FB_MKREC(2, 6, 3); break;   // B^T
        case 30410: FB_MKREC(3, 4, 10); break;
        case 31004: FB_MKREC(3, 10, 4); break;
        default: set_error("gather path: unsupported combination of row and column elements"); return FEDDB200_ELOGIC;
        }
#undef FB_MKREC
        c->launches++;
        FB_CUDA(cudaGetLastError());
    }
    std::vector<int8_t> rtype(n_rows);
    std::vector<uint64_t> sig(n_rows);
    FB_CUDA(cudaMemcpyAsync(rtype.data(), rtype_d, n_rows, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaMemcpyAsync(sig.data(), sig_d, sizeof(uint64_t) * n_rows, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(rtype_d);
    cudaFree(sig_d);
    // launch order: (row type, capacity) buckets; inside a bucket rows of the same stencil class (equal
    // signature) are adjacent, so the threads of a warp address their accumulator rows identically
    std::vector<int32_t> perm(n_rows);
    for (int64_t r = 0; r < n_rows; r++) perm[r] = (int32_t)r;
    // FEDDB200_MERGE_RING=1 (tuning aid): all ring-ordered edge rows in ONE bucket, so one launch sweeps 72 % of the matrix in
    // address order instead of one launch per row length
    const bool merge_ring = FB_ENV_INT("FEDDB200_MERGE_RING", 0) != 0;
    int ring_cap = 4;
    if (merge_ring)
        for (int64_t r = 0; r < n_rows; r++)
            if ((rtype[r] & 3) == 1) ring_cap = std::max(ring_cap, ((int)(p->rowptr_h[r + 1] - p->rowptr_h[r]) + 3) & ~3);
    auto cap = [&](int64_t r) {
        if (merge_ring && (rtype[r] & 3) == 1) return ring_cap;
        const int l = (int)(p->rowptr_h[r + 1] - p->rowptr_h[r]);
        return std::max(4, (l + 3) & ~3);
    };
    auto key = [&](int32_t r) { return (int64_t)(r >= p->n_owned ? 0 : 1) * 1000000 + (int64_t)(rtype[r] & 3) * 100000 + cap(r); };
    // ... but only inside chunks of consecutive rows, so that a launch still sweeps the mesh (and the geometry
    // lines in L2) once instead of once per stencil class
    int chunk_shift = 12;
    if (const char *f = getenv("FEDDB200_CLASS_CHUNK_SHIFT")) chunk_shift = atoi(f); // tuning aid (<0: no class sort)
    std::sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) {
        const int64_t ka = key(a), kb = key(b);
        if (ka != kb) return ka < kb;
        if (chunk_shift >= 0) {
            const int32_t ca = a >> chunk_shift, cb = b >> chunk_shift;
            if (ca != cb) return ca < cb;
            if (sig[a] != sig[b]) return sig[a] < sig[b];
        }
        return a < b;
    });
    p->buckets.clear();
    for (int64_t s = 0; s < n_rows;) {
        int64_t e = s;
        while (e < n_rows && key(perm[e]) == key(perm[s])) e++;
        p->buckets.push_back({(int)(rtype[perm[s]] & 3), perm[s] >= p->n_owned ? 1 : 0, cap(perm[s]), s, e - s});
        s = e;
    }
    p->bucket_order.resize(p->buckets.size());
    for (size_t i = 0; i < p->buckets.size(); i++) p->bucket_order[i] = (int)i;
    std::stable_sort(p->bucket_order.begin(), p->bucket_order.end(), [&](int x, int y) {
        return p->buckets[x].count * (int64_t)p->buckets[x].lcap > p->buckets[y].count * (int64_t)p->buckets[y].lcap; });
    {
        std::vector<int64_t> inc_ptr(n_rows + 1);
        FB_CUDA(cudaMemcpy(inc_ptr.data(), p->inc_ptr_d, sizeof(int64_t) * (n_rows + 1), cudaMemcpyDeviceToHost));
        std::vector<RowInfo> info(n_rows);
        for (int64_t q = 0; q < n_rows; q++) {
            const int32_t r = perm[q];
            info[q].base = p->rowptr_h[r];
            info[q].k0 = inc_ptr[r];
            info[q].len = (int32_t)(p->rowptr_h[r + 1] - p->rowptr_h[r]);
            info[q].ninc = (int32_t)(inc_ptr[r + 1] - inc_ptr[r]);
            // bits 0-1 row type, bit 4: the row has positions without a local contribution, bit 5: last row of the fragment
            // protocol (its tail is stored directly), bits 32-63: the row
            info[q].pad = (int64_t)(uint8_t)rtype[r] | (r == p->n_owned - 1 ? 32 : 0) | (int64_t)r << 32;
            for (int j = 0; j < 8; j++) info[q].e[j] = 0;
        }
        FB_CUDA(cudaMalloc(&p->rowinfo_d, sizeof(RowInfo) * std::max<int64_t>(n_rows, 1)));
        FB_CUDA(cudaMemcpy(p->rowinfo_d, info.data(), sizeof(RowInfo) * n_rows, cudaMemcpyHostToDevice));
        if (dim == 3 && nl == 10 && nlc == 10 && n_rows > 0) { // 3D P2: look-ahead elements of k_gather_s
            // (+ 4: k_gather_s copies the aligned 16-byte chunk that holds ahead[k])
            FB_CUDA(cudaMalloc(&p->ahead_d, sizeof(uint32_t) * (p->n_inc + 4)));
            FB_CUDA(cudaMemsetAsync(p->ahead_d, 0, sizeof(uint32_t) * (p->n_inc + 4), c->stream));
            k_ring_ahead<<<(unsigned)std::min<int64_t>((n_rows + 127) / 128, 148 * 64), 128, 0, c->stream>>>(n_rows, (RowInfo *)p->rowinfo_d, p->rec_d, p->ahead_d);
            c->launches++;
            FB_CUDA(cudaGetLastError());
            FB_CUDA(cudaStreamSynchronize(c->stream));
        }
        const int rc_task = build_task_programs(p, info);
        if (rc_task != FEDDB200_OK) return rc_task;
        const int rc_fan = build_fan_records(p, info);
        if (rc_fan != FEDDB200_OK) return rc_fan;
        const int rc_star = build_star_tiles(p, info);
        if (rc_star != FEDDB200_OK) return rc_star;
    }
    const int gs = dim == 3 ? 16 : 8;
    FB_CUDA(cudaMalloc(&p->geom_d, sizeof(double) * std::max<int64_t>(p->rm->ne * gs, 1)));
    {   // fragment protocol of the write-out (kernels.cuh): needs runs of at least 8 doubles
        int64_t min_len = 1 << 30;
        for (int64_t r = 0; r < n_rows; r++) min_len = std::min<int64_t>(min_len, p->rowptr_h[r + 1] - p->rowptr_h[r]);
        if (n_rows > 1 && min_len >= 8) FB_CUDA(cudaMalloc(&p->frag_d, sizeof(double) * 8 * n_rows));
    }
    p->gather_ready = true;
    return FEDDB200_OK;
}

// canonical coefficient table from the operator's own quadrature rule (Laplace / elasticity use the
// Grad-Grad degree): r[type][jc][s][t] = sum_q w_q c_row,s(q) c_col,t(q), where c are the
// coefficients of grad(lambda_v) in the basis gradients:  vertex v: 4 lambda_v - 1 on v;
// edge (p,q): 4 lambda_q on p, 4 lambda_p on q;  P1: 1.
void canon_table(const OpTables &t, int dim, int nl, CanonR &R)
{
    std::memset(&R, 0, sizeof(R));
    const int nv = dim + 1;
    const bool p2 = nl > nv;
    static const int E2[3][2] = {{0, 1}, {1, 2}, {0, 2}};
    static const int E3[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    auto edge = [&](int j, int k) { return dim == 2 ? E2[j - nv][k] : E3[j - nv][k]; };
    for (int type = 0; type < (p2 ? 2 : 1); type++)
        for (int jc = 0; jc < nl; jc++)
            for (int s = 0; s < (type == 0 ? 1 : 2); s++)
                for (int tt = 0; tt < (jc < nv ? 1 : 2); tt++) {
                    double sum = 0.0;
                    for (int q = 0; q < t.nq; q++) {
                        const double *lam = &t.lam[q * 4];
                        double crow, ccol;
                        if (!p2) { crow = 1.0; ccol = 1.0; }
                        else {
                            crow = type == 0 ? 4.0 * lam[0] - 1.0 : 4.0 * lam[1 - s];
                            ccol = jc < nv ? 4.0 * lam[jc] - 1.0 : 4.0 * lam[edge(jc, 1 - tt)];
                        }
                        sum += t.w[q] * crow * ccol;
                    }
                    R.r[type][jc][s][tt] = sum;
                }
}

constexpr int64_t kSmallBucketRows = 4096;   // buckets up to this many rows run on the side stream (see launch_gather_t)

template <int OPG, int DIM, int NL>
int launch_gather_t(feddb200_ctx *c, const feddb200_pat *p, GatherArgs &G)
{
    using S = GatherShape<OPG, DIM>;
    const int64_t blocks_geom = (p->rm->ne + 255) / 256;
    const int phase = c->row_phase;
    const bool want_ghost = phase == FEDDB200_ROWS_ALL || phase == FEDDB200_ROWS_GHOST || phase == FEDDB200_ROWS_GHOST_ONLY;
    const bool want_owned = phase == FEDDB200_ROWS_ALL || phase == FEDDB200_ROWS_OWNED;
    // points path of the scalar Laplace rows of 3D P2 patterns (kernels.cuh: geo_from_points): no geometry lines at all, unless
    // one of the alternative row kernels (tuning aids, they read geometry lines) is switched on
    bool use_pts = false;
    if constexpr (OPG == 0 && DIM == 3 && NL == 10) {
        // OFF by default: parity-green (the GPU suite ran with it), but config 2 measured 3.19 ms with it against 3.08 ms with the
        // geometry lines -- 70 % fewer DRAM reads buy nothing, the Laplace rows are not bandwidth-bound either
        static const bool pts_env = [] { const char *f = getenv("FEDDB200_LAPLACE_POINTS"); return f && atoi(f) != 0; }(); // tuning aid
        use_pts = pts_env && !star_enabled() && p->n_inc > 0 && p->rm->nn < ((int64_t)1 << 32);
        for (const Bucket &b : p->buckets)
            if (b.fan_W > 0 && p->fanrec_d) use_pts = false;
        if (use_pts && !p->vtx_d) {   // once per pattern
            feddb200_pat *pm = const_cast<feddb200_pat *>(p);
            FB_CUDA(cudaMalloc(&pm->vtx_d, sizeof(uint4) * p->n_inc));
            FB_CUDA(cudaMalloc(&pm->coords4_d, sizeof(double) * 4 * std::max<int64_t>(p->rm->nn, 1)));
            k_make_vtx<<<(unsigned)std::min<int64_t>((p->n_inc + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(
                p->n_inc, p->rec_d, p->rm->conn_d, (uint4 *)p->vtx_d);
            c->launches++;
        }
    }
    if (p->rm->ne > 0 && phase != FEDDB200_ROWS_OWNED && phase != FEDDB200_ROWS_GHOST_ONLY) {
        if (use_pts)
            k_coords4<<<(unsigned)std::min<int64_t>((p->rm->nn + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(p->rm->nn, p->rm->coords_d, p->coords4_d);
        else
            k_geom<DIM, NL><<<(unsigned)blocks_geom, 256, 0, c->stream>>>(p->rm->ne, p->rm->conn_d, p->rm->coords_d, p->geom_d);
        c->launches++;
    }
    G.vtx = use_pts ? (const uint4 *)p->vtx_d : nullptr;
    G.coords4 = use_pts ? p->coords4_d : nullptr;
    if (phase == FEDDB200_ROWS_GEOM) { FB_CUDA(cudaGetLastError()); return FEDDB200_OK; }
    const size_t budget = c->smem_optin - 1024;
    // the bucket launches are independent of one another and can be forked over the side streams (measured on
    // B200: slower than back-to-back launches -- concurrent sweeps of the mesh compete for L2 -- so off by default)
    int n_side = std::max(0, std::min(feddb200_ctx::kSide, FB_ENV_INT("FEDDB200_SIDE_STREAMS", 0))); // tuning aid
    if (p->buckets.size() < 2) n_side = 0;
    // ... except the SMALL buckets (boundary and corner rows: a handful of blocks that run for 5-50 us each on an
    // otherwise empty GPU): they go to one side stream and run underneath the large launches
    static const bool small_aside = [] { const char *f = getenv("FEDDB200_SMALL_ASIDE"); return !f || atoi(f) != 0; }(); // tuning aid
    const bool aside = n_side == 0 && small_aside && p->buckets.size() >= 2;
    if (n_side > 0 || aside) {
        FB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        for (int i = 0; i < std::max(n_side, 1); i++) FB_CUDA(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
    }
    // star kernel: all fan / task tiles of the phase in one address-ordered persistent launch (star_kernels.cuh)
    bool star_done = false;
    if constexpr (DIM == 3 && NL == 10) {
        if (star_enabled() && (p->star_n[0] > 0 || p->star_n[1] > 0)) {
            constexpr int TPRs = OPG == 1 ? DIM : 1, NBs = OPG == 1 ? DIM : 1, NVs = OPG == 1 ? 9 : 1;
            size_t wb = 0;
            for (const Bucket &b : p->buckets) {
                if (!b.in_star) continue;
                const size_t pitch = (size_t)((TPRs * NBs * b.lcap + 2) & ~1);
                size_t need;
                if (b.type == 0) need = (size_t)b.npt * pitch * 8 + (size_t)kTaskMaxTets * kTaskVec * 24;
                else need = (std::max<size_t>((size_t)b.fan_npw * pitch, (size_t)3 * NVs * 33 + 1) + 1 & ~(size_t)1) * 8;
                wb = std::max(wb, need);
            }
            wb = (wb + 15) & ~(size_t)15;
            const int nts = 64;
            const size_t smem_s = wb * (nts / 32);
            if (smem_s <= budget) {
                int per_sm = 1;
                { const int rc_k = kernel_cfg(c, k_star<OPG>, nts, smem_s, budget, &per_sm); if (rc_k != FEDDB200_OK) return rc_k; }
                for (int g = 1; g >= 0; g--) {   // ghost rows first
                    if (p->star_n[g] == 0) continue;
                    if ((g == 0 && !want_owned) || (g == 1 && !want_ghost)) continue;
                    StarArgs SA;
                    SA.G = G;
                    SA.G.zero = 0;
                    SA.G.nseg = g ? c->n_ghost_seg : 0;
                    for (int i = 0; i < kMaxGhostSeg; i++) { SA.G.seg_begin[i] = c->ghost_seg_begin[i]; SA.G.seg_ptr[i] = c->ghost_seg_ptr[i]; }
                    SA.G.seg_begin[kMaxGhostSeg] = c->ghost_seg_begin[kMaxGhostSeg];
                    SA.tiles = (const uint4 *)p->star_tiles_d[g]; SA.n_tiles = p->star_n[g];
                    SA.fanrec = (const uint4 *)p->fanrec_d; SA.tileblk = (const uint4 *)p->tiletet_d; SA.tasks = p->tasks_d;
                    SA.warp_bytes = (int)wb;
                    const int64_t blocks_s = std::min<int64_t>((SA.n_tiles + nts / 32 - 1) / (nts / 32), (int64_t)std::max(per_sm, 1) * c->sm_count);
                    k_star<OPG><<<(unsigned)blocks_s, nts, smem_s, c->stream>>>(SA);
                    c->launches++;
                    FB_CUDA(cudaGetLastError());
                }
                star_done = true;
            }
        }
    }
    // fragment protocol of the write-out (kernels.cuh): owned rows of k_ring / k_task / k_gather launches; the alternative
    // row kernels (k_star, k_fan, k_gather_s: tuning aids) and replicated scalar rows store plainly.  OFF by default: parity
    // green, but config 3 measured 2.555 ms with it against 2.486 ms without -- the row kernels are not limited by the write
    // path (FEDDB200_FRAG=1 switches it on)
    static const bool frag_env = [] { const char *f = getenv("FEDDB200_FRAG"); return f && atoi(f) != 0; }(); // tuning aid
    bool frag_on = frag_env && p->frag_d != nullptr && !star_done && !(OPG == 0 && G.vec_dim != 0) && p->n_owned > 1;
    if (frag_on)
        for (const Bucket &b : p->buckets) {
            if (b.ghost) continue;
            if (b.type == 1 && b.fan_W > 0 && p->fanrec_d) frag_on = false;
            if (DIM == 3 && NL == 10 && OPG == 1 && b.type == 0 && !(b.tile_count > 0 && p->tasks_d)) frag_on = false;   // k_gather_s
        }
    int turn = 0;
    for (const int bi : p->bucket_order) {
        const Bucket &b = p->buckets[bi];
        if ((!b.ghost && !want_owned) || (b.ghost && !want_ghost)) continue;
        if (star_done && b.in_star) continue;
        G.frag = (frag_on && !b.ghost) ? p->frag_d : nullptr;
        cudaStream_t st = n_side > 0 ? c->side[turn++ % n_side] : ((aside && b.count <= kSmallBucketRows) ? c->side[0] : c->stream);
        G.zero = 0;
        G.start = b.start; G.count = b.count;
        // ghost rows go straight into their owners' receive buffers (peer memory) when targets are set
        G.nseg = b.ghost ? c->n_ghost_seg : 0;
        for (int i = 0; i < kMaxGhostSeg; i++) { G.seg_begin[i] = c->ghost_seg_begin[i]; G.seg_ptr[i] = c->ghost_seg_ptr[i]; }
        G.seg_begin[kMaxGhostSeg] = c->ghost_seg_begin[kMaxGhostSeg];
        int rc;
        if (b.type == 1) {
            // ring-ordered edge rows (3D P2): one thread per CSR row, one shared-memory row per thread (pitch odd)
            if constexpr (DIM == 3 && NL == 10) {
                constexpr int TPR = OPG == 1 ? DIM : 1;
                constexpr int NBr = OPG == 1 ? DIM : 1;
                if (b.fan_W > 0 && p->fanrec_d) {
                    // fan kernel: one lane per incident element, fan_W lanes per row node (star_kernels.cuh)
                    FanArgs F;
                    F.G = G;
                    F.G.pitch = (TPR * NBr * b.lcap + 2) & ~1;
                    F.fanrec = reinterpret_cast<const uint4 *>(p->fanrec_d + b.fan_off * 32 * 8); F.W = b.fan_W; F.npw = b.fan_npw;
                    F.ntiles = (b.count + b.fan_npw - 1) / b.fan_npw;
                    const int ntf = 64;
                    size_t wd = std::max<size_t>((size_t)b.fan_npw * F.G.pitch, (size_t)3 * (OPG == 1 ? 9 : 1) * 33 + 1);
                    wd = (wd + 1) & ~(size_t)1;
                    const size_t smem_f = (wd * 8 + kFanStageB) * (ntf / 32);
                    if (smem_f <= budget) {
                        auto launch_fan = [&](auto kernel) -> int {
                            int per_sm = 1;
                            { const int rc_k = kernel_cfg(c, kernel, ntf, smem_f, budget, &per_sm); if (rc_k != FEDDB200_OK) return rc_k; }
                            const int64_t blocks_f = std::min<int64_t>((F.ntiles + ntf / 32 - 1) / (ntf / 32), (int64_t)std::max(per_sm, 1) * c->sm_count);
                            kernel<<<(unsigned)blocks_f, ntf, smem_f, st>>>(F);
                            c->launches++;
                            FB_CUDA(cudaGetLastError());
                            return FEDDB200_OK;
                        };
                        rc = b.fan_W == 4 ? launch_fan(k_fan<OPG, 4>) : (b.fan_W == 6 ? launch_fan(k_fan<OPG, 6>) : launch_fan(k_fan<OPG, 0>));
                        if (rc != FEDDB200_OK) return rc;
                        continue;
                    }
                }
                const int nt = std::max(32, std::min(64, FB_ENV_INT("FEDDB200_RING_NT", 64) & ~31)); // tuning aid
                const int npt = TPR == 3 ? FB_RING_NPT3 : 32 / TPR;   // row nodes per warp tile
                // node pitch: room for the TPR dof rows + the phase shift; among the next candidates the one with the fewest
                // bank conflicts when the threads of a half-warp store to the same position of their rows
                const int n_row = NBr * std::max(1, b.lcap - 1), need = TPR * NBr * b.lcap + 1; // typical row: lcap rounds L up to a multiple of 4
                int pitch = need, best_conf = 1 << 30;
                for (int cand = need; cand < need + 16; cand++) {
                    int hist[16] = {0}, conf = 0;
                    for (int t = 0; t < 16; t++) hist[((t / TPR) * cand + (t % TPR) * n_row) & 15]++;
                    for (int k = 0; k < 16; k++) conf += hist[k] > 1 ? hist[k] - 1 : 0;
                    if (conf < best_conf) { best_conf = conf; pitch = cand; }
                }
                const size_t smem = (size_t)pitch * 8 * npt * (nt / 32);
                G.pitch = pitch;
                const int64_t tiles = ((b.count + npt - 1) / npt + nt / 32 - 1) / (nt / 32); // in blocks
                FB_LOGIC(smem > budget, "row too long for the gather path's shared-memory rows; use the coloured or atomic mode");
                // persistent blocks: as many as are resident at once (times a small factor that evens out the tail)
                auto launch_ring = [&](auto kernel) -> int {
                    int per_sm = 1;
                    { const int rc_k = kernel_cfg(c, kernel, nt, smem, budget, &per_sm); if (rc_k != FEDDB200_OK) return rc_k; }
                    const int waves = std::max(1, FB_ENV_INT("FEDDB200_RING_WAVES", 1)); // tuning aid
                    const int64_t blocks = std::min<int64_t>(tiles, (int64_t)std::max(per_sm, 1) * c->sm_count * waves);
                    kernel<<<(unsigned)blocks, nt, smem, st>>>(G);
                    c->launches++;
                    FB_CUDA(cudaGetLastError());
                    return FEDDB200_OK;
                };
                if constexpr (OPG == 0) rc = use_pts ? launch_ring(k_ring<0, true>) : launch_ring(k_ring<0, false>);
                else rc = launch_ring(k_ring<OPG, false>);
            } else { set_error("ring rows exist for 3D P2 patterns only"); rc = FEDDB200_ELOGIC; }
        } else {
            if constexpr (DIM == 3 && NL == 10) {
                // vertex-node rows of 3D P2: block-task kernel (star_kernels.cuh), one warp per tile of row nodes
                // (elasticity only: for the scalar Laplace operator the task kernel measured slower than k_gather, 1.81 vs 1.63 ms at M = 80)
                if (OPG == 1 && b.type == 0 && b.tile_count > 0 && p->tasks_d) {
                    constexpr int TPRt = OPG == 1 ? DIM : 1, NBt = OPG == 1 ? DIM : 1;
                    TaskArgs T;
                    T.G = G;
                    T.G.pitch = (TPRt * NBt * b.lcap + 2) & ~1;   // room for the phase shift, even (keeps the stage planes 16-byte aligned)
                    T.tileblk = reinterpret_cast<const uint4 *>((const char *)p->tiletet_d + (size_t)b.tile_start * kTileBlkB); T.n_tiles = b.tile_count;
                    T.tasks = p->tasks_d; T.npt = b.npt;
                    T.max_tets = (std::max(1, b.task_max_tets) + 1) & ~1 /* even: keeps the areas 16-byte aligned */; T.max_passes = std::max(1, b.task_max_passes);
                    const int ntk = 64;
                    const size_t smem_t = task_warp_bytes(b.npt, T.G.pitch, T.max_tets, T.max_passes) * (ntk / 32);
                    if (smem_t <= budget) {
                        int per_sm = 1;
                        { const int rc_k = kernel_cfg(c, k_task<OPG>, ntk, smem_t, budget, &per_sm); if (rc_k != FEDDB200_OK) return rc_k; }
                        const int64_t blocks_t = std::min<int64_t>((b.tile_count + ntk / 32 - 1) / (ntk / 32), (int64_t)std::max(per_sm, 1) * c->sm_count);
                        k_task<OPG><<<(unsigned)blocks_t, ntk, smem_t, st>>>(T);
                        c->launches++;
                        FB_CUDA(cudaGetLastError());
                        continue;
                    }
                }
            }
            // accumulators: 32 rows of NBL*L doubles per block, `pitch` doubles apart with pitch == NBL (mod 16), see k_gather
            int pitch = S::NBL * b.lcap;
            while ((pitch & 15) != (S::NBL & 15)) pitch++;
            const size_t smem = (size_t)pitch * 8 * 32;
            FB_LOGIC(smem > budget, "row too long for the gather path's shared-memory accumulators; use the coloured or atomic mode");
            const int nt = S::NT;
            G.pitch = pitch;
            if constexpr (OPG == 1 && DIM == 3 && NL == 10) {
                // vertex-node rows of 3D P2 elasticity: streamed inputs (k_gather_s), 3 row nodes per warp
                static const bool stream_ok = [] { const char *f = getenv("FEDDB200_GATHER_STREAM"); return !f || atoi(f) != 0; }(); // tuning aid
                const int nts = std::max(32, std::min(128, FB_ENV_INT("FEDDB200_GATHER_NT", 64) & ~31)); // tuning aid
                const size_t smem_s = ((((size_t)pitch * 8 * 9 + 15) & ~(size_t)15) + 3 * kRsNodeB) * (nts / 32);
                if (stream_ok && b.type == 0 && p->ahead_d && smem_s <= budget) {
                    int per_sm = 1;
                    { const int rc_k = kernel_cfg(c, k_gather_s, nts, smem_s, budget, &per_sm); if (rc_k != FEDDB200_OK) return rc_k; }
                    const int64_t tiles = ((b.count + 2) / 3 + nts / 32 - 1) / (nts / 32); // in blocks
                    const int64_t blocks_s = std::min<int64_t>(tiles, (int64_t)std::max(per_sm, 1) * c->sm_count);
                    k_gather_s<<<(unsigned)blocks_s, nts, smem_s, st>>>(G);
                    c->launches++;
                    FB_CUDA(cudaGetLastError());
                    continue;
                }
            }
            const int64_t blocks = (b.count * S::CPR + nt - 1) / nt;
            auto launch = [&](auto kernel) -> int {
                { const int rc_k = kernel_cfg(c, kernel, 0, 0, budget, nullptr); if (rc_k != FEDDB200_OK) return rc_k; }
                kernel<<<(unsigned)blocks, nt, smem, st>>>(G);
                c->launches++;
                FB_CUDA(cudaGetLastError());
                return FEDDB200_OK;
            };
            if constexpr (OPG == 0 && DIM == 3 && NL == 10) {
                if (use_pts) rc = b.type == 0 ? launch(k_gather<0, 3, 10, 0, true>) : launch(k_gather<0, 3, 10, 1, true>);
                else rc = b.type == 0 ? launch(k_gather<0, 3, 10, 0>) : launch(k_gather<0, 3, 10, 1>);
            } else
            if (b.type == 0) rc = launch(k_gather<OPG, DIM, NL, 0>);
            else {
                if constexpr (NL > DIM + 1) rc = launch(k_gather<OPG, DIM, NL, 1>);
                else { set_error("edge-node row in a P1 pattern"); rc = FEDDB200_ELOGIC; }
            }
        }
        if (rc != FEDDB200_OK) return rc;
    }
    for (int i = 0; i < std::max(n_side, aside ? 1 : 0); i++) {
        FB_CUDA(cudaEventRecord(c->ev_join[i], c->side[i]));
        FB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
    }
    if (frag_on && want_owned) {
        const int64_t n_bound = p->n_owned - 1;
        k_stitch<<<(unsigned)std::min<int64_t>((n_bound + 255) / 256, (int64_t)c->sm_count * 8), 256, 0, c->stream>>>(
            n_bound, p->rowptr_d, OPG == 1 ? DIM * DIM : 1, G.values, p->frag_d);
        c->launches++;
        FB_CUDA(cudaGetLastError());
    }
    return FEDDB200_OK;
}

template <int OPG>
int launch_gather(feddb200_ctx *c, const feddb200_pat *p, GatherArgs &G)
{
    switch (elem_index(p->rm->dim, p->rm->nloc)) {
    case 0: return launch_gather_t<OPG, 2, 3>(c, p, G);
    case 1: return launch_gather_t<OPG, 2, 6>(c, p, G);
    case 2: return launch_gather_t<OPG, 3, 4>(c, p, G);
    case 3: return launch_gather_t<OPG, 3, 10>(c, p, G);
    }
    set_error("unsupported element");
    return FEDDB200_ELOGIC;
}

// coefficient tensors of the operator gather kernels (kernels.cuh: OpCoef), from the operators' own quadrature
// tables: c_{jt} = coefficient of G_t in grad phi_j (P2 vertex v: 4 lambda_v - 1 on v; edge (p,q): 4 lambda_q on p,
// 4 lambda_p on q; P1: 1); canonical row node i' = vertex 0 (type 0) or edge (0,1) (type 1).
void op_coefficients(const OpTables &t, int dim, int nl_vel, OpCoef &C, int what /*0 N+W, 1 B, 2 BT, 3 mass*/)
{
    const int nv = dim + 1;
    const bool p2 = nl_vel > nv;
    static const int E2[3][2] = {{0, 1}, {1, 2}, {0, 2}};
    static const int E3[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    auto sv = [&](int j, int k) { return j < nv ? j : (dim == 2 ? E2[j - nv][k] : E3[j - nv][k]); };
    auto cgrad = [&](const double *lam, int j, int tt) { // coefficient of G_{sv(j,tt)} in grad phi_j
        if (!p2) return 1.0;
        return j < nv ? 4.0 * lam[j] - 1.0 : 4.0 * lam[sv(j, 1 - tt)];
    };
    const int ntypes = p2 ? 2 : 1;
    for (int q = 0; q < t.nq; q++) {
        const double *lam = &t.lam[q * 4];
        const double w = t.w[q];
        if (what == 0) {
            const double *phi = &t.phi[q * t.np];
            for (int type = 0; type < ntypes; type++) {
                const int ir = type == 0 ? 0 : nv;
                for (int j = 0; j < nl_vel; j++)
                    for (int v = 0; v < nv; v++) C.MW[type][j][v] += w * lam[v] * phi[ir] * phi[j];
            }
        } else if (what == 3) {
            const double *phi = &t.phi[q * t.np];
            for (int type = 0; type < ntypes; type++)
                for (int j = 0; j < nl_vel; j++) C.MM[type][j] += w * phi[type == 0 ? 0 : nv] * phi[j];
        } else if (what == 1) {
            const double psi0 = t.phi[q * t.np + 0];
            for (int j = 0; j < nl_vel; j++)
                for (int tt = 0; tt < (j < nv ? 1 : 2); tt++) C.BC[j][tt] += w * psi0 * cgrad(lam, j, tt);
        } else {
            for (int type = 0; type < ntypes; type++) {
                const int ir = type == 0 ? 0 : nv;
                for (int j = 0; j < t.np && j < 4; j++)
                    for (int s2 = 0; s2 < (type == 0 ? 1 : 2); s2++) C.BTC[type][j][s2] += w * t.phi[q * t.np + j] * cgrad(lam, ir, s2);
            }
        }
    }
}

// coefficient tensors of k_sloc (kernels.cuh), natural local order, from the operators' own quadrature rules:
//   TNF[i][m][j][t] = sum_q w phi_i phi_m c_{jt}   (advection rule)      RLF[i][j][s][t] = sum_q w c_{is} c_{jt}   (Grad-Grad rule)
// with c_{jt} the coefficient of G_sv(j,t) in grad phi_j (P2 vertex v: 4 lambda_v - 1; edge (p,q): 4 lambda_q on p, 4 lambda_p
// on q; P1: 1).  Layout: TNF[nl][nl][nl][2] | RLF[nl][nl][2][2].
void sloc_tables(const OpTables &ta, const OpTables *tl, int dim, int nl, std::vector<double> &out)
{
    const int nv = dim + 1;
    const bool p2 = nl > nv;
    static const int E2[3][2] = {{0, 1}, {1, 2}, {0, 2}};
    static const int E3[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    auto sv = [&](int j, int k) { return j < nv ? j : (dim == 2 ? E2[j - nv][k] : E3[j - nv][k]); };
    auto cgrad = [&](const double *lam, int j, int tt) {
        if (!p2) return 1.0;
        return j < nv ? 4.0 * lam[j] - 1.0 : 4.0 * lam[sv(j, 1 - tt)];
    };
    const size_t ntn = (size_t)nl * nl * nl * 2, nrl = (size_t)nl * nl * 4;
    out.assign(ntn + nrl, 0.0);
    for (int q = 0; q < ta.nq; q++) {
        const double *lam = &ta.lam[q * 4], *phi = &ta.phi[q * ta.np];
        for (int i = 0; i < nl; i++)
            for (int m = 0; m < nl; m++)
                for (int j = 0; j < nl; j++)
                    for (int tt = 0; tt < (j < nv ? 1 : 2); tt++)
                        out[(((size_t)i * nl + m) * nl + j) * 2 + tt] += ta.w[q] * phi[i] * phi[m] * cgrad(lam, j, tt);
    }
    if (tl)
        for (int q = 0; q < tl->nq; q++) {
            const double *lam = &tl->lam[q * 4];
            for (int i = 0; i < nl; i++)
                for (int j = 0; j < nl; j++)
                    for (int ss = 0; ss < (i < nv ? 1 : 2); ss++)
                        for (int tt = 0; tt < (j < nv ? 1 : 2); tt++)
                            out[ntn + (((size_t)i * nl + j) * 2 + ss) * 2 + tt] += tl->w[q] * cgrad(lam, i, ss) * cgrad(lam, j, tt);
        }
}

template <int OPX, int DIM, int NLR, int NL>
int launch_gatherx_t(feddb200_ctx *c, feddb200_pat *p, const feddb200_mesh *vm, const double *u_d, double *values_d, int vec_dim,
                     const OpCoef &C)
{
    using S = OpXShape<OPX, DIM>;
    constexpr int NLV = OPX == X_BT ? NLR : NL; // nodes of the velocity space
    const int64_t ne = p->rm->ne;
    const int phase = c->row_phase;
    const bool want_ghost = phase == FEDDB200_ROWS_ALL || phase == FEDDB200_ROWS_GHOST || phase == FEDDB200_ROWS_GHOST_ONLY;
    const bool want_owned = phase == FEDDB200_ROWS_ALL || phase == FEDDB200_ROWS_OWNED;
    // element pre-passes (geometry, |det| grad u, scalar local matrices): once per assembly, in its first phase
    if (ne > 0 && phase != FEDDB200_ROWS_OWNED && phase != FEDDB200_ROWS_GHOST_ONLY) {
        k_geom<DIM, NLV><<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(ne, vm->conn_d, vm->coords_d, p->geom_d);
        c->launches++;
        if constexpr (S::NEEDS_U) {
            if (!p->dt_d) FB_CUDA(cudaMalloc(&p->dt_d, sizeof(double) * 4 * DIM * DIM * ne));
            if constexpr (OPX != X_ADV) {   // |det| grad u at the vertices: W(u)
                k_udata<DIM, NLV><<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(ne, vm->conn_d, p->geom_d, u_d, nullptr, p->dt_d, 1);
                c->launches++;
            }
            if constexpr (OPX != X_ADVU) {  // scalar local matrices: N(u) [+ the viscous term]
                if (!p->sloc_d) FB_CUDA(cudaMalloc(&p->sloc_d, sizeof(double) * NLV * NLV * ne));
                if (!p->sloc_tab_d) {       // once per pattern
                    OpTables ta, tl;
                    if (build_tables(ta, OP_ADV, DIM, NLV, NLV) != 0 || build_tables(tl, OP_LAP, DIM, NLV, NLV) != 0) { set_error("no tables"); return FEDDB200_ELOGIC; }
                    std::vector<double> tab;
                    sloc_tables(ta, &tl, DIM, NLV, tab);
                    FB_CUDA(cudaMalloc(&p->sloc_tab_d, sizeof(double) * tab.size()));
                    FB_CUDA(cudaMemcpy(p->sloc_tab_d, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
                }
                const int nts = std::max(32, std::min(128, FB_ENV_INT("FEDDB200_SLOC_NT", 128) & ~31)); // tuning aid
                const size_t smem_s = sizeof(double) * ((size_t)NLV * NLV * NLV * 2 + (size_t)NLV * NLV * 4 + (size_t)16 * nts);
                k_sloc<DIM, NLV><<<(unsigned)((ne + nts - 1) / nts), nts, smem_s, c->stream>>>(
                    ne, vm->conn_d, p->geom_d, u_d, p->sloc_tab_d, OPX == X_NSJ ? C.c0 : 0.0, OPX == X_NSJ ? C.c1 : 1.0, p->sloc_d);
                c->launches++;
            }
            FB_CUDA(cudaGetLastError());
        }
    }
    if (phase == FEDDB200_ROWS_GEOM) return FEDDB200_OK;
    GatherXArgs G;
    G.rowinfo = (const RowInfo *)p->rowinfo_d; G.rec = p->rec_d; G.geom = p->geom_d; G.sloc = p->sloc_d; G.dt = p->dt_d;
    G.values = values_d; G.vec_dim = vec_dim; G.C = C;
    const size_t budget = c->smem_optin - 1024;
    // small buckets on a side stream, underneath the large launches (see launch_gather_t)
    static const bool small_aside = [] { const char *f = getenv("FEDDB200_SMALL_ASIDE"); return !f || atoi(f) != 0; }(); // tuning aid
    const bool aside = small_aside && p->buckets.size() >= 2;
    if (aside) {
        FB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        FB_CUDA(cudaStreamWaitEvent(c->side[0], c->ev_fork, 0));
    }
    static const bool percomp = [] { const char *f = getenv("FEDDB200_GATHERW"); return !f || atoi(f) != 0; }(); // tuning aid
    for (const int bi : p->bucket_order) {
        const Bucket &b = p->buckets[bi];
        if ((!b.ghost && !want_owned) || (b.ghost && !want_ghost)) continue;
        cudaStream_t st = (aside && b.count <= kSmallBucketRows) ? c->side[0] : c->stream;
        G.start = b.start; G.count = b.count;
        if constexpr (OPX == X_ADVU || OPX == X_NSJ) {
            // W(u) / Navier-Stokes block: per-component threads (k_gatherw), 32 dof rows per block
            int pitch = DIM * b.lcap;
            while ((pitch & 15) != (DIM & 15)) pitch++;
            const size_t smem_w = (size_t)pitch * 8 * 32;
            if (percomp && smem_w <= budget) {
                G.pitch = pitch;
                const int ntw = 32 * DIM;
                const int64_t blocks_w = (b.count * DIM * DIM + ntw - 1) / ntw;
                auto launch_w = [&](auto kernel) -> int {
                    { const int rc_k = kernel_cfg(c, kernel, 0, 0, budget, nullptr); if (rc_k != FEDDB200_OK) return rc_k; }
                    kernel<<<(unsigned)blocks_w, ntw, smem_w, st>>>(G);
                    c->launches++;
                    FB_CUDA(cudaGetLastError());
                    return FEDDB200_OK;
                };
                int rc_w;
                if (b.type == 0) rc_w = launch_w(k_gatherw<OPX, DIM, NL, 0>);
                else {
                    if constexpr (NLR > DIM + 1) rc_w = launch_w(k_gatherw<OPX, DIM, NL, 1>);
                    else { set_error("edge-node row in a P1 row space"); rc_w = FEDDB200_ELOGIC; }
                }
                if (rc_w != FEDDB200_OK) return rc_w;
                continue;
            }
        }
        G.pitch = (S::NB * b.lcap) | 1;
        int nt = 64;
        while (nt > 32 && (size_t)G.pitch * 8 * nt > budget) nt -= 32;
        const size_t smem = (size_t)G.pitch * 8 * nt;
        FB_LOGIC(smem > budget, "row too long for the gather path's shared-memory accumulators; use the coloured or atomic mode");
        const int64_t blocks = (b.count * S::RD + nt - 1) / nt;
        auto launch = [&](auto kernel) -> int {
            { const int rc_k = kernel_cfg(c, kernel, 0, 0, budget, nullptr); if (rc_k != FEDDB200_OK) return rc_k; }
            kernel<<<(unsigned)blocks, nt, smem, st>>>(G);
            c->launches++;
            FB_CUDA(cudaGetLastError());
            return FEDDB200_OK;
        };
        int rc;
        if (b.type == 0) rc = launch(k_gatherx<OPX, DIM, NLR, NL, 0>);
        else {
            if constexpr (NLR > DIM + 1) rc = launch(k_gatherx<OPX, DIM, NLR, NL, 1>);
            else { set_error("edge-node row in a P1 row space"); rc = FEDDB200_ELOGIC; }
        }
        if (rc != FEDDB200_OK) return rc;
    }
    if (aside) {
        FB_CUDA(cudaEventRecord(c->ev_join[0], c->side[0]));
        FB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
    }
    return FEDDB200_OK;
}

// returns 1 if this (operator, element) combination has no gather kernel (the caller falls back to the coloured mode)
int launch_gatherx(feddb200_ctx *c, feddb200_pat *p, int op, const feddb200_mesh *vm, const double *u_d, double c0, double c1,
                   double c2, double *values_d, int vec_dim, int *handled)
{
    const int dim = p->rm->dim, nr = p->rm->nloc, nc = p->cm->nloc;
    const int combo = dim * 10000 + nr * 100 + nc;
    *handled = 0;
    const bool square = (op == OP_ADV || op == OP_ADVU || op == OP_NSJ || op == OP_MASS);
    if (square && !(combo == 20303 || combo == 20606 || combo == 30404 || combo == 31010)) return FEDDB200_OK;
    if (op == OP_B && !(combo == 20306 || combo == 30410 || combo == 20303 || combo == 30404)) return FEDDB200_OK;
    if (op == OP_BT && !(combo == 20603 || combo == 31004 || combo == 20303 || combo == 30404)) return FEDDB200_OK;
    int rc = ensure_gather(p);
    if (rc != FEDDB200_OK) return rc;
    // coefficient tensors: passed to the kernels by value (kernel parameter space)
    OpCoef C;
    std::memset(&C, 0, sizeof(C));
    OpTables t;
    if (op == OP_MASS) {
        if (build_tables(t, OP_MASS, dim, nr, nr) != 0) { set_error("no tables"); return FEDDB200_ELOGIC; }
        op_coefficients(t, dim, nr, C, 3);
    } else if (square) {
        if (build_tables(t, OP_ADV, dim, nr, nr) != 0) { set_error("no tables"); return FEDDB200_ELOGIC; }
        op_coefficients(t, dim, nr, C, 0);
    } else {
        const int nvel = op == OP_B ? nc : nr, npre = op == OP_B ? nr : nc;
        if (build_tables(t, op, dim, nvel, npre) != 0) { set_error("no tables"); return FEDDB200_ELOGIC; }
        op_coefficients(t, dim, nvel, C, op == OP_B ? 1 : 2);
    }
    C.c0 = c0; C.c1 = c1; C.c2 = c2;
    *handled = 1;
#define FB_GX(OPX, D, NR, NC) return launch_gatherx_t<OPX, D, NR, NC>(c, p, vm, u_d, values_d, vec_dim, C)
    switch (op) {
    case OP_ADV:
        switch (combo) { case 20303: FB_GX(X_ADV, 2, 3, 3); case 20606: FB_GX(X_ADV, 2, 6, 6); case 30404: FB_GX(X_ADV, 3, 4, 4); case 31010: FB_GX(X_ADV, 3, 10, 10); }
        break;
    case OP_MASS:
        switch (combo) { case 20303: FB_GX(X_MASS, 2, 3, 3); case 20606: FB_GX(X_MASS, 2, 6, 6); case 30404: FB_GX(X_MASS, 3, 4, 4); case 31010: FB_GX(X_MASS, 3, 10, 10); }
        break;
    case OP_ADVU:
        switch (combo) { case 20303: FB_GX(X_ADVU, 2, 3, 3); case 20606: FB_GX(X_ADVU, 2, 6, 6); case 30404: FB_GX(X_ADVU, 3, 4, 4); case 31010: FB_GX(X_ADVU, 3, 10, 10); }
        break;
    case OP_NSJ:
        switch (combo) { case 20303: FB_GX(X_NSJ, 2, 3, 3); case 20606: FB_GX(X_NSJ, 2, 6, 6); case 30404: FB_GX(X_NSJ, 3, 4, 4); case 31010: FB_GX(X_NSJ, 3, 10, 10); }
        break;
    case OP_B:
        switch (combo) { case 20306: FB_GX(X_B, 2, 3, 6); case 30410: FB_GX(X_B, 3, 4, 10); case 20303: FB_GX(X_B, 2, 3, 3); case 30404: FB_GX(X_B, 3, 4, 4); }
        break;
    case OP_BT:
        switch (combo) { case 20603: FB_GX(X_BT, 2, 6, 3); case 31004: FB_GX(X_BT, 3, 10, 4); case 20303: FB_GX(X_BT, 2, 3, 3); case 30404: FB_GX(X_BT, 3, 4, 4); }
        break;
    }
#undef FB_GX
    set_error("gather path: internal dispatch error");
    return FEDDB200_ELOGIC;
}

// common driver of every assembly call
int run_op(feddb200_ctx *c, const feddb200_pat *pc, int op, const double *u_d, double c0, double c1, double c2,
           int vec_field, double *values_d, const double *coef_d = nullptr)
{
    FB_LOGIC(!c || !pc, "assembly: null argument");
    FB_LOGIC(pc->ctx != c, "assembly: pattern belongs to another context");
    feddb200_pat *p = const_cast<feddb200_pat *>(pc);
    FB_CUDA(cudaSetDevice(c->device));
    const feddb200_mesh *rm = p->rm, *cm = p->cm;
    const int dim = rm->dim, nr = rm->nloc, nc = cm->nloc;
    const bool square = (op != OP_B && op != OP_BT);
    FB_LOGIC(square && nr != nc, "assembly: this operator needs the same FE space for rows and columns");
    FB_LOGIC((op == OP_ADV || op == OP_ADVU || op == OP_NSJ) && !u_d, "assembly: velocity vector is null");
    const feddb200_mesh *vm = (op == OP_B) ? cm : rm; // velocity / geometry mesh
    const int nv = vm->nloc, np = (op == OP_B) ? nr : nc;
    const OpTables *tab_d = nullptr;
    OpTables tab_h;
    int rc = get_tables(c, op, dim, nv, np, &tab_d, &tab_h);
    if (rc != FEDDB200_OK) return rc;

    int64_t factor;
    switch (op) {
    case OP_LAP:  factor = vec_field ? dim : 1; break;
    case OP_MASS: factor = vec_field ? dim : 1; break;
    case OP_ADV:  factor = dim; break;
    case OP_B: case OP_BT: factor = dim; break;
    default:      factor = (int64_t)dim * dim; break;
    }
    const int64_t nnz = factor * p->nnz;
    FB_LOGIC(nnz > 0 && !values_d, "assembly: values pointer is null");
    if (nnz == 0) return FEDDB200_OK;

    int mode = c->mode;
    // a coefficient that varies inside the elements needs the quadrature loop of the element-row kernels
    if (coef_d && mode == FEDDB200_SCATTER_GATHER) mode = FEDDB200_SCATTER_COLOURED;
    // ghost targets (feddb200_set_ghost_targets): only the row-gather kernels of the Laplace / elasticity operators store their
    // ghost rows through out_ptr into the owners' buffers; every other path would leave those buffers untouched and the owners
    // would add stale data -- refuse instead of assembling a wrong matrix
    FB_LOGIC(c->n_ghost_seg > 0 && c->row_phase != FEDDB200_ROWS_OWNED &&
                 !(mode == FEDDB200_SCATTER_GATHER && (op == OP_LAP || op == OP_ELAS) && !coef_d && rm->nloc == cm->nloc),
             "assembly: ghost targets are set, but this operator / scatter mode does not store ghost rows to peer memory; "
             "use the NCCL exchange (assemble_overlapped) for it");
    if (mode == FEDDB200_SCATTER_GATHER && !(op == OP_LAP || op == OP_ELAS)) {
        int handled = 0;
        if (FB_ENV_INT("FEDDB200_NO_GATHERX", 0) == 0) { // tuning aid
            rc = launch_gatherx(c, p, op, vm, u_d, c0, c1, c2, values_d, (op == OP_MASS && vec_field) ? dim : 0, &handled);
            if (rc != FEDDB200_OK || handled) return rc;
        }
        mode = FEDDB200_SCATTER_COLOURED; // no gather kernel for this combination of elements
    }
    if (mode == FEDDB200_SCATTER_GATHER && rm->nloc != cm->nloc) mode = FEDDB200_SCATTER_COLOURED;

    if (mode == FEDDB200_SCATTER_GATHER) {
        rc = ensure_gather(p);
        if (rc != FEDDB200_OK) return rc;
        GatherArgs G;
        G.rowinfo = (const RowInfo *)p->rowinfo_d; G.rec = p->rec_d; G.geom = p->geom_d; G.ahead = p->ahead_d; G.frag = nullptr; G.vtx = nullptr; G.coords4 = nullptr;
        G.c0 = c0; G.c1 = c1; G.values = values_d; G.vec_dim = (op == OP_LAP && vec_field) ? dim : 0;
        canon_table(tab_h, dim, nr, G.R);
        return op == OP_LAP ? launch_gather<0>(c, p, G) : launch_gather<1>(c, p, G);
    }

    if (c->row_phase == FEDDB200_ROWS_OWNED || c->row_phase == FEDDB200_ROWS_GEOM) return FEDDB200_OK; // element-wise modes do everything in the ROWS_GHOST[_ONLY] call
    ElemArgs A;
    A.conn_r = rm->conn_d; A.conn_c = cm->conn_d; A.conn_v = vm->conn_d; A.coords = vm->coords_d;
    A.row_lid = p->row_lid_d; A.rowptr = p->rowptr_d; A.pos = p->pos_d; A.pos_stride = p->pos_stride;
    A.elems = nullptr; A.n_items = rm->ne; A.u = u_d; A.coef = coef_d; A.c0 = c0; A.c1 = c1; A.c2 = c2; A.tab = tab_d;
    A.values = values_d; A.vec_dim = ((op == OP_LAP || op == OP_MASS) && vec_field) ? dim : 0;
    if (nnz > 0) FB_CUDA(cudaMemsetAsync(values_d, 0, sizeof(double) * nnz, c->stream));
    if (mode == FEDDB200_SCATTER_ATOMIC) return launch_elem_op(c, op, dim, nr, nc, A, true);
    rc = ensure_colouring(p);
    if (rc != FEDDB200_OK) return rc;
    for (int k = 0; k < p->n_colours; k++) {
        A.elems = p->colour_perm_d + p->colour_ptr[k];
        A.n_items = p->colour_ptr[k + 1] - p->colour_ptr[k];
        rc = launch_elem_op(c, op, dim, nr, nc, A, false);
        if (rc != FEDDB200_OK) return rc;
    }
    return FEDDB200_OK;
}

__global__ void k_unpack_add(double *__restrict__ values, const double *__restrict__ recv, const int64_t *__restrict__ slot, int64_t n)
{
    // slots of one exchange are distinct per sender but two senders may hit the same slot
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(values + slot[t], recv[t]);
}

// BCBuilder::setLocalRowOne / setLocalRowZero (core/General/BCBuilder_def.hpp:653-709) on the resident CSR values:
// one warp per owned row node; for every dof a selected by the node's mask the dof row is zeroed and, on a diagonal
// block, its diagonal entry set to one.
__global__ void k_dirichlet_rows(int64_t n_owned, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                 const uint8_t *__restrict__ mask, int rd, int cd, int diag_layout, int diagonal_block,
                                 double *__restrict__ values)
{
    const int lane = threadIdx.x & 31;
    const int64_t I = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (I >= n_owned) return;
    const unsigned m = mask[I];
    if (m == 0) return;
    const int64_t b0 = rowptr[I];
    const int L = (int)(rowptr[I + 1] - b0);
    const int per = diag_layout ? 1 : cd;          // values per column node in a dof row
    int pd = -1;                                   // position of the row node's own column (owned node I <-> column I)
    if (diagonal_block) {
        int lo = 0, hi = L;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (colind[b0 + mid] < (int32_t)I) lo = mid + 1; else hi = mid; }
        if (lo < L && colind[b0 + lo] == (int32_t)I) pd = lo;
    }
    for (int a = 0; a < rd; a++) {
        if (!((m >> a) & 1u)) continue;
        double *row = values + (int64_t)rd * per * b0 + (int64_t)a * per * L;
        for (int x = lane; x < per * L; x += 32) row[x] = 0.0;
        __syncwarp();
        if (lane == 0 && pd >= 0) row[diag_layout ? pd : pd * cd + a] = 1.0;
        __syncwarp();
    }
}

// FE::assemblyRHS (FE_def.hpp:4694-4766), constant source: one thread per pattern row (node); the row's incidences
// are walked in ascending element order -- the order in which the reference's element loop adds into valuesRhs[row].
struct RhsArgs {
    int64_t n_rows;
    const int64_t *inc_ptr;
    const int32_t *inc;
    const double *geom;
    int det_off, gs, dofs, vec_field;
    double c[MAXN], f[3];
    double *rhs;
};
__global__ void k_rhs(const RhsArgs A)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= A.n_rows) return;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t k = A.inc_ptr[r]; k < A.inc_ptr[r + 1]; k++) {
        const int32_t code = A.inc[k];
        const double adet = A.geom[(int64_t)(code >> 4) * A.gs + A.det_off];
        double value = A.c[code & 15];
        if (!A.vec_field) acc[0] += value * (adet * A.f[0]);          // value *= absDetB * valueFunc[0]   (:4750)
        else {
            value *= adet;                                            // value *= absDetB                  (:4755)
            for (int d = 0; d < A.dofs; d++) acc[d] += value * A.f[d];
        }
    }
    for (int d = 0; d < A.dofs; d++) A.rhs[r * A.dofs + d] = acc[d];
}

__global__ void k_scale(double *__restrict__ v, int64_t n, double a)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) v[t] *= a;
}

// host-pointer wrapper: grow-only device scratch owned by the context, H2D of u, D2H of the values
int scratch(feddb200_ctx *c, int which, int64_t bytes, double **out)
{
    if ((int64_t)c->scratch_bytes[which] < bytes) {
        if (c->scratch_d[which]) FB_CUDA(cudaFree(c->scratch_d[which]));
        c->scratch_d[which] = nullptr;
        c->scratch_bytes[which] = 0;
        FB_CUDA(cudaMalloc(&c->scratch_d[which], (size_t)std::max<int64_t>(bytes, 8)));
        c->scratch_bytes[which] = (size_t)bytes;
    }
    *out = (double *)c->scratch_d[which];
    return FEDDB200_OK;
}

template <class F>
int with_host_buffers(feddb200_ctx *c, const feddb200_pat *p, int64_t nnz, const double *u, int64_t nu, double *values, F &&run)
{
    FB_LOGIC(!c || !p || (nnz > 0 && !values), "assembly: null argument");
    FB_CUDA(cudaSetDevice(c->device));
    double *v_d = nullptr, *u_d = nullptr;
    int rc = scratch(c, 0, sizeof(double) * nnz, &v_d);
    if (rc != FEDDB200_OK) return rc;
    if (u) {
        rc = scratch(c, 1, sizeof(double) * nu, &u_d);
        if (rc != FEDDB200_OK) return rc;
        FB_CUDA(cudaMemcpyAsync(u_d, u, sizeof(double) * nu, cudaMemcpyHostToDevice, c->stream));
    }
    rc = run(u_d, v_d);
    if (rc == FEDDB200_OK && nnz > 0) {
        FB_CUDA(cudaMemcpyAsync(values, v_d, sizeof(double) * nnz, cudaMemcpyDeviceToHost, c->stream));
        FB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return rc;
}

} // namespace

extern "C" int feddb200_assemble_laplace_d(feddb200_ctx *c, const feddb200_pat *p, int vec_field, double *v)
{
    return run_op(c, p, OP_LAP, nullptr, 0, 0, 0, vec_field, v);
}
extern "C" int feddb200_assemble_mass_d(feddb200_ctx *c, const feddb200_pat *p, int vec_field, double *v)
{
    return run_op(c, p, OP_MASS, nullptr, 0, 0, 0, vec_field, v);
}
extern "C" int feddb200_stress_quadrature(int dim, int nloc, int *nq, double *ref_points, double *weights)
{
    FB_LOGIC(!nq, "stress_quadrature: null argument");
    OpTables t;
    FB_LOGIC(build_tables(t, OP_ELAS, dim, nloc, nloc) != 0, "stress_quadrature: no quadrature rule for this element");
    *nq = t.nq;
    for (int q = 0; q < t.nq; q++) {
        if (ref_points)
            for (int d = 0; d < dim; d++) ref_points[q * dim + d] = t.lam[q * 4 + 1 + d]; // reference point = (lambda_1, .., lambda_dim)
        if (weights) weights[q] = t.w[q];
    }
    return FEDDB200_OK;
}
extern "C" int feddb200_assemble_stress_d(feddb200_ctx *c, const feddb200_pat *p, double coef_const, const double *coef_d, double *v)
{
    // the symmetric-gradient block is the mu-part of the elasticity operator: lambda = 0, mu = the coefficient
    return run_op(c, p, OP_ELAS, nullptr, 0.0, coef_d ? 1.0 : coef_const, 0, 0, v, coef_d);
}
extern "C" int feddb200_assemble_bdstab_d(feddb200_ctx *c, const feddb200_pat *p, double *v)
{
    FB_LOGIC(!p, "null pattern");
    const int dim = p->rm->dim;
    FB_LOGIC(p->rm->nloc != dim + 1 || p->cm->nloc != dim + 1, "Only implemented for P1. Q1 is equivalent but we need to adjust scaling for the reference element.");
    // refElementSize, refElementScale (FE_def.hpp:2183-2190)
    return run_op(c, p, OP_MASS, nullptr, dim == 2 ? 0.5 : 1. / 6., dim == 2 ? 1. / 9. : 1. / 16., 0, 0, v);
}
extern "C" int feddb200_assemble_linelas_d(feddb200_ctx *c, const feddb200_pat *p, double lambda, double mu, double *v)
{
    return run_op(c, p, OP_ELAS, nullptr, lambda, mu, 0, 0, v);
}
extern "C" int feddb200_assemble_advection_d(feddb200_ctx *c, const feddb200_pat *p, const double *u, double *v)
{
    return run_op(c, p, OP_ADV, u, 0, 0, 0, 0, v);
}
extern "C" int feddb200_assemble_advection_in_u_d(feddb200_ctx *c, const feddb200_pat *p, const double *u, double *v)
{
    return run_op(c, p, OP_ADVU, u, 0, 0, 0, 0, v);
}
extern "C" int feddb200_assemble_div_divT_d(feddb200_ctx *c, const feddb200_pat *pB, const feddb200_pat *pBT, double *vB, double *vBT)
{
    FB_LOGIC((vB && !pB) || (vBT && !pBT), "assemble_div_divT: output given without its pattern");
    if (vB) { const int rc = run_op(c, pB, OP_B, nullptr, 0, 0, 0, 0, vB); if (rc) return rc; }
    if (vBT) { const int rc = run_op(c, pBT, OP_BT, nullptr, 0, 0, 0, 0, vBT); if (rc) return rc; }
    return FEDDB200_OK;
}
extern "C" int feddb200_assemble_ns_jacobian_d(feddb200_ctx *c, const feddb200_pat *p, double rho, double nu,
                                               const double *u, int newton, double *v)
{
    return run_op(c, p, OP_NSJ, u, rho * nu, rho, newton ? rho : 0.0, 0, v);
}

extern "C" int feddb200_assemble_laplace(feddb200_ctx *c, const feddb200_pat *p, int vec_field, double *values)
{
    FB_LOGIC(!p, "null pattern");
    const int64_t nnz = (vec_field ? p->rm->dim : 1) * p->nnz;
    return with_host_buffers(c, p, nnz, nullptr, 0, values,
                             [&](double *, double *v_d) { return feddb200_assemble_laplace_d(c, p, vec_field, v_d); });
}
extern "C" int feddb200_assemble_mass(feddb200_ctx *c, const feddb200_pat *p, int vec_field, double *values)
{
    FB_LOGIC(!p, "null pattern");
    const int64_t nnz = (vec_field ? p->rm->dim : 1) * p->nnz;
    return with_host_buffers(c, p, nnz, nullptr, 0, values,
                             [&](double *, double *v_d) { return feddb200_assemble_mass_d(c, p, vec_field, v_d); });
}
extern "C" int feddb200_assemble_stress(feddb200_ctx *c, const feddb200_pat *p, double coef_const, const double *coef, int64_t n_coef,
                                        double *values)
{
    FB_LOGIC(!p, "null pattern");
    if (coef) {
        int nq = 0;
        int rc = feddb200_stress_quadrature(p->rm->dim, p->rm->nloc, &nq, nullptr, nullptr);
        if (rc != FEDDB200_OK) return rc;
        FB_LOGIC(n_coef != p->rm->ne * nq, "assemble_stress: the coefficient array must hold one value per element and quadrature point");
    }
    const int64_t nnz = (int64_t)p->rm->dim * p->rm->dim * p->nnz;
    return with_host_buffers(c, p, nnz, coef, coef ? n_coef : 0, values,
                             [&](double *coef_d, double *v_d) { return feddb200_assemble_stress_d(c, p, coef_const, coef ? coef_d : nullptr, v_d); });
}
extern "C" int feddb200_assemble_bdstab(feddb200_ctx *c, const feddb200_pat *p, double *values)
{
    FB_LOGIC(!p, "null pattern");
    return with_host_buffers(c, p, p->nnz, nullptr, 0, values, [&](double *, double *v_d) { return feddb200_assemble_bdstab_d(c, p, v_d); });
}
extern "C" int feddb200_assemble_linelas(feddb200_ctx *c, const feddb200_pat *p, double lambda, double mu, double *values)
{
    FB_LOGIC(!p, "null pattern");
    const int64_t nnz = (int64_t)p->rm->dim * p->rm->dim * p->nnz;
    return with_host_buffers(c, p, nnz, nullptr, 0, values,
                             [&](double *, double *v_d) { return feddb200_assemble_linelas_d(c, p, lambda, mu, v_d); });
}
extern "C" int feddb200_assemble_advection(feddb200_ctx *c, const feddb200_pat *p, const double *u, double *values)
{
    FB_LOGIC(!p || !u, "null argument");
    return with_host_buffers(c, p, (int64_t)p->rm->dim * p->nnz, u, (int64_t)p->rm->dim * p->rm->nn, values,
                             [&](double *u_d, double *v_d) { return feddb200_assemble_advection_d(c, p, u_d, v_d); });
}
extern "C" int feddb200_assemble_advection_in_u(feddb200_ctx *c, const feddb200_pat *p, const double *u, double *values)
{
    FB_LOGIC(!p || !u, "null argument");
    return with_host_buffers(c, p, (int64_t)p->rm->dim * p->rm->dim * p->nnz, u, (int64_t)p->rm->dim * p->rm->nn, values,
                             [&](double *u_d, double *v_d) { return feddb200_assemble_advection_in_u_d(c, p, u_d, v_d); });
}
extern "C" int feddb200_assemble_div_divT(feddb200_ctx *c, const feddb200_pat *pB, const feddb200_pat *pBT, double *vB, double *vBT)
{
    if (vB) {
        FB_LOGIC(!pB, "null pattern");
        const int rc = with_host_buffers(c, pB, (int64_t)pB->rm->dim * pB->nnz, nullptr, 0, vB, [&](double *, double *v_d) {
            return feddb200_assemble_div_divT_d(c, pB, nullptr, v_d, nullptr);
        });
        if (rc) return rc;
    }
    if (vBT) {
        FB_LOGIC(!pBT, "null pattern");
        const int rc = with_host_buffers(c, pBT, (int64_t)pBT->rm->dim * pBT->nnz, nullptr, 0, vBT, [&](double *, double *v_d) {
            return feddb200_assemble_div_divT_d(c, nullptr, pBT, nullptr, v_d);
        });
        if (rc) return rc;
    }
    return FEDDB200_OK;
}
extern "C" int feddb200_assemble_ns_jacobian(feddb200_ctx *c, const feddb200_pat *p, double rho, double nu, const double *u,
                                             int newton, double *values)
{
    FB_LOGIC(!p || !u, "null argument");
    return with_host_buffers(c, p, (int64_t)p->rm->dim * p->rm->dim * p->nnz, u, (int64_t)p->rm->dim * p->rm->nn, values,
                             [&](double *u_d, double *v_d) { return feddb200_assemble_ns_jacobian_d(c, p, rho, nu, u_d, newton, v_d); });
}

// ---- peer memory (one process per GPU, CUDA IPC) and ghost-row targets -----------------------------------
extern "C" int feddb200_ipc_alloc(feddb200_ctx *c, int64_t bytes, void **ptr_d, unsigned char *handle64)
{
    FB_LOGIC(!c || !ptr_d || !handle64 || bytes < 0, "ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaMalloc(ptr_d, (size_t)std::max<int64_t>(bytes, 256)));
    cudaIpcMemHandle_t h;
    FB_CUDA(cudaIpcGetMemHandle(&h, *ptr_d));
    std::memcpy(handle64, &h, 64);
    return FEDDB200_OK;
}

extern "C" int feddb200_ipc_open(feddb200_ctx *c, const unsigned char *handle64, void **ptr_d)
{
    FB_LOGIC(!c || !ptr_d || !handle64, "ipc_open: bad arguments");
    FB_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    FB_CUDA(cudaIpcOpenMemHandle(ptr_d, h, cudaIpcMemLazyEnablePeerAccess));
    return FEDDB200_OK;
}

extern "C" int feddb200_ipc_close(feddb200_ctx *c, void *ptr_d)
{
    FB_LOGIC(!c, "ipc_close: bad arguments");
    if (!ptr_d) return FEDDB200_OK;
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaIpcCloseMemHandle(ptr_d));
    return FEDDB200_OK;
}

extern "C" int feddb200_set_ghost_targets(feddb200_ctx *c, int nseg, const int64_t *seg_begin, void *const *seg_ptr_d)
{
    FB_LOGIC(!c || nseg < 0 || nseg > kMaxGhostSeg || (nseg > 0 && (!seg_begin || !seg_ptr_d)), "set_ghost_targets: bad arguments");
    for (int i = 0; i < nseg; i++) {
        FB_LOGIC(seg_begin[i + 1] < seg_begin[i], "set_ghost_targets: segment offsets must ascend");
        FB_LOGIC(seg_begin[i + 1] > seg_begin[i] && !seg_ptr_d[i], "set_ghost_targets: non-empty segment without a target");
    }
    c->n_ghost_seg = nseg;
    for (int i = 0; i <= kMaxGhostSeg; i++) c->ghost_seg_begin[i] = nseg > 0 ? seg_begin[std::min(i, nseg)] : 0; // unused tail: empty
    for (int i = 0; i < kMaxGhostSeg; i++) c->ghost_seg_ptr[i] = i < nseg ? static_cast<double *>(seg_ptr_d[i]) : nullptr;
    return FEDDB200_OK;
}

extern "C" int feddb200_unpack_add_d(feddb200_ctx *c, double *values_d, const double *recv_d, const int64_t *slot_d, int64_t n)
{
    FB_LOGIC(!c || (n > 0 && (!values_d || !recv_d || !slot_d)) || n < 0, "unpack_add: bad arguments");
    if (n == 0) return FEDDB200_OK;
    FB_CUDA(cudaSetDevice(c->device));
    k_unpack_add<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, c->stream>>>(values_d, recv_d, slot_d, n);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

extern "C" int feddb200_assemble_rhs_d(feddb200_ctx *c, const feddb200_pat *pc, int vec_field, int deg_func,
                                       const double *value_func, double *rhs_d)
{
    FB_LOGIC(!c || !pc || !value_func || !rhs_d, "assemble_rhs: null argument");
    feddb200_pat *p = const_cast<feddb200_pat *>(pc);
    const feddb200_mesh *m = p->rm;
    FB_CUDA(cudaSetDevice(c->device));
    RhsArgs A;
    std::memset(&A, 0, sizeof(A));
    FB_LOGIC(rhs_coefficients(m->dim, m->nloc, deg_func, A.c) != 0, "assemble_rhs: no quadrature rule for this FE type and function degree");
    if (p->n_rows == 0) return FEDDB200_OK;
    const int gs = m->dim == 3 ? 16 : 8;
    if (!p->geom_d) FB_CUDA(cudaMalloc(&p->geom_d, sizeof(double) * std::max<int64_t>(m->ne * gs, 1)));
    if (m->ne > 0) {
        const unsigned blocks = (unsigned)((m->ne + 255) / 256);
        switch (elem_index(m->dim, m->nloc)) {
        case 0: k_geom<2, 3><<<blocks, 256, 0, c->stream>>>(m->ne, m->conn_d, m->coords_d, p->geom_d); break;
        case 1: k_geom<2, 6><<<blocks, 256, 0, c->stream>>>(m->ne, m->conn_d, m->coords_d, p->geom_d); break;
        case 2: k_geom<3, 4><<<blocks, 256, 0, c->stream>>>(m->ne, m->conn_d, m->coords_d, p->geom_d); break;
        case 3: k_geom<3, 10><<<blocks, 256, 0, c->stream>>>(m->ne, m->conn_d, m->coords_d, p->geom_d); break;
        default: set_error("unsupported element"); return FEDDB200_ELOGIC;
        }
        c->launches++;
    }
    A.n_rows = p->n_rows; A.inc_ptr = p->inc_ptr_d; A.inc = p->inc_d; A.geom = p->geom_d;
    A.gs = gs; A.det_off = m->dim == 3 ? 3 : 6;
    A.vec_field = vec_field ? 1 : 0; A.dofs = vec_field ? m->dim : 1;
    for (int d = 0; d < A.dofs; d++) A.f[d] = value_func[d];
    A.rhs = rhs_d;
    k_rhs<<<(unsigned)((p->n_rows + 255) / 256), 256, 0, c->stream>>>(A);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

extern "C" int feddb200_assemble_rhs(feddb200_ctx *c, const feddb200_pat *p, int vec_field, int deg_func, const double *value_func,
                                     double *rhs)
{
    FB_LOGIC(!c || !p || !rhs, "assemble_rhs: null argument");
    const int64_t n = p->n_rows * (vec_field ? p->rm->dim : 1);
    double *r_d = nullptr;
    int rc = scratch(c, 0, sizeof(double) * n, &r_d);
    if (rc != FEDDB200_OK) return rc;
    rc = feddb200_assemble_rhs_d(c, p, vec_field, deg_func, value_func, r_d);
    if (rc == FEDDB200_OK && n > 0) {
        FB_CUDA(cudaMemcpyAsync(rhs, r_d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        FB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return rc;
}

extern "C" int feddb200_set_dirichlet_rows_d(feddb200_ctx *c, const feddb200_pat *p, int rd, int cd, int mode,
                                             const uint8_t *node_mask_d, int diagonal_block, double *values_d)
{
    FB_LOGIC(!c || !p || !node_mask_d || !values_d, "set_dirichlet_rows: null argument");
    FB_LOGIC(!((mode == FEDDB200_BLOCK_SCALAR && rd == 1 && cd == 1) || (mode == FEDDB200_BLOCK_DIAG && rd == cd && rd >= 1 && rd <= 3) ||
               (mode == FEDDB200_BLOCK_FULL && rd >= 1 && rd <= 3 && cd >= 1 && cd <= 3)),
             "set_dirichlet_rows: unsupported dof layout");
    FB_LOGIC(diagonal_block && rd != cd, "set_dirichlet_rows: a diagonal block has as many row as column dofs");
    if (p->n_owned == 0) return FEDDB200_OK;
    FB_CUDA(cudaSetDevice(c->device));
    const int64_t threads = p->n_owned * 32;
    k_dirichlet_rows<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(p->n_owned, p->rowptr_d, p->colind_d, node_mask_d, rd, cd,
                                                                             mode != FEDDB200_BLOCK_FULL, diagonal_block, values_d);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

__global__ void k_dirichlet_rhs(int64_t n, int dofs, const uint8_t *__restrict__ mask, const double *__restrict__ bc, double *__restrict__ rhs)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * dofs; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t node = t / dofs;
        const int a = (int)(t - node * dofs);
        if ((mask[node] >> a) & 1) rhs[t] = bc[t];
    }
}

extern "C" int feddb200_set_dirichlet_rhs_d(feddb200_ctx *c, int64_t n_nodes, int dofs, const uint8_t *node_mask_d, const double *bc_values_d,
                                            double *rhs_d)
{
    FB_LOGIC(!c || n_nodes < 0 || dofs < 1 || dofs > 3 || (n_nodes > 0 && (!node_mask_d || !bc_values_d || !rhs_d)), "set_dirichlet_rhs: bad arguments");
    if (n_nodes == 0) return FEDDB200_OK;
    FB_CUDA(cudaSetDevice(c->device));
    k_dirichlet_rhs<<<(unsigned)std::min<int64_t>((n_nodes * dofs + 255) / 256, 148 * 16), 256, 0, c->stream>>>(n_nodes, dofs, node_mask_d, bc_values_d, rhs_d);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

extern "C" int feddb200_scale_d(feddb200_ctx *c, double *values_d, int64_t n, double alpha)
{
    FB_LOGIC(!c || (n > 0 && !values_d) || n < 0, "scale: bad arguments");
    if (n == 0) return FEDDB200_OK;
    FB_CUDA(cudaSetDevice(c->device));
    k_scale<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, c->stream>>>(values_d, n, alpha);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}
