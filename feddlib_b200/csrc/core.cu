// Context, mesh upload and the one-time device sparsity builder of libfeddb200.so.
//
// The builder replaces what Tpetra does dynamically behind Matrix::insertGlobalValues /
// fillComplete (reference: feddlib/core/LinearAlgebra/Matrix_def.hpp:46-51, 88-92, 192-199;
// semantics restated in SURVEY.md Appendix C): the CSR graph is the set of (row node, col node)
// pairs of all elements, rows in row-map order, columns ascending by column-map local index.
// Everything below runs on the device: key generation -> radix sort -> unique -> rowptr ->
// per-(element,i,j) position-in-row map (the scatter map) -> row->element incidence lists
// (the gather map).  Only the greedy element colouring runs on the host (lazily, coloured
// mode only).
#include <algorithm>
#include <cub/cub.cuh>
#include <cstring>
#include <mutex>

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>
#include <cctype>
#include <cstring>

#include "common.cuh"

namespace fb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
} // namespace fb

using namespace fb;

extern "C" const char *feddb200_last_error(void) { return fb::g_err.c_str(); }

extern "C" int feddb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int feddb200_create(feddb200_ctx **out, int device)
{
    FB_LOGIC(out == nullptr, "feddb200_create: null output pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        fb::set_error("feddb200_create: no CUDA device available (this engine has no CPU fallback)");
        return FEDDB200_ERUNTIME;
    }
    FB_LOGIC(device < 0 || device >= n, "feddb200_create: invalid device ordinal");
    FB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        fb::set_error("feddb200_create: device is not sm_100 (Blackwell); libfeddb200 is built for sm_100a only");
        return FEDDB200_ERUNTIME;
    }
    auto *c = new feddb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    FB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int i = 0; i < feddb200_ctx::kSide; i++) {
        FB_CUDA(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
        FB_CUDA(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    FB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    FB_CUDA(cudaMalloc(&c->tab_d, OP_COUNT * sizeof(OpTables)));
    *out = c;
    return FEDDB200_OK;
}

extern "C" void feddb200_destroy(feddb200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->tab_d);
    cudaFree(c->scratch_d[0]);
    cudaFree(c->scratch_d[1]);
    for (int i = 0; i < feddb200_ctx::kSide; i++) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int feddb200_set_stream(feddb200_ctx *c, void *s)
{
    FB_LOGIC(!c, "null context");
    c->stream = (cudaStream_t)s; // NULL selects the legacy default stream, as in the CUDA runtime
    return FEDDB200_OK;
}

extern "C" int feddb200_use_own_stream(feddb200_ctx *c)
{
    FB_LOGIC(!c, "null context");
    c->stream = c->own_stream;
    return FEDDB200_OK;
}

extern "C" int feddb200_set_scatter_mode(feddb200_ctx *c, int mode)
{
    FB_LOGIC(!c, "null context");
    FB_LOGIC(mode < 0 || mode > 2, "feddb200_set_scatter_mode: mode must be 0 (atomic), 1 (coloured) or 2 (gather)");
    c->mode = mode;
    return FEDDB200_OK;
}
extern "C" int feddb200_get_scatter_mode(const feddb200_ctx *c) { return c ? c->mode : -1; }
extern "C" int feddb200_set_row_phase(feddb200_ctx *c, int phase)
{
    FB_LOGIC(!c, "null context");
    FB_LOGIC(phase < 0 || phase > 4, "feddb200_set_row_phase: phase must be 0 (all), 1 (ghost rows), 2 (owned rows), 3 (geometry) or 4 (ghost rows only)");
    c->row_phase = phase;
    return FEDDB200_OK;
}

extern "C" int feddb200_synchronize(feddb200_ctx *c)
{
    FB_LOGIC(!c, "null context");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FEDDB200_OK;
}
extern "C" int64_t feddb200_launch_count(const feddb200_ctx *c) { return c ? c->launches : -1; }

extern "C" int feddb200_dev_alloc(feddb200_ctx *c, void **p, int64_t bytes)
{
    FB_LOGIC(!c || !p || bytes < 0, "feddb200_dev_alloc: bad arguments");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaMalloc(p, (size_t)std::max<int64_t>(bytes, 8)));
    return FEDDB200_OK;
}
extern "C" int feddb200_dev_free(feddb200_ctx *c, void *p)
{
    FB_LOGIC(!c, "null context");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaFree(p));
    return FEDDB200_OK;
}
extern "C" int feddb200_bind_host_numa(feddb200_ctx *c, int *numa_node)
{
    FB_LOGIC(!c, "feddb200_bind_host_numa: null context");
    if (numa_node) *numa_node = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), c->device) != cudaSuccess) { cudaGetLastError(); return FEDDB200_OK; }
    for (char *q = bus; *q; q++) *q = (char)tolower(*q);
    std::string base = std::string("/sys/bus/pci/devices/") + bus;
    int node = -1;
    if (FILE *f = fopen((base + "/numa_node").c_str(), "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return FEDDB200_OK;
    char list[4096] = {0};
    FILE *f = fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r");
    if (!f) return FEDDB200_OK;
    const bool got = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!got) return FEDDB200_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_cpus = 0;
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {   // "0-31,64-95"
        int a = -1, b = -1;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1)
            for (int x = a; x <= b && x < CPU_SETSIZE; x++) { CPU_SET(x, &set); n_cpus++; }
    }
    if (n_cpus == 0) return FEDDB200_OK;
    cpu_set_t cur;   // keep only CPUs this process may use (cgroup / taskset limits)
    if (sched_getaffinity(0, sizeof(cur), &cur) == 0) {
        cpu_set_t both;
        CPU_AND(&both, &set, &cur);
        if (CPU_COUNT(&both) == 0) return FEDDB200_OK;
        set = both;
    }
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return FEDDB200_OK;
    if (numa_node) *numa_node = node;
    return FEDDB200_OK;
}
extern "C" int feddb200_host_alloc(feddb200_ctx *c, void **p, int64_t bytes)
{
    FB_LOGIC(!c || !p || bytes < 0, "feddb200_host_alloc: bad arguments");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaHostAlloc(p, (size_t)std::max<int64_t>(bytes, 8), cudaHostAllocDefault));
    return FEDDB200_OK;
}
extern "C" int feddb200_host_free(feddb200_ctx *, void *p)
{
    FB_CUDA(cudaFreeHost(p));
    return FEDDB200_OK;
}
extern "C" int feddb200_copy_h2d(feddb200_ctx *c, void *dst, const void *src, int64_t bytes)
{
    FB_LOGIC(!c, "null context");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FEDDB200_OK;
}
extern "C" int feddb200_copy_d2h(feddb200_ctx *c, void *dst, const void *src, int64_t bytes)
{
    FB_LOGIC(!c, "null context");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FEDDB200_OK;
}

// ---------------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------------
static bool valid_elem(int dim, int nloc)
{
    // nloc == 1: the P0 pressure space of assemblyDivAndDivT (one pseudo-node per element; its ids are the element map,
    // FE_def.hpp:1954-1957).  2D only: FE::phi has no P0 case for dim 3 (FE_def.hpp:5037 ff.)
    return (dim == 2 && (nloc == 1 || nloc == 3 || nloc == 6)) || (dim == 3 && (nloc == 4 || nloc == 10));
}

extern "C" int feddb200_mesh_upload(feddb200_ctx *c, feddb200_mesh **out, int dim, int nloc, int64_t ne,
                                    const int32_t *conn, int64_t nn, const double *coords)
{
    FB_LOGIC(!c || !out, "feddb200_mesh_upload: null context/output");
    FB_LOGIC(!valid_elem(dim, nloc),
             "feddb200_mesh_upload: only P1/P2 triangles (3/6 nodes), tetrahedra (4/10 nodes) and the 2D P0 pseudo-mesh (1 node per element) are implemented");
    FB_LOGIC(ne < 0 || nn < 0 || (ne > 0 && !conn) || (nn > 0 && !coords), "feddb200_mesh_upload: bad sizes/pointers");
    FB_LOGIC(ne >= (int64_t(1) << 27), "feddb200_mesh_upload: more than 2^27 elements per GPU are not supported");
    for (int64_t k = 0; k < ne * nloc; k++)
        FB_LOGIC(conn[k] < 0 || conn[k] >= nn, "feddb200_mesh_upload: connectivity entry out of range");
    FB_CUDA(cudaSetDevice(c->device));
    auto *m = new feddb200_mesh();
    m->ctx = c; m->device = c->device; m->dim = dim; m->nloc = nloc; m->ne = ne; m->nn = nn;
    m->conn_h.assign(conn, conn + ne * nloc);
    FB_CUDA(cudaMalloc(&m->conn_d, std::max<size_t>(8, sizeof(int32_t) * ne * nloc)));
    FB_CUDA(cudaMalloc(&m->coords_d, std::max<size_t>(8, sizeof(double) * nn * dim)));
    FB_CUDA(cudaMemcpyAsync(m->conn_d, conn, sizeof(int32_t) * ne * nloc, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaMemcpyAsync(m->coords_d, coords, sizeof(double) * nn * dim, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    *out = m;
    return FEDDB200_OK;
}

extern "C" int feddb200_mesh_update_coords(feddb200_ctx *c, feddb200_mesh *m, const double *coords)
{
    FB_LOGIC(!c || !m || !coords, "feddb200_mesh_update_coords: null argument");
    FB_CUDA(cudaSetDevice(c->device));
    FB_CUDA(cudaMemcpyAsync(m->coords_d, coords, sizeof(double) * m->nn * m->dim, cudaMemcpyHostToDevice, c->stream));
    return FEDDB200_OK;
}

extern "C" void feddb200_mesh_free(feddb200_mesh *m)
{
    if (!m) return;
    cudaSetDevice(m->device);   // not m->ctx->device: bindings may free a mesh after feddb200_destroy
    cudaFree(m->conn_d);
    cudaFree(m->coords_d);
    delete m;
}

// ---------------------------------------------------------------------------------------
// pattern build kernels
// ---------------------------------------------------------------------------------------
static constexpr uint64_t kDropKey = ~uint64_t(0);

__global__ void k_make_keys(int64_t ne, int nr, int nc, const int32_t *__restrict__ conn_r,
                            const int32_t *__restrict__ conn_c, const int32_t *__restrict__ row_lid,
                            const int32_t *__restrict__ col_lid, uint64_t *__restrict__ keys)
{
    const int64_t n = ne * nr * nc;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nr * nc);
        const int ij = (int)(t - e * (nr * nc));
        const int i = ij / nc, j = ij - i * nc;
        int32_t r = conn_r[e * nr + i], c = conn_c[e * nc + j];
        if (row_lid) r = row_lid[r];
        if (col_lid) c = col_lid[c];
        keys[t] = (r < 0 || c < 0) ? kDropKey : ((uint64_t)(uint32_t)r << 32) | (uint32_t)c;
    }
}

__global__ void k_extra_keys(int64_t n, const int32_t *__restrict__ er, const int32_t *__restrict__ ec,
                             uint64_t *__restrict__ keys)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        keys[t] = ((uint64_t)(uint32_t)er[t] << 32) | (uint32_t)ec[t];
}

__global__ void k_split_unique(int64_t nnz, const uint64_t *__restrict__ ukeys, int32_t *__restrict__ colind)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x)
        colind[t] = (int32_t)(uint32_t)(ukeys[t] & 0xffffffffu);
}

// rowptr[r] = first unique key with row >= r  (lower bound)
__global__ void k_rowptr(int64_t n_rows, int64_t nnz, const uint64_t *__restrict__ ukeys, int64_t *__restrict__ rowptr)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t)r << 32;
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
        }
        rowptr[r] = lo;
    }
}

__global__ void k_make_pos(int64_t ne, int nr, int nc, int stride, const int32_t *__restrict__ conn_r,
                           const int32_t *__restrict__ conn_c, const int32_t *__restrict__ row_lid,
                           const int32_t *__restrict__ col_lid, const int64_t *__restrict__ rowptr,
                           const int32_t *__restrict__ colind, uint16_t *__restrict__ pos)
{
    const int64_t n = ne * nr * nc;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nr * nc);
        const int ij = (int)(t - e * (nr * nc));
        const int i = ij / nc, j = ij - i * nc;
        int32_t r = conn_r[e * nr + i], c = conn_c[e * nc + j];
        if (row_lid) r = row_lid[r];
        if (col_lid) c = col_lid[c];
        uint16_t p = 0xffff;
        if (r >= 0 && c >= 0) {
            const int64_t b = rowptr[r];
            int64_t lo = b, hi = rowptr[r + 1];
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (colind[mid] < c) lo = mid + 1; else hi = mid;
            }
            p = (uint16_t)(lo - b);
        }
        pos[(e * nr + i) * stride + j] = p;
    }
}

__global__ void k_make_inc(int64_t ne, int nr, const int32_t *__restrict__ conn_r, const int32_t *__restrict__ row_lid,
                           int32_t n_rows, int32_t *__restrict__ keys, int32_t *__restrict__ vals)
{
    const int64_t n = ne * nr;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / nr;
        const int i = (int)(t - e * nr);
        int32_t r = conn_r[t];
        if (row_lid) r = row_lid[r];
        keys[t] = r < 0 ? n_rows : r; // dropped rows sort to the end
        vals[t] = (int32_t)((e << 4) | i);
    }
}

__global__ void k_lower_bound32(int64_t n_rows, int64_t n, const int32_t *__restrict__ sorted, int64_t *__restrict__ ptr)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted[mid] < (int32_t)r) lo = mid + 1; else hi = mid;
        }
        ptr[r] = lo;
    }
}

static int grid_for(int64_t n, int threads = 256)
{
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 32));
}

static int bits_for(int64_t v)
{
    int b = 1;
    while ((int64_t(1) << b) <= v && b < 63) b++;
    return b;
}

extern "C" int feddb200_pattern_build(feddb200_ctx *c, feddb200_pat **out, const feddb200_mesh *rm,
                                      const feddb200_mesh *cm, int64_t n_rows, int64_t n_owned,
                                      const int32_t *row_lid, int64_t n_cols, const int32_t *col_lid,
                                      int64_t n_extra, const int32_t *extra_row, const int32_t *extra_col)
{
    FB_LOGIC(!c || !out || !rm || !cm, "feddb200_pattern_build: null argument");
    FB_LOGIC(rm->ne != cm->ne || rm->dim != cm->dim,
             "feddb200_pattern_build: row and column meshes must share the element list and dimension");
    FB_LOGIC(n_extra < 0 || (n_extra > 0 && (!extra_row || !extra_col)), "feddb200_pattern_build: bad extra entries");
    if (!row_lid) { n_rows = rm->nn; n_owned = rm->nn; }
    if (!col_lid) { n_cols = cm->nn; }
    FB_LOGIC(n_rows < 0 || n_owned < 0 || n_owned > n_rows || n_cols < 0, "feddb200_pattern_build: bad row/col counts");
    FB_LOGIC(n_rows >= (int64_t(1) << 31) - 1 || n_cols >= (int64_t(1) << 31) - 1,
             "feddb200_pattern_build: local index range exceeds int32 (LO)");
    if (row_lid)
        for (int64_t k = 0; k < rm->nn; k++) FB_LOGIC(row_lid[k] >= n_rows, "feddb200_pattern_build: row_lid out of range");
    if (col_lid)
        for (int64_t k = 0; k < cm->nn; k++)
            FB_LOGIC(col_lid[k] < 0 || col_lid[k] >= n_cols, "feddb200_pattern_build: col_lid out of range");
    for (int64_t k = 0; k < n_extra; k++)
        FB_LOGIC(extra_row[k] < 0 || extra_row[k] >= n_owned || extra_col[k] < 0 || extra_col[k] >= n_cols,
                 "feddb200_pattern_build: extra entry out of range (extra rows must be owned rows)");
    FB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int nr = rm->nloc, nc = cm->nloc;
    const int64_t ne = rm->ne;

    auto *p = new feddb200_pat();
    p->ctx = c; p->device = c->device; p->rm = rm; p->cm = cm;
    p->n_rows = n_rows; p->n_owned = n_owned; p->n_cols = n_cols;
    p->pos_stride = (nc + 1) & ~1;
    // every early return below (FB_CUDA / FB_LOGIC) frees the half-built pattern and the device temporaries
    struct Guard {
        feddb200_pat *pat;
        std::vector<void **> tmps;
        ~Guard() { for (void **t : tmps) if (*t) { cudaFree(*t); *t = nullptr; } if (pat) feddb200_pat_free(pat); }
    } guard{p, {}};
    auto tmp_of = [&guard](auto **pp) { guard.tmps.push_back(reinterpret_cast<void **>(pp)); };

    int32_t *row_lid_d = nullptr, *col_lid_d = nullptr, *er_d = nullptr, *ec_d = nullptr;
    tmp_of(&col_lid_d); tmp_of(&er_d); tmp_of(&ec_d);   // row_lid_d belongs to the pattern
    // temporaries of the sort / unique / incidence steps (function scope: the guard outlives the blocks that use them)
    void *sort_tmp = nullptr, *uniq_tmp = nullptr, *inc_tmp = nullptr;
    int64_t *count_d = nullptr;
    int32_t *k_a = nullptr, *k_b = nullptr, *v_a = nullptr, *v_b = nullptr;
    tmp_of(&sort_tmp); tmp_of(&uniq_tmp); tmp_of(&inc_tmp); tmp_of(&count_d); tmp_of(&k_a); tmp_of(&k_b); tmp_of(&v_a); tmp_of(&v_b);
    if (row_lid) {
        FB_CUDA(cudaMalloc(&row_lid_d, sizeof(int32_t) * std::max<int64_t>(rm->nn, 1)));
        FB_CUDA(cudaMemcpyAsync(row_lid_d, row_lid, sizeof(int32_t) * rm->nn, cudaMemcpyHostToDevice, st));
    }
    if (col_lid) {
        FB_CUDA(cudaMalloc(&col_lid_d, sizeof(int32_t) * std::max<int64_t>(cm->nn, 1)));
        FB_CUDA(cudaMemcpyAsync(col_lid_d, col_lid, sizeof(int32_t) * cm->nn, cudaMemcpyHostToDevice, st));
    }
    p->row_lid_d = row_lid_d;

    // 1. keys
    const int64_t nkeys = ne * nr * nc + n_extra;
    uint64_t *keys_a = nullptr, *keys_b = nullptr;
    tmp_of(&keys_a); tmp_of(&keys_b);
    FB_CUDA(cudaMalloc(&keys_a, sizeof(uint64_t) * std::max<int64_t>(nkeys, 1)));
    FB_CUDA(cudaMalloc(&keys_b, sizeof(uint64_t) * std::max<int64_t>(nkeys, 1)));
    if (ne > 0) {
        k_make_keys<<<grid_for(ne * nr * nc), 256, 0, st>>>(ne, nr, nc, rm->conn_d, cm->conn_d, row_lid_d, col_lid_d, keys_a);
        c->launches++;
    }
    if (n_extra > 0) {
        FB_CUDA(cudaMalloc(&er_d, sizeof(int32_t) * n_extra));
        FB_CUDA(cudaMalloc(&ec_d, sizeof(int32_t) * n_extra));
        FB_CUDA(cudaMemcpyAsync(er_d, extra_row, sizeof(int32_t) * n_extra, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(ec_d, extra_col, sizeof(int32_t) * n_extra, cudaMemcpyHostToDevice, st));
        k_extra_keys<<<grid_for(n_extra), 256, 0, st>>>(n_extra, er_d, ec_d, keys_a + ne * nr * nc);
        c->launches++;
    }
    // 2. sort (dropped keys = all ones sort last)
    {
        cub::DoubleBuffer<uint64_t> db(keys_a, keys_b);
        size_t tmp_bytes = 0;
        FB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, nkeys, 0, 64, st));
        FB_CUDA(cudaMalloc(&sort_tmp, std::max<size_t>(tmp_bytes, 8)));
        FB_CUDA(cub::DeviceRadixSort::SortKeys(sort_tmp, tmp_bytes, db, nkeys, 0, 64, st));
        FB_CUDA(cudaStreamSynchronize(st));
        cudaFree(sort_tmp); sort_tmp = nullptr;
        if (db.Current() != keys_a) std::swap(keys_a, keys_b);
    }
    // 3. unique
    int64_t nuniq = 0;
    {
        FB_CUDA(cudaMalloc(&count_d, sizeof(int64_t)));
        size_t tmp_bytes = 0;
        FB_CUDA(cub::DeviceSelect::Unique(nullptr, tmp_bytes, keys_a, keys_b, count_d, nkeys, st));
        FB_CUDA(cudaMalloc(&uniq_tmp, std::max<size_t>(tmp_bytes, 8)));
        FB_CUDA(cub::DeviceSelect::Unique(uniq_tmp, tmp_bytes, keys_a, keys_b, count_d, nkeys, st));
        FB_CUDA(cudaMemcpyAsync(&nuniq, count_d, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        FB_CUDA(cudaStreamSynchronize(st));
        cudaFree(uniq_tmp); uniq_tmp = nullptr;
        cudaFree(count_d); count_d = nullptr;
        if (nuniq > 0) { // the drop key, if present, is the last unique key
            uint64_t last = 0;
            FB_CUDA(cudaMemcpy(&last, keys_b + (nuniq - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
            if (last == kDropKey) nuniq--;
        }
    }
    p->nnz = nuniq;
    cudaFree(keys_a); keys_a = nullptr;
    // 4. colind, rowptr
    FB_CUDA(cudaMalloc(&p->colind_d, sizeof(int32_t) * std::max<int64_t>(nuniq, 1)));
    FB_CUDA(cudaMalloc(&p->rowptr_d, sizeof(int64_t) * (n_rows + 1)));
    if (nuniq > 0) { k_split_unique<<<grid_for(nuniq), 256, 0, st>>>(nuniq, keys_b, p->colind_d); c->launches++; }
    k_rowptr<<<grid_for(n_rows + 1), 256, 0, st>>>(n_rows, nuniq, keys_b, p->rowptr_d);
    c->launches++;
    p->rowptr_h.resize(n_rows + 1);
    FB_CUDA(cudaMemcpyAsync(p->rowptr_h.data(), p->rowptr_d, sizeof(int64_t) * (n_rows + 1), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    cudaFree(keys_b); keys_b = nullptr;
    p->nnz_owned = p->rowptr_h[n_owned];
    int max_len = 0;
    for (int64_t r = 0; r < n_rows; r++) max_len = std::max<int>(max_len, (int)(p->rowptr_h[r + 1] - p->rowptr_h[r]));
    p->max_len = max_len;
    if (max_len >= 0xffff) {
        fb::set_error("feddb200_pattern_build: a node row has >= 65535 entries (position map is 16 bit)");
        return FEDDB200_ELOGIC;   // the guard frees the pattern
    }
    // 5. position map
    FB_CUDA(cudaMalloc(&p->pos_d, sizeof(uint16_t) * std::max<int64_t>(ne * nr * p->pos_stride, 1)));
    if (ne > 0) {
        FB_CUDA(cudaMemsetAsync(p->pos_d, 0xff, sizeof(uint16_t) * ne * nr * p->pos_stride, st));
        k_make_pos<<<grid_for(ne * nr * nc), 256, 0, st>>>(ne, nr, nc, p->pos_stride, rm->conn_d, cm->conn_d, row_lid_d,
                                                          col_lid_d, p->rowptr_d, p->colind_d, p->pos_d);
        c->launches++;
    }
    // 6. incidences (row -> (element, local index)), stable sort keeps elements ascending
    {
        const int64_t n = ne * nr;
        FB_CUDA(cudaMalloc(&k_a, sizeof(int32_t) * std::max<int64_t>(n, 1)));
        FB_CUDA(cudaMalloc(&k_b, sizeof(int32_t) * std::max<int64_t>(n, 1)));
        FB_CUDA(cudaMalloc(&v_a, sizeof(int32_t) * std::max<int64_t>(n, 1)));
        FB_CUDA(cudaMalloc(&v_b, sizeof(int32_t) * std::max<int64_t>(n, 1)));
        if (n > 0) {
            k_make_inc<<<grid_for(n), 256, 0, st>>>(ne, nr, rm->conn_d, row_lid_d, (int32_t)n_rows, k_a, v_a);
            c->launches++;
        }
        cub::DoubleBuffer<int32_t> dk(k_a, k_b), dv(v_a, v_b);
        size_t tmp_bytes = 0;
        FB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, bits_for(n_rows), st));
        FB_CUDA(cudaMalloc(&inc_tmp, std::max<size_t>(tmp_bytes, 8)));
        FB_CUDA(cub::DeviceRadixSort::SortPairs(inc_tmp, tmp_bytes, dk, dv, n, 0, bits_for(n_rows), st));
        FB_CUDA(cudaMalloc(&p->inc_ptr_d, sizeof(int64_t) * (n_rows + 1)));
        k_lower_bound32<<<grid_for(n_rows + 1), 256, 0, st>>>(n_rows, n, dk.Current(), p->inc_ptr_d);
        c->launches++;
        int64_t n_inc = 0;
        FB_CUDA(cudaMemcpyAsync(&n_inc, p->inc_ptr_d + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        FB_CUDA(cudaStreamSynchronize(st));
        p->n_inc = n_inc;
        FB_CUDA(cudaMalloc(&p->inc_d, sizeof(int32_t) * std::max<int64_t>(n_inc, 1)));
        FB_CUDA(cudaMemcpyAsync(p->inc_d, dv.Current(), sizeof(int32_t) * n_inc, cudaMemcpyDeviceToDevice, st));
        FB_CUDA(cudaStreamSynchronize(st));
    }
    guard.pat = nullptr;   // success: the caller owns the pattern; the guard frees what is left of the temporaries
    *out = p;
    return FEDDB200_OK;
}

extern "C" void feddb200_pat_free(feddb200_pat *p)
{
    if (!p) return;
    cudaSetDevice(p->device);   // not p->ctx->device: bindings may free a pattern after feddb200_destroy
    cudaFree(p->rowptr_d); cudaFree(p->colind_d); cudaFree(p->row_lid_d); cudaFree(p->pos_d);
    cudaFree(p->inc_ptr_d); cudaFree(p->inc_d); cudaFree(p->row_perm_d); cudaFree(p->colour_perm_d);
    cudaFree(p->rec_d); cudaFree(p->geom_d); cudaFree(p->frag_d); cudaFree(p->vtx_d); cudaFree(p->coords4_d); cudaFree(p->rowinfo_d); cudaFree(p->ahead_d); cudaFree(p->task_tiles_d); cudaFree(p->tasks_d); cudaFree(p->tiletet_d); cudaFree(p->fanrec_d); cudaFree(p->star_tiles_d[0]); cudaFree(p->star_tiles_d[1]); cudaFree(p->sloc_d); cudaFree(p->sloc_tab_d); cudaFree(p->dt_d);
    delete p;
}

extern "C" int feddb200_pattern_info(const feddb200_pat *p, int64_t *n_rows, int64_t *n_owned, int64_t *n_cols,
                                     int64_t *nnz, int64_t *nnz_owned, int32_t *max_len, int32_t *n_colours)
{
    FB_LOGIC(!p, "null pattern");
    if (n_rows) *n_rows = p->n_rows;
    if (n_owned) *n_owned = p->n_owned;
    if (n_cols) *n_cols = p->n_cols;
    if (nnz) *nnz = p->nnz;
    if (nnz_owned) *nnz_owned = p->nnz_owned;
    if (max_len) *max_len = p->max_len;
    if (n_colours) {
        if (p->n_colours == 0 && fb::ensure_colouring(const_cast<feddb200_pat *>(p)) != 0) return FEDDB200_ERUNTIME;
        *n_colours = p->n_colours;
    }
    return FEDDB200_OK;
}

extern "C" int feddb200_pattern_get_nodes(feddb200_ctx *c, const feddb200_pat *p, int64_t *rowptr, int32_t *colind)
{
    FB_LOGIC(!c || !p, "null argument");
    FB_CUDA(cudaSetDevice(c->device));
    if (rowptr) std::memcpy(rowptr, p->rowptr_h.data(), sizeof(int64_t) * (p->n_rows + 1));
    if (colind && p->nnz > 0) FB_CUDA(cudaMemcpy(colind, p->colind_d, sizeof(int32_t) * p->nnz, cudaMemcpyDeviceToHost));
    return FEDDB200_OK;
}

static int64_t layout_factor(int rd, int cd, int mode)
{
    if (mode == FEDDB200_BLOCK_SCALAR) return (rd == 1 && cd == 1) ? 1 : -1;
    if (mode == FEDDB200_BLOCK_DIAG) return (rd == cd && rd >= 1 && rd <= 3) ? rd : -1;
    if (mode == FEDDB200_BLOCK_FULL) return (rd >= 1 && rd <= 3 && cd >= 1 && cd <= 3) ? (int64_t)rd * cd : -1;
    return -1;
}

extern "C" int64_t feddb200_pattern_nnz(const feddb200_pat *p, int rd, int cd, int mode)
{
    if (!p) return -1;
    const int64_t f = layout_factor(rd, cd, mode);
    return f < 0 ? -1 : f * p->nnz;
}
extern "C" int64_t feddb200_pattern_nnz_owned(const feddb200_pat *p, int rd, int cd, int mode)
{
    if (!p) return -1;
    const int64_t f = layout_factor(rd, cd, mode);
    return f < 0 ? -1 : f * p->nnz_owned;
}

// dof-level CSR: row rd*I+a holds, for every node entry p of row I, the columns cd*J+b
// (full: b = 0..cd-1; diag: b = a only)
__global__ void k_expand(int64_t n_rows, int rd, int cd, int diag, const int64_t *__restrict__ rowptr,
                         const int32_t *__restrict__ colind, int64_t *__restrict__ rowptr_out,
                         int32_t *__restrict__ colind_out)
{
    const int per = diag ? 1 : cd;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_rows * rd; t += nthreads) {
        const int64_t I = t / rd;
        const int a = (int)(t - I * rd);
        const int64_t b0 = rowptr[I], L = rowptr[I + 1] - b0;
        const int64_t start = (int64_t)rd * per * b0 + (int64_t)a * per * L;
        rowptr_out[t] = start;
        for (int64_t q = 0; q < L; q++) {
            const int32_t J = colind[b0 + q];
            if (diag) colind_out[start + q] = cd * J + a;
            else for (int b = 0; b < cd; b++) colind_out[start + q * cd + b] = cd * J + b;
        }
        if (t == n_rows * rd - 1) rowptr_out[n_rows * rd] = (int64_t)rd * per * rowptr[n_rows];
    }
}

extern "C" int feddb200_pattern_expand(feddb200_ctx *c, const feddb200_pat *p, int rd, int cd, int mode,
                                       int64_t *rowptr, int32_t *colind)
{
    FB_LOGIC(!c || !p, "null argument");
    const int64_t f = layout_factor(rd, cd, mode);
    FB_LOGIC(f < 0, "feddb200_pattern_expand: unsupported dof layout");
    FB_LOGIC((int64_t)cd * p->n_cols >= (int64_t(1) << 31), "feddb200_pattern_expand: dof column index exceeds int32");
    FB_CUDA(cudaSetDevice(c->device));
    const int64_t nr = p->n_rows * rd, nnz = f * p->nnz;
    int64_t *rp_d = nullptr;
    int32_t *ci_d = nullptr;
    FB_CUDA(cudaMalloc(&rp_d, sizeof(int64_t) * (nr + 1)));
    FB_CUDA(cudaMalloc(&ci_d, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    if (nr > 0) {
        k_expand<<<grid_for(nr), 256, 0, c->stream>>>(p->n_rows, rd, cd, mode == FEDDB200_BLOCK_DIAG, p->rowptr_d,
                                                      p->colind_d, rp_d, ci_d);
        c->launches++;
    } else {
        FB_CUDA(cudaMemsetAsync(rp_d, 0, sizeof(int64_t), c->stream));
    }
    if (rowptr) FB_CUDA(cudaMemcpyAsync(rowptr, rp_d, sizeof(int64_t) * (nr + 1), cudaMemcpyDeviceToHost, c->stream));
    if (colind && nnz > 0) FB_CUDA(cudaMemcpyAsync(colind, ci_d, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(rp_d);
    cudaFree(ci_d);
    return FEDDB200_OK;
}

// ---------------------------------------------------------------------------------------
// greedy element colouring (host, lazy): two elements conflict when they share a node of the
// denser of the two meshes, which covers every shared (row node, col node) value slot.
// ---------------------------------------------------------------------------------------
namespace fb {
int ensure_colouring(feddb200_pat *p)
{
    if (p->n_colours > 0 || p->rm->ne == 0) return 0;
    const feddb200_mesh *m = p->rm->nloc >= p->cm->nloc ? p->rm : p->cm;
    const int64_t ne = m->ne;
    const int nl = m->nloc;
    constexpr int W = 4; // up to 256 colours
    std::vector<uint64_t> mask((size_t)m->nn * W, 0);
    std::vector<int32_t> colour(ne);
    int ncol = 0;
    for (int64_t e = 0; e < ne; e++) {
        uint64_t forb[W] = {0, 0, 0, 0};
        for (int i = 0; i < nl; i++) {
            const uint64_t *mk = &mask[(size_t)m->conn_h[e * nl + i] * W];
            for (int w = 0; w < W; w++) forb[w] |= mk[w];
        }
        int col = -1;
        for (int w = 0; w < W && col < 0; w++)
            if (~forb[w]) col = w * 64 + __builtin_ctzll(~forb[w]);
        if (col < 0) { set_error("element colouring needs more than 256 colours"); return FEDDB200_ERUNTIME; }
        colour[e] = col;
        ncol = std::max(ncol, col + 1);
        for (int i = 0; i < nl; i++) mask[(size_t)m->conn_h[e * nl + i] * W + col / 64] |= uint64_t(1) << (col % 64);
    }
    p->colour_ptr.assign(ncol + 1, 0);
    for (int64_t e = 0; e < ne; e++) p->colour_ptr[colour[e] + 1]++;
    for (int k = 0; k < ncol; k++) p->colour_ptr[k + 1] += p->colour_ptr[k];
    std::vector<int32_t> perm(ne);
    std::vector<int64_t> fill(p->colour_ptr.begin(), p->colour_ptr.end() - 1);
    for (int64_t e = 0; e < ne; e++) perm[fill[colour[e]]++] = (int32_t)e;
    cudaSetDevice(p->ctx->device);
    if (cudaMalloc(&p->colour_perm_d, sizeof(int32_t) * ne) != cudaSuccess ||
        cudaMemcpy(p->colour_perm_d, perm.data(), sizeof(int32_t) * ne, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("element colouring: device allocation/copy failed");
        return FEDDB200_ERUNTIME;
    }
    p->n_colours = ncol;
    return 0;
}
} // namespace fb
