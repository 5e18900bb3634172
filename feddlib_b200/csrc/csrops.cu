// CSR algebra on resident matrices: the two passes the reference makes over every assembled
// matrix before the solve, both of which re-insert every entry through insertGlobalValues on the CPU:
//   Matrix::addMatrix        -> Xpetra TwoMatrixAdd  B := alpha*A + beta*B   (core/LinearAlgebra/Matrix_def.hpp:281-287;
//                               NavierStokes_def.hpp:303-304, 312-313)
//   BlockMatrix::merge       -> monolithic [A B^T; B C] with block offsets   (core/LinearAlgebra/BlockMatrix_def.hpp:119-289)
// Here they are two-phase (symbolic: row lengths -> exclusive scan; numeric: fill) kernels over device CSR arrays with
// ascending column indices per row (what feddb200_pattern_expand emits).  One warp per row; columns of the result stay
// ascending; explicit zeros of the inputs are kept, like Tpetra keeps every inserted entry.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace fb {
namespace {

__device__ __forceinline__ int lower_bound_i32(const int32_t *a, int n, int32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Union of two ascending rows, one warp per row.  An entry of A lands at  i + lb_B(a_i) - (#matches before i),
// an unmatched entry of B at  j + lb_A(b_j) - (#matches before j);  matched entries of B fold into A's slot.
// NUMERIC = false: only the row length is produced (len[r] = nA + nB - #matches).
template <bool NUMERIC>
__global__ void __launch_bounds__(256) k_csr_add(int64_t n_rows, double alpha, const int64_t *__restrict__ rpA,
                                                 const int32_t *__restrict__ ciA, const double *__restrict__ vA, double beta,
                                                 const int64_t *__restrict__ rpB, const int32_t *__restrict__ ciB,
                                                 const double *__restrict__ vB, const int64_t *__restrict__ rpC,
                                                 int64_t *__restrict__ len, int32_t *__restrict__ ciC, double *__restrict__ vC)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const int64_t a0 = rpA[r], b0 = rpB[r];
        const int nA = (int)(rpA[r + 1] - a0), nB = (int)(rpB[r + 1] - b0);
        const int32_t *A = ciA + a0, *B = ciB + b0;
        int matches = 0;
        for (int base = 0; base < nA; base += 32) {           // entries of A
            const int i = base + lane;
            int lb = 0, hit = 0;
            if (i < nA) {
                const int32_t key = A[i];
                lb = lower_bound_i32(B, nB, key);
                hit = lb < nB && B[lb] == key;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if constexpr (NUMERIC) {
                if (i < nA) {
                    const int64_t p = rpC[r] + i + lb - (matches + __popc(m & ((1u << lane) - 1u)));
                    ciC[p] = A[i];
                    // beta*B scaled first, alpha*A summed into it (TwoMatrixAdd): two rounded products and one sum, no FMA
                    const double pa = __dmul_rn(alpha, vA[a0 + i]);
                    vC[p] = hit ? __dadd_rn(__dmul_rn(beta, vB[b0 + lb]), pa) : pa;
                }
            }
            matches += __popc(m);
        }
        if constexpr (!NUMERIC) {
            if (lane == 0) len[r] = nA + nB - matches;
        } else {
            int seen = 0;
            for (int base = 0; base < nB; base += 32) {       // unmatched entries of B
                const int j = base + lane;
                int lb = 0, hit = 0;
                if (j < nB) {
                    const int32_t key = B[j];
                    lb = lower_bound_i32(A, nA, key);
                    hit = lb < nA && A[lb] == key;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (j < nB && !hit) {
                    const int64_t p = rpC[r] + j + lb - (seen + __popc(m & ((1u << lane) - 1u)));
                    ciC[p] = B[j];
                    vC[p] = beta * vB[b0 + j];
                }
                seen += __popc(m);
            }
        }
    }
}

constexpr int kMaxBlocks = 4;   // block rows / columns of a BlockMatrix (the reference's problems use 2 or 3)
struct MergeArgs {
    int nb;
    int64_t n_rows[kMaxBlocks], row_off[kMaxBlocks + 1];
    int32_t col_off[kMaxBlocks];
    const int64_t *rp[kMaxBlocks][kMaxBlocks];
    const int32_t *ci[kMaxBlocks][kMaxBlocks];
    const double *v[kMaxBlocks][kMaxBlocks];
};

// merged row = the rows of the blocks (i, 0), (i, 1), ... back to back, columns shifted by the block-column offset
// (BlockMatrix_def.hpp:252-270).  One warp per merged row.
template <bool NUMERIC>
__global__ void __launch_bounds__(256) k_block_merge(const MergeArgs M, const int64_t *__restrict__ rpM, int64_t *__restrict__ len,
                                                     int32_t *__restrict__ ciM, double *__restrict__ vM)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = M.row_off[M.nb];
    for (int64_t R = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; R < total; R += warps) {
        int i = 0;
        while (i + 1 < M.nb && R >= M.row_off[i + 1]) i++;
        const int64_t r = R - M.row_off[i];
        int64_t out = NUMERIC ? rpM[R] : 0;
        for (int j = 0; j < M.nb; j++) {
            const int64_t *rp = M.rp[i][j];
            if (!rp) continue;
            const int64_t k0 = rp[r];
            const int n = (int)(rp[r + 1] - k0);
            if constexpr (NUMERIC) {
                const int32_t off = M.col_off[j];
                for (int x = lane; x < n; x += 32) {
                    ciM[out + x] = M.ci[i][j][k0 + x] + off;
                    vM[out + x] = M.v[i][j][k0 + x];
                }
            }
            out += n;
        }
        if constexpr (!NUMERIC) {
            if (lane == 0) len[R] = out;
        }
    }
}

int scan_lengths(feddb200_ctx *c, int64_t n, int64_t *rp_d /* [n+1]: lengths in rp_d[1..n] */, int64_t *nnz)
{
    // inclusive scan of rp[1..n] in place; rp[0] = 0
    FB_CUDA(cudaMemsetAsync(rp_d, 0, sizeof(int64_t), c->stream));
    if (n > 0) {
        size_t tmp_bytes = 0;
        FB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, rp_d + 1, rp_d + 1, n, c->stream));
        void *tmp = nullptr;
        FB_CUDA(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
        cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, rp_d + 1, rp_d + 1, n, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nnz, rp_d + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        FB_CUDA(e);
    } else *nnz = 0;
    return FEDDB200_OK;
}

unsigned warp_grid(int64_t rows) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, 148 * 32)); }

} // namespace
} // namespace fb

using namespace fb;

extern "C" int feddb200_csr_add_symbolic_d(feddb200_ctx *c, int64_t n_rows, const int64_t *rpA_d, const int32_t *ciA_d,
                                           const int64_t *rpB_d, const int32_t *ciB_d, int64_t *rpC_d, int64_t *nnzC)
{
    FB_LOGIC(!c || n_rows < 0 || !rpA_d || !rpB_d || !rpC_d || !nnzC, "csr_add_symbolic: bad arguments");
    FB_CUDA(cudaSetDevice(c->device));
    if (n_rows > 0) {
        k_csr_add<false><<<warp_grid(n_rows), 256, 0, c->stream>>>(n_rows, 0.0, rpA_d, ciA_d, nullptr, 0.0, rpB_d, ciB_d, nullptr,
                                                                  nullptr, rpC_d + 1, nullptr, nullptr);
        c->launches++;
        FB_CUDA(cudaGetLastError());
    }
    return scan_lengths(c, n_rows, rpC_d, nnzC);
}

extern "C" int feddb200_csr_add_numeric_d(feddb200_ctx *c, int64_t n_rows, double alpha, const int64_t *rpA_d, const int32_t *ciA_d,
                                          const double *vA_d, double beta, const int64_t *rpB_d, const int32_t *ciB_d,
                                          const double *vB_d, const int64_t *rpC_d, int32_t *ciC_d, double *vC_d)
{
    FB_LOGIC(!c || n_rows < 0 || !rpA_d || !rpB_d || !rpC_d, "csr_add_numeric: bad arguments");
    if (n_rows == 0) return FEDDB200_OK;
    FB_LOGIC(!ciC_d || !vC_d || !vA_d || !vB_d, "csr_add_numeric: null value or index array");
    FB_CUDA(cudaSetDevice(c->device));
    k_csr_add<true><<<warp_grid(n_rows), 256, 0, c->stream>>>(n_rows, alpha, rpA_d, ciA_d, vA_d, beta, rpB_d, ciB_d, vB_d, rpC_d,
                                                             nullptr, ciC_d, vC_d);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}

static int fill_merge_args(MergeArgs &M, int nb, const int64_t *n_rows, const int32_t *n_cols, const int64_t *const *rp,
                           const int32_t *const *ci, const double *const *v, bool numeric)
{
    FB_LOGIC(nb < 1 || nb > kMaxBlocks || !n_rows || !n_cols || !rp, "block_merge: bad arguments (1 to 4 block rows)");
    std::memset(&M, 0, sizeof(M));
    M.nb = nb;
    int64_t ro = 0, co = 0;
    for (int i = 0; i < nb; i++) {
        FB_LOGIC(n_rows[i] < 0 || n_cols[i] < 0, "block_merge: negative size");
        M.n_rows[i] = n_rows[i]; M.row_off[i] = ro; M.col_off[i] = (int32_t)co;
        ro += n_rows[i]; co += n_cols[i];
        FB_LOGIC(co > 0x7fffffff, "block_merge: merged column index exceeds int32");
        for (int j = 0; j < nb; j++) {
            M.rp[i][j] = rp[i * nb + j];
            if (numeric && M.rp[i][j]) {
                FB_LOGIC(!ci || !v || !ci[i * nb + j] || !v[i * nb + j], "block_merge: block without column indices or values");
                M.ci[i][j] = ci[i * nb + j]; M.v[i][j] = v[i * nb + j];
            }
        }
    }
    M.row_off[nb] = ro;
    return FEDDB200_OK;
}

extern "C" int feddb200_block_merge_symbolic_d(feddb200_ctx *c, int nb, const int64_t *n_rows, const int32_t *n_cols,
                                               const int64_t *const *rowptr_d, int64_t *rpM_d, int64_t *nnzM)
{
    FB_LOGIC(!c || !rpM_d || !nnzM, "block_merge_symbolic: bad arguments");
    MergeArgs M;
    int rc = fill_merge_args(M, nb, n_rows, n_cols, rowptr_d, nullptr, nullptr, false);
    if (rc != FEDDB200_OK) return rc;
    FB_CUDA(cudaSetDevice(c->device));
    const int64_t total = M.row_off[nb];
    if (total > 0) {
        k_block_merge<false><<<warp_grid(total), 256, 0, c->stream>>>(M, nullptr, rpM_d + 1, nullptr, nullptr);
        c->launches++;
        FB_CUDA(cudaGetLastError());
    }
    return scan_lengths(c, total, rpM_d, nnzM);
}

extern "C" int feddb200_block_merge_numeric_d(feddb200_ctx *c, int nb, const int64_t *n_rows, const int32_t *n_cols,
                                              const int64_t *const *rowptr_d, const int32_t *const *colind_d,
                                              const double *const *values_d, const int64_t *rpM_d, int32_t *ciM_d, double *vM_d)
{
    FB_LOGIC(!c || !rpM_d, "block_merge_numeric: bad arguments");
    MergeArgs M;
    int rc = fill_merge_args(M, nb, n_rows, n_cols, rowptr_d, colind_d, values_d, true);
    if (rc != FEDDB200_OK) return rc;
    const int64_t total = M.row_off[nb];
    if (total == 0) return FEDDB200_OK;
    FB_LOGIC(!ciM_d || !vM_d, "block_merge_numeric: null output");
    FB_CUDA(cudaSetDevice(c->device));
    k_block_merge<true><<<warp_grid(total), 256, 0, c->stream>>>(M, rpM_d, nullptr, ciM_d, vM_d);
    c->launches++;
    FB_CUDA(cudaGetLastError());
    return FEDDB200_OK;
}
