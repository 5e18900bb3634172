// oracle/_ref BlockMatrix driver (test infrastructure): the REFERENCE's own BlockMatrix::merge (determineGlobalOffsets,
// mergeBlockNew) and BlockMap::merge -- sliced by extract_bm.py at build time, never copied into the repo -- against the reference's
// own BlockMatrix_decl.hpp / BlockMap_decl.hpp and mock containers (include_bm/bm_mocks.hpp), behind a C interface.
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "bm_mocks.hpp"
#include "feddlib/core/LinearAlgebra/BlockMap_decl.hpp"
#include "feddlib/core/LinearAlgebra/BlockMatrix_decl.hpp"
#include "bm_subset.inc"

using namespace FEDD;
typedef long long GOx;
typedef Map<int, GOx, default_no> Map_t;
typedef Matrix<double, int, GOx, default_no> Matrix_t;
typedef BlockMatrix<double, int, GOx, default_no> BlockMatrix_t;

static thread_local std::string g_bm_err;
extern "C" const char *ref_bm_last_error(void) { return g_bm_err.c_str(); }

struct Merged {
    std::vector<long long> rowptr, rowgid, colgid;
    std::vector<double> values;
};

// nb x nb blocks; block k = i * nb + j: rowptr[k] (NULL: absent), colind[k], values[k]; row gids of block row i: row_gid[i]
// (n_rows[i]), column gids of block column j as seen by block (i, j): col_gid[k] (n_cols[k]).  Returns a handle.
extern "C" void *ref_bm_merge(int nb, const int64_t *n_rows, const long long *const *row_gid, const long long *const *rowptr,
                              const int *const *colind, const double *const *values, const int64_t *n_cols, const long long *const *col_gid)
{
    try {
        Teuchos::RCP<BlockMatrix_t> S(new BlockMatrix_t((unsigned)nb));
        for (int i = 0; i < nb; i++) {
            Teuchos::RCP<const Map_t> rowMap(new Map_t(row_gid[i], (size_t)n_rows[i]));
            for (int j = 0; j < nb; j++) {
                const int k = i * nb + j;
                if (!rowptr[k]) continue;
                Teuchos::RCP<const Map_t> colMap(new Map_t(col_gid[k], (size_t)n_cols[k]));
                S->addBlock(Teuchos::rcp(new Matrix_t(rowptr[k], colind[k], values[k], rowMap, colMap)), i, j);
            }
        }
        S->merge();
        Merged *M = new Merged;
        const Matrix_t &A = *S->mergedMatrix_;
        M->rowptr.push_back(0);
        for (size_t r = 0; r < A.ins_.size(); r++) {
            M->rowgid.push_back(A.row_->gids_[r]);
            for (const auto &e : A.ins_[r]) { M->colgid.push_back(e.first); M->values.push_back(e.second); }
            M->rowptr.push_back((long long)M->colgid.size());
        }
        return M;
    } catch (const std::exception &e) { g_bm_err = e.what(); return nullptr; }
}
extern "C" int64_t ref_bm_rows(const void *h) { return (int64_t)static_cast<const Merged *>(h)->rowgid.size(); }
extern "C" int64_t ref_bm_nnz(const void *h) { return (int64_t)static_cast<const Merged *>(h)->colgid.size(); }
extern "C" void ref_bm_get(const void *h, long long *rowptr, long long *rowgid, long long *colgid, double *values)
{
    const Merged &M = *static_cast<const Merged *>(h);
    std::memcpy(rowptr, M.rowptr.data(), M.rowptr.size() * 8);
    std::memcpy(rowgid, M.rowgid.data(), M.rowgid.size() * 8);
    std::memcpy(colgid, M.colgid.data(), M.colgid.size() * 8);
    std::memcpy(values, M.values.data(), M.values.size() * 8);
}
extern "C" void ref_bm_free(void *h) { delete static_cast<Merged *>(h); }
