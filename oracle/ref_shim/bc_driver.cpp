// oracle/_ref BC driver (test infrastructure): compiles the REFERENCE's own BCBuilder members -- sliced by extract_bc.py from
// /root/reference/feddlib/core/General/BCBuilder_def.hpp at build time, never copied into the repo -- against the reference's own
// BCBuilder_decl.hpp and mock containers (include_bc/bc_mocks.hpp), and exposes setSystem / setRHS through a C interface.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iterator>

#include "feddlib/core/General/BCBuilder_decl.hpp"
#include "bc_subset.inc"

using namespace FEDD;
typedef long long GOx;
typedef BCBuilder<double, int, GOx, default_no> BC_t;
typedef Domain<double, int, GOx, default_no> Domain_t;
typedef Map<int, GOx, default_no> Map_t;
typedef Matrix<double, int, GOx, default_no> Matrix_t;
typedef BlockMatrix<double, int, GOx, default_no> BlockMatrix_t;
typedef MultiVector<double, int, GOx, default_no> MV_t;
typedef BlockMultiVector<double, int, GOx, default_no> BlockMV_t;

static thread_local std::string g_bc_err;
extern "C" const char *ref_bc_last_error(void) { return g_bc_err.c_str(); }

typedef void (*bc_func_c)(double *x, double *res, double t, const double *parameters);

static Teuchos::RCP<Domain_t> make_domain(int dim, int64_t n_nodes, const int *flags, const double *points, const int64_t *node_gid)
{
    Teuchos::RCP<Domain_t> d(new Domain_t());
    d->dim_ = dim;
    d->flags_ = Teuchos::rcp(new std::vector<int>(flags, flags + n_nodes));
    d->points_ = Teuchos::rcp(new std::vector<std::vector<double> >(n_nodes, std::vector<double>(dim, 0.0)));
    if (points)
        for (int64_t k = 0; k < n_nodes; k++)
            for (int c = 0; c < dim; c++) (*d->points_)[k][c] = points[k * dim + c];
    std::vector<GOx> g(node_gid, node_gid + n_nodes);
    d->mapUnique_ = Teuchos::RCP<const Map_t>(new Map_t(g.data(), g.size()));
    return d;
}

static void add_bcs(BC_t &bc, const Teuchos::RCP<Domain_t> &dom, int n_bc, const int *flag, const int *block, const char *const *type,
                    const int *dofs, bc_func_c func, const double *params, int n_params)
{
    std::vector<double> pv(params, params + n_params);
    for (int k = 0; k < n_bc; k++) {
        BC_t::BC_func_Type f = func ? BC_t::BC_func_Type(func) : BC_t::BC_func_Type();
        if (n_params > 0) bc.addBC(f, flag[k], block[k], dom, type[k], dofs[k], pv);
        else bc.addBC(f, flag[k], block[k], dom, type[k], dofs[k]);
    }
}

// BCBuilder::setSystem on an nb x nb block system over ONE node set (every block's rows = dofs_of_block[i] * node + d):
// block (i,j) is CSR k = i * nb + j: rowptr[k] (NULL: block absent), colind[k], values[k] (in/out), col_gid[k] / ncols[k].
extern "C" int ref_bc_set_system(int dim, int64_t n_nodes, const int *node_flags, const int64_t *node_gid,
                                 int n_bc, const int *bc_flag, const int *bc_block, const char *const *bc_type, const int *bc_dofs,
                                 int nb, const int *dofs_of_block, const long long *const *rowptr, const int *const *colind,
                                 double *const *values, const int64_t *const *col_gid, const int64_t *ncols)
{
    try {
        Teuchos::RCP<Domain_t> dom = make_domain(dim, n_nodes, node_flags, nullptr, node_gid);
        BC_t bc;
        add_bcs(bc, dom, n_bc, bc_flag, bc_block, bc_type, bc_dofs, nullptr, nullptr, 0);
        Teuchos::RCP<BlockMatrix_t> S(new BlockMatrix_t(nb));
        for (int i = 0; i < nb; i++) {
            std::vector<GOx> rg((size_t)n_nodes * dofs_of_block[i]);
            for (int64_t n = 0; n < n_nodes; n++)
                for (int d = 0; d < dofs_of_block[i]; d++) rg[(size_t)n * dofs_of_block[i] + d] = (GOx)dofs_of_block[i] * node_gid[n] + d;
            Teuchos::RCP<const Map_t> rowMap(new Map_t(rg.data(), rg.size()));
            for (int j = 0; j < nb; j++) {
                const int k = i * nb + j;
                if (!rowptr[k]) continue;
                std::vector<GOx> cg(col_gid[k], col_gid[k] + ncols[k]);
                Teuchos::RCP<const Map_t> colMap(new Map_t(cg.data(), cg.size()));
                S->addBlock(Teuchos::rcp(new Matrix_t(rowptr[k], colind[k], values[k], rowMap, colMap)), i, j);
            }
        }
        bc.setSystem(S);
        for (int k = 0; k < nb * nb; k++)
            if (rowptr[k] && S->blocks_[k]->resumes_ != S->blocks_[k]->completes_) throw std::runtime_error("resumeFill without fillComplete");
        return 0;
    } catch (const std::exception &e) { g_bc_err = e.what(); return 1; }
}

// BCBuilder::setRHS on nb block vectors over one node set: rhs[i] has dofs_of_block[i] * n_nodes entries (in/out)
extern "C" int ref_bc_set_rhs(int dim, int64_t n_nodes, const int *node_flags, const double *points, const int64_t *node_gid,
                              int n_bc, const int *bc_flag, const int *bc_block, const char *const *bc_type, const int *bc_dofs,
                              bc_func_c func, const double *params, int n_params, int nb, const int *dofs_of_block, double *const *rhs, double t)
{
    try {
        Teuchos::RCP<Domain_t> dom = make_domain(dim, n_nodes, node_flags, points, node_gid);
        BC_t bc;
        add_bcs(bc, dom, n_bc, bc_flag, bc_block, bc_type, bc_dofs, func, params, n_params);
        Teuchos::RCP<BlockMV_t> V(new BlockMV_t());
        for (int i = 0; i < nb; i++) V->blocks_.push_back(Teuchos::rcp(new MV_t(rhs[i], (size_t)n_nodes * dofs_of_block[i])));
        bc.setRHS(V, t);
        return 0;
    } catch (const std::exception &e) { g_bc_err = e.what(); return 1; }
}
