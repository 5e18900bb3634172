// oracle/_ref driver (test infrastructure): compiles the REFERENCE's own hot-path routines -- sliced by
// extract.py from /root/reference/feddlib/core/FE/FE_def.hpp at build time, never copied into the repo --
// against the reference's own FE_decl.hpp / SmallMatrix.hpp / FEDDCore.hpp and mock Trilinos containers, and
// exposes them through the same C interface as the oracle restatement (fedd_oracle.c).
#include <cstring>

#include "feddlib/core/FE/FE_decl.hpp"
#include "fe_subset.inc"

using namespace FEDD;
typedef long long GOx;
typedef FE<double, int, GOx, default_no> FE_t;
typedef Domain<double, int, GOx, default_no> Domain_t;
typedef Map<int, GOx, default_no> Map_t;
typedef Matrix<double, int, GOx, default_no> Matrix_t;
typedef MultiVector<double, int, GOx, default_no> MV_t;

static thread_local std::string g_err;

static Teuchos::RCP<Domain_t> make_domain(int dim, const char *fe, int64_t ne, const int32_t *conn, int nloc,
                                          const double *coords, int64_t nn, const int64_t *gid)
{
    Teuchos::RCP<Domain_t> d(new Domain_t(dim, fe));
    d->elementsC_ = Teuchos::rcp(new Elements());
    for (int64_t e = 0; e < ne; e++)
        d->elementsC_->addElement(FiniteElement(std::vector<int>(conn + e * nloc, conn + (e + 1) * nloc)));
    d->pointsRep_ = Teuchos::rcp(new std::vector<std::vector<double> >(nn, std::vector<double>(dim)));
    if (coords)
        for (int64_t k = 0; k < nn; k++)
            for (int c = 0; c < dim; c++) (*d->pointsRep_)[k][c] = coords[k * dim + c];
    std::vector<GOx> g(gid, gid + nn);
    d->mapRepeated_ = Teuchos::RCP<const Map_t>(new Map_t(g.data(), g.size()));
    // a "P0" domain is described by one pseudo-node per element: its ids are the element map (rows of B, FE_def.hpp:1954, 2012)
    if (std::string(fe) == "P0") d->elementMap_ = d->mapRepeated_;
    return d;
}

extern "C" const char *ref_last_error(void) { return g_err.c_str(); }

// op: 0 assemblyLaplace, 1 assemblyLaplaceVecField, 2 assemblyLinElasXDim, 3 assemblyAdvectionVecField,
//     4 assemblyAdvectionInUVecField, 5 assemblyDivAndDivT, 6 assemblyDivAndDivTFast,
//     7 assemblyMass "Scalar", 8 assemblyMass "Vector", 9 assemblyBDStabilization
extern "C" int ref_assemble(int op, int dim, const char *fe1, const char *fe2, int64_t ne,
                            const int32_t *conn1, int nloc1, const double *coords1, int64_t nn1, const int64_t *gid1,
                            const int32_t *conn2, int nloc2, int64_t nn2, const int64_t *gid2,
                            const double *u, double lambda, double mu, fo_matrix *A, fo_matrix *B)
{
    try {
        FE_t fe;
        Teuchos::RCP<Domain_t> d1 = make_domain(dim, fe1, ne, conn1, nloc1, coords1, nn1, gid1);
        fe.addFE(d1);
        Teuchos::RCP<Matrix_t> mA(new Matrix_t(A));
        std::string t1(fe1);
        switch (op) {
        case 0: fe.assemblyLaplace(dim, t1, 2, mA, true); break;
        case 1: fe.assemblyLaplaceVecField(dim, t1, 2, mA, true); break;
        case 2: fe.assemblyLinElasXDim(dim, t1, mA, lambda, mu, true); break;
        case 7: fe.assemblyMass(dim, t1, std::string("Scalar"), mA, true); break;
        case 8: fe.assemblyMass(dim, t1, std::string("Vector"), mA, true); break;
        case 9: fe.assemblyBDStabilization(dim, t1, mA, true); break;
        case 3:
        case 4: {
            Teuchos::RCP<MV_t> mv(new MV_t(u, (std::size_t)dim * nn1));
            if (op == 3) fe.assemblyAdvectionVecField(dim, t1, mA, mv, true);
            else fe.assemblyAdvectionInUVecField(dim, t1, mA, mv, true);
            break;
        }
        case 5:
        case 6: {
            std::string t2(fe2);
            Teuchos::RCP<Domain_t> d2 = make_domain(dim, fe2, ne, conn2, nloc2, nullptr, nn2, gid2);
            if (t2 != t1) fe.addFE(d2);
            Teuchos::RCP<Matrix_t> mB(new Matrix_t(B));
            Matrix_t::MapConstPtr_Type m1, m2;
            if (op == 5) fe.assemblyDivAndDivT(dim, t1, t2, 2, mA, mB, m1, m2, true);
            else fe.assemblyDivAndDivTFast(dim, t1, t2, 2, mA, mB, m1, m2, true);
            break;
        }
        default: g_err = "unknown op"; return -1;
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -2;
    }
}

// FE::assemblyStress with the coefficient callback func(xyz, user) (CoeffFunc_Type = boost::function<double(double*, int*)>)
extern "C" int ref_assemble_stress(int dim, const char *fe1, int64_t ne, const int32_t *conn1, int nloc1, const double *coords1,
                                   int64_t nn1, const int64_t *gid1, double (*func)(const double *, void *), void *user, fo_matrix *A)
{
    try {
        FE_t fe;
        Teuchos::RCP<Domain_t> d1 = make_domain(dim, fe1, ne, conn1, nloc1, coords1, nn1, gid1);
        fe.addFE(d1);
        Teuchos::RCP<Matrix_t> mA(new Matrix_t(A));
        CoeffFunc_Type f = [func, user](double *x, int *) { return func(x, user); };
        int *dummy = nullptr;
        fe.assemblyStress(dim, std::string(fe1), mA, f, dummy, true);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -2;
    }
}

// FE::assemblyRHS with a constant source: rhs[nn * dofs] (repeated vector, zero on entry)
extern "C" int ref_assemble_rhs(int dim, const char *fe1, int64_t ne, const int32_t *conn1, int nloc1, const double *coords1,
                                int64_t nn1, int vec_field, int deg_func, const double *value_func, double *rhs)
{
    try {
        FE_t fe;
        std::vector<int64_t> gid(nn1);
        for (int64_t k = 0; k < nn1; k++) gid[k] = k;
        Teuchos::RCP<Domain_t> d1 = make_domain(dim, fe1, ne, conn1, nloc1, coords1, nn1, gid.data());
        fe.addFE(d1);
        const std::size_t n = (std::size_t)nn1 * (vec_field ? dim : 1);
        std::vector<double> zero(n, 0.0);
        Teuchos::RCP<MV_t> mv(new MV_t(zero.data(), n));
        std::vector<double> f(value_func, value_func + dim);
        RhsFunc_Type func = [f, dim](double *, double *res, double *) { for (int d = 0; d < dim; d++) res[d] = f[d]; };
        std::vector<double> para(2, 0.0);
        para[1] = (double)deg_func; // "last parameter should always be the degree" (FE_def.hpp:4715)
        fe.assemblyRHS(dim, std::string(fe1), mv, std::string(vec_field ? "Vector" : "Scalar"), func, para);
        Teuchos::ArrayRCP<const double> out = mv->getData(0);
        for (std::size_t k = 0; k < n; k++) rhs[k] = out[k];
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -2;
    }
}

// reference tables, for direct comparison with the restatement
extern "C" int ref_get_dphi(int dim, const char *fe, int deg, double *dphi, double *w)
{
    try {
        FE_t f;
        vec3D_dbl_ptr_Type D;
        vec_dbl_ptr_Type W = Teuchos::rcp(new vec_dbl_Type(0));
        f.getDPhi(D, W, dim, std::string(fe), deg);
        std::size_t k = 0;
        for (auto &q : *D) for (auto &i : q) for (double v : i) dphi[k++] = v;
        for (std::size_t q = 0; q < W->size(); q++) w[q] = (*W)[q];
        return (int)W->size();
    } catch (const std::exception &e) { g_err = e.what(); return -2; }
}
extern "C" int ref_get_phi(int dim, const char *fe, int deg, double *phi, double *w)
{
    try {
        FE_t f;
        vec2D_dbl_ptr_Type P;
        vec_dbl_ptr_Type W = Teuchos::rcp(new vec_dbl_Type(0));
        f.getPhi(P, W, dim, std::string(fe), deg);
        std::size_t k = 0;
        for (auto &q : *P) for (double v : q) phi[k++] = v;
        for (std::size_t q = 0; q < W->size(); q++) w[q] = (*W)[q];
        return (int)W->size();
    } catch (const std::exception &e) { g_err = e.what(); return -2; }
}
extern "C" int ref_determine_degree(int dim, const char *fe1, const char *fe2, int t1, int t2, int extra)
{
    FE_t f;
    return (int)f.determineDegree((UN)dim, std::string(fe1), std::string(fe2), (FE_t::VarType)t1, (FE_t::VarType)t2, (UN)extra);
}
