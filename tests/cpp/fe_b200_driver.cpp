// fe_b200_driver -- C++ side-by-side check of the drop-in boundary, shaped like the reference's own test driver
// feddlib/core/FE/tests/fe.cpp:60-101 (build a Domain, FE::addFE, allocate A on the unique map, call assemblyXxx):
// the SAME Domain object goes through
//   (1) FEDDLib's own FE<SC,LO,GO,NO> routines (oracle/_ref/libfedd_ref.so: FE_def.hpp compiled unmodified against
//       the mock Trilinos containers of oracle/ref_shim -- the checker), and
//   (2) FEDD::FE_b200 (feddlib_b200/csrc/host/FE_b200.hpp -> libfeddb200.so -> CUDA kernels),
// and the two CSR matrices are compared: pattern exactly, values to 1e-12 relative Frobenius error.
// Test infrastructure (may use oracle/); built by __graft_entry__.build(), run by tests/test_gpu_cpp_host.py.
//
//   fe_b200_driver <mesh.bin>      mesh.bin: int64 header {dim, nloc1, ne, nn1, nloc2, nn2}, conn1 int32, coords f64,
//                                  gid1 int64, conn2 int32, gid2 int64, u f64[dim*nn1]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "fedd_mocks.hpp"

#include "FE_b200.hpp"

extern "C" {
fo_matrix *fo_matrix_new(int64_t nrows, int32_t cap_hint);
void fo_matrix_free(fo_matrix *A);
int64_t fo_matrix_nnz(const fo_matrix *A);
void fo_get_csr(const fo_matrix *A, int64_t *rowptr, int64_t *colgid, double *vals);
int ref_assemble(int op, int dim, const char *fe1, const char *fe2, int64_t ne, const int32_t *conn1, int nloc1,
                 const double *coords1, int64_t nn1, const int64_t *gid1, const int32_t *conn2, int nloc2, int64_t nn2,
                 const int64_t *gid2, const double *u, double lambda, double mu, fo_matrix *A, fo_matrix *B);
int ref_assemble_rhs(int dim, const char *fe1, int64_t ne, const int32_t *conn1, int nloc1, const double *coords1, int64_t nn1,
                     int vec_field, int deg_func, const double *value_func, double *rhs);
int ref_assemble_stress(int dim, const char *fe1, int64_t ne, const int32_t *conn1, int nloc1, const double *coords1, int64_t nn1,
                        const int64_t *gid1, double (*func)(const double *, void *), void *user, fo_matrix *A);
const char *ref_last_error(void);
}

// coefficient functions of the assemblyStress cases: user = dimension
static double stress_one(const double *, void *) { return 1.0; }
static double stress_varying(const double *x, void *user)
{
    return 1.0 + 0.5 * x[0] + 0.25 * x[1] * x[1] + (*static_cast<int *>(user) > 2 ? 0.125 * x[2] : 0.0);
}

using namespace FEDD;
typedef long long GOx;
typedef int NOx;
typedef Domain<double, int, GOx, NOx> Domain_t;
typedef Map<int, GOx, NOx> Map_t;
typedef Matrix<double, int, GOx, NOx> Matrix_t;
typedef MultiVector<double, int, GOx, NOx> MV_t;
typedef b200::LocalCsr<double, int, GOx> Csr_t;

// seat step for the mock Matrix: keep the CSR beside the matrix object
static std::map<const void *, Csr_t> g_seated;
namespace FEDD { namespace b200 {
void seat_csr(Teuchos::RCP<Matrix_t> &A, Csr_t &csr, Teuchos::RCP<const Map_t>, Teuchos::RCP<const Map_t>, bool callFillComplete)
{
    g_seated[A.get()] = csr;
    if (callFillComplete) A->fillCompleteCalls_++;
}
}} // namespace FEDD::b200

static Teuchos::RCP<Domain_t> make_domain(int dim, const char *fe, int64_t ne, const int32_t *conn, int nloc, const double *coords,
                                          int64_t nn, const int64_t *gid)
{
    Teuchos::RCP<Domain_t> d(new Domain_t(dim, fe));
    d->elementsC_ = Teuchos::rcp(new Elements());
    for (int64_t e = 0; e < ne; e++) d->elementsC_->addElement(FiniteElement(std::vector<int>(conn + e * nloc, conn + (e + 1) * nloc)));
    d->pointsRep_ = Teuchos::rcp(new std::vector<std::vector<double> >(nn, std::vector<double>(dim, 0.0)));
    if (coords)
        for (int64_t k = 0; k < nn; k++)
            for (int c = 0; c < dim; c++) (*d->pointsRep_)[k][c] = coords[k * dim + c];
    std::vector<GOx> g(gid, gid + nn);
    d->mapRepeated_ = Teuchos::RCP<const Map_t>(new Map_t(g.data(), g.size()));
    return d;
}

struct RefCsr { std::vector<int64_t> rowptr, col; std::vector<double> val; };
static RefCsr ref_csr(fo_matrix *A, int64_t nrows)
{
    RefCsr r;
    r.rowptr.resize(nrows + 1); r.col.resize(fo_matrix_nnz(A)); r.val.resize(r.col.size());
    fo_get_csr(A, r.rowptr.data(), r.col.data(), r.val.data());
    return r;
}

// rows of `ours` are local rows (row-map order); global dof row = rowDofs * gidRow[local node] + d
static bool compare(const char *name, const Csr_t &ours, const RefCsr &ref, const std::vector<int64_t> &gidRow, int rowDofs)
{
    double num = 0.0, den = 0.0;
    const std::vector<std::int64_t> &rowptr = ours.pattern->rowptr;
    const std::vector<std::int32_t> &colind = ours.pattern->colind;
    const std::vector<GOx> &colmap = ours.pattern->colmap;
    const double *values = ours.values.get();
    const int64_t nLocalRows = (int64_t)rowptr.size() - 1;
    bool pattern_ok = nLocalRows == (int64_t)gidRow.size() * rowDofs;
    for (int64_t lr = 0; lr < nLocalRows && pattern_ok; lr++) {
        const int64_t gr = rowDofs * gidRow[lr / rowDofs] + lr % rowDofs;
        const int64_t a0 = rowptr[lr], a1 = rowptr[lr + 1], b0 = ref.rowptr[gr], b1 = ref.rowptr[gr + 1];
        if (a1 - a0 != b1 - b0) { pattern_ok = false; break; }
        std::vector<std::pair<GOx, double> > row;
        for (int64_t k = a0; k < a1; k++) row.push_back(std::make_pair(colmap[colind[k]], values[k]));
        std::sort(row.begin(), row.end());
        for (int64_t k = 0; k < a1 - a0; k++) {
            if (row[k].first != ref.col[b0 + k]) { pattern_ok = false; break; }
            const double d = row[k].second - ref.val[b0 + k];
            num += d * d; den += ref.val[b0 + k] * ref.val[b0 + k];
        }
    }
    const double err = std::sqrt(num) / (den > 0 ? std::sqrt(den) : 1.0);
    const bool ok = pattern_ok && err <= 1e-12;
    std::printf("%-28s nnz %10lld  pattern %s  rel.Frobenius %.3e  %s\n", name, (long long)ours.nnz,
                pattern_ok ? "exact" : "MISMATCH", err, ok ? "PASS" : "FAIL");
    return ok;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s mesh.bin\n", argv[0]); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    int64_t h[6];
    if (std::fread(h, sizeof(int64_t), 6, f) != 6) return 2;
    const int dim = (int)h[0], nloc1 = (int)h[1], nloc2 = (int)h[4];
    const int64_t ne = h[2], nn1 = h[3], nn2 = h[5];
    std::vector<int32_t> conn1(ne * nloc1), conn2(ne * nloc2);
    std::vector<double> xyz(nn1 * dim), u(dim * nn1);
    std::vector<int64_t> gid1(nn1), gid2(nn2);
    bool rd = std::fread(conn1.data(), 4, conn1.size(), f) == conn1.size() && std::fread(xyz.data(), 8, xyz.size(), f) == xyz.size() &&
              std::fread(gid1.data(), 8, gid1.size(), f) == gid1.size() && std::fread(conn2.data(), 4, conn2.size(), f) == conn2.size() &&
              std::fread(gid2.data(), 8, gid2.size(), f) == gid2.size() && std::fread(u.data(), 8, u.size(), f) == u.size();
    std::fclose(f);
    if (!rd) { std::fprintf(stderr, "short mesh file\n"); return 2; }
    const char *fe1 = (nloc1 == dim + 1) ? "P1" : "P2", *fe2 = (nloc2 == dim + 1) ? "P1" : "P2";
    const double lambda = 8.0e6, mu = 2.0e6;
    const int64_t nglob1 = *std::max_element(gid1.begin(), gid1.end()) + 1, nglob2 = *std::max_element(gid2.begin(), gid2.end()) + 1;

    bool all = true;
    try {
        // --- our side: one FE_b200 with both spaces registered, as the Stokes / Navier-Stokes problems do
        FE_b200<double, int, GOx, NOx> fe;
        Teuchos::RCP<Domain_t> dV = make_domain(dim, fe1, ne, conn1.data(), nloc1, xyz.data(), nn1, gid1.data());
        Teuchos::RCP<Domain_t> dP = make_domain(dim, fe2, ne, conn2.data(), nloc2, xyz.data(), nn2, gid2.data());
        fe.addFE(dV);
        if (std::string(fe1) != fe2) fe.addFE(dP);
        Teuchos::RCP<MV_t> uMV(new MV_t(u.data(), u.size()));

        struct Case { const char *name; int op; int rowDofs; };
        const Case cases[] = {{"assemblyLaplace", 0, 1}, {"assemblyLaplaceVecField", 1, dim}, {"assemblyLinElasXDim", 2, dim},
                              {"assemblyAdvectionVecField", 3, dim}, {"assemblyAdvectionInUVecField", 4, dim},
                              {"assemblyMass Scalar", 7, 1}, {"assemblyMass Vector", 8, dim}};
        std::vector<Case> todo(std::begin(cases), std::end(cases));
        if (std::string(fe1) == "P1") todo.push_back({"assemblyBDStabilization", 9, 1});
        todo.push_back({"assemblyStress (func = 1)", 10, dim});
        todo.push_back({"assemblyStress (func varies)", 11, dim});
        for (const Case &c : todo) {
            // reference
            fo_matrix *rA = fo_matrix_new(c.rowDofs * nglob1, 64), *rB = fo_matrix_new(1, 8);
            int dimv = dim;
            const int rrc = c.op >= 10
                ? ref_assemble_stress(dim, fe1, ne, conn1.data(), nloc1, xyz.data(), nn1, gid1.data(), c.op == 10 ? stress_one : stress_varying, &dimv, rA)
                : ref_assemble(c.op, dim, fe1, fe1, ne, conn1.data(), nloc1, xyz.data(), nn1, gid1.data(), nullptr, 0, 0, nullptr, u.data(),
                               lambda, mu, rA, rB);
            if (rrc != 0) { std::printf("reference failed: %s\n", ref_last_error()); return 1; }
            RefCsr ref = ref_csr(rA, c.rowDofs * nglob1);
            fo_matrix_free(rA); fo_matrix_free(rB);
            // ours
            Teuchos::RCP<Matrix_t> A(new Matrix_t(nullptr));
            switch (c.op) {
            case 0: fe.assemblyLaplace(dim, fe1, 2, A); break;
            case 1: fe.assemblyLaplaceVecField(dim, fe1, 2, A); break;
            case 2: fe.assemblyLinElasXDim(dim, fe1, A, lambda, mu); break;
            case 3: fe.assemblyAdvectionVecField(dim, fe1, A, uMV, true); break;
            case 4: fe.assemblyAdvectionInUVecField(dim, fe1, A, uMV, true); break;
            case 7: fe.assemblyMass(dim, fe1, "Scalar", A); break;
            case 8: fe.assemblyMass(dim, fe1, "Vector", A); break;
            case 9: fe.assemblyBDStabilization(dim, fe1, A); break;
            case 10: fe.assemblyStress(dim, fe1, A, [](double *x, int *) { return stress_one(x, nullptr); }, nullptr); break;
            case 11: fe.assemblyStress(dim, fe1, A, [&dimv](double *x, int *) { return stress_varying(x, &dimv); }, nullptr); break;
            }
            all = compare(c.name, g_seated[A.get()], ref, gid1, c.rowDofs) && all;
            all = (A->fillCompleteCalls_ == 1) && all;
        }
        if (std::string(fe1) != fe2) { // B and B^T (velocity fe1, pressure fe2)
            fo_matrix *rB = fo_matrix_new(nglob2, 64), *rBT = fo_matrix_new(dim * nglob1, 64);
            if (ref_assemble(5, dim, fe1, fe2, ne, conn1.data(), nloc1, xyz.data(), nn1, gid1.data(), conn2.data(), nloc2, nn2, gid2.data(),
                             nullptr, 0, 0, rB, rBT) != 0) { std::printf("reference failed: %s\n", ref_last_error()); return 1; }
            RefCsr refB = ref_csr(rB, nglob2), refBT = ref_csr(rBT, dim * nglob1);
            fo_matrix_free(rB); fo_matrix_free(rBT);
            Teuchos::RCP<Matrix_t> B(new Matrix_t(nullptr)), BT(new Matrix_t(nullptr));
            fe.assemblyDivAndDivT(dim, fe1, fe2, 2, B, BT, Teuchos::RCP<const Map_t>(), Teuchos::RCP<const Map_t>(), true);
            all = compare("assemblyDivAndDivT: B", g_seated[B.get()], refB, gid2, 1) && all;
            all = compare("assemblyDivAndDivT: BT", g_seated[BT.get()], refBT, gid1, dim) && all;
        }
        if (dim == 2) { // B and B^T with P0 pressure: rows / columns on the element map (FE_def.hpp:1954-1957, 2012-2013, 2039-2040)
            std::vector<int32_t> conn0((std::size_t)ne);
            std::vector<int64_t> egid((std::size_t)ne);
            for (int64_t e = 0; e < ne; e++) { conn0[(std::size_t)e] = (int32_t)e; egid[(std::size_t)e] = (e * 7919) % ne; }   // a permuted element map
            if (ne % 7919 == 0) for (int64_t e = 0; e < ne; e++) egid[(std::size_t)e] = e;
            Teuchos::RCP<Domain_t> d0 = make_domain(dim, "P0", ne, conn0.data(), 1, nullptr, ne, egid.data());
            d0->elementMap_ = d0->mapRepeated_;
            fe.addFE(d0);
            fo_matrix *rB = fo_matrix_new(ne, 64), *rBT = fo_matrix_new(dim * nglob1, 64);
            if (ref_assemble(5, dim, fe1, "P0", ne, conn1.data(), nloc1, xyz.data(), nn1, gid1.data(), conn0.data(), 1, ne, egid.data(),
                             nullptr, 0, 0, rB, rBT) != 0) { std::printf("reference failed: %s\n", ref_last_error()); return 1; }
            RefCsr refB = ref_csr(rB, ne), refBT = ref_csr(rBT, dim * nglob1);
            fo_matrix_free(rB); fo_matrix_free(rBT);
            Teuchos::RCP<Matrix_t> B(new Matrix_t(nullptr)), BT(new Matrix_t(nullptr));
            fe.assemblyDivAndDivT(dim, fe1, "P0", 2, B, BT, Teuchos::RCP<const Map_t>(), Teuchos::RCP<const Map_t>(), true);
            all = compare("assemblyDivAndDivT (P0 pressure): B", g_seated[B.get()], refB, egid, 1) && all;
            all = compare("assemblyDivAndDivT (P0 pressure): BT", g_seated[BT.get()], refBT, gid1, dim) && all;
        }
        // assemblyRHS (constant source), "Scalar" and "Vector"
        for (int vec = 0; vec < 2; vec++) {
            const double f[3] = {1.5, -2.0, 0.25};
            const std::size_t n = (std::size_t)nn1 * (vec ? dim : 1);
            std::vector<double> want(n, 0.0), zero(n, 0.0);
            if (ref_assemble_rhs(dim, fe1, ne, conn1.data(), nloc1, xyz.data(), nn1, vec, 1, f, want.data()) != 0) { std::printf("reference failed: %s\n", ref_last_error()); return 1; }
            Teuchos::RCP<MV_t> a(new MV_t(zero.data(), n));
            std::vector<double> para(2, 0.0); para[1] = 1.0;
            auto func = [&](double *, double *res, double *) { for (int d = 0; d < dim; d++) res[d] = f[d]; };
            fe.assemblyRHS(dim, fe1, a, vec ? "Vector" : "Scalar", func, para);
            Teuchos::ArrayRCP<const double> got = a->getData(0);
            double num = 0.0, den = 0.0;
            for (std::size_t k = 0; k < n; k++) { num += (got[k] - want[k]) * (got[k] - want[k]); den += want[k] * want[k]; }
            const double err = std::sqrt(num / den);
            std::printf("%-28s n   %10lld                 rel. 2-norm   %.3e  %s\n", vec ? "assemblyRHS Vector" : "assemblyRHS Scalar", (long long)n, err, err <= 1e-12 ? "PASS" : "FAIL");
            all = (err <= 1e-12) && all;
        }
        // error behaviour of the boundary (FE_def.hpp:610, 6950)
        bool threw = false;
        try { Teuchos::RCP<Matrix_t> A(new Matrix_t(nullptr)); fe.assemblyLaplace(dim, "P0", 2, A); } catch (const std::logic_error &) { threw = true; }
        all = threw && all;
        threw = false;
        try { Teuchos::RCP<Matrix_t> A(new Matrix_t(nullptr)); fe.assemblyLaplace(dim == 2 ? 3 : 2, fe1, 2, A); } catch (const std::logic_error &) { threw = true; }
        all = threw && all;
        std::printf("error behaviour (P0 / missing addFE -> std::logic_error)  %s\n", threw ? "PASS" : "FAIL");
        std::printf("kernel launches: %lld\n", (long long)fe.launchCount());
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 1;
    }
    std::printf(all ? "ALL PASS\n" : "SOME FAILED\n");
    return all ? 0 : 1;
}
