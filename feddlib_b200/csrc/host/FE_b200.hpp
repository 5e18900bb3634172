// FE_b200.hpp -- C++ host side of the drop-in: FEDD::FE_b200<SC,LO,GO,NO> keeps the member signatures of
// FEDD::FE<SC,LO,GO,NO> for the element-wise matrix assembly hot path (feddlib/core/FE/FE_decl.hpp:116,
// 130-135, 153-157, 194-224, 252-257) and forwards the work to libfeddb200.so through the C ABI of
// include/feddb200.h.  What it reads of the FEDDLib containers is exactly what the reference routines read
// (Domain::getDimension/getFEType/getElementsC/getPointsRepeated/getMapRepeated, Elements::numberElements/
// getElement, FiniteElement::getVectorNodeList, Map::getNodeNumElements/getGlobalElement,
// MultiVector::getData), so it compiles unchanged against
//   * FEDDLib + Trilinos (define FEDD_B200_TRILINOS) -- the returned matrix is a
//     fill-complete Tpetra::CrsMatrix on the caller's row map, built from the engine's CSR arrays, and
//   * the mock containers of oracle/ref_shim (tests/cpp/fe_b200_driver.cpp), where the reference's own
//     FE_def.hpp routines run side by side on the same Domain objects.
// Include the FEDDLib headers (or the mocks) BEFORE this header; it includes nothing of FEDDLib itself.
//
// Error behaviour mirrors the reference: std::logic_error for unsupported FE types / missing addFE
// (FE_def.hpp:610, 676, 2746, 6950), std::runtime_error for engine failures.  There is no CPU fallback.
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <tuple>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "feddb200.h"
#include "feddb200_halo.h"

namespace FEDD {
namespace b200 {

inline void check(int rc)
{
    if (rc == FEDDB200_OK) return;
    const std::string msg = feddb200_last_error();
    if (rc == FEDDB200_ELOGIC) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

// Structure of one assembled matrix: local rows in row-map order, local column indices into `colmap` (global dof ids of
// the column map, owned first).  Built ONCE per (pattern, dof layout) and shared by every matrix assembled on it.
template <class GO>
struct CsrStructure {
    std::vector<std::int64_t> rowptr;
    std::vector<std::int32_t> colind;
    std::vector<GO> colmap;
};

// Page-locked value buffers (feddb200_host_alloc) are expensive to create, so they are pooled: a buffer belongs to the
// matrix it was handed to (shared ownership) and returns to the pool when the last owner lets go of it.  In a Newton or
// time loop the previous Jacobian is released before or right after the next assembly, so the pool settles at one or
// two buffers and every device-to-host copy of the values runs at the PCIe rate.
class PinnedPool : public std::enable_shared_from_this<PinnedPool> {
  public:
    explicit PinnedPool(feddb200_ctx *ctx) : ctx_(ctx) {}
    ~PinnedPool() { retire(); }
    std::shared_ptr<double> take(std::size_t n)
    {
        std::size_t best = free_.size();
        for (std::size_t k = 0; k < free_.size(); k++)
            if (free_[k].second >= n && (best == free_.size() || free_[k].second < free_[best].second)) best = k;
        double *p = nullptr;
        std::size_t cap = n;
        if (best < free_.size()) {
            p = free_[best].first; cap = free_[best].second;
            free_.erase(free_.begin() + (std::ptrdiff_t)best);
        } else {
            void *raw = nullptr;
            check_(feddb200_host_alloc(ctx_, &raw, (std::int64_t)(sizeof(double) * std::max<std::size_t>(n, 1))));
            p = static_cast<double *>(raw);
        }
        std::weak_ptr<PinnedPool> home = shared_from_this();
        return std::shared_ptr<double>(p, [home, cap](double *q) {
            if (std::shared_ptr<PinnedPool> h = home.lock()) { if (h->ctx_) { h->free_.push_back(std::make_pair(q, cap)); return; } }
            feddb200_host_free(nullptr, q);
        });
    }
    // called before the engine context goes away: buffers still held by matrices free themselves later
    void retire()
    {
        for (auto &f : free_) feddb200_host_free(nullptr, f.first);
        free_.clear();
        ctx_ = nullptr;
    }
    std::size_t idle() const { return free_.size(); }

  private:
    static void check_(int rc)
    {
        if (rc != FEDDB200_OK) throw std::runtime_error(feddb200_last_error());
    }
    feddb200_ctx *ctx_;
    std::vector<std::pair<double *, std::size_t> > free_;
};

// CSR of one assembled matrix, as handed to the seat step: the shared structure + this matrix's values on the host
template <class SC, class LO, class GO>
struct LocalCsr {
    std::shared_ptr<const CsrStructure<GO> > pattern;
    std::shared_ptr<double> values;   // page-locked, pooled
    std::size_t nnz = 0;
};

#ifdef FEDD_B200_TRILINOS
// Teuchos deallocator that only holds a shared owner of the memory an ArrayRCP views
template <class T, class Owner>
struct KeepAlive {
    typedef T ptr_t;
    explicit KeepAlive(const std::shared_ptr<Owner> &o) : owner(o) {}
    void free(T *) { owner.reset(); }
    std::shared_ptr<Owner> owner;
};

// Seat step for the real containers: wrap the CSR in a Tpetra::CrsMatrix on the caller's row map and re-seat
// `A` (passed as MatrixPtr_Type&, as in the reference).  domain/range default to the row map, like
// Matrix::fillComplete() (Matrix_def.hpp:192-194); B / B^T pass (map1,map2) / (map2,map1) (FE_def.hpp:2052-2055).
template <class SC, class LO, class GO, class NO>
void seat_csr(Teuchos::RCP<Matrix<SC, LO, GO, NO> > &A, LocalCsr<SC, LO, GO> &csr,
              Teuchos::RCP<const Map<LO, GO, NO> > domainMap, Teuchos::RCP<const Map<LO, GO, NO> > rangeMap,
              bool callFillComplete)
{
    typedef Tpetra::Map<LO, GO, NO> TMap;
    typedef Tpetra::CrsMatrix<SC, LO, GO, NO> TCrs;
    Teuchos::RCP<const Map<LO, GO, NO> > rowMapF = A->getMap();
    Teuchos::RCP<const TMap> rowMap = Xpetra::toTpetra(rowMapF->getXpetraMap());
    Teuchos::RCP<const TMap> colMap = Teuchos::rcp(new TMap(Teuchos::OrdinalTraits<Tpetra::global_size_t>::invalid(),
                                                              Teuchos::ArrayView<const GO>(csr.pattern->colmap.data(), csr.pattern->colmap.size()),
                                                              rowMap->getIndexBase(), rowMap->getComm()));
    // no copies: Tpetra's arrays are views of the shared structure and of this matrix's pooled value buffer; each view
    // keeps its owner alive (Teuchos deallocator) and lets go of it when Tpetra releases the array
    static_assert(sizeof(std::size_t) == sizeof(std::int64_t) && sizeof(LO) == sizeof(std::int32_t), "index types");
    const CsrStructure<GO> &P = *csr.pattern;
    Teuchos::ArrayRCP<std::size_t> rp = Teuchos::arcp(reinterpret_cast<std::size_t *>(const_cast<std::int64_t *>(P.rowptr.data())), 0,
                                                     (Teuchos::Ordinal)P.rowptr.size(), KeepAlive<std::size_t, const CsrStructure<GO> >(csr.pattern), true);
    Teuchos::ArrayRCP<LO> ci = Teuchos::arcp(reinterpret_cast<LO *>(const_cast<std::int32_t *>(P.colind.data())), 0,
                                             (Teuchos::Ordinal)P.colind.size(), KeepAlive<LO, const CsrStructure<GO> >(csr.pattern), true);
    Teuchos::ArrayRCP<SC> va = Teuchos::arcp(reinterpret_cast<SC *>(csr.values.get()), 0, (Teuchos::Ordinal)csr.nnz,
                                             KeepAlive<SC, double>(csr.values), true);
    Teuchos::RCP<TCrs> T = Teuchos::rcp(new TCrs(rowMap, colMap, rp, ci, va));
    if (callFillComplete) {
        Teuchos::RCP<const TMap> dom = domainMap.is_null() ? rowMap : Xpetra::toTpetra(domainMap->getXpetraMap());
        Teuchos::RCP<const TMap> ran = rangeMap.is_null() ? rowMap : Xpetra::toTpetra(rangeMap->getXpetraMap());
        T->expertStaticFillComplete(dom, ran);
    }
    Teuchos::RCP<Xpetra::CrsMatrix<SC, LO, GO, NO> > xCrs = Teuchos::rcp(new Xpetra::TpetraCrsMatrix<SC, LO, GO, NO>(T));
    Teuchos::RCP<Xpetra::Matrix<SC, LO, GO, NO> > xMat = Teuchos::rcp(new Xpetra::CrsMatrixWrap<SC, LO, GO, NO>(xCrs));
    A = Teuchos::rcp(new Matrix<SC, LO, GO, NO>(xMat));
}
#endif

} // namespace b200

namespace b200 {
// vec2D_dbl_Type (one std::vector per point) -> flat array; the pointer chase over millions of small vectors is the slowest
// host step of a re-upload of the mesh points, so it runs on a few threads
template <class Points>
inline void flatten_points(const Points &points, std::int64_t nn, int dim, double *xyz)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = nn < (1 << 16) ? 1 : (int)std::max(1u, std::min(16u, hw ? hw : 1u));
    auto part = [&](int t) {
        const std::int64_t k0 = nn * t / nt, k1 = nn * (t + 1) / nt;
        for (std::int64_t k = k0; k < k1; k++)
            for (int c = 0; c < dim; c++) xyz[(std::size_t)k * dim + c] = points[(std::size_t)k][(std::size_t)c];
    };
    if (nt == 1) { part(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(part, t);
    for (std::thread &t : th) t.join();
}
template <class Points>
inline void flatten_points(const Points &points, std::int64_t nn, int dim, std::vector<double> &xyz)
{
    xyz.resize((std::size_t)nn * dim);
    flatten_points(points, nn, dim, xyz.data());
}
} // namespace b200

template <class SC = double, class LO = int, class GO = long long, class NO = int>
class FE_b200 {
  public:
    typedef Domain<SC, LO, GO, NO> Domain_Type;
    typedef Teuchos::RCP<const Domain_Type> DomainConstPtr_Type;
    typedef Matrix<SC, LO, GO, NO> Matrix_Type;
    typedef Teuchos::RCP<Matrix_Type> MatrixPtr_Type;
    typedef Map<LO, GO, NO> Map_Type;
    typedef Teuchos::RCP<const Map_Type> MapConstPtr_Type;
    typedef MultiVector<SC, LO, GO, NO> MultiVector_Type;
    typedef Teuchos::RCP<MultiVector_Type> MultiVectorPtr_Type;

    explicit FE_b200(bool /*saveAssembly*/ = false, int device = 0) : ctx_(nullptr)
    {
        b200::check(feddb200_create(&ctx_, device));
        feddb200_bind_host_numa(ctx_, nullptr);   // this rank's thread and its page-locked buffers next to its GPU
        pool_ = std::make_shared<b200::PinnedPool>(ctx_);
    }
    ~FE_b200()
    {
        pool_->retire();
        for (auto &kv : plans_) feddb200_halo_free(kv.second->halo);
        for (auto &kv : pats_) feddb200_pat_free(kv.second);
        for (auto &s : slots_) { feddb200_mesh_free(s.mesh); if (s.xyz_pinned) feddb200_host_free(ctx_, s.xyz_pinned); }
        feddb200_destroy(ctx_);
    }
    FE_b200(const FE_b200 &) = delete;
    FE_b200 &operator=(const FE_b200 &) = delete;

    // FE::addFE (FE_def.hpp:64-72): registers the domain; here also the one-time mesh upload
    void addFE(DomainConstPtr_Type domain)
    {
        Slot s;
        s.domain = domain;
        s.dim = domain->getDimension();
        s.FEType = domain->getFEType();
        auto elements = domain->getElementsC();
        auto points = domain->getPointsRepeated();
        s.ne = elements->numberElements();
        if (s.FEType == "P0") {
            // pressure space of assemblyDivAndDivT (FE_def.hpp:1954-1957): one pseudo-node per element, numbered by the
            // element map; the node lists and points of the domain are not read
            s.nloc = 1;
            s.nn = s.ne;
            s.p0 = true;
            std::vector<std::int32_t> conn0((std::size_t)s.ne);
            for (std::int64_t T = 0; T < s.ne; T++) conn0[(std::size_t)T] = (std::int32_t)T;
            std::vector<double> zeros((std::size_t)std::max<std::int64_t>(s.ne, 1) * s.dim, 0.0);
            b200::check(feddb200_mesh_upload(ctx_, &s.mesh, s.dim, 1, s.ne, conn0.data(), s.nn, zeros.data()));
            slots_.push_back(s);
            return;
        }
        s.nloc = s.ne > 0 ? (int)elements->getElement(0).getVectorNodeList().size() : nloc_of(s.dim, s.FEType);
        s.nn = (std::int64_t)points->size();
        std::vector<std::int32_t> conn((std::size_t)s.ne * s.nloc);
        for (std::int64_t T = 0; T < s.ne; T++) {
            const std::vector<int> &nodes = elements->getElement((int)T).getVectorNodeList();
            for (int i = 0; i < s.nloc; i++) conn[(std::size_t)T * s.nloc + i] = nodes[i];
        }
        std::vector<double> xyz;
        b200::flatten_points(*points, s.nn, s.dim, xyz);
        b200::check(feddb200_mesh_upload(ctx_, &s.mesh, s.dim, s.nloc, s.ne, conn.data(), s.nn, xyz.data()));
        slots_.push_back(s);
    }

    // Multi-rank use (one FE_b200 per rank / GPU).  The communicator is two callbacks (include/feddb200_halo.h): an MPI build
    // passes MPI_Alltoallv on the problem's Teuchos communicator, the tests an in-process one.  Set it BEFORE addFE, and
    // register every domain with the rank that owns each repeated node (the unique map of the reference:
    // Map::buildUniqueMap, core/LinearAlgebra/Map_def.hpp:184-212; with Trilinos: mapUnique->getRemoteIndexList).
    // Rows then live on the unique map, columns on the Tpetra column map (owned, then remotes by (owner, gid)), and every
    // assembly ends with the host-side globalAssemble of Matrix::fillComplete (Matrix_def.hpp:192-199): the entries this
    // rank computed for rows of other ranks travel to their owners and are added there.
    void setCommunicator(const feddb200_comm &comm) { comm_ = comm; }
    void addFE(DomainConstPtr_Type domain, const std::vector<int> &ownerRanks)
    {
        addFE(domain);
        if ((std::int64_t)ownerRanks.size() != slots_.back().nn) throw std::logic_error("addFE: one owner rank per repeated node");
        slots_.back().owner.assign(ownerRanks.begin(), ownerRanks.end());
    }
    int numRanks() const { return comm_.size; }

    // re-reads the points of a registered domain and uploads them (moving meshes: the Geometry / FSI problems displace the
    // mesh between assemblies, problems/specific/FSI_def.hpp; connectivity and pattern stay)
    void updatePoints(int loc)
    {
        Slot &s = slots_.at((std::size_t)loc);
        auto points = s.domain->getPointsRepeated();
        // flat page-locked staging buffer, kept with the slot: no 67 MB allocation + page faults per call (config 3) and the
        // upload runs at PCIe speed instead of through the driver's bounce buffer (measured 37 ms -> see bench `cpp_host`)
        if (!s.xyz_pinned) {
            void *q = nullptr;
            b200::check(feddb200_host_alloc(ctx_, &q, (std::int64_t)sizeof(double) * std::max<std::int64_t>(s.nn * s.dim, 1)));
            s.xyz_pinned = static_cast<double *>(q);
        }
        else b200::check(feddb200_synchronize(ctx_));   // the previous upload from this buffer has left the host
        b200::flatten_points(*points, s.nn, s.dim, s.xyz_pinned);
        b200::check(feddb200_mesh_update_coords(ctx_, s.mesh, s.xyz_pinned));
    }

    void setScatterMode(int mode) { b200::check(feddb200_set_scatter_mode(ctx_, mode)); }

    // FE::assemblyRHS (FE_def.hpp:4694-4766): constant source; func is evaluated once on the host (:4735), the last
    // entry of funcParameter is the degree of the function (:4716); the result is ADDED to the repeated vector `a`
    template <class RhsFunc>
    void assemblyRHS(int dim, std::string FEType, MultiVectorPtr_Type a, std::string fieldType, RhsFunc func,
                     std::vector<SC> &funcParameter)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        if (a.is_null()) throw std::runtime_error("MultiVector in assemblyConstRHS is null.");
        if (a->getNumVectors() > 1) throw std::logic_error("Implement for numberMV > 1 .");
        const bool vec = fieldType == "Vector";
        if (!vec && fieldType != "Scalar") throw std::logic_error("Invalid field type.");
        const int loc = checkFE(dim, FEType);
        const int degFunc = (int)(funcParameter[funcParameter.size() - 1] + 1.e-14);
        double x = 0.0;
        std::vector<double> valueFunc(3, 0.0);
        func(&x, &valueFunc[0], &funcParameter[0]);
        feddb200_pat *p = pattern(loc, loc);
        std::vector<double> rhs((std::size_t)slots_[loc].nn * (vec ? dim : 1));
        b200::check(feddb200_assemble_rhs(ctx_, p, vec ? 1 : 0, degFunc, valueFunc.data(), rhs.data()));
        Teuchos::ArrayRCP<SC> valuesRhs = a->getDataNonConst(0);
        for (std::size_t k = 0; k < rhs.size(); k++) valuesRhs[k] += rhs[k];
    }

    // FE::assemblyStress (FE_def.hpp:2407-2735): func is evaluated on the host at the physical quadrature points
    // x_k = B q_k + p_1 of every element, where the reference evaluates it (:2479-2484, :2599-2608)
    template <class CoeffFunc>   // the reference's CoeffFunc_Type = boost::function<double(double *x, int *parameters)>
    void assemblyStress(int dim, std::string FEType, MatrixPtr_Type &A, CoeffFunc func, int *parameters,
                        bool callFillComplete = true)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        const auto &el = *slots_[loc].domain->getElementsC();
        const auto &pts = *slots_[loc].domain->getPointsRepeated();
        const int nloc = slots_[loc].nloc;
        int nq = 0;
        b200::check(feddb200_stress_quadrature(dim, nloc, &nq, nullptr, nullptr));
        std::vector<double> q((std::size_t)nq * dim);
        b200::check(feddb200_stress_quadrature(dim, nloc, &nq, q.data(), nullptr));
        const std::size_t ne = el.numberElements();
        std::vector<double> coef(ne * nq);
        bool constant = true;
        for (std::size_t T = 0; T < ne; T++) {
            const std::vector<int> &nodes = el.getElement(T).getVectorNodeList();
            const std::vector<double> &p1 = pts.at(nodes[0]);
            for (int k = 0; k < nq; k++) {
                double x[3] = {0., 0., 0.};
                for (int r = 0; r < dim; r++)
                    for (int c = 0; c < dim; c++) x[c] += (pts.at(nodes[r + 1])[c] - p1[c]) * q[(std::size_t)k * dim + r];
                for (int c = 0; c < dim; c++) x[c] += p1[c];
                coef[T * nq + k] = func(x, parameters);
                constant = constant && coef[T * nq + k] == coef[0];
            }
        }
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_FULL, csr);
        if (constant) b200::check(feddb200_assemble_stress(ctx_, p, ne ? coef[0] : 1.0, nullptr, 0, csr.values.get()));
        else b200::check(feddb200_assemble_stress(ctx_, p, 1.0, coef.data(), (int64_t)coef.size(), csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_FULL, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyBDStabilization (FE_def.hpp:2151-2220): P1 only, like the reference (:2156)
    void assemblyBDStabilization(int dim, std::string FEType, MatrixPtr_Type &A, bool callFillComplete = true)
    {
        if (FEType != "P1")
            throw std::logic_error("Only implemented for P1. Q1 is equivalent but we need to adjust scaling for the reference element.");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        b200::check(feddb200_assemble_bdstab(ctx_, p, csr.values.get()));
        globalAssemble(p, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyMass (FE_def.hpp:454-521): fieldType "Scalar" or "Vector" (same value on the dim diagonal blocks)
    void assemblyMass(int dim, std::string FEType, std::string fieldType, MatrixPtr_Type &A, bool callFillComplete = true)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = checkFE(dim, FEType);
        const bool vec = fieldType == "Vector";
        if (!vec && fieldType != "Scalar") throw std::logic_error("Specify valid vieldType for assembly of mass matrix.");
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        if (vec) expand(p, loc, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        else expand(p, loc, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        b200::check(feddb200_assemble_mass(ctx_, p, vec ? 1 : 0, csr.values.get()));
        if (vec) globalAssemble(p, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        else globalAssemble(p, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyLaplace (FE_def.hpp:604-667); `degree` is ignored exactly as in the reference (:626)
    void assemblyLaplace(int dim, std::string FEType, int /*degree*/, MatrixPtr_Type &A, bool callFillComplete = true,
                         int FELocExternal = -1)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = FELocExternal < 0 ? checkFE(dim, FEType) : FELocExternal;
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        b200::check(feddb200_assemble_laplace(ctx_, p, 0, csr.values.get()));
        globalAssemble(p, 1, 1, FEDDB200_BLOCK_SCALAR, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyLaplaceVecField (FE_def.hpp:670-734)
    void assemblyLaplaceVecField(int dim, std::string FEType, int /*degree*/, MatrixPtr_Type &A, bool callFillComplete = true)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        b200::check(feddb200_assemble_laplace(ctx_, p, 1, csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyLinElasXDim (FE_def.hpp:2739-3040)
    void assemblyLinElasXDim(int dim, std::string FEType, MatrixPtr_Type &A, double lambda, double mu, bool callFillComplete = true)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_FULL, csr);
        b200::check(feddb200_assemble_linelas(ctx_, p, lambda, mu, csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_FULL, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyAdvectionVecField (FE_def.hpp:1685-1836); u is the repeated, node-wise interleaved velocity
    void assemblyAdvectionVecField(int dim, std::string FEType, MatrixPtr_Type &A, MultiVectorPtr_Type u, bool callFillComplete)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        if (u->getNumVectors() > 1) throw std::logic_error("Implement for numberMV > 1 .");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        Teuchos::ArrayRCP<const SC> uArray = u->getData(0);
        b200::check(feddb200_assemble_advection(ctx_, p, (uArray.size() ? &uArray[0] : nullptr), csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_DIAG, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyAdvectionInUVecField (FE_def.hpp:1839-1929)
    void assemblyAdvectionInUVecField(int dim, std::string FEType, MatrixPtr_Type &A, MultiVectorPtr_Type u, bool callFillComplete)
    {
        if (FEType == "P0") throw std::logic_error("Not implemented for P0");
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_FULL, csr);
        Teuchos::ArrayRCP<const SC> uArray = u->getData(0);
        b200::check(feddb200_assemble_advection_in_u(ctx_, p, (uArray.size() ? &uArray[0] : nullptr), csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_FULL, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::assemblyDivAndDivT (FE_def.hpp:1932-2057): FEType1 = velocity space, FEType2 = pressure space; rows of
    // Bmat are pressure dofs, fillComplete(map1, map2) / (map2, map1) (:2052-2055)
    void assemblyDivAndDivT(int dim, std::string FEType1, std::string FEType2, int /*degree*/, MatrixPtr_Type &Bmat,
                            MatrixPtr_Type &BTmat, MapConstPtr_Type map1, MapConstPtr_Type map2, bool callFillComplete = true)
    {
        if (FEType2 == "P1-disc" || FEType2 == "P1-disc-global")
            throw std::logic_error("assemblyDivAndDivT: P1-disc pressure is not implemented in the B200 engine (the reference pairs it with Q2 meshes)");
        // FE::phi has a P0 case for dim 1 and 2 only (FE_def.hpp:4955, 4993): in 3D the reference reads an uninitialised value
        if (FEType2 == "P0" && dim != 2) throw std::logic_error("assemblyDivAndDivT: P0 pressure is implemented for dim == 2 only (as in the reference)");
        const int loc1 = checkFE(dim, FEType1), loc2 = checkFE(dim, FEType2);
        feddb200_pat *pB = pattern(loc2, loc1), *pBT = pattern(loc1, loc2);
        b200::LocalCsr<SC, LO, GO> cB, cBT;
        expand(pB, loc1, 1, dim, FEDDB200_BLOCK_FULL, cB);
        expand(pBT, loc2, dim, 1, FEDDB200_BLOCK_FULL, cBT);
        b200::check(feddb200_assemble_div_divT(ctx_, pB, pBT, cB.values.get(), cBT.values.get()));
        seat_csr(Bmat, cB, map1, map2, callFillComplete);
        seat_csr(BTmat, cBT, map2, map1, callFillComplete);
    }
    // FE::assemblyDivAndDivTFast (FE_def.hpp:2061-2148): same matrices, single-entry inserts in the reference
    void assemblyDivAndDivTFast(int dim, std::string FEType1, std::string FEType2, int degree, MatrixPtr_Type &Bmat,
                                MatrixPtr_Type &BTmat, MapConstPtr_Type map1, MapConstPtr_Type map2, bool callFillComplete = true)
    {
        assemblyDivAndDivT(dim, FEType1, FEType2, degree, Bmat, BTmat, map1, map2, callFillComplete);
    }

    // Fused (0,0) block of the Navier-Stokes system: rho*nu*LaplaceVecField + rho*N(u) [+ rho*W(u)], the three
    // assemblies and two TwoMatrixAdd of NavierStokes::reAssemble (problems/specific/NavierStokes_def.hpp:140-152,
    // 297-313) in one pass on the union pattern.  Not a member of the reference's FE; optional fast path.
    void assemblyNavierStokesJacobian(int dim, std::string FEType, MatrixPtr_Type &A, MultiVectorPtr_Type u, double rho,
                                      double nu, bool newton, bool callFillComplete = true)
    {
        const int loc = checkFE(dim, FEType);
        feddb200_pat *p = pattern(loc, loc);
        b200::LocalCsr<SC, LO, GO> csr;
        expand(p, loc, dim, dim, FEDDB200_BLOCK_FULL, csr);
        Teuchos::ArrayRCP<const SC> uArray = u->getData(0);
        b200::check(feddb200_assemble_ns_jacobian(ctx_, p, rho, nu, (uArray.size() ? &uArray[0] : nullptr), newton ? 1 : 0, csr.values.get()));
        globalAssemble(p, dim, dim, FEDDB200_BLOCK_FULL, csr);
        seat_csr(A, csr, MapConstPtr_Type(), MapConstPtr_Type(), callFillComplete);
    }

    // FE::checkFE (FE_def.hpp:6932-6953): last registered domain with this dimension and FE type
    int checkFE(int dim, std::string FEType)
    {
        int loc = -1;
        for (std::size_t i = 0; i < slots_.size(); i++)
            if (slots_[i].dim == dim && slots_[i].FEType == FEType) loc = (int)i;
        if (loc < 0)
            throw std::logic_error("Combination of dimenson(2/3) and FE Type(P1/P2) not defined yet. Use addFE(domain)");
        return loc;
    }

    std::int64_t launchCount() const { return feddb200_launch_count(ctx_); }

  private:
    struct Slot {
        DomainConstPtr_Type domain;
        feddb200_mesh *mesh = nullptr;
        int dim = 0, nloc = 0;
        std::int64_t ne = 0, nn = 0;
        std::string FEType;
        std::vector<std::int32_t> owner;   // multi-rank: owning rank of every repeated node
        bool p0 = false;                   // P0 space: "nodes" are the elements, numbered by Domain::getElementMap
        double *xyz_pinned = nullptr;      // page-locked staging of the flat points (updatePoints)
    };

    // multi-rank plan of one pattern + the state of its node-pattern callback
    struct Plan {
        FE_b200 *self = nullptr;
        int loc = 0;
        feddb200_halo *halo = nullptr;
        feddb200_pat *pat = nullptr;          // pattern of the last callback (the final one stays)
        std::vector<std::int64_t> rowptr;
        std::vector<std::int32_t> colind;
    };
    static int pattern_callback(void *user, const std::int32_t *row_lid, std::int64_t n_rows, std::int64_t n_owned, const std::int32_t *col_lid,
                                std::int64_t n_cols, const std::int32_t *extra_row, const std::int32_t *extra_col, std::int64_t n_extra,
                                const std::int64_t **rowptr, const std::int32_t **colind)
    {
        Plan &P = *static_cast<Plan *>(user);
        FE_b200 &F = *P.self;
        if (P.pat) { feddb200_pat_free(P.pat); P.pat = nullptr; }
        const feddb200_mesh *m = F.slots_[(std::size_t)P.loc].mesh;
        int rc = feddb200_pattern_build(F.ctx_, &P.pat, m, m, n_rows, n_owned, row_lid, n_cols, col_lid, n_extra, extra_row, extra_col);
        if (rc != FEDDB200_OK) return rc;
        std::int64_t nnz = 0;
        rc = feddb200_pattern_info(P.pat, nullptr, nullptr, nullptr, &nnz, nullptr, nullptr, nullptr);
        if (rc != FEDDB200_OK) return rc;
        P.rowptr.resize((std::size_t)n_rows + 1);
        P.colind.resize((std::size_t)std::max<std::int64_t>(nnz, 1));
        rc = feddb200_pattern_get_nodes(F.ctx_, P.pat, P.rowptr.data(), P.colind.data());
        *rowptr = P.rowptr.data();
        *colind = P.colind.data();
        return rc;
    }

    static int nloc_of(int dim, const std::string &fe)
    {
        if (fe == "P1") return dim + 1;
        if (fe == "P2") return dim == 2 ? 6 : 10;
        throw std::logic_error("FE_b200: only P1/P2 triangles and tetrahedra are implemented");
    }

    // one pattern per (row space, column space), built on first use (single rank: the unique map lists the
    // repeated nodes in order, Map_def.hpp:201-206; the multi-rank row/column maps are described in INTEGRATION.md)
    feddb200_pat *pattern(int rowLoc, int colLoc)
    {
        const std::pair<int, int> key(rowLoc, colLoc);
        auto it = pats_.find(key);
        if (it != pats_.end()) return it->second;
        feddb200_pat *p = nullptr;
        if (comm_.size > 1) {
            if (rowLoc != colLoc)
                throw std::logic_error("FE_b200: patterns over two FE spaces (B, B^T) are single-rank in this version");
            Slot &s = slots_[(std::size_t)rowLoc];
            if ((std::int64_t)s.owner.size() != s.nn) throw std::logic_error("FE_b200: multi-rank use needs addFE(domain, ownerRanks)");
            std::unique_ptr<Plan> P(new Plan());
            P->self = this; P->loc = rowLoc;
            MapConstPtr_Type mapRep = s.domain->getMapRepeated();
            std::vector<std::int64_t> gid((std::size_t)s.nn);
            for (std::int64_t k = 0; k < s.nn; k++) gid[(std::size_t)k] = (std::int64_t)mapRep->getGlobalElement((LO)k);
            b200::check(feddb200_halo_create(&P->halo, &comm_, s.nn, gid.data(), s.owner.data(), &FE_b200::pattern_callback, P.get()));
            p = P->pat;
            plans_[p] = std::move(P);
        } else
        b200::check(feddb200_pattern_build(ctx_, &p, slots_[rowLoc].mesh, slots_[colLoc].mesh, 0, 0, nullptr, 0, nullptr, 0,
                                           nullptr, nullptr));
        pats_[key] = p;
        return p;
    }

    // dof-level CSR + column map of global dof ids (node-wise numbering dofs*g + d, Map_def.hpp:95-108): expanded and
    // downloaded once per (pattern, layout), then shared; the values come from the pinned pool
    void expand(feddb200_pat *p, int colLoc, int rowDofs, int colDofs, int mode, b200::LocalCsr<SC, LO, GO> &csr)
    {
        const std::int64_t nnz = feddb200_pattern_nnz(p, rowDofs, colDofs, mode);
        const std::tuple<feddb200_pat *, int, int, int> key(p, rowDofs, colDofs, mode);
        auto it = structures_.find(key);
        auto pl = plans_.find(p);
        if (it == structures_.end()) {
            std::int64_t nRows = 0, nCols = 0, nOwned = 0;
            b200::check(feddb200_pattern_info(p, &nRows, &nOwned, &nCols, nullptr, nullptr, nullptr, nullptr));
            std::shared_ptr<b200::CsrStructure<GO> > st = std::make_shared<b200::CsrStructure<GO> >();
            st->rowptr.resize((std::size_t)nRows * rowDofs + 1);
            st->colind.resize((std::size_t)nnz);
            b200::check(feddb200_pattern_expand(ctx_, p, rowDofs, colDofs, mode, st->rowptr.data(), st->colind.data()));
            if (pl == plans_.end()) {
                MapConstPtr_Type mapRep = slots_[colLoc].p0 ? slots_[colLoc].domain->getElementMap() : slots_[colLoc].domain->getMapRepeated();
                st->colmap.resize((std::size_t)nCols * colDofs);
                for (std::int64_t j = 0; j < nCols; j++)
                    for (int d = 0; d < colDofs; d++)
                        st->colmap[(std::size_t)j * colDofs + d] = (GO)colDofs * mapRep->getGlobalElement((LO)j) + d;
            } else {
                // multi-rank: the matrix has the OWNED rows (unique map) and the columns of the Tpetra column map; the ghost
                // rows behind them only exist in the value buffer until the globalAssemble step has shipped them
                const std::int64_t nnzOwned = feddb200_pattern_nnz_owned(p, rowDofs, colDofs, mode);
                st->rowptr.resize((std::size_t)nOwned * rowDofs + 1);
                st->colind.resize((std::size_t)nnzOwned);
                std::int64_t nColmap = 0;
                const std::int64_t *cg = static_cast<const std::int64_t *>(feddb200_halo_array(pl->second->halo, 4, &nColmap));
                st->colmap.resize((std::size_t)nColmap * colDofs);
                for (std::int64_t j = 0; j < nColmap; j++)
                    for (int d = 0; d < colDofs; d++) st->colmap[(std::size_t)j * colDofs + d] = (GO)colDofs * (GO)cg[j] + d;
            }
            it = structures_.insert(std::make_pair(key, std::shared_ptr<const b200::CsrStructure<GO> >(st))).first;
        }
        csr.pattern = it->second;
        csr.values = pool_->take((std::size_t)nnz);     // owned + ghost values
        csr.nnz = pl == plans_.end() ? (std::size_t)nnz : (std::size_t)feddb200_pattern_nnz_owned(p, rowDofs, colDofs, mode);
    }

    // the globalAssemble step of Matrix::fillComplete (multi-rank): ghost-row values to their owners, added there
    void globalAssemble(feddb200_pat *p, int rowDofs, int colDofs, int mode, b200::LocalCsr<SC, LO, GO> &csr)
    {
        auto pl = plans_.find(p);
        if (pl == plans_.end()) return;
        b200::check(feddb200_halo_export_add(pl->second->halo, &comm_, rowDofs, colDofs, mode, (std::int64_t)csr.nnz, csr.values.get()));
    }

    feddb200_ctx *ctx_;
    std::vector<Slot> slots_;
    std::map<std::pair<int, int>, feddb200_pat *> pats_;
    std::map<std::tuple<feddb200_pat *, int, int, int>, std::shared_ptr<const b200::CsrStructure<GO> > > structures_;
    std::shared_ptr<b200::PinnedPool> pool_;
    feddb200_comm comm_ = {nullptr, 0, 1, nullptr};
    std::map<feddb200_pat *, std::unique_ptr<Plan> > plans_;
};

} // namespace FEDD
