// fe_b200_mr_driver -- multi-rank C++ side-by-side check of the drop-in boundary.  R ranks run as threads of this process, one
// FEDD::FE_b200 each (own engine context on the same GPU), talking through an in-process communicator with the two callbacks
// an MPI build would supply (include/feddb200_halo.h).  Reference: FEDDLib's own FE routines (oracle/_ref) inserting the
// elements of ALL ranks into one global matrix -- what Tpetra's insertGlobalValues + fillComplete/globalAssemble produce
// (Matrix_def.hpp:88-92, 192-199).  Checked: every global row is owned by exactly one rank, its pattern equals the reference
// row (columns through the rank's column map), values within 1e-12 relative Frobenius error.
// Test infrastructure (may use oracle/); built by __graft_entry__.build(), run by tests/test_gpu_cpp_host.py.
//
//   fe_b200_mr_driver <parts.bin>   int64 header {dim, nloc, R}; per rank: int64 {ne, nn}, conn int32[ne*nloc], coords f64[nn*dim],
//                                   gid int64[nn], owner int32[nn], u f64[dim*nn]
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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
const char *ref_last_error(void);
}

using namespace FEDD;
typedef long long GOx;
typedef int NOx;
typedef Domain<double, int, GOx, NOx> Domain_t;
typedef Map<int, GOx, NOx> Map_t;
typedef Matrix<double, int, GOx, NOx> Matrix_t;
typedef MultiVector<double, int, GOx, NOx> MV_t;
typedef b200::LocalCsr<double, int, GOx> Csr_t;

static std::mutex g_seat_mutex;
static std::map<const void *, Csr_t> g_seated;
namespace FEDD { namespace b200 {
void seat_csr(Teuchos::RCP<Matrix_t> &A, Csr_t &csr, Teuchos::RCP<const Map_t>, Teuchos::RCP<const Map_t>, bool)
{
    std::lock_guard<std::mutex> lock(g_seat_mutex);
    g_seated[A.get()] = csr;
}
}} // namespace FEDD::b200

// ---- in-process communicator: all-to-all-v between the threads ----
struct Shared {
    int size;
    std::vector<std::vector<std::vector<int64_t> > > mail;   // [src][dst]
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0, generation = 0;
    explicit Shared(int n) : size(n), mail(n, std::vector<std::vector<int64_t> >(n)) {}
    void barrier()
    {
        std::unique_lock<std::mutex> lock(m);
        const int gen = generation;
        if (++waiting == size) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lock, [&] { return gen != generation; });
    }
};
struct RankComm { Shared *s; int rank; std::vector<int64_t> recv; };
static int a2a(void *user, const int64_t *send, const int64_t *counts, const int64_t **recv, int64_t *rcounts)
{
    RankComm &c = *static_cast<RankComm *>(user);
    int64_t off = 0;
    for (int d = 0; d < c.s->size; d++) { c.s->mail[c.rank][d].assign(send + off, send + off + counts[d]); off += counts[d]; }
    c.s->barrier();
    c.recv.clear();
    for (int src = 0; src < c.s->size; src++) {
        const std::vector<int64_t> &v = c.s->mail[src][c.rank];
        rcounts[src] = (int64_t)v.size();
        c.recv.insert(c.recv.end(), v.begin(), v.end());
    }
    c.s->barrier();
    if (c.recv.empty()) c.recv.push_back(0);
    *recv = c.recv.data();
    return 0;
}

struct Part {
    int64_t ne, nn;
    std::vector<int32_t> conn, owner;
    std::vector<double> xyz, u;
    std::vector<int64_t> gid;
};

static Teuchos::RCP<Domain_t> make_domain(int dim, const char *fe, const Part &P, int nloc)
{
    Teuchos::RCP<Domain_t> d(new Domain_t(dim, fe));
    d->elementsC_ = Teuchos::rcp(new Elements());
    for (int64_t e = 0; e < P.ne; e++) d->elementsC_->addElement(FiniteElement(std::vector<int>(P.conn.begin() + e * nloc, P.conn.begin() + (e + 1) * nloc)));
    d->pointsRep_ = Teuchos::rcp(new std::vector<std::vector<double> >(P.nn, std::vector<double>(dim, 0.0)));
    for (int64_t k = 0; k < P.nn; k++)
        for (int c = 0; c < dim; c++) (*d->pointsRep_)[k][c] = P.xyz[k * dim + c];
    std::vector<GOx> g(P.gid.begin(), P.gid.end());
    d->mapRepeated_ = Teuchos::RCP<const Map_t>(new Map_t(g.data(), g.size()));
    return d;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s parts.bin\n", argv[0]); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    int64_t h[3];
    if (std::fread(h, 8, 3, f) != 3) return 2;
    const int dim = (int)h[0], nloc = (int)h[1], R = (int)h[2];
    std::vector<Part> parts(R);
    int64_t nglob = 0;
    for (Part &P : parts) {
        int64_t q[2];
        if (std::fread(q, 8, 2, f) != 2) return 2;
        P.ne = q[0]; P.nn = q[1];
        P.conn.resize(P.ne * nloc); P.xyz.resize(P.nn * dim); P.gid.resize(P.nn); P.owner.resize(P.nn); P.u.resize(P.nn * dim);
        bool ok = std::fread(P.conn.data(), 4, P.conn.size(), f) == P.conn.size() && std::fread(P.xyz.data(), 8, P.xyz.size(), f) == P.xyz.size() &&
                  std::fread(P.gid.data(), 8, P.gid.size(), f) == P.gid.size() && std::fread(P.owner.data(), 4, P.owner.size(), f) == P.owner.size() &&
                  std::fread(P.u.data(), 8, P.u.size(), f) == P.u.size();
        if (!ok) { std::fprintf(stderr, "short parts file\n"); return 2; }
        for (int64_t g : P.gid) nglob = std::max(nglob, g + 1);
    }
    std::fclose(f);
    const char *fe = nloc == dim + 1 ? "P1" : "P2";
    const double lambda = 8.0e6, mu = 2.0e6;
    struct Case { const char *name; int op; int rd; };
    const Case cases[] = {{"assemblyLaplace", 0, 1}, {"assemblyLaplaceVecField", 1, dim}, {"assemblyLinElasXDim", 2, dim},
                          {"assemblyAdvectionVecField", 3, dim}, {"assemblyAdvectionInUVecField", 4, dim}, {"assemblyMass Vector", 8, dim}};
    bool all = true;
    for (const Case &c : cases) {
        // reference: all ranks insert into one global matrix
        fo_matrix *rA = fo_matrix_new(c.rd * nglob, 64), *rB = fo_matrix_new(1, 8);
        for (const Part &P : parts)
            if (ref_assemble(c.op, dim, fe, fe, P.ne, P.conn.data(), nloc, P.xyz.data(), P.nn, P.gid.data(), nullptr, 0, 0, nullptr, P.u.data(), lambda, mu, rA, rB) != 0) {
                std::printf("reference failed: %s\n", ref_last_error());
                return 1;
            }
        std::vector<int64_t> rrp(c.rd * nglob + 1), rcol(fo_matrix_nnz(rA));
        std::vector<double> rval(rcol.size());
        fo_get_csr(rA, rrp.data(), rcol.data(), rval.data());
        fo_matrix_free(rA); fo_matrix_free(rB);

        Shared shared(R);
        std::vector<Csr_t> got(R);
        std::vector<std::string> errs(R);
        std::vector<std::thread> th;
        for (int r = 0; r < R; r++)
            th.emplace_back([&, r] {
                try {
                    RankComm rc{&shared, r, {}};
                    feddb200_comm comm = {&rc, r, R, a2a};
                    FE_b200<double, int, GOx, NOx> fb;
                    fb.setCommunicator(comm);
                    Teuchos::RCP<Domain_t> d = make_domain(dim, fe, parts[r], nloc);
                    fb.addFE(d, std::vector<int>(parts[r].owner.begin(), parts[r].owner.end()));
                    Teuchos::RCP<MV_t> uMV(new MV_t(parts[r].u.data(), parts[r].u.size()));
                    Teuchos::RCP<Matrix_t> A(new Matrix_t(nullptr));
                    switch (c.op) {
                    case 0: fb.assemblyLaplace(dim, fe, 2, A); break;
                    case 1: fb.assemblyLaplaceVecField(dim, fe, 2, A); break;
                    case 2: fb.assemblyLinElasXDim(dim, fe, A, lambda, mu); break;
                    case 3: fb.assemblyAdvectionVecField(dim, fe, A, uMV, true); break;
                    case 4: fb.assemblyAdvectionInUVecField(dim, fe, A, uMV, true); break;
                    case 8: fb.assemblyMass(dim, fe, "Vector", A); break;
                    }
                    std::lock_guard<std::mutex> lock(g_seat_mutex);
                    got[r] = g_seated[A.get()];
                } catch (const std::exception &e) { errs[r] = e.what(); }
            });
        for (std::thread &t : th) t.join();
        for (int r = 0; r < R; r++)
            if (!errs[r].empty()) { std::printf("%s: rank %d failed: %s\n", c.name, r, errs[r].c_str()); return 1; }

        std::vector<int> seen(c.rd * nglob, 0);
        bool pattern_ok = true;
        double num = 0.0, den = 0.0;
        long long nnz_total = 0;
        for (int r = 0; r < R && pattern_ok; r++) {
            std::vector<int64_t> ugid;   // unique map = owned nodes in repeated order
            for (int64_t k = 0; k < parts[r].nn; k++)
                if (parts[r].owner[k] == r) ugid.push_back(parts[r].gid[k]);
            const std::vector<std::int64_t> &rp = got[r].pattern->rowptr;
            const std::vector<std::int32_t> &ci = got[r].pattern->colind;
            const std::vector<GOx> &cm = got[r].pattern->colmap;
            const double *val = got[r].values.get();
            if ((int64_t)rp.size() - 1 != (int64_t)ugid.size() * c.rd) { pattern_ok = false; break; }
            nnz_total += (long long)got[r].nnz;
            for (int64_t lr = 0; lr + 1 < (int64_t)rp.size() && pattern_ok; lr++) {
                const int64_t gr = c.rd * ugid[lr / c.rd] + lr % c.rd;
                seen[gr]++;
                const int64_t a0 = rp[lr], a1 = rp[lr + 1], b0 = rrp[gr], b1 = rrp[gr + 1];
                if (a1 - a0 != b1 - b0) { pattern_ok = false; break; }
                std::vector<std::pair<GOx, double> > row;
                for (int64_t k = a0; k < a1; k++) row.push_back(std::make_pair(cm[ci[k]], val[k]));
                std::sort(row.begin(), row.end());
                for (int64_t k = 0; k < a1 - a0; k++) {
                    if (row[k].first != rcol[b0 + k]) { pattern_ok = false; break; }
                    const double dd = row[k].second - rval[b0 + k];
                    num += dd * dd; den += rval[b0 + k] * rval[b0 + k];
                }
            }
        }
        for (int v : seen) pattern_ok = pattern_ok && v == 1;
        pattern_ok = pattern_ok && nnz_total == (long long)rcol.size();
        const double err = std::sqrt(num) / (den > 0 ? std::sqrt(den) : 1.0);
        const bool ok = pattern_ok && err <= 1e-12;
        std::printf("%-30s ranks %d  nnz %10lld  pattern %s  rel.Frobenius %.3e  %s\n", c.name, R, nnz_total, pattern_ok ? "identical" : "DIFFERS", err, ok ? "PASS" : "FAIL");
        all = all && ok;
        g_seated.clear();
    }
    std::printf(all ? "ALL PASS\n" : "SOME FAILED\n");
    return all ? 0 : 1;
}
