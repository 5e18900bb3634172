// bench_fe_b200 -- end-to-end timing of the C++ drop-in: FEDD::FE_b200::assemblyLinElasXDim on the built-in structured cube
// (P2 tetrahedra, H/h = M), host containers in, fill-complete host CSR out.  Every timed call includes the upload of the
// points (FE_b200::updatePoints: pageable std::vector -> device), the assembly on the GPU and the download of the CSR values
// into a pooled page-locked buffer that the returned matrix owns.  One JSON line on stdout.
//   usage: bench_fe_b200 [M=70] [steps=3] [device=0]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "standalone.hpp"

using namespace FEDD;
typedef long long GOx;
typedef Domain<double, int, GOx, int> Domain_t;
typedef Matrix<double, int, GOx, int> Matrix_t;

// the reference's structured cube (MeshStructured::buildMesh3D, MeshStructured_def.hpp:808-994): M^3 cells of 6 Kuhn
// tetrahedra, P2 nodes on the half grid, node order: 4 vertices, then the edge midpoints (0,1),(1,2),(0,2),(0,3),(1,3),(2,3)
static Teuchos::RCP<Domain_t> cube_p2(int M)
{
    static const int TET[6][4] = {{1, 0, 5, 7}, {4, 0, 5, 7}, {1, 0, 3, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 6, 7}};
    static const int MID[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    const int n = 2 * M + 1;
    Teuchos::RCP<Domain_t> d(new Domain_t(3, "P2"));
    d->elementsC_ = Teuchos::rcp(new Elements());
    d->elementsC_->reserve((std::size_t)6 * M * M * M);
    const double h = 1.0 / M;
    d->pointsRep_ = Teuchos::rcp(new std::vector<std::vector<double> >((std::size_t)n * n * n, std::vector<double>(3)));
    std::vector<GOx> gid((std::size_t)n * n * n);
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                const std::size_t id = ((std::size_t)k * n + j) * n + i;
                (*d->pointsRep_)[id][0] = 0.5 * h * i; (*d->pointsRep_)[id][1] = 0.5 * h * j; (*d->pointsRep_)[id][2] = 0.5 * h * k;
                gid[id] = (GOx)id;
            }
    d->mapRepeated_ = Teuchos::RCP<const Map<int, GOx, int> >(new Map<int, GOx, int>(gid.data(), gid.size()));
    for (int t = 0; t < M; t++)
        for (int s = 0; s < M; s++)
            for (int r = 0; r < M; r++)
                for (int q = 0; q < 6; q++) {
                    int v[4][3], nodes[10];
                    for (int a = 0; a < 4; a++) {
                        const int c = TET[q][a];
                        v[a][0] = 2 * (r + (c & 1)); v[a][1] = 2 * (s + ((c >> 1) & 1)); v[a][2] = 2 * (t + ((c >> 2) & 1));
                        nodes[a] = (v[a][2] * n + v[a][1]) * n + v[a][0];
                    }
                    for (int e = 0; e < 6; e++) {
                        const int *a = v[MID[e][0]], *b = v[MID[e][1]];
                        nodes[4 + e] = (((a[2] + b[2]) / 2) * n + (a[1] + b[1]) / 2) * n + (a[0] + b[0]) / 2;
                    }
                    d->elementsC_->addElement(FiniteElement(nodes, 10));
                }
    return d;
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    const int M = argc > 1 ? std::atoi(argv[1]) : 70, steps = argc > 2 ? std::atoi(argv[2]) : 3, device = argc > 3 ? std::atoi(argv[3]) : 0;
    try {
        Teuchos::RCP<Domain_t> dom = cube_p2(M);
        FE_b200<double, int, GOx, int> fe(false, device);
        double t0 = now();
        fe.addFE(Teuchos::RCP<const Domain_t>(dom));
        const double t_add = now() - t0;
        Teuchos::RCP<Matrix_t> A;
        t0 = now();
        fe.assemblyLinElasXDim(3, "P2", A, 8.0e6, 2.0e6, true);   // first call: pattern build + expand + pinned pool
        const double t_first = now() - t0;
        double best = 1e30, sum = 0.0, sum_points = 0.0;
        for (int k = 0; k < steps; k++) {
            A = Teuchos::RCP<Matrix_t>();                          // the caller lets go of the previous matrix (Newton / time loop)
            t0 = now();
            fe.updatePoints(0);
            const double t1 = now();
            fe.assemblyLinElasXDim(3, "P2", A, 8.0e6, 2.0e6, true);
            const double dt = now() - t0;
            best = dt < best ? dt : best;
            sum += dt;
            sum_points += t1 - t0;
        }
        double chk = 0.0;
        const std::size_t nchk = A->nnz < ((std::size_t)1 << 20) ? A->nnz : ((std::size_t)1 << 20);
        for (std::size_t k = 0; k < nchk; k++) chk += A->values.get()[k];
        const long long ne = (long long)6 * M * M * M;
        std::printf("{\"host\": \"FE_b200 (C++)\", \"M\": %d, \"elements\": %lld, \"nnz\": %lld, \"steps\": %d, \"ms_per_step\": %.3f, \"ms_best\": %.3f, "
                    "\"update_points_ms\": %.3f, \"addFE_s\": %.3f, \"first_call_s\": %.3f, \"h2d_bytes_per_step\": %lld, \"d2h_bytes_per_step\": %lld, "
                    "\"checksum_first_1Mi_values\": %.10g, \"launches\": %lld}\n",
                    M, ne, (long long)A->nnz, steps, 1e3 * sum / steps, 1e3 * best, 1e3 * sum_points / steps, t_add, t_first,
                    (long long)dom->pointsRep_->size() * 24, (long long)A->nnz * 8, chk, (long long)fe.launchCount());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "bench_fe_b200: %s\n", e.what());
        return 1;
    }
    return 0;
}
