// Device code of the assembly hot path (sm_100a).  Three scatter strategies share the math:
//
//  * element-row kernels (k_elem): one thread per (element, local row[, row dof]).  The thread
//    evaluates its row of the local matrix by quadrature exactly as the reference loops do
//    (FE_def.hpp:637-661 etc.) and adds it to the CSR values either with RED.ADD.F64 atomics
//    (FEDDB200_SCATTER_ATOMIC) or, launched once per element colour, with plain
//    read-modify-write (FEDDB200_SCATTER_COLOURED, deterministic).
//  * row-gather kernels (k_gather): output-stationary.  One thread per CSR row component walks
//    the elements incident to its row node in a fixed order, accumulates in lane-private
//    shared-memory banks and writes every CSR value exactly once with vector stores -- no
//    atomics, no memset, bitwise reproducible.  (FEDDB200_SCATTER_GATHER)
#pragma once
#include "common.cuh"

namespace fb {

// -----------------------------------------------------------------------------------------
// affine map of an element: B[i][j] = x_{j+1}[i] - x_0[i], Binv = adj(B)/det, |det|
// (reference: FE_def.hpp:5342-5357 buildTransformation, SmallMatrix.hpp:306-358)
// -----------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void affine_map(const int32_t *__restrict__ el, const double *__restrict__ coords,
                                           double (&Binv)[DIM][DIM], double &adet)
{
    double B[DIM][DIM], x0[DIM];
    const int64_t n0 = el[0];
#pragma unroll
    for (int i = 0; i < DIM; i++) x0[i] = coords[n0 * DIM + i];
#pragma unroll
    for (int j = 0; j < DIM; j++) {
        const int64_t nj = el[j + 1];
#pragma unroll
        for (int i = 0; i < DIM; i++) B[i][j] = coords[nj * DIM + i] - x0[i];
    }
    if constexpr (DIM == 2) {
        const double det = B[0][0] * B[1][1] - B[1][0] * B[0][1];
        const double r = 1.0 / det;
        Binv[0][0] = B[1][1] * r;
        Binv[0][1] = -B[0][1] * r;
        Binv[1][0] = -B[1][0] * r;
        Binv[1][1] = B[0][0] * r;
        adet = fabs(det);
    } else {
        const double det = B[0][0] * B[1][1] * B[2][2] + B[0][1] * B[1][2] * B[2][0] + B[0][2] * B[1][0] * B[2][1] -
                           B[2][0] * B[1][1] * B[0][2] - B[2][1] * B[1][2] * B[0][0] - B[2][2] * B[1][0] * B[0][1];
        const double r = 1.0 / det;
        Binv[0][0] = (B[1][1] * B[2][2] - B[1][2] * B[2][1]) * r;
        Binv[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * r;
        Binv[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * r;
        Binv[1][0] = (B[1][2] * B[2][0] - B[1][0] * B[2][2]) * r;
        Binv[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * r;
        Binv[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * r;
        Binv[2][0] = (B[1][0] * B[2][1] - B[1][1] * B[2][0]) * r;
        Binv[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * r;
        Binv[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * r;
        adet = fabs(det);
    }
}

// physical gradient g[d] = sum_c ghat[c] * Binv[c][d]   (FE_def.hpp:83-96 applyBTinv)
template <int DIM>
__device__ __forceinline__ void push_grad(const double *__restrict__ ghat, const double (&Binv)[DIM][DIM], double (&g)[DIM])
{
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < DIM; c++) s += ghat[c] * Binv[c][d];
        g[d] = s;
    }
}

struct ElemArgs {
    const int32_t *conn_r, *conn_c, *conn_v; // row / column / velocity(geometry) connectivity
    const double *coords;                    // points of the velocity mesh
    const int32_t *row_lid;                  // nullable
    const int64_t *rowptr;
    const uint16_t *pos;
    int pos_stride;
    const int32_t *elems;                    // nullable: element list of this launch (one colour)
    int64_t n_items;
    const double *u;
    const double *coef;                      // nullable: per (element, quadrature point) coefficient (assemblyStress)
    double c0, c1, c2;                       // lambda,mu | rho*nu, rho, rho(newton)
    const OpTables *tab;
    double *values;
    int vec_dim;                             // LAP: 0 scalar, dim = block-diagonal copies
    int atomic;
};

template <bool ATOMIC>
__device__ __forceinline__ void add_to(double *p, double v)
{
    if constexpr (ATOMIC) atomicAdd(p, v);
    else *p += v;
}

// threads per row node
template <int OP, int DIM> struct RowDofs { static constexpr int value = (OP == OP_ELAS || OP == OP_ADVU || OP == OP_BT || OP == OP_NSJ) ? DIM : 1; };

template <int OP, int DIM, int NR, int NC, bool ATOMIC>
__global__ void __launch_bounds__(128) k_elem(const ElemArgs A)
{
    constexpr int RD = RowDofs<OP, DIM>::value;
    constexpr int NV = (OP == OP_B) ? NC : NR;  // nodes of the velocity (gradient) space
    constexpr int NP = (OP == OP_B) ? NR : NC;  // nodes of the value space (pressure for B/BT)
    __shared__ OpTables T;
    {
        const double *src = reinterpret_cast<const double *>(A.tab);
        double *dst = reinterpret_cast<double *>(&T);
        for (int k = threadIdx.x; k < (int)(sizeof(OpTables) / sizeof(double)); k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= A.n_items * (NR * RD)) return;
    const int64_t item = t / (NR * RD);
    const int rem = (int)(t - item * (NR * RD));
    const int i = rem / RD;
    const int a = rem - i * RD; // row dof handled by this thread (0 when RD == 1)
    const int64_t e = A.elems ? A.elems[item] : item;

    int32_t I = A.conn_r[e * NR + i];
    if (A.row_lid) I = A.row_lid[I];
    if (I < 0) return;

    const int32_t *cv = A.conn_v + e * NV;
    double Binv[DIM][DIM], adet;
    affine_map<DIM>(cv, A.coords, Binv, adet);

    const int nq = T.nq;
    const int64_t base = A.rowptr[I];
    const int64_t L = A.rowptr[I + 1] - base;
    const uint16_t *pos = A.pos + (e * NR + i) * A.pos_stride;

    if constexpr (OP == OP_LAP) {
        double acc[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) acc[j] = 0.0;
        for (int q = 0; q < nq; q++) {
            double gi[DIM];
            push_grad<DIM>(&T.dphi[(q * NV + i) * DIM], Binv, gi);
            const double w = T.w[q];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                double gj[DIM];
                push_grad<DIM>(&T.dphi[(q * NV + j) * DIM], Binv, gj);
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) s += w * gi[d] * gj[d];
                acc[j] += s;
            }
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double v = acc[j] * adet;
            const int64_t p = pos[j];
            if (A.vec_dim == 0) add_to<ATOMIC>(A.values + base + p, v);
            else for (int d = 0; d < DIM; d++) add_to<ATOMIC>(A.values + DIM * base + d * L + p, v);
        }
    } else if constexpr (OP == OP_MASS) {
        // M_ij = |det| sum_q w_q phi_i phi_j  (FE_def.hpp:493-500); "Vector": the same value on the DIM diagonal blocks
#pragma unroll
        for (int j = 0; j < NC; j++) {
            double v = 0.0;
            for (int q = 0; q < nq; q++) v += T.w[q] * T.phi[q * NP + i] * T.phi[q * NP + j];
            v *= adet;
            v -= A.c0 * adet * A.c1;   // assemblyBDStabilization (FE_def.hpp:2207): refElementSize * absDetB * refElementScale; mass: c0 = 0
            const int64_t p = pos[j];
            if (A.vec_dim == 0) add_to<ATOMIC>(A.values + base + p, v);
            else for (int d = 0; d < DIM; d++) add_to<ATOMIC>(A.values + DIM * base + d * L + p, v);
        }
    } else if constexpr (OP == OP_ELAS) {
        const double lambda = A.c0, mu = A.c1;
        double acc[NC][DIM];
#pragma unroll
        for (int j = 0; j < NC; j++)
#pragma unroll
            for (int b = 0; b < DIM; b++) acc[j][b] = 0.0;
        for (int q = 0; q < nq; q++) {
            double gi[DIM];
            push_grad<DIM>(&T.dphi[(q * NV + i) * DIM], Binv, gi);
            // assemblyStress (FE_def.hpp:2515-2518, 2654-2664): funcvalue * weight in front of the tensor product
            const double w = A.coef ? A.coef[e * nq + q] * T.w[q] : T.w[q];
            double gia = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) gia = (d == a) ? gi[d] : gia;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                double gj[DIM];
                push_grad<DIM>(&T.dphi[(q * NV + j) * DIM], Binv, gj);
                double dot = 0.0, gja = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) { dot += gi[d] * gj[d]; gja = (d == a) ? gj[d] : gja; }
#pragma unroll
                for (int b = 0; b < DIM; b++) {
                    // K^{ab}_ij = mu (delta_ab g_i.g_j + g_i[b] g_j[a]) + lambda g_i[a] g_j[b]  (SURVEY A.4,
                    // algebraically equal to the epsilon-tensor form of FE_def.hpp:2939-2993)
                    const double k = mu * ((b == a ? dot : 0.0) + gi[b] * gja) + lambda * gia * gj[b];
                    acc[j][b] += w * k;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int64_t p = pos[j];
#pragma unroll
            for (int b = 0; b < DIM; b++)
                add_to<ATOMIC>(A.values + (int64_t)DIM * DIM * base + (int64_t)a * DIM * L + DIM * p + b, adet * acc[j][b]);
        }
    } else if constexpr (OP == OP_ADV) {
        double ul[NV][DIM];
#pragma unroll
        for (int m = 0; m < NV; m++)
#pragma unroll
            for (int d = 0; d < DIM; d++) ul[m][d] = A.u[(int64_t)DIM * cv[m] + d];
        double acc[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) acc[j] = 0.0;
        for (int q = 0; q < nq; q++) {
            double uq[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) uq[d] = 0.0;
#pragma unroll
            for (int m = 0; m < NV; m++) {
                const double ph = T.phi[q * NP + m];
#pragma unroll
                for (int d = 0; d < DIM; d++) uq[d] += ul[m][d] * ph;
            }
            const double f = T.w[q] * T.phi[q * NP + i];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                double gj[DIM];
                push_grad<DIM>(&T.dphi[(q * NV + j) * DIM], Binv, gj);
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) s += uq[d] * gj[d];
                acc[j] += f * s;
            }
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double v = acc[j] * adet;
            const int64_t p = pos[j];
            for (int d = 0; d < DIM; d++) add_to<ATOMIC>(A.values + DIM * base + d * L + p, v);
        }
    } else if constexpr (OP == OP_ADVU || OP == OP_NSJ) {
        // thread handles row dof d1 = a.  W^{d1 d2}_ij = |det| sum_q w_q D_q[d2][d1] phi_i phi_j with
        // D_q[d2][d1] = sum_m u_{m,d1} g_m(q)[d2]  (FE_def.hpp:1887-1912)
        double ua[NV];
#pragma unroll
        for (int m = 0; m < NV; m++) ua[m] = A.u[(int64_t)DIM * cv[m] + a];
        double acc[NC][DIM];
#pragma unroll
        for (int j = 0; j < NC; j++)
#pragma unroll
            for (int b = 0; b < DIM; b++) acc[j][b] = 0.0;
        for (int q = 0; q < nq; q++) {
            double Dh[DIM], D[DIM];
#pragma unroll
            for (int c = 0; c < DIM; c++) Dh[c] = 0.0;
#pragma unroll
            for (int m = 0; m < NV; m++)
#pragma unroll
                for (int c = 0; c < DIM; c++) Dh[c] += ua[m] * T.dphi[(q * NV + m) * DIM + c];
            push_grad<DIM>(Dh, Binv, D);
            const double wq = T.w[q];
            const double phi_i = T.phi[q * NP + i];
            if constexpr (OP == OP_ADVU) {
                const double f = wq * phi_i;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double fj = f * T.phi[q * NP + j];
#pragma unroll
                    for (int b = 0; b < DIM; b++) acc[j][b] += fj * D[b];
                }
            } else {
                // fused (0,0) block: c0*Laplace (diag) + c1*N(u) (diag) + c2*W(u) (full)
                double gi[DIM], uq[DIM];
                push_grad<DIM>(&T.dphi[(q * NV + i) * DIM], Binv, gi);
#pragma unroll
                for (int d = 0; d < DIM; d++) uq[d] = 0.0;
#pragma unroll
                for (int m = 0; m < NV; m++) {
                    const double ph = T.phi[q * NP + m];
#pragma unroll
                    for (int d = 0; d < DIM; d++) uq[d] += A.u[(int64_t)DIM * cv[m] + d] * ph;
                }
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    double gj[DIM];
                    push_grad<DIM>(&T.dphi[(q * NV + j) * DIM], Binv, gj);
                    double lap = 0.0, adv = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) { lap += gi[d] * gj[d]; adv += uq[d] * gj[d]; }
                    const double diag = wq * (A.c0 * lap + A.c1 * phi_i * adv);
                    const double fj = A.c2 * wq * phi_i * T.phi[q * NP + j];
#pragma unroll
                    for (int b = 0; b < DIM; b++) acc[j][b] += (b == a ? diag : 0.0) + fj * D[b];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int64_t p = pos[j];
#pragma unroll
            for (int b = 0; b < DIM; b++)
                add_to<ATOMIC>(A.values + (int64_t)DIM * DIM * base + (int64_t)a * DIM * L + DIM * p + b, adet * acc[j][b]);
        }
    } else if constexpr (OP == OP_B) {
        // row = pressure node i, cols = (velocity node j, d)   (FE_def.hpp:1991-2017)
        double acc[NC][DIM];
#pragma unroll
        for (int j = 0; j < NC; j++)
#pragma unroll
            for (int d = 0; d < DIM; d++) acc[j][d] = 0.0;
        for (int q = 0; q < nq; q++) {
            const double f = T.w[q] * T.phi[q * NP + i];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                double gj[DIM];
                push_grad<DIM>(&T.dphi[(q * NV + j) * DIM], Binv, gj);
#pragma unroll
                for (int d = 0; d < DIM; d++) acc[j][d] += f * gj[d];
            }
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int64_t p = pos[j];
#pragma unroll
            for (int d = 0; d < DIM; d++) add_to<ATOMIC>(A.values + (int64_t)DIM * base + DIM * p + d, adet * acc[j][d]);
        }
    } else if constexpr (OP == OP_BT) {
        // row = (velocity node i, dof a), cols = pressure nodes j   (FE_def.hpp:2021-2047)
        double acc[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) acc[j] = 0.0;
        for (int q = 0; q < nq; q++) {
            double gi[DIM];
            push_grad<DIM>(&T.dphi[(q * NV + i) * DIM], Binv, gi);
            double gia = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) gia = (d == a) ? gi[d] : gia;
            const double f = T.w[q] * gia;
#pragma unroll
            for (int j = 0; j < NC; j++) acc[j] += f * T.phi[q * NP + j];
        }
#pragma unroll
        for (int j = 0; j < NC; j++)
            add_to<ATOMIC>(A.values + (int64_t)DIM * base + (int64_t)a * L + pos[j], adet * acc[j]);
    }
}

// -----------------------------------------------------------------------------------------
// row-gather path
// -----------------------------------------------------------------------------------------
// per-element geometry cache, one aligned line per element:
//   3D: [e][v = 0..3] = (G_v.x, G_v.y, G_v.z, |det B|)   32 B per vertex, 128 B per element
//   2D: [e] = (G_0.x, G_0.y, G_1.x, G_1.y, G_2.x, G_2.y, |det B|, |det B|)   64 B per element
// with G_v = grad lambda_v, the gradients of the barycentric coordinates.
template <int DIM> struct GeomStride { static constexpr int value = DIM == 3 ? 16 : 8; };

__device__ __forceinline__ void st_v4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int DIM, int NL>
__global__ void __launch_bounds__(256) k_geom(int64_t ne, const int32_t *__restrict__ conn,
                                              const double *__restrict__ coords, double *__restrict__ geom)
{
    constexpr int GS = GeomStride<DIM>::value;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    double Binv[DIM][DIM], adet;
    affine_map<DIM>(conn + e * NL, coords, Binv, adet);
    double *g = geom + e * GS;
    // grad lambda_k = row k-1 of Binv (k >= 1), grad lambda_0 = -(sum of the others)
    double G[DIM + 1][DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; k++) { G[k + 1][d] = Binv[k][d]; s -= Binv[k][d]; }
        G[0][d] = s;
    }
    // whole 32-byte sectors per store instruction (st.global.v4.f64): 4 (2) stores per element instead of 16 (8)
    if constexpr (DIM == 3) {
#pragma unroll
        for (int v = 0; v < 4; v++) st_v4(g + 4 * v, G[v][0], G[v][1], G[v][2], adet);
    } else {
        st_v4(g, G[0][0], G[0][1], G[1][0], G[1][1]);
        st_v4(g + 4, G[2][0], G[2][1], adet, adet);
    }
}

// canonical relabelling of an element seen from local node i: the vertex permutation pi puts
// node i at canonical vertex 0 (vertex nodes) or on canonical edge (0,1) (edge nodes)
template <int DIM>
__host__ __device__ inline void canon_perm(int i, int (&pi)[DIM + 1])
{
    constexpr int NVTX = DIM + 1;
    int first = i, second = -1;
    if (i >= NVTX) {
        const int e = i - NVTX;
        if constexpr (DIM == 2) { const int E[3][2] = {{0, 1}, {1, 2}, {0, 2}}; first = E[e][0]; second = E[e][1]; }
        else { const int E[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}}; first = E[e][0]; second = E[e][1]; }
    }
    int n = 0;
    pi[n++] = first;
    if (second >= 0) pi[n++] = second;
    for (int v = 0; v < NVTX; v++)
        if (v != first && v != second) pi[n++] = v;
}

// actual local node index of canonical local node jc under permutation pi
template <int DIM>
__host__ __device__ inline int canon_node(const int (&pi)[DIM + 1], int jc)
{
    constexpr int NVTX = DIM + 1;
    if (jc < NVTX) return pi[jc];
    const int e = jc - NVTX;
    int a, b;
    if constexpr (DIM == 2) { const int E[3][2] = {{0, 1}, {1, 2}, {0, 2}}; a = pi[E[e][0]]; b = pi[E[e][1]]; }
    else { const int E[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}}; a = pi[E[e][0]]; b = pi[E[e][1]]; }
    if (a > b) { const int t = a; a = b; b = t; }
    if constexpr (DIM == 2) return NVTX + (a == 0 ? (b == 1 ? 0 : 2) : 1);
    else return NVTX + (a == 0 ? (b == 1 ? 0 : (b == 2 ? 2 : 3)) : (a == 1 ? (b == 2 ? 1 : 4) : 5));
}

// -----------------------------------------------------------------------------------------
// Per-incidence gather records (one-time, pattern build).  One record per (row, incident
// element): the positions, in the row, of the element's canonical local nodes, the element
// index and the canonical vertex permutation -- packed so that a thread fetches it with ONE
// vector load:
//   NL <= 4 : 16 B = pos[4] u16 | element u32 | perm u32
//   NL 6, 10: 32 B = pos[NL] u16 (words 0-4) | element u32 | perm u32 | natural-index word
// (NL = nodes of the COLUMN space).  perm word: bits 0-7 pi(3)<<6 | pi(2)<<4 | pi(1)<<2 | pi(0), bits 8-9 ring
// flags (k_ring), bits 16-23 natural local index of canonical nodes 8, 9; the last word holds the natural local
// index (4 bits each) of canonical nodes 0..7 -- the operators with a velocity argument use it to fetch the
// element's nodal values in canonical order.
// -----------------------------------------------------------------------------------------
template <int NL> struct RecWords { static constexpr int value = NL <= 4 ? 4 : 8; };

template <int NL>
__host__ __device__ inline void rec_pack_code(uint32_t *w, uint32_t e, uint32_t perm, uint64_t natidx = 0)
{
    if constexpr (NL <= 4) { w[2] = e; w[3] = perm; }
    else { w[5] = e; w[6] = perm | ((uint32_t)(natidx >> 32) & 0xffu) << 16; w[7] = (uint32_t)natidx; }
}
template <int NL>
__device__ __forceinline__ uint32_t rec_elem(const uint32_t (&w)[RecWords<NL>::value])
{
    if constexpr (NL <= 4) return w[2];
    else return w[5];
}
template <int NL>
__device__ __forceinline__ uint32_t rec_perm(const uint32_t (&w)[RecWords<NL>::value])
{
    if constexpr (NL <= 4) return w[3];
    else return w[6];
}
// natural local index of canonical node jc (P1: the permutation itself)
template <int NL>
__device__ __forceinline__ int rec_natidx(const uint32_t (&w)[RecWords<NL>::value], int jc)
{
    if constexpr (NL <= 4) return (int)((w[3] >> (2 * jc)) & 3);
    else return jc < 8 ? (int)((w[7] >> (4 * jc)) & 15) : (int)((w[6] >> (16 + 4 * (jc - 8))) & 15);
}

__host__ __device__ inline uint64_t sig_mix(uint64_t h, uint32_t v)
{
    h ^= v;
    h *= 0x100000001b3ull;
    return h ^ (h >> 29);
}

// Ring order of the tetrahedra around an edge (3D P2 edge-node rows).  The star of a mesh edge (v0, v1) is a
// union of fans / closed rings of tetrahedra (v0, v1, w_r, w_{r+1}); consecutive tetrahedra share the face
// (v0, v1, w_{r+1}).  Ordering the row's incidences along these chains, with the canonical vertex order
// (v0, v1, w_in, w_out), lets the gather kernel carry the partial sums of the shared face in registers and write
// every other value exactly once -- no shared-memory read-modify-write (k_ring below).
//   order[m]  index (into the row's ascending incidence list) of the m-th incidence in chain order
//   rperm[m]  canonical vertex permutation  l(v0) | l(v1)<<2 | l(w_in)<<4 | l(w_out)<<6
//   rflag[m]  what happens to the contributions of the out-face (v0, v1, w_out):
//             0 carried to the next incidence, 1 stored (chain ends on a boundary face), 2 added to the
//             stored partial of the chain's first in-face (the ring closes)
// Returns false for stars the scheme does not cover (a face shared by more than two tetrahedra, duplicate or
// degenerate elements, more than RING_KMAX incidences): such rows use the generic kernel.
constexpr int RING_KMAX = 32;
__device__ inline bool ring_order(int ninc, const int32_t *__restrict__ inc_row, const int32_t *__restrict__ conn,
                                  int *order, uint32_t *rperm, uint32_t *rflag)
{
    const int E3[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
    int32_t wa[RING_KMAX], wb[RING_KMAX];
    int8_t lp[RING_KMAX], lq[RING_KMAX], la[RING_KMAX], lb[RING_KMAX];
    bool used[RING_KMAX];
    for (int m = 0; m < ninc; m++) {
        const int32_t code = inc_row[m];
        const int64_t e = code >> 4;
        const int ed = (code & 15) - 4;
        int p = E3[ed][0], q = E3[ed][1];
        if (conn[e * 10 + p] > conn[e * 10 + q]) { const int t = p; p = q; q = t; }
        int o[2], n = 0;
        for (int v = 0; v < 4; v++)
            if (v != p && v != q) o[n++] = v;
        lp[m] = (int8_t)p; lq[m] = (int8_t)q; la[m] = (int8_t)o[0]; lb[m] = (int8_t)o[1];
        wa[m] = conn[e * 10 + o[0]]; wb[m] = conn[e * 10 + o[1]];
        used[m] = false;
        if (wa[m] == wb[m]) return false;
    }
    auto count = [&](int32_t w) { int c = 0; for (int m = 0; m < ninc; m++) c += (wa[m] == w) + (wb[m] == w); return c; };
    for (int m = 0; m < ninc; m++) {
        if (count(wa[m]) > 2 || count(wb[m]) > 2) return false;
        for (int m2 = m + 1; m2 < ninc; m2++)
            if ((wa[m] == wa[m2] && wb[m] == wb[m2]) || (wa[m] == wb[m2] && wb[m] == wa[m2])) return false;
    }
    int n_out = 0;
    while (n_out < ninc) {
        int start = -1, low = -1;
        bool in_is_a = true;
        for (int m = 0; m < ninc && start < 0; m++) {
            if (used[m]) continue;
            if (low < 0) low = m;
            if (count(wa[m]) == 1) { start = m; in_is_a = true; }
            else if (count(wb[m]) == 1) { start = m; in_is_a = false; }
        }
        const bool closed = start < 0;
        if (closed) { start = low; in_is_a = true; }
        int cur = start;
        int32_t w_in = in_is_a ? wa[cur] : wb[cur];
        const int32_t first_in = w_in;
        for (;;) {
            used[cur] = true;
            const bool cin_a = wa[cur] == w_in;
            const int l_in = cin_a ? la[cur] : lb[cur], l_out = cin_a ? lb[cur] : la[cur];
            const int32_t w_out = cin_a ? wb[cur] : wa[cur];
            int nxt = -1;
            for (int m = 0; m < ninc && nxt < 0; m++)
                if (!used[m] && (wa[m] == w_out || wb[m] == w_out)) nxt = m;
            uint32_t mode;
            if (nxt >= 0) mode = 0;
            else if (closed) { if (w_out != first_in || cur == start) return false; mode = 2; }
            else mode = 1;
            order[n_out] = cur;
            rperm[n_out] = (uint32_t)lp[cur] | ((uint32_t)lq[cur] << 2) | ((uint32_t)l_in << 4) | ((uint32_t)l_out << 6);
            rflag[n_out] = mode;
            n_out++;
            if (nxt < 0) break;
            cur = nxt;
            w_in = w_out;
        }
    }
    return true;
}

// Chain order of the star of a VERTEX (3D P2 vertex-node rows, k_gather_s).  The tetrahedra around the row vertex I
// are grouped into fans around its edges (I, J) -- greedily, the neighbour J shared by the most tetrahedra not yet
// placed first -- and every fan is walked like an edge ring (ring_order above): canonical vertices (I, J, w_in, w_out),
// consecutive tetrahedra share the face (I, J, w_out).  Along a chain the columns J and mid(I, J) accumulate in
// registers and the contributions to the out-face group are carried to the next incidence, so an incidence costs 4
// shared-memory updates instead of 9 (plus 5 where a chain ends).
//   rflag[m]  0: the out-face group is carried to the next incidence;  1: the chain ends here (everything is flushed)
// Any decomposition is valid for the kernel; fans the walk does not cover (a face shared by more than two tetrahedra,
// duplicate elements) are emitted as chains of length one.
__device__ inline void vchain_order(int ninc, const int32_t *__restrict__ inc_row, const int32_t *__restrict__ conn,
                                    int *order, uint32_t *rperm, uint32_t *rflag)
{
    int8_t l0[RING_KMAX], lo[RING_KMAX][3];
    int32_t go[RING_KMAX][3];
    bool used[RING_KMAX];
    for (int m = 0; m < ninc; m++) {
        const int32_t code = inc_row[m];
        const int64_t e = code >> 4;
        const int i0 = code & 15;
        l0[m] = (int8_t)i0;
        int n = 0;
        for (int v = 0; v < 4; v++)
            if (v != i0) { lo[m][n] = (int8_t)v; go[m][n] = conn[e * 10 + v]; n++; }
        used[m] = false;
    }
    int n_out = 0;
    while (n_out < ninc) {
        // the neighbour vertex shared by the most unplaced tetrahedra (ties: the smallest id)
        int32_t J = -1;
        int best = 0;
        for (int m = 0; m < ninc; m++) {
            if (used[m]) continue;
            for (int x = 0; x < 3; x++) {
                const int32_t g = go[m][x];
                int c = 0;
                for (int m2 = 0; m2 < ninc; m2++)
                    if (!used[m2]) c += (go[m2][0] == g) + (go[m2][1] == g) + (go[m2][2] == g);
                if (c > best || (c == best && g < J)) { best = c; J = g; }
            }
        }
        // the fan around (I, J)
        int mem[RING_KMAX], nm = 0;
        int8_t lJ[RING_KMAX], la[RING_KMAX], lb[RING_KMAX];
        int32_t wa[RING_KMAX], wb[RING_KMAX];
        bool done[RING_KMAX];
        for (int m = 0; m < ninc; m++) {
            if (used[m]) continue;
            int x = -1;
            for (int t = 0; t < 3; t++)
                if (go[m][t] == J) x = t;
            if (x < 0) continue;
            const int y = x == 0 ? 1 : 0, z = x == 2 ? 1 : 2;
            mem[nm] = m; lJ[nm] = lo[m][x]; la[nm] = lo[m][y]; lb[nm] = lo[m][z]; wa[nm] = go[m][y]; wb[nm] = go[m][z];
            done[nm] = false;
            nm++;
        }
        auto count = [&](int32_t w) { int c = 0; for (int k = 0; k < nm; k++) c += (wa[k] == w) + (wb[k] == w); return c; };
        bool ok = true;
        for (int k = 0; k < nm && ok; k++) {
            if (wa[k] == wb[k] || wa[k] == J || wb[k] == J || count(wa[k]) > 2 || count(wb[k]) > 2) ok = false;
            for (int k2 = k + 1; k2 < nm && ok; k2++)
                if ((wa[k] == wa[k2] && wb[k] == wb[k2]) || (wa[k] == wb[k2] && wb[k] == wa[k2])) ok = false;
        }
        auto emit = [&](int k, int l_in, int l_out, uint32_t mode) {
            const int m = mem[k];
            order[n_out] = m;
            rperm[n_out] = (uint32_t)l0[m] | ((uint32_t)lJ[k] << 2) | ((uint32_t)l_in << 4) | ((uint32_t)l_out << 6);
            rflag[n_out] = mode;
            n_out++;
            used[m] = true;
            done[k] = true;
        };
        if (!ok) {
            for (int k = 0; k < nm; k++) emit(k, la[k], lb[k], 1u);
            continue;
        }
        int placed = 0;
        while (placed < nm) {
            // start at a tetrahedron with a face no other member shares (open fan), else anywhere (closed ring)
            int start = -1, low = -1;
            bool in_is_a = true;
            for (int k = 0; k < nm && start < 0; k++) {
                if (done[k]) continue;
                if (low < 0) low = k;
                int ca = 0, cb = 0;
                for (int k2 = 0; k2 < nm; k2++)
                    if (!done[k2]) { ca += (wa[k2] == wa[k]) + (wb[k2] == wa[k]); cb += (wa[k2] == wb[k]) + (wb[k2] == wb[k]); }
                if (ca == 1) { start = k; in_is_a = true; }
                else if (cb == 1) { start = k; in_is_a = false; }
            }
            if (start < 0) { start = low; in_is_a = true; }
            int cur = start;
            int32_t w_in = in_is_a ? wa[cur] : wb[cur];
            for (;;) {
                const bool cin_a = wa[cur] == w_in;
                const int l_in = cin_a ? la[cur] : lb[cur], l_out = cin_a ? lb[cur] : la[cur];
                const int32_t w_out = cin_a ? wb[cur] : wa[cur];
                done[cur] = true;
                int nxt = -1;
                for (int k = 0; k < nm && nxt < 0; k++)
                    if (!done[k] && (wa[k] == w_out || wb[k] == w_out)) nxt = k;
                emit(cur, l_in, l_out, nxt >= 0 ? 0u : 1u);
                placed++;
                if (nxt < 0) break;
                cur = nxt;
                w_in = w_out;
            }
        }
    }
}

// rtype[row] = 0 vertex-node row / 1 edge-node row in ring order (3D P2) / 2 edge-node row, generic;
// sig[row] = hash of the row's stencil shape (row length, permutations, flags and positions of all
// incidences, NOT the element indices): rows with equal signatures address their accumulators identically
// and are grouped into the same warps (bank-conflict-free shared-memory accesses).
template <int DIM, int NLR, int NL>
__global__ void k_make_records(int64_t n_rows, const int64_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                               const int64_t *__restrict__ rowptr, const uint16_t *__restrict__ pos, int pos_stride,
                               const int32_t *__restrict__ conn, int use_ring,
                               uint32_t *__restrict__ rec, int8_t *__restrict__ rtype, uint64_t *__restrict__ sig)
{
    // NLR / NL: local nodes of the row / column space (equal for the square operators)
    constexpr int RW = RecWords<NL>::value;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kb = inc_ptr[r];
        const int ninc = (int)(inc_ptr[r + 1] - kb);
        const int ty = (ninc > 0 && (inc[kb] & 15) > DIM) ? 1 : 0;
        uint64_t h = 1469598103934665603ull;
        h = sig_mix(h, (uint32_t)(rowptr[r + 1] - rowptr[r]));
        bool ring = false, vchain = false;   // edge row in ring order / vertex row in chain order
        uint64_t seen[4] = {0, 0, 0, 0}; // positions that receive a local contribution (rows of up to 256 nodes)
        int order[RING_KMAX];
        uint32_t rperm[RING_KMAX], rflag[RING_KMAX];
        if constexpr (DIM == 3 && NL == 10 && NLR == 10) {
            if (use_ring && ty == 1 && ninc <= RING_KMAX) ring = ring_order(ninc, inc + kb, conn, order, rperm, rflag);
            if (ty == 0 && ninc > 0 && ninc <= RING_KMAX) { vchain_order(ninc, inc + kb, conn, order, rperm, rflag); vchain = true; }
        }
        for (int m = 0; m < ninc; m++) {
            const int32_t code = inc[kb + ((ring || vchain) ? order[m] : m)];
            const int64_t e = code >> 4;
            const int i = code & 15;
            int pi[DIM + 1];
            uint32_t bits = 0;
            if (ring || vchain) {
                bits = rperm[m];
                for (int v = 0; v <= DIM; v++) pi[v] = (int)((bits >> (2 * v)) & 3);
                bits |= rflag[m] << 8;
            } else {
                canon_perm<DIM>(i, pi);
                for (int v = 0; v <= DIM; v++) bits |= (uint32_t)pi[v] << (2 * v);
                if (DIM == 3 && NL == 10 && NLR == 10 && ty == 0) bits |= 1u << 8; // vertex row, not ordered: chains of length one
            }
            uint32_t w[RW];
            for (int x = 0; x < RW; x++) w[x] = 0;
            uint64_t natidx = 0;
            for (int jc = 0; jc < NL; jc++) {
                const int jn = canon_node<DIM>(pi, jc);
                natidx |= (uint64_t)jn << (4 * jc);
                const uint32_t p = pos[(e * NLR + i) * pos_stride + jn];
                w[jc >> 1] |= p << (16 * (jc & 1));
                if (p < 256) seen[p >> 6] |= uint64_t(1) << (p & 63);
            }
            h = sig_mix(h, bits);
            for (int x = 0; x < (NL + 1) / 2; x++) h = sig_mix(h, w[x]);
            rec_pack_code<NL>(w, (uint32_t)e, bits, natidx);
            for (int x = 0; x < RW; x++) rec[(kb + m) * RW + x] = w[x];
        }
        // bit 4: some position of the row gets no local contribution (entries contributed by other ranks only,
        // or an empty row): the write-once ring kernel must zero its shared-memory row first
        const int len = (int)(rowptr[r + 1] - rowptr[r]);
        const int covered = __popcll(seen[0]) + __popcll(seen[1]) + __popcll(seen[2]) + __popcll(seen[3]);
        const int holes = (len > 256 || covered < len) ? 16 : 0;
        rtype[r] = (int8_t)((ty == 0 ? 0 : (ring ? 1 : 2)) | holes);
        sig[r] = h;
    }
}

// Canonical coefficient table: K_{i',j'} = sum_{s,t} r[type][j'][s][t] * E(s, sv(j',t)) where
// s runs over the canonical support of the row function (vertex 0 | edge (0,1)) and t over the
// support of column function j' (a vertex function uses t = 0 only).
struct CanonR {
    double r[2][MAXN][2][2];
};

// per-row record of the gather path (bucket order): one aligned 32-byte load per thread
struct RowInfo {
    int64_t base;   // rowptr[row]   (node level)
    int64_t k0;     // first incidence
    int32_t len;    // node-row length
    int32_t ninc;   // number of incidences
    int64_t pad;    // bits 0-1 row type, bit 4 the row has positions without a local contribution, bit 5 the row's tail goes
                    // straight to the values array (last row of the fragment protocol), bits 32-63 the row (index in the pattern)
    uint32_t e[8];  // elements of the row's first incidences -- lets k_gather_s request the geometry lines of a
                    // row's first incidences from the row record alone (second 32-byte half)
};

constexpr int kMaxGhostSeg = 8;   // ranks of one box

struct GatherArgs {
    const RowInfo *rowinfo;   // rows of this launch: rowinfo[start .. start+count), bucket order
    int64_t start, count;
    const uint32_t *rec;      // incidence records
    const double *geom;
    double c0, c1;            // lambda, mu
    double *values;
    int pitch;                // doubles between consecutive shared-memory accumulator rows (== NBL mod 16)
    int zero;                 // always 0 (opaque to the compiler; used to order loads after their buffer's last use)
    int vec_dim;              // LAP: 0 scalar, DIM = replicate to the DIM block-diagonal dof rows
    // ghost-row launches of a multi-GPU assembly (feddb200_set_ghost_targets): values at offsets
    // [seg_begin[o], seg_begin[o+1]) are written to seg_ptr[o] + (offset - seg_begin[o]) -- the receive buffer of the
    // row's owner, mapped over NVLink -- instead of values + offset.  nseg == 0: everything goes to `values`.
    int nseg;
    int64_t seg_begin[kMaxGhostSeg + 1];
    double *seg_ptr[kMaxGhostSeg];
    const uint32_t *ahead;    // k_gather_s: element of the incidence kRsAhead places further along the same row
    double *frag;             // fragment protocol (below): two 32-byte slots per node row; null: plain stores
    // points path of the scalar Laplace rows (3D P2): canonical vertex ids of every incidence + padded coordinates; the
    // geometry is recomputed per incidence instead of read as a 128-byte line (see geo_from_points)
    const uint4 *vtx;         // [n_inc] canonical vertices (node ids of the geometry mesh)
    const double *coords4;    // [nn][4] (x, y, z, 0)
    CanonR R;
};

// Fragment protocol of the write-out.  The values of a row node (all its dof rows) are one contiguous RUN of the values
// array, 8-byte aligned, so its first and last 32-byte sector are shared with the neighbouring runs -- rows of another type,
// written by another launch milliseconds later.  A sector that reaches DRAM partially written costs a read-modify-write
// there: measured on B200 (tools/microbench_frag.cu, microbench_chunk.cu) runs of 1944 bytes written in an order in which
// neighbours are not written together reach 3.6 TB/s, the same runs 32-byte aligned 5.4 TB/s, in address order 5.7 TB/s.
// So every launch writes only the WHOLE sectors of its runs; the doubles in front of the first / behind the last sector
// boundary go, as whole 32-byte stores, into the run's head / tail slot of a side buffer (frag[2 row], frag[2 row + 1],
// values at their position inside the sector), and k_stitch composes every shared sector from the tail slot of the left
// run and the head slot of the right run after the last launch (4.5 TB/s for the three passes together).
// Rows flagged with bit 5 of RowInfo::pad (the last row of the protocol: nothing behind it completes the sector) store
// their tail doubles directly.  Runs are at least 8 doubles long (checked when the maps are built).
__device__ __forceinline__ void st_v4(double *p, double a, double b, double c, double d);
struct RunSplit {
    int hc, tc;      // doubles in front of the first / behind the last sector boundary
};
__device__ __forceinline__ RunSplit run_split(const double *outp, int total)
{
    const int mis = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 3);
    RunSplit s;
    s.hc = (4 - mis) & 3;
    s.tc = (mis + total) & 3;
    return s;
}
// head / tail doubles of the run [src, src + total) (shared or generic memory) into the slots of `row`
__device__ __forceinline__ void frag_head(double *frag, int64_t row, const double *src, int hc)
{
    const int mis = 4 - hc;   // hc in 1..3
    st_v4(frag + 8 * row, 0.0, mis <= 1 ? src[1 - mis] : 0.0, mis <= 2 ? src[2 - mis] : 0.0, src[3 - mis]);
}
__device__ __forceinline__ void frag_tail(double *frag, int64_t row, const double *src_tail, int tc)
{
    st_v4(frag + 8 * row + 4, src_tail[0], tc > 1 ? src_tail[1] : 0.0, tc > 2 ? src_tail[2] : 0.0, 0.0);
}
// one thread per boundary between the runs of rows r and r + 1 (r + 1 < n_rows_protocol)
__global__ void __launch_bounds__(256) k_stitch(int64_t n_bound, const int64_t *__restrict__ rowptr, int mult, double *__restrict__ values,
                                                const double *__restrict__ frag)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_bound; r += (int64_t)gridDim.x * blockDim.x) {
        double *e = values + (int64_t)mult * rowptr[r + 1];
        const int t = (int)((reinterpret_cast<uintptr_t>(e) >> 3) & 3);
        if (t == 0) continue;
        const double4 a = *reinterpret_cast<const double4 *>(frag + 8 * r + 4), b = *reinterpret_cast<const double4 *>(frag + 8 * (r + 1));
        st_v4(e - t, a.x, t > 1 ? a.y : b.y, t > 2 ? a.z : b.z, b.w);
    }
}

// destination of the values at offset `off` of the values array (see GatherArgs::nseg)
__device__ __forceinline__ double *out_ptr(const GatherArgs &A, int64_t off)
{
    if (A.nseg == 0) return A.values + off;
    int o = 0;
#pragma unroll 1
    while (o + 1 < A.nseg && off >= A.seg_begin[o + 1]) o++;
    return A.seg_ptr[o] + (off - A.seg_begin[o]);
}

// TMA bulk store shared -> global (1-D, no tensor map): both addresses 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_store(void *gptr, const void *sptr, int bytes)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sptr);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr), "r"(s), "r"(bytes) : "memory");
}
// L2 eviction policies (tuning aids FB_HINT_ST / FB_HINT_GEOM): the assembled values are written once and never read
// again by the assembly (evict-first), the geometry lines are re-read by every row that touches the element (evict-last)
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_store_hint(void *gptr, const void *sptr, int bytes, uint64_t pol)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sptr);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gptr), "r"(s), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_out(double *p, double v)
{
#ifdef FB_HINT_ST
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy_evict_first()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void bulk_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifdef FB_GATHERW_PF_L1
#define FB_PF_ELEM prefetch_l1
#else
#define FB_PF_ELEM prefetch_l2
#endif
// prefetch.global.L2 brings in ONE 32-byte sector (ncu: 12 sectors per warp-wide prefetch of ten 128-byte lines); whole
// lines / spans go through the bulk (TMA) prefetch: address 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ld_v4(const double *p, double (&v)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// geometry lines (re-read by every incident row)
__device__ __forceinline__ void ld_v4g(const double *p, double (&v)[4])
{
#ifdef FB_HINT_GEOM
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p), "l"(policy_evict_last()));
#else
    ld_v4(p, v);
#endif
}
__device__ __forceinline__ void ld_v8u(const uint32_t *p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

// support (canonical vertices) of canonical local node jc
template <int DIM>
__device__ __forceinline__ constexpr int canon_sv(int jc, int t)
{
    constexpr int NVTX = DIM + 1;
    if (jc < NVTX) return jc;
    if constexpr (DIM == 2) { constexpr int E[3][2] = {{0, 1}, {1, 2}, {0, 2}}; return E[jc - NVTX][t]; }
    else { constexpr int E[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}}; return E[jc - NVTX][t]; }
}

template <int NL>
struct IncRec {
    uint32_t w[RecWords<NL>::value];
};
template <int DIM>
struct IncGeo {
    double G[DIM + 1][4];     // canonical vertex v: (Gx, Gy[, Gz], |det|) -- 2D uses [0],[1] and [3]
};

template <int NL>
__device__ __forceinline__ void load_rec(const GatherArgs &A, int64_t k, IncRec<NL> &R)
{
    constexpr int RW = RecWords<NL>::value;
    const uint32_t *p = A.rec + k * RW;
    if constexpr (RW == 8) ld_v8u(p, R.w);
    else {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        R.w[0] = a.x; R.w[1] = a.y; R.w[2] = a.z; R.w[3] = a.w;
    }
}

template <int DIM, int NL>
__device__ __forceinline__ void load_geo(const GatherArgs &A, const IncRec<NL> &R, IncGeo<DIM> &D)
{
    constexpr int GS = GeomStride<DIM>::value;
    const uint32_t perm = rec_perm<NL>(R.w);
    const double *g = A.geom + (int64_t)rec_elem<NL>(R.w) * GS;
    if constexpr (DIM == 3) {
#pragma unroll
        for (int v = 0; v < 4; v++) ld_v4g(g + 4 * ((perm >> (2 * v)) & 3), D.G[v]);
    } else {
        const double ad = __ldg(g + 6);
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(g) + ((perm >> (2 * v)) & 3));
            D.G[v][0] = t.x; D.G[v][1] = t.y; D.G[v][2] = 0.0; D.G[v][3] = ad;
        }
    }
}

// Points path (scalar Laplace rows of 3D P2 meshes).  A Laplace row moves 160 bytes of input (32-byte record + 128-byte
// geometry line, re-read from DRAM about 1.7 times per element because the output stream evicts it from L2) for every
// 80 bytes it writes; the launches run at the DRAM traffic ceiling of this access pattern (ncu: 2.3 GB read + 0.9 GB
// written in 1.02 ms for the largest bucket of config 2).  Here an incidence carries the node ids of its four canonical
// vertices (16 bytes, k_make_vtx) and the kernel recomputes (grad lambda_v, |det|) from the padded coordinates, which
// stay in L2 (33 MB for a million vertices): 48 instead of 160 bytes of DRAM reads per incidence, no k_geom pass, ~55 more
// FP64 instructions per incidence on a pipe that was 8 % busy.  The gradients are those of k_geom up to rounding (the
// origin of the affine map is the canonical instead of the natural first vertex).
__global__ void k_make_vtx(int64_t n_inc, const uint32_t *__restrict__ rec, const int32_t *__restrict__ conn, uint4 *__restrict__ vtx)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_inc; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t e = rec[k * 8 + 5], perm = rec[k * 8 + 6];
        const int32_t *el = conn + (int64_t)e * 10;
        vtx[k] = make_uint4((uint32_t)el[perm & 3], (uint32_t)el[(perm >> 2) & 3], (uint32_t)el[(perm >> 4) & 3], (uint32_t)el[(perm >> 6) & 3]);
    }
}
__global__ void k_coords4(int64_t nn, const double *__restrict__ coords, double *__restrict__ coords4)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x)
        st_v4(coords4 + 4 * n, coords[3 * n], coords[3 * n + 1], coords[3 * n + 2], 0.0);
}
// issue the four coordinate loads of an incidence into the geometry buffer (converted in place by geo_from_points)
__device__ __forceinline__ void load_points(const GatherArgs &A, const uint4 v, IncGeo<3> &D, int dep = 0)
{
    ld_v4g(A.coords4 + 4 * (int64_t)v.x + dep, D.G[0]);
    ld_v4g(A.coords4 + 4 * (int64_t)v.y + dep, D.G[1]);
    ld_v4g(A.coords4 + 4 * (int64_t)v.z + dep, D.G[2]);
    ld_v4g(A.coords4 + 4 * (int64_t)v.w + dep, D.G[3]);
}
__device__ __forceinline__ void geo_from_points(IncGeo<3> &D)
{
    double B[3][3];   // B[i][j] = x_{j+1}[i] - x_0[i]
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) B[i][j] = D.G[j + 1][i] - D.G[0][i];
    const double det = B[0][0] * B[1][1] * B[2][2] + B[0][1] * B[1][2] * B[2][0] + B[0][2] * B[1][0] * B[2][1] -
                       B[2][0] * B[1][1] * B[0][2] - B[2][1] * B[1][2] * B[0][0] - B[2][2] * B[1][0] * B[0][1];
    const double r = 1.0 / det, ad = fabs(det);
    // grad lambda_k = row k-1 of B^-1 (k >= 1), grad lambda_0 = -(sum of the others)
    D.G[1][0] = (B[1][1] * B[2][2] - B[1][2] * B[2][1]) * r;
    D.G[1][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * r;
    D.G[1][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * r;
    D.G[2][0] = (B[1][2] * B[2][0] - B[1][0] * B[2][2]) * r;
    D.G[2][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * r;
    D.G[2][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * r;
    D.G[3][0] = (B[1][0] * B[2][1] - B[1][1] * B[2][0]) * r;
    D.G[3][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * r;
    D.G[3][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * r;
#pragma unroll
    for (int d = 0; d < 3; d++) D.G[0][d] = -(D.G[1][d] + D.G[2][d] + D.G[3][d]);
#pragma unroll
    for (int v = 0; v < 4; v++) D.G[v][3] = ad;
}

// Row-gather kernel (output-stationary): every CSR value is written exactly once, no atomics, no memset,
// bitwise reproducible.
//
// One thread per scalar COMPONENT row: OPG 0 (Laplace) one thread per row node; OPG 1 (elasticity) DIM*DIM
// threads per row node, thread (I, a, b) owns the values K^{ab}_{I,J} of every column node J.  The thread
// walks the elements incident to its row node (records: positions | element | canonical permutation),
// loads the element's geometry line in canonical vertex order (the row node is canonical vertex 0 or
// canonical edge (0,1), so the code is fully static), forms
//     e[s][w] = |det| G_s^T M G_w,   M = mu (delta_ab I + e_b e_a^T) + lambda e_a e_b^T   (Laplace: M = I)
// for the row support s and the 4 (3) canonical vertices w, and adds  sum_{s,t} R_j^{st} e[s][sv(j,t)]  to
// its accumulator of column node j (R: the operator's own quadrature applied to the P2 gradient
// coefficients, see canon_table in api.cu).
//
// Accumulators live in shared memory laid out exactly like the CSR rows: one row of NBL*L doubles per dof
// row (I, a), `pitch` doubles apart, thread (I, a, b) touching entries NBL*p + b.  With pitch == NBL (mod 16)
// the 16 threads of a half-warp that address the same column position hit 16 distinct 8-byte banks, and a
// node's own threads never collide, so the data-dependent read-modify-write is (nearly) conflict-free on
// any mesh.  The write-out is then a plain block-cooperative copy with full-line coalesced stores.
// Small per-thread state (one accumulator row per DIM threads, ~110 registers) keeps ~16 warps resident
// per SM, which is what hides the latency of the dependent record -> geometry -> accumulate chain.
#ifndef FB_GATHER_MINBLOCKS
#define FB_GATHER_MINBLOCKS 5
#endif
template <int OPG, int DIM> struct GatherShape {
    static constexpr int NBL = OPG == 1 ? DIM : 1;        // values per column node in a shared-memory row
    static constexpr int CPR = OPG == 1 ? DIM * DIM : 1;  // threads per row node
    static constexpr int NT = 32 * NBL;                   // threads per block: 32 accumulator rows
};

template <int OPG, int DIM, int NL, int TYPE, bool PTS = false>
__global__ void __launch_bounds__(GatherShape<OPG, DIM>::NT, FB_GATHER_MINBLOCKS) k_gather(const GatherArgs A)
{
    static_assert(!PTS || (OPG == 0 && DIM == 3), "points path: scalar Laplace rows of 3D meshes");
    using S = GatherShape<OPG, DIM>;
    constexpr int NBL = S::NBL, CPR = S::CPR, NT = S::NT;
    constexpr int NVTX = DIM + 1;
    constexpr int NS = TYPE == 0 ? 1 : 2;    // canonical support size of the row function
    constexpr int JD = TYPE == 0 ? 0 : NVTX; // canonical index of the row node itself
    constexpr int ROWS = 32;                 // accumulator rows per block
    extern __shared__ double acc[];          // [ROWS][pitch]
    __shared__ int64_t s_off[ROWS];
    __shared__ int64_t s_flags[ROWS];
    __shared__ int s_n[ROWS];
    const int tid = threadIdx.x;
    const int pitch = A.pitch;
    const int64_t t = blockIdx.x * (int64_t)NT + tid;
    const bool live = t < A.count * CPR;
    const int64_t rloc = live ? t / CPR : 0;
    const int comp = live ? (int)(t - rloc * CPR) : 0;
    const int a = comp / NBL, b = comp - a * NBL;
    int64_t base = 0, k0 = 0, rflags = 0;
    int L = 0, ninc = 0;
    if (live) {
        double raw[4];
        ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + rloc), raw);
        base = __double_as_longlong(raw[0]);
        k0 = __double_as_longlong(raw[1]);
        const int64_t ln = __double_as_longlong(raw[2]);
        L = (int)(ln & 0xffffffff);
        ninc = (int)(ln >> 32);
        rflags = __double_as_longlong(raw[3]);
    }
    const int64_t k1 = k0 + ninc, kl = k1 - 1;
    IncRec<NL> rc, rn, r2;
    IncGeo<DIM> g;
    uint4 vn = make_uint4(0u, 0u, 0u, 0u);   // points path: canonical vertices of the next incidence
    if (ninc > 0) {
        load_rec<NL>(A, k0, rc);
        load_rec<NL>(A, k0 + 1 < kl ? k0 + 1 : kl, rn);
        load_rec<NL>(A, k0 + 2 < kl ? k0 + 2 : kl, r2);
        if constexpr (PTS) {
            const uint4 vc = __ldg(A.vtx + k0);
            vn = __ldg(A.vtx + (k0 + 1 < kl ? k0 + 1 : kl));
            load_points(A, vc, g);
        } else load_geo<DIM, NL>(A, rc, g);
    }
    // zero the block's accumulators while the first loads are in flight; publish the row table
    for (int x = tid; x < ROWS * pitch; x += NT) acc[x] = 0.0;
    const int row = tid / NBL;               // accumulator row of this thread inside the block
    if (b == 0) {
        s_n[row] = live ? NBL * L : 0;
        s_off[row] = OPG == 1 ? (int64_t)CPR * base + (int64_t)a * NBL * L : base;
        s_flags[row] = rflags;
    }
    __syncthreads();
    double *my = acc + (size_t)row * pitch + b;

    if (ninc > 0) {
        // M^T columns, scaled per incidence by |det|
        double Mm[DIM][DIM];
        if constexpr (OPG == 1) {
#pragma unroll
            for (int c = 0; c < DIM; c++)
#pragma unroll
                for (int d = 0; d < DIM; d++)
                    Mm[c][d] = ((a == b && c == d) ? A.c1 : 0.0) + ((c == b && d == a) ? A.c1 : 0.0) + ((c == a && d == b) ? A.c0 : 0.0);
        }
        double dacc = 0.0;
        const int pdiag = (int)((rc.w[JD >> 1] >> (16 * (JD & 1))) & 0xffffu) * NBL;
        for (int64_t k = k0; k < k1; k++) {
            if constexpr (PTS) geo_from_points(g);
            // e[s][w] from the current geometry buffer
            double e[NS][NVTX];
            const double adet = g.G[0][3];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                double v[DIM];
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                    if constexpr (OPG == 1) {
                        double x = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; c++) x += g.G[s][c] * Mm[c][d];
                        v[d] = x * adet;
                    } else v[d] = g.G[s][d] * adet;
                }
#pragma unroll
                for (int w = 0; w < NVTX; w++) {
                    double x = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) x += v[d] * g.G[w][d];
                    e[s][w] = x;
                }
            }
            // the geometry buffer is consumed: refill it for the next incidence (the address is made to depend
            // on e so the loads cannot be scheduled above the computation that waits for the previous ones),
            // and fetch the record three incidences ahead.  Indices are clamped to the row's last incidence.
            IncRec<NL> r3;
            {
                const int dep = __double2hiint(e[0][0]) & A.zero;
                GatherArgs const &AA = A;
                const uint32_t perm = rec_perm<NL>(rn.w);
                const double *gp = AA.geom + (int64_t)rec_elem<NL>(rn.w) * GeomStride<DIM>::value + dep;
                if constexpr (PTS) {
                    load_points(A, vn, g, dep);
                    vn = __ldg(A.vtx + (k + 2 < kl ? k + 2 : kl));
                } else
                if constexpr (DIM == 3) {
#pragma unroll
                    for (int v = 0; v < 4; v++) ld_v4g(gp + 4 * ((perm >> (2 * v)) & 3), g.G[v]);
                } else {
                    const double ad = __ldg(gp + 6);
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        const double2 tt = __ldg(reinterpret_cast<const double2 *>(gp) + ((perm >> (2 * v)) & 3));
                        g.G[v][0] = tt.x; g.G[v][1] = tt.y; g.G[v][2] = 0.0; g.G[v][3] = ad;
                    }
                }
                load_rec<NL>(A, k + 3 < kl ? k + 3 : kl, r3);
            }
            // the row node itself receives a contribution from every incidence: accumulated in a register
#pragma unroll
            for (int s = 0; s < NS; s++) {
                dacc += A.R.r[TYPE][JD][s][0] * e[s][canon_sv<DIM>(JD, 0)];
                if (JD >= NVTX) dacc += A.R.r[TYPE][JD][s][1] * e[s][canon_sv<DIM>(JD, 1)];
            }
            // distinct canonical nodes hit distinct row positions: load all accumulators, add, store all
            constexpr int NO = NL - 1;
            double *ptr[NO];
            double old[NO];
#pragma unroll
            for (int jj = 0; jj < NO; jj++) {
                const int jc = jj + (jj >= JD ? 1 : 0);
                ptr[jj] = my + ((rc.w[jc >> 1] >> (16 * (jc & 1))) & 0xffffu) * NBL;
                old[jj] = *ptr[jj];
            }
#pragma unroll
            for (int jj = 0; jj < NO; jj++) {
                const int jc = jj + (jj >= JD ? 1 : 0);
                double v = old[jj];
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    v += A.R.r[TYPE][jc][s][0] * e[s][canon_sv<DIM>(jc, 0)];
                    if (jc >= NVTX) v += A.R.r[TYPE][jc][s][1] * e[s][canon_sv<DIM>(jc, 1)];
                }
                *ptr[jj] = v;
            }
            rc = rn; rn = r2; r2 = r3;
        }
        my[pdiag] = dacc;
    }
    __syncthreads();

    // write-out: every accumulator row is one CSR row (scalar row / elasticity dof row (I, a)); the
    // block-diagonal vector Laplacian replicates it DIM times.  Warps stride over the block's rows.
    const int nrep = (OPG == 0 && A.vec_dim != 0) ? A.vec_dim : 1;
    const int lane = tid & 31;
    for (int r = tid >> 5; r < ROWS; r += NT / 32) {
        const int nr = s_n[r];
        const double *src = acc + (size_t)r * pitch;
        if (A.frag != nullptr && A.nseg == 0) {
            // fragment protocol: this accumulator row is dof row a = r % NBL of its node; the node's run starts with
            // dof row 0 and ends with dof row NBL - 1 (the boundaries between them are written by this block together)
            if (nr == 0) continue;
            double *out = A.values + s_off[r];
            const int64_t fl = s_flags[r];
            const int a_r = r % NBL;
            const RunSplit sp = run_split(out, nr);
            const int lo = a_r == 0 ? sp.hc : 0;
            const bool tail_frag = a_r == NBL - 1 && sp.tc != 0 && (fl & 32) == 0;
            const int hi = tail_frag ? nr - sp.tc : nr;
#pragma unroll 4
            for (int x = lo + lane; x < hi; x += 32) st_out(out + x, src[x]);
            if (lane == 0 && lo != 0) frag_head(A.frag, fl >> 32, src, sp.hc);
            if (lane == 1 && tail_frag) frag_tail(A.frag, fl >> 32, src + hi, sp.tc);
            continue;
        }
        for (int d = 0; d < nrep; d++) {
            double *out = out_ptr(A, (int64_t)nrep * s_off[r]) + (int64_t)d * nr;
#pragma unroll 4
            for (int x = lane; x < nr; x += 32) st_out(out + x, src[x]);
        }
    }
}

// Ring kernel: 3D P2 edge-node rows whose incidences are in chain order (ring_order above).  One thread per
// CSR row: OPG 0 (Laplace) row node, OPG 1 (elasticity) dof row (I, a) with NB = 3 values per column node.
// With canonical vertices (v0, v1, w_in, w_out) the 10 column nodes of an incidence split into
//   {0, 1, 4}  v0, v1 and the row node itself: every incidence contributes  -> register accumulators
//   {2, 6, 5}  nodes of the in-face  (v0, v1, w_in):  carry + contribution is final -> stored once
//   {3, 7, 8}  nodes of the out-face (v0, v1, w_out): becomes the carry of the next incidence
//   {9}        node (w_in, w_out), this tetrahedron only -> stored once
// so apart from the closing face of a ring there is no shared-memory read at all.  The thread's shared-memory
// row is laid out exactly like its CSR row (`pitch` odd: conflict-free when the threads of a half-warp store
// to the same position) and is copied out by the warp with full-line coalesced stores.
#ifndef FB_RING_MINBLOCKS
#define FB_RING_MINBLOCKS 5
#endif

#ifndef FB_RING_UNROLL
#define FB_RING_UNROLL 1
#endif
#ifndef FB_RING_PF_AT
#define FB_RING_PF_AT 1      // main-loop iteration after which the next tile's lines are requested into L2
#endif
// L2 prefetch of k_ring (tuning aid).  0: one sector of the geometry line two incidences ahead (default), 1: that whole line
// (bulk prefetch), 2: everything the next tile reads, a tile ahead (bulk prefetches), 3: 1 + 2, 4: all four sectors of that line
// by plain prefetches (config 3 2.484 ms against 2.465, config 2 3.24 against 3.10 ms), 5: none (2.60 / 3.07 ms), 6: sectors
// 0 and 2 (2.48 / 3.16 ms) -- one sector two incidences ahead is the optimum inside a row.  Measured on B200, config 3:
// 2.49 / 2.71 / 2.65 / 2.91 ms per assembly -- every additional prefetched byte makes the step SLOWER: the ring launches move
// 3.5-3.8 TB/s of DRAM traffic (writes + reads), the ceiling of this write pattern (tools/microbench_window.cu), so they are
// bound by DRAM traffic, not by the latency their long-scoreboard stalls suggest.
// 7: mode 0 + the incidence records of the NEXT tile's rows by plain prefetches from inside the main loop: config 3
// 2.465 -> 2.379 ms, config 2 3.10 -> 3.02 ms (same box).  8 (default): 7 + the first geometry sector of the next tile's
// first two incidences: config 3 2.470 (mode 0) / 2.387 (7) / 2.346 ms (8) on one box.  What helps is taking the DRAM misses
// out of the dependent chain row record -> incidence records -> geometry at every tile start -- with few, plain requests.
#ifndef FB_RING_PF_MODE
#define FB_RING_PF_MODE 8
#endif
constexpr int kRingUnroll = FB_RING_UNROLL;
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc);
__device__ __forceinline__ void cp_async_commit();
template <int N> __device__ __forceinline__ void cp_async_wait();
template <int OPG, bool PTS = false>
__device__ __forceinline__ void ring_tiles(const GatherArgs &A, double *acc)
{
    static_assert(!PTS || OPG == 0, "points path: scalar Laplace rows");
#if FB_RING_PF_MODE == 2 || FB_RING_PF_MODE == 3
    __shared__ uint4 s_en[2][64];   // elements of the next tile's rows (second half of their row records), per lane
#endif
    constexpr int DIM = 3, NL = 10, NVTX = 4;
    constexpr int TPR = OPG == 1 ? DIM : 1;
    constexpr int NB = OPG == 1 ? DIM : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int pitch = A.pitch;
    // a tile = the NPT row nodes of ONE WARP; lane < NPT * TPR owns dof a = lane % TPR of the tile's node lane / TPR, so
    // the TPR dof rows of a node are always in the same warp (they are contiguous in the values array and leave in ONE
    // bulk store) and warps never wait for one another
#ifndef FB_RING_NPT3
#define FB_RING_NPT3 10     // row nodes per warp tile of the elasticity ring rows (3 lanes each); must be even (16-byte aligned
                            // warp areas).  8 nodes = 12 instead of 10 resident warps at L = 27, 24 active lanes: measured 2.51 vs 2.47 ms
#endif
    constexpr int NPT = TPR == 3 ? FB_RING_NPT3 : 32 / TPR;
    const int slot = lane / TPR;
    const bool lane_used = lane < NPT * TPR;
    const int64_t ntiles = (A.count + NPT - 1) / NPT;
    acc += (size_t)(tid >> 5) * NPT * A.pitch;        // this warp's shared-memory rows
    const int64_t tile_step = (int64_t)gridDim.x * (NT >> 5);
    const int nrep = (OPG == 0 && A.vec_dim != 0) ? A.vec_dim : 1;
    const double mu = A.c1, lam = A.c0;
    double *wbase = acc + (size_t)(tid - lane) * pitch; // the warp's 32 rows
    auto posof = [&](const IncRec<NL> &r, int jc) { return (int)((r.w[jc >> 1] >> (16 * (jc & 1))) & 0xffffu) * NB; };

    // Persistent blocks: a block walks tiles of blockDim.x rows.  The row record of the NEXT tile is fetched at
    // the start of the current one, its first incidence records after the current main loop and its first
    // geometry line while the bulk stores of the current tile drain, so the dependent chain
    // row record -> incidence record -> geometry is off the critical path for all but a block's first tile.
    int64_t tile = (int64_t)blockIdx.x * (NT >> 5) + (tid >> 5);
    if (tile >= ntiles) return;
    double raw[4] = {0.0, 0.0, 0.0, 0.0};
    IncRec<NL> rc, rn, r2;
    IncGeo<DIM> g;
    uint4 vc = make_uint4(0u, 0u, 0u, 0u), vn = vc;   // points path: canonical vertices of this / the next incidence
    {
        const int64_t node = tile * NPT + slot;
        if (lane_used && node < A.count) {
            ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + node), raw);
            const int64_t k0 = __double_as_longlong(raw[1]);
            const int ninc = (int)(__double_as_longlong(raw[2]) >> 32);
            if (ninc > 0) {
                const int64_t kl = k0 + ninc - 1;
                load_rec<NL>(A, k0, rc);
                load_rec<NL>(A, k0 + 1 < kl ? k0 + 1 : kl, rn);
                load_rec<NL>(A, k0 + 2 < kl ? k0 + 2 : kl, r2);
                if constexpr (PTS) {
                    vc = __ldg(A.vtx + k0);
                    vn = __ldg(A.vtx + (k0 + 1 < kl ? k0 + 1 : kl));
                    load_points(A, vc, g);
                } else load_geo<DIM, NL>(A, rc, g);
            }
        }
    }
    for (;;) {
        const bool live = lane_used && tile * NPT + slot < A.count;
        const int a = lane - slot * TPR;
#ifdef FB_RING_UNITVEC_DOF
        const double ea[3] = {a == 0 ? 1.0 : 0.0, a == 1 ? 1.0 : 0.0, a == 2 ? 1.0 : 0.0};   // unit vector of the row dof
#endif
        const int64_t base = __double_as_longlong(raw[0]);
        const int64_t k0 = __double_as_longlong(raw[1]);
        const int L = live ? (int)(__double_as_longlong(raw[2]) & 0xffffffff) : 0;
        const int ninc = live ? (int)(__double_as_longlong(raw[2]) >> 32) : 0;
        const bool holes = live && (__double_as_longlong(raw[3]) & 16) != 0;
        // row record of the next tile
        const int64_t tile_n = tile + tile_step;
        const int64_t node_n = tile_n * NPT + slot;
        const bool live_n = lane_used && tile_n < ntiles && node_n < A.count;
        double rawn[4] = {0.0, 0.0, 0.0, 0.0};
        if (live_n) ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + node_n), rawn);
#if FB_RING_PF_MODE == 8
        uint2 en01 = make_uint2(0u, 0u);   // elements of the next row's first two incidences
        if (live_n) en01 = __ldg(reinterpret_cast<const uint2 *>(A.rowinfo + A.start + node_n) + 4);
#endif
#if FB_RING_PF_MODE == 2 || FB_RING_PF_MODE == 3
        // ... and the elements of its rows (second half of the row record): their geometry lines, the rows' incidence
        // records and the row records of the tile after the next are requested into L2 a whole tile ahead (do_prefetch
        // below, issued by the node's first lane from inside the main loop, once this load has landed).  These lines are
        // streamed (no reuse): without the request every tile waits for DRAM three times in a row -- row record, first
        // incidence records, first geometry line (ncu source page: 10 %, 12 % and 31 % of the kernel's stall samples).
        // (the element list waits in shared memory, not in registers: the main loop has none to spare)
        const bool pf_lane = live_n && lane == slot * TPR;
        if (pf_lane) {
            const char *src = reinterpret_cast<const char *>(A.rowinfo + A.start + node_n) + 32;
            cp_async16(&s_en[0][tid], src);
            cp_async16(&s_en[1][tid], src + 16);
        }
        cp_async_commit();
        bool pf_done = !pf_lane;
        auto do_prefetch = [&]() {
            const int64_t k0p = __double_as_longlong(rawn[1]);
            const int nincp = (int)(__double_as_longlong(rawn[2]) >> 32);
            cp_async_wait<0>();
            const uint4 e0 = s_en[0][tid], e1 = s_en[1][tid];
            const uint32_t en[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < nincp) prefetch_l2_bulk(A.geom + (int64_t)en[j] * GeomStride<DIM>::value, GeomStride<DIM>::value * 8);
            if (nincp > 0) prefetch_l2_bulk(A.rec + k0p * RecWords<NL>::value, (uint32_t)nincp * RecWords<NL>::value * 4);
            const int64_t node_nn = node_n + tile_step * NPT;
            if (node_nn < A.count) prefetch_l2_bulk(A.rowinfo + A.start + node_nn, (uint32_t)sizeof(RowInfo));
            pf_done = true;
        };
#endif
        const int n = NB * L;                         // values of this thread's dof row
        const int64_t off_node = (int64_t)TPR * NB * nrep * base; // first value of the node's TPR dof rows (contiguous)
        // the node's rows are stored back to back in shared memory (node pitch A.pitch), shifted by one double where
        // that gives them the 16-byte phase of their destination (TMA bulk stores need both sides 16-byte aligned)
        double *const outp = out_ptr(A, off_node);
        const int head = (int)((reinterpret_cast<uintptr_t>(outp) >> 3) & 1);
        double *nodep = acc + (size_t)slot * pitch + ((slot * pitch + head) & 1);
        double *my = nodep + a * n;
        if (__any_sync(FULL, holes)) { // rare: rows with positions no local element contributes to
            for (int x = lane; x < NPT * pitch; x += 32) acc[x] = 0.0;
            __syncwarp();
        }

        if (ninc > 0) {
            const int64_t k1 = k0 + ninc, kl = k1 - 1;
            const int p_v0 = posof(rc, 0), p_v1 = posof(rc, 1), p_self = posof(rc, 4);
            double accE[3][NB], carry[3][NB];
#pragma unroll
            for (int x = 0; x < 3; x++)
#pragma unroll
                for (int b = 0; b < NB; b++) { accE[x][b] = 0.0; carry[x][b] = 0.0; }
            constexpr int JE[3] = {0, 1, 4}, JIN[3] = {2, 6, 5}, JOUT[3] = {3, 7, 8};
#pragma unroll(kRingUnroll)
            for (int64_t k = k0; k < k1; k++) {
                if constexpr (PTS) geo_from_points(g);
                // E[s][w][b] for the two row-support vertices s = v0, v1 and the four canonical vertices w
                double E[2][NVTX][NB];
                {
                    const double adet = g.G[0][3];
                    if constexpr (OPG == 0) {
#pragma unroll
                        for (int s = 0; s < 2; s++)
#pragma unroll
                            for (int w = 0; w < NVTX; w++) {
                                double dot = 0.0;
#pragma unroll
                                for (int d = 0; d < DIM; d++) dot += g.G[s][d] * g.G[w][d];
                                E[s][w][0] = dot * adet;
                            }
                    } else {
                        const double mud = mu * adet, lamd = lam * adet;
                        // the row dof a differs from lane to lane: register selects (42 FSEL + predicated moves per incidence).
                        // FB_RING_UNITVEC_DOF replaces them by products with the unit vector e_a (36 more FP64 instructions, 70
                        // fewer others): measured SLOWER, 2.54 against 2.47 ms -- an FP64 instruction holds the issue port of its
                        // sub-partition for two cycles (16 lanes), so the loop costs 2 x 200 FP64 + 210 other slots and moving
                        // work onto the FP64 pipe is the wrong direction
                        double Ga[NVTX];
#pragma unroll
#ifndef FB_RING_UNITVEC_DOF
                        for (int w = 0; w < NVTX; w++) Ga[w] = a == 0 ? g.G[w][0] : (a == 1 ? g.G[w][1] : g.G[w][2]);
#else
                        for (int w = 0; w < NVTX; w++) Ga[w] = fma(ea[2], g.G[w][2], fma(ea[1], g.G[w][1], ea[0] * g.G[w][0]));
#endif
#pragma unroll
                        for (int s = 0; s < 2; s++) {
                            const double ls = lamd * Ga[s];
#pragma unroll
                            for (int w = 0; w < NVTX; w++) {
                                double dot = 0.0;
#pragma unroll
                                for (int d = 0; d < DIM; d++) dot += g.G[s][d] * g.G[w][d];
                                const double mdot = mud * dot, mga = mud * Ga[w];
                                // E^{ab}_{sw} = mu (delta_ab G_s.G_w + G_s[b] G_w[a]) + lambda G_s[a] G_w[b]
#pragma unroll
#ifndef FB_RING_UNITVEC_DOF
                                for (int b = 0; b < DIM; b++) E[s][w][b] = (a == b ? mdot : 0.0) + mga * g.G[s][b] + ls * g.G[w][b];
#else
                                for (int b = 0; b < DIM; b++) E[s][w][b] = ea[b] * mdot + mga * g.G[s][b] + ls * g.G[w][b];
#endif
                            }
                        }
                    }
                }
                // refill the geometry buffer for the next incidence (ordered after E by a data dependence so the
                // loads are not scheduled above the wait for the previous ones); record three incidences ahead
                IncRec<NL> r3;
                {
                    const int dep = __double2hiint(E[0][0][0]) & A.zero;
                    const uint32_t perm = rec_perm<NL>(rn.w);
#ifdef FB_WHATIF_GEOM0   // timing experiment only (wrong values): every geometry load hits L1
                    const double *gp = A.geom + (int64_t)(rec_elem<NL>(rn.w) & 63) * GeomStride<DIM>::value + dep;
#else
                    const double *gp = A.geom + (int64_t)rec_elem<NL>(rn.w) * GeomStride<DIM>::value + dep;
#endif
                    if constexpr (PTS) {
                        load_points(A, vn, g, dep);
                        vn = __ldg(A.vtx + (k + 2 < kl ? k + 2 : kl));
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; v++) ld_v4g(gp + 4 * ((perm >> (2 * v)) & 3), g.G[v]);
                    }
#ifdef FB_WHATIF_REC0    // timing experiment only (wrong values): every record load hits L1
                    load_rec<NL>(A, k0 + ((k + 3 - k0) & 1), r3);
#else
                    load_rec<NL>(A, k + 3 < kl ? k + 3 : kl, r3);
#endif
#if FB_RING_PF_MODE == 4
                    // all four sectors of the geometry line two incidences ahead, plain (LSU) prefetches: the dof lanes of a node
                    // share them (a = 0: sectors 0 and 3, a = 1: 1, a = 2: 2); a scalar row requests all four itself
                    if constexpr (!PTS) {
                        const double *pg = A.geom + (int64_t)rec_elem<NL>(r2.w) * GeomStride<DIM>::value;
                        if constexpr (TPR == 1) { prefetch_l2(pg); prefetch_l2(pg + 4); prefetch_l2(pg + 8); prefetch_l2(pg + 12); }
                        else { prefetch_l2(pg + 4 * a); if (a == 0) prefetch_l2(pg + 12); }
                    }
#elif FB_RING_PF_MODE == 6
                    // both 64-byte halves of the line (sectors 0 and 2)
                    if constexpr (!PTS) {
                        const double *pg = A.geom + (int64_t)rec_elem<NL>(r2.w) * GeomStride<DIM>::value;
                        if constexpr (TPR == 1) { prefetch_l2(pg); prefetch_l2(pg + 8); }
                        else if (a < 2) prefetch_l2(pg + 8 * a);
                    }
#elif FB_RING_PF_MODE == 0 || FB_RING_PF_MODE == 7 || FB_RING_PF_MODE == 8
                    if constexpr (!PTS) prefetch_l2(A.geom + (int64_t)rec_elem<NL>(r2.w) * GeomStride<DIM>::value);
#elif FB_RING_PF_MODE == 1 || FB_RING_PF_MODE == 3
                    // the whole geometry line of the incidence two places on (one request per node)
                    if (lane == slot * TPR) prefetch_l2_bulk(A.geom + (int64_t)rec_elem<NL>(r2.w) * GeomStride<DIM>::value, GeomStride<DIM>::value * 8);
#endif
                }
                auto contrib = [&](int jc, int b) {
                    double v = 0.0;
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        v += A.R.r[1][jc][s][0] * E[s][canon_sv<DIM>(jc, 0)][b];
                        if (jc >= NVTX) v += A.R.r[1][jc][s][1] * E[s][canon_sv<DIM>(jc, 1)][b];
                    }
                    return v;
                };
                const uint32_t mode = (rec_perm<NL>(rc.w) >> 8) & 3;
#pragma unroll
                for (int x = 0; x < 3; x++)
#pragma unroll
                    for (int b = 0; b < NB; b++) accE[x][b] += contrib(JE[x], b);
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    double *p = my + posof(rc, JIN[x]);
#pragma unroll
                    for (int b = 0; b < NB; b++) p[b] = carry[x][b] + contrib(JIN[x], b);
                }
                {
                    double *p = my + posof(rc, 9);
#pragma unroll
                    for (int b = 0; b < NB; b++) p[b] = contrib(9, b);
                }
#pragma unroll
                for (int x = 0; x < 3; x++)
#pragma unroll
                    for (int b = 0; b < NB; b++) carry[x][b] = contrib(JOUT[x], b);
                if (mode != 0) { // the chain ends here: the out-face is final (1) or closes the ring (2)
#pragma unroll
                    for (int x = 0; x < 3; x++) {
                        double *p = my + posof(rc, JOUT[x]);
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            p[b] = carry[x][b] + (mode == 2 ? p[b] : 0.0);
                            carry[x][b] = 0.0;
                        }
                    }
                }
                rc = rn; rn = r2; r2 = r3;
#if FB_RING_PF_MODE == 2 || FB_RING_PF_MODE == 3
                if (!pf_done && k >= k0 + FB_RING_PF_AT) do_prefetch();
#elif FB_RING_PF_MODE == 7 || FB_RING_PF_MODE == 8
                // the incidence records of the NEXT tile's row (one 32-byte sector each, plain prefetches shared by the dof lanes
                // of the node), once its row record has landed: every tile otherwise starts with a DRAM miss on them
                if (k == k0 + FB_RING_PF_AT && live_n) {
                    const int64_t k0p = __double_as_longlong(rawn[1]);
                    const int nincp = (int)(__double_as_longlong(rawn[2]) >> 32);
                    for (int j = lane - slot * TPR; j < nincp; j += TPR) prefetch_l2(A.rec + (k0p + j) * RecWords<NL>::value);
#if FB_RING_PF_MODE == 8
                    // ... and the first sector of the geometry lines of its first two incidences (RowInfo::e)
                    if (lane - slot * TPR < 2 && lane - slot * TPR < nincp)
                        prefetch_l2(A.geom + (int64_t)(lane - slot * TPR == 0 ? en01.x : en01.y) * GeomStride<DIM>::value);
#endif
                }
#endif
            }
#pragma unroll
            for (int b = 0; b < NB; b++) { my[p_v0 + b] = accE[0][b]; my[p_v1 + b] = accE[1][b]; my[p_self + b] = accE[2][b]; }
        }
#if FB_RING_PF_MODE == 2 || FB_RING_PF_MODE == 3
        if (!pf_done) do_prefetch();
#endif

        // first incidence records of the next tile (their addresses come from the row record fetched above)
        const int64_t k0n = __double_as_longlong(rawn[1]);
        const int nincn = live_n ? (int)(__double_as_longlong(rawn[2]) >> 32) : 0;
        if (nincn > 0) {
            const int64_t kln = k0n + nincn - 1;
            load_rec<NL>(A, k0n, rc);
            load_rec<NL>(A, k0n + 1 < kln ? k0n + 1 : kln, rn);
            load_rec<NL>(A, k0n + 2 < kln ? k0n + 2 : kln, r2);
            if constexpr (PTS) {
                vc = __ldg(A.vtx + k0n);
                vn = __ldg(A.vtx + (k0n + 1 < kln ? k0n + 1 : kln));
            }
        }

        // write-out: the node's TPR dof rows are one contiguous run of the values array and of shared memory: ONE TMA
        // bulk store per node (SASS UBLKCP) moves the 16-byte aligned interior; the at most two odd doubles go by plain
        // stores.  (One store per dof row is limited by the rate of bulk operations: 45 cycles per SM each.)
        __syncwarp();
        if (a == 0 && n > 0 && A.frag != nullptr && A.nseg == 0) {
            // fragment protocol: whole sectors by one bulk store, the head / tail doubles into the row's slots
            bulk_fence();
            const int total = TPR * n;
            const int64_t fl = __double_as_longlong(raw[3]);
            const RunSplit sp = run_split(outp, total);
            const bool tail_plain = (fl & 32) != 0;
            const int body_n = total - sp.hc - sp.tc;
#ifndef FB_WHATIF_NOSTORE   // timing experiment only: no write-out
            if (body_n > 0) bulk_store(outp + sp.hc, nodep + sp.hc, body_n * 8);
#endif
            if (sp.hc) frag_head(A.frag, fl >> 32, nodep, sp.hc);
            if (sp.tc) {
                if (tail_plain) { for (int x = total - sp.tc; x < total; x++) outp[x] = nodep[x]; }
                else frag_tail(A.frag, fl >> 32, nodep + total - sp.tc, sp.tc);
            }
        } else
        if (a == 0 && n > 0) {
            bulk_fence(); // make the generic-proxy shared-memory writes visible to the async proxy
            const int total = TPR * n;
#pragma unroll 1
            for (int d = 0; d < nrep; d++) {
                double *out = outp + (int64_t)d * total;
                const int h = (int)((reinterpret_cast<uintptr_t>(out) >> 3) & 1);
                if (OPG == 1 || h == head) {
                    const int body_n = (total - h) & ~1;
                    if (h) out[0] = nodep[0];
#ifdef FB_HINT_ST
                    if (body_n > 0) bulk_store_hint(out + h, nodep + h, body_n * 8, policy_evict_first());
#elif !defined(FB_WHATIF_NOSTORE)
                    if (body_n > 0) bulk_store(out + h, nodep + h, body_n * 8);
#endif
                    if (h + body_n < total) out[total - 1] = nodep[total - 1];
                } else { // replicated scalar row whose copy has the other phase: plain stores
                    for (int x = 0; x < total; x++) out[x] = nodep[x];
                }
            }
        }
        if (nincn > 0) {   // next tile's first geometry line lands while the stores drain
            if constexpr (PTS) load_points(A, vc, g);
            else load_geo<DIM, NL>(A, rc, g);
        }
        bulk_commit_wait_read(); // the bulk stores read this warp's shared memory: wait before it is reused
        __syncwarp();
        if (tile_n >= ntiles) break;
        tile = tile_n;
#pragma unroll
        for (int x = 0; x < 4; x++) raw[x] = rawn[x];
    }
}


template <int OPG, bool PTS = false>
__global__ void __launch_bounds__(64, FB_RING_MINBLOCKS) k_ring(const GatherArgs A)
{
    extern __shared__ double acc[];          // [blockDim.x][pitch]
    // (a warp-uniform, compile-time row dof -- 32 row nodes per 96-thread block -- removes the
    // delta_ab selects but makes every geometry load touch 32 lines instead of 11: measured 3.49 ms vs 3.08 ms)
    ring_tiles<OPG, PTS>(A, acc);
}

// -----------------------------------------------------------------------------------------
// Streamed inputs (k_gather_s below).  Every row node owns a small ring of kRsD slots in shared memory; a slot holds
// everything one incidence needs
//     [ record 32 B | geometry line 128 B | element of the incidence kRsAhead places further on, 4 B ]
// and is filled by asynchronous copies (cp.async, SASS LDGSTS) issued kRsAhead = kRsD - 1 incidences before it is
// consumed; the lanes of a node share the 11 chunks.  No register is held across the latency, the main loop contains
// no global load, and the request stream of a row continues into the row the lane processes next (its first elements
// come with the row record, RowInfo::e), so a warp only waits for memory on its very first tile.  The element
// look-ahead word removes the record -> geometry dependence: the address of a geometry line is known kRsAhead
// incidences early without a second level of prefetching.
//
// Completion: a lane commits one cp.async group per incidence it consumes, so the slot of an incidence was filled
// kRsAhead groups earlier and `wait_group kRsAhead - 1` suffices; only where a row's first incidences were requested
// late (first tile, or the previous row had fewer than kRsAhead incidences) the lane waits for everything.
// -----------------------------------------------------------------------------------------
#ifndef FB_RS_D
#define FB_RS_D 3
#endif
constexpr int kRsD = FB_RS_D;                 // ring slots per row node
constexpr int kRsAhead = kRsD - 1;            // incidences between request and use
constexpr int kRsSlotB = 176;                 // bytes per slot: 32 + 128 + 4, rounded up to 16
// bytes per node ring, padded to 16 * (8 k + 1): with nine lanes per node the 16-byte chunk that lane l copies then
// falls into bank group l mod 8, so the eight lanes of a quarter-warp never collide while the nodes use the same slot
constexpr int kRsNodeB = 16 * (((kRsD * kRsSlotB / 16 + 6) / 8) * 8 + 1);
static_assert(kRsNodeB >= kRsD * kRsSlotB && (kRsNodeB / 16) % 8 == 1, "ring stride");
static_assert(kRsD >= 2 && kRsD <= 8, "ring depth");

__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t sdst, const void *gsrc)   // destination as a shared-memory address
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_sh(uint32_t sdst, const void *gsrc, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// one-time (pattern build): look-ahead elements of every row -- RowInfo::e (first incidences) and ahead[k]
__global__ void k_ring_ahead(int64_t n_rows, RowInfo *__restrict__ info, const uint32_t *__restrict__ rec, uint32_t *__restrict__ ahead)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_rows; q += (int64_t)gridDim.x * blockDim.x) {
        RowInfo &R = info[q];
        for (int j = 0; j < 8; j++) R.e[j] = j < R.ninc ? rec[(R.k0 + j) * 8 + 5] : 0u;
        for (int m = 0; m < R.ninc; m++) ahead[R.k0 + m] = m + kRsAhead < R.ninc ? rec[(R.k0 + m + kRsAhead) * 8 + 5] : 0u;
    }
}

// -----------------------------------------------------------------------------------------
// Vertex-node rows of 3D P2 elasticity with streamed inputs: the arithmetic of k_gather<1,3,10,0> (one thread per
// component row (I, a, b), read-modify-write accumulators laid out like the CSR rows, pitch == 3 (mod 16): conflict
// free) behind the input rings above.  k_gather is limited by the L1 data pipe, and 2/3 of its load is the global-load
// return path: each of the nine threads of a node pulls the same 160 bytes per incidence into its registers
// (LDG.256 x 5 = 40 wavefront-equivalents per warp and incidence).  Here the record and the geometry line enter shared
// memory ONCE per node and incidence and are read back with shared-memory loads whose lanes share addresses
// (LDS.128 x 10 = 20 wavefronts: ncu counts two per instruction for three distinct addresses).
// A warp owns NPT = 3 row nodes (27 lanes); warps are persistent and never wait for one another.
// -----------------------------------------------------------------------------------------
#ifndef FB_GS_MINBLOCKS
#define FB_GS_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(64, FB_GS_MINBLOCKS) k_gather_s(const GatherArgs A)
{
    constexpr int DIM = 3, NVTX = 4, NL = 10, NB = 3, CPR = 9, NPT = 3, ROWS = NPT * DIM;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double acc[];          // per warp: ROWS accumulator rows of `pitch` doubles, then NPT rings
    const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
    const int pitch = A.pitch;
    const int slot = lane / CPR;             // node of this lane inside the warp's tile (3 = the five spare lanes)
    const bool lane_used = lane < NPT * CPR;
    const int comp = lane - slot * CPR;
    const int a = comp / NB, b = comp - a * NB;
    const unsigned nmask = lane_used ? (0x1ffu << (CPR * slot)) : (0x1fu << 27);
    const int64_t ntiles = (A.count + NPT - 1) / NPT;
    const size_t rows_bytes = ((size_t)ROWS * pitch * 8 + 15) & ~(size_t)15;   // the rings are 16-byte aligned
    const size_t warp_bytes = rows_bytes + (size_t)NPT * kRsNodeB;
    char *const wbase = reinterpret_cast<char *>(acc) + (size_t)(tid >> 5) * warp_bytes;
    double *const rows = reinterpret_cast<double *>(wbase);
    char *const ring = wbase + rows_bytes + (size_t)(lane_used ? slot : 0) * kRsNodeB;
    const int64_t tile_step = (int64_t)gridDim.x * (NT >> 5);

    // this lane's share of the 11 chunks of a slot: chunk `comp` (record: 0, 1; geometry: 2..8), lane 0 also the
    // last geometry chunk, lane 1 the look-ahead word
    // (per-lane constants: source array and offset of the lane's chunk, shared-memory address of its destination;
    // the look-ahead word travels as the aligned 16-byte chunk of the `ahead` array that holds it)
    const bool is_rec = comp < 2;
    const char *const src0 = is_rec ? reinterpret_cast<const char *>(A.rec) + 16 * comp
                                    : reinterpret_cast<const char *>(A.geom) + 16 * (comp - 2);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t dst0 = ring_s + 16 * comp, dst1 = ring_s + (comp == 0 ? 144 : 160);
    auto request = [&](int rs, int64_t k, uint32_t e) {
        const uint32_t so = rs * kRsSlotB;
        const int64_t line = (int64_t)e * 128;
#ifdef FB_HINT_GEOM
        // records and look-ahead words are read once (evict-first), geometry lines by every incident row (evict-last)
        cp_async16_sh(dst0 + so, src0 + (is_rec ? k * 32 : line), is_rec ? policy_evict_first() : policy_evict_last());
        if (is_rec)
            cp_async16_sh(dst1 + so, comp == 0 ? reinterpret_cast<const char *>(A.geom) + line + 112
                                               : reinterpret_cast<const char *>(A.ahead + (k & ~(int64_t)3)),
                          comp == 0 ? policy_evict_last() : policy_evict_first());
#else
        cp_async16_s(dst0 + so, src0 + (is_rec ? k * 32 : line));
        if (is_rec)
            cp_async16_s(dst1 + so, comp == 0 ? reinterpret_cast<const char *>(A.geom) + line + 112
                                              : reinterpret_cast<const char *>(A.ahead + (k & ~(int64_t)3)));
#endif
    };
    auto wrap = [](int x) { return x >= kRsD ? x - kRsD : x; };

    // M^T of this component: e = |det| G_s^T M G_w,  M = mu (delta_ab I + e_b e_a^T) + lambda e_a e_b^T
    double Mm[DIM][DIM];
#pragma unroll
    for (int c = 0; c < DIM; c++)
#pragma unroll
        for (int d = 0; d < DIM; d++)
            Mm[c][d] = ((a == b && c == d) ? A.c1 : 0.0) + ((c == b && d == a) ? A.c1 : 0.0) + ((c == a && d == b) ? A.c0 : 0.0);

    int64_t tile = (int64_t)blockIdx.x * (NT >> 5) + (tid >> 5);
    if (tile >= ntiles) return;
    double raw[4] = {0.0, 0.0, 0.0, 0.0};
    {
        const int64_t node = tile * NPT + slot;
        if (lane_used && node < A.count) ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + node), raw);
    }
    int s0 = 0, nis = 0;
    bool careful = true;                     // the previous row was too short to request this row's first incidences in time

    for (;;) {
        const bool live = lane_used && tile * NPT + slot < A.count;
        const int64_t base = __double_as_longlong(raw[0]);
        const int64_t k0 = __double_as_longlong(raw[1]);
        const int L = live ? (int)(__double_as_longlong(raw[2]) & 0xffffffff) : 0;
        const int ninc = live ? (int)(__double_as_longlong(raw[2]) >> 32) : 0;
        const int64_t tile_n = tile + tile_step;
        const int64_t node_n = tile_n * NPT + slot;
        const bool live_n = lane_used && tile_n < ntiles && node_n < A.count;
        double rawn[4] = {0.0, 0.0, 0.0, 0.0};
        uint4 en = make_uint4(0u, 0u, 0u, 0u);
        if (live_n) {
            ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + node_n), rawn);
            en = __ldg(reinterpret_cast<const uint4 *>(A.rowinfo[A.start + node_n].e));
        }
        int nisn = 0;
        const int sn0 = (s0 + ninc) % kRsD;

        const int row = (lane_used ? slot : 0) * DIM + a;
        double *my = rows + (size_t)row * pitch + b;
        // the warp's rows were copied out by the warp itself (below): zero them for this tile
        for (int x = lane; x < ROWS * pitch; x += 32) rows[x] = 0.0;
        __syncwarp();
        const int nmax = __reduce_max_sync(FULL, ninc);

        if (ninc > 0 && nis < ninc && nis < kRsAhead) { // catch-up (first tile, short rows)
#pragma unroll
            for (int j = 0; j < kRsAhead; j++)
                if (j >= nis && j < ninc) {
                    const int rs = wrap(s0 + j);
                    request(rs, k0 + j, __ldg(A.rowinfo[A.start + tile * NPT + slot].e + j));
                }
            nis = ninc < kRsAhead ? ninc : kRsAhead;
            cp_async_commit();
            careful = true;
        }
        double dacc = 0.0, accJ = 0.0, accIJ = 0.0, carry[3] = {0.0, 0.0, 0.0};
        int pdiag = 0;
        int cs = s0;
#pragma unroll 1
        for (int m = 0; m < nmax; m++) {
            if (m >= ninc) continue;
            // one group per incidence: the slot of incidence m was requested kRsAhead groups ago, so kRsAhead - 1 newer
            // groups may stay pending -- except at the start of a row whose first incidences were requested late
            if (careful && m < kRsAhead) cp_async_wait<0>();
            else cp_async_wait<kRsAhead - 1>();
            __syncwarp(nmask);
            const char *sl = ring + cs * kRsSlotB;
            {
                const int j = m + kRsAhead;
                if (j < ninc) {
                    const int rs = wrap(cs + kRsAhead);
                    request(rs, k0 + j, *reinterpret_cast<const uint32_t *>(sl + 160 + 4 * (int)((k0 + m) & 3)));
                    nis = j + 1;
                } else if (live_n) {
                    const int jn = j - ninc;
                    const int nincn = (int)(__double_as_longlong(rawn[2]) >> 32);
                    const int64_t k0n = __double_as_longlong(rawn[1]);
                    const int lim = nincn < kRsAhead ? nincn : kRsAhead;
#pragma unroll
                    for (int q = 0; q < kRsAhead; q++)
                        if (q >= nisn && q <= jn && q < lim) {
                            const int rs = wrap(sn0 + q);
                            request(rs, k0n + q, q == 0 ? en.x : (q == 1 ? en.y : (q == 2 ? en.z : en.w)));
                        }
                    if (jn + 1 > nisn) nisn = jn + 1 < lim ? jn + 1 : lim;
                }
                cp_async_commit();
            }
            const uint4 w0 = *reinterpret_cast<const uint4 *>(sl);
            const uint4 w1 = *reinterpret_cast<const uint4 *>(sl + 16);
            const uint32_t pw[5] = {w0.x, w0.y, w0.z, w0.w, w1.x};
            const uint32_t perm = w1.z;
            auto posof = [&](int jc) { return (int)((pw[jc >> 1] >> (16 * (jc & 1))) & 0xffffu) * NB; };
            if (m == 0) pdiag = posof(0);
            double G[NVTX][DIM], adet;
#pragma unroll
            for (int v = 0; v < NVTX; v++) {
                const char *gp = sl + 32 + 32 * ((perm >> (2 * v)) & 3);
                const double2 x = *reinterpret_cast<const double2 *>(gp);
                const double2 y = *reinterpret_cast<const double2 *>(gp + 16);
                G[v][0] = x.x; G[v][1] = x.y; G[v][2] = y.x;
                if (v == 0) adet = y.y;
            }
            // e[w] = |det| G_0^T M G_w (the row function is canonical vertex 0)
            double e[NVTX];
            {
                double v[DIM];
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                    double x = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; c++) x += G[0][c] * Mm[c][d];
                    v[d] = x * adet;
                }
#pragma unroll
                for (int w = 0; w < NVTX; w++) {
                    double x = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) x += v[d] * G[w][d];
                    e[w] = x;
                }
            }
            // chain order (vchain_order): canonical vertices (I, J, w_in, w_out).  J and mid(I, J) accumulate in
            // registers along the chain, the out-face group {3, 7, 8} is carried to the next incidence, where it is
            // the in-face group {2, 6, 5}; only that group and node 9 are updated in shared memory -- 4 updates
            // per incidence instead of 9, plus 5 where the chain ends
            auto contrib = [&](int jc) {
                double v = A.R.r[0][jc][0][0] * e[canon_sv<DIM>(jc, 0)];
                if (jc >= NVTX) v += A.R.r[0][jc][0][1] * e[canon_sv<DIM>(jc, 1)];
                return v;
            };
            const uint32_t mode = (perm >> 8) & 3;
            dacc += contrib(0);
            accJ += contrib(1);
            accIJ += contrib(4);
            {
                double *p2 = my + posof(2), *p6 = my + posof(6), *p5 = my + posof(5), *p9 = my + posof(9);
                const double o2 = *p2, o6 = *p6, o5 = *p5, o9 = *p9;
                *p2 = o2 + (carry[0] + contrib(2));
                *p6 = o6 + (carry[1] + contrib(6));
                *p5 = o5 + (carry[2] + contrib(5));
                *p9 = o9 + contrib(9);
            }
            carry[0] = contrib(3); carry[1] = contrib(7); carry[2] = contrib(8);
            if (mode != 0) {
                double *p3 = my + posof(3), *p7 = my + posof(7), *p8 = my + posof(8), *p1 = my + posof(1), *p4 = my + posof(4);
                const double o3 = *p3, o7 = *p7, o8 = *p8, o1 = *p1, o4 = *p4;
                *p3 = o3 + carry[0]; *p7 = o7 + carry[1]; *p8 = o8 + carry[2]; *p1 = o1 + accJ; *p4 = o4 + accIJ;
                carry[0] = carry[1] = carry[2] = accJ = accIJ = 0.0;
            }
            cs = wrap(cs + 1);
        }
        if (ninc > 0) my[pdiag] = dacc;

        // write-out: the three dof rows of a node are one contiguous run of 9 L values in the values array (ghost rows:
        // in the owner's receive buffer); the warp copies them with full-line coalesced stores
        __syncwarp();
#pragma unroll 1
        for (int sl = 0; sl < NPT; sl++) {
            const int64_t base_s = __shfl_sync(FULL, base, sl * CPR);
            const int L_s = __shfl_sync(FULL, L, sl * CPR);
            if (L_s == 0) continue;
            const int n_s = NB * L_s;
            double *out = out_ptr(A, (int64_t)CPR * base_s);
#pragma unroll
            for (int aa = 0; aa < DIM; aa++) {
                const double *src = rows + (size_t)(sl * DIM + aa) * pitch;
                for (int x = lane; x < n_s; x += 32) st_out(out + aa * n_s + x, src[x]);
            }
        }
        __syncwarp();
        if (tile_n >= ntiles) break;
        tile = tile_n;
        s0 = cs;
        nis = nisn;
        careful = ninc < kRsAhead;
#pragma unroll
        for (int x = 0; x < 4; x++) raw[x] = rawn[x];
    }
    cp_async_wait<0>();
}

// =========================================================================================
// Row-gather kernels of the operators with their own coefficient tensors: advection N(u)
// (FE_def.hpp:1685-1836), advection-in-u W(u) (:1839-1929), the fused Navier-Stokes (0,0) block, and the
// divergence pair B / B^T (:1932-2148).  In barycentric form (grad phi_j = sum_t c_jt(lambda) G_t, P2 values
// quadratic in lambda) every local entry is a contraction of per-element data with a CONSTANT tensor that the
// host forms once from the operator's own quadrature rule (api.cu: op_coefficients), e.g.
//   N_ij     = |det| sum_{m,t} TN[i][m][j][t] (u_m . G_t)           160 FMA per local row instead of 15 points
//   W_ij^ab  = |det| sum_v     MW[i][j][v] D^(v)[b][a]              (D = sum_m u_m (x) grad phi_m is affine)
//   B_i,(j,d)= |det| sum_t     BC[j][t] G_t[d]
// One thread per CSR dof row (I, a); accumulators in a shared-memory row laid out like the CSR row.
//
// The SCALAR part of N(u) and of the fused Navier-Stokes block -- the same number on the three diagonal entries of a node
// block -- is evaluated ONCE PER ELEMENT by k_sloc (element-stationary, one thread per element):
//   S_e[i][j] = c0 |det| sum_{s,t} RLF[i][j][s][t] G_sv(i,s).G_sv(j,t)  +  c1 |det| sum_{m,t} TNF[i][m][j][t] (u_m . G_sv(j,t))
// (natural local order, 800 bytes per P2 tetrahedron), and the row kernels read their row of it (ten 8-byte loads per
// incidence) instead of contracting the 160-term tensor in every one of the 3 x 10 dof-row threads that meet the element:
// k_gatherx<X_NSJ> went from 246 registers / 900 instructions per incidence to the W(u) contraction alone.
// =========================================================================================
struct OpCoef {
    double MW[2][MAXN][4];        // [row type][j'][v']       sum_q w lambda_v' phi_i' phi_j'
    double BC[MAXN][2];           // B:   [j'][t']            sum_q w psi_0 c_{j' t'}
    double BTC[2][4][2];          // B^T: [row type][j'][s']  sum_q w psi_j' c_{i' s'}
    double MM[2][MAXN];           // mass: [row type][j']     sum_q w phi_i' phi_j'
    double c0, c1, c2;            // rho*nu, rho, rho (Newton) of the fused block
};

// per-element velocity data (natural local order), written by k_udata after k_geom:
//   uel[e][m]    = (u_m.x, u_m.y, u_m.z, 0)
//   dt[e][a][b]  = |det| * (D^(v)[d2 = b][d1 = a] for the element's vertices v = 0..3)
template <int DIM, int NL>
__global__ void __launch_bounds__(256) k_udata(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ geom,
                                               const double *__restrict__ u, double *__restrict__ uel, double *__restrict__ dt,
                                               int want_dt)
{
    constexpr int NVTX = DIM + 1, GS = GeomStride<DIM>::value;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    double U[NL][DIM], G[NVTX][DIM];
#pragma unroll
    for (int m = 0; m < NL; m++) {
        const int64_t n = conn[e * NL + m];
#pragma unroll
        for (int d = 0; d < DIM; d++) U[m][d] = u[n * DIM + d];
        if (uel != nullptr) st_v4(uel + (e * NL + m) * 4, U[m][0], U[m][1], DIM == 3 ? U[m][DIM - 1] : 0.0, 0.0);
    }
    if (!want_dt) return;
    const double *g = geom + e * GS;
    double adet;
    if constexpr (DIM == 3) {
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
            for (int d = 0; d < 3; d++) G[v][d] = g[4 * v + d];
        adet = g[3];
    } else {
#pragma unroll
        for (int v = 0; v < 3; v++) { G[v][0] = g[2 * v]; G[v][1] = g[2 * v + 1]; }
        adet = g[6];
    }
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) {
            double *o = dt + ((e * DIM + a) * DIM + b) * 4;
            double ov[4];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                double s = 0.0;
                if (v < NVTX) {
                    if constexpr (NL == NVTX) { // P1: the gradient of u is constant
#pragma unroll
                        for (int w = 0; w < NVTX; w++) s += U[w][a] * G[w][b];
                    } else {
                        // grad phi_w(e_v) = (4 delta_wv - 1) G_w;  edge (p,q) containing v: 4 G_{other end}
#pragma unroll
                        for (int w = 0; w < NVTX; w++) s += (w == v ? 3.0 : -1.0) * U[w][a] * G[w][b];
#pragma unroll
                        for (int ed = 0; ed < NL - NVTX; ed++) {
                            const int p = canon_sv<DIM>(NVTX + ed, 0), q = canon_sv<DIM>(NVTX + ed, 1);
                            if (p == v) s += 4.0 * U[NVTX + ed][a] * G[q][b];
                            if (q == v) s += 4.0 * U[NVTX + ed][a] * G[p][b];
                        }
                    }
                }
                ov[v] = s * adet;
            }
            st_v4(o, ov[0], ov[1], ov[2], ov[3]);
        }
}

// natural vertex support of local node i (run-time index): vertex functions sit on one vertex, edge functions on two
template <int DIM>
__device__ __forceinline__ int sv_rt(int i, int t)
{
    constexpr int NVTX = DIM + 1;
    // packed 2-bit fields of canon_sv's edge tables
    constexpr uint32_t P0 = DIM == 2 ? (0u | 1u << 2 | 0u << 4) : (0u | 1u << 2 | 0u << 4 | 0u << 6 | 1u << 8 | 2u << 10);
    constexpr uint32_t P1 = DIM == 2 ? (1u | 2u << 2 | 2u << 4) : (1u | 2u << 2 | 2u << 4 | 3u << 6 | 3u << 8 | 3u << 10);
    if (i < NVTX) return i;
    return (int)(((t == 0 ? P0 : P1) >> (2 * (i - NVTX))) & 3u);
}

// Element-stationary pre-pass of N(u) and of the fused Navier-Stokes block: the scalar local matrix S_e (see the header of
// this section), one thread per element, after k_geom.  tab = TNF[NL][NL][NL][2] | RLF[NL][NL][2][2] (api.cu: sloc_tables),
// staged in shared memory and read with warp-uniform addresses; the element's Gram matrix |det| G_a.G_b sits in a
// per-thread shared-memory column because the row loop indexes it at run time.
// (Tried: the tensors as immediate constant-bank operands with every loop unrolled -- no loads at all, but 58 KB of
// straight-line code: 0.83 ms instead of 0.50 ms for 750 k elements, instruction-cache bound.  Register caps through
// __launch_bounds__(128, 3 | 4): 168 / 128 registers, both slower than the 134 the compiler picks on its own.)
template <int DIM, int NL>
__global__ void __launch_bounds__(128) k_sloc(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ geom,
                                              const double *__restrict__ u, const double *__restrict__ tab, double c0, double c1,
                                              double *__restrict__ sloc)
{
    constexpr int NVTX = DIM + 1, GS = GeomStride<DIM>::value;
    constexpr int NTN = NL * NL * NL * 2, NRL = NL * NL * 4;
    constexpr bool P2 = NL > NVTX;
    extern __shared__ double sm_sloc[];
    double *const tn = sm_sloc, *const rl = sm_sloc + NTN, *const gmall = rl + NRL;
    const int NT = blockDim.x, tid = threadIdx.x;
    for (int x = tid; x < NTN + NRL; x += NT) sm_sloc[x] = tab[x];
    __syncthreads();
    const int64_t e = blockIdx.x * (int64_t)NT + tid;
    if (e >= ne) return;
    double G[NVTX][DIM], adet;
    {
        const double *g = geom + e * GS;
        if constexpr (DIM == 3) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                double t4[4];
                ld_v4(g + 4 * v, t4);
                G[v][0] = t4[0]; G[v][1] = t4[1]; G[v][2] = t4[2];
                if (v == 0) adet = t4[3];
            }
        } else {
#pragma unroll
            for (int v = 0; v < 3; v++) { G[v][0] = g[2 * v]; G[v][1] = g[2 * v + 1]; }
            adet = g[6];
        }
    }
    double *const gm = gmall + tid;   // gm[(a * 4 + b) * NT]
#pragma unroll
    for (int a = 0; a < NVTX; a++)
#pragma unroll
        for (int b = 0; b < NVTX; b++) {
            double x = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) x += G[a][d] * G[b][d];
            gm[(a * 4 + b) * NT] = x * adet;
        }
    double s[NL][NVTX];   // |det| u_m . G_t
#pragma unroll
    for (int m = 0; m < NL; m++) {
        const int64_t n = conn[e * NL + m];
        double um[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) um[d] = u != nullptr ? u[n * DIM + d] : 0.0;
#pragma unroll
        for (int t = 0; t < NVTX; t++) {
            double x = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) x += um[d] * G[t][d];
            s[m][t] = x * adet;
        }
    }
    double *out = sloc + e * (int64_t)(NL * NL);
#ifndef FB_SLOC_UNROLL
#define FB_SLOC_UNROLL 1
#endif
    constexpr int kSlocUnroll = FB_SLOC_UNROLL;
#pragma unroll(kSlocUnroll)
    for (int i = 0; i < NL; i++) {
        double acc[NL];
#pragma unroll
        for (int j = 0; j < NL; j++) acc[j] = 0.0;
        const double *T = tn + i * (NL * NL * 2);
#pragma unroll
        for (int m = 0; m < NL; m++)
#pragma unroll
            for (int j = 0; j < NL; j++) {
                const double2 c = *reinterpret_cast<const double2 *>(T + (m * NL + j) * 2);
                acc[j] = fma(c.x, s[m][canon_sv<DIM>(j, 0)], acc[j]);
                if (P2 && j >= NVTX) acc[j] = fma(c.y, s[m][canon_sv<DIM>(j, 1)], acc[j]);
            }
        const int i0 = sv_rt<DIM>(i, 0), i1 = sv_rt<DIM>(i, 1);
        const double *R = rl + i * (NL * 4);
#pragma unroll
        for (int j = 0; j < NL; j++) {
            const double2 r0 = *reinterpret_cast<const double2 *>(R + j * 4), r1 = *reinterpret_cast<const double2 *>(R + j * 4 + 2);
            double lap = r0.x * gm[(i0 * 4 + canon_sv<DIM>(j, 0)) * NT];
            lap = fma(r1.x, gm[(i1 * 4 + canon_sv<DIM>(j, 0)) * NT], lap);      // zero coefficient for a vertex row function
            if (P2 && j >= NVTX) {
                lap = fma(r0.y, gm[(i0 * 4 + canon_sv<DIM>(j, 1)) * NT], lap);
                lap = fma(r1.y, gm[(i1 * 4 + canon_sv<DIM>(j, 1)) * NT], lap);
            }
            acc[j] = c0 * lap + c1 * acc[j];
        }
#pragma unroll
        for (int j = 0; j < NL; j++) out[i * NL + j] = acc[j];
    }
}

enum OpX { X_ADV = 0, X_ADVU = 1, X_NSJ = 2, X_B = 3, X_BT = 4, X_MASS = 5 };
template <int OPX, int DIM> struct OpXShape {
    static constexpr int RD = (OPX == X_ADV || OPX == X_B || OPX == X_MASS) ? 1 : DIM;   // threads (row dofs) per row node
    static constexpr int NB = (OPX == X_ADV || OPX == X_BT || OPX == X_MASS) ? 1 : DIM;  // values per column node in a thread's row
    static constexpr bool SQUARE = OPX == X_ADV || OPX == X_ADVU || OPX == X_NSJ || OPX == X_MASS;
    static constexpr bool NEEDS_U = OPX == X_ADV || OPX == X_ADVU || OPX == X_NSJ;
};

struct GatherXArgs {
    const RowInfo *rowinfo;
    int64_t start, count;
    const uint32_t *rec;
    const double *geom, *sloc, *dt;   // geometry lines, scalar local matrices of k_sloc, |det| grad u of k_udata
    double *values;
    int pitch;                // doubles per thread in shared memory (odd)
    int vec_dim;              // mass: 0 scalar, DIM = replicate to the DIM block-diagonal dof rows
    OpCoef C;                 // coefficient tensors of the operator: kernel parameter (constant bank), so every launch carries
                              // its own copy -- contexts, streams and host threads never share coefficient state
};

// the warp copies the 32 shared-memory rows of its threads to their CSR rows with coalesced stores
__device__ __forceinline__ void warp_write_rows(const double *wbase, int pitch, int lane, int n, int64_t off, int nrep, double *values)
{
    const double *src = wbase + lane;
#pragma unroll 1
    for (int r = 0; r < 32; r++, src += pitch) {
        const int nr = __shfl_sync(0xffffffffu, n, r);
        const int64_t o = __shfl_sync(0xffffffffu, off, r);
#pragma unroll 1
        for (int d = 0; d < nrep; d++) {
            double *out = values + o + (int64_t)d * nr + lane;
            if (lane < nr) out[0] = src[0];
            if (lane + 32 < nr) out[32] = src[32];
            if (lane + 64 < nr) out[64] = src[64];
            for (int x = lane + 96; x < nr; x += 32) out[x - lane] = src[x - lane];
        }
    }
}

// NLR / NL: local nodes of the row / column space;  TYPE: 0 vertex-node rows, 1 edge-node rows (of the row space)
template <int OPX, int DIM, int NLR, int NL, int TYPE>
__global__ void __launch_bounds__(64, 4) k_gatherx(const GatherXArgs A)
{
    using S = OpXShape<OPX, DIM>;
    constexpr int RD = S::RD, NB = S::NB, NVTX = DIM + 1;
    constexpr int NS = TYPE == 0 ? 1 : 2;
    constexpr int JD = !S::SQUARE ? -1 : (TYPE == 0 ? 0 : NVTX); // column that is the row node itself
    constexpr int GS = GeomStride<DIM>::value;
    constexpr bool P2C = NL > NVTX;                                // P2 column space
    extern __shared__ double acc[];                                // [blockDim.x][pitch]
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, pitch = A.pitch;
    const int64_t t = blockIdx.x * (int64_t)NT + tid;
    const bool live = t < A.count * RD;
    const int64_t rloc = live ? t / RD : 0;
    const int a = live ? (int)(t - rloc * RD) : 0;
    int64_t base = 0, k0 = 0;
    int L = 0, ninc = 0;
    if (live) {
        double raw[4];
        ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + rloc), raw);
        base = __double_as_longlong(raw[0]);
        k0 = __double_as_longlong(raw[1]);
        const int64_t ln = __double_as_longlong(raw[2]);
        L = (int)(ln & 0xffffffff);
        ninc = (int)(ln >> 32);
    }
    double *wbase = acc + (size_t)(tid - lane) * pitch;
    double *my = acc + (size_t)tid * pitch;
    for (int x = lane; x < 32 * pitch; x += 32) wbase[x] = 0.0;
    __syncwarp();

    GatherArgs RA; // only .rec / .geom are used by the shared load helpers
    RA.rec = A.rec; RA.geom = A.geom;
    double dacc[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) dacc[b] = 0.0;
    int pdiag = 0;
    // fused block (where it runs through this kernel): records two incidences ahead and the next incidence's row of S_e
    // requested into L2 one iteration early (see k_gatherw).  Not for N(u) alone: 1.65 -> 1.53 ms on the structured cube
    // (M = 50) but 4.03 -> 4.75 ms on config 4's unstructured mesh (FB_GATHERX_ADV_AHEAD switches it on)
#ifdef FB_GATHERX_ADV_AHEAD
    constexpr bool AHEAD = OPX == X_ADV || OPX == X_NSJ;
#else
    constexpr bool AHEAD = OPX == X_NSJ;
#endif
    IncRec<NL> rc, rn, r2;
    if (ninc > 0) {
        load_rec<NL>(RA, k0, rc);
        if constexpr (AHEAD) load_rec<NL>(RA, k0 + (ninc > 1 ? 1 : 0), rn);
    }
    for (int k = 0; k < ninc; k++) {
        if constexpr (AHEAD) {
            load_rec<NL>(RA, k0 + (k + 2 < ninc ? k + 2 : ninc - 1), r2);
#ifndef FB_GATHERW_NO_PREFETCH
            if (k + 1 < ninc) {
                const double *sn = A.sloc + ((int64_t)rec_elem<NL>(rn.w) * NL + rec_natidx<NL>(rn.w, JD < 0 ? 0 : JD)) * NL;
#pragma unroll
                for (int x = 0; x < NL; x += 4) FB_PF_ELEM(sn + x);
            }
#endif
        } else if (k + 1 < ninc) load_rec<NL>(RA, k0 + k + 1, rn);
        const uint32_t perm = rec_perm<NL>(rc.w);
        const int64_t e = rec_elem<NL>(rc.w);
        IncGeo<DIM> g;
        if constexpr (OPX == X_MASS) g.G[0][3] = __ldg(RA.geom + e * GS + (DIM == 3 ? 3 : 6)); // only |det| is needed
        else if constexpr (OPX == X_B || OPX == X_BT) load_geo<DIM, NL>(RA, rc, g);
        double val[NL][NB];
#pragma unroll
        for (int j = 0; j < NL; j++)
#pragma unroll
            for (int b = 0; b < NB; b++) val[j][b] = 0.0;

        if constexpr (OPX == X_ADV || OPX == X_NSJ) {
            // scalar part: the row of the element's scalar local matrix (k_sloc) that belongs to the row node, natural order
            const double *srow = A.sloc + (e * NL + rec_natidx<NL>(rc.w, JD)) * NL;
#pragma unroll
            for (int j = 0; j < NL; j++) {
                const double sv = __ldg(srow + rec_natidx<NL>(rc.w, j));
                if constexpr (OPX == X_ADV) val[j][0] = sv;
                else {
#pragma unroll
                    for (int b = 0; b < NB; b++) val[j][b] = (b == a) ? sv : 0.0;
                }
            }
        }
        if constexpr (OPX == X_ADVU || OPX == X_NSJ) {
            // W^{ab}_{i'j'} = sum_{v'} MW[j'][v'] dt[a][b][pi(v')]
            const double cw = OPX == X_NSJ ? A.C.c2 : 1.0;
#pragma unroll
            for (int b = 0; b < DIM; b++) {
                double dn[4], dv[NVTX];
                ld_v4(A.dt + ((e * DIM + a) * DIM + b) * 4, dn);
#pragma unroll
                for (int v = 0; v < NVTX; v++) {
                    const int pv = (perm >> (2 * v)) & 3;
                    dv[v] = pv == 0 ? dn[0] : (pv == 1 ? dn[1] : (pv == 2 ? dn[2] : dn[3]));
                }
#pragma unroll
                for (int j = 0; j < NL; j++) {
                    double x = 0.0;
#pragma unroll
                    for (int v = 0; v < NVTX; v++) x += A.C.MW[TYPE][j][v] * dv[v];
                    val[j][b] += cw * x;
                }
            }
        }
        if constexpr (OPX == X_MASS) {
            const double adet = g.G[0][3];
#pragma unroll
            for (int j = 0; j < NL; j++) val[j][0] = A.C.MM[TYPE][j] * adet - A.C.c0 * adet * A.C.c1; // c0 = 0: mass; else BD stabilisation
        }
        if constexpr (OPX == X_B) {
            // row = pressure vertex (canonical vertex 0); B_{i',(j',d)} = |det| sum_t' BC[j'][t'] G_{t'}[d]
            const double adet = g.G[0][3];
#pragma unroll
            for (int j = 0; j < NL; j++)
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                    double x = A.C.BC[j][0] * g.G[canon_sv<DIM>(j, 0)][d];
                    if (P2C && j >= NVTX) x += A.C.BC[j][1] * g.G[canon_sv<DIM>(j, 1)][d];
                    val[j][d] = x * adet;
                }
        }
        if constexpr (OPX == X_BT) {
            // row = velocity node i' dof a, columns = the element's pressure vertices (canonical order)
            const double adet = g.G[0][3];
#pragma unroll
            for (int j = 0; j < NL; j++) {
                double x = 0.0;
#pragma unroll
                for (int s2 = 0; s2 < NS; s2++) {
                    const double ga = a == 0 ? g.G[s2][0] : (a == 1 ? g.G[s2][1] : g.G[s2][2]);
                    x += A.C.BTC[TYPE][j][s2] * ga;
                }
                val[j][0] = x * adet;
            }
        }
        // accumulate: the row node's own column in registers, the others in the thread's shared-memory row
#pragma unroll
        for (int j = 0; j < NL; j++) {
            if (j == JD) {
                if (k == 0) pdiag = (int)((rc.w[j >> 1] >> (16 * (j & 1))) & 0xffffu) * NB;
#pragma unroll
                for (int b = 0; b < NB; b++) dacc[b] += val[j][b];
            } else {
                double *p = my + ((rc.w[j >> 1] >> (16 * (j & 1))) & 0xffffu) * NB;
#pragma unroll
                for (int b = 0; b < NB; b++) p[b] += val[j][b];
            }
        }
        rc = rn;
        if constexpr (AHEAD) rn = r2;
    }
    if (JD >= 0 && ninc > 0) {
#pragma unroll
        for (int b = 0; b < NB; b++) my[pdiag + b] = dacc[b];
    }
    __syncwarp();
    const int n = live ? NB * L : 0;
    const int nrep = OPX == X_ADV ? DIM : ((OPX == X_MASS && A.vec_dim != 0) ? A.vec_dim : 1);
    const int64_t off = (OPX == X_ADV || OPX == X_MASS) ? (int64_t)nrep * base : (int64_t)RD * NB * base + (int64_t)a * n;
    warp_write_rows(wbase, pitch, lane, n, off, nrep, A.values);
}

// Per-component variant of the row-gather kernel for W(u) and the fused Navier-Stokes block (3 x 3 node blocks): one thread
// per COMPONENT row (I, a, b), DIM * DIM threads per row node -- it contracts one row of |det| grad u (one 32-byte sector)
// with MW for the NL column nodes (4 FMA each) and owns the entries NB p + b of the shared-memory row of dof row (I, a),
// which is laid out like the CSR row (pitch == NB (mod 16): the threads of a node never collide, see k_gather).  A third of
// the registers and of the shared memory per thread of k_gatherx: 12 instead of 4 resident warps on the vertex-node rows
// (config 4 Navier-Stokes block 9.1 -> 5.8 ms).
template <int OPX, int DIM, int NL, int TYPE>
__global__ void __launch_bounds__(32 * DIM, 4) k_gatherw(const GatherXArgs A)
{
    static_assert(OPX == X_ADVU || OPX == X_NSJ, "W(u) / Navier-Stokes block only");
    constexpr int NB = DIM, CPR = DIM * DIM, NT = 32 * NB, ROWS = 32, NVTX = DIM + 1;
    constexpr int JD = TYPE == 0 ? 0 : NVTX;
    extern __shared__ double acc[];                  // [ROWS][pitch]
    __shared__ int64_t s_off[ROWS];
    __shared__ int s_n[ROWS];
    const int tid = threadIdx.x, pitch = A.pitch;
    const int64_t t = blockIdx.x * (int64_t)NT + tid;
    const bool live = t < A.count * CPR;
    const int64_t rloc = live ? t / CPR : 0;
    const int comp = live ? (int)(t - rloc * CPR) : 0;
    const int a = comp / NB, b = comp - a * NB;
    int64_t base = 0, k0 = 0;
    int L = 0, ninc = 0;
    if (live) {
        double raw[4];
        ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + rloc), raw);
        base = __double_as_longlong(raw[0]);
        k0 = __double_as_longlong(raw[1]);
        const int64_t ln = __double_as_longlong(raw[2]);
        L = (int)(ln & 0xffffffff);
        ninc = (int)(ln >> 32);
    }
    GatherArgs RA; // only .rec is used by the shared load helper
    RA.rec = A.rec;
    // records two incidences ahead; the element data of the NEXT incidence (its sector of |det| grad u, its row of S_e) is
    // requested into L2 one iteration early -- no registers held across the latency (ncu before: 60 % of the main loop's
    // stall samples sat on the first uses of these two loads)
    // (Navier-Stokes block only: W(u) alone has one load per incidence and measured 5 % slower with the look-ahead)
    constexpr bool AHEAD = OPX == X_NSJ;
    IncRec<NL> rc, rn, r2;
    if (ninc > 0) {
        load_rec<NL>(RA, k0, rc);
        if constexpr (AHEAD) load_rec<NL>(RA, k0 + (ninc > 1 ? 1 : 0), rn);
    }
    for (int x = tid; x < ROWS * pitch; x += NT) acc[x] = 0.0;
    const int row = tid / NB;
    if (b == 0) {
        s_n[row] = live ? NB * L : 0;
        s_off[row] = (int64_t)CPR * base + (int64_t)a * NB * L;
    }
    __syncthreads();
    double *my = acc + (size_t)row * pitch + b;
    const double cw = OPX == X_NSJ ? A.C.c2 : 1.0;
    double dacc = 0.0;
    int pdiag = 0;
    for (int k = 0; k < ninc; k++) {
        if constexpr (AHEAD) load_rec<NL>(RA, k0 + (k + 2 < ninc ? k + 2 : ninc - 1), r2);
        else if (k + 1 < ninc) load_rec<NL>(RA, k0 + k + 1, rn);
#ifndef FB_GATHERW_NO_PREFETCH
        if (AHEAD && k + 1 < ninc) {
            const int64_t en = rec_elem<NL>(rn.w);
            FB_PF_ELEM(A.dt + ((en * DIM + a) * DIM + b) * 4);
            if (OPX == X_NSJ && a == b) {
                const double *sn = A.sloc + (en * NL + rec_natidx<NL>(rn.w, JD)) * NL;
                FB_PF_ELEM(sn); FB_PF_ELEM(sn + 4); FB_PF_ELEM(sn + 8);
            }
        }
#endif
        const uint32_t perm = rec_perm<NL>(rc.w);
        const int64_t e = rec_elem<NL>(rc.w);
        // |det| grad u (component a, b) at the canonical vertices: four 8-byte loads of one sector, addressed through the
        // permutation (a 32-byte load + register selects costs 24 instructions more)
        double dv[NVTX];
        {
            const double *dp = A.dt + ((e * DIM + a) * DIM + b) * 4;
#ifdef FB_GATHERW_V4   // A/B: one 32-byte load + register selects
            double dn[4];
            ld_v4(dp, dn);
#pragma unroll
            for (int v = 0; v < NVTX; v++) {
                const int pv = (perm >> (2 * v)) & 3;
                dv[v] = cw * (pv == 0 ? dn[0] : (pv == 1 ? dn[1] : (pv == 2 ? dn[2] : dn[3])));
            }
#else
#pragma unroll
            for (int v = 0; v < NVTX; v++) dv[v] = cw * __ldg(dp + ((perm >> (2 * v)) & 3));
#endif
        }
        double val[NL];
        if (OPX == X_NSJ && a == b) {   // scalar part (k_sloc): the diagonal components only
            const double *srow = A.sloc + (e * NL + rec_natidx<NL>(rc.w, JD)) * NL;
#pragma unroll
            for (int j = 0; j < NL; j++) val[j] = __ldg(srow + rec_natidx<NL>(rc.w, j));
        } else {
#pragma unroll
            for (int j = 0; j < NL; j++) val[j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NL; j++) {
            double x = val[j];
#pragma unroll
            for (int v = 0; v < NVTX; v++) x = fma(A.C.MW[TYPE][j][v], dv[v], x);
            val[j] = x;
        }
        if (k == 0) pdiag = (int)((rc.w[JD >> 1] >> (16 * (JD & 1))) & 0xffffu) * NB;
        dacc += val[JD];
        // distinct canonical nodes hit distinct row positions: load all, add, store all
        constexpr int NO = NL - 1;
        double *ptr[NO];
        double old[NO];
#pragma unroll
        for (int jj = 0; jj < NO; jj++) {
            const int jc = jj + (jj >= JD ? 1 : 0);
            ptr[jj] = my + ((rc.w[jc >> 1] >> (16 * (jc & 1))) & 0xffffu) * NB;
            old[jj] = *ptr[jj];
        }
#pragma unroll
        for (int jj = 0; jj < NO; jj++) {
            const int jc = jj + (jj >= JD ? 1 : 0);
            *ptr[jj] = old[jj] + val[jc];
        }
        rc = rn;
        if constexpr (AHEAD) rn = r2;
    }
    if (ninc > 0) my[pdiag] = dacc;
    __syncthreads();
    const int lane = tid & 31;
    for (int r = tid >> 5; r < ROWS; r += NT / 32) {
        const int nr = s_n[r];
        const double *src = acc + (size_t)r * pitch;
        double *out = A.values + s_off[r];
#pragma unroll 4
        for (int x = lane; x < nr; x += 32) out[x] = src[x];
    }
}

} // namespace fb
