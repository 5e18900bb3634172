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
            const double w = T.w[q];
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
    if constexpr (DIM == 3) {
#pragma unroll
        for (int v = 0; v < 4; v++) {
            g[4 * v + 0] = G[v][0]; g[4 * v + 1] = G[v][1]; g[4 * v + 2] = G[v][2]; g[4 * v + 3] = adet;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 3; v++) { g[2 * v] = G[v][0]; g[2 * v + 1] = G[v][1]; }
        g[6] = adet; g[7] = adet;
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

// position words per incidence: 16 bit per canonical local node, padded to a 16-byte multiple
template <int NL> struct PoscStride { static constexpr int value = NL <= 4 ? 4 : (NL <= 8 ? 8 : 16); };

// Per-incidence gather records (one-time, pattern build):
//   incp[k]  = (element << 8) | (pi(3)<<6 | pi(2)<<4 | pi(1)<<2 | pi(0))   canonical vertex permutation
//   posc[k][jc] = position, in the row of the incidence's row node, of canonical local node jc
//   rtype[row]  = 0 vertex node / 1 edge node
template <int DIM, int NL>
__global__ void k_canon_pos(int64_t n_rows, const int64_t *__restrict__ inc_ptr, const int32_t *__restrict__ inc,
                            const int64_t *__restrict__ rowptr, const uint16_t *__restrict__ pos, int pos_stride,
                            uint16_t *__restrict__ posc, uint32_t *__restrict__ incp, int8_t *__restrict__ rtype)
{
    constexpr int PS = PoscStride<NL>::value;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int8_t ty = 0;
        // first-touch bitmap of the row's positions (rows longer than 1024 nodes are zero-initialised instead)
        uint64_t seen[16];
        for (int w = 0; w < 16; w++) seen[w] = 0;
        const int len = (int)(rowptr[r + 1] - rowptr[r]);
        const bool track = len <= 1024;
        int ntouched = 0;
        for (int64_t k = inc_ptr[r]; k < inc_ptr[r + 1]; k++) {
            const int32_t code = inc[k];
            const int64_t e = code >> 4;
            const int i = code & 15;
            ty = i > DIM ? 1 : 0;
            int pi[DIM + 1];
            canon_perm<DIM>(i, pi);
            uint32_t bits = 0;
            for (int v = 0; v <= DIM; v++) bits |= (uint32_t)pi[v] << (2 * v);
            incp[k] = ((uint32_t)e << 8) | bits;
            for (int jc = 0; jc < PS; jc++) {
                uint16_t w = 0;
                if (jc < NL) {
                    w = pos[(e * NL + i) * pos_stride + canon_node<DIM>(pi, jc)];
                    if (track && !((seen[w >> 6] >> (w & 63)) & 1)) {
                        seen[w >> 6] |= uint64_t(1) << (w & 63);
                        ntouched++;
                        w |= 0x8000; // first contribution to this position: store instead of add
                    }
                }
                posc[k * PS + jc] = w;
            }
        }
        // bit 0: row type; bit 1: some position gets no local contribution -> accumulators need zero-init
        rtype[r] = ty | ((!track || ntouched < len) ? 2 : 0);
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
    int64_t pad;
};

struct GatherArgs {
    const RowInfo *rowinfo;   // rows of this launch: rowinfo[start .. start+count), bucket order
    int64_t start, count;
    const uint32_t *incp;
    const uint16_t *posc;
    const double *geom;
    double c0, c1;            // lambda, mu
    double *values;
    int lcap;
    int vec_dim;              // LAP: 0 scalar, DIM = replicate to the DIM block-diagonal dof rows
    CanonR R;
};

__device__ __forceinline__ void st_v4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ld_v4(const double *p, double (&v)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
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

// one incidence's operands, as loaded from global memory
template <int DIM, int NL>
struct IncData {
    double G[DIM + 1][4];                       // canonical vertex v: (Gx, Gy[, Gz], |det|) -- 2D uses [0],[1] and [3]
    uint32_t P[PoscStride<NL>::value / 2];      // packed 16-bit positions
};

template <int DIM, int NL>
__device__ __forceinline__ void load_inc(const GatherArgs &A, int64_t k, uint32_t code, IncData<DIM, NL> &D)
{
    constexpr int GS = GeomStride<DIM>::value;
    constexpr int PW = PoscStride<NL>::value / 2;
    const double *g = A.geom + (int64_t)(code >> 8) * GS;
    if constexpr (DIM == 3) {
#pragma unroll
        for (int v = 0; v < 4; v++) ld_v4(g + 4 * ((code >> (2 * v)) & 3), D.G[v]);
    } else {
        const double ad = __ldg(g + 6);
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(g) + ((code >> (2 * v)) & 3));
            D.G[v][0] = t.x; D.G[v][1] = t.y; D.G[v][2] = 0.0; D.G[v][3] = ad;
        }
    }
    const uint32_t *pw = reinterpret_cast<const uint32_t *>(A.posc) + k * PW;
    if constexpr (PW == 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(pw)), b = __ldg(reinterpret_cast<const uint4 *>(pw) + 1);
        D.P[0] = a.x; D.P[1] = a.y; D.P[2] = a.z; D.P[3] = a.w; D.P[4] = b.x; D.P[5] = b.y; D.P[6] = b.z; D.P[7] = b.w;
    } else if constexpr (PW == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(pw));
        D.P[0] = a.x; D.P[1] = a.y; D.P[2] = a.z; D.P[3] = a.w;
    } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(pw));
        D.P[0] = a.x; D.P[1] = a.y;
    }
}

__device__ __forceinline__ double sel3(const double (&v)[4], int c) { return c == 0 ? v[0] : (c == 1 ? v[1] : v[2]); }

// One incidence of one thread: evaluate the thread's part of local row i' (canonical) and add it to the
// lane-private accumulators.  OPG 0: Laplace, 1 value per column node.  OPG 1: elasticity, the thread
// owns row dof `a` and produces the DIM column dofs b of every column node.
template <int OPG, int DIM, int NL, int TYPE>
__device__ __forceinline__ void gather_incidence(const GatherArgs &A, const IncData<DIM, NL> &D, int a, double *my, int NT,
                                                 double (&dacc)[OPG == 1 ? DIM : 1])
{
    constexpr int NVTX = DIM + 1;
    constexpr int NS = TYPE == 0 ? 1 : 2;   // canonical support size of the row function
    constexpr int NB = OPG == 1 ? DIM : 1;  // values per column node
    const double adet = D.G[0][3];
    // E[s][w][b] for canonical row-support vertex s and canonical vertex w
    double E[NS][NVTX][NB];
    if constexpr (OPG == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int w = 0; w < NVTX; w++) {
                double dot = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) dot += D.G[s][d] * D.G[w][d];
                E[s][w][0] = dot * adet;
            }
    } else {
        const double mu = A.c1 * adet, lam = A.c0 * adet;
        double Ga[NVTX];
#pragma unroll
        for (int w = 0; w < NVTX; w++) Ga[w] = sel3(D.G[w], a);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const double ls = lam * Ga[s];
#pragma unroll
            for (int w = 0; w < NVTX; w++) {
                double dot = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; d++) dot += D.G[s][d] * D.G[w][d];
                const double mdot = mu * dot, mga = mu * Ga[w];
                // E^{ab}_{sw} = mu (delta_ab G_s.G_w + G_s[b] G_w[a]) + lambda G_s[a] G_w[b]
#pragma unroll
                for (int b = 0; b < DIM; b++) E[s][w][b] = (a == b ? mdot : 0.0) + mga * D.G[s][b] + ls * D.G[w][b];
            }
        }
    }
    // the row node itself (canonical node 0 of a vertex row, canonical edge (0,1) of an edge row) receives a
    // contribution from every incidence: it is accumulated in registers
    constexpr int JD = TYPE == 0 ? 0 : NVTX;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        double v = dacc[b];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            v += A.R.r[TYPE][JD][s][0] * E[s][canon_sv<DIM>(JD, 0)][b];
            if (JD >= NVTX) v += A.R.r[TYPE][JD][s][1] * E[s][canon_sv<DIM>(JD, 1)][b];
        }
        dacc[b] = v;
    }
    // distinct canonical nodes hit distinct row positions: per chunk of column nodes load all
    // accumulators, add, store all (keeps the shared-memory round trips independent).  Bit 15 of a
    // position word marks the first contribution to that position: it is stored, not added.
    constexpr int NO = NL - 1;               // off-diagonal column nodes
    constexpr int JB = NO <= 5 ? NO : (NO == 9 ? 5 : 3);
#pragma unroll
    for (int j0 = 0; j0 < NO; j0 += JB) {
        int idx[JB];
        double old[JB][NB];
#pragma unroll
        for (int jj = 0; jj < JB; jj++) {
            if (j0 + jj < NO) {
                const int jc = (j0 + jj) + ((j0 + jj) >= JD ? 1 : 0);
                const uint32_t w = (D.P[jc >> 1] >> (16 * (jc & 1))) & 0xffffu;
                idx[jj] = (int)(w & 0x7fffu) * (NB * NT);
#ifdef FB_FIRST_TOUCH
                const bool first = (w & 0x8000u) != 0;
#else
                const bool first = false;
#endif
#pragma unroll
                for (int b = 0; b < NB; b++) old[jj][b] = first ? 0.0 : my[idx[jj] + b * NT];
            }
        }
#pragma unroll
        for (int jj = 0; jj < JB; jj++) {
            if (j0 + jj < NO) {
                const int jc = (j0 + jj) + ((j0 + jj) >= JD ? 1 : 0);
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    double v = old[jj][b];
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        v += A.R.r[TYPE][jc][s][0] * E[s][canon_sv<DIM>(jc, 0)][b];
                        if (jc >= NVTX) v += A.R.r[TYPE][jc][s][1] * E[s][canon_sv<DIM>(jc, 1)][b];
                    }
                    my[idx[jj] + b * NT] = v;
                }
            }
        }
    }
}

// OPG: 0 Laplace (one thread per row node), 1 elasticity (DIM threads per row node: one per row dof)
// TYPE: 0 vertex rows, 1 edge rows
#ifndef FB_GATHER_MINBLOCKS
#define FB_GATHER_MINBLOCKS 3
#endif
template <int OPG, int DIM, int NL, int TYPE>
__global__ void __launch_bounds__(128, FB_GATHER_MINBLOCKS) k_gather(const GatherArgs A)
{
    constexpr int TPR = OPG == 1 ? DIM : 1; // threads per row node
    constexpr int NB = OPG == 1 ? DIM : 1;  // accumulators per (thread, column node)
    extern __shared__ double acc[];         // [NB*lcap][blockDim.x], lane-private banks
    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int64_t t = blockIdx.x * (int64_t)NT + tid;
    if (t >= A.count * TPR) return;
    const int64_t rloc = t / TPR;
    const int a = (int)(t - rloc * TPR);
    int64_t base, k0;
    int L, ninc, flags;
    {
        double raw[4];
        ld_v4(reinterpret_cast<const double *>(A.rowinfo + A.start + rloc), raw);
        base = __double_as_longlong(raw[0]);
        k0 = __double_as_longlong(raw[1]);
        const int64_t ln = __double_as_longlong(raw[2]);
        L = (int)(ln & 0xffffffff);
        ninc = (int)(ln >> 32);
        flags = (int)__double_as_longlong(raw[3]);
    }
    double *my = acc + tid;
    constexpr int GS = GeomStride<DIM>::value;
    constexpr int PW = PoscStride<NL>::value / 2;

    if (ninc > 0) {
        // two-deep software pipeline on ping-pong register buffers plus an L2 prefetch three incidences
        // ahead; prefetch indices are clamped to the row's last incidence so the loads are unconditional
        IncData<DIM, NL> bufA, bufB;
        const int64_t k1 = k0 + ninc, kl = k1 - 1;
        const uint32_t c0 = A.incp[k0];
        uint32_t code_n = A.incp[k0 + 1 < kl ? k0 + 1 : kl];
        uint32_t code_f = A.incp[k0 + 3 < kl ? k0 + 3 : kl];
        load_inc<DIM, NL>(A, k0, c0, bufA);
#ifdef FB_FIRST_TOUCH
        if (flags & 2)
#endif
            for (int p = 0; p < NB * L; p++) my[p * NT] = 0.0;
        double dacc[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) dacc[b] = 0.0;
        // position of the row node in its own row (same in every incidence)
        constexpr int JD = TYPE == 0 ? 0 : DIM + 1;
        const int pdiag = (int)((bufA.P[JD >> 1] >> (16 * (JD & 1))) & 0x7fffu) * (NB * NT);
        for (int64_t k = k0; k < k1; k += 2) {
            const int64_t kb = k + 1 < kl ? k + 1 : kl;
            load_inc<DIM, NL>(A, kb, code_n, bufB);
            code_n = A.incp[k + 2 < kl ? k + 2 : kl];
            prefetch_l2(A.geom + (int64_t)(code_f >> 8) * GS);
            prefetch_l2(reinterpret_cast<const uint32_t *>(A.posc) + (k + 3 < kl ? k + 3 : kl) * PW);
            code_f = A.incp[k + 4 < kl ? k + 4 : kl];
            gather_incidence<OPG, DIM, NL, TYPE>(A, bufA, a, my, NT, dacc);
            if (k + 1 < k1) {
                const int64_t ka = k + 2 < kl ? k + 2 : kl;
                load_inc<DIM, NL>(A, ka, code_n, bufA);
                code_n = A.incp[k + 3 < kl ? k + 3 : kl];
                prefetch_l2(A.geom + (int64_t)(code_f >> 8) * GS);
                prefetch_l2(reinterpret_cast<const uint32_t *>(A.posc) + (k + 4 < kl ? k + 4 : kl) * PW);
                code_f = A.incp[k + 5 < kl ? k + 5 : kl];
                gather_incidence<OPG, DIM, NL, TYPE>(A, bufB, a, my, NT, dacc);
            }
        }
#pragma unroll
        for (int b = 0; b < NB; b++) my[pdiag + b * NT] = dacc[b];
    } else {
        for (int p = 0; p < NB * L; p++) my[p * NT] = 0.0;
    }

    // write-out: the thread's accumulators are exactly one CSR row (scalar / elasticity dof row (I, a)),
    // contiguous in the values array; the block-diagonal vector Laplacian replicates it DIM times
    const int nrep = (OPG == 0 && A.vec_dim != 0) ? A.vec_dim : 1;
    const int n = NB * L;
    for (int d = 0; d < nrep; d++) {
        double *out = OPG == 1 ? A.values + (int64_t)DIM * DIM * base + (int64_t)a * n
                               : A.values + (int64_t)nrep * base + (int64_t)d * L;
        int p = 0;
        while (p < n && (reinterpret_cast<uintptr_t>(out + p) & 31)) { out[p] = my[p * NT]; p++; }
        for (; p + 4 <= n; p += 4) st_v4(out + p, my[p * NT], my[(p + 1) * NT], my[(p + 2) * NT], my[(p + 3) * NT]);
        for (; p < n; p++) out[p] = my[p * NT];
    }
}

} // namespace fb
