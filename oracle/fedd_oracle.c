/*
 * fedd_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the FEDDLib element-wise assembly hot path, written to follow
 * the reference loop order and floating-point expression order.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY STATUS: the reference ships no golden matrices, known-answer tests or expected norms for
 * this path, so the restatement is pinned by outputs of the reference itself run here:
 * oracle/_ref (`make -C oracle ref`) compiles FEDDLib's own FE_def.hpp hot-path routines, sliced from
 * /root/reference at build time, against mock Trilinos containers; tests/test_oracle_vs_ref.py requires
 * BITWISE equality of every operator, table and quadrature degree between the two.  Analytic
 * known-answer tests (tests/test_oracle_kat.py) pin both from the other side.
 * The one thing neither can pin is Tpetra's own insert/fillComplete (Trilinos is absent, version
 * unpinned by the reference): it is emulated below as append -> stable sort -> sequential sum.
 *
 * Reference lines restated (all under /root/reference/feddlib/core):
 *   FE/FE_def.hpp:83-112     applyBTinv
 *   FE/FE_def.hpp:604-667    assemblyLaplace
 *   FE/FE_def.hpp:670-734    assemblyLaplaceVecField
 *   FE/FE_def.hpp:1685-1836  assemblyAdvectionVecField (live branch 1759-1832)
 *   FE/FE_def.hpp:1839-1929  assemblyAdvectionInUVecField
 *   FE/FE_def.hpp:1932-2057  assemblyDivAndDivT
 *   FE/FE_def.hpp:2061-2148  assemblyDivAndDivTFast
 *   FE/FE_def.hpp:2739-3040  assemblyLinElasXDim
 *   FE/FE_def.hpp:4931-4944  epsilonTensor
 *   FE/FE_def.hpp:4947-5088  phi (P1/P2, 2D/3D)
 *   FE/FE_def.hpp:5342-5368  buildTransformation
 *   FE/FE_def.hpp:5431-5562  determineDegree (3 overloads)
 *   FE/FE_def.hpp:5565-5714  gradPhi (P1/P2, 2D/3D)
 *   FE/FE_def.hpp:6023-6460  getQuadratureValues (2D deg 1,2,5; 3D deg 1,3,5)
 *   FE/FE_def.hpp:6730-6929  getPhi / getDPhi
 *   General/SmallMatrix.hpp:164-215, 306-358  innerProduct, trace, computeInverse, computeDet
 *   LinearAlgebra/Matrix_def.hpp:46-51,88-92,192-199  Matrix ctor / insertGlobalValues / fillComplete
 *     (the accumulate itself lives in Trilinos Tpetra, version unpinned by the reference:
 *      append on insert, per-row sort by column + duplicate summation on fillComplete;
 *      restated here as append -> stable sort by column -> sequential sum)
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off: no FMA contraction, so the
 *                              expression order written here is the order evaluated)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_MAXQ 16   /* max quadrature points used on this path (15-pt tet rule) */
#define FO_MAXN 10   /* max local nodes (P2 tet) */

/* ------------------------------------------------------------------------------------ */
/* Matrix emulation: Matrix(map,numEntries) + insertGlobalValues + fillComplete          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int64_t *cols;
    double  *vals;
    int32_t  n, cap;
} fo_row;

typedef struct fo_matrix {
    int64_t nrows;
    int32_t cap_hint;
    fo_row *rows;
    int     filled;
} fo_matrix;

fo_matrix *fo_matrix_new(int64_t nrows, int32_t cap_hint)
{
    fo_matrix *A = (fo_matrix *)calloc(1, sizeof(fo_matrix));
    A->nrows = nrows;
    A->cap_hint = cap_hint > 0 ? cap_hint : 8;
    A->rows = (fo_row *)calloc((size_t)(nrows > 0 ? nrows : 1), sizeof(fo_row));
    return A;
}

void fo_matrix_free(fo_matrix *A)
{
    if (!A) return;
    for (int64_t r = 0; r < A->nrows; r++) { free(A->rows[r].cols); free(A->rows[r].vals); }
    free(A->rows);
    free(A);
}

/* Matrix_def.hpp:88-92 -> Xpetra/Tpetra insertGlobalValues: append, duplicates kept. */
int fo_insert(fo_matrix *A, int64_t row, int32_t n, const int64_t *cols, const double *vals)
{
    if (row < 0 || row >= A->nrows) return -1;
    fo_row *R = &A->rows[row];
    if (R->n + n > R->cap) {
        int32_t nc = R->cap ? R->cap : A->cap_hint;
        while (nc < R->n + n) nc *= 2;
        R->cols = (int64_t *)realloc(R->cols, (size_t)nc * sizeof(int64_t));
        R->vals = (double *)realloc(R->vals, (size_t)nc * sizeof(double));
        R->cap = nc;
    }
    memcpy(R->cols + R->n, cols, (size_t)n * sizeof(int64_t));
    memcpy(R->vals + R->n, vals, (size_t)n * sizeof(double));
    R->n += n;
    return 0;
}

/* stable merge sort of (col,val) pairs by col */
static void fo_sort_row(int64_t *c, double *v, int32_t n, int64_t *tc, double *tv)
{
    for (int32_t w = 1; w < n; w *= 2) {
        for (int32_t lo = 0; lo < n; lo += 2 * w) {
            int32_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int32_t a = lo, b = mid, k = lo;
            while (a < mid && b < hi) {
                if (c[b] < c[a]) { tc[k] = c[b]; tv[k++] = v[b++]; }
                else             { tc[k] = c[a]; tv[k++] = v[a++]; }
            }
            while (a < mid) { tc[k] = c[a]; tv[k++] = v[a++]; }
            while (b < hi)  { tc[k] = c[b]; tv[k++] = v[b++]; }
        }
        memcpy(c, tc, (size_t)n * sizeof(int64_t));
        memcpy(v, tv, (size_t)n * sizeof(double));
    }
}

/* Matrix_def.hpp:192-199 -> fillComplete: per row sort by column, merge duplicates by
 * summation in insertion order (stable).  Returns nnz. */
int64_t fo_fill_complete(fo_matrix *A)
{
    int32_t maxn = 0;
    for (int64_t r = 0; r < A->nrows; r++) if (A->rows[r].n > maxn) maxn = A->rows[r].n;
    int64_t *tc = (int64_t *)malloc((size_t)(maxn + 1) * sizeof(int64_t));
    double  *tv = (double *)malloc((size_t)(maxn + 1) * sizeof(double));
    int64_t nnz = 0;
    for (int64_t r = 0; r < A->nrows; r++) {
        fo_row *R = &A->rows[r];
        if (R->n == 0) continue;
        fo_sort_row(R->cols, R->vals, R->n, tc, tv);
        int32_t k = 0;
        for (int32_t i = 1; i < R->n; i++) {
            if (R->cols[i] == R->cols[k]) R->vals[k] += R->vals[i];
            else { k++; R->cols[k] = R->cols[i]; R->vals[k] = R->vals[i]; }
        }
        R->n = k + 1;
        nnz += R->n;
    }
    free(tc); free(tv);
    A->filled = 1;
    return nnz;
}

int64_t fo_matrix_nnz(const fo_matrix *A)
{
    int64_t nnz = 0;
    for (int64_t r = 0; r < A->nrows; r++) nnz += A->rows[r].n;
    return nnz;
}

/* rowptr[nrows+1], colgid[nnz], vals[nnz] (global column ids) */
void fo_get_csr(const fo_matrix *A, int64_t *rowptr, int64_t *colgid, double *vals)
{
    int64_t p = 0;
    for (int64_t r = 0; r < A->nrows; r++) {
        rowptr[r] = p;
        memcpy(colgid + p, A->rows[r].cols, (size_t)A->rows[r].n * sizeof(int64_t));
        memcpy(vals + p, A->rows[r].vals, (size_t)A->rows[r].n * sizeof(double));
        p += A->rows[r].n;
    }
    rowptr[A->nrows] = p;
}

/* Matrix_def.hpp:257 scale (used by the problem classes after assembly) */
void fo_matrix_scale(fo_matrix *A, double s)
{
    for (int64_t r = 0; r < A->nrows; r++)
        for (int32_t i = 0; i < A->rows[r].n; i++) A->rows[r].vals[i] *= s;
}

/* ------------------------------------------------------------------------------------ */
/* FE type helpers                                                                        */
/* ------------------------------------------------------------------------------------ */
enum { FO_STD = 0, FO_GRAD = 1 };

static int fo_is(const char *a, const char *b) { return strcmp(a, b) == 0; }

/* nodes per element; -1 if unsupported on this path */
int fo_nloc(int dim, const char *fe)
{
    if (fo_is(fe, "P0")) return dim == 2 ? 1 : -1;   /* FE_def.hpp:6766-6769: one local "point", intFE = 0 (2D; see fo_phi) */
    if (fo_is(fe, "P1")) return dim + 1;
    if (fo_is(fe, "P2")) return dim == 2 ? 6 : (dim == 3 ? 10 : -1);
    return -1;
}
static int fo_intfe(const char *fe) { return fo_is(fe, "P0") ? 0 : (fo_is(fe, "P1") ? 1 : (fo_is(fe, "P2") ? 2 : -1)); }

/* FE_def.hpp:5431-5512 */
int fo_determine_degree2(int dim, const char *fe1, const char *fe2, int t1, int t2, int extra)
{
    (void)dim;
    int d1 = 0, d2 = 0;
    if (fo_is(fe1, "P0")) d1 = 0;
    else if (fo_is(fe1, "P1")) d1 = (t1 == FO_STD) ? 1 : 0;
    else if (fo_is(fe1, "P2")) d1 = (t1 == FO_STD) ? 2 : 1;
    if (fo_is(fe2, "P0")) d2 = 0;
    else if (fo_is(fe2, "P1")) d2 = (t2 == FO_STD) ? 1 : 0;
    else if (fo_is(fe2, "P2")) d2 = (t2 == FO_STD) ? 2 : 1;
    int deg = d1 + d2 + extra;
    if (deg == 0) deg = 1;
    return deg;
}
/* FE_def.hpp:5515-5543 */
int fo_determine_degree1(int dim, const char *fe, int t)
{
    (void)dim;
    int deg = 0;
    if (fo_is(fe, "P0")) deg = 0;
    else if (fo_is(fe, "P1")) deg = (t == FO_STD) ? 1 : 0;
    else if (fo_is(fe, "P2")) deg = (t == FO_STD) ? 2 : 1;
    if (deg == 0) deg = 1;
    return deg;
}

/* FE_def.hpp:6023-6460.  pts[n][dim], w[n]; returns n or -1. */
int fo_quadrature(int dim, int deg, double *pts, double *w)
{
    if (dim == 2) {
        if (deg == 3 || deg == 4) deg = 5;
        if (deg == 6) deg = 7;
        switch (deg) {
        case 1:
            pts[0] = 1 / 3.; pts[1] = 1 / 3.; w[0] = 1 / 2.;
            return 1;
        case 2: {
            double a = 1 / 6.;
            pts[0] = 0.5; pts[1] = 0.5;
            pts[2] = 0.;  pts[3] = 0.5;
            pts[4] = 0.5; pts[5] = 0.;
            w[0] = a; w[1] = a; w[2] = a;
            return 3;
        }
        case 5: {
            double a = 0.470142064105115, b = 0.101286507323456;
            double P1 = 0.066197076394253, P2 = 0.062969590272413;
            pts[0] = 1 / 3.;       pts[1] = 1 / 3.;
            pts[2] = a;            pts[3] = a;
            pts[4] = 1 - 2. * a;   pts[5] = a;
            pts[6] = a;            pts[7] = 1 - 2. * a;
            pts[8] = b;            pts[9] = b;
            pts[10] = 1 - 2. * b;  pts[11] = b;
            pts[12] = b;           pts[13] = 1 - 2. * b;
            w[0] = 9 / 80.;
            w[1] = P1; w[2] = P1; w[3] = P1;
            w[4] = P2; w[5] = P2; w[6] = P2;
            return 7;
        }
        default: return -1; /* 28-pt degree-7 rule is never reached on this path */
        }
    }
    if (dim == 3) {
        if (deg == 2) deg = 3;
        if (deg == 4) deg = 5;
        switch (deg) {
        case 1:
            pts[0] = 0.25; pts[1] = 0.25; pts[2] = 0.25; w[0] = 1 / 6.;
            return 1;
        case 3: {
            double a = .25, b = 1. / 6., c = .5;
            double P[5][3] = {{a, a, a}, {b, b, b}, {b, b, c}, {b, c, b}, {c, b, b}};
            memcpy(pts, P, sizeof(P));
            w[0] = -2. / 15.;
            w[1] = 3. / 40.; w[2] = 3. / 40.; w[3] = 3. / 40.; w[4] = 3. / 40.;
            return 5;
        }
        case 5: {
            double a = 0.25;
            double b1 = (7. + sqrt(15.)) / 34., b2 = (7. - sqrt(15.)) / 34.;
            double c1 = (13. - 3. * sqrt(15.)) / 34., c2 = (13. + 3. * sqrt(15.)) / 34.;
            double d = (5. - sqrt(15.)) / 20., e = (5. + sqrt(15.)) / 20.;
            double P[15][3] = {
                {a, a, a},
                {b1, b1, b1}, {b1, b1, c1}, {b1, c1, b1}, {c1, b1, b1},
                {b2, b2, b2}, {b2, b2, c2}, {b2, c2, b2}, {c2, b2, b2},
                {d, d, e}, {d, e, d}, {e, d, d}, {d, e, e}, {e, d, e}, {e, e, d}};
            memcpy(pts, P, sizeof(P));
            double P1 = (2665. - 14. * sqrt(15.)) / 226800.;
            double P2 = (2665. + 14. * sqrt(15.)) / 226800.;
            double b = 5. / 567.;
            w[0] = 8. / 405.;
            w[1] = P1; w[2] = P1; w[3] = P1; w[4] = P1;
            w[5] = P2; w[6] = P2; w[7] = P2; w[8] = P2;
            for (int k = 9; k < 15; k++) w[k] = b;
            return 15;
        }
        default: return -1;
        }
    }
    return -1;
}

/* FE_def.hpp:4991-5088 (P1/P2 cases) */
double fo_phi(int dim, int intFE, int i, const double *p)
{
    /* P0: `case 0: //P0` of FE::phi exists for dim 1 and 2 only (FE_def.hpp:4955, 4993); in 3D the reference leaves *value
     * uninitialised (no case 0 below :5037), so P0 is a 2D feature -- the callers reject it for dim 3 */
    if (intFE == 0 && i == 0 && dim == 2) return 1.;
    if (dim == 2) {
        if (intFE == 1) {
            switch (i) {
            case 0: return (1. - p[0] - p[1]);
            case 1: return p[0];
            case 2: return p[1];
            }
        } else if (intFE == 2) {
            switch (i) {
            case 0: return -(1. - p[0] - p[1]) * (1 - 2. * (1 - p[0] - p[1]));
            case 1: return -p[0] * (1 - 2 * p[0]);
            case 2: return -p[1] * (1 - 2 * p[1]);
            case 3: return 4 * p[0] * (1 - p[0] - p[1]);
            case 4: return 4 * p[0] * p[1];
            case 5: return 4 * p[1] * (1 - p[0] - p[1]);
            }
        }
    } else if (dim == 3) {
        if (intFE == 1) {
            switch (i) {
            case 0: return (1. - p[0] - p[1] - p[2]);
            case 1: return p[0];
            case 2: return p[1];
            case 3: return p[2];
            }
        } else if (intFE == 2) {
            switch (i) {
            case 0: return (1. - p[0] - p[1] - p[2]) * (1 - 2 * p[0] - 2 * p[1] - 2 * p[2]);
            case 1: return p[0] * (2 * p[0] - 1);
            case 2: return p[1] * (2 * p[1] - 1);
            case 3: return p[2] * (2 * p[2] - 1);
            case 4: return 4 * p[0] * (1 - p[0] - p[1] - p[2]);
            case 5: return 4 * p[0] * p[1];
            case 6: return 4 * p[1] * (1 - p[0] - p[1] - p[2]);
            case 7: return 4 * p[2] * (1 - p[0] - p[1] - p[2]);
            case 8: return 4 * p[0] * p[2];
            case 9: return 4 * p[1] * p[2];
            }
        }
    }
    return NAN;
}

/* FE_def.hpp:5570-5714 (P1/P2 cases) */
void fo_grad_phi(int dim, int intFE, int i, const double *p, double *v)
{
    if (dim == 2) {
        if (intFE == 1) {
            switch (i) {
            case 0: v[0] = -1.; v[1] = -1.; break;
            case 1: v[0] = 1.;  v[1] = 0.;  break;
            case 2: v[0] = 0.;  v[1] = 1.;  break;
            }
        } else {
            switch (i) {
            case 0: v[0] = 1. - 4. * (1 - p[0] - p[1]); v[1] = 1. - 4. * (1 - p[0] - p[1]); break;
            case 1: v[0] = 4. * p[0] - 1; v[1] = 0.; break;
            case 2: v[0] = 0.; v[1] = 4. * p[1] - 1; break;
            case 3: v[0] = 4 * (1. - 2 * p[0] - p[1]); v[1] = -4 * p[0]; break;
            case 4: v[0] = 4. * p[1]; v[1] = 4. * p[0]; break;
            case 5: v[0] = -4. * p[1]; v[1] = 4 * (1. - p[0] - 2 * p[1]); break;
            }
        }
    } else {
        if (intFE == 1) {
            switch (i) {
            case 0: v[0] = -1.; v[1] = -1.; v[2] = -1.; break;
            case 1: v[0] = 1.;  v[1] = 0.;  v[2] = 0.;  break;
            case 2: v[0] = 0.;  v[1] = 1.;  v[2] = 0.;  break;
            case 3: v[0] = 0.;  v[1] = 0.;  v[2] = 1.;  break;
            }
        } else {
            switch (i) {
            case 0:
                v[0] = -3. + 4. * p[0] + 4. * p[1] + 4. * p[2];
                v[1] = -3. + 4. * p[0] + 4. * p[1] + 4. * p[2];
                v[2] = -3. + 4. * p[0] + 4. * p[1] + 4. * p[2];
                break;
            case 1: v[0] = 4. * p[0] - 1; v[1] = 0.; v[2] = 0.; break;
            case 2: v[0] = 0.; v[1] = 4. * p[1] - 1; v[2] = 0.; break;
            case 3: v[0] = 0.; v[1] = 0.; v[2] = 4. * p[2] - 1; break;
            case 4: v[0] = 4. - 8. * p[0] - 4. * p[1] - 4. * p[2]; v[1] = -4. * p[0]; v[2] = -4. * p[0]; break;
            case 5: v[0] = 4. * p[1]; v[1] = 4. * p[0]; v[2] = 0.; break;
            case 6: v[0] = -4. * p[1]; v[1] = 4. - 4. * p[0] - 8. * p[1] - 4. * p[2]; v[2] = -4. * p[1]; break;
            case 7: v[0] = -4. * p[2]; v[1] = -4. * p[2]; v[2] = 4. - 4. * p[0] - 4. * p[1] - 8. * p[2]; break;
            case 8: v[0] = 4. * p[2]; v[1] = 0.; v[2] = 4. * p[0]; break;
            case 9: v[0] = 0.; v[1] = 4. * p[2]; v[2] = 4. * p[1]; break;
            }
        }
    }
}

/* FE_def.hpp:6730-6836: phi[nq][nloc], weights[nq]; returns nq */
int fo_get_phi(int dim, const char *fe, int deg, double *phi, double *w)
{
    double pts[FO_MAXQ * 3];
    int nq = fo_quadrature(dim, deg, pts, w);
    int nloc = fo_nloc(dim, fe), intFE = fo_intfe(fe);
    if (nq < 0 || nloc < 0) return -1;
    for (int k = 0; k < nq; k++)
        for (int i = 0; i < nloc; i++) phi[k * nloc + i] = fo_phi(dim, intFE, i, pts + k * dim);
    return nq;
}

/* FE_def.hpp:6846-6929: dphi[nq][nloc][dim], weights[nq]; returns nq */
int fo_get_dphi(int dim, const char *fe, int deg, double *dphi, double *w)
{
    double pts[FO_MAXQ * 3];
    int nq = fo_quadrature(dim, deg, pts, w);
    int nloc = fo_nloc(dim, fe), intFE = fo_intfe(fe);
    if (nq < 0 || nloc < 0) return -1;
    for (int k = 0; k < nq; k++)
        for (int i = 0; i < nloc; i++) fo_grad_phi(dim, intFE, i, pts + k * dim, dphi + (k * nloc + i) * dim);
    return nq;
}

/* ------------------------------------------------------------------------------------ */
/* SmallMatrix + transformation                                                           */
/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:5342-5357: B[i][j] = x_{elem[j+1]}[i] - x_{elem[0]}[i] */
static void fo_build_transformation(int dim, const int32_t *el, const double *coords, double B[3][3])
{
    int32_t i0 = el[0];
    for (int j = 0; j < dim; j++) {
        int32_t idx = el[j + 1];
        for (int i = 0; i < dim; i++) B[i][j] = coords[(int64_t)idx * dim + i] - coords[(int64_t)i0 * dim + i];
    }
}

/* SmallMatrix.hpp:338-358 */
static double fo_det(int dim, double v[3][3])
{
    if (dim == 2) return v[0][0] * v[1][1] - v[1][0] * v[0][1];
    return v[0][0] * v[1][1] * v[2][2] +
           v[0][1] * v[1][2] * v[2][0] +
           v[0][2] * v[1][0] * v[2][1] -
           v[2][0] * v[1][1] * v[0][2] -
           v[2][1] * v[1][2] * v[0][0] -
           v[2][2] * v[1][0] * v[0][1];
}

/* SmallMatrix.hpp:306-335 */
static double fo_inverse(int dim, double v[3][3], double inv[3][3])
{
    double det = fo_det(dim, v);
    if (dim == 2) {
        inv[0][0] = v[1][1] / det;
        inv[0][1] = (-v[0][1]) / det;
        inv[1][0] = (-v[1][0]) / det;
        inv[1][1] = v[0][0] / det;
    } else {
        inv[0][0] = (v[1][1] * v[2][2] - v[1][2] * v[2][1]) / det;
        inv[0][1] = (v[0][2] * v[2][1] - v[0][1] * v[2][2]) / det;
        inv[0][2] = (v[0][1] * v[1][2] - v[0][2] * v[1][1]) / det;
        inv[1][0] = (v[1][2] * v[2][0] - v[1][0] * v[2][2]) / det;
        inv[1][1] = (v[0][0] * v[2][2] - v[0][2] * v[2][0]) / det;
        inv[1][2] = (v[0][2] * v[1][0] - v[0][0] * v[1][2]) / det;
        inv[2][0] = (v[1][0] * v[2][1] - v[1][1] * v[2][0]) / det;
        inv[2][1] = (v[0][1] * v[2][0] - v[0][0] * v[2][1]) / det;
        inv[2][2] = (v[0][0] * v[1][1] - v[0][1] * v[1][0]) / det;
    }
    return det;
}

/* FE_def.hpp:83-96: out[w][i][d1] += in[w][i][d2] * Binv[d2][d1]  (out pre-zeroed) */
static void fo_apply_btinv(int dim, int nq, int nloc, const double *in, double *out, double Binv[3][3])
{
    for (int w = 0; w < nq; w++)
        for (int i = 0; i < nloc; i++)
            for (int d1 = 0; d1 < dim; d1++) {
                double acc = 0.;
                for (int d2 = 0; d2 < dim; d2++) acc += in[(w * nloc + i) * dim + d2] * Binv[d2][d1];
                out[(w * nloc + i) * dim + d1] = acc;
            }
}

/* per-element geometry shared by all routines: B, Binv, |det|, transformed gradients */
static double fo_element_geometry(int dim, int nq, int nloc, const int32_t *el, const double *coords,
                                  const double *dphi, double *dphiT)
{
    double B[3][3], Binv[3][3];
    fo_build_transformation(dim, el, coords, B);
    double detB = fo_inverse(dim, B, Binv);
    fo_apply_btinv(dim, nq, nloc, dphi, dphiT, Binv);
    return fabs(detB);
}

/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:604-667 assemblyLaplace                                                     */
/* ------------------------------------------------------------------------------------ */
int fo_assembly_laplace(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                        const int64_t *gid, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int deg = fo_determine_degree2(dim, fe, fe, FO_GRAD, FO_GRAD, 0);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0) return -1;
    double value[FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                value[j] = 0.;
                for (int q = 0; q < nq; q++)
                    for (int d = 0; d < dim; d++)
                        value[j] += w[q] * dT[(q * nloc + i) * dim + d] * dT[(q * nloc + j) * dim + d];
                value[j] *= absDetB;
                indices[j] = gid[el[j]];
            }
            fo_insert(A, gid[el[i]], nloc, indices, value);
        }
    }
    return 0;
}

/* FE_def.hpp:670-734 assemblyLaplaceVecField */
int fo_assembly_laplace_vecfield(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                                 const int64_t *gid, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int deg = fo_determine_degree2(dim, fe, fe, FO_GRAD, FO_GRAD, 0);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0) return -1;
    double value[FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                value[j] = 0.;
                for (int q = 0; q < nq; q++)
                    for (int d = 0; d < dim; d++)
                        value[j] += w[q] * dT[(q * nloc + i) * dim + d] * dT[(q * nloc + j) * dim + d];
                value[j] *= absDetB;
            }
            for (int d = 0; d < dim; d++) {
                for (int j = 0; j < nloc; j++) indices[j] = dim * gid[el[j]] + d;
                fo_insert(A, dim * gid[el[i]] + d, nloc, indices, value);
            }
        }
    }
    return 0;
}

/* FE_def.hpp:4694-4766 assemblyRHS: constant source term f (the reference evaluates func once, :4735), added into the
 * repeated vector with LOCAL node ids (:4751, :4758).  deg = determineDegree(FEType, Std) + degFunc (:4716-4717). */
int fo_assembly_rhs(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                    int vec_field, int deg_func, const double *value_func, double *rhs)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ];
    int deg = fo_determine_degree1(dim, fe, FO_STD) + deg_func;
    int nq = fo_get_phi(dim, fe, deg, phi, w);
    if (nq < 0) return -1;
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double B[3][3];
        fo_build_transformation(dim, el, coords, B);
        double absDetB = fabs(fo_det(dim, B));
        for (int i = 0; i < nloc; i++) {
            double value = 0.;
            for (int q = 0; q < nq; q++) value += w[q] * phi[q * nloc + i];
            if (!vec_field) {
                value *= absDetB * value_func[0];
                rhs[el[i]] += value;
            } else {
                value *= absDetB;
                for (int d = 0; d < dim; d++) rhs[(int64_t)dim * el[i] + d] += value * value_func[d];
            }
        }
    }
    return 0;
}

/* FE_def.hpp:454-521 assemblyMass (fieldType "Scalar": vec_field = 0, "Vector": vec_field = 1).
 * deg = determineDegree(Std, Std) (:474); only |det B| of the affine map is used (:489-491). */
int fo_assembly_mass(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                     const int64_t *gid, int vec_field, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ];
    int deg = fo_determine_degree2(dim, fe, fe, FO_STD, FO_STD, 0);
    int nq = fo_get_phi(dim, fe, deg, phi, w);
    if (nq < 0) return -1;
    double value[FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double B[3][3];
        fo_build_transformation(dim, el, coords, B);
        double absDetB = fabs(fo_det(dim, B));
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                value[j] = 0.;
                for (int q = 0; q < nq; q++) value[j] += w[q] * phi[q * nloc + i] * phi[q * nloc + j];
                value[j] *= absDetB;
                if (!vec_field) indices[j] = gid[el[j]];
            }
            if (!vec_field) fo_insert(A, gid[el[i]], nloc, indices, value);
            else
                for (int d = 0; d < dim; d++) {
                    for (int j = 0; j < nloc; j++) indices[j] = dim * gid[el[j]] + d;
                    fo_insert(A, dim * gid[el[i]] + d, nloc, indices, value);
                }
        }
    }
    return 0;
}

/* FE_def.hpp:2151-2220 assemblyBDStabilization (P1 only, :2156): the P1 mass entry minus the mean-value term,
 *   value[j] = (sum_w weights[w] phi[w][i] phi[w][j]) * absDetB - refElementSize * absDetB * refElementScale   (:2203-2207)
 * with refElementSize / refElementScale = 1/2, 1/9 (2D) and 1/6, 1/16 (3D) (:2183-2190); deg = determineDegree(Std, Std). */
int fo_assembly_bdstab(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                       const int64_t *gid, fo_matrix *A)
{
    if (strcmp(fe, "P1") != 0) return -1; /* "Only implemented for P1." */
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ];
    int deg = fo_determine_degree2(dim, fe, fe, FO_STD, FO_STD, 0);
    int nq = fo_get_phi(dim, fe, deg, phi, w);
    if (nq < 0) return -1;
    double refElementSize, refElementScale;
    if (dim == 2) { refElementSize = 0.5; refElementScale = 1. / 9.; }
    else if (dim == 3) { refElementSize = 1. / 6.; refElementScale = 1. / 16.; }
    else return -1;
    double value[FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double B[3][3];
        fo_build_transformation(dim, el, coords, B);
        double absDetB = fabs(fo_det(dim, B));
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                value[j] = 0.;
                for (int q = 0; q < nq; q++) value[j] += w[q] * phi[q * nloc + i] * phi[q * nloc + j];
                value[j] *= absDetB;
                value[j] -= refElementSize * absDetB * refElementScale;
                indices[j] = gid[el[j]];
            }
            fo_insert(A, gid[el[i]], nloc, indices, value);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:2739-3040 assemblyLinElasXDim                                               */
/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:4931-4944 */
static void fo_epsilon_tensor(int dim, const double *g, double eps[3][3], int activeDof)
{
    for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) {
            eps[i][j] = 0.;
            if (i == activeDof) eps[i][j] += 0.5 * g[j];
            if (j == activeDof) eps[i][j] += 0.5 * g[i];
        }
}
/* SmallMatrix.hpp:164-179 */
static double fo_inner(int dim, double a[3][3], double b[3][3])
{
    if (dim == 2)
        return a[0][0] * b[0][0] + a[0][1] * b[0][1] + a[1][0] * b[1][0] + a[1][1] * b[1][1];
    return a[0][0] * b[0][0] + a[0][1] * b[0][1] + a[0][2] * b[0][2] +
           a[1][0] * b[1][0] + a[1][1] * b[1][1] + a[1][2] * b[1][2] +
           a[2][0] * b[2][0] + a[2][1] * b[2][1] + a[2][2] * b[2][2];
}
/* SmallMatrix.hpp:200-215 */
static double fo_trace(int dim, double a[3][3])
{
    if (dim == 2) return a[0][0] + a[1][1];
    return a[0][0] + a[1][1] + a[2][2];
}

int fo_assembly_linelas(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                        const int64_t *gid, double lambda, double mu, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int deg = fo_determine_degree2(dim, fe, fe, FO_GRAD, FO_GRAD, 0);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0) return -1;
    double epsI[3][3][3], epsJ[3][3][3];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                double v[3][3] = {{0}};
                for (int k = 0; k < nq; k++) {
                    for (int a = 0; a < dim; a++) fo_epsilon_tensor(dim, dT + (k * nloc + i) * dim, epsI[a], a);
                    for (int b = 0; b < dim; b++) fo_epsilon_tensor(dim, dT + (k * nloc + j) * dim, epsJ[b], b);
                    for (int a = 0; a < dim; a++)
                        for (int b = 0; b < dim; b++) {
                            double res = fo_inner(dim, epsI[a], epsJ[b]);
                            double tr_i = fo_trace(dim, epsI[a]);
                            double tr_j = fo_trace(dim, epsJ[b]);
                            v[a][b] = v[a][b] + w[k] * (2 * mu * res + lambda * tr_j * tr_i);
                        }
                }
                for (int a = 0; a < dim; a++)
                    for (int b = 0; b < dim; b++) v[a][b] = absDetB * v[a][b];
                int64_t glob_j = dim * gid[el[j]];
                int64_t glob_i = dim * gid[el[i]];
                /* insertion order of the reference: for each column b, rows a = 0..dim-1,
                 * one entry per call (FE_def.hpp:2877-2885, 3016-3031) */
                for (int b = 0; b < dim; b++) {
                    int64_t col = glob_j + b;
                    for (int a = 0; a < dim; a++) fo_insert(A, glob_i + a, 1, &col, &v[a][b]);
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:2407-2735 assemblyStress: the symmetric-gradient viscous block with a coefficient function,
 *   v_ab += func(xyz_k) * weights[k] * (e_a(phi_i) : e_b(phi_j)),   v_ab = absDetB * v_ab,
 * where e_a(phi_i) has the transformed gradient of phi_i in row a (:2496, :2511; 3D :2624-2626, :2636-2638, :2648-2650)
 * and e_b(phi_j) is that row plus its transpose (:2485-2494, :2499-2509; 3D :2618-2622, :2629-2633, :2641-2645);
 * innerProduct sums all entries row by row (SmallMatrix.hpp:164-190).  xyz_k = B * quadPts[k] + p1 (:2479-2484,
 * :2599-2608).  func is evaluated once per (i, j, k) in the reference; it is a pure function of xyz.
 * Insertion: one entry per call, for each column dof b the rows a = 0..dim-1 (:2547-2556, :2706-2723). */
typedef double (*fo_coeff_func)(const double *xyz, void *user);
int fo_assembly_stress(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                       const int64_t *gid, fo_coeff_func func, void *user, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3], pts[FO_MAXQ * 3], wq[FO_MAXQ];
    int deg = fo_determine_degree2(dim, fe, fe, FO_GRAD, FO_GRAD, 0);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0 || fo_quadrature(dim, deg, pts, wq) != nq) return -1;
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double B[3][3];
        fo_build_transformation(dim, el, coords, B);
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        const double *p1 = coords + (int64_t)el[0] * dim;
        double fv[FO_MAXQ];
        for (int k = 0; k < nq; k++) {
            double xyz[3] = {0., 0., 0.};
            for (int r = 0; r < dim; r++)
                for (int c = 0; c < dim; c++) xyz[c] += B[c][r] * pts[k * dim + r];
            for (int c = 0; c < dim; c++) xyz[c] += p1[c];
            fv[k] = func(xyz, user);
        }
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                double v[3][3] = {{0}};
                for (int k = 0; k < nq; k++) {
                    const double *gi = dT + (k * nloc + i) * dim, *gj = dT + (k * nloc + j) * dim;
                    double ei[3][3][3], ej[3][3][3];
                    memset(ei, 0, sizeof(ei)); memset(ej, 0, sizeof(ej));
                    for (int a = 0; a < dim; a++)
                        for (int c = 0; c < dim; c++) {
                            ei[a][a][c] = gi[c];                         /* e_a i: row a = grad phi_i */
                            if (dim == 2) {                              /* tmpRes1 + tmpRes2 (:2485-2509) */
                                ej[a][a][c] += gj[c];
                                ej[a][c][a] += gj[c];
                            } else {                                     /* explicit entries (:2618-2645) */
                                ej[a][a][c] = (c == a) ? 2. * gj[c] : gj[c];
                                ej[a][c][a] = (c == a) ? 2. * gj[c] : gj[c];
                            }
                        }
                    for (int a = 0; a < dim; a++)
                        for (int b = 0; b < dim; b++) v[a][b] = v[a][b] + fv[k] * w[k] * fo_inner(dim, ei[a], ej[b]);
                }
                for (int a = 0; a < dim; a++)
                    for (int b = 0; b < dim; b++) v[a][b] = absDetB * v[a][b];
                int64_t glob_j = dim * gid[el[j]];
                int64_t glob_i = dim * gid[el[i]];
                for (int b = 0; b < dim; b++) {
                    int64_t col = glob_j + b;
                    for (int a = 0; a < dim; a++) fo_insert(A, glob_i + a, 1, &col, &v[a][b]);
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:1759-1832 assemblyAdvectionVecField (N); u is node-wise interleaved on the   */
/* repeated map and indexed with LOCAL ids: u[dim*node + d]                                */
/* ------------------------------------------------------------------------------------ */
int fo_assembly_advection(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                          const int64_t *gid, const double *u, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int extraDeg = fo_determine_degree1(dim, fe, FO_STD);
    int deg = fo_determine_degree2(dim, fe, fe, FO_GRAD, FO_STD, extraDeg);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0 || fo_get_phi(dim, fe, deg, phi, w) != nq) return -1;
    double uLoc[3][FO_MAXQ], value[FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        for (int q = 0; q < nq; q++)
            for (int d = 0; d < dim; d++) {
                uLoc[d][q] = 0.;
                for (int i = 0; i < nloc; i++) uLoc[d][q] += u[(int64_t)dim * el[i] + d] * phi[q * nloc + i];
            }
        for (int i = 0; i < nloc; i++) {
            for (int j = 0; j < nloc; j++) {
                value[j] = 0.;
                for (int q = 0; q < nq; q++)
                    for (int d = 0; d < dim; d++)
                        value[j] += w[q] * uLoc[d][q] * phi[q * nloc + i] * dT[(q * nloc + j) * dim + d];
                value[j] *= absDetB;
            }
            for (int d = 0; d < dim; d++) {
                for (int j = 0; j < nloc; j++) indices[j] = dim * gid[el[j]] + d;
                fo_insert(A, dim * gid[el[i]] + d, nloc, indices, value);
            }
        }
    }
    return 0;
}

/* FE_def.hpp:1839-1929 assemblyAdvectionInUVecField (W) */
int fo_assembly_advection_in_u(int dim, const char *fe, int64_t ne, const int32_t *conn, const double *coords,
                               const int64_t *gid, const double *u, fo_matrix *A)
{
    int nloc = fo_nloc(dim, fe);
    if (nloc < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int extraDeg = fo_determine_degree1(dim, fe, FO_GRAD);
    int deg = fo_determine_degree2(dim, fe, fe, FO_STD, FO_STD, extraDeg);
    int nq = fo_get_dphi(dim, fe, deg, dphi, w);
    if (nq < 0 || fo_get_phi(dim, fe, deg, phi, w) != nq) return -1;
    double duLoc[FO_MAXQ][3][3], value[FO_MAXN * 3];
    int64_t indices[FO_MAXN * 3];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *el = conn + T * nloc;
        double absDetB = fo_element_geometry(dim, nq, nloc, el, coords, dphi, dT);
        memset(duLoc, 0, sizeof(duLoc));
        for (int q = 0; q < nq; q++)
            for (int d1 = 0; d1 < dim; d1++)
                for (int i = 0; i < nloc; i++) {
                    double ui = u[(int64_t)dim * el[i] + d1];
                    for (int d2 = 0; d2 < dim; d2++) duLoc[q][d2][d1] += ui * dT[(q * nloc + i) * dim + d2];
                }
        for (int i = 0; i < nloc; i++)
            for (int d1 = 0; d1 < dim; d1++) {
                for (int j = 0; j < nloc; j++)
                    for (int d2 = 0; d2 < dim; d2++) {
                        double v = 0.;
                        for (int q = 0; q < nq; q++)
                            v += w[q] * duLoc[q][d2][d1] * phi[q * nloc + i] * phi[q * nloc + j];
                        v *= absDetB;
                        value[dim * j + d2] = v;
                        indices[dim * j + d2] = dim * gid[el[j]] + d2;
                    }
                fo_insert(A, dim * gid[el[i]] + d1, dim * nloc, indices, value);
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FE_def.hpp:1932-2057 assemblyDivAndDivT / 2061-2148 assemblyDivAndDivTFast             */
/* fe1 = velocity space (gradients), fe2 = pressure space (values); both element lists    */
/* share the element index T.  Positive sign (callers scale by -1).                       */
/* ------------------------------------------------------------------------------------ */
int fo_assembly_div_divT(int dim, const char *fe1, const char *fe2, int64_t ne,
                         const int32_t *conn1, const double *coords1, const int64_t *gid1,
                         const int32_t *conn2, const int64_t *gid2,
                         fo_matrix *Bm, fo_matrix *BTm, int fast)
{
    int n1 = fo_nloc(dim, fe1), n2 = fo_nloc(dim, fe2);
    if (n1 < 0 || n2 < 0) return -1;
    double dphi[FO_MAXQ * FO_MAXN * 3], phi[FO_MAXQ * FO_MAXN], w[FO_MAXQ], dT[FO_MAXQ * FO_MAXN * 3];
    int deg = fo_determine_degree2(dim, fe1, fe2, FO_GRAD, FO_STD, 0);
    int nq = fo_get_dphi(dim, fe1, deg, dphi, w);
    if (nq < 0 || fo_get_phi(dim, fe2, deg, phi, w) != nq) return -1;
    double valueVec[3][FO_MAXN];
    int64_t indices[FO_MAXN];
    for (int64_t T = 0; T < ne; T++) {
        const int32_t *e1 = conn1 + T * n1, *e2 = conn2 + T * n2;
        double absDetB = fo_element_geometry(dim, nq, n1, e1, coords1, dphi, dT);
        if (fast) {
            for (int i = 0; i < n2; i++) {
                int64_t row = gid2[e2[i]];
                for (int j = 0; j < n1; j++)
                    for (int d = 0; d < dim; d++) {
                        double v = 0.;
                        for (int q = 0; q < nq; q++) v += w[q] * phi[q * n2 + i] * dT[(q * n1 + j) * dim + d];
                        v *= absDetB;
                        int64_t col = dim * gid1[e1[j]] + d;
                        fo_insert(Bm, row, 1, &col, &v);
                        fo_insert(BTm, col, 1, &row, &v);
                    }
            }
            continue;
        }
        for (int i = 0; i < n2; i++) {
            for (int j = 0; j < n1; j++) {
                for (int d = 0; d < dim; d++) valueVec[d][j] = 0.;
                for (int q = 0; q < nq; q++)
                    for (int d = 0; d < dim; d++)
                        valueVec[d][j] += w[q] * phi[q * n2 + i] * dT[(q * n1 + j) * dim + d];
                for (int d = 0; d < dim; d++) valueVec[d][j] *= absDetB;
            }
            for (int d = 0; d < dim; d++) {
                for (int j = 0; j < n1; j++) indices[j] = dim * gid1[e1[j]] + d;
                fo_insert(Bm, gid2[e2[i]], n1, indices, valueVec[d]);
            }
        }
        for (int i = 0; i < n1; i++) {
            for (int j = 0; j < n2; j++) {
                for (int d = 0; d < dim; d++) valueVec[d][j] = 0.;
                for (int q = 0; q < nq; q++)
                    for (int d = 0; d < dim; d++)
                        valueVec[d][j] += w[q] * phi[q * n2 + j] * dT[(q * n1 + i) * dim + d];
                for (int d = 0; d < dim; d++) valueVec[d][j] *= absDetB;
            }
            for (int j = 0; j < n2; j++) indices[j] = gid2[e2[j]];
            for (int d = 0; d < dim; d++) fo_insert(BTm, dim * gid1[e1[i]] + d, n2, indices, valueVec[d]);
        }
    }
    return 0;
}
