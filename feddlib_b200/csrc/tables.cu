// Reference-element tables (host side): Lagrange P1/P2 bases on the unit simplex written in
// barycentric form, and the simplex quadrature rules FEDDLib selects for each operator.
//
// Behaviour follows (reference, /root/reference/feddlib/core/FE/FE_def.hpp):
//   determineDegree  :5431-5562   P1: Std 1 / Grad 0, P2: Std 2 / Grad 1, sum(+extra), 0 -> 1
//   getQuadratureValues :6023-6460  2D: deg 1 (1 pt), 2 (3 pts), 3..5 (7 pts);
//                                   3D: deg 1 (1 pt), 2..3 (5 pts, negative centre weight), 4..5 (15 pts)
//   phi :4991-5088, gradPhi :5570-5714  local node order: vertices, then edge nodes
//                                   (0,1),(1,2),(0,2)[,(0,3),(1,3),(2,3)]   
// The bases are evaluated through barycentric coordinates (lambda_0 = 1 - sum x, lambda_k = x_k):
//   P1: phi_v = lambda_v;  P2 vertex: lambda_v (2 lambda_v - 1);  P2 edge (a,b): 4 lambda_a lambda_b
// which is the same polynomial the reference hard-codes per case.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace fb {

static const int kEdges2[3][2] = {{0, 1}, {1, 2}, {0, 2}};
static const int kEdges3[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};

static int fe_order(int dim, int nloc)
{
    if (dim == 2 && nloc == 1) return 0;   // P0 (pressure of assemblyDivAndDivT, 2D)
    if (nloc == dim + 1) return 1;
    if ((dim == 2 && nloc == 6) || (dim == 3 && nloc == 10)) return 2;
    return -1;
}

// quadrature on the reference simplex; weights include the simplex volume
static int quad_rule(int dim, int deg, double pts[][3], double *w)
{
    if (dim == 2) {
        if (deg <= 1) {
            pts[0][0] = pts[0][1] = 1.0 / 3.0;
            w[0] = 0.5;
            return 1;
        }
        if (deg == 2) { // edge midpoints
            const double p[3][2] = {{0.5, 0.5}, {0.0, 0.5}, {0.5, 0.0}};
            for (int q = 0; q < 3; q++) { pts[q][0] = p[q][0]; pts[q][1] = p[q][1]; w[q] = 1.0 / 6.0; }
            return 3;
        }
        if (deg <= 5) { // Radon 7-point rule, constants as tabulated by the reference
            const double a = 0.470142064105115, b = 0.101286507323456;
            const double wa = 0.066197076394253, wb = 0.062969590272413;
            const double p[7][2] = {{1.0 / 3.0, 1.0 / 3.0}, {a, a}, {1 - 2. * a, a}, {a, 1 - 2. * a},
                                    {b, b}, {1 - 2. * b, b}, {b, 1 - 2. * b}};
            const double ww[7] = {9.0 / 80.0, wa, wa, wa, wb, wb, wb};
            for (int q = 0; q < 7; q++) { pts[q][0] = p[q][0]; pts[q][1] = p[q][1]; w[q] = ww[q]; }
            return 7;
        }
        return -1;
    }
    if (deg <= 1) {
        pts[0][0] = pts[0][1] = pts[0][2] = 0.25;
        w[0] = 1.0 / 6.0;
        return 1;
    }
    if (deg <= 3) { // 5-point rule with negative centre weight
        const double b = 1.0 / 6.0, c = 0.5;
        const double p[5][3] = {{0.25, 0.25, 0.25}, {b, b, b}, {b, b, c}, {b, c, b}, {c, b, b}};
        const double ww[5] = {-2.0 / 15.0, 3.0 / 40.0, 3.0 / 40.0, 3.0 / 40.0, 3.0 / 40.0};
        for (int q = 0; q < 5; q++) { for (int d = 0; d < 3; d++) pts[q][d] = p[q][d]; w[q] = ww[q]; }
        return 5;
    }
    if (deg <= 5) { // 15-point degree-5 rule
        const double s15 = std::sqrt(15.0);
        const double b1 = (7. + s15) / 34., b2 = (7. - s15) / 34.;
        const double c1 = (13. - 3. * s15) / 34., c2 = (13. + 3. * s15) / 34.;
        const double d = (5. - s15) / 20., e = (5. + s15) / 20.;
        const double p[15][3] = {{0.25, 0.25, 0.25},
                                 {b1, b1, b1}, {b1, b1, c1}, {b1, c1, b1}, {c1, b1, b1},
                                 {b2, b2, b2}, {b2, b2, c2}, {b2, c2, b2}, {c2, b2, b2},
                                 {d, d, e}, {d, e, d}, {e, d, d}, {d, e, e}, {e, d, e}, {e, e, d}};
        const double w1 = (2665. - 14. * s15) / 226800., w2 = (2665. + 14. * s15) / 226800., w3 = 5. / 567.;
        for (int q = 0; q < 15; q++) {
            for (int k = 0; k < 3; k++) pts[q][k] = p[q][k];
            w[q] = q == 0 ? 8. / 405. : (q < 5 ? w1 : (q < 9 ? w2 : w3));
        }
        return 15;
    }
    return -1;
}

static void bary(int dim, const double *x, double *lam)
{
    lam[0] = 1.0;
    for (int d = 0; d < dim; d++) { lam[0] -= x[d]; lam[d + 1] = x[d]; }
}

static double basis(int dim, int order, int i, const double *x)
{
    double lam[4];
    bary(dim, x, lam);
    if (order == 0) return 1.0;             // FE::phi `case 0: //P0` (FE_def.hpp:4993)
    if (order == 1) return lam[i];
    if (i <= dim) return lam[i] * (2.0 * lam[i] - 1.0);
    const int *e = dim == 2 ? kEdges2[i - 3] : kEdges3[i - 4];
    return 4.0 * lam[e[0]] * lam[e[1]];
}

static void basis_grad(int dim, int order, int i, const double *x, double *g)
{
    double lam[4];
    bary(dim, x, lam);
    // d lambda_v / d x_c
    auto dl = [&](int v, int c) { return v == 0 ? -1.0 : (v == c + 1 ? 1.0 : 0.0); };
    for (int c = 0; c < dim; c++) {
        if (order == 0) g[c] = 0.0;
        else if (order == 1) g[c] = dl(i, c);
        else if (i <= dim) g[c] = (4.0 * lam[i] - 1.0) * dl(i, c);
        else {
            const int *e = dim == 2 ? kEdges2[i - 3] : kEdges3[i - 4];
            g[c] = 4.0 * (lam[e[0]] * dl(e[1], c) + lam[e[1]] * dl(e[0], c));
        }
    }
}

// quadrature degree each operator asks for (FE_def.hpp:474, 626, 1770-1772, 1859-1861, 1963, 2758)
static int op_degree(int op, int ov, int op_)
{
    const int gradv = ov - 1, stdv = ov, stdp = op_;
    int deg;
    switch (op) {
    case OP_LAP:
    case OP_ELAS: deg = gradv + gradv; break;
    case OP_ADV:  deg = gradv + stdv + (stdv == 0 ? 1 : stdv); break;      // extra = degree of u_h
    case OP_ADVU:
    case OP_NSJ:  deg = stdv + stdv + (gradv == 0 ? 1 : gradv); break;     // extra = degree of grad u_h (0 -> 1)
    case OP_B:
    case OP_BT:   deg = gradv + stdp; break;
    case OP_MASS: deg = stdv + stdv; break;                                 // FE_def.hpp:474
    default: return -1;
    }
    return deg == 0 ? 1 : deg;
}

// c[i] = sum_q w_q phi_i(q) for the rule of degree determineDegree(FEType, Std) + deg_func (FE_def.hpp:4716-4717, 4746-4747)
int rhs_coefficients(int dim, int nloc, int deg_func, double *c)
{
    const int order = fe_order(dim, nloc);
    if (order < 0 || deg_func < 0) return -1;
    int deg = order + deg_func;
    if (deg == 0) deg = 1;
    double pts[MAXQ][3], w[MAXQ];
    const int nq = quad_rule(dim, deg, pts, w);
    if (nq < 0) return -1;
    for (int i = 0; i < nloc; i++) {
        double v = 0.0;
        for (int q = 0; q < nq; q++) v += w[q] * basis(dim, order, i, pts[q]);
        c[i] = v;
    }
    return 0;
}

int build_tables(OpTables &t, int op, int dim, int nloc_v, int nloc_p)
{
    std::memset(&t, 0, sizeof(t));
    const int ov = fe_order(dim, nloc_v), opp = fe_order(dim, nloc_p);
    if (ov < 0 || opp < 0) return -1;
    const int deg = op_degree(op, ov, opp);
    double pts[MAXQ][3];
    const int nq = quad_rule(dim, deg, pts, t.w);
    if (nq < 0) return -1;
    t.nq = nq; t.nv = nloc_v; t.np = nloc_p; t.dim = dim;
    for (int q = 0; q < nq; q++) {
        bary(dim, pts[q], &t.lam[q * 4]);
        for (int i = 0; i < nloc_v; i++) basis_grad(dim, ov, i, pts[q], &t.dphi[(q * nloc_v + i) * dim]);
        for (int i = 0; i < nloc_p; i++) t.phi[q * nloc_p + i] = basis(dim, opp, i, pts[q]);
    }
    return 0;
}

} // namespace fb
