/* TEST INFRASTRUCTURE ONLY -- see ldm_oracle.h.
 *
 * PARITY UNPINNED beyond the uniform state.  The reference has no counterpart of what this file computes: it closes the
 * third moment with the power-1 closure T(xi, xi') = C(xi) C(xi') / N, under which the triple integral factorises into
 * C(xi) * WC / N with ONE scalar WC (S:265-274, S:483-488).  BASELINE.json's north_star names the integrals
 * that do NOT factorise -- (w*C)(xi) and the integral of w(xi') C(xi') C(xi + xi') at every xi -- which are what the
 * other two closures of the same paper (Law & Dieckmann 2000, the text around eq. 14.10; Dieckmann & Law 2000) need:
 *
 *   power-3 (Kirkwood):   T = C(xi) C(xi') C(xi' - xi) / N^3
 *   power-2 (symmetric):  T = 1/2 [ C(xi) C(xi')/N + C(xi) C(xi' - xi)/N + C(xi') C(xi' - xi)/N - N^3 ]
 *
 * This file is the CPU statement of that extension, written in the reference's own style so that it can be read
 * against it: the same quadrature nodes and loop order as MakeIntegralWithPairDensity (S:297-316), the same
 * BilinearInterpolation with SquaredN beyond the cutoff (ldmo_bilinear), the same doubled form of the second-moment
 * equation (S:464-505), only the term of S:486-487 replaced.  What pins it: at C == N^2 all three closures give the
 * reference's own right-hand side (the golden vectors of tests/golden/ computed by the reference), and power-1 selected
 * here is ldmo_rhs bit for bit (tests/test_oracle_golden.py).  Away from the uniform state it is a definition, not a
 * restatement.
 *
 * C is even in both coordinates and stored on the non-negative quadrant, w is even: the integral over xi' of
 * w(xi') C(xi') C(xi' - xi) equals the one of w(xi') C(xi') C(xi + xi'), and C(xi') is read at (|xi'_x|, |xi'_y|) --
 * the sample positions of the reference's WC integral at shift (0, 0).
 */
#include "ldm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* the kernel-node coordinate of quadrature index i (S:307-308): +kappa[i+1] for i >= 0, -kappa[-i-1] below */
static double signed_offset(const double* kg, int i) { return i >= 0 ? kg[i + 1] : -kg[-i - 1]; }

/* at node (x, y), over the quadrature of w (extent mW, S:300-302):
 *   with_own_density == 0:  (w*C)(xi)  = h^2 sum w(xi') C~(xi + xi')
 *   with_own_density != 0:  K(xi)      = h^2 sum w(xi') C~(xi') C~(xi + xi')
 * C~ = bilinear C inside the cutoff, N^2 beyond (SquaredN, the type every experiment of the reference passes). */
static double w_integral(const ldmo_problem* p, int x, int y, const double* C, double asym, int with_own_density)
{
    const int K = p->K, m = p->mW;
    const double* mg = p->moments_grid;
    const double* kg = p->kernels_grid;
    double integral = 0;
    for (int i = 1 - m; i < m - 1; i += 2) {
        for (int j = 1 - m; j < m - 1; j += 2) {
            const double kx = signed_offset(kg, i), ky = signed_offset(kg, j);
            const double sum_x = fabs(mg[x] + kx), sum_y = fabs(mg[y] + ky);
            const double shifted = ldmo_bilinear(mg, p->G, C, sum_x, sum_y, p->cutoff, 1, asym);
            double weight = p->W[(size_t)abs(i + 1) * K + abs(j + 1)];
            if (with_own_density)
                weight = weight * ldmo_bilinear(mg, p->G, C, fabs(mg[0] + kx), fabs(mg[0] + ky), p->cutoff, 1, asym);
            integral += weight * shifted;
        }
    }
    integral *= p->h * p->h;
    return integral;
}

/* h^2 sum w over the same quadrature: the midpoint-rule integral of w (Properties::CheckKernels, S:147-183, holds it to 1 +- 1e-8) */
double ldmo_w_mass(const ldmo_problem* p)
{
    const int K = p->K, m = p->mW;
    double integral = 0;
    for (int i = 1 - m; i < m - 1; i += 2)
        for (int j = 1 - m; j < m - 1; j += 2) integral += p->W[(size_t)abs(i + 1) * K + abs(j + 1)];
    return integral * (p->h * p->h);
}

void ldmo_closure_integrals(const ldmo_problem* p, const double* state, double* w_conv_C, double* K_out)
{
    const int G = p->G;
    const double N = state[0];
    const double* C = state + 1;
    const double asym = N * N;
    #pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) {
            if (w_conv_C) w_conv_C[(size_t)x * G + y] = w_integral(p, x, y, C, asym, 0);
            if (K_out) K_out[(size_t)x * G + y] = w_integral(p, x, y, C, asym, 1);
        }
}

/* Equation::operator() (S:431-513) with the closure term of S:486-487 exchanged.  closure: 1 power-1 (the reference),
 * 2 power-2 symmetric, 3 power-3. */
void ldmo_rhs_closure(const ldmo_problem* p, const double* state, int closure, double* deriv)
{
    const int G = p->G;
    const size_t P = (size_t)G * G;
    const double N = state[0];
    const double* C = state + 1;
    double WC;
    double* MC = (double*)malloc(sizeof(double) * P);
    double* WcC = (double*)malloc(sizeof(double) * P);
    double* Kc = (double*)malloc(sizeof(double) * P);
    ldmo_integrals(p, state, &WC, MC);
    if (closure != 1) ldmo_closure_integrals(p, state, closure == 2 ? WcC : NULL, Kc);
    const double mass = ldmo_w_mass(p);

    double dN = N * (p->b - p->d);                                        /* S:446 */
    double d_prime_sum = 0;
    d_prime_sum = d_prime_sum + p->dd * WC;                               /* S:450 */
    dN = dN - d_prime_sum;                                                /* S:451 */
    deriv[0] = dN;

    for (int x = 0; x < G; x++) {
        for (int y = 0; y < G; y++) {
            const size_t q = (size_t)x * G + y;
            const double c = C[q];
            const double mc = MC[q];
            double result = p->b * mc + p->b * mc;                        /* S:464 */
            result += 2 * p->b * N * ldmo_bilinear(p->kernels_grid, p->K, p->M,
                p->moments_grid[x], p->moments_grid[y], p->sigmaM * 3, 0, 0.0);   /* S:469 */
            result -= p->d * c;                                           /* S:480 */
            result -= p->d * c;                                           /* S:481 */
            if (closure == 1) {
                result -= p->dd * WC * c / N;                             /* S:486 */
                result -= p->dd * WC * c / N;                             /* S:487 */
            } else {
                double triplets;                                          /* integral of w(xi') T(xi, xi') */
                if (closure == 3) triplets = c * Kc[q] / (N * N * N);
                else triplets = 0.5 * (WC * c / N + c * WcC[q] / N + Kc[q] / N - N * N * N * mass);
                result -= p->dd * triplets;
                result -= p->dd * triplets;
            }
            const double w = ldmo_bilinear(p->kernels_grid, p->K, p->W,
                p->moments_grid[x], p->moments_grid[y], p->sigmaW * 3, 0, 0.0);
            result -= p->dd * c * w;                                      /* S:490 */
            result -= p->dd * c * w;                                      /* S:498 */
            deriv[1 + q] = result;
        }
    }
    free(MC); free(WcC); free(Kc);
}

/* EulerSolver::Solve (S:557-592) on that right-hand side: full history of the values; returns the number of points */
int ldmo_solve_closure(const ldmo_problem* p, int closure, double t_max, double dt, const double* init, double* times, double* values)
{
    const size_t n = 1 + (size_t)p->G * p->G;
    double* value = (double*)malloc(sizeof(double) * n);
    double* derivative = (double*)malloc(sizeof(double) * n);
    memcpy(value, init, sizeof(double) * n);
    double time = 0;
    ldmo_rhs_closure(p, value, closure, derivative);
    int k = 0;
    for (;;) {
        if (times) times[k] = time;
        if (values) memcpy(values + (size_t)k * n, value, sizeof(double) * n);
        time += dt;
        if (!(time < t_max)) break;
        for (size_t q = 0; q < n; q++) {
            double v = value[q] + derivative[q] * dt;
            if (v < 0) v = 0;
            value[q] = v;
        }
        ldmo_rhs_closure(p, value, closure, derivative);
        k++;
    }
    free(value); free(derivative);
    return k + 1;
}
