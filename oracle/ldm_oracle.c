/* TEST INFRASTRUCTURE ONLY -- see ldm_oracle.h.  Plain-C restatement of the reference's hot path,
 * one species, every floating-point operation in the reference's order (compile with
 * -ffp-contract=off and no -ffast-math; oracle/Makefile does).
 * S:n = Solver.cpp line n, T:n = types.h line n of the reference. */
#include "ldm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double LDMO_PI = 3.1415927;   /* S:23 (sic) */
static const double LDMO_EPSILON = 1E-08;  /* S:24 */

static int g_threads = 0;
void ldmo_set_threads(int n) { g_threads = n; }

/* The log-stretched node set for a given node count (S:69 / S:73).  Vector::setLinSpaced(n,0,n-1)
 * yields min + i*step with step = (n-1)/(n-1) = 1 and appends max verbatim (T:120-130); the Vector
 * operators then apply, left to right: unary minus (x * -1, T:175-183), / n, + 1, log, unary
 * minus, * 0.01, * n, abs. */
static void moments_nodes(int n, double* g)
{
    double step = ((double)(n - 1) - 0.0) / ((double)n - 1.0);
    for (int i = 0; i < n; i++) {
        double k = (i < n - 1) ? 0.0 + i * step : (double)(n - 1);
        double v = -1 * k;
        v = v / n;
        v = v + 1;
        v = log(v);
        v = -1 * v;
        v = v * 0.01;
        v = v * n;
        g[i] = fabs(v);
    }
}

/* Properties::GenerateKernel (S:98-145): truncated Gaussian on the K x K kernel nodes, renormalised
 * by the midpoint rule over the odd nodes. */
static int generate_kernel(const ldmo_problem* p, double s, double R, double* kernel)
{
    const int K = p->K;
    const double* grid = p->kernels_grid;
    const double step = p->h;
    int max_non_zero_num = 0;
    for (int x = 0; x < K; x++) {
        for (int y = 0; y < K; y++) {
            const double r = hypot(grid[x], grid[y]);                     /* S:114 */
            if (r <= R) {
                double norm = s * sqrt(2 * LDMO_PI) * (1 - exp((-R * R) / (2 * s * s)));   /* S:120 */
                double weight = exp((-r * r) / (2 * s * s));              /* S:121 */
                kernel[(size_t)x * K + y] = weight / norm;
                if (kernel[(size_t)x * K + y] != 0) {                     /* S:125-126 */
                    if (x + 1 > max_non_zero_num) max_non_zero_num = x + 1;
                    if (y + 1 > max_non_zero_num) max_non_zero_num = y + 1;
                }
            } else {
                kernel[(size_t)x * K + y] = 0;
            }
        }
    }
    if (K % 2 == 1 && max_non_zero_num % 2 == 0 && max_non_zero_num != 2) max_non_zero_num++;   /* S:131 */
    if (max_non_zero_num == 1) max_non_zero_num = 2;                      /* S:132 */

    double integral = 0;                                                  /* S:135-138 */
    for (int x = 1 - max_non_zero_num; x < max_non_zero_num - 1; x += 2)
        for (int y = 1 - max_non_zero_num; y < max_non_zero_num - 1; y += 2)
            integral += kernel[(size_t)abs(x + 1) * K + abs(y + 1)];
    integral *= step * step;                                              /* S:141 */
    for (size_t q = 0; q < (size_t)K * K; q++) kernel[q] = kernel[q] / integral;   /* S:142, T:411-420 */
    return max_non_zero_num;
}

ldmo_problem* ldmo_generate(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G)
{
    ldmo_problem* p = (ldmo_problem*)calloc(1, sizeof(ldmo_problem));
    p->sigmaW = sigma_w; p->sigmaM = sigma_m; p->b = b; p->d = d; p->dd = d_prime;

    /* GenerateGrids (S:63-95) */
    double cutoff_distance = (sigma_w + sigma_m) * 5;
    double max_radius = (sigma_w > sigma_m ? sigma_w : sigma_m) * 5;
    int n = forced_G > 0 ? forced_G : 5;
    double* g = (double*)malloc(sizeof(double) * (size_t)n);
    moments_nodes(n, g);
    if (forced_G <= 0) {
        while (g[n - 1] < cutoff_distance - 0.01) {                       /* S:70-74 */
            n++;
            g = (double*)realloc(g, sizeof(double) * (size_t)n);
            moments_nodes(n, g);
        }
    }
    double parameter = cutoff_distance / g[n - 1];                        /* S:75 */
    for (int i = 0; i < n; i++) g[i] = g[i] * parameter;                  /* S:76 */

    int K = (n - 1) * 2 - 1;                                              /* S:79 */
    if ((sigma_w < sigma_m ? sigma_w : sigma_m) < 0.015) {                /* S:80-84 */
        K = (K - 2) * 4 - 1;
        K = (K - 1) * 2 - 1;
    }
    double* kg = (double*)malloc(sizeof(double) * (size_t)K);             /* setLinSpaced(K, 0, max_radius), T:120-130 */
    {
        double step = (max_radius - 0.0) / ((double)K - 1.0);
        for (int i = 0; i < K - 1; i++) kg[i] = 0.0 + i * step;
        kg[K - 1] = max_radius;
    }
    p->G = n; p->K = K;
    p->cutoff = g[n - 1];                                                 /* S:91 */
    p->moments_grid = g;
    p->kernels_grid = kg;
    p->h = kg[2] - kg[0];                                                 /* S:94 */

    /* Generate (S:205-239): radius 3 sigma (Matrix/Vector * 3) */
    p->W = (double*)malloc(sizeof(double) * (size_t)K * K);
    p->M = (double*)malloc(sizeof(double) * (size_t)K * K);
    p->mW = generate_kernel(p, sigma_w, sigma_w * 3, p->W);
    p->mM = generate_kernel(p, sigma_m, sigma_m * 3, p->M);
    return p;
}

void ldmo_free(ldmo_problem* p)
{
    if (!p) return;
    free(p->moments_grid); free(p->kernels_grid); free(p->M); free(p->W); free(p);
}

/* Extrapolation (S:355-410): Zero, SquaredN, Alpha as the reference computes it -- `2 / (grid.size() + 1)` is an integer
 * division (S:377), so alpha == 0, every predicted value equals values[0][0] and the result is values[0][0] -- and
 * Exponential (S:392-408). */
static double extrapolate(const double* values, int n, int type, double asymptotic_value, const double* grid, double distance,
                          double max_distance)
{
    if (fabs(values[n - 1] - asymptotic_value) < LDMO_EPSILON)            /* S:357: values[0].last() */
        return asymptotic_value;
    if (type == 2) {                                                      /* Alpha, S:375-389 */
        const double alpha = 2 / (n + 1);                                 /* integer division, as in the reference */
        double predicted = values[0];
        for (int k = 1; k < n; k++) predicted = alpha * values[k] + (1 - alpha) * predicted;
        return alpha * asymptotic_value + (1 - alpha) * predicted;
    }
    if (type == 3) {                                                      /* Exponential, S:392-408 */
        const int accuracy = 5;
        double decay_avg = 0;
        for (int k = 1; k < accuracy; k++) {
            double pair_distance = max_distance - grid[n - 1 - k];
            double new_decay_value = fabs((values[n - 1 - k] - asymptotic_value) / (values[n - 1] - asymptotic_value));
            new_decay_value = log(new_decay_value) / pair_distance;
            decay_avg += new_decay_value;
        }
        decay_avg = decay_avg / (accuracy - 1);
        return asymptotic_value + (values[n - 1] - asymptotic_value) * exp(-1 * decay_avg * (distance - max_distance));
    }
    return type == 1 ? asymptotic_value : 0.0;                            /* SquaredN : Zero */
}

static int g_c_extrapolation = 1;                                         /* what Equation::operator() is called with */
void ldmo_set_extrapolation(int type) { g_c_extrapolation = type; }

double ldmo_bilinear(const double* grid, int n, const double* values, double x, double y,
                     double max_distance, int extrapolation, double asymptotic_value)
{
    double distance = hypot(x, y);                                        /* S:329 */
    if (distance > max_distance)
        return extrapolate(values, n, extrapolation, asymptotic_value, grid, distance, max_distance);
    int i = 1, j = 1;                                                     /* S:334-336 */
    while (x > grid[i]) i++;
    while (y > grid[j]) j++;
    double a = grid[i - 1], b = grid[i], c = grid[j - 1], d = grid[j];
    double wx = (x - a) / (b - a);                                        /* S:344-345 */
    double wy = (y - c) / (d - c);
    double value = values[(size_t)(i - 1) * n + (j - 1)] * (1 - wx) * (1 - wy) +   /* S:348-351 */
                   values[(size_t)i * n + (j - 1)] * wx * (1 - wy) +
                   values[(size_t)(i - 1) * n + j] * (1 - wx) * wy +
                   values[(size_t)i * n + j] * wx * wy;
    return value;
}

double ldmo_pair_integral(const ldmo_problem* p, int x_shift, int y_shift, int max_non_zero_num,
                          const double* kernel, const double* C, double asymptotic_value)
{
    const int K = p->K;
    const double* mg = p->moments_grid;
    const double* kg = p->kernels_grid;
    double integral = 0;
    for (int i = 1 - max_non_zero_num; i < max_non_zero_num - 1; i += 2) {       /* S:300 */
        for (int j = 1 - max_non_zero_num; j < max_non_zero_num - 1; j += 2) {   /* S:302 */
            const double sum_x = fabs(mg[x_shift] + (i >= 0 ? kg[i + 1] : -kg[-i - 1]));   /* S:307 */
            const double sum_y = fabs(mg[y_shift] + (j >= 0 ? kg[j + 1] : -kg[-j - 1]));   /* S:308 */
            double pair_density = ldmo_bilinear(mg, p->G, C, sum_x, sum_y, p->cutoff, g_c_extrapolation, asymptotic_value);
            integral += kernel[(size_t)abs(i + 1) * K + abs(j + 1)] * pair_density;       /* S:311 */
        }
    }
    integral *= p->h * p->h;                                              /* S:314: pow(step, 2) */
    return integral;
}

void ldmo_integrals(const ldmo_problem* p, const double* state, double* WC, double* MC)
{
    const int G = p->G;
    const double N = state[0];
    const double* C = state + 1;
    const double asym = N * N;                                            /* S:271, S:289 */
    if (WC) *WC = ldmo_pair_integral(p, 0, 0, p->mW, p->W, C, asym);      /* S:265-274 */
    if (MC) {
        int threads = g_threads;
#ifdef _OPENMP
        if (threads <= 0) threads = omp_get_max_threads();
#endif
        (void)threads;
        #pragma omp parallel for collapse(2) schedule(dynamic, 4) num_threads(threads)
        for (int x = 0; x < G; x++)                                       /* S:287-289 */
            for (int y = 0; y < G; y++)
                MC[(size_t)x * G + y] = ldmo_pair_integral(p, x, y, p->mM, p->M, C, asym);
    }
}

void ldmo_kernels_at_nodes(const ldmo_problem* p, double* m_at_nodes, double* w_at_nodes)
{
    const int G = p->G;
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) {
            m_at_nodes[(size_t)x * G + y] = ldmo_bilinear(p->kernels_grid, p->K, p->M,
                p->moments_grid[x], p->moments_grid[y], p->sigmaM * 3, 0, 0.0);   /* S:469-476 */
            w_at_nodes[(size_t)x * G + y] = ldmo_bilinear(p->kernels_grid, p->K, p->W,
                p->moments_grid[x], p->moments_grid[y], p->sigmaW * 3, 0, 0.0);   /* S:490-497 */
        }
}

void ldmo_rhs(const ldmo_problem* p, const double* state, double* deriv)
{
    const int G = p->G;
    const double N = state[0];
    const double* C = state + 1;
    double WC;
    double* MC = (double*)malloc(sizeof(double) * (size_t)G * G);
    ldmo_integrals(p, state, &WC, MC);                                    /* S:442-443 */

    double dN = N * (p->b - p->d);                                        /* S:446 */
    double d_prime_sum = 0;
    d_prime_sum = d_prime_sum + p->dd * WC;                               /* S:450 */
    dN = dN - d_prime_sum;                                                /* S:451 */
    deriv[0] = dN;

    for (int x = 0; x < G; x++) {
        for (int y = 0; y < G; y++) {
            const double c = C[(size_t)x * G + y];
            const double mc = MC[(size_t)x * G + y];
            double result = p->b * mc + p->b * mc;                        /* S:464 (i == j) */
            result += 2 * p->b * N * ldmo_bilinear(p->kernels_grid, p->K, p->M,
                p->moments_grid[x], p->moments_grid[y], p->sigmaM * 3, 0, 0.0);   /* S:469 */
            result -= p->d * c;                                           /* S:480 */
            result -= p->d * c;                                           /* S:481 */
            result -= p->dd * WC * c / N;                                 /* S:486 */
            result -= p->dd * WC * c / N;                                 /* S:487 */
            const double w = ldmo_bilinear(p->kernels_grid, p->K, p->W,
                p->moments_grid[x], p->moments_grid[y], p->sigmaW * 3, 0, 0.0);
            result -= p->dd * c * w;                                      /* S:490 */
            result -= p->dd * c * w;                                      /* S:498 */
            deriv[1 + (size_t)x * G + y] = result;
        }
    }
    free(MC);
}

int ldmo_count_points(double t_max, double dt)
{
    int n = 1;
    double time = 0;
    time += dt;                                                           /* S:574 */
    while (time < t_max) { n++; time += dt; }                             /* S:575, S:588 */
    return n;
}

int ldmo_solve(const ldmo_problem* p, double t_max, double dt, const double* init, int full_history,
               double* times, double* values, double* derivs, double* observables)
{
    const int G = p->G;
    const size_t n = 1 + (size_t)G * G;
    double* value = (double*)malloc(sizeof(double) * n);
    double* derivative = (double*)malloc(sizeof(double) * n);
    memcpy(value, init, sizeof(double) * n);
    double time = 0;                                                      /* S:565 */
    ldmo_rhs(p, value, derivative);                                       /* S:567 */
    int k = 0;
    for (;;) {
        if (times) times[k] = time;
        if (observables) { observables[2 * k] = value[0]; observables[2 * k + 1] = value[G]; }   /* S:1011-1014 */
        if (full_history) {
            if (values) memcpy(values + (size_t)k * n, value, sizeof(double) * n);
            if (derivs) memcpy(derivs + (size_t)k * n, derivative, sizeof(double) * n);
        }
        time += dt;                                                       /* S:574 / S:588 */
        if (!(time < t_max)) break;                                       /* S:575 */
        for (size_t q = 0; q < n; q++) {
            double v = value[q] + derivative[q] * dt;                     /* S:577, T:186-220: multiply, then add */
            if (v < 0) v = 0;                                             /* S:578-580 */
            value[q] = v;
        }
        ldmo_rhs(p, value, derivative);                                   /* S:582 */
        k++;
    }
    if (!full_history) {
        if (values) memcpy(values, value, sizeof(double) * n);
        if (derivs) memcpy(derivs, derivative, sizeof(double) * n);
    }
    free(value); free(derivative);
    return k + 1;
}

int ldmo_count_points_rk4(double t_max, double dt)
{
    int n = 1;
    double time = 0;
    while (time < t_max) { time += dt / 2; time += dt / 2; n++; }         /* S:627-634 */
    return n;
}

/* RungeKutta4thSolver::Solve (S:597-654): present in the reference, never instantiated by its experiments */
int ldmo_solve_rk4(const ldmo_problem* p, double t_max, double dt, const double* init, double* times, double* values, double* derivs)
{
    const int G = p->G;
    const size_t n = 1 + (size_t)G * G;
    double* value = (double*)malloc(sizeof(double) * n);
    double* stage = (double*)malloc(sizeof(double) * n);
    double* c[4];
    for (int i = 0; i < 4; i++) c[i] = (double*)malloc(sizeof(double) * n);
    memcpy(value, init, sizeof(double) * n);
    double time = 0;
    int k = 0;
    if (times) times[0] = time;
    if (values) memcpy(values, value, sizeof(double) * n);
    while (time < t_max) {                                                /* S:627 */
        ldmo_rhs(p, value, c[0]);                                         /* S:629, S:631 (evaluated twice there) */
        if (derivs) memcpy(derivs + (size_t)k * n, c[0], sizeof(double) * n);
        time += dt / 2;
        for (size_t q = 0; q < n; q++) stage[q] = value[q] + c[0][q] * dt / 2;   /* S:633 */
        ldmo_rhs(p, stage, c[1]);
        for (size_t q = 0; q < n; q++) stage[q] = value[q] + c[1][q] * dt / 2;   /* S:634 */
        ldmo_rhs(p, stage, c[2]);
        time += dt / 2;
        for (size_t q = 0; q < n; q++) stage[q] = value[q] + c[2][q] * dt;       /* S:636 */
        ldmo_rhs(p, stage, c[3]);
        for (size_t q = 0; q < n; q++) {
            double sum = c[0][q];                                         /* S:637-640 */
            sum = sum + c[1][q] * 2;
            sum = sum + c[2][q] * 2;
            sum = sum + c[3][q];
            double v = value[q] + sum * dt / 6;                           /* S:641 */
            if (v < 0) v = 0;                                             /* S:643-645 */
            value[q] = v;
        }
        k++;
        if (times) times[k] = time;
        if (values) memcpy(values + (size_t)k * n, value, sizeof(double) * n);
    }
    ldmo_rhs(p, value, c[0]);                                             /* S:652 */
    if (derivs) memcpy(derivs + (size_t)k * n, c[0], sizeof(double) * n);
    free(value); free(stage);
    for (int i = 0; i < 4; i++) free(c[i]);
    return k + 1;
}


/* ---- several species ---- */
void ldmo_rhs_multi(const ldmo_multi* p, const double* state, double* deriv)
{
    const int s = p->species, G = p->G, K = p->K;
    const size_t P = (size_t)G * G, KK = (size_t)K * K;
    const double* N = state;
    ldmo_problem shared;                                   /* what ldmo_pair_integral / ldmo_bilinear read: the grids */
    memset(&shared, 0, sizeof shared);
    shared.G = G; shared.K = K; shared.h = p->h; shared.cutoff = p->cutoff;
    shared.moments_grid = (double*)p->moments_grid; shared.kernels_grid = (double*)p->kernels_grid;

    double* WC = (double*)malloc(sizeof(double) * (size_t)s * s);
    double* MC = (double*)malloc(sizeof(double) * (size_t)s * s * P);
    for (int i = 0; i < s; i++)                                                          /* S:265-274 */
        for (int j = 0; j < s; j++)
            WC[i * s + j] = ldmo_pair_integral(&shared, 0, 0, p->mW[i * s + j], p->W + (size_t)(i * s + j) * KK,
                                               state + s + (size_t)(i * s + j) * P, N[i] * N[j]);
    for (int i = 0; i < s; i++)                                                          /* S:277-295 */
        for (int j = 0; j < s; j++) {
            const double* C = state + s + (size_t)(i * s + j) * P;
            double* out = MC + (size_t)(i * s + j) * P;
            #pragma omp parallel for collapse(2) schedule(dynamic, 4)
            for (int x = 0; x < G; x++)
                for (int y = 0; y < G; y++)
                    out[(size_t)x * G + y] = ldmo_pair_integral(&shared, x, y, p->mM[i], p->M + (size_t)i * KK, C, N[i] * N[j]);
        }
    for (int i = 0; i < s; i++) {                                                        /* S:446-451 */
        double dN = N[i] * (p->b[i] - p->d[i]);
        double d_prime_sum = 0;
        for (int j = 0; j < s; j++) d_prime_sum = d_prime_sum + p->dd[i * s + j] * WC[i * s + j];
        deriv[i] = dN - d_prime_sum;
    }
    for (int i = 0; i < s; i++)                                                          /* S:454-508 */
        for (int j = 0; j < s; j++) {
            const double* Cij = state + s + (size_t)(i * s + j) * P;
            const double* Cji = state + s + (size_t)(j * s + i) * P;
            const double* MCij = MC + (size_t)(i * s + j) * P;
            const double* MCji = MC + (size_t)(j * s + i) * P;
            for (int x = 0; x < G; x++)
                for (int y = 0; y < G; y++) {
                    const size_t q = (size_t)x * G + y;
                    double result = p->b[i] * MCij[q] + p->b[j] * MCji[q];
                    if (i == j)
                        result += 2 * p->b[i] * N[i] * ldmo_bilinear(p->kernels_grid, K, p->M + (size_t)i * KK,
                            p->moments_grid[x], p->moments_grid[y], p->sigmaM[i] * 3, 0, 0.0);
                    result -= p->d[i] * Cij[q];
                    result -= p->d[j] * Cji[q];
                    for (int k = 0; k < s; k++) {
                        result -= p->dd[i * s + k] * WC[i * s + k] * Cij[q] / N[i];
                        result -= p->dd[j * s + k] * WC[j * s + k] * Cji[q] / N[j];
                    }
                    result -= p->dd[i * s + j] * Cij[q] * ldmo_bilinear(p->kernels_grid, K, p->W + (size_t)(i * s + j) * KK,
                        p->moments_grid[x], p->moments_grid[y], p->sigmaW[i * s + j] * 3, 0, 0.0);
                    result -= p->dd[j * s + i] * Cji[q] * ldmo_bilinear(p->kernels_grid, K, p->W + (size_t)(j * s + i) * KK,
                        p->moments_grid[x], p->moments_grid[y], p->sigmaW[j * s + i] * 3, 0, 0.0);
                    deriv[s + (size_t)(i * s + j) * P + q] = result;
                }
        }
    free(WC); free(MC);
}

int ldmo_solve_multi(const ldmo_multi* p, double t_max, double dt, const double* init, double* times, double* values)
{
    const size_t n = (size_t)p->species + (size_t)p->species * p->species * p->G * p->G;
    double* value = (double*)malloc(sizeof(double) * n);
    double* derivative = (double*)malloc(sizeof(double) * n);
    memcpy(value, init, sizeof(double) * n);
    double time = 0;
    ldmo_rhs_multi(p, value, derivative);
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
        ldmo_rhs_multi(p, value, derivative);
        k++;
    }
    free(value); free(derivative);
    return k + 1;
}
