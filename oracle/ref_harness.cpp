// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Harness around the UNMODIFIED reference translation unit.  The reference has no library
// boundary (one TU with main(), Solver.cpp:1323), so the harness renames its main and includes
// the file where it lies under /root/reference; nothing of the reference is copied into this
// repository.  Built by oracle/Makefile into oracle/_ref/ (git-ignored), twice:
//   libldm_ref.so        the reference exactly as shipped (g++ -O3 -fopenmp, README.md:14)
//   libldm_ref_byref.so  the same source streamed through sed so that the five by-value matrix
//                        parameters (Solver.cpp:265,277,297,318-320,355) become references;
//                        bit-identical results (asserted by tests/golden/make_golden.py),
//                        ~30x faster, used where the shipped build would need hours.
//
// Everything here is a thin extern "C" view of the reference's own classes:
//   FileReader (Solver.cpp:665-816), Properties (40-242), Equation (246-514),
//   EulerSolver (541-593).
#include <cstdio>
#include <cstring>
// every standard header the reference pulls in, parsed BEFORE the access-opening defines below
#include <iostream>
#include <string>
#include <cmath>
#include <fstream>
#include <sstream>
#include <cassert>
#include <time.h>
#include <omp.h>
#include <vector>
#include <numeric>
#include <algorithm>

#define main ldm_reference_main
#define private public            // reach Equation's private integrals for finer-grained goldens
#define class struct              // ... including the members declared before any access specifier
#include LDM_REF_SOLVER           // "/root/reference/Euler Solver ver.5.3/Solver.cpp" (or the sed stream)
#undef class
#undef private
#undef main

namespace {

struct QuietCout {                 // the reference narrates on std::cout (Solver.cpp:85, 572, ...)
    std::streambuf* saved;
    std::ostringstream sink;
    QuietCout() : saved(std::cout.rdbuf(sink.rdbuf())) {}
    ~QuietCout() { std::cout.rdbuf(saved); }
};

struct RefHandle {
    Properties props;
    SolverResults results;
};

ExtrapolationType g_type = SquaredN;      // what Equation::operator() / the solvers are called with (ref_set_extrapolation)

}  // namespace

extern "C" {

// route 0: FileReader::GetProperties (Solver.cpp:745), 1: GetWithoutSigmaW(sigma) (769),
// 2: GetWithoutSigmaM(sigma) (794)
void* ref_props_from_file(const char* path, int route, double sigma)
{
    QuietCout quiet;
    FileReader reader{std::string(path)};
    auto* h = new RefHandle;
    switch (route) {
        case 0: h->props = reader.GetProperties(false); break;
        case 1: h->props = reader.GetWithoutSigmaW(sigma, false); break;
        case 2: h->props = reader.GetWithoutSigmaM(sigma, false); break;
        default: delete h; return nullptr;
    }
    return h;
}

// One-species Properties straight from numbers, through the reference's own Properties::Generate
// (Solver.cpp:186).  forced_G <= 0 keeps the reference's grid search.  forced_G > 0 reproduces
// Properties::Generate + GenerateGrids (63-95) with the `while` search (70-74) skipped and
// num_for_moments pinned -- the synthetic-grid configurations of BASELINE.json; every expression
// below is the reference's, evaluated by the reference's own Vector operators.
void* ref_props_from_numbers(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G)
{
    QuietCout quiet;
    auto* h = new RefHandle;
    Matrix sW(1, sigma_w);
    Vector sM(1, sigma_m);
    Vector vb(1, b), vd(1, d);
    Matrix dd(1, d_prime);
    if (forced_G <= 0) {
        h->props = Properties::Generate(1, sW, sM, vb, vd, dd);
        return h;
    }
    Properties& p = h->props;
    p.SpeciesNumber = 1;
    p.SigmaM = sM; p.SigmaW = sW; p.b = vb; p.d = vd; p.d_prime = dd;
    {
        CalcType cutoff_distance = (sigma_w + sigma_m) * 5;
        CalcType max_radius = std::max(sigma_w, sigma_m) * 5;
        int num_for_moments = forced_G;
        auto moments_grid = Vector::abs(-Vector::log(-Vector::setLinSpaced(num_for_moments, 0, num_for_moments - 1) / num_for_moments + 1) * 0.01 * num_for_moments);
        CalcType parameter = cutoff_distance / moments_grid.last();
        moments_grid = moments_grid * parameter;
        int num_for_kernels = (num_for_moments - 1) * 2 - 1;
        if (std::min(sigma_w, sigma_m) < 0.015) {
            num_for_kernels = (num_for_kernels - 2) * 4 - 1;
            num_for_kernels = (num_for_kernels - 1) * 2 - 1;
        }
        auto kernels_grid = Vector::setLinSpaced(num_for_kernels, 0, max_radius);
        p.CutoffDistance = moments_grid.last();
        p.MomentsGrid = moments_grid;
        p.KernelsGrid = kernels_grid;
        p.KernelsGridStep = kernels_grid[2] - kernels_grid[0];
    }
    MatrixOfMatrix w(1); Matrix max_num_w(1);
    auto kernel_W = p.GenerateKernel(sigma_w, sigma_w * 3, Normal);
    w(0, 0) = kernel_W.first; max_num_w[0][0] = kernel_W.second;
    VectorOfMatrix m(1); Vector max_num_m(1);
    auto kernel_M = p.GenerateKernel(sigma_m, sigma_m * 3, Normal);
    m(0) = kernel_M.first; max_num_m[0] = kernel_M.second;
    p.M = m; p.W = w; p.MaxNumForW = max_num_w; p.MaxNumForM = max_num_m;
    return h;
}

void ref_props_free(void* handle) { delete static_cast<RefHandle*>(handle); }

// several species: MaxNumForM[i] (s), MaxNumForW[i][j] (s*s), and the tables M(i) (s*K*K), W(i,j) (s*s*K*K)
void ref_props_multi(void* handle, int* mM, int* mW, double* M, double* W)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int s = p.SpeciesNumber, K = p.KernelsGrid.size();
    for (int i = 0; i < s; i++) {
        mM[i] = (int)p.MaxNumForM[i];
        for (int x = 0; x < K; x++)
            for (int y = 0; y < K; y++) M[((size_t)i * K + x) * K + y] = p.M(i)[x][y];
        for (int j = 0; j < s; j++) {
            mW[i * s + j] = (int)p.MaxNumForW[i][j];
            for (int x = 0; x < K; x++)
                for (int y = 0; y < K; y++) W[(((size_t)i * s + j) * K + x) * K + y] = p.W(i, j)[x][y];
        }
    }
}

// ints: [species, G, K, MaxNumForM[0], MaxNumForW[0][0]]; reals: [h, cutoff, sigmaM, sigmaW, b, d, d']
void ref_props_scalars(void* handle, int* ints, double* reals)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    ints[0] = p.SpeciesNumber;
    ints[1] = p.MomentsGrid.size();
    ints[2] = p.KernelsGrid.size();
    ints[3] = (int)p.MaxNumForM[0];
    ints[4] = (int)p.MaxNumForW[0][0];
    reals[0] = p.KernelsGridStep;
    reals[1] = p.CutoffDistance;
    reals[2] = p.SigmaM[0];
    reals[3] = p.SigmaW[0][0];
    reals[4] = p.b[0];
    reals[5] = p.d[0];
    reals[6] = p.d_prime[0][0];
}

void ref_props_grids(void* handle, double* moments_grid, double* kernels_grid)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    for (int k = 0; k < p.MomentsGrid.size(); k++) moments_grid[k] = p.MomentsGrid[k];
    for (int k = 0; k < p.KernelsGrid.size(); k++) kernels_grid[k] = p.KernelsGrid[k];
}

// K*K row-major each
void ref_props_tables(void* handle, double* M, double* W)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int K = p.KernelsGrid.size();
    for (int x = 0; x < K; x++)
        for (int y = 0; y < K; y++) {
            M[(size_t)x * K + y] = p.M(0)[x][y];
            W[(size_t)x * K + y] = p.W(0, 0)[x][y];
        }
}

// Equation::operator() (Solver.cpp:431); state/deriv hold 1 + G*G doubles (layout Solver.cpp:255-262)
void ref_rhs(void* handle, const double* state, double* deriv)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int n = p.SpeciesNumber + p.SpeciesNumber * p.SpeciesNumber * p.MomentsGrid.size() * p.MomentsGrid.size();
    Equation equation(p);
    Vector out = equation(0, Vector(std::vector<CalcType>(state, state + n)), g_type);
    for (int k = 0; k < n; k++) deriv[k] = out[k];
}

// The two integrals on their own (Solver.cpp:265, 277): WC scalar and MC (G*G row-major), one species.
void ref_integrals(void* handle, const double* state, double* WC, double* MC)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int G = p.MomentsGrid.size();
    Equation equation(p);
    equation.C_extrapolation_type = SquaredN;
    Vector N(1, state[0]);
    MatrixOfMatrix C(1);
    Matrix c(G);
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) c[x][y] = state[1 + (size_t)x * G + y];
    C(0, 0) = c;
    Matrix wc = equation.CalculateIntegralOfWAndC(C, N);
    *WC = wc[0][0];
    if (MC) {
        MatrixOfMatrix mc = equation.CalculateIntegralOfMAndC(C, N);
        for (int x = 0; x < G; x++)
            for (int y = 0; y < G; y++) MC[(size_t)x * G + y] = mc(0, 0)[x][y];
    }
}

// A bounded sample of the hot loop (Solver.cpp:277-295): the reference's own per-node integral
// MakeIntegralWithPairDensity (Solver.cpp:297-316) at the nodes of the rows first, first + stride, ... -- what
// CalculateIntegralOfMAndC does for them, nothing else.  out: one row of G values per sampled row.  Returns the row count.
// (bench.py's reference arm: a full right-hand side on the synthetic grids takes tens of seconds.)
int ref_mc_rows(void* handle, const double* state, int first, int stride, double* out)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int G = p.MomentsGrid.size();
    Equation equation(p);
    equation.C_extrapolation_type = g_type;
    Matrix c(G);
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) c[x][y] = state[1 + (size_t)x * G + y];
    int rows = 0;
    for (int x = first; x < G; x += stride, rows++)
        for (int y = 0; y < G; y++)
            out[(size_t)rows * G + y] = equation.MakeIntegralWithPairDensity(x, y, p.MaxNumForM[0], p.M(0), c, state[0] * state[0]);
    return rows;
}

// The time-invariant kernel look-ups at the moment nodes (Solver.cpp:469-476, 490-497): G*G each.
void ref_kernels_at_nodes(void* handle, double* m_at_nodes, double* w_at_nodes)
{
    Properties& p = static_cast<RefHandle*>(handle)->props;
    const int G = p.MomentsGrid.size();
    Equation equation(p);
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) {
            m_at_nodes[(size_t)x * G + y] = equation.BilinearInterpolation(
                p.KernelsGrid, p.M(0), p.MomentsGrid[x], p.MomentsGrid[y], p.SigmaM[0] * 3, Zero);
            w_at_nodes[(size_t)x * G + y] = equation.BilinearInterpolation(
                p.KernelsGrid, p.W(0, 0), p.MomentsGrid[x], p.MomentsGrid[y], p.SigmaW[0][0] * 3, Zero);
        }
}

// EulerSolver::Solve (Solver.cpp:557) from the reference's initial condition
// GenerateEquilibriumValuesForNAndC(n0) (416) or from a caller-supplied state (init != NULL).
// Returns the number of stored points; the history stays in the handle.
int ref_solve(void* handle, double t_max, double dt, double n0, const double* init)
{
    QuietCout quiet;
    RefHandle* h = static_cast<RefHandle*>(handle);
    Properties& p = h->props;
    const int n = p.SpeciesNumber + p.SpeciesNumber * p.SpeciesNumber * p.MomentsGrid.size() * p.MomentsGrid.size();
    Equation equation(p);
    EulerSolver solver;
    Vector start = init ? Vector(std::vector<CalcType>(init, init + n))
                        : equation.GenerateEquilibriumValuesForNAndC(n0);
    h->results = solver.Solve(equation, g_type, t_max, dt, start, false);
    return (int)h->results.Values.size();
}

// ExtrapolationType (Solver.cpp:28) for ref_rhs / ref_solve / ref_solve_rk4: 0 Zero, 1 SquaredN, 2 Alpha, 3 Exponential
void ref_set_extrapolation(int type) { g_type = (ExtrapolationType)type; }

// RungeKutta4thSolver::Solve (Solver.cpp:597-654) -- present in the reference, never instantiated by its experiments
int ref_solve_rk4(void* handle, double t_max, double dt, double n0, const double* init)
{
    QuietCout quiet;
    RefHandle* h = static_cast<RefHandle*>(handle);
    Properties& p = h->props;
    const int n = p.SpeciesNumber + p.SpeciesNumber * p.SpeciesNumber * p.MomentsGrid.size() * p.MomentsGrid.size();
    Equation equation(p);
    RungeKutta4thSolver solver;
    Vector start = init ? Vector(std::vector<CalcType>(init, init + n))
                        : equation.GenerateEquilibriumValuesForNAndC(n0);
    h->results = solver.Solve(equation, g_type, t_max, dt, start, false);
    return (int)h->results.Values.size();
}

void ref_result_times(void* handle, double* times)
{
    RefHandle* h = static_cast<RefHandle*>(handle);
    for (int k = 0; k < h->results.Times.size(); k++) times[k] = h->results.Times[k];
}

// values/derivs: 1 + G*G doubles of stored point k (either may be NULL)
void ref_result_point(void* handle, int k, double* values, double* derivs)
{
    RefHandle* h = static_cast<RefHandle*>(handle);
    if (values) { Vector& v = h->results.Values.at(k); for (int q = 0; q < v.size(); q++) values[q] = v[q]; }
    if (derivs) { Vector& v = h->results.Derivatives.at(k); for (int q = 0; q < v.size(); q++) derivs[q] = v[q]; }
}

// The experiment menu itself, driven by scripted keys (LDM_REF_KEYS, see stubs/conio.h); the
// reference writes the literal file name "Data\VisData.csv" in the cwd (Solver.cpp:983).
int ref_run_menu(void) { return ldm_reference_main(); }

}  // extern "C"
