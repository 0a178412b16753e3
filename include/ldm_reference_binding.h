/* ldm_reference_binding.h -- the reference-side binding of libldm_b200.so: what a maintainer of
 * Kintobor1370/Euler-_Solver_-OpenMP- ("Euler Solver ver.5.3") adds to Solver.cpp so that its two seams
 *
 *     Vector        Equation::operator()(CalcType times, Vector vals, ExtrapolationType chosen_type)       Solver.cpp:431
 *     SolverResults EulerSolver::Solve(Equation, ExtrapolationType, CalcType time_max, CalcType step,
 *                                      Vector initial_value, bool debug_print_enabled)                     Solver.cpp:557
 *
 * run on the GPU.  The header is C++ and header-only; it is written against the reference's own containers (types.h:
 * Vector, Matrix, VectorOfMatrix, MatrixOfMatrix) and its Properties / SolverResults structs WITHOUT naming them --
 * everything is a template -- so it can be included right under `#include "types.h"` (Solver.cpp:14), before those
 * structs are defined.  The patch to Solver.cpp is four lines (applied to the reference's own file
 * by oracle/patch_binding.py, compiled by oracle/Makefile into oracle/_ref/solver_bound and run through the reference's
 * own menu by tests/test_reference_binding.py):
 *
 *     #include "ldm_reference_binding.h"                                                    // under Solver.cpp:14
 *     Properties& LdmProps() { return props; }                                              // class Equation, under Solver.cpp:413
 *     return ldm_binding::Rhs(props, vals, (int)chosen_type);                               // first line of Equation::operator(), Solver.cpp:432
 *     return ldm_binding::Solve(solver_results, current_equation.LdmProps(), (int)type, time_max, step,
 *                               initial_value, debug_print_enabled);                        // first line of EulerSolver::Solve, Solver.cpp:565
 *
 * One species or several (the state vector keeps the reference's layout, Solver.cpp:255-262).  Errors keep the
 * reference's behaviour: it asserts, so does the binding.
 */
#ifndef LDM_REFERENCE_BINDING_H
#define LDM_REFERENCE_BINDING_H

#include <cassert>
#include <cstdio>
#include <iostream>
#include <vector>

#include "ldm.h"

namespace ldm_binding {

/* the flat view of struct Properties (Solver.cpp:40-60) that ldm_create takes; `keep` owns the arrays */
struct Flat {
    ldm_problem problem;
    std::vector<double> grid, kgrid, M, W, sigma_m, sigma_w, b, d, d_prime;
    std::vector<int> mM, mW;
};

template <class Props>
inline void Flatten(Props& p, Flat& f)
{
    const int s = p.SpeciesNumber, G = p.MomentsGrid.size(), K = p.KernelsGrid.size();
    f.grid = p.MomentsGrid.get_vector();
    f.kgrid = p.KernelsGrid.get_vector();
    f.M.assign((size_t)s * K * K, 0.0);
    f.W.assign((size_t)s * s * K * K, 0.0);
    for (int i = 0; i < s; i++) {
        for (int x = 0; x < K; x++)
            for (int y = 0; y < K; y++) f.M[((size_t)i * K + x) * K + y] = p.M(i)[x][y];
        for (int j = 0; j < s; j++)
            for (int x = 0; x < K; x++)
                for (int y = 0; y < K; y++) f.W[(((size_t)i * s + j) * K + x) * K + y] = p.W(i, j)[x][y];
    }
    for (int i = 0; i < s; i++) {
        f.mM.push_back((int)p.MaxNumForM[i]);
        f.sigma_m.push_back(p.SigmaM[i]);
        f.b.push_back(p.b[i]);
        f.d.push_back(p.d[i]);
        for (int j = 0; j < s; j++) {
            f.mW.push_back((int)p.MaxNumForW[i][j]);
            f.sigma_w.push_back(p.SigmaW[i][j]);
            f.d_prime.push_back(p.d_prime[i][j]);
        }
    }
    ldm_problem& q = f.problem;
    q = ldm_problem();
    q.species = s; q.G = G; q.K = K;
    q.mM = f.mM[0]; q.mW = f.mW[0];
    q.h = p.KernelsGridStep; q.cutoff = p.CutoffDistance;
    q.sigma_m = f.sigma_m[0]; q.sigma_w = f.sigma_w[0];
    q.b = f.b[0]; q.d = f.d[0]; q.d_prime = f.d_prime[0];
    q.moments_grid = f.grid.data(); q.kernels_grid = f.kgrid.data(); q.M = f.M.data(); q.W = f.W.data();
    if (s > 1) {
        q.mM_of = f.mM.data(); q.mW_of = f.mW.data(); q.sigma_m_of = f.sigma_m.data(); q.sigma_w_of = f.sigma_w.data();
        q.b_of = f.b.data(); q.d_of = f.d.data(); q.d_prime_of = f.d_prime.data();
    }
}

inline void Check(int code, const char* what)
{
    if (code != LDM_OK) std::fprintf(stderr, "libldm_b200: %s: %s\n", what, ldm_last_error());
    assert(code == LDM_OK);                     /* the reference has no error channel but assert */
    (void)what;
}

template <class Props>
inline ldm_handle Create(Props& p, int chosen_type)
{
    Flat flat;
    Flatten(p, flat);
    ldm_handle h = 0;
    Check(ldm_create(&flat.problem, /*device*/ 0, &h), "ldm_create");
    Check(ldm_set_extrapolation(h, chosen_type), "ldm_set_extrapolation");     /* enum ExtrapolationType, Solver.cpp:28: same numbering */
    return h;
}

/* Equation::operator() (Solver.cpp:431-513) */
template <class Props, class Vec>
inline Vec Rhs(Props& p, Vec& vals, int chosen_type)
{
    ldm_handle h = Create(p, chosen_type);
    std::vector<double> in = vals.get_vector(), out(in.size());
    Check(ldm_rhs(h, in.data(), out.data()), "ldm_rhs");
    ldm_destroy(h);
    return Vec(out);
}

/* EulerSolver::Solve (Solver.cpp:557-592): the whole loop on the device; SolverResults filled as the reference fills it */
template <class Results, class Props, class Vec>
inline Results Solve(Results& results, Props& p, int chosen_type, double time_max, double step, Vec& initial_value, bool debug_print_enabled)
{
    ldm_handle h = Create(p, chosen_type);
    const int n = ldm_count_points(time_max, step);
    std::vector<double> init = initial_value.get_vector();
    const size_t width = init.size();
    std::vector<double> t((size_t)n), v((size_t)n * width), d((size_t)n * width);
    int got = 0;
    Check(ldm_solve(h, init.data(), step, time_max, LDM_RECORD_ALL, t.data(), v.data(), d.data(), &got), "ldm_solve");
    ldm_destroy(h);
    for (int k = 0; k < got; k++) {
        results.Times.add(t[(size_t)k]);
        results.Values.emplace_back(Vec(std::vector<double>(v.begin() + (size_t)k * width, v.begin() + (size_t)(k + 1) * width)));
        results.Derivatives.emplace_back(Vec(std::vector<double>(d.begin() + (size_t)k * width, d.begin() + (size_t)(k + 1) * width)));
        if (debug_print_enabled) std::cout << "   t = " << t[(size_t)k] << ";   N = " << v[(size_t)k * width] << '\n';   /* Solver.cpp:571-586 */
    }
    return results;
}

}  /* namespace ldm_binding */

#endif
