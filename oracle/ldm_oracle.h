/* TEST INFRASTRUCTURE ONLY -- the CPU oracle of the Law-Dieckmann hot path.
 *
 * A plain-C restatement of the reference's algorithm (one species), operation for operation, so
 * that it rounds exactly like the reference's `g++ -O3 -fopenmp` x86-64 build.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (euler_solver_openmp_b200/) never does.
 *
 * PARITY PINNING: the reference's own tests hold no vectors for this path,
 * so the oracle is pinned against the reference itself, compiled unmodified from
 * /root/reference into oracle/_ref/ (oracle/Makefile) -- bitwise on grids, kernel tables, RHS and
 * whole trajectories (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 * Citations: S:n = "/root/reference/Euler Solver ver.5.3/Solver.cpp" line n, T:n = types.h line n.
 */
#ifndef LDM_ORACLE_H
#define LDM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldmo_problem {
    int G;                /* MomentsGrid.size()            S:57  */
    int K;                /* KernelsGrid.size()            S:59  */
    int mM, mW;           /* MaxNumForM[0], MaxNumForW[0][0]  S:51-52 */
    double h;             /* KernelsGridStep               S:60  */
    double cutoff;        /* CutoffDistance                S:58  */
    double sigmaM, sigmaW, b, d, dd;
    double* moments_grid; /* G   */
    double* kernels_grid; /* K   */
    double* M;            /* K*K row-major, M(0)           S:48  */
    double* W;            /* K*K row-major, W(0,0)         S:49  */
} ldmo_problem;

/* Properties::Generate for one species (S:186-241).  forced_G > 0 pins num_for_moments and skips
 * the search loop S:70-74 (the synthetic-grid configurations); forced_G <= 0 is the reference. */
ldmo_problem* ldmo_generate(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G);
void ldmo_free(ldmo_problem* p);

/* Equation::BilinearInterpolation + Extrapolation (S:318-410); extrapolation: 0 Zero, 1 SquaredN. */
double ldmo_bilinear(const double* grid, int n, const double* values, double x, double y,
                     double max_distance, int extrapolation, double asymptotic_value);

/* MakeIntegralWithPairDensity (S:297-316) */
double ldmo_pair_integral(const ldmo_problem* p, int x_shift, int y_shift, int max_non_zero_num,
                          const double* kernel, const double* C, double asymptotic_value);

/* CalculateIntegralOfWAndC / CalculateIntegralOfMAndC (S:265-295); MC may be NULL */
void ldmo_integrals(const ldmo_problem* p, const double* state, double* WC, double* MC);

/* the time-invariant kernel look-ups at the moment nodes (S:469-476, S:490-497); G*G each */
void ldmo_kernels_at_nodes(const ldmo_problem* p, double* m_at_nodes, double* w_at_nodes);

/* Equation::operator() (S:431-513); state, deriv: 1 + G*G doubles */
void ldmo_rhs(const ldmo_problem* p, const double* state, double* deriv);

/* number of stored points of EulerSolver::Solve's `time += step; while (time < time_max)` (S:565-589) */
int ldmo_count_points(double t_max, double dt);

/* EulerSolver::Solve (S:557-592).  init: 1 + G*G doubles.  times: n_points.  values/derivs: either
 * NULL, or n_points*(1+G*G) (full history) when full_history != 0, or 1+G*G (last point only).
 * observables (may be NULL): n_points*2 = N and C[0][G-1] at every point (S:1011-1015).
 * Returns the number of points. */
int ldmo_solve(const ldmo_problem* p, double t_max, double dt, const double* init, int full_history,
               double* times, double* values, double* derivs, double* observables);

/* the ExtrapolationType Equation::operator() is called with for C (S:28, S:431): 0 Zero, 1 SquaredN (default; the only
 * one any experiment passes), 2 Alpha, 3 Exponential.  Applies to ldmo_integrals / ldmo_rhs / the solvers. */
void ldmo_set_extrapolation(int type);

/* RungeKutta4thSolver::Solve (S:597-654): full history, values/derivs n_points*(1+G*G) (either may be NULL) */
int ldmo_count_points_rk4(double t_max, double dt);
int ldmo_solve_rk4(const ldmo_problem* p, double t_max, double dt, const double* init, double* times, double* values, double* derivs);

/* ---- several species (S:265-295, S:446-511 with SpeciesNumber > 1).  State: [N_0..N_{s-1} | C_00 | C_01 | ...], each C_ij
 * G*G row-major (S:255-262).  Tables and grids as the reference's Properties holds them (shared grids; M(i): s tables,
 * W(i,j): s*s tables, K*K row-major each). ---- */
typedef struct ldmo_multi {
    int species, G, K;
    double h, cutoff;
    const int* mM;             /* s     MaxNumForM[i]    */
    const int* mW;             /* s*s   MaxNumForW[i][j] */
    const double* sigmaM;      /* s */
    const double* sigmaW;      /* s*s */
    const double* b;           /* s */
    const double* d;           /* s */
    const double* dd;          /* s*s */
    const double* moments_grid;
    const double* kernels_grid;
    const double* M;           /* s * K*K */
    const double* W;           /* s*s * K*K */
} ldmo_multi;
void ldmo_rhs_multi(const ldmo_multi* p, const double* state, double* deriv);
/* EulerSolver::Solve (S:557-592) on such a state: full history, values n_points * width; returns the number of points */
int ldmo_solve_multi(const ldmo_multi* p, double t_max, double dt, const double* init, double* times, double* values);

/* ---- EXTENSION, parity unpinned beyond the uniform state (ldm_oracle_closure.c): the integrals that do not factorise --
 * (w*C)(xi) and K(xi) = integral of w(xi') C(xi') C(xi + xi') at every node (G*G each, either may be NULL) -- and the
 * right-hand side / Euler solve with the closure term of S:486-487 exchanged: closure 1 power-1 (the reference; ldmo_rhs
 * bit for bit), 2 power-2 symmetric, 3 power-3.  SquaredN beyond the cutoff. ---- */
double ldmo_w_mass(const ldmo_problem* p);
void ldmo_closure_integrals(const ldmo_problem* p, const double* state, double* w_conv_C, double* K_out);
void ldmo_rhs_closure(const ldmo_problem* p, const double* state, int closure, double* deriv);
int ldmo_solve_closure(const ldmo_problem* p, int closure, double t_max, double dt, const double* init, double* times, double* values);

/* threads the integrals' outer loop may use (results are identical for any count) */
void ldmo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
