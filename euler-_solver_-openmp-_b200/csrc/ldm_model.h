// Host-side model setup for the B200 Law-Dieckmann step: everything that is time-invariant.
//
// Two layers:
//   * generate_problem(): the reference's Properties::Generate / GenerateGrids / GenerateKernel
//     (Solver.cpp:63-241) for one species, reproduced bit for bit on the host (same libm calls, same
//     operation order; built with -ffp-contract=off).
//   * build_tables(): what the CUDA kernels consume -- per-axis sample tables (abscissa, interval,
//     weight, cutoff threshold), the odd-node quadrants of the kernel tables, the kernel values at the
//     moment nodes, and the linear operator that turns C into the scalar integral WC.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include "../../include/ldm.h"

namespace ldm {

struct OwnedProblem {             // backing store of an ldm_problem handed out by ldm_problem_generate(_multi)
    ldm_problem view{};
    std::vector<double> moments_grid, kernels_grid, M, W;
    std::vector<int> mM_of, mW_of;                                     // species > 1
    std::vector<double> sigma_m_of, sigma_w_of, b_of, d_of, d_prime_of;
};

// Properties::Generate (Solver.cpp:186-241), one species.  Throws std::invalid_argument on the
// reference's assert conditions (S:77, S:88, S:140).
void generate_problem(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G,
                      OwnedProblem& out);

// The same for several species (Solver.cpp:186-241 with num > 1): grids from the largest sigmas, one W table per pair,
// one M table per species.
void generate_problem_multi(int species, const double* sigma_w, const double* sigma_m, const double* b, const double* d,
                            const double* d_prime, int forced_G, OwnedProblem& out);

// A one-species view of species pair (i, j) of a multi-species problem: M = M(i), W = W(i,j), rates of species i.
ldm_problem species_view(const ldm_problem& p, int i, int j);

// Properties::CheckKernels (Solver.cpp:147-183)
void check_kernels(const ldm_problem& p, double& integral_w, double& integral_m);

// Signed quadrature offsets of MakeIntegralWithPairDensity (Solver.cpp:300-308) for one kernel.
struct OffsetSet {
    int n_off = 0;                 // number of signed offsets per axis: m-1 (m odd) or 1 (m == 2)
    int A = 0;                     // distinct kernel nodes used: (m-1)/2, or 1
    std::vector<double> offset;    // n_off: +kappa[i+1] (i >= 0) or -kappa[-i-1] (i < 0), in the reference's loop order
    std::vector<int> node;         // n_off: kernel node |i+1|
    std::vector<int> alpha;        // n_off: compact index of that node in the quadrant table (0..A-1)
    std::vector<double> quadrant;  // A*A: kernel[node_a][node_b], row-major
};

// Per moment node x and signed offset o (row-major [G][n_off]).
struct AxisTable {
    std::vector<double> s;         // sample abscissa |g_x + offset|  (S:307)
    std::vector<int> upper;        // interval upper index i of S:335 (first i >= 1 with s <= g[i]); 0 when s > cutoff
    std::vector<double> weight;    // (s - g[i-1]) / (g[i] - g[i-1])  (S:344)
    std::vector<double> limit;     // largest partner abscissa t among ALL samples with hypot(s, t) <= cutoff
                                   // (S:329-330 decided by the host's hypot); -1 when none
};

struct Tables {
    int G = 0, K = 0;
    double h = 0, hh = 0, cutoff = 0;
    double b = 0, d = 0, dd = 0;
    OffsetSet offM;                // the m-kernel quadrature
    AxisTable axisM;               // ... on the moment nodes
    double s_max = 0;              // largest abscissa value that is still <= cutoff (or cutoff)
    std::vector<double> far_inside;    // G: largest sample abscissa of a node that is still inside the grid (-1: none)
    std::vector<double> m_nodes;   // G*G: BilinearInterpolation(KernelsGrid, M, g_x, g_y, 3 sigma_M, Zero)  (S:469)
    std::vector<double> w_nodes;   // G*G: same for W (S:490)
    std::vector<double> wc_operator;   // G*G: WC = h^2 * (sum wc_operator .* C + N^2 * wc_outside)   (S:265-274)
    double wc_outside = 0;
    std::vector<double> wc_out_w, wc_out_distance;   // the samples behind wc_outside one by one: weight and hypot(sx, sy) (S:329),
                                                     // for the one extrapolation type that looks at the distance (Exponential, S:392-408)
    std::string census;            // diagnostics: how close any sample comes to the cutoff circle
};

void build_tables(const ldm_problem& p, Tables& t);

// Structure of the m-kernel quadrant that the prefix-sum path (ldm_scan.cu) exploits.  A truncated Gaussian
// (Solver.cpp:113-123) is separable inside its disc: Q[beta][alpha] = q[beta] q[alpha] for alpha <= a_star[beta] and 0
// beyond, with a_star non-increasing in beta (the quarter circle r <= 3 sigma as a staircase).  Then
//     H_x[beta][:] = sum_alpha Q[beta][alpha] E_x[alpha][:] = q[beta] * (prefix sum over alpha <= a_star[beta] of q[alpha] E_x[alpha][:])
// and the dense contraction collapses to one running sum along alpha.  detect_separable() verifies the structure on the
// table it is given (any K x K table may arrive through ldm_create); `ok` is false when it does not hold to `tolerance`
// relative, and the dense contraction is used instead.
struct SeparableKernel {
    bool ok = false;
    double max_rel_dev = 0;        // max |Q - q q^T| / (q q^T) over the non-zero entries
    std::vector<double> q;         // A
    std::vector<int> a_star;       // A: last non-zero column of row beta (-1: empty row)
    std::vector<int> beta_lo, beta_hi;   // A: rows beta in [beta_lo[a], beta_hi[a]) have a_star[beta] == a
};
void detect_separable(const OffsetSet& off, double tolerance, SeparableKernel& out);

// y-side work list of the prefix-sum path for column chunks of width bn.  For every (chunk, block of 32 consecutive
// nodes y) the signed offsets o whose interpolation pair (upper-1, upper) touches a column of the chunk form up to
// kScanRanges intervals; the intervals of a chunk are cut into PIECES of about equal length, dealt out to the warps of a
// CTA in order (the far chunks of a stretched grid own most pairs of only a few y blocks: whole blocks per warp would
// leave most warps idle).  Every piece belongs to one y block; the pieces of a block and the pieces of a warp are
// contiguous in the list.
constexpr int kScanRanges = 3;
struct ScanPieces {
    std::vector<int> chunk_begin;      // chunks + 1: first piece of each chunk
    std::vector<int> piece;            // 4 ints per piece: y block, first offset, end offset, 0
    std::vector<int> warp_begin;       // [chunks][warps + 1], relative to the chunk's first piece
    std::vector<int> block_begin;      // [chunks][ceil(G/32) + 1], relative to the chunk's first piece
    int max_pieces = 0;                // largest piece count of a chunk (sizes the per-piece partial sums)
};
void build_scan_pieces(const Tables& t, int bn, int warps, ScanPieces& out);

// The exact correction of the factored path as a sparse linear operator (time-invariant).  The factored sum
// covers every sample inside the grid square; the reference replaces the samples beyond the cutoff circle by
// N^2 (Solver.cpp:329-331), so for each node that owns such "ring" samples
//     ring[node] = sum_e weight[e] * C[node_index[e]]          (CSR row of that node)
// collects m * (bilinear weights) of exactly those samples.  The far part of the moment grid is coarse, so the
// ~10^6..10^8 ring samples collapse onto few distinct grid nodes (141 k entries for the 16 384-point grid).
struct RingOperator {
    std::vector<int> point;        // per CSR row: local_row * G + y
    std::vector<int> row_begin;    // rows + 1
    std::vector<int> grid_row_begin;   // local moment-grid rows + 1: first CSR row of each
    std::vector<int> node;         // C index in gathered layout: row_slot(p) * G + q
    std::vector<double> weight;
    long long samples = 0;         // ring samples enumerated
};
void build_ring_operator(const Tables& t, int rank, int world, RingOperator& out);

// Equation::BilinearInterpolation (Solver.cpp:318-353) with Zero / SquaredN extrapolation (355-373)
double bilinear(const double* grid, int n, const double* values, double x, double y, double max_distance,
                bool squared_n, double asymptotic_value);

}  // namespace ldm
