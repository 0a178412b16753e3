// Pointwise pieces shared by the closing kernels: the reference's operation order, no contraction.
// S:n = "/root/reference/Euler Solver ver.5.3/Solver.cpp" line n.
#pragma once
#include "ldm_device.h"

namespace ldm {

__device__ __forceinline__ double warp_sum_fixed(double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);   // fixed butterfly
    return v;
}

// One row of the WC functional (S:265-274): <wc_operator[x][:], c_row>, lanes striding the row, explicit fma so that
// every kernel that forms it rounds identically.  Call with a full warp; every lane returns the sum.
__device__ __forceinline__ double wc_row_dot(const double* __restrict__ op_row, const double* c_row, int G, int lane)
{
    double s = 0.0;
    for (int y = lane; y < G; y += 32) s = fma(op_row[y], c_row[y], s);
    return warp_sum_fixed(s);
}

// Sum of the G row dots, same order everywhere (one warp, lanes striding the rows).
__device__ __forceinline__ double wc_total(const double* wc_rows, int G, int lane)
{
    double s = 0.0;
    for (int r = lane; r < G; r += 32) s += wc_rows[r];
    return warp_sum_fixed(s);
}

// What the pair density is taken to be beyond the cutoff (Extrapolation, S:355-389), per step: the asymptote N^2 whenever
// C[0][G-1] is within 1e-8 of it (S:357) or the type is SquaredN (the only type any experiment passes); else 0 (Zero) or
// C[0][0] (Alpha: `2 / (grid.size() + 1)` is an integer division in the reference, S:377, so its exponential smoothing
// never moves off values[0][0]).  C: the full pair density (row 0 is slot 0 in every layout).
__device__ __forceinline__ double outside_value(const DevProblem& P, double N, const double* C)
{
    const double asymptote = N * N;
    if (P.extrapolation == 1 || P.extrapolation == 3) return asymptote;      // (Exponential goes through outside_sum)
    if (fabs(C[P.G - 1] - asymptote) < 1e-8) return asymptote;
    return P.extrapolation == 0 ? 0.0 : C[0];
}

// `inside` plus what the samples beyond the cutoff contribute, their kernel weights adding up to `weight`: a constant times
// the weight for Zero / SquaredN / Alpha; for Exponential (S:392-408) the asymptote times the weight plus
// (C[0][G-1] - N^2) * sum of kernel * exp(-decay * (distance - cutoff)), the latter sum in `exp_sum` (scalars[3] is 0 when
// the asymptote test of S:357 holds).  Every operation is rounded on its own: all kernels that close a step -- persistent,
// per-step, sharded -- go through here and must agree bit for bit.
__device__ __forceinline__ double add_outside(const DevProblem& P, double inside, double N, const double* C, const double* scalars,
                                              double weight, double exp_sum)
{
    if (P.extrapolation == 3)
        return __dadd_rn(inside, __dadd_rn(__dmul_rn(__dmul_rn(N, N), weight), __dmul_rn(scalars[3], exp_sum)));
    return __dadd_rn(inside, __dmul_rn(outside_value(P, N, C), weight));
}

// dN/dt (S:446-451) and the clamped Euler update of N (S:577-580); thread 0 of one CTA
__device__ __forceinline__ void first_moment_update(const DevProblem& P, double N, double WC, double dt, double* scalars, double* N_next)
{
    double dN = __dmul_rn(N, __dsub_rn(P.b, P.d));                          // S:446
    dN = __dsub_rn(dN, __dadd_rn(0.0, __dmul_rn(P.dd, WC)));                // S:450-451
    scalars[0] = WC;
    scalars[1] = dN;
    if (N_next) {
        const double v1 = __dadd_rn(N, __dmul_rn(dN, dt));                  // S:577
        N_next[0] = v1 < 0.0 ? 0.0 : v1;                                    // S:578-580
    }
}

__device__ __forceinline__ double second_moment_rhs(const DevProblem& P, double mc, double c, double N, double WC,
                                                    double m_node, double w_node)
{
    double r = __dadd_rn(__dmul_rn(P.b, mc), __dmul_rn(P.b, mc));                   // S:464 (i == j: both terms)
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.b), N), m_node));         // S:469
    const double death = __dmul_rn(P.d, c);
    r = __dsub_rn(r, death);                                                        // S:480
    r = __dsub_rn(r, death);                                                        // S:481
    const double closure = __ddiv_rn(__dmul_rn(__dmul_rn(P.dd, WC), c), N);         // S:486
    r = __dsub_rn(r, closure);
    r = __dsub_rn(r, closure);                                                      // S:487
    const double local = __dmul_rn(__dmul_rn(P.dd, c), w_node);                     // S:490
    r = __dsub_rn(r, local);
    r = __dsub_rn(r, local);                                                        // S:498
    return r;
}


}  // namespace ldm
