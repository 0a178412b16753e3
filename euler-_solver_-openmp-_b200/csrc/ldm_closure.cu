// The closures whose triple integral does not factorise (include/ldm.h ldm_set_closure) -- an extension: the reference
// closes the third moment with the power-1 closure, whose integral is C(xi) * WC / N with one scalar WC
// (Solver.cpp:265-274, 483-488).  The power-2 (symmetric) and power-3 closures of the same paper need, at every node xi,
//     (w*C)(xi) = h^2 sum_{xi'} w(xi') C~(xi + xi')            K(xi) = h^2 sum_{xi'} w(xi') C~(xi') C~(xi + xi')
// over the quadrature of the reference's WC integral (MakeIntegralWithPairDensity, S:297-316, extent MaxNumForW; C~ =
// bilinear C inside the cutoff, N^2 beyond: SquaredN).  The table w(xi') C~(xi') changes every step and is not separable,
// so neither the prefix sums nor a fixed contraction apply: the samples are enumerated, one warp per node, with the
// per-axis tables of the w quadrature (abscissa, interval, weight, exact cutoff threshold) laid out exactly like the m
// side's (ldm_model.cpp build_tables on a view of the problem with W in M's place).  Three kernels per right-hand side
// behind the m*C integral and the first-moment equation:
//   own_density_kernel   C~ at the (mW-1)^2 samples of the WC integral (shift (0,0)) -> the factor C~(xi') of K
//   w_sums_kernel<OWN>   the sums above at every node (w*C only where the power-2 closure asks for it)
//   closing_kernel       dC/dt in the reference's operation order with the term of S:486-487 exchanged, Euler update, clamp
// FP64 CUDA-core work: data-dependent 4-tap gathers; C (<= 2 MB) stays in L1/L2.  S:n = Solver.cpp line n.
#include "ldm_context.h"
#include "ldm_pointwise.cuh"

namespace ldm {

namespace {

constexpr int kNodesPerBlock = 8;          // one warp per node y, 8 consecutive y of one row x per CTA

__device__ __forceinline__ double bilinear_inside(const double* __restrict__ c_lo, const double* __restrict__ c_hi, int uy,
                                                  double wx, double wy)
{
    // a(1-w) + b w along x, then along y: the form of sample_sum_kernel (ldm_kernels.cu)
    const double ux_w = 1.0 - wx;
    const double lo = c_lo[uy - 1] * ux_w + c_hi[uy - 1] * wx;
    const double hi = c_lo[uy] * ux_w + c_hi[uy] * wx;
    return lo * (1.0 - wy) + hi * wy;
}

// C~ at the samples of the WC integral: (s[0][oi], s[0][oj]) -- node 0 is xi = 0, so these are (|xi'_x|, |xi'_y|)
__global__ void __launch_bounds__(256)
own_density_kernel(DevProblem W, const double* __restrict__ C, const double* __restrict__ N_cur, double* __restrict__ own)
{
    const int n_off = W.n_off, G = W.G;
    const int at = blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= n_off * n_off) return;
    const int oi = at / n_off, oj = at - oi * n_off;
    const double N = N_cur[0];
    double v = N * N;                                            // beyond the cutoff (S:329-331, SquaredN)
    if (W.s[oj] <= W.limit[oi]) {                                // row 0 of the tables: node 0
        const int ux = W.upper[oi], uy = W.upper[oj];
        v = bilinear_inside(C + (size_t)(ux - 1) * G, C + (size_t)ux * G, uy, W.weight[oi], W.weight[oj]);
    }
    own[at] = v;
}

// out[x][y] = hh * (sum over inside samples of t * bilinear(C) + N^2 * sum over outside samples of t),
// t = w (OWN false: the convolution w*C) or w * own[oi][oj] (OWN true: K)
template <bool OWN>
__global__ void __launch_bounds__(32 * kNodesPerBlock)
w_sums_kernel(DevProblem W, const double* __restrict__ C, const double* __restrict__ own, const double* __restrict__ N_cur,
              double* __restrict__ out)
{
    const int G = W.G, n_off = W.n_off, A = W.A;
    const int x = blockIdx.y;
    const int y = blockIdx.x * kNodesPerBlock + threadIdx.y;
    if (y >= G) return;
    const int lane = threadIdx.x;
    const double* lim_row = W.limit + (size_t)x * n_off;
    const double* wx_row = W.weight + (size_t)x * n_off;
    const int* ux_row = W.upper + (size_t)x * n_off;
    const double* sy_row = W.s + (size_t)y * n_off;
    const double* wy_row = W.weight + (size_t)y * n_off;
    const int* uy_row = W.upper + (size_t)y * n_off;

    double inside = 0.0, outside = 0.0;
    for (int oi = 0; oi < n_off; oi++) {
        const double lim = lim_row[oi];
        const int ux = ux_row[oi];
        const double wx = wx_row[oi];
        const double* q_row = W.quadrant + (size_t)W.alpha[oi] * A;
        const double* own_row = OWN ? own + (size_t)oi * n_off : nullptr;
        const double* c_lo = C + (size_t)(ux > 0 ? ux - 1 : 0) * G;
        const double* c_hi = C + (size_t)ux * G;
        for (int oj = lane; oj < n_off; oj += 32) {
            double t = q_row[W.alpha[oj]];
            if (OWN) t *= own_row[oj];
            if (sy_row[oj] <= lim) inside += t * bilinear_inside(c_lo, c_hi, uy_row[oj], wx, wy_row[oj]);
            else outside += t;
        }
    }
    inside = warp_sum_fixed(inside);
    outside = warp_sum_fixed(outside);
    if (lane == 0) {
        const double N = N_cur[0];
        out[(size_t)x * G + y] = __dmul_rn(__dadd_rn(inside, __dmul_rn(__dmul_rn(N, N), outside)), W.hh);
    }
}

// Solver.cpp:459-508 with the term of S:486-487 exchanged, then value + derivative * step and the clamp (S:577-580).
// Every operation rounded on its own, in the order of oracle/ldm_oracle_closure.c.
__global__ void __launch_bounds__(256)
closing_kernel(DevProblem P, const double* __restrict__ C, const double* __restrict__ inside_sum, const double* __restrict__ N_cur,
               const double* __restrict__ scalars, double dt, int closure, double w_mass, const double* __restrict__ w_conv,
               const double* __restrict__ k_int, double* __restrict__ deriv_out, double* __restrict__ C_next)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.G * P.G) return;
    const double N = N_cur[0], WC = scalars[0], c = C[q];
    const double mc = __dmul_rn(add_outside(P, inside_sum[q], N, C, scalars, P.out_weight[q], 0.0), P.hh);      // S:311, S:314
    double r = __dadd_rn(__dmul_rn(P.b, mc), __dmul_rn(P.b, mc));                           // S:464
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.b), N), P.m_nodes[q]));           // S:469
    const double death = __dmul_rn(P.d, c);
    r = __dsub_rn(r, death);                                                                // S:480
    r = __dsub_rn(r, death);                                                                // S:481
    const double cube = __dmul_rn(__dmul_rn(N, N), N);
    double triplets;                                                                        // integral of w(xi') T(xi, xi')
    if (closure == LDM_CLOSURE_POWER3) {
        triplets = __ddiv_rn(__dmul_rn(c, k_int[q]), cube);
    } else {
        double sum = __dadd_rn(__ddiv_rn(__dmul_rn(WC, c), N), __ddiv_rn(__dmul_rn(c, w_conv[q]), N));
        sum = __dadd_rn(sum, __ddiv_rn(k_int[q], N));
        sum = __dsub_rn(sum, __dmul_rn(cube, w_mass));
        triplets = __dmul_rn(0.5, sum);
    }
    const double term = __dmul_rn(P.dd, triplets);
    r = __dsub_rn(r, term);
    r = __dsub_rn(r, term);                                                                 // in place of S:486-487
    const double local = __dmul_rn(__dmul_rn(P.dd, c), P.w_nodes[q]);                       // S:490
    r = __dsub_rn(r, local);
    r = __dsub_rn(r, local);                                                                // S:498
    if (deriv_out) deriv_out[q] = r;
    if (C_next) {
        const double v = __dadd_rn(c, __dmul_rn(r, dt));                                    // S:577
        C_next[q] = v < 0.0 ? 0.0 : v;                                                      // S:578-580
    }
}

// the tables of the w quadrature on the device, once per handle
void ensure_closure_tables(ldm_context* c)
{
    if (c->closure_ready) return;
    if (c->w_seed.W.empty()) throw std::invalid_argument("the handle holds no w table (several species: not built)");
    Tables t;
    build_tables(c->w_seed.view, t);                  // the view carries W in M's place: offM / axisM are the w quadrature's
    DevProblem& w = c->closure_w;
    w = DevProblem{};
    w.G = t.G; w.n_off = t.offM.n_off; w.A = t.offM.A;
    w.world = 1; w.rank = 0; w.rows_per_rank = t.G;
    w.extrapolation = LDM_EXTRAPOLATION_SQUARED_N;
    w.cutoff = t.cutoff; w.hh = t.hh;
    w.s = c->upload(t.axisM.s);
    w.limit = c->upload(t.axisM.limit);
    w.weight = c->upload(t.axisM.weight);
    w.upper = c->upload(t.axisM.upper);
    w.alpha = c->upload(t.offM.alpha);
    w.quadrant = c->upload(t.offM.quadrant);
    double sum = 0;                                   // the reference's loop order (S:300-302, S:311)
    for (int oi = 0; oi < w.n_off; oi++)
        for (int oj = 0; oj < w.n_off; oj++) sum += t.offM.quadrant[(size_t)t.offM.alpha[oi] * w.A + t.offM.alpha[oj]];
    c->closure_w_sum = sum;
    c->d_own_density = c->alloc<double>((size_t)w.n_off * w.n_off);
    c->d_w_conv = c->alloc<double>(c->P());
    c->d_k_int = c->alloc<double>(c->P());
    c->closure_ready = true;
}

// (w*C)(xi) (when wanted) and K(xi) at the state (N_cur, C), natural layout
void closure_integrals_on_device(ldm_context* c, const double* C, const double* N_cur, bool want_w_conv, cudaStream_t s)
{
    ensure_closure_tables(c);
    const DevProblem& w = c->closure_w;
    const int samples = w.n_off * w.n_off;
    own_density_kernel<<<(samples + 255) / 256, 256, 0, s>>>(w, C, N_cur, c->d_own_density);
    const dim3 block(32, kNodesPerBlock), grid((w.G + kNodesPerBlock - 1) / kNodesPerBlock, w.G);
    if (want_w_conv) w_sums_kernel<false><<<grid, block, 0, s>>>(w, C, nullptr, N_cur, c->d_w_conv);
    w_sums_kernel<true><<<grid, block, 0, s>>>(w, C, c->d_own_density, N_cur, c->d_k_int);
    c->launches += want_w_conv ? 3 : 2;
    LDM_CUDA(cudaGetLastError());
}

}  // namespace

void closure_close(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, double dt, double* deriv_out,
                   double* C_next, cudaStream_t s)
{
    if (!deriv_out && !C_next) return;                // integrals only
    const bool power2 = c->closure == LDM_CLOSURE_POWER2_SYMMETRIC;
    closure_integrals_on_device(c, C, N_cur, power2, s);
    const int n = P.G * P.G;
    closing_kernel<<<(n + 255) / 256, 256, 0, s>>>(P, C, c->d_sum, N_cur, c->d_scalars, dt, c->closure, c->closure_w_sum * P.hh,
                                                   c->d_w_conv, c->d_k_int, deriv_out, C_next);
    c->launches += 1;
    LDM_CUDA(cudaGetLastError());
}

}  // namespace ldm

using namespace ldm;

extern "C" {

int ldm_set_closure(ldm_handle h, int closure)
{
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    if (closure != LDM_CLOSURE_POWER1 && closure != LDM_CLOSURE_POWER2_SYMMETRIC && closure != LDM_CLOSURE_POWER3)
        return fail(LDM_ERR_ARGUMENT, "unknown closure");
    if (closure == LDM_CLOSURE_POWER1) { h->closure = closure; return LDM_OK; }
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for one species");
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature run on one rank (ldm_shard_configure(h, 0, 1) to go back)");
    if (h->dev.extrapolation != LDM_EXTRAPOLATION_SQUARED_N)
        return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for SquaredN extrapolation");
    const int rc = guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        ensure_closure_tables(h);
        LDM_CUDA(cudaStreamSynchronize(h->stream));
    });
    if (rc == LDM_OK) h->closure = closure;
    return rc;
}

int ldm_get_closure(ldm_handle h) { return h ? h->closure : LDM_ERR_ARGUMENT; }

int ldm_closure_integrals(ldm_handle h, const double* state, double* w_conv_C, double* K)
{
    if (!h || !state) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for one species");
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure)");
    if (h->dev.extrapolation != LDM_EXTRAPOLATION_SQUARED_N)
        return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for SquaredN extrapolation");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        const size_t P = h->P();
        h->cur = 0;
        LDM_CUDA(cudaMemcpyAsync(h->d_N, state, sizeof(double), cudaMemcpyHostToDevice, s));
        LDM_CUDA(cudaMemcpyAsync(h->d_C[0], state + 1, P * sizeof(double), cudaMemcpyHostToDevice, s));
        closure_integrals_on_device(h, h->d_C[0], h->d_N, w_conv_C != nullptr, s);
        if (w_conv_C) LDM_CUDA(cudaMemcpyAsync(w_conv_C, h->d_w_conv, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (K) LDM_CUDA(cudaMemcpyAsync(K, h->d_k_int, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        LDM_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
