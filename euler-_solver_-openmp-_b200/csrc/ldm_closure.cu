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
//   own_density_kernel   w(xi') C~(xi') at the (mW-1)^2 samples of the WC integral (shift (0,0)): the table K convolves with
//   w_sums_kernel<BOTH>  K at every node, and w*C from the same samples where the power-2 closure asks for it; the rows of C
//                        interpolated along x are staged in shared memory once per CTA (16 nodes of one grid row)
//   closing_kernel       dC/dt in the reference's operation order with the term of S:486-487 exchanged, Euler update, clamp
// FP64 CUDA-core work: data-dependent 4-tap gathers; C (<= 2 MB) stays in L1/L2.  S:n = Solver.cpp line n.
#include "ldm_context.h"
#include "ldm_pointwise.cuh"

namespace ldm {

namespace {

constexpr int kWarpsPerBlock = 16;         // one warp per node y: 16 consecutive y of one row x per CTA
constexpr int kTileRows = 8;               // sample rows (offsets along x) whose x-interpolated C rows are staged at a time
constexpr size_t kTileBytes = 48u << 10;   // shared memory a CTA may take without opting in

// a(1-w) + b w, the form of sample_sum_kernel (ldm_kernels.cu): exact at w = 0 and w = 1, like S:348-351
__device__ __forceinline__ double lerp2(double a, double b, double w) { return a * (1.0 - w) + b * w; }

// t[oi][oj] = w(xi') C~(xi') at the samples of the WC integral: xi' = (s[0][oi], s[0][oj]) -- node 0 is xi = 0, so these
// are (|xi'_x|, |xi'_y|).  The time-varying table K(xi) is a convolution with.
__global__ void __launch_bounds__(256)
own_density_kernel(DevProblem W, const double* __restrict__ C, const double* __restrict__ N_cur, double* __restrict__ t)
{
    const int n_off = W.n_off, G = W.G;
    const int at = blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= n_off * n_off) return;
    const int oi = at / n_off, oj = at - oi * n_off;
    const double N = N_cur[0];
    double v = N * N;                                            // beyond the cutoff (S:329-331, SquaredN)
    if (W.s[oj] <= W.limit[oi]) {                                // row 0 of the tables: node 0
        const int ux = W.upper[oi], uy = W.upper[oj];
        const double wx = W.weight[oi];
        const double* c_lo = C + (size_t)row_slot(W, ux - 1) * G;
        const double* c_hi = C + (size_t)row_slot(W, ux) * G;
        v = lerp2(lerp2(c_lo[uy - 1], c_hi[uy - 1], wx), lerp2(c_lo[uy], c_hi[uy], wx), W.weight[oj]);
    }
    t[at] = W.quadrant[(size_t)W.alpha[oi] * W.A + W.alpha[oj]] * v;
}

// K(xi) -- and with BOTH also (w*C)(xi) from the same samples -- at the nodes (x, y) of one row x:
//     hh * (sum over inside samples of t * bilinear(C) + N^2 * sum over outside samples of t),   t = w C~(xi')  or  w.
// The interpolation along x is the same for every y of the row: for `rows` sample rows at a time the CTA stages
// E[oi][:] = (1 - wx) C[ux-1][:] + wx C[ux][:] in shared memory (coalesced reads of two rows of C), then every warp
// walks its y's samples of those rows: two shared-memory taps and one interpolation along y per sample instead of four
// gathers from global memory and three interpolations.  Lanes stride the offsets along y; fixed reduction order.
// The kernel is bound by the load/store data pipe (ncu at 16 384 points: 72-78 % of its peak, profiles/r02o_ncu_w_sums_g128.csv).
// What a lane needs of ITS offsets along y -- abscissa, interval, weight, table column: the same for every sample row -- is
// held in registers where the quadrature is at most 64 offsets wide (TILES = 2: every experiment of the reference's 14.4;
// 14.4a: 64.8 -> 57.6 us per step); wider quadratures read them per sample (TILES = 0): holding 5 or 8 tiles was measured
// and is slower (16 384 points: 955 against 811 us per step -- 62-100 registers per thread cost a third of the resident
// warps, and this kernel lives on them: `long_scoreboard` is its first stall reason).  Both variants add the same numbers
// in the same order.
template <bool BOTH, int TILES>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
w_sums_kernel(DevProblem W, const double* __restrict__ C, const double* __restrict__ t, const double* __restrict__ N_cur,
              int rows, double* __restrict__ k_out, double* __restrict__ w_out)
{
    extern __shared__ double tile[];                             // [rows][G]
    const int G = W.G, n_off = W.n_off, A = W.A;
    const int x = W.rank + W.world * blockIdx.y;                 // this rank's rows (all rows when world == 1)
    const int y = blockIdx.x * kWarpsPerBlock + threadIdx.y;
    const bool live = y < G;                                     // the warps beyond the row's end still build and wait
    const int lane = threadIdx.x, thread = threadIdx.y * 32 + lane;
    const double* lim_row = W.limit + (size_t)x * n_off;
    const double* wx_row = W.weight + (size_t)x * n_off;
    const int* ux_row = W.upper + (size_t)x * n_off;
    const double* sy_row = W.s + (size_t)(live ? y : 0) * n_off;
    const double* wy_row = W.weight + (size_t)(live ? y : 0) * n_off;
    const int* uy_row = W.upper + (size_t)(live ? y : 0) * n_off;

    constexpr int kHeld = TILES > 0 ? TILES : 1;
    double sy_held[kHeld], wy_held[kHeld];
    int uy_held[kHeld], column_held[kHeld];
    if (TILES > 0) {
        #pragma unroll
        for (int k = 0; k < kHeld; k++) {
            const int oj = lane + 32 * k;
            const bool mine = oj < n_off;
            sy_held[k] = mine ? sy_row[oj] : 0.0;
            wy_held[k] = mine ? wy_row[oj] : 0.0;
            uy_held[k] = mine ? uy_row[oj] : 1;
            column_held[k] = (BOTH && mine) ? W.alpha[oj] : 0;
        }
    }

    double k_inside = 0.0, k_outside = 0.0, w_inside = 0.0, w_outside = 0.0;
    for (int first = 0; first < n_off; first += rows) {
        const int n_rows = min(rows, n_off - first);
        for (int e = thread; e < n_rows * G; e += 32 * kWarpsPerBlock) {
            const int r = e / G, col = e - r * G;
            const int ux = ux_row[first + r];
            tile[e] = lerp2(C[(size_t)row_slot(W, ux > 0 ? ux - 1 : 0) * G + col], C[(size_t)row_slot(W, ux) * G + col], wx_row[first + r]);
        }
        __syncthreads();
        if (live) {
            for (int r = 0; r < n_rows; r++) {
                const int oi = first + r;
                const double lim = lim_row[oi];
                const double* t_row = t + (size_t)oi * n_off;
                const double* q_row = W.quadrant + (size_t)W.alpha[oi] * A;
                const double* E = tile + (size_t)r * G;
                if (TILES > 0) {
                    #pragma unroll
                    for (int k = 0; k < kHeld; k++) {
                        const int oj = lane + 32 * k;
                        if (oj < n_off) {
                            const double tk = t_row[oj];
                            const double tw = BOTH ? q_row[column_held[k]] : 0.0;
                            if (sy_held[k] <= lim) {
                                const double v = lerp2(E[uy_held[k] - 1], E[uy_held[k]], wy_held[k]);
                                k_inside += tk * v;
                                if (BOTH) w_inside += tw * v;
                            } else {
                                k_outside += tk;
                                if (BOTH) w_outside += tw;
                            }
                        }
                    }
                } else {
                    for (int oj = lane; oj < n_off; oj += 32) {
                        const double tk = t_row[oj];
                        const double tw = BOTH ? q_row[W.alpha[oj]] : 0.0;
                        if (sy_row[oj] <= lim) {
                            const int uy = uy_row[oj];
                            const double v = lerp2(E[uy - 1], E[uy], wy_row[oj]);
                            k_inside += tk * v;
                            if (BOTH) w_inside += tw * v;
                        } else {
                            k_outside += tk;
                            if (BOTH) w_outside += tw;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    k_inside = warp_sum_fixed(k_inside);
    k_outside = warp_sum_fixed(k_outside);
    if (BOTH) { w_inside = warp_sum_fixed(w_inside); w_outside = warp_sum_fixed(w_outside); }
    if (live && lane == 0) {
        const double N = N_cur[0], asymptote = __dmul_rn(N, N);
        k_out[(size_t)blockIdx.y * G + y] = __dmul_rn(__dadd_rn(k_inside, __dmul_rn(asymptote, k_outside)), W.hh);      // shard-local
        if (BOTH) w_out[(size_t)blockIdx.y * G + y] = __dmul_rn(__dadd_rn(w_inside, __dmul_rn(asymptote, w_outside)), W.hh);
    }
}

int g_forced_tiles = -1;                   // ldm_tuning_set("closure_tiles"): tests reach every variant on small grids

template <bool BOTH>
void launch_w_sums(const DevProblem& w, const double* C, const double* t, const double* N_cur, double* k_out, double* w_out, cudaStream_t s)
{
    const int rows = (int)std::min<size_t>(kTileRows, kTileBytes / ((size_t)w.G * sizeof(double)));
    if (rows < 1) throw std::invalid_argument("a row of C does not fit the shared-memory tile of the closure integrals");
    const dim3 block(32, kWarpsPerBlock), grid((w.G + kWarpsPerBlock - 1) / kWarpsPerBlock, local_rows(w));
    const size_t smem = (size_t)rows * w.G * sizeof(double);
    int tiles = w.n_off <= 64 ? 2 : 0;
    if (g_forced_tiles >= 0) tiles = g_forced_tiles * 32 >= w.n_off ? g_forced_tiles : 0;
    if (tiles == 2) w_sums_kernel<BOTH, 2><<<grid, block, smem, s>>>(w, C, t, N_cur, rows, k_out, w_out);
    else w_sums_kernel<BOTH, 0><<<grid, block, smem, s>>>(w, C, t, N_cur, rows, k_out, w_out);
}

// Solver.cpp:459-508 with the term of S:486-487 exchanged, then value + derivative * step and the clamp (S:577-580).
// Every operation rounded on its own, in the order of oracle/ldm_oracle_closure.c.
__global__ void __launch_bounds__(256)
closing_kernel(DevProblem P, const double* __restrict__ C, const double* __restrict__ inside_sum, const double* __restrict__ N_cur,
               const double* __restrict__ scalars, double dt, int closure, double w_mass, const double* __restrict__ w_conv,
               const double* __restrict__ k_int, double* __restrict__ deriv_out, double* __restrict__ C_next)
{
    const int ql = blockIdx.x * blockDim.x + threadIdx.x;        // shard-local index: inside_sum, w_conv, k_int, deriv_out, C_next
    if (ql >= local_rows(P) * P.G) return;
    const int lx = ql / P.G, y = ql - lx * P.G;
    const int q = (P.rank + P.world * lx) * P.G + y;             // natural index of this node: the time-invariant tables
    const double N = N_cur[0], WC = scalars[0];
    const double c = C[(size_t)(P.rank * P.rows_per_rank + lx) * P.G + y];                  // gathered layout
    const double mc = __dmul_rn(add_outside(P, inside_sum[ql], N, C, scalars, P.out_weight[q], 0.0), P.hh);     // S:311, S:314
    double r = __dadd_rn(__dmul_rn(P.b, mc), __dmul_rn(P.b, mc));                           // S:464
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.b), N), P.m_nodes[q]));           // S:469
    const double death = __dmul_rn(P.d, c);
    r = __dsub_rn(r, death);                                                                // S:480
    r = __dsub_rn(r, death);                                                                // S:481
    const double cube = __dmul_rn(__dmul_rn(N, N), N);
    double triplets;                                                                        // integral of w(xi') T(xi, xi')
    if (closure == LDM_CLOSURE_POWER3) {
        triplets = __ddiv_rn(__dmul_rn(c, k_int[ql]), cube);
    } else {
        double sum = __dadd_rn(__ddiv_rn(__dmul_rn(WC, c), N), __ddiv_rn(__dmul_rn(c, w_conv[ql]), N));
        sum = __dadd_rn(sum, __ddiv_rn(k_int[ql], N));
        sum = __dsub_rn(sum, __dmul_rn(cube, w_mass));
        triplets = __dmul_rn(0.5, sum);
    }
    const double term = __dmul_rn(P.dd, triplets);
    r = __dsub_rn(r, term);
    r = __dsub_rn(r, term);                                                                 // in place of S:486-487
    const double local = __dmul_rn(__dmul_rn(P.dd, c), P.w_nodes[q]);                       // S:490
    r = __dsub_rn(r, local);
    r = __dsub_rn(r, local);                                                                // S:498
    if (deriv_out) deriv_out[ql] = r;
    if (C_next) {
        const double v = __dadd_rn(c, __dmul_rn(r, dt));                                    // S:577
        C_next[ql] = v < 0.0 ? 0.0 : v;                                                     // S:578-580
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
void closure_integrals_on_device(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, bool want_w_conv, cudaStream_t s)
{
    ensure_closure_tables(c);
    DevProblem& w = c->closure_w;
    w.world = P.world; w.rank = P.rank; w.rows_per_rank = P.rows_per_rank;      // xi-sharding: C in gathered layout, this rank's rows
    const int samples = w.n_off * w.n_off;
    own_density_kernel<<<(samples + 255) / 256, 256, 0, s>>>(w, C, N_cur, c->d_own_density);
    if (want_w_conv) launch_w_sums<true>(w, C, c->d_own_density, N_cur, c->d_k_int, c->d_w_conv, s);
    else launch_w_sums<false>(w, C, c->d_own_density, N_cur, c->d_k_int, nullptr, s);
    c->launches += 2;
    LDM_CUDA(cudaGetLastError());
}

}  // namespace

void closure_force_tiles(int tiles) { g_forced_tiles = (tiles == 0 || tiles == 2) ? tiles : -1; }

void closure_close(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, double dt, double* deriv_out,
                   double* C_next, cudaStream_t s)
{
    if (!deriv_out && !C_next) return;                // integrals only
    const bool power2 = c->closure == LDM_CLOSURE_POWER2_SYMMETRIC;
    closure_integrals_on_device(c, P, C, N_cur, power2, s);
    const int n = local_rows(P) * P.G;
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
    if (h->dev.extrapolation != LDM_EXTRAPOLATION_SQUARED_N)
        return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for SquaredN extrapolation");
    if (h->peer_connected) return fail(LDM_ERR_UNSUPPORTED, "the fused exchange (ldm_shard_peer_*) carries the reference's closure only: use ldm_shard_step");
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
        closure_integrals_on_device(h, h->dev, h->d_C[0], h->d_N, w_conv_C != nullptr, s);
        if (w_conv_C) LDM_CUDA(cudaMemcpyAsync(w_conv_C, h->d_w_conv, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (K) LDM_CUDA(cudaMemcpyAsync(K, h->d_k_int, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        LDM_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
