// CUDA kernels of the Law-Dieckmann step for sm_100a (FP64 CUDA-core work: data-dependent 4-tap
// gathers and a dense FP64 contraction -- no tensor-core shape here).
//
// S:n = "/root/reference/Euler Solver ver.5.3/Solver.cpp" line n.
#include "ldm_device.h"
#include "ldm_pointwise.cuh"

namespace ldm {

namespace {

constexpr int kWarp = 32;
constexpr int kRowsPerBlock = 8;          // y nodes per CTA in sample_sum (one warp each)

__device__ __forceinline__ double warp_sum(double v)
{
    // fixed butterfly: the same association order on every run
    #pragma unroll
    for (int off = kWarp / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ------------------------------------------------------------------------------------------------
// sample_sum: for output node (x, y), enumerate the quadrature samples (o_i, o_j) of
// MakeIntegralWithPairDensity (S:297-316).  A warp owns one y; lanes stride over o_j; the loop over
// o_i is warp-uniform.  The sample coordinates are separable -- s(x,o_i) and s(y,o_j) come from the
// same per-axis table -- and the cutoff test hypot(sx,sy) > cutoff (S:329-330) is the exact
// host-computed comparison  sy <= limit(x,o_i).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kWarp * kRowsPerBlock)
sample_sum_kernel(DevProblem P, const double* __restrict__ C, double* __restrict__ out, const double* __restrict__ scalars)
{
    const int G = P.G, n_off = P.n_off, A = P.A;
    const int x = P.rank + P.world * blockIdx.y;       // this rank's rows (all rows when world == 1)
    const int y = blockIdx.x * kRowsPerBlock + threadIdx.y;
    if (y >= G) return;
    const int lane = threadIdx.x;

    const double* lim_row = P.limit + (size_t)x * n_off;
    const double* wx_row = P.weight + (size_t)x * n_off;
    const int* ux_row = P.upper + (size_t)x * n_off;
    const double* sy_row = P.s + (size_t)y * n_off;
    const double* sx_row = P.s + (size_t)x * n_off;
    const double decay = MODE == SAMPLES_OUT_EXP ? scalars[2] : 0.0;
    const double* wy_row = P.weight + (size_t)y * n_off;
    const int* uy_row = P.upper + (size_t)y * n_off;

    // largest abscissa this y can produce: g_y + largest offset = the last entry of its row
    const double sy_far = fmax(sy_row[0], sy_row[n_off - 1]);

    double acc = 0.0;
    for (int oi = 0; oi < n_off; oi++) {
        const double lim = lim_row[oi];
        const int ux = ux_row[oi];
        if (MODE == SAMPLES_INSIDE && lim < 0.0) continue;              // nothing inside on this sample row
        if (MODE == SAMPLES_RING && (ux == 0 || lim >= sy_far)) continue;   // no sample between circle and square edge
        const double wx = wx_row[oi], ux_w = 1.0 - wx;   // a(1-w) + b w: exact at w = 0 and w = 1, like S:348-351
        const double* q_row = P.quadrant + (size_t)P.alpha[oi] * A;
        const double* c_lo = C + (size_t)row_slot(P, ux > 0 ? ux - 1 : 0) * G;
        const double* c_hi = C + (size_t)row_slot(P, ux) * G;
        for (int oj = lane; oj < n_off; oj += kWarp) {
            const double sy = sy_row[oj];
            const double m = q_row[P.alpha[oj]];
            if (MODE == SAMPLES_OUT_WEIGHT) {
                if (!(sy <= lim)) acc += m;
                continue;
            }
            if (MODE == SAMPLES_OUT_EXP) {          // S:406: exp(-1 * decay_avg * (distance - max_distance)), distance = hypot (S:329)
                if (!(sy <= lim) && m != 0.0) acc += m * exp(-1 * decay * (hypot(sx_row[oi], sy) - P.cutoff));
                continue;
            }
            const int uy = uy_row[oj];
            bool take;
            if (MODE == SAMPLES_INSIDE) take = (sy <= lim);
            else take = (uy != 0) && !(sy <= lim);                      // inside the square, outside the circle
            if (take) {
                const double wy = wy_row[oj];
                const double c00 = c_lo[uy - 1], c01 = c_lo[uy];
                const double c10 = c_hi[uy - 1], c11 = c_hi[uy];
                const double lo = c00 * ux_w + c10 * wx;
                const double hi = c01 * ux_w + c11 * wx;
                acc += m * (lo * (1.0 - wy) + hi * wy);
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)blockIdx.y * G + y] = acc;
}

// ------------------------------------------------------------------------------------------------
// WC (S:265-274) as a dot product with the precomputed operator, and the first-moment equation
// dN = N(b-d) - d' WC (S:446-451) with its Euler update and clamp (S:577-580).  One CTA,
// fixed-shape tree reduction.
// ------------------------------------------------------------------------------------------------
constexpr int kWcThreads = 1024;

__global__ void __launch_bounds__(kWcThreads)
wc_kernel(DevProblem P, const double* __restrict__ C, const double* __restrict__ N_cur, double* __restrict__ N_next,
          double* __restrict__ scalars, double dt)
{
    __shared__ double part[kWcThreads / kWarp];
    const int n = P.G * P.G;
    double acc = 0.0;
    for (int q = threadIdx.x; q < n; q += kWcThreads) {
        const int x = q / P.G, y = q - x * P.G;
        acc += P.wc_operator[q] * C[(size_t)row_slot(P, x) * P.G + y];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < kWarp) {
        double v = part[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            const double N = N_cur[0];
            const double WC = __dmul_rn(add_outside(P, v, N, C, scalars, P.wc_outside, scalars[4]), P.hh);   // S:311, S:314
            double dN = __dmul_rn(N, __dsub_rn(P.b, P.d));                          // S:446
            dN = __dsub_rn(dN, __dadd_rn(0.0, __dmul_rn(P.dd, WC)));                // S:450-451
            scalars[0] = WC;
            scalars[1] = dN;
            if (N_next) {
                double v1 = __dadd_rn(N, __dmul_rn(dN, dt));                        // S:577
                N_next[0] = v1 < 0.0 ? 0.0 : v1;                                    // S:578-580
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pointwise assembly of dC (S:459-508), in the reference's operation order with no contraction, then
// value + derivative*step and the clamp (S:577-580).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pointwise_kernel(DevProblem P, const double* __restrict__ C, const double* __restrict__ inside_sum,
                 const double* __restrict__ N_cur, const double* __restrict__ scalars, double dt,
                 double* __restrict__ deriv_out, double* __restrict__ C_next, const double* __restrict__ out_exp)
{
    const int n = local_rows(P) * P.G;
    const int ql = blockIdx.x * blockDim.x + threadIdx.x;       // shard-local index
    if (ql >= n) return;
    const int lx = ql / P.G, y = ql - lx * P.G;
    const int q = (P.rank + P.world * lx) * P.G + y;            // natural index of this node
    const double N = N_cur[0];
    const double WC = scalars[0];
    const double c = C[(size_t)(P.rank * P.rows_per_rank + lx) * P.G + y];
    const double mc = __dmul_rn(add_outside(P, inside_sum[ql], N, C, scalars, P.out_weight[q], out_exp ? out_exp[ql] : 0.0), P.hh);   // S:311, S:314
    const double dc = second_moment_rhs(P, mc, c, N, WC, P.m_nodes[q], P.w_nodes[q]);
    if (deriv_out) deriv_out[ql] = dc;
    if (C_next) {
        const double v = __dadd_rn(c, __dmul_rn(dc, dt));                           // S:577
        C_next[ql] = v < 0.0 ? 0.0 : v;                                             // S:578-580
    }
}

// ------------------------------------------------------------------------------------------------
// Ring correction of the factored path: a sparse matrix-vector product with the time-invariant operator built on
// the host (ldm_model.cpp build_ring_operator).  One warp per CSR row, fixed reduction order.
// ------------------------------------------------------------------------------------------------
constexpr int kRingWarps = 8;

template <bool BATCH>
__global__ void __launch_bounds__(kWarp * kRingWarps)
ring_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    const DevProblem& P = S.P;
    const double* __restrict__ C = S.C;
    double* __restrict__ ring = S.ring;
    const int row = blockIdx.x * kRingWarps + (threadIdx.x >> 5);
    if (row >= P.n_ring) return;
    const int lane = threadIdx.x & 31;
    const int begin = P.ring_begin[row], end = P.ring_begin[row + 1];
    double acc = 0.0;
    for (int e = begin + lane; e < end; e += kWarp) acc = fma(P.ring_weight[e], C[P.ring_node[e]], acc);
    acc = warp_sum(acc);
    if (lane == 0) ring[P.ring_point[row]] = acc;
}

// ------------------------------------------------------------------------------------------------
// Closing kernel of the factored path (see ldm_device.h).  Every CTA reduces the G per-row dot products of WC
// in the same fixed order, so all CTAs hold the same WC; block 0 also advances the first moment.
// ------------------------------------------------------------------------------------------------
template <bool BATCH>
__global__ void __launch_bounds__(256)
finish_factored_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    const DevProblem& P = S.P;
    if (BATCH && (int)(blockIdx.x * blockDim.x) >= local_rows(P) * P.G) return;      // batch grids are padded
    const double* __restrict__ C = S.C;
    const int tiles = S.tiles * S.chunks;
    const double* __restrict__ partial = S.partial;
    const double* __restrict__ ring = S.ring;
    const double* __restrict__ wc_rows = S.wc_rows;
    const double* __restrict__ N_cur = S.N_cur;
    double* __restrict__ N_next = S.N_next;
    double* __restrict__ scalars = S.scalars;
    const double dt = S.dt;
    double* __restrict__ sum_out = S.sum_out;
    double* __restrict__ deriv_out = S.deriv_out;
    double* __restrict__ C_next = S.C_next;
    __shared__ double wc_shared;
    const int G = P.G;
    if (threadIdx.x < kWarp) {
        const double v = wc_total(wc_rows, G, threadIdx.x);
        if (threadIdx.x == 0) wc_shared = v;
    }
    __syncthreads();
    const double N = N_cur[0];
    const double WC = __dmul_rn(add_outside(P, wc_shared, N, C, scalars, P.wc_outside, scalars[4]), P.hh);   // S:311, S:314
    if (threadIdx.x == 0 && blockIdx.x == 0) first_moment_update(P, N, WC, dt, scalars, N_next);

    const int n = local_rows(P) * G;
    const int ql = blockIdx.x * blockDim.x + threadIdx.x;
    if (ql < n) {
        const int lx = ql / G, y = ql - lx * G;
        const int q = (P.rank + P.world * lx) * G + y;
        double inside = 0.0;
        for (int t = 0; t < tiles; t++) inside += partial[(size_t)t * n + ql];
        inside -= ring[ql];
        if (sum_out) sum_out[ql] = inside;
        const size_t slot = (size_t)(P.rank * P.rows_per_rank + lx) * G + y;          // this node in the gathered layout
        const double c = C[slot];
        const double mc = __dmul_rn(add_outside(P, inside, N, C, scalars, P.out_weight[q], S.out_exp ? S.out_exp[ql] : 0.0), P.hh);
        const double dc = second_moment_rhs(P, mc, c, N, WC, P.m_nodes[q], P.w_nodes[q]);
        if (deriv_out) deriv_out[ql] = dc;
        double v = __dadd_rn(c, __dmul_rn(dc, dt));
        v = v < 0.0 ? 0.0 : v;
        if (C_next) C_next[ql] = v;
        for (int r = 0; r < S.peer_count; r++) S.peer_C_next[r][slot] = v;           // the all-gather: NVLink peer stores
    }
    if (S.peer_count > 0) {
        // release: the CTA's peer stores are made visible system-wide -- ONE fence per CTA, by thread 0 behind the CTA
        // barrier (a fence is cumulative over what the barrier ordered before it, and fences are the expensive part of
        // any hand-over on this chip) -- then the last CTA of the grid publishes this rank's epoch in every rank's flag array
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int arrived = atomicAdd(S.done_counter, 1u);
            if (arrived == gridDim.x - 1) {
                *S.done_counter = 0u;
                __threadfence_system();
                for (int r = 0; r < S.peer_count; r++)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(S.peer_flags[r] + P.rank), "l"(S.epoch) : "memory");
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Several species (see ldm_device.h MultiArgs).  One thread per node, all pairs; the first warps of every CTA reduce
// the pairs' WC row dots in the same fixed order, so all CTAs hold the same WC matrix; block 0 also closes the first
// moments.  Operation order of Solver.cpp:446-508, nothing contracted.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finish_multi_kernel(const MultiArgs A)
{
    const int s = A.species, G = A.G, n = G * G, pairs = s * s;
    __shared__ double wc_raw[kMaxSpecies * kMaxSpecies];
    for (int p = threadIdx.x >> 5; p < pairs; p += blockDim.x >> 5) {
        const double v = wc_total(A.pair[p].wc_rows, G, threadIdx.x & 31);
        if ((threadIdx.x & 31) == 0) wc_raw[p] = v;
    }
    __syncthreads();
    double N[kMaxSpecies], WC[kMaxSpecies * kMaxSpecies], beyond[kMaxSpecies * kMaxSpecies];
    #pragma unroll
    for (int i = 0; i < kMaxSpecies; i++) N[i] = i < s ? A.N_cur[i] : 1.0;
    #pragma unroll
    for (int p = 0; p < kMaxSpecies * kMaxSpecies; p++) {
        if (p >= pairs) { WC[p] = 0.0; beyond[p] = 0.0; continue; }
        const int i = p / s, j = p - i * s;
        const double asymptote = N[i] * N[j];                                       // S:271, S:289
        double ext = asymptote;
        if (A.extrapolation != 1 && !(fabs(A.pair[p].C[G - 1] - asymptote) < 1e-8))  // S:357; Zero / Alpha (S:361-389)
            ext = A.extrapolation == 0 ? 0.0 : A.pair[p].C[0];
        beyond[p] = ext;
        WC[p] = (wc_raw[p] + ext * A.pair[p].wc_outside) * A.hh;                    // S:311, S:314
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < s) {
        const int i = threadIdx.x;
        double dN = __dmul_rn(N[i], __dsub_rn(A.b[i], A.d[i]));                     // S:446
        double d_prime_sum = 0.0;
        for (int j = 0; j < s; j++) d_prime_sum = __dadd_rn(d_prime_sum, __dmul_rn(A.dd[i * s + j], WC[i * s + j]));   // S:448-450
        dN = __dsub_rn(dN, d_prime_sum);                                            // S:451
        A.dN_out[i] = dN;
        if (A.N_next) {
            const double v = __dadd_rn(N[i], __dmul_rn(dN, A.dt));                  // S:577
            A.N_next[i] = v < 0.0 ? 0.0 : v;                                        // S:578-580
        }
    }
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double MC[kMaxSpecies * kMaxSpecies], C[kMaxSpecies * kMaxSpecies];
    #pragma unroll
    for (int p = 0; p < kMaxSpecies * kMaxSpecies; p++) {
        if (p >= pairs) { MC[p] = 0.0; C[p] = 0.0; continue; }
        const MultiPair& P = A.pair[p];
        double inside = 0.0;
        for (int t = 0; t < P.slabs; t++) inside += P.partial[(size_t)t * n + q];
        inside -= P.ring[q];
        MC[p] = (inside + beyond[p] * P.out_weight[q]) * A.hh;
        C[p] = P.C[q];
    }
    for (int i = 0; i < s; i++)
        for (int j = 0; j < s; j++) {
            const int ij = i * s + j, ji = j * s + i;
            double r = __dadd_rn(__dmul_rn(A.b[i], MC[ij]), __dmul_rn(A.b[j], MC[ji]));             // S:464
            if (i == j) r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(2.0, A.b[i]), N[i]), A.pair[ij].m_nodes[q]));   // S:469
            r = __dsub_rn(r, __dmul_rn(A.d[i], C[ij]));                                             // S:480
            r = __dsub_rn(r, __dmul_rn(A.d[j], C[ji]));                                             // S:481
            for (int k = 0; k < s; k++) {
                r = __dsub_rn(r, __ddiv_rn(__dmul_rn(__dmul_rn(A.dd[i * s + k], WC[i * s + k]), C[ij]), N[i]));   // S:486
                r = __dsub_rn(r, __ddiv_rn(__dmul_rn(__dmul_rn(A.dd[j * s + k], WC[j * s + k]), C[ji]), N[j]));   // S:487
            }
            r = __dsub_rn(r, __dmul_rn(__dmul_rn(A.dd[ij], C[ij]), A.pair[ij].w_nodes[q]));         // S:490
            r = __dsub_rn(r, __dmul_rn(__dmul_rn(A.dd[ji], C[ji]), A.pair[ji].w_nodes[q]));         // S:498
            if (A.pair[ij].deriv_out) A.pair[ij].deriv_out[q] = r;
            if (A.pair[ij].C_next) {
                const double v = __dadd_rn(C[ij], __dmul_rn(r, A.dt));                              // S:577
                A.pair[ij].C_next[q] = v < 0.0 ? 0.0 : v;                                           // S:578-580
            }
        }
}

// One warp: lane r spins until rank r has published `epoch` (acquire, system scope).  A time budget (~9 s) turns a lost
// peer into an error status instead of a hung GPU.
__global__ void peer_wait_kernel(const unsigned long long* flags, int world, unsigned long long epoch, int* status)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long seen = 0;
    const long long t0 = clock64();
    for (unsigned int polls = 1;; polls++) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + r) : "memory");
        if (seen >= epoch) return;
        if ((polls & 0x3ffu) == 0 && clock64() - t0 > (1LL << 34)) break;
    }
    atomicExch(status, 1 + r);
}

// ------------------------------------------------------------------------------------------------
// RungeKutta4thSolver::Solve (S:597-654) on state vectors laid out like the reference's, [N | C]: the stage
// states value + (c * step / divisor) (S:633-636) and the closing combination with its clamp (S:637-645), each
// operation rounded separately as the reference's Vector operators do (T:186-265).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rk4_stage_kernel(int n, const double* __restrict__ value, const double* __restrict__ c, double step, double divisor,
                 double* __restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = __dadd_rn(value[q], __ddiv_rn(__dmul_rn(c[q], step), divisor));
}

__global__ void __launch_bounds__(256)
rk4_close_kernel(int n, const double* __restrict__ value, const double* __restrict__ c0, const double* __restrict__ c1,
                 const double* __restrict__ c2, const double* __restrict__ c3, double step, double* __restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double c = c0[q];                                                   // S:637
    c = __dadd_rn(c, __dmul_rn(c1[q], 2.0));                            // S:638
    c = __dadd_rn(c, __dmul_rn(c2[q], 2.0));                            // S:639
    c = __dadd_rn(c, c3[q]);                                            // S:640
    const double v = __dadd_rn(value[q], __ddiv_rn(__dmul_rn(c, step), 6.0));   // S:641
    out[q] = v < 0.0 ? 0.0 : v;                                         // S:643-645
}

// ------------------------------------------------------------------------------------------------
// Exponential extrapolation (S:392-408), the per-step scalars: the decay rate averaged over the four nodes before the
// last of row 0 of C, the factor C[0][G-1] - N^2 (0 when the asymptote test of S:357 holds: the reference then returns
// the asymptote), and the sum over the WC samples beyond the cutoff.  One CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
exponential_prepare_kernel(DevProblem P, const double* __restrict__ C, const double* __restrict__ N_cur, double* __restrict__ scalars)
{
    __shared__ double part[8];
    __shared__ double decay_shared;
    const double N = N_cur[0], asymptote = N * N;
    const double last = C[P.G - 1];                                     // values[0].last(): row 0 is slot 0 in every layout
    const bool at_asymptote = fabs(last - asymptote) < 1e-8;            // S:357
    if (threadIdx.x == 0) {
        double decay_avg = 0.0;
        for (int k = 1; k < 5; k++) {                                   // accuracy = 5 (S:394-404)
            double value = fabs((C[P.G - 1 - k] - asymptote) / (last - asymptote));
            value = log(value) / P.tail_gap[k - 1];
            decay_avg += value;
        }
        decay_avg = decay_avg / 4;
        decay_shared = at_asymptote ? 0.0 : decay_avg;
    }
    __syncthreads();
    const double decay = decay_shared;
    double acc = 0.0;
    for (int e = threadIdx.x; e < P.wc_out_n; e += blockDim.x) acc += P.wc_out_w[e] * exp(-1 * decay * (P.wc_out_distance[e] - P.cutoff));
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) total += part[w];
        scalars[2] = decay;
        scalars[3] = at_asymptote ? 0.0 : last - asymptote;
        scalars[4] = total;
    }
}

__global__ void record_observables_kernel(DevProblem P, const double* __restrict__ N_cur, const double* __restrict__ C,
                                          double* __restrict__ obs_row)
{
    obs_row[0] = N_cur[0];                                               // Values[k][0]   (S:1011)
    obs_row[1] = C[(size_t)row_slot(P, 0) * P.G + (P.G - 1)];            // Values[k][G] = C[0][G-1]  (S:1014)
}

}  // namespace

void launch_finish_multi(const MultiArgs& args, cudaStream_t stream)
{
    const int n = args.G * args.G;
    finish_multi_kernel<<<(n + 255) / 256, 256, 0, stream>>>(args);
}

void launch_peer_wait(const unsigned long long* flags, int world, unsigned long long epoch, int* status, cudaStream_t stream)
{
    peer_wait_kernel<<<1, 32, 0, stream>>>(flags, world, epoch, status);
}

void launch_record_observables(const DevProblem& P, const double* N_cur, const double* C, double* obs_row, cudaStream_t stream)
{
    record_observables_kernel<<<1, 1, 0, stream>>>(P, N_cur, C, obs_row);
}

void launch_rk4_stage(int n, const double* value, const double* c, double step, double divisor, double* out, cudaStream_t stream)
{
    rk4_stage_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, value, c, step, divisor, out);
}

void launch_rk4_close(int n, const double* value, const double* c0, const double* c1, const double* c2, const double* c3, double step,
                      double* out, cudaStream_t stream)
{
    rk4_close_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, value, c0, c1, c2, c3, step, out);
}

void launch_ring(const StepArgs& one, cudaStream_t stream)
{
    const int n = one.P.n_ring;
    if (n > 0) ring_kernel<false><<<(n + kRingWarps - 1) / kRingWarps, kWarp * kRingWarps, 0, stream>>>(one, nullptr);
}

void launch_ring_batch(const StepArgs* d_batch, int n, int max_ring_rows, cudaStream_t stream)
{
    static const StepArgs none{};
    if (max_ring_rows > 0)
        ring_kernel<true><<<dim3((max_ring_rows + kRingWarps - 1) / kRingWarps, 1, n), kWarp * kRingWarps, 0, stream>>>(none, d_batch);
}

void launch_finish_factored(const StepArgs& one, cudaStream_t stream)
{
    const int n = local_rows(one.P) * one.P.G;
    finish_factored_kernel<false><<<(n + 255) / 256, 256, 0, stream>>>(one, nullptr);
}

void launch_finish_factored_batch(const StepArgs* d_batch, int n, int max_points, cudaStream_t stream)
{
    static const StepArgs none{};
    finish_factored_kernel<true><<<dim3((max_points + 255) / 256, 1, n), 256, 0, stream>>>(none, d_batch);
}

namespace {
__global__ void __launch_bounds__(256) state_moves_kernel(const StateMove* __restrict__ moves, double* stage, const bool to_members)
{
    const StateMove m = moves[blockIdx.y];
    double* at = stage + m.offset;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m.count + 1; i += (long long)gridDim.x * blockDim.x) {
        double* mine = i == 0 ? m.N : m.C + (i - 1);
        if (to_members) *mine = at[i]; else at[i] = *mine;
    }
}
}  // namespace

void launch_state_moves(const StateMove* d_moves, int n, double* stage, bool to_members, cudaStream_t stream)
{
    if (n > 0) state_moves_kernel<<<dim3(8, n), 256, 0, stream>>>(d_moves, stage, to_members);
}

void launch_exponential_prepare(const DevProblem& P, const double* C, const double* N_cur, double* scalars, cudaStream_t stream)
{
    exponential_prepare_kernel<<<1, 256, 0, stream>>>(P, C, N_cur, scalars);
}

void launch_sample_sum(const DevProblem& P, SampleMode mode, const double* C, double* out, cudaStream_t stream, const double* scalars)
{
    dim3 block(kWarp, kRowsPerBlock);
    dim3 grid((P.G + kRowsPerBlock - 1) / kRowsPerBlock, local_rows(P));
    switch (mode) {
        case SAMPLES_INSIDE: sample_sum_kernel<SAMPLES_INSIDE><<<grid, block, 0, stream>>>(P, C, out, scalars); break;
        case SAMPLES_RING: sample_sum_kernel<SAMPLES_RING><<<grid, block, 0, stream>>>(P, C, out, scalars); break;
        case SAMPLES_OUT_WEIGHT: sample_sum_kernel<SAMPLES_OUT_WEIGHT><<<grid, block, 0, stream>>>(P, C, out, scalars); break;
        case SAMPLES_OUT_EXP: sample_sum_kernel<SAMPLES_OUT_EXP><<<grid, block, 0, stream>>>(P, C, out, scalars); break;
    }
}

void launch_wc(const DevProblem& P, const double* C, const double* N_cur, double* N_next, double* scalars, double dt,
               cudaStream_t stream)
{
    wc_kernel<<<1, kWcThreads, 0, stream>>>(P, C, N_cur, N_next, scalars, dt);
}

void launch_pointwise(const DevProblem& P, const double* C, const double* inside_sum, const double* N_cur,
                      const double* scalars, double dt, double* deriv_out, double* C_next, cudaStream_t stream, const double* out_exp)
{
    const int n = local_rows(P) * P.G;
    pointwise_kernel<<<(n + 255) / 256, 256, 0, stream>>>(P, C, inside_sum, N_cur, scalars, dt, deriv_out, C_next, out_exp);
}

}  // namespace ldm
