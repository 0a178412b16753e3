// The factored evaluation of the m*C integral (Solver.cpp:277-316) for sm_100a.
//
// The reference evaluates, for every node (x,y), a double sum over quadrature offsets (o_i, o_j) of
// m[a_i][a_j] * bilinear(C)(s(x,o_i), s(y,o_j)).  Bilinear interpolation is a product of two 1-D
// interpolations and the sample abscissae are separable, so with L[n][o] the 2-tap interpolation row of
// node n / offset o the sum over all samples inside the grid square is
//
//     ext[x][y] = sum_{beta} sum_{o_j in beta} L[y][o_j] . H_x[beta][:],      H_x = Q . E_x,
//     E_x[alpha][q] = sum_{o_i in alpha} (L[x][o_i] . C[:, q])
//
// with Q the odd-node quadrant of the kernel table (A x A, symmetric) and "o in alpha" the (at most two)
// signed offsets that share kernel node alpha.  H_x = Q E_x is a dense FP64 GEMM (A x A x G per row x):
// P*A*A fused multiply-adds per step instead of ~2 P (2A)^2 for the sample-by-sample sum.  The samples that
// lie inside the square but outside the cutoff circle (where the reference substitutes N^2, Solver.cpp:329-331)
// are removed afterwards by the exact ring correction (sample_sum<SAMPLES_RING>), so
//     inside_sum = ext - ring.
//
// One CTA owns one row x, one tile of BM beta-rows and one chunk of BN columns q ("slab" = tile x chunk): it
// generates its E chunk on the fly from C (4 taps per element), accumulates the BM x BN tile of H in registers
// (plus the single halo column q0-1 that the first tap pairs of the chunk need), stages it in shared memory and
// contracts it with the y-side interpolation rows.  Slabs are independent; finish_factored_kernel adds them up.
#include <cooperative_groups.h>

#include "../../include/ldm.h"
#include "ldm_device.h"
#include "ldm_pointwise.cuh"

namespace ldm {

// Optional phase timeline (build with -DLDM_PHASE_CLOCKS): thread 0 of every CTA stores clock64() at the phase
// boundaries into the scalars array's debug tail; read back by tools/phase_clocks.py.  Off in the product build.
#ifdef LDM_PHASE_CLOCKS
__device__ long long g_phase_clock[148 * 4][12];
#define PHASE_MARK(i) do { if (threadIdx.x == 0) g_phase_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 4)][i] = clock64(); } while (0)
#define PHASE_WAIT_BEGIN() const long long wait_t0_ = clock64()
#define PHASE_WAIT_END(slot, who) do { if (threadIdx.x == (who)) g_phase_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 4)][slot] += clock64() - wait_t0_; } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_WAIT_BEGIN() do { } while (0)
#define PHASE_WAIT_END(slot, who) do { } while (0)
#endif

namespace {

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);   // fixed butterfly
    return v;
}

template <int BM, int BN, int TM, int TN, int BK, bool MMA = false>
struct FactoredConfig {
    static constexpr int TYN = BM / TM;            // thread rows
    static constexpr int TXN = BN / TN;            // thread columns
    static constexpr int NT = TYN * TXN;           // threads per CTA
    static constexpr int HS_LD = BN + 2;           // H staging row: 1 halo column + BN + 1 pad
    // operand tiles in shared memory.  Thread-tile path: Q as [k][BM], E as [k][BN].  Tensor-core path (mma.sync
    // m8n8k4 f64): Q as [beta][k] and E as [k][col] with row strides that spread the fragment reads over all banks.
    static constexpr int LDQ = MMA ? BK + 4 : BM;
    static constexpr int LDE = MMA ? BN + 4 : BN;
    static constexpr int Q_TILE = MMA ? BM * LDQ : BK * BM;    // doubles per buffer
    static constexpr int E_TILE = BK * LDE;
    static constexpr int MB = BM / 8;                          // 8-row MMA blocks of the tile
    static constexpr int WN = BN / (NT / 32);                  // columns owned by one warp
    static constexpr int NB = WN / 8;
    static_assert(!MMA || (BM % 8 == 0 && WN % 8 == 0 && BK % 4 == 0), "MMA tiling");
    static constexpr int E_PER_THREAD = (BK * BN) / NT;
    static constexpr int Q_PER_THREAD = (BK * BM + NT - 1) / NT;
    static_assert(BM % TM == 0 && BN % TN == 0 && TN % 2 == 0, "tile shape");
    static_assert((BK * BN) % NT == 0, "E chunk must divide evenly over the threads");
    static_assert(NT % BN == 0 && BK % (NT / BN) == 0, "a thread keeps one E column");
    // an upper bound that is monotone in G and A (a batch launch sizes shared memory for its largest member):
    // G * (1 + split) <= G + max(NT, G)
    static size_t smem_bytes(int G, int A)
    {
        const size_t a_pad = (size_t)((A + BK - 1) / BK) * BK;
        return sizeof(double) * ((size_t)2 * Q_TILE + (size_t)2 * E_TILE + (size_t)BM * HS_LD + (size_t)((G + (NT > G ? NT : G) + 2 * BK + 3) & ~3) + 4 * a_pad) +
               sizeof(int) * 4 * a_pad;
    }
};

// offsets sharing kernel node alpha: (A-1-alpha, A+alpha) for an odd extent, the single offset 0 for m == 2
__device__ __forceinline__ int folded_offset(int A, int n_fold, int alpha, int f)
{
    return n_fold == 1 ? 0 : (f == 0 ? A - 1 - alpha : A + alpha);
}

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// The work of one CTA: row `lx` of this rank, slab `slab` (= beta tile * chunks + column chunk).  A device function so
// that both the one-launch-per-step kernel and the persistent multi-step kernel run exactly the same code.
template <int BM, int BN, int TM, int TN, int BK, bool MMA, bool WS>
__device__ __forceinline__ void factored_body(const StepArgs& S, const int slab, const int lx, const bool build_taps = true,
                                              const int cached_w_rows = -1)
{
    const DevProblem& P = S.P;
    // no __restrict__ on the state: inside the persistent kernel C is rewritten every step by other CTAs, so these
    // loads must stay ordinary (coherent after grid.sync), never the read-only (.nc) path
    const double* C = S.C;
    double* partial = S.partial;
    double* wc_rows = S.wc_rows;
    const int tile = slab / S.chunks, chunk = slab - tile * S.chunks;
    PHASE_MARK(0);
#ifdef LDM_PHASE_CLOCKS
    if (threadIdx.x == 0) { g_phase_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 4)][7] = 0; g_phase_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 4)][8] = 0; }
    __syncthreads();
#endif
    using Cfg = FactoredConfig<BM, BN, TM, TN, BK, MMA>;
    constexpr int TYN = Cfg::TYN, TXN = Cfg::TXN, NT = Cfg::NT, HS_LD = Cfg::HS_LD;
    constexpr int LDQ = Cfg::LDQ, LDE = Cfg::LDE, Q_TILE = Cfg::Q_TILE, E_TILE = Cfg::E_TILE;
    constexpr int K_STRIDE = NT / BN;                     // E rows covered by one pass of the CTA's threads
    const int G = P.G, A = P.A, n_off = P.n_off;
    const int n_fold = n_off / A;
    const int x = P.rank + P.world * lx;
    const int b0 = tile * BM;                             // first beta row of this tile
    const int q0 = chunk * BN;                            // first column of this chunk
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int n_kc = (A + BK - 1) / BK;
    const int A_pad = n_kc * BK;

    extern __shared__ __align__(32) double smem[];
    double* Qs = smem;                                    // [2][Q_TILE]
    double* Es = Qs + 2 * Q_TILE;                         // [2][E_TILE]
    double* Hs = Es + 2 * E_TILE;                         // [BM][HS_LD]
    double* acc_y = Hs + BM * HS_LD;                      // [G]
    const int split = NT >= 2 * G ? NT / G : 1;           // threads sharing one y in the y-side pass
    double* part_y = acc_y + G;                           // [split][G]
    double* Eh = acc_y + G * (1 + split);                 // [2][BK] E values of the halo column q0-1
    double* tap_w = acc_y + ((G * (1 + split) + 2 * BK + 3) & ~3);   // [A_pad][4] (32-byte aligned): weights of the four C rows feeding E[alpha]
    int* tap_row = reinterpret_cast<int*>(tap_w + 4 * A_pad);   // [A_pad][4]: their offsets (slot * G)
    // Persistent kernel only (cached_w_rows >= 0): the y-side interpolation table is the same every step, and streaming it
    // from L2 in all CTAs at once is what bounds that phase -- so the 16-bit interval indices of the whole table and the
    // weights of its first cached_w_rows offsets are kept in the shared memory left over, filled on the first step.
    unsigned short* u_cache = reinterpret_cast<unsigned short*>(tap_row + 4 * A_pad);          // [n_off][G]
    double* w_cache = reinterpret_cast<double*>(u_cache + (((size_t)n_off * G + 3) & ~(size_t)3));   // [cached_w_rows][G]
    if (cached_w_rows >= 0 && build_taps) {
        for (int i = tid; i < n_off * G; i += NT) u_cache[i] = (unsigned short)P.upper_t[i];
        for (int i = tid; i < cached_w_rows * G; i += NT) w_cache[i] = P.weight_t[i];
    }

    // x-side interpolation taps of this row, once per CTA: E[alpha][q] = sum_t tap_w[alpha][t] * C[tap_row[alpha][t] + q]
    // (time-invariant per row: the persistent kernel builds them once and keeps them in shared memory across steps)
    for (int t = tid; build_taps && t < A_pad * 2; t += NT) {
        const int alpha = t >> 1, f = t & 1;
        double w_lo = 0.0, w_hi = 0.0;
        int r_lo = 0, r_hi = 0;
        if (alpha < A && f < n_fold) {
            const int o = folded_offset(A, n_fold, alpha, f);
            const int u = P.upper[(size_t)x * n_off + o];
            if (u != 0) {
                const double w = P.weight[(size_t)x * n_off + o];
                w_lo = 1.0 - w; w_hi = w;
                r_lo = row_slot(P, u - 1) * G; r_hi = row_slot(P, u) * G;
            }
        }
        tap_w[alpha * 4 + f * 2] = w_lo; tap_w[alpha * 4 + f * 2 + 1] = w_hi;
        tap_row[alpha * 4 + f * 2] = r_lo; tap_row[alpha * 4 + f * 2 + 1] = r_hi;
    }
    for (int y = tid; y < G; y += NT) acc_y[y] = 0.0;
    __syncthreads();
    PHASE_MARK(1);

    double c_raw[Cfg::E_PER_THREAD][4];                   // the four C values behind each E element of the next chunk
    double q_reg[Cfg::Q_PER_THREAD];
    double h_raw[4] = {0.0, 0.0, 0.0, 0.0};               // threads < BK: the same for the halo column q0-1
    double h_acc = 0.0;                                   // threads < BM: H[b0 + tid][q0-1]
    const bool want_halo = q0 > 0;
    const int e_col = tid % BN, e_k0 = tid / BN;          // this thread's E column and first row within a chunk

    {
        const int q = q0 + e_col;
        const double* c_col = C + (q < G ? q : G - 1);    // clamped: columns beyond G are multiplied by 0 below
        const double live = q < G ? 1.0 : 0.0;

        // Next chunk of E and Q in two halves: `issue_chunk` only issues the global loads (raw C values, Q entries) so that
        // they fly under the contraction of the current chunk; `finish_chunk` does the interpolation arithmetic and the
        // shared-memory stores afterwards.  (Doing the arithmetic up front would park every warp on L2 latency.)
        auto issue_chunk = [&](int kc) {
            #pragma unroll
            for (int r = 0; r < Cfg::E_PER_THREAD; r++) {
                const int alpha = kc * BK + r * K_STRIDE + e_k0;           // < A_pad; rows beyond A carry zero weights
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + alpha * 4);
                c_raw[r][0] = c_col[ra.x]; c_raw[r][1] = c_col[ra.y]; c_raw[r][2] = c_col[ra.z]; c_raw[r][3] = c_col[ra.w];
            }
            if (want_halo && tid < BK) {
                const int alpha = kc * BK + tid;
                const double* c_h = C + (q0 - 1);
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + alpha * 4);
                h_raw[0] = c_h[ra.x]; h_raw[1] = c_h[ra.y]; h_raw[2] = c_h[ra.z]; h_raw[3] = c_h[ra.w];
            }
            #pragma unroll
            for (int r = 0; r < Cfg::Q_PER_THREAD; r++) {
                const int e = r * NT + tid;
                // Q is symmetric, so it can be read along whichever index the shared-memory layout wants contiguous:
                // consecutive threads take consecutive beta ([k][BM] layout) or consecutive alpha ([beta][k] layout)
                const int k = MMA ? e % BK : e / BM, bm = MMA ? e / BK : e % BM;
                const int beta = b0 + bm, alpha = kc * BK + k;
                q_reg[r] = (e < BK * BM && beta < A && alpha < A)
                               ? (MMA ? P.quadrant[(size_t)beta * A + alpha] : P.quadrant[(size_t)alpha * A + beta]) : 0.0;
            }
        };
        auto finish_chunk = [&](int kc, int buf) {
            #pragma unroll
            for (int r = 0; r < Cfg::E_PER_THREAD; r++) {
                const int alpha = kc * BK + r * K_STRIDE + e_k0;
                const double4 wa = *reinterpret_cast<const double4*>(tap_w + alpha * 4);
                const double v = (c_raw[r][0] * wa.x + c_raw[r][1] * wa.y) + (c_raw[r][2] * wa.z + c_raw[r][3] * wa.w);
                Es[buf * E_TILE + (r * K_STRIDE + e_k0) * LDE + e_col] = v * live;
            }
            if (want_halo && tid < BK) {
                const double4 wa = *reinterpret_cast<const double4*>(tap_w + (kc * BK + tid) * 4);
                Eh[buf * BK + tid] = (h_raw[0] * wa.x + h_raw[1] * wa.y) + (h_raw[2] * wa.z + h_raw[3] * wa.w);
            }
            #pragma unroll
            for (int r = 0; r < Cfg::Q_PER_THREAD; r++) {
                const int e = r * NT + tid;
                if (e < BK * BM) Qs[buf * Q_TILE + (MMA ? (e / BK) * LDQ + e % BK : e)] = q_reg[r];
            }
        };

        // accumulators: a TM x TN register tile per thread, or MB x NB tensor-core fragments (2 doubles each) per warp
        constexpr int ACC_A = MMA ? Cfg::MB : TM, ACC_B = MMA ? Cfg::NB * 2 : TN;
        double acc[ACC_A][ACC_B];
        #pragma unroll
        for (int i = 0; i < ACC_A; i++)
            #pragma unroll
            for (int j = 0; j < ACC_B; j++) acc[i][j] = 0.0;
        const int lane_g = (tid & 31) >> 2, lane_t = tid & 3;       // MMA fragment coordinates of this lane
        const int wn0 = (tid >> 5) * Cfg::WN;                        // first column of this warp's strip

        issue_chunk(0);
        finish_chunk(0, 0);
        __syncthreads();
        PHASE_MARK(2);
        for (int kc = 0; kc < n_kc; kc++) {
            const int buf = kc & 1;
            if (kc + 1 < n_kc) issue_chunk(kc + 1);              // global loads in flight under the contraction
            const double* qs = Qs + buf * Q_TILE;
            const double* es = Es + buf * E_TILE;
            if constexpr (MMA) {
                // FP64 tensor cores: D(8x8) += A(8x4) B(4x8); lane (g, t) holds A[g][t], B[t][g], D[g][2t..2t+1]
                #pragma unroll
                for (int k4 = 0; k4 < BK; k4 += 4) {
                    double af[Cfg::MB], bf[Cfg::NB];
                    #pragma unroll
                    for (int mb = 0; mb < Cfg::MB; mb++) af[mb] = qs[(mb * 8 + lane_g) * LDQ + k4 + lane_t];
                    #pragma unroll
                    for (int nb = 0; nb < Cfg::NB; nb++) bf[nb] = es[(k4 + lane_t) * LDE + wn0 + nb * 8 + lane_g];
                    #pragma unroll
                    for (int mb = 0; mb < Cfg::MB; mb++)
                        #pragma unroll
                        for (int nb = 0; nb < Cfg::NB; nb++)
                            dmma_8x8x4(*reinterpret_cast<double(*)[2]>(&acc[mb][nb * 2]), af[mb], bf[nb]);
                }
            } else {
                #pragma unroll
                for (int k = 0; k < BK; k++) {
                    double a[TM], b[TN];
                    #pragma unroll
                    for (int i = 0; i < TM; i++) a[i] = qs[k * BM + i * TYN + ty];
                    #pragma unroll
                    for (int j = 0; j < TN; j += 2) {            // 16-byte loads, consecutive threads -> consecutive 16 B
                        const double2 v = *reinterpret_cast<const double2*>(es + k * BN + (j / 2) * (2 * TXN) + tx * 2);
                        b[j] = v.x; b[j + 1] = v.y;
                    }
                    #pragma unroll
                    for (int i = 0; i < TM; i++)
                        #pragma unroll
                        for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                }
            }
            if (want_halo && tid < BM) {
                #pragma unroll
                for (int k = 0; k < BK; k++) h_acc = fma(qs[MMA ? tid * LDQ + k : k * BM + tid], Eh[buf * BK + k], h_acc);
            }
            if (kc + 1 < n_kc) finish_chunk(kc + 1, buf ^ 1);
            __syncthreads();
        }

        PHASE_MARK(3);
        // stage the H tile: Hs[row][1 + col]; column 0 is the halo column q0-1
        if (tid < BM) Hs[tid * HS_LD] = h_acc;
        if constexpr (MMA) {
            #pragma unroll
            for (int mb = 0; mb < Cfg::MB; mb++)
                #pragma unroll
                for (int nb = 0; nb < Cfg::NB; nb++) {
                    double* dst = Hs + (mb * 8 + lane_g) * HS_LD + 1 + wn0 + nb * 8 + lane_t * 2;
                    dst[0] = acc[mb][nb * 2];
                    dst[1] = acc[mb][nb * 2 + 1];
                }
        } else {
            #pragma unroll
            for (int i = 0; i < TM; i++)
                #pragma unroll
                for (int j = 0; j < TN; j++)
                    Hs[(i * TYN + ty) * HS_LD + 1 + (j / 2) * (2 * TXN) + tx * 2 + (j & 1)] = acc[i][j];
        }
        __syncthreads();

    }
    {
        PHASE_MARK(4);
        // y side: every tap pair (u-1, u) whose upper node u falls in this chunk.  Consecutive threads own consecutive y
        // (the transposed tables upper_t / weight_t make those reads contiguous); `split` threads share one y, each
        // taking every split-th beta row, so every thread runs a long loop of independent loads.
        // y side: every tap pair (u-1, u) whose upper node u falls in this chunk.  Consecutive threads own consecutive y
        // (the transposed tables make those reads contiguous); `split` threads share one y, each taking every split-th
        // beta row, so every thread runs a long branch-free loop of independent table reads.  Measured (diagnostic builds):
        // the table reads alone take 12.3 k cycles per CTA at G = 128 (233 KB at ~19 B/clk/SM, 8 warps of loads in
        // flight), the shared-memory gathers alone 5.5 k, together 12.4 k -- the phase is bound by global-load
        // throughput at this occupancy.  Tried without gain: bigger load batches, staggering the rows per CTA, rebuilding
        // the weights from node tables in shared memory (slower: scattered 8-byte gathers).
        // A sample beyond the grid is tabulated as (u = 0, w = 0): in chunk 0 it reads 1 * halo + 0 * H[0], and the halo
        // column of chunk 0 is zero, so no validity test is needed; pairs of other chunks drop out by select.
        const int rows_here = min(BM, A - b0);
        const double* hbase = Hs + 1 - q0;
        const bool whole_row = (S.chunks == 1);            // every tap pair lands in this chunk: no range test needed
        for (int idx = tid; idx < G * split; idx += NT) {
            const int part = idx / G, y = idx - part * G;
            double s = 0.0;
            for (int f = 0; f < n_fold; f++) {
                const int o0 = folded_offset(A, n_fold, b0, f);
                const int dir = (n_fold == 2 && f == 0) ? -1 : 1;
                const int stride = dir * G;
                if (cached_w_rows >= 0) {
                    // whole_row is implied (the persistent kernel runs one slab per row)
                    const unsigned short* uc = u_cache + (size_t)o0 * G + y;
                    #pragma unroll 8
                    for (int bm = part; bm < rows_here; bm += split) {
                        const int o = o0 + dir * bm;
                        const int u = uc[bm * stride];
                        const double w = o < cached_w_rows ? w_cache[(size_t)o * G + y] : P.weight_t[(size_t)o * G + y];
                        const double* hrow = hbase + bm * HS_LD;
                        s += hrow[u - 1] * (1.0 - w) + hrow[u] * w;
                    }
                    continue;
                }
                const int* up = P.upper_t + (size_t)o0 * G + y;
                const double* wp = P.weight_t + (size_t)o0 * G + y;
                if (whole_row) {
                    #pragma unroll 8
                    for (int bm = part; bm < rows_here; bm += split) {
                        const int u = up[bm * stride];
                        const double w = wp[bm * stride];
                        const double* hrow = hbase + bm * HS_LD;
                        s += hrow[u - 1] * (1.0 - w) + hrow[u] * w;
                    }
                } else {
                    #pragma unroll 8
                    for (int bm = part; bm < rows_here; bm += split) {
                        const int u = up[bm * stride];
                        const double w = wp[bm * stride];
                        const bool live = (unsigned)(u - q0) < (unsigned)BN;
                        const int uc = live ? u : q0;
                        const double* hrow = hbase + bm * HS_LD;
                        const double v = hrow[uc - 1] * (1.0 - w) + hrow[uc] * w;
                        s += live ? v : 0.0;
                    }
                }
            }
            part_y[part * G + y] = s;
        }
        __syncthreads();
        for (int y = tid; y < G; y += NT) {
            double s = acc_y[y];
            for (int part = 0; part < split; part++) s += part_y[part * G + y];     // fixed order
            acc_y[y] = s;
        }
        __syncthreads();
    }

    PHASE_MARK(5);
    double* out = partial + ((size_t)slab * local_rows(P) + lx) * G;
    for (int y = tid; y < G; y += NT) out[y] = acc_y[y];

    // WC = <wc_operator, C> (Solver.cpp:265-274) row by row, folded into this launch.  Every rank reduces the FULL
    // operator (so no all-reduce is needed): the first-slab CTAs of this rank share all G rows round-robin.
    if (slab == 0 && wc_rows != nullptr) {
        const int n_local = local_rows(P);
        int slot = 0;
        for (int xr = lx; xr < G; xr += n_local, slot++) {
            if (slot % (NT / 32) != tid / 32) continue;          // one warp per row
            const double s = wc_row_dot(P.wc_operator + (size_t)xr * G, C + (size_t)row_slot(P, xr) * G, G, tid % 32);
            if (tid % 32 == 0) wc_rows[xr] = s;
        }
    }
    PHASE_MARK(6);
}

template <int BM, int BN, int TM, int TN, int BK, bool BATCH, bool MMA, bool WS>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
factored_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    if (BATCH && ((int)blockIdx.x >= S.tiles * S.chunks || (int)blockIdx.y >= local_rows(S.P))) return;   // batch grids are padded
    factored_body<BM, BN, TM, TN, BK, MMA, WS>(S, blockIdx.x, blockIdx.y);
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent multi-step kernel (cooperative launch, one CTA per SM): ALL Euler steps of a solve inside one launch.
// Per step: CTAs 0..G-1 run the factored contraction of their row while the spare CTAs apply the sparse ring
// correction; grid-wide barrier; every CTA closes its share of the nodes (closure, Euler update, clamp, history);
// grid-wide barrier.  For grids that fit one slab per row (G <= 128, A <= 80, one rank) this removes the three kernel
// boundaries per step -- the small default grids of the reference are latency bound, not FLOP bound.
// ------------------------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, int BK>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
persistent_kernel(const PersistentArgs A)
{
    constexpr int NT = (BM / TM) * (BN / TN);
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double wc_shared;
    StepArgs S = A.step;
    const DevProblem& P = S.P;
    const int G = P.G, n = G * G, tid = threadIdx.x;
    int parity = A.parity;
    for (int k = 0; k < A.n_steps; k++) {
        S.C = A.C[parity]; S.C_next = A.C[parity ^ 1];
        S.N_cur = A.N + parity; S.N_next = A.N + (parity ^ 1);
        const long long point = A.first_point + k;                 // index of the state this step starts from
        if (blockIdx.x == 0 && tid == 0 && A.observables) {
            A.observables[2 * point] = S.N_cur[0];                                  // Values[k][0]       (Solver.cpp:1011)
            A.observables[2 * point + 1] = S.C[(size_t)0 * G + (G - 1)];            // Values[k][G]       (Solver.cpp:1014)
        }
        if ((int)blockIdx.x < G) {
            factored_body<BM, BN, TM, TN, BK, true, false>(S, 0, blockIdx.x, k == 0, A.cached_w_rows);
        } else {
            const int warps = (gridDim.x - G) * (NT / 32);
            for (int row = (blockIdx.x - G) * (NT / 32) + tid / 32; row < P.n_ring; row += warps) {
                double acc = 0.0;
                for (int en = P.ring_begin[row] + tid % 32; en < P.ring_begin[row + 1]; en += 32) acc += P.ring_weight[en] * S.C[P.ring_node[en]];
                acc = warp_sum(acc);
                if (tid % 32 == 0) S.ring[P.ring_point[row]] = acc;
            }
        }
        grid.sync();
        if (tid < 32) {
            const double v = wc_total(S.wc_rows, G, tid);
            if (tid == 0) wc_shared = v;
        }
        __syncthreads();
        const double N = S.N_cur[0];
        const double WC = __dmul_rn(add_outside(P, wc_shared, N, S.C, S.scalars, P.wc_outside, 0.0), P.hh);   // Solver.cpp:311, 314
        double* values_next = A.values ? A.values + (point + 1) * A.stride : nullptr;
        double* derivs_here = A.derivs ? A.derivs + point * A.stride : nullptr;
        if (blockIdx.x == 0 && tid == 0) {
            first_moment_update(P, N, WC, S.dt, S.scalars, S.N_next);
            if (values_next) values_next[0] = S.N_next[0];
            if (derivs_here) derivs_here[0] = S.scalars[1];
        }
        for (int q = blockIdx.x * NT + tid; q < n; q += gridDim.x * NT) {
            const double inside = S.partial[q] - S.ring[q];
            const double c = S.C[q];
            const double mc = __dmul_rn(add_outside(P, inside, N, S.C, S.scalars, P.out_weight[q], 0.0), P.hh);
            const double dc = second_moment_rhs(P, mc, c, N, WC, P.m_nodes[q], P.w_nodes[q]);
            double v = __dadd_rn(c, __dmul_rn(dc, S.dt));                          // Solver.cpp:577
            v = v < 0.0 ? 0.0 : v;                                                 // Solver.cpp:578-580
            S.C_next[q] = v;
            if (values_next) values_next[1 + q] = v;
            if (derivs_here) derivs_here[1 + q] = dc;
        }
        grid.sync();
        parity ^= 1;
    }
}

template <int BM, int BN, int TM, int TN, int BK, bool BATCH, bool MMA, bool WS>
void launch_variant(const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, int A, cudaStream_t stream)
{
    using Cfg = FactoredConfig<BM, BN, TM, TN, BK, MMA>;
    const size_t smem = Cfg::smem_bytes(G, A);
    static bool configured[64] = {};                 // the opt-in is per device
    int device = 0;
    cudaGetDevice(&device);
    if (!configured[device & 63]) {
        cudaFuncSetAttribute(factored_kernel<BM, BN, TM, TN, BK, BATCH, MMA, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[device & 63] = true;
    }
    static const StepArgs none{};
    factored_kernel<BM, BN, TM, TN, BK, BATCH, MMA, WS><<<grid, Cfg::NT, smem, stream>>>(one ? *one : none, d_batch);
}

template <bool BATCH>
void dispatch(const FactoredPlan& shape, const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, int A, cudaStream_t stream)
{
    // the contraction runs on the FP64 tensor cores (mma.sync m8n8k4); the DFMA register-tile instances of factored_body
    // (template parameter MMA = false) were measured slower and are not built
    if (shape.tn == 4) {          // twice the threads per tile (16 warps on an 80 x 128 tile)
        if (shape.bm == 128 && shape.bn == 128) return launch_variant<128, 128, 8, 4, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
        if (shape.bm == 80 && shape.bn == 128) return launch_variant<80, 128, 5, 4, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
        if (shape.bm == 80 && shape.bn == 64) return launch_variant<80, 64, 5, 4, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
        if (shape.bm == 40 && shape.bn == 128) return launch_variant<40, 128, 5, 4, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
        return launch_variant<40, 64, 5, 4, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
    }
    if (shape.bm == 80 && shape.bn == 128) return launch_variant<80, 128, 5, 8, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
    if (shape.bm == 80 && shape.bn == 64) return launch_variant<80, 64, 5, 8, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
    if (shape.bm == 40 && shape.bn == 128) return launch_variant<40, 128, 5, 8, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
    return launch_variant<40, 64, 5, 8, 16, BATCH, true, false>(one, d_batch, grid, G, A, stream);
}

}  // namespace

template <int BM, int BN, int TN, int TM = 5>
cudaError_t launch_persistent_variant(const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid)
{
    using Cfg = FactoredConfig<BM, BN, TM, TN, 16, true>;
    auto kernel = persistent_kernel<BM, BN, TM, TN, 16>;
    // shared memory: the step's working set, then the 16-bit y-side index table, then as many weight rows as still fit
    const int G = args.step.P.G, n_off = args.step.P.n_off;
    const size_t base = Cfg::smem_bytes(G, args.step.P.A);
    const size_t u_bytes = (((size_t)n_off * G + 3) & ~(size_t)3) * sizeof(unsigned short);
    const size_t budget = 226 * 1024;
    PersistentArgs with_cache = args;
    with_cache.cached_w_rows = -1;
    size_t smem = base;
    if (base + u_bytes <= budget && G <= 65535) {
        const int rows = (int)std::min<size_t>((size_t)n_off, (budget - base - u_bytes) / ((size_t)G * sizeof(double)));
        with_cache.cached_w_rows = rows;
        smem = base + u_bytes + (size_t)rows * G * sizeof(double);
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e != cudaSuccess) return e;
    int device = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::NT, smem);
    if (e != cudaSuccess) return e;
    if (max_grid) *max_grid = sms * per_sm;
    if (grid <= 0) return cudaSuccess;                  // query only
    if (grid > sms * per_sm) return cudaErrorCooperativeLaunchTooLarge;
    void* params[] = {&with_cache};
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(Cfg::NT), params, smem, stream);
}

cudaError_t launch_persistent(const FactoredPlan& plan, const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid)
{
    if (plan.bm == 128) return launch_persistent_variant<128, 128, 4, 8>(args, grid, stream, max_grid);
    if (plan.tn == 4) {
        if (plan.bm == 80 && plan.bn == 128) return launch_persistent_variant<80, 128, 4>(args, grid, stream, max_grid);
        if (plan.bm == 80 && plan.bn == 64) return launch_persistent_variant<80, 64, 4>(args, grid, stream, max_grid);
        if (plan.bm == 40 && plan.bn == 128) return launch_persistent_variant<40, 128, 4>(args, grid, stream, max_grid);
        return launch_persistent_variant<40, 64, 4>(args, grid, stream, max_grid);
    }
    if (plan.bm == 80 && plan.bn == 128) return launch_persistent_variant<80, 128, 8>(args, grid, stream, max_grid);
    if (plan.bm == 80 && plan.bn == 64) return launch_persistent_variant<80, 64, 8>(args, grid, stream, max_grid);
    if (plan.bm == 40 && plan.bn == 128) return launch_persistent_variant<40, 128, 8>(args, grid, stream, max_grid);
    return launch_persistent_variant<40, 64, 8>(args, grid, stream, max_grid);
}

FactoredPlan plan_factored(const DevProblem& P)
{
    FactoredPlan plan;
    // one column chunk when the grid allows it, because every
    // chunk's CTA scans the whole y-side table
    const int A = P.A;
    // Tile height: every beta tile regenerates E and re-scans the y-side table, a fixed cost of roughly 40 rows' worth of
    // contraction, so taller tiles win unless they pad too much (measured at G = 512, A = 243: 4 x 80 rows 2.9 ms,
    // 7 x 40 rows 4.4 ms).  128-row tiles need the 16-warp tensor-core variant and 128-column chunks.
    plan.bm = A <= 40 ? 40 : 80;
    if (P.G > 64 && A > 80) {
        const int cost80 = ((A + 79) / 80) * (80 + 40), cost128 = ((A + 127) / 128) * (128 + 40);
        if (cost128 < cost80) plan.bm = 128;
    }
    plan.bn = P.G <= 64 ? 64 : 128;
    plan.tn = 4;       // 16 warps on an 80 x 128 tile: twice the table reads in flight in the y-side phase (12.4 k -> 8.9 k cycles)
    // shape overrides of the tests (ldm_tuning_set)
    const Tuning& t = tuning();
    if (t.factored_bn == 64 || t.factored_bn == 128) plan.bn = t.factored_bn;
    if (t.factored_tn == 4 || t.factored_tn == 8) plan.tn = t.factored_tn;
    if (t.factored_bm == 40 || t.factored_bm == 80 || (t.factored_bm == 128 && P.G > 64)) plan.bm = t.factored_bm;
    if (plan.bm == 128) { plan.tn = 4; plan.bn = 128; }   // the only 128-row kernel
    plan.tiles = (A + plan.bm - 1) / plan.bm;
    plan.chunks = (P.G + plan.bn - 1) / plan.bn;
    return plan;
}

#ifdef LDM_PHASE_CLOCKS
extern "C" LDM_API int ldm_debug_phase_clocks(long long* out, int rows)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_phase_clock, sizeof(long long) * 12 * (size_t)rows);
}
#endif

void launch_factored(const FactoredPlan& plan, const StepArgs& one, cudaStream_t stream)
{
    dispatch<false>(plan, &one, nullptr, dim3(plan.slabs(), local_rows(one.P), 1), one.P.G, one.P.A, stream);
}

void launch_factored_batch(const FactoredPlan& shape, const StepArgs* d_batch, int n, int max_slabs, int max_rows, int max_G, int max_A,
                           cudaStream_t stream)
{
    dispatch<true>(shape, nullptr, d_batch, dim3(max_slabs, max_rows, n), max_G, max_A, stream);
}

}  // namespace ldm
