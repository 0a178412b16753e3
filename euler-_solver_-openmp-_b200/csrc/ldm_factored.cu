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
#include <cstdlib>

#include "ldm_device.h"

namespace ldm {

namespace {

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);   // fixed butterfly
    return v;
}

template <int BM, int BN, int TM, int TN, int BK>
struct FactoredConfig {
    static constexpr int TYN = BM / TM;            // thread rows
    static constexpr int TXN = BN / TN;            // thread columns
    static constexpr int NT = TYN * TXN;           // threads per CTA
    static constexpr int HS_LD = BN + 2;           // H staging row: 1 halo column + BN + 1 pad
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
        return sizeof(double) * ((size_t)2 * BK * BM + (size_t)2 * BK * BN + (size_t)BM * HS_LD + (size_t)((G + (NT > G ? NT : G) + 2 * BK + 3) & ~3) + 4 * a_pad) +
               sizeof(int) * 4 * a_pad;
    }
};

// offsets sharing kernel node alpha: (A-1-alpha, A+alpha) for an odd extent, the single offset 0 for m == 2
__device__ __forceinline__ int folded_offset(int A, int n_fold, int alpha, int f)
{
    return n_fold == 1 ? 0 : (f == 0 ? A - 1 - alpha : A + alpha);
}

template <int BM, int BN, int TM, int TN, int BK, bool BATCH>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
factored_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    const DevProblem& P = S.P;
    if (BATCH && ((int)blockIdx.x >= S.tiles * S.chunks || (int)blockIdx.y >= local_rows(P))) return;   // batch grids are padded
    const double* __restrict__ C = S.C;
    double* __restrict__ partial = S.partial;
    double* __restrict__ wc_rows = S.wc_rows;
    const int tile = blockIdx.x / S.chunks, chunk = blockIdx.x - tile * S.chunks;
    using Cfg = FactoredConfig<BM, BN, TM, TN, BK>;
    constexpr int TYN = Cfg::TYN, TXN = Cfg::TXN, NT = Cfg::NT, HS_LD = Cfg::HS_LD;
    constexpr int K_STRIDE = NT / BN;                     // E rows covered by one pass of the CTA's threads
    const int G = P.G, A = P.A, n_off = P.n_off;
    const int n_fold = n_off / A;
    const int lx = blockIdx.y;
    const int x = P.rank + P.world * lx;
    const int b0 = tile * BM;                             // first beta row of this tile
    const int q0 = chunk * BN;                            // first column of this chunk
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int n_kc = (A + BK - 1) / BK;
    const int A_pad = n_kc * BK;

    extern __shared__ __align__(32) double smem[];
    double* Qs = smem;                                    // [2][BK][BM]
    double* Es = Qs + 2 * BK * BM;                        // [2][BK][BN]
    double* Hs = Es + 2 * BK * BN;                        // [BM][HS_LD]
    double* acc_y = Hs + BM * HS_LD;                      // [G]
    const int split = NT >= 2 * G ? NT / G : 1;           // threads sharing one y in the y-side pass
    double* part_y = acc_y + G;                           // [split][G]
    double* Eh = acc_y + G * (1 + split);                 // [2][BK] E values of the halo column q0-1
    double* tap_w = acc_y + ((G * (1 + split) + 2 * BK + 3) & ~3);   // [A_pad][4] (32-byte aligned): weights of the four C rows feeding E[alpha]
    int* tap_row = reinterpret_cast<int*>(tap_w + 4 * A_pad);   // [A_pad][4]: their offsets (slot * G)

    // x-side interpolation taps of this row, once per CTA: E[alpha][q] = sum_t tap_w[alpha][t] * C[tap_row[alpha][t] + q]
    for (int t = tid; t < A_pad * 2; t += NT) {
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

    double e_reg[Cfg::E_PER_THREAD];
    double q_reg[Cfg::Q_PER_THREAD];
    double eh_reg = 0.0;                                  // threads < BK: E[alpha][q0-1]
    double h_acc = 0.0;                                   // threads < BM: H[b0 + tid][q0-1]
    const bool want_halo = q0 > 0;
    const int e_col = tid % BN, e_k0 = tid / BN;          // this thread's E column and first row within a chunk

    {
        const int q = q0 + e_col;
        const double* c_col = C + (q < G ? q : G - 1);    // clamped: columns beyond G are multiplied by 0 below
        const double live = q < G ? 1.0 : 0.0;

        auto load_chunk = [&](int kc) {
            #pragma unroll
            for (int r = 0; r < Cfg::E_PER_THREAD; r++) {
                const int alpha = kc * BK + r * K_STRIDE + e_k0;           // < A_pad; rows beyond A carry zero weights
                const double4 wa = *reinterpret_cast<const double4*>(tap_w + alpha * 4);
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + alpha * 4);
                e_reg[r] = (c_col[ra.x] * wa.x + c_col[ra.y] * wa.y) + (c_col[ra.z] * wa.z + c_col[ra.w] * wa.w);
            }
            if (want_halo && tid < BK) {
                const int alpha = kc * BK + tid;
                const double* c_h = C + (q0 - 1);
                const double4 wa = *reinterpret_cast<const double4*>(tap_w + alpha * 4);
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + alpha * 4);
                eh_reg = (c_h[ra.x] * wa.x + c_h[ra.y] * wa.y) + (c_h[ra.z] * wa.z + c_h[ra.w] * wa.w);
            }
            #pragma unroll
            for (int r = 0; r < Cfg::Q_PER_THREAD; r++) {
                const int e = r * NT + tid;
                const int k = e / BM, bm = e - k * BM;     // Q is symmetric: read it alpha-major so consecutive threads
                const int beta = b0 + bm, alpha = kc * BK + k;   // take consecutive beta (coalesced, conflict-free store)
                q_reg[r] = (e < BK * BM && beta < A && alpha < A) ? P.quadrant[(size_t)alpha * A + beta] : 0.0;
            }
        };
        auto store_chunk = [&](int buf) {
            #pragma unroll
            for (int r = 0; r < Cfg::E_PER_THREAD; r++)
                Es[buf * BK * BN + (r * K_STRIDE + e_k0) * BN + e_col] = e_reg[r] * live;
            if (want_halo && tid < BK) Eh[buf * BK + tid] = eh_reg;
            #pragma unroll
            for (int r = 0; r < Cfg::Q_PER_THREAD; r++) {
                const int e = r * NT + tid;
                if (e < BK * BM) Qs[buf * BK * BM + e] = q_reg[r];
            }
        };

        double acc[TM][TN];
        #pragma unroll
        for (int i = 0; i < TM; i++)
            #pragma unroll
            for (int j = 0; j < TN; j++) acc[i][j] = 0.0;

        load_chunk(0);
        store_chunk(0);
        __syncthreads();
        for (int kc = 0; kc < n_kc; kc++) {
            const int buf = kc & 1;
            if (kc + 1 < n_kc) load_chunk(kc + 1);               // global loads in flight under the FMAs
            const double* qs = Qs + buf * BK * BM;
            const double* es = Es + buf * BK * BN;
            #pragma unroll
            for (int k = 0; k < BK; k++) {
                double a[TM], b[TN];
                #pragma unroll
                for (int i = 0; i < TM; i++) a[i] = qs[k * BM + i * TYN + ty];
                #pragma unroll
                for (int j = 0; j < TN; j += 2) {                // 16-byte loads, consecutive threads -> consecutive 16 B
                    const double2 v = *reinterpret_cast<const double2*>(es + k * BN + (j / 2) * (2 * TXN) + tx * 2);
                    b[j] = v.x; b[j + 1] = v.y;
                }
                #pragma unroll
                for (int i = 0; i < TM; i++)
                    #pragma unroll
                    for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
            if (want_halo && tid < BM) {
                #pragma unroll
                for (int k = 0; k < BK; k++) h_acc = fma(qs[k * BM + tid], Eh[buf * BK + k], h_acc);
            }
            if (kc + 1 < n_kc) store_chunk(buf ^ 1);
            __syncthreads();
        }

        // stage the H tile: Hs[row][1 + col]; column 0 is the halo column q0-1
        if (tid < BM) Hs[tid * HS_LD] = h_acc;
        #pragma unroll
        for (int i = 0; i < TM; i++)
            #pragma unroll
            for (int j = 0; j < TN; j++)
                Hs[(i * TYN + ty) * HS_LD + 1 + (j / 2) * (2 * TXN) + tx * 2 + (j & 1)] = acc[i][j];
        __syncthreads();

        // y side: every tap pair (u-1, u) whose upper node u falls in this chunk.  Consecutive threads own consecutive y
        // (the transposed tables upper_t / weight_t make those reads contiguous); `split` threads share one y, each
        // taking every split-th beta row, so every thread runs a long loop of independent loads.
        const int rows_here = min(BM, A - b0);
        const double* hbase = Hs + 1 - q0;
        for (int idx = tid; idx < G * split; idx += NT) {
            const int part = idx / G, y = idx - part * G;
            double s = 0.0;
            for (int f = 0; f < n_fold; f++) {
                const int o0 = folded_offset(A, n_fold, b0, f);
                const int dir = (n_fold == 2 && f == 0) ? -1 : 1;
                // batches of kBatch table reads issued together (they come from L2), then the shared-memory gathers;
                // branch-free: a tap pair outside this chunk is redirected to column q0 with weight 0
                constexpr int kBatch = 8;
                for (int bm0 = part; bm0 < rows_here; bm0 += split * kBatch) {
                    int u[kBatch];
                    double w[kBatch];
                    #pragma unroll
                    for (int r = 0; r < kBatch; r++) {
                        const int bm = min(bm0 + r * split, rows_here - 1);
                        const size_t c = (size_t)(o0 + dir * bm) * G + y;
                        u[r] = P.upper_t[c];
                        w[r] = P.weight_t[c];
                    }
                    #pragma unroll
                    for (int r = 0; r < kBatch; r++) {
                        const int bm = bm0 + r * split;
                        const bool live = bm < rows_here && u[r] >= q0 && u[r] < q0 + BN && u[r] != 0;
                        const int uc = live ? u[r] : q0;
                        const double* hrow = hbase + min(bm, rows_here - 1) * HS_LD;
                        const double v = hrow[uc - 1] * (1.0 - w[r]) + hrow[uc] * w[r];
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

    double* out = partial + ((size_t)blockIdx.x * local_rows(P) + lx) * G;
    for (int y = tid; y < G; y += NT) out[y] = acc_y[y];

    // WC = <wc_operator, C> (Solver.cpp:265-274) row by row, folded into this launch: the first beta-tile of local
    // row lx covers rows lx*world .. lx*world+world-1 (every rank reduces the full operator, so no all-reduce)
    if (blockIdx.x == 0 && wc_rows != nullptr) {
        for (int r = tid / 32; r < P.world; r += NT / 32) {
            const int xr = lx * P.world + r;
            if (xr >= G) continue;
            const double* op = P.wc_operator + (size_t)xr * G;
            const double* crow = C + (size_t)row_slot(P, xr) * G;
            double s = 0.0;
            for (int y = tid % 32; y < G; y += 32) s += op[y] * crow[y];
            s = warp_sum(s);
            if (tid % 32 == 0) wc_rows[xr] = s;
        }
    }
}

template <int BM, int BN, int TM, int TN, int BK, bool BATCH>
void launch_variant(const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, int A, cudaStream_t stream)
{
    using Cfg = FactoredConfig<BM, BN, TM, TN, BK>;
    const size_t smem = Cfg::smem_bytes(G, A);
    static bool configured[64] = {};                 // the opt-in is per device
    int device = 0;
    cudaGetDevice(&device);
    if (!configured[device & 63]) {
        cudaFuncSetAttribute(factored_kernel<BM, BN, TM, TN, BK, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[device & 63] = true;
    }
    static const StepArgs none{};
    factored_kernel<BM, BN, TM, TN, BK, BATCH><<<grid, Cfg::NT, smem, stream>>>(one ? *one : none, d_batch);
}

template <bool BATCH>
void dispatch(const FactoredPlan& shape, const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, int A, cudaStream_t stream)
{
    if (shape.bm == 80) {
        if (shape.bn == 64) launch_variant<80, 64, 5, 8, 16, BATCH>(one, d_batch, grid, G, A, stream);
        else launch_variant<80, 128, 5, 8, 16, BATCH>(one, d_batch, grid, G, A, stream);
    } else {
        if (shape.bn == 64) launch_variant<40, 64, 5, 8, 16, BATCH>(one, d_batch, grid, G, A, stream);
        else launch_variant<40, 128, 5, 8, 16, BATCH>(one, d_batch, grid, G, A, stream);
    }
}

}  // namespace

FactoredPlan plan_factored(const DevProblem& P)
{
    FactoredPlan plan;
    // beta-tile height that wastes the fewest padded rows; one column chunk when the grid allows it, because every
    // chunk's CTA scans the whole y-side table (LDM_FACTORED_BM / LDM_FACTORED_BN override, for tuning)
    const int A = P.A;
    const int pad80 = ((A + 79) / 80) * 80, pad40 = ((A + 39) / 40) * 40;
    plan.bm = pad40 < pad80 ? 40 : 80;
    plan.bn = P.G <= 64 ? 64 : 128;
    if (const char* env = std::getenv("LDM_FACTORED_BN")) {
        const int v = std::atoi(env);
        if (v == 64 || v == 128) plan.bn = v;
    }
    if (const char* env = std::getenv("LDM_FACTORED_BM")) {
        const int v = std::atoi(env);
        if (v == 40 || v == 80) plan.bm = v;
    }
    plan.tiles = (A + plan.bm - 1) / plan.bm;
    plan.chunks = (P.G + plan.bn - 1) / plan.bn;
    return plan;
}

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
