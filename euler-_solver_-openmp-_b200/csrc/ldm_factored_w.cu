// Factored m*C integral, second formulation: the x-side interpolation is folded into the kernel table once.
//
// ldm_factored.cu computes H_x = Q . E_x with E_x = L_x . C regenerated from C every step (4 gathered C values and 7
// FP64 operations per E element, again for every beta tile).  Q and the interpolation rows L_x are both time-invariant,
// so their product is precomputed at set-up:
//
//     W_x = Q . L_x            (A x K_x, K_x = number of moment-grid rows p that row x's samples touch)
//     H_x = W_x . C[p_lo(x) .. p_lo(x)+K_x) [:, :]
//
// Per step the contraction is then a plain dense FP64 GEMM whose two operands are copied straight from global memory
// (W_x rows, C rows) into shared memory with cp.async -- no gathers, no interpolation arithmetic, no L2 re-reads of C
// per beta tile -- and run on the FP64 tensor cores (mma.sync m8n8k4).  Rows far from the origin touch few grid rows
// (K_x small), rows near it up to ~0.8 G, so the FLOP count is about that of the E-form.  The y-side contraction and
// everything after it are the same as in ldm_factored.cu.
//
// Storage of W: for every x, a_pad rows (a multiple of the tile height) of K_x doubles (K_x a multiple of 16, zero
// padded), 6.6 MB for the 16 384-point grid and ~0.5 GB for the 262 144-point grid.
#include <algorithm>
#include <cstdlib>

#include "ldm_device.h"
#include "ldm_pointwise.cuh"

namespace ldm {

#ifdef LDM_PHASE_CLOCKS
extern __device__ long long g_phase_clock[148 * 4][12];
#define PHASE_MARK(i) do { if (threadIdx.x == 0) g_phase_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 4)][i] = clock64(); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

namespace {

constexpr int kBK = 16;          // K chunk
constexpr int kStages = 3;       // cp.async pipeline depth

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ int folded_offset(int A, int n_fold, int alpha, int f)
{
    return n_fold == 1 ? 0 : (f == 0 ? A - 1 - alpha : A + alpha);
}

template <int BM, int BN>
struct WConfig {
    static constexpr int NT = (BM / 40) * (BN / 64) * 64;      // 256 threads for 80 x 128
    static constexpr int WARPS = NT / 32;
    static constexpr int LDQ = kBK + 4;                        // W tile [beta][k]: fragment reads hit all banks
    static constexpr int LDE = BN + 8;                         // C tile [k][col]
    static constexpr int HS_LD = BN + 2;                       // staged H tile: halo column + BN + pad
    static constexpr int Q_TILE = BM * LDQ, E_TILE = kBK * LDE;
    static constexpr int MB = BM / 8, WN = BN / WARPS, NB = WN / 8;
    static_assert(WN % 8 == 0, "warp strip");
    static size_t smem_bytes(int G)
    {
        return sizeof(double) * ((size_t)kStages * (Q_TILE + E_TILE + kBK) + (size_t)BM * HS_LD + (size_t)((G + (NT > G ? NT : G) + 3) & ~3));
    }
};

template <int BM, int BN, bool BATCH>
__global__ void __launch_bounds__(WConfig<BM, BN>::NT)
factored_w_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    using Cfg = WConfig<BM, BN>;
    constexpr int NT = Cfg::NT, LDQ = Cfg::LDQ, LDE = Cfg::LDE, HS_LD = Cfg::HS_LD, Q_TILE = Cfg::Q_TILE, E_TILE = Cfg::E_TILE;
    constexpr int MB = Cfg::MB, NB = Cfg::NB, WN = Cfg::WN;
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    const DevProblem& P = S.P;
    if (BATCH && ((int)blockIdx.x >= S.tiles * S.chunks || (int)blockIdx.y >= local_rows(P))) return;   // batch grids are padded
    const double* __restrict__ C = S.C;
    const int tile = blockIdx.x / S.chunks, chunk = blockIdx.x - tile * S.chunks;
    PHASE_MARK(0);
    const int G = P.G, A = P.A, n_off = P.n_off;
    const int n_fold = n_off / A;
    const int lx = blockIdx.y;
    const int x = P.rank + P.world * lx;
    const int b0 = tile * BM, q0 = chunk * BN;
    const int tid = threadIdx.x;
    const int Kx = P.w_k[x], p_lo = P.w_plo[x];
    const int n_kc = Kx / kBK;
    const double* __restrict__ Wx = P.W + P.w_offset[x] + (size_t)b0 * Kx;

    extern __shared__ __align__(32) double smem[];
    double* Qs = smem;                                    // [stages][BM][LDQ]
    double* Es = Qs + kStages * Q_TILE;                   // [stages][BK][LDE]
    double* Eh = Es + kStages * E_TILE;                   // [stages][BK] column q0-1 of the C rows (halo)
    double* Hs = Eh + kStages * kBK;                      // [BM][HS_LD]
    double* acc_y = Hs + BM * HS_LD;                      // [G]
    const int split = NT >= 2 * G ? NT / G : 1;
    double* part_y = acc_y + G;                           // [split][G]

    const int cols = min(BN, G - q0);                     // live columns of this chunk
    const bool want_halo = q0 > 0;
    for (int i = tid; i < kStages * E_TILE; i += NT) Es[i] = 0.0;      // columns beyond the grid stay zero for good
    for (int y = tid; y < G; y += NT) acc_y[y] = 0.0;
    __syncthreads();

    // One K chunk = 16 columns of W_x (BM rows) and the 16 matching rows of C (this chunk's columns): plain copies.
    const bool c_aligned = ((G & 1) == 0);                // even row length: C rows start 16-byte aligned
    auto load_stage = [&](int kc, int st) {
        double* qs = Qs + st * Q_TILE;
        for (int i = tid; i < BM * (kBK / 2); i += NT) {
            const int row = i / (kBK / 2), piece = i - row * (kBK / 2);
            cp_async_16(qs + row * LDQ + piece * 2, Wx + (size_t)row * Kx + kc * kBK + piece * 2);
        }
        double* es = Es + st * E_TILE;
        if (c_aligned) {
            const int pieces = cols / 2;                  // 16-byte pieces per row (cols is even when G and BN are)
            for (int i = tid; i < kBK * pieces; i += NT) {
                const int k = i / pieces, piece = i - k * pieces;
                const int p = min(p_lo + kc * kBK + k, G - 1);              // rows past K_x meet zero W columns
                cp_async_16(es + k * LDE + piece * 2, C + (size_t)row_slot(P, p) * G + q0 + piece * 2);
            }
        } else {
            for (int i = tid; i < kBK * cols; i += NT) {
                const int k = i / cols, col = i - k * cols;
                const int p = min(p_lo + kc * kBK + k, G - 1);
                cp_async_8(es + k * LDE + col, C + (size_t)row_slot(P, p) * G + q0 + col);
            }
        }
        if (want_halo && tid < kBK) {
            const int p = min(p_lo + kc * kBK + tid, G - 1);
            cp_async_8(Eh + st * kBK + tid, C + (size_t)row_slot(P, p) * G + q0 - 1);
        }
    };

    double acc[MB][NB * 2];
    #pragma unroll
    for (int i = 0; i < MB; i++)
        #pragma unroll
        for (int j = 0; j < NB * 2; j++) acc[i][j] = 0.0;
    double h_acc = 0.0;                                   // threads < BM: H[b0 + tid][q0 - 1]
    const int lane_g = (tid & 31) >> 2, lane_t = tid & 3;
    const int wn0 = (tid >> 5) * WN;

    PHASE_MARK(1);
    load_stage(0, 0);
    cp_async_commit();
    if (n_kc > 1) load_stage(1, 1);
    cp_async_commit();
    for (int kc = 0; kc < n_kc; kc++) {
        cp_async_wait<1>();                               // chunk kc has landed (at most the next one is still in flight)
        __syncthreads();
        if (kc == 0) PHASE_MARK(2);                                  // ... for every thread, and everybody is done with chunk kc-1
        if (kc + 2 < n_kc) load_stage(kc + 2, (kc + 2) % kStages);
        cp_async_commit();
        const double* qs = Qs + (kc % kStages) * Q_TILE;
        const double* es = Es + (kc % kStages) * E_TILE;
        // FP64 tensor cores: D(8x8) += A(8x4) B(4x8); lane (g, t) holds A[g][t], B[t][g], D[g][2t..2t+1]
        #pragma unroll
        for (int k4 = 0; k4 < kBK; k4 += 4) {
            double af[MB], bf[NB];
            #pragma unroll
            for (int mb = 0; mb < MB; mb++) af[mb] = qs[(mb * 8 + lane_g) * LDQ + k4 + lane_t];
            #pragma unroll
            for (int nb = 0; nb < NB; nb++) bf[nb] = es[(k4 + lane_t) * LDE + wn0 + nb * 8 + lane_g];
            #pragma unroll
            for (int mb = 0; mb < MB; mb++)
                #pragma unroll
                for (int nb = 0; nb < NB; nb++)
                    dmma_8x8x4(*reinterpret_cast<double(*)[2]>(&acc[mb][nb * 2]), af[mb], bf[nb]);
        }
        if (want_halo && tid < BM) {
            const double* eh = Eh + (kc % kStages) * kBK;
            #pragma unroll
            for (int k = 0; k < kBK; k++) h_acc = fma(qs[tid * LDQ + k], eh[k], h_acc);
        }
    }
    cp_async_wait<0>();
    PHASE_MARK(3);

    // stage the H tile: Hs[row][1 + col]; column 0 is the halo column q0-1 (zero for the first chunk)
    if (tid < BM) Hs[tid * HS_LD] = h_acc;
    #pragma unroll
    for (int mb = 0; mb < MB; mb++)
        #pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            double* dst = Hs + (mb * 8 + lane_g) * HS_LD + 1 + wn0 + nb * 8 + lane_t * 2;
            dst[0] = acc[mb][nb * 2];
            dst[1] = acc[mb][nb * 2 + 1];
        }
    __syncthreads();

    PHASE_MARK(4);
    // y side (see ldm_factored.cu): every tap pair (u-1, u) whose upper node falls in this chunk; consecutive threads
    // own consecutive y, `split` threads share one y.  A sample beyond the grid is tabulated as (u = 0, w = 0): in
    // chunk 0 it reads 1 * halo + 0 * H[0] = 0; pairs of other chunks drop out by select.
    const int rows_here = min(BM, A - b0);
    const double* hbase = Hs + 1 - q0;
    const bool whole_row = (S.chunks == 1);
    for (int idx = tid; idx < G * split; idx += NT) {
        const int part = idx / G, y = idx - part * G;
        double s = 0.0;
        for (int f = 0; f < n_fold; f++) {
            const int o0 = folded_offset(A, n_fold, b0, f);
            const int dir = (n_fold == 2 && f == 0) ? -1 : 1;
            const int* up = P.upper_t + (size_t)o0 * G + y;
            const double* wp = P.weight_t + (size_t)o0 * G + y;
            const int stride = dir * G;
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
    PHASE_MARK(5);
    double* out = S.partial + ((size_t)blockIdx.x * local_rows(P) + lx) * G;
    for (int y = tid; y < G; y += NT) {
        double s = 0.0;
        for (int part = 0; part < split; part++) s += part_y[part * G + y];     // fixed order
        out[y] = s;
    }

    // WC = <wc_operator, C> row by row (Solver.cpp:265-274): the first-slab CTAs of this rank share all G rows
    if (blockIdx.x == 0 && S.wc_rows != nullptr) {
        const int n_local = local_rows(P);
        int slot = 0;
        for (int xr = lx; xr < G; xr += n_local, slot++) {
            if (slot % (NT / 32) != tid / 32) continue;          // one warp per row
            const double s = wc_row_dot(P.wc_operator + (size_t)xr * G, C + (size_t)row_slot(P, xr) * G, G, tid % 32);
            if (tid % 32 == 0) S.wc_rows[xr] = s;
        }
    }
    PHASE_MARK(6);
}

// W_x = Q . L_x, one thread per (x, beta) row: walks the kernel nodes alpha in a fixed order and scatters
// Q[beta][alpha] * (1-w, w) onto the two grid rows of each of alpha's samples -- no atomics, deterministic.
__global__ void __launch_bounds__(128)
build_w_kernel(DevProblem P, double* __restrict__ W, int a_pad)
{
    const int x = blockIdx.y;
    const int beta = blockIdx.x * blockDim.x + threadIdx.x;
    if (beta >= P.A || P.w_offset[x] < 0) return;             // rows of other ranks are not held
    const int A = P.A, n_off = P.n_off, n_fold = n_off / A;
    const int Kx = P.w_k[x], p_lo = P.w_plo[x];
    double* row = W + P.w_offset[x] + (size_t)beta * Kx;
    for (int alpha = 0; alpha < A; alpha++) {
        const double q = P.quadrant[(size_t)beta * A + alpha];
        for (int f = 0; f < n_fold; f++) {
            const int o = folded_offset(A, n_fold, alpha, f);
            const int u = P.upper[(size_t)x * n_off + o];
            if (u == 0) continue;                            // sample beyond the grid
            const double w = P.weight[(size_t)x * n_off + o];
            row[u - 1 - p_lo] += q * (1.0 - w);
            row[u - p_lo] += q * w;
        }
    }
    (void)a_pad;
}

template <int BM, int BN, bool BATCH>
void launch_w_variant(const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, cudaStream_t stream)
{
    using Cfg = WConfig<BM, BN>;
    const size_t smem = Cfg::smem_bytes(G);
    static bool configured[64] = {};
    int device = 0;
    cudaGetDevice(&device);
    if (!configured[device & 63]) {
        cudaFuncSetAttribute(factored_w_kernel<BM, BN, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured[device & 63] = true;
    }
    static const StepArgs none{};
    factored_w_kernel<BM, BN, BATCH><<<grid, Cfg::NT, smem, stream>>>(one ? *one : none, d_batch);
}

template <bool BATCH>
void dispatch_w(const FactoredPlan& shape, const StepArgs* one, const StepArgs* d_batch, dim3 grid, int G, cudaStream_t stream)
{
    if (shape.bm == 80 && shape.bn == 128) return launch_w_variant<80, 128, BATCH>(one, d_batch, grid, G, stream);
    if (shape.bm == 80 && shape.bn == 64) return launch_w_variant<80, 64, BATCH>(one, d_batch, grid, G, stream);
    if (shape.bm == 40 && shape.bn == 128) return launch_w_variant<40, 128, BATCH>(one, d_batch, grid, G, stream);
    return launch_w_variant<40, 64, BATCH>(one, d_batch, grid, G, stream);
}

}  // namespace

void launch_factored_w(const FactoredPlan& plan, const StepArgs& one, cudaStream_t stream)
{
    dispatch_w<false>(plan, &one, nullptr, dim3(plan.slabs(), local_rows(one.P), 1), one.P.G, stream);
}

void launch_factored_w_batch(const FactoredPlan& shape, const StepArgs* d_batch, int n, int max_slabs, int max_rows, int max_G,
                             cudaStream_t stream)
{
    dispatch_w<true>(shape, nullptr, d_batch, dim3(max_slabs, max_rows, n), max_G, stream);
}

void launch_build_w(const DevProblem& P, double* W, int a_pad, cudaStream_t stream)
{
    build_w_kernel<<<dim3((P.A + 127) / 128, P.G), 128, 0, stream>>>(P, W, a_pad);
}

}  // namespace ldm
