// The prefix-sum evaluation of the m*C integral (Solver.cpp:277-316) for sm_100a.
//
// Same factorisation as ldm_factored.cu --
//
//     ext[x][y] = sum_{beta} sum_{o_j in beta} L[y][o_j] . H_x[beta][:],      H_x = Q . E_x,
//     E_x[alpha][q] = sum_{o_i in alpha} (L[x][o_i] . C[:, q])
//
// -- but for the kernel tables the reference actually builds (Solver.cpp:98-145: a Gaussian truncated to the disc
// r <= 3 sigma) Q is separable inside the disc: Q[beta][alpha] = q[beta] q[alpha] for alpha <= a_star[beta], 0 beyond,
// a_star the quarter circle as a staircase (ldm_model.h SeparableKernel, verified per table at ldm_create).  So
//
//     H_x[beta][:] = q[beta] * R_x[a_star[beta]][:],      R_x[a][:] = sum_{alpha <= a} q[alpha] E_x[alpha][:]
//
// and the dense A x A x G contraction per row becomes ONE running sum along alpha: P*A additions per step instead of
// P*A*A fused multiply-adds.  What remains is gather work: 4 loads per element of E, 2 shared-memory gathers per
// (node y, signed offset) on the y side.
//
// One CTA owns one row x and one chunk of bn columns q (runtime shape: the H tile is A x bn doubles in shared memory):
//   1. tap tables of the row (time-invariant; resident across steps in the persistent kernel), q folded into the weights;
//   2. threads = (column, group): E on the fly from C (coalesced along q), running sum along alpha (the groups of a column
//      exchange block sums once per super-chunk of alphas), and whenever alpha is the last in-disc column of some rows
//      beta, H[beta][col] = q[beta] * sum;
//   3. y side: the (y block, offset interval) pieces that touch this chunk are dealt out evenly over the warps
//      (ldm_model.h ScanPieces); lanes = 32 consecutive nodes y read the transposed tables coalesced, gather the two H
//      columns of their interpolation pair and accumulate; per-piece partial sums are added up in list order;
//   4. per-chunk partial sums to global memory; finish_factored_kernel adds the chunks, subtracts the ring correction
//      and closes the step.
#include <algorithm>
#include <cstdlib>

#include "ldm_device.h"
#include "ldm_pointwise.cuh"

#ifndef LDM_Y_UNROLL
#define LDM_Y_UNROLL 4
#endif
#ifndef LDM_Y_UNROLL_WIDE
#define LDM_Y_UNROLL_WIDE 6
#endif

namespace ldm {

// Optional phase timeline (debug build `make phase-clocks`): thread 0 of every CTA stores clock64() at the phase
// boundaries; read back by tools/scan_clocks.py.  Off in the product build.
#ifdef LDM_PHASE_CLOCKS
__device__ long long g_scan_clock[148 * 8][12];
#define SCAN_MARK(i) do { if (threadIdx.x == 0) g_scan_clock[(blockIdx.y * gridDim.x + blockIdx.x) % (148 * 8)][i] = clock64(); } while (0)
#else
#define SCAN_MARK(i) do { } while (0)
#endif

namespace {

constexpr int kScanThreadsMax = 512;       // launch bound; a CTA runs ng * bn threads (scan_pick_threads)
constexpr int kScanLead = 2;              // doubles in front of column 0 of an H row (keeps column 0 16-byte aligned); index
                                          // kScanLead - 1 is the zero column that the "beyond the grid" pair reads
// E rows whose loads are in flight together per thread (4 loads each): 2 in the 256-thread CTAs that small problems use
// (fewer registers: four of those CTAs share an SM); in the 512-thread CTAs 4 or 5 -- whichever pads the A alphas to
// fewer whole super-chunks of ng * unroll, and 5 on a tie (one block-sum exchange fewer: A = 76, ng = 4 takes 4 super-chunks
// of 20 instead of 5 of 16; A = 245, ng = 8 stays with 8 of 32, where 5 would pad to 280).  A function of the CTA shape and A
// alone, so host and device agree and every path of a handle sums in the same order.
__host__ __device__ inline int scan_unroll(int NT, int A, int ng)
{
    if (NT <= 256) return 2;
    const int pad4 = ((A + ng * 4 - 1) / (ng * 4)) * ng * 4, pad5 = ((A + ng * 5 - 1) / (ng * 5)) * ng * 5;
    return pad5 <= pad4 ? 5 : 4;
}

struct ScanLayout {                       // byte offsets into dynamic shared memory, identical on host and device
    int ld, a_pad, n_yblk, part_stride, ng;
    size_t hs, part, rowsum, rowbuf, bsum, tap_w, qv, tap_row, sel, end;
};

__host__ __device__ inline size_t align32(size_t v) { return (v + 31) & ~(size_t)31; }

// `pieces`: the largest piece count of a chunk's y-side work list (one partial sum per piece and lane)
__host__ __device__ inline ScanLayout scan_layout(int G, int A, int bn, int NT, int pieces)
{
    ScanLayout L;
    L.ld = bn + kScanLead;
    L.n_yblk = (G + 31) / 32;
    L.part_stride = L.n_yblk * 32;
    L.ng = NT / bn > 1 ? NT / bn : 1;                     // thread groups of bn columns sharing the alphas of a super-chunk
    const int span = L.ng * scan_unroll(NT, A, L.ng);
    L.a_pad = ((A + span - 1) / span) * span;             // whole super-chunks (zero-weight rows behind A)
    size_t at = 0;
    L.hs = at;      at = align32(at + sizeof(double) * (size_t)A * L.ld);
    L.part = at;    at = align32(at + sizeof(double) * 32 * (size_t)pieces);
    L.rowsum = at;  at = align32(at + sizeof(double) * (size_t)L.part_stride);
    L.rowbuf = at;  at = align32(at + sizeof(double) * 2 * (size_t)L.part_stride);   // persistent kernel: the new row of C, the row of the WC functional
    L.bsum = at;    at = align32(at + sizeof(double) * 2 * (size_t)L.ng * bn);
    L.tap_w = at;   at = align32(at + sizeof(double) * 4 * (size_t)L.a_pad);
    L.qv = at;      at = align32(at + sizeof(double) * (size_t)L.a_pad);
    L.tap_row = at; at = align32(at + sizeof(int) * 4 * (size_t)L.a_pad);
    L.sel = at;     at = align32(at + sizeof(int2) * (size_t)L.a_pad);
    L.end = at;
    return L;
}

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);   // fixed butterfly
    return v;
}

// One CSR row of the ring correction (ldm_model.h RingOperator) by one warp: lanes stride the entries, four gathers in
// flight, accumulation in entry order with explicit fma -- the same bits as ring_kernel (ldm_kernels.cu).
__device__ __forceinline__ double ring_row_dot(const DevProblem& P, const double* C, int row, int lane)
{
    const int end = P.ring_begin[row + 1];
    int en = P.ring_begin[row] + lane;
    double acc = 0.0;
    for (; en + 96 < end; en += 128) {
        const double w0 = P.ring_weight[en], w1 = P.ring_weight[en + 32], w2 = P.ring_weight[en + 64], w3 = P.ring_weight[en + 96];
        const double c0 = C[P.ring_node[en]], c1 = C[P.ring_node[en + 32]], c2 = C[P.ring_node[en + 64]], c3 = C[P.ring_node[en + 96]];
        acc = fma(w0, c0, acc); acc = fma(w1, c1, acc); acc = fma(w2, c2, acc); acc = fma(w3, c3, acc);
    }
    for (; en < end; en += 32) acc = fma(P.ring_weight[en], C[P.ring_node[en]], acc);
    return warp_sum(acc);
}

// offsets sharing kernel node alpha: (A-1-alpha, A+alpha) for an odd extent, the single offset 0 for m == 2
__device__ __forceinline__ int folded_offset(int A, int n_fold, int alpha, int f)
{
    return n_fold == 1 ? 0 : (f == 0 ? A - 1 - alpha : A + alpha);
}

// (node, offset) pairs of the y side whose table reads are in flight together per lane: LDM_Y_UNROLL in the 256-thread CTAs
// (64 registers), LDM_Y_UNROLL_WIDE in the 512-thread ones (measured: 4 and 6)
constexpr int kYUnroll = LDM_Y_UNROLL;

// One run of consecutive offsets of the y side in which nothing changes but the position: the H row moves by dh doubles
// per offset (-ld on the minus side, where beta falls as the offset grows, +ld on the plus side, 0 for the single offset
// of m == 2), the tables by one row of G entries.  Pointer increments only; the weight source (shared-memory cache or
// global table) is a template parameter so that every load knows its address space.
template <bool MASKED, int YU, typename WPtr>
__device__ __forceinline__ double y_side_run(double s, const double* hrow, int dh, const unsigned short* up, WPtr wp, int G, int n,
                                             int q0, int bn)
{
    #pragma unroll YU
    for (int i = 0; i < n; i++) {
        const int u = (int)*up;
        const double w = *wp;
        if (MASKED) {
            const bool ok_a = (unsigned)(u - 1 - q0) < (unsigned)bn, ok_b = (unsigned)(u - q0) < (unsigned)bn;
            const double a = hrow[ok_a ? u - 1 : q0], b = hrow[ok_b ? u : q0];
            s += (ok_a ? a * (1.0 - w) : 0.0) + (ok_b ? b * w : 0.0);
        } else {
            // hrow[u - 1] * (1 - w) + hrow[u] * w, the roundings spelled out (product, complement, fma, sum)
            s = __dadd_rn(s, __fma_rn(hrow[u - 1], __dadd_rn(1.0, -w), __dmul_rn(hrow[u], w)));
        }
        hrow += dh; up += G; wp += G;
    }
    return s;
}

// y side over one interval of offsets [lo, hi) for one lane's node y, accumulated in offset order.  MASKED: the chunk holds
// only columns [q0, q0 + bn), the two taps of a pair are taken separately.  CACHED (persistent kernel): 16-bit interval
// indices and the first cached_w_rows weight rows live in shared memory.  The interval is cut where the side changes
// (offset A) and where the cached weight rows end, so that the loops themselves carry no selects.
template <bool MASKED, bool CACHED, int YU>
__device__ __forceinline__ double y_side_interval(const DevProblem& P, const double* Hs, int ld, int q0, int bn, int lo, int hi,
                                                  int y, const unsigned short* u_cache, const double* w_cache, int cached_w_rows)
{
    const int G = P.G, A = P.A;
    const bool single = P.n_off == 1;
    double s = 0.0;
    for (int o = lo; o < hi;) {
        const bool minus = !single && o < A;
        const bool w_shared = CACHED && o < cached_w_rows;
        int end = hi;
        if (minus) end = min(end, A);
        if (w_shared) end = min(end, cached_w_rows);
        const int beta = single ? 0 : (minus ? A - 1 - o : o - A);
        const int dh = single ? 0 : (minus ? -ld : ld);
        const size_t at = (size_t)o * G + y;
        const double* hrow = Hs + beta * ld + kScanLead - q0;
        const unsigned short* up = (CACHED ? u_cache : P.upper_t16) + at;
        if (w_shared) s = y_side_run<MASKED, YU>(s, hrow, dh, up, w_cache + at, G, end - o, q0, bn);
        else s = y_side_run<MASKED, YU>(s, hrow, dh, up, P.weight_t + at, G, end - o, q0, bn);
        o = end;
    }
    return s;
}

// The work of one CTA: row `lx` of this rank, column chunk `chunk`.
// fused (persistent kernel): the per-node sums stay in shared memory (row_sum) for the closing phase of the same CTA, nothing is written to global memory and the WC rows are left to the caller.
// `mid` runs between the two phases (the persistent kernel requests there what its closing phase will need, so that
// those loads fly under the y side).
struct NoMid { __device__ void operator()() const {} };

template <int kScanUnroll, typename Mid = NoMid>
__device__ __forceinline__ void scan_body(const StepArgs& S, const int chunk, const int lx, const bool build_tables,
                                          const int cached_w_rows, const bool fused = false, Mid mid = Mid())
{
    const int NT = blockDim.x;                             // = ng * bn; scan_unroll(NT, A, ng) == kScanUnroll
    const DevProblem& P = S.P;
    // no __restrict__ / read-only path on the state: inside the persistent kernel other CTAs rewrite C every step
    const double* C = S.C;
    const int G = P.G, A = P.A, n_off = P.n_off, bn = P.scan_bn;
    const int n_fold = n_off / A;
    const int x = P.rank + P.world * lx;
    const int q0 = chunk * bn;
    const int tid = threadIdx.x;
    const ScanLayout L = scan_layout(G, A, bn, NT, P.scan_max_pieces);
    const int ld = L.ld;

    extern __shared__ __align__(32) unsigned char scan_smem[];
    double* Hs = reinterpret_cast<double*>(scan_smem + L.hs);            // [A][ld]
    double* part_y = reinterpret_cast<double*>(scan_smem + L.part);      // [pieces][32] partial sums of the y side
    double* row_sum = reinterpret_cast<double*>(scan_smem + L.rowsum);   // [G] fused mode: the finished per-node sums
    double* bsum = reinterpret_cast<double*>(scan_smem + L.bsum);        // [2][ng][bn] block sums of the current super-chunk
    double* tap_w = reinterpret_cast<double*>(scan_smem + L.tap_w);      // [a_pad][4] q[alpha] * weights of the four C rows feeding E[alpha]
    double* qv = reinterpret_cast<double*>(scan_smem + L.qv);            // [a_pad]
    int* tap_row = reinterpret_cast<int*>(scan_smem + L.tap_row);        // [a_pad][4] their offsets (slot * G)
    int2* sel = reinterpret_cast<int2*>(scan_smem + L.sel);              // [a_pad] rows beta closed by alpha
    unsigned short* u_cache = reinterpret_cast<unsigned short*>(scan_smem + L.end);                    // [n_off][G]
    double* w_cache = reinterpret_cast<double*>(scan_smem + align32(L.end + sizeof(unsigned short) * (size_t)n_off * G));

    SCAN_MARK(0);
    if (build_tables) {
        if (cached_w_rows >= 0) {
            for (int i = tid; i < n_off * G; i += NT) u_cache[i] = P.upper_t16[i];
            for (int i = tid; i < cached_w_rows * G; i += NT) w_cache[i] = P.weight_t[i];
        }
        for (int t = tid; t < L.a_pad * 2; t += NT) {
            const int alpha = t >> 1, f = t & 1;
            double w_lo = 0.0, w_hi = 0.0;
            int r_lo = 0, r_hi = 0;
            if (alpha < A && f < n_fold) {
                const int o = folded_offset(A, n_fold, alpha, f);
                const int u = P.upper[(size_t)x * n_off + o];
                if (u != 0) {
                    const double w = P.weight[(size_t)x * n_off + o], qa = P.qvec[alpha];
                    w_lo = qa * (1.0 - w); w_hi = qa * w;
                    r_lo = row_slot(P, u - 1) * G; r_hi = row_slot(P, u) * G;
                }
            }
            tap_w[alpha * 4 + f * 2] = w_lo; tap_w[alpha * 4 + f * 2 + 1] = w_hi;
            tap_row[alpha * 4 + f * 2] = r_lo; tap_row[alpha * 4 + f * 2 + 1] = r_hi;
        }
        for (int a = tid; a < L.a_pad; a += NT) {
            qv[a] = a < A ? P.qvec[a] : 0.0;
            sel[a] = a < A ? P.beta_range[a] : make_int2(0, 0);
        }
        for (int b = tid; b < A; b += NT) Hs[b * ld + kScanLead - 1] = 0.0;   // the zero column (never written again)
        // empty rows of the table (a suffix: a_star never grows with beta) are never closed: zero them once
        for (int i = P.scan_live_rows * ld + tid; i < A * ld; i += NT) Hs[i] = 0.0;
        __syncthreads();
    }

    SCAN_MARK(1);
    // ---- E on the fly, running sum along alpha, H rows closed as their last in-disc column passes ----
    // Threads = (column, group).  The alphas are walked in super-chunks of ng * kScanUnroll: group g takes kScanUnroll
    // consecutive alphas of every super-chunk, the groups exchange their block sums through shared memory (one barrier per
    // super-chunk, double buffered), so every thread knows the running sum in front of its block and the H rows it closes
    // are final -- no second sweep over the tile.  A warp reads 32 consecutive columns of one C row per load (two
    // 128-byte lines: the L1 wavefront count matters as much as the bytes).  This phase is bound by the L2 throughput of
    // the whole chip -- 4 loads of C per element of E and no reuse, the moment grid being finer than the quadrature near
    // the origin -- so the raw C values of the next super-chunk are requested before the barrier.
    {
        const int col = tid % bn, grp = tid / bn;
        const bool active = grp < L.ng;
        const int span = L.ng * kScanUnroll;
        const int n_sc = L.a_pad / span;
        const int q = q0 + col;
        const bool live = q < G;
        const double* c = C + (live ? q : G - 1);
        // keep the column pointer as ONE 64-bit value (the compiler otherwise folds the column into every tap offset and
        // spends four integer instructions per load on 64-bit sums); it still points to global memory
        asm volatile("" : "+l"(c));
        __builtin_assume(__isGlobal(c));
        double* hcol = Hs + kScanLead + col;
        double base = 0.0;                                 // sum over all alphas of the super-chunks before (same in every group)
        // the barrier of a super-chunk only joins the ng warps that share the same 32 columns (named barriers 1..bn/32):
        // the column sets drift apart, so the loads of one set fly while another exchanges its block sums
        const int set_id = 1 + col / 32, set_threads = L.ng * 32;
        double raw[kScanUnroll][4];
        auto issue = [&](int sc) {
            #pragma unroll
            for (int k = 0; k < kScanUnroll; k++) {
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + (sc * span + grp * kScanUnroll + k) * 4);
                raw[k][0] = c[(unsigned)ra.x]; raw[k][1] = c[(unsigned)ra.y]; raw[k][2] = c[(unsigned)ra.z]; raw[k][3] = c[(unsigned)ra.w];   // unsigned: one IMAD.WIDE.U32 per address
            }
        };
        if (active) {
            issue(0);
            for (int sc = 0; sc < n_sc; sc++) {
                double prefix[kScanUnroll];
                double* mine = bsum + (sc & 1) * L.ng * bn;
                double run = 0.0;
                #pragma unroll
                for (int k = 0; k < kScanUnroll; k++) {
                    const double4 wa = *reinterpret_cast<const double4*>(tap_w + (sc * span + grp * kScanUnroll + k) * 4);
                    run += (raw[k][0] * wa.x + raw[k][1] * wa.y) + (raw[k][2] * wa.z + raw[k][3] * wa.w);
                    prefix[k] = run;
                }
                mine[grp * bn + col] = run;
                if (sc + 1 < n_sc) issue(sc + 1);
                asm volatile("bar.sync %0, %1;" ::"r"(set_id), "r"(set_threads) : "memory");
                double before = 0.0, all = 0.0;
                for (int g = 0; g < L.ng; g++) {             // fixed order
                    const double v = mine[g * bn + col];
                    if (g < grp) before += v;
                    all += v;
                }
                const double start = base + before;
                #pragma unroll
                for (int k = 0; k < kScanUnroll; k++) {
                    const int2 rows = sel[sc * span + grp * kScanUnroll + k];       // empty beyond A
                    const double v = live ? start + prefix[k] : 0.0;
                    double* hp = hcol + rows.x * ld;
                    #pragma unroll 1                           // one row per alpha on average: no unrolled body, no remainder logic
                    for (int r = rows.x; r < rows.y; r++, hp += ld) *hp = v * qv[r];
                }
                base += all;
            }
        }
    }
    __syncthreads();
    SCAN_MARK(3);
    mid();

    // ---- y side ----
    {
        const int warp = tid >> 5, lane = tid & 31, warps = NT / 32;
        const bool masked = P.scan_chunks > 1;
        // (node, offset) pairs in flight per lane: the 256-thread CTAs live within 64 registers, the 512-thread ones can afford more
        constexpr int kYU = kScanUnroll == 2 ? kYUnroll : LDM_Y_UNROLL_WIDE;
        const int4* pieces = P.scan_piece + P.scan_chunk_begin[chunk];
        const int* warp_begin = P.scan_warp_begin + chunk * (warps + 1);
        const int* block_begin = P.scan_block_begin + chunk * (L.n_yblk + 1);
        for (int pi = warp_begin[warp]; pi < warp_begin[warp + 1]; pi++) {
            const int4 pc = pieces[pi];                        // y block, first offset, end offset
            const int y = min(pc.x * 32 + lane, G - 1);
            double s;
            if (cached_w_rows >= 0) s = y_side_interval<false, true, kYU>(P, Hs, ld, q0, bn, pc.y, pc.z, y, u_cache, w_cache, cached_w_rows);
            else if (masked) s = y_side_interval<true, false, kYU>(P, Hs, ld, q0, bn, pc.y, pc.z, y, nullptr, nullptr, 0);
            else s = y_side_interval<false, false, kYU>(P, Hs, ld, q0, bn, pc.y, pc.z, y, nullptr, nullptr, 0);
            part_y[pi * 32 + lane] = s;
        }
        __syncthreads();
        SCAN_MARK(4);
        double* out = S.partial + ((size_t)chunk * local_rows(P) + lx) * G;
        for (int y = tid; y < G; y += NT) {
            double s = 0.0;
            for (int pi = block_begin[y >> 5]; pi < block_begin[(y >> 5) + 1]; pi++) s += part_y[pi * 32 + (y & 31)];   // fixed order
            if (fused) row_sum[y] = s;
            else out[y] = s;
        }
    }

    // WC = <wc_operator, C> (Solver.cpp:265-274) row by row, folded into this launch.  Every rank reduces the FULL
    // operator (no all-reduce needed): the first-chunk CTAs of this rank share all G rows round-robin.
    if (chunk == 0 && S.wc_rows != nullptr && !fused) {
        const int n_local = local_rows(P);
        int slot = 0;
        for (int xr = lx; xr < G; xr += n_local, slot++) {
            if (slot % (NT / 32) != tid / 32) continue;          // one warp per row
            const double s = wc_row_dot(P.wc_operator + (size_t)xr * G, C + (size_t)row_slot(P, xr) * G, G, tid % 32);
            if (tid % 32 == 0) S.wc_rows[xr] = s;
        }
    }
    __syncthreads();       // shared memory is reused by the next row / step of this CTA
    SCAN_MARK(5);
}

// scan_body with the unroll factor of this CTA shape: U = 4 or 5 when the launch picked the kernel instance by it (the
// single-problem kernels: each instance keeps its own register allocation), U = 0 to decide here by a uniform branch (batch
// launches, whose members may differ, and the one-launch exchange option)
template <bool SMALL, int U, typename Mid = NoMid>
__device__ __forceinline__ void scan_body_of(const StepArgs& S, const int chunk, const int lx, const bool build_tables,
                                             const int cached_w_rows, const bool fused = false, Mid mid = Mid())
{
    if (SMALL) { scan_body<2>(S, chunk, lx, build_tables, cached_w_rows, fused, mid); return; }
    if (U != 0) { scan_body<(U == 0 ? 4 : U)>(S, chunk, lx, build_tables, cached_w_rows, fused, mid); return; }
    const int ng = (int)blockDim.x / S.P.scan_bn > 1 ? (int)blockDim.x / S.P.scan_bn : 1;
    if (scan_unroll((int)blockDim.x, S.P.A, ng) == 5) scan_body<5>(S, chunk, lx, build_tables, cached_w_rows, fused, mid);
    else scan_body<4>(S, chunk, lx, build_tables, cached_w_rows, fused, mid);
}

// SMALL: 256-thread CTAs capped at 64 registers (four per SM) for problems whose working set is small; otherwise 512
// threads with the registers they want.
template <bool BATCH, bool SMALL, int U>
__global__ void __launch_bounds__(SMALL ? 256 : kScanThreadsMax, SMALL ? 4 : 1)
scan_kernel(const StepArgs one, const StepArgs* __restrict__ batch)
{
    // grid = (rows, chunks): the CTAs are issued chunk by chunk, most expensive chunk first (on a stretched grid the far
    // chunks own twice the y-side pairs of the near ones), so the tail of the launch is filled with the cheap ones
    const StepArgs& S = BATCH ? batch[blockIdx.z] : one;
    if (BATCH && ((int)blockIdx.y >= S.P.scan_chunks || (int)blockIdx.x >= local_rows(S.P))) return;   // batch grids are padded
    scan_body_of<SMALL, U>(S, S.P.scan_chunk_order[blockIdx.y], blockIdx.x, true, -1);
}

// Spin until *counter >= target: relaxed polls (an acquire load in the loop would pay for its fence on every poll), then
// ONE acquire load orders everything that follows behind the writers' releases (measured: cheaper than a fence here).
__device__ __forceinline__ void wait_for_count(const unsigned int* counter, unsigned int target)
{
    unsigned int seen;
    do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
}

// Hand-over of one ring value between CTAs without a fence (see persistent_scan_kernel): the slot holds kRingEmpty until
// the value is published.  A computed NaN never has this bit pattern, so a state that has gone NaN does not hang the poll
// (and the poll is bounded anyway).
constexpr unsigned long long kRingEmpty = ~0ULL;

__device__ __forceinline__ void publish_ring(double* slot, double value)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(slot), "d"(value) : "memory");
}

__device__ __forceinline__ double take_ring(double* slot)
{
    unsigned long long bits = kRingEmpty;
    for (long long spin = 0; spin < 400000000LL && bits == kRingEmpty; spin++)
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(slot) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(kRingEmpty) : "memory");
    return __longlong_as_double((long long)bits);
}

// Grid-wide barrier of the persistent kernel: a monotone arrival counter in global memory (zeroed before the launch; the
// cooperative launch guarantees that every CTA is resident).  After the CTA barrier thread 0 arrives with a releasing
// reduction (cumulative over the CTA's writes, and nothing waits for a returned value) and polls until all CTAs of
// barrier number `index` have arrived.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int index)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        wait_for_count(counter, (index + 1u) * gridDim.x);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// xi-sharded step with the exchange fused in, ONE launch per step (ldm_shard_peer_step).  One-dimensional grid of
// local_rows * (1 + chunks) CTAs, issued in this order:
//   the first local_rows CTAs   CTA x: its share of the G row dots of the WC functional and of the ring correction, both
//                               counted in global memory as they finish;
//   then row by row             the prefix-sum integral of (local row, chunk), the chunks of a row next to each other;
//                               the chunk CTA that arrives LAST on its row closes the row -- closure, Euler update, clamp
//                               -- and stores it straight into the next replica of EVERY rank over NVLink (peer-mapped
//                               memory), so rows travel while the rows behind them are still being computed; the CTA
//                               that closes the last row publishes this rank's epoch in every rank's flag array.
// The launch is preceded by peer_wait_kernel (one warp: all ranks have published the previous step, so this rank's current
// replica is complete and nobody reads the replicas this step writes to any more) -- polling the flags from every CTA
// instead costs more than that launch (measured: 8 system-scope loads per CTA, 15 CTAs per SM).
// ------------------------------------------------------------------------------------------------------------------
template <bool SMALL>
__global__ void __launch_bounds__(SMALL ? 256 : kScanThreadsMax, 1)
scan_exchange_kernel(const StepArgs S)
{
    const DevProblem& P = S.P;
    const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = P.G, n_local = local_rows(P);
    const bool helper = (int)blockIdx.x < n_local;
    const int lx = helper ? (int)blockIdx.x : ((int)blockIdx.x - n_local) / P.scan_chunks;
    const int chunk_slot = helper ? 0 : ((int)blockIdx.x - n_local) % P.scan_chunks;
    const unsigned int step = (unsigned int)S.epoch;                        // 1-based count of steps since the upload
    unsigned int* wc_done = S.sync_words;
    unsigned int* rows_closed = S.sync_words + 1;
    unsigned int* row_arrivals = S.sync_words + 2;
    unsigned int* ring_done = S.sync_words + 2 + P.rows_per_rank;
    __shared__ int closes_row;
    __shared__ double wc_shared;

    if (helper) {
        // WC = <wc_operator, C> row by row (every rank reduces the FULL operator: no all-reduce), then this CTA's share
        // of the ring correction (CSR rows dealt out evenly by entry count)
        int slot = 0, mine = 0;
        for (int xr = lx; xr < G; xr += n_local, slot++) {
            if (slot % (NT / 32) == warp) {
                const double s = wc_row_dot(P.wc_operator + (size_t)xr * G, S.C + (size_t)row_slot(P, xr) * G, G, lane);
                if (lane == 0) S.wc_rows[xr] = s;
            }
            mine++;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(wc_done), "r"((unsigned int)mine) : "memory");
        for (int row = P.ring_cta_begin[lx] + warp; row < P.ring_cta_begin[lx + 1]; row += NT / 32) {
            const double acc = ring_row_dot(P, S.C, row, lane);
            if (lane == 0) {
                const int point = P.ring_point[row];
                S.ring[point] = acc;
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ring_done + point / G), "r"(1u) : "memory");
            }
        }
        return;
    }

    StepArgs body = S;
    body.wc_rows = nullptr;                                                 // the helper CTAs own the WC rows here
    scan_body_of<SMALL, 0>(body, P.scan_chunk_order[chunk_slot], lx, true, -1);
    __threadfence();                                                        // this thread's partial sums, before the arrival
    __syncthreads();
    if (tid == 0) closes_row = (atomicAdd(row_arrivals + lx, 1u) + 1u == step * (unsigned int)P.scan_chunks);
    __syncthreads();
    if (!closes_row) return;

    // ---- close row lx: everything finish_factored_kernel does for it, same arithmetic ----
    if (tid == 0) {
        wait_for_count(wc_done, step * (unsigned int)G);
        wait_for_count(ring_done + lx, step * (unsigned int)(P.ring_row_begin[lx + 1] - P.ring_row_begin[lx]));
    }
    __syncthreads();
    if (warp == 0) {
        double t = 0.0;
        for (int r = lane; r < G; r += 32) t += __ldcg(S.wc_rows + r);      // wc_total(): the same order
        t = warp_sum(t);
        if (lane == 0) wc_shared = t;
    }
    __syncthreads();
    const double N = S.N_cur[0];
    const double WC = __dmul_rn(add_outside(P, wc_shared, N, S.C, S.scalars, P.wc_outside, 0.0), P.hh);   // Solver.cpp:311, 314
    if (lx == 0 && tid == 0) first_moment_update(P, N, WC, S.dt, S.scalars, S.N_next);
    const int n = n_local * G;
    for (int y = tid; y < G; y += NT) {
        const int ql = lx * G + y;
        const int q = (P.rank + P.world * lx) * G + y;
        double inside = 0.0;
        for (int t = 0; t < P.scan_chunks; t++) inside += __ldcg(S.partial + (size_t)t * n + ql);
        inside -= __ldcg(S.ring + ql);
        const size_t slot = (size_t)(P.rank * P.rows_per_rank + lx) * G + y;   // this node in the gathered layout
        const double c = S.C[slot];
        const double mc = __dmul_rn(add_outside(P, inside, N, S.C, S.scalars, P.out_weight[q], 0.0), P.hh);
        const double dc = second_moment_rhs(P, mc, c, N, WC, P.m_nodes[q], P.w_nodes[q]);
        double v = __dadd_rn(c, __dmul_rn(dc, S.dt));                       // Solver.cpp:577
        v = v < 0.0 ? 0.0 : v;                                              // Solver.cpp:578-580
        for (int r = 0; r < S.peer_count; r++) S.peer_C_next[r][slot] = v;  // the all-gather: NVLink peer stores
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && atomicAdd(rows_closed, 1u) + 1u == step * (unsigned int)n_local) {
        __threadfence_system();
        for (int r = 0; r < S.peer_count; r++)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(S.peer_flags[r] + P.rank), "l"(S.epoch) : "memory");
    }
}

// Persistent multi-step kernel (cooperative launch, one CTA per moment-grid row): all Euler steps of a solve inside one
// launch, for grids whose rows are one chunk each.  Per step every CTA
//   1. sums the G row dots of the WC functional (written by the closing phases of the step before),
//   2. computes its share of the sparse ring correction (dealt out evenly over all CTAs, counted per grid row), then runs
//      scan_body for its row (per-node sums stay in shared memory),
//   3. closes its own row -- closure, Euler update, clamp, history -- and forms its row dot of WC for the NEXT state,
// and ONE grid barrier ends the step.  Tap tables and the y-side tables stay in shared memory across steps.
template <bool SMALL, int U>
__global__ void __launch_bounds__(SMALL ? 256 : kScanThreadsMax, 1)
persistent_scan_kernel(const PersistentArgs A)
{
    const int NT = blockDim.x;
    __shared__ double wc_shared;
    StepArgs S = A.step;
    const DevProblem& P = S.P;
    const int G = P.G, tid = threadIdx.x, x = blockIdx.x;          // one rank, natural layout: row x = CTA x
    const ScanLayout L = scan_layout(G, P.A, P.scan_bn, NT, P.scan_max_pieces);
    extern __shared__ __align__(32) unsigned char scan_smem[];
    double* row_sum = reinterpret_cast<double*>(scan_smem + L.rowsum);            // [G] after scan_body(fused)
    int parity = A.parity;
    unsigned int barrier_index = 0;

    // prologue: the row dots of WC for the initial state
    if (tid < 32) {
        const double s = wc_row_dot(P.wc_operator + (size_t)x * G, A.C[parity] + (size_t)x * G, G, tid);
        if (tid == 0) A.wc_rows[parity * G + x] = s;
    }
    grid_barrier(A.barrier, barrier_index++);

    double* c_new = reinterpret_cast<double*>(scan_smem + L.rowbuf);              // [G] this row at the new time
    double* op_row = c_new + L.part_stride;                                        // [G] this row of the WC functional
    for (int y = tid; y < G; y += NT) op_row[y] = P.wc_operator[(size_t)x * G + y];
    const int warp = tid >> 5, lane = tid & 31, last_warp = NT / 32 - 1;
    const bool closer = tid < G;                                                   // thread y closes node (x, y); G <= bn <= NT
    const int q = x * G + (closer ? tid : 0);
    const bool has_ring = closer && P.ring_present[q] != 0;                       // node (x, y) owns ring samples
    const double ow = P.out_weight[q], mn = P.m_nodes[q], wn = P.w_nodes[q];      // time-invariant: registers for the whole launch

    // This CTA's share of the ring correction (the far rows of the grid own almost all ring entries -- row G-1 alone an
    // eighth of them -- so the CSR rows are dealt out evenly by entry count over ALL CTAs, whichever grid row they belong
    // to).  The share is time-invariant: its node indices and weights are copied to shared memory once, one CSR row per
    // warp, so that a step costs ONE round trip to L2 (the gathers of C) instead of a chain of four.
    constexpr int kRingPerLane = 8;
    const int share_r0 = P.ring_cta_begin[x], share_r1 = P.ring_cta_begin[x + 1];
    const int share_e0 = P.ring_begin[share_r0];
    const bool ring_cached = A.ring_cache_offset >= 0;
    int* rc_node = reinterpret_cast<int*>(scan_smem + (ring_cached ? A.ring_cache_offset : 0));
    double* rc_weight = reinterpret_cast<double*>(rc_node + ((P.ring_share_entries + 1) & ~1));
    const int my_row = share_r0 + warp;
    int my_begin = 0, my_end = 0, my_point = 0;
    if (ring_cached) {
        for (int e = tid; e < P.ring_begin[share_r1] - share_e0; e += NT) { rc_node[e] = P.ring_node[share_e0 + e]; rc_weight[e] = P.ring_weight[share_e0 + e]; }
        if (my_row < share_r1) { my_begin = P.ring_begin[my_row] - share_e0; my_end = P.ring_begin[my_row + 1] - share_e0; my_point = P.ring_point[my_row]; }
        __syncthreads();
    }

    for (int k = 0; k < A.n_steps; k++) {
        SCAN_MARK(10);
        S.C = A.C[parity]; S.C_next = A.C[parity ^ 1];
        S.N_cur = A.N + parity; S.N_next = A.N + (parity ^ 1);
        const long long point = A.first_point + k;                 // index of the state this step starts from
        if (x == 0 && tid == 0 && A.observables) {
            A.observables[2 * point] = S.N_cur[0];                                  // Values[k][0]       (Solver.cpp:1011)
            A.observables[2 * point + 1] = S.C[(size_t)0 * G + (G - 1)];            // Values[k][G]       (Solver.cpp:1014)
        }
        // The ring share.  Hand-over without fences or counters (a releasing reduction costs ~4 k cycles here): the values go
        // to this step's half of a ping-pong buffer whose entries hold an all-ones NaN until they are written; the owner
        // of a node polls its entry, takes the value and puts the NaN back for the step after next (two grid barriers lie
        // between that store and the next write of the entry).  One 8-byte relaxed store, one 8-byte relaxed load.
        double* ring_slot = A.ring_pingpong + (size_t)(k & 1) * G * G;
        if (ring_cached) {
            double cv[kRingPerLane];
            #pragma unroll
            for (int j = 0; j < kRingPerLane; j++) {                 // all gathers of the row in flight at once
                const int e = my_begin + lane + 32 * j;
                cv[j] = e < my_end ? S.C[rc_node[e]] : 0.0;
            }
            double acc = 0.0;
            #pragma unroll
            for (int j = 0; j < kRingPerLane; j++) {                 // the accumulation order of ring_row_dot / ring_kernel
                const int e = my_begin + lane + 32 * j;
                if (e < my_end) acc = fma(rc_weight[e], cv[j], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0 && my_row < share_r1) publish_ring(ring_slot + my_point, acc);
        } else {
            for (int row = share_r0 + warp; row < share_r1; row += NT / 32) {
                const double acc = ring_row_dot(P, S.C, row, lane);
                if (lane == 0) publish_ring(ring_slot + P.ring_point[row], acc);
            }
        }
        SCAN_MARK(11);
        // Everything the closing phase reads from global memory is requested between the two phases of scan_body and
        // lands while the y side runs: the old C of this row, its ring correction (the owner first checks that all
        // CSR rows of its grid row have been counted -- a whole E phase after they were dealt out, so no real wait),
        // N, and the G row dots of the WC functional (written by the closing phases of the step before).
        double c_old = 0.0, ring_v = 0.0, N = 0.0, wc_part[4] = {0.0, 0.0, 0.0, 0.0};
        const double* wc_rows_now = A.wc_rows + parity * G;
        auto mid = [&]() {
            if (closer) {
                if (has_ring) ring_v = take_ring(ring_slot + q);
                c_old = S.C[q];
            }
            N = S.N_cur[0];
            if (warp == last_warp) {
                #pragma unroll
                for (int j = 0; j < 4; j++) wc_part[j] = lane + 32 * j < G ? wc_rows_now[lane + 32 * j] : 0.0;
            }
        };
        scan_body_of<SMALL, U>(S, 0, x, k == 0, A.cached_w_rows, true, mid);   // ends with a CTA barrier
        SCAN_MARK(6);
        if (warp == last_warp) {                                   // wc_total(): the same order, from the registers
            double t = 0.0;
            #pragma unroll
            for (int j = 0; j < 4; j++) if (lane + 32 * j < G) t += wc_part[j];
            t = warp_sum(t);
            if (lane == 0) wc_shared = t;
        }
        __syncthreads();
        const double WC = __dmul_rn(add_outside(P, wc_shared, N, S.C, S.scalars, P.wc_outside, 0.0), P.hh);   // Solver.cpp:311, 314
        double* values_next = A.values ? A.values + (point + 1) * A.stride : nullptr;
        double* derivs_here = A.derivs ? A.derivs + point * A.stride : nullptr;
        if (x == 0 && tid == 0) {
            first_moment_update(P, N, WC, S.dt, S.scalars, S.N_next);
            if (values_next) values_next[0] = S.N_next[0];
            if (derivs_here) derivs_here[0] = S.scalars[1];
        }
        if (closer) {
            const double inside = (0.0 + row_sum[tid]) - ring_v;
            const double mc = __dmul_rn(add_outside(P, inside, N, S.C, S.scalars, ow, 0.0), P.hh);
            const double dc = second_moment_rhs(P, mc, c_old, N, WC, mn, wn);
            double v = __dadd_rn(c_old, __dmul_rn(dc, S.dt));                      // Solver.cpp:577
            v = v < 0.0 ? 0.0 : v;                                                 // Solver.cpp:578-580
            S.C_next[q] = v;
            c_new[tid] = v;
            if (values_next) values_next[1 + q] = v;
            if (derivs_here) derivs_here[1 + q] = dc;
        }
        __syncthreads();
        if (warp == 0) {     // this row's dot of the WC functional for the next state
            const double s = wc_row_dot(op_row, c_new, G, lane);
            if (lane == 0) A.wc_rows[(parity ^ 1) * G + x] = s;
        }
        SCAN_MARK(8);
        grid_barrier(A.barrier, barrier_index++);
        SCAN_MARK(9);
        parity ^= 1;
    }
}

constexpr size_t kSmemBudget = 226 * 1024;

void configure_once(const void* kernel, bool& done)
{
    if (!done) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
        done = true;
    }
}

}  // namespace

#ifdef LDM_PHASE_CLOCKS
extern "C" int ldm_debug_scan_clocks(long long* out, int rows)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_scan_clock, sizeof(long long) * 12 * (size_t)rows);
}
#endif

size_t scan_smem_bytes(int G, int A, int bn, int nt, int max_pieces) { return scan_layout(G, A, bn, nt, max_pieces).end; }

// no chunk has more pieces than this (every y block at most kScanRanges intervals, every warp at most one extra cut)
static int piece_bound(int G, int nt) { return 3 * ((G + 31) / 32) + nt / 32; }

// Threads per CTA = ng * bn (ng groups of bn columns).  Small working sets (several CTAs fit one SM's shared memory)
// take 256-thread CTAs, everything else 257..512 threads.
int scan_pick_threads(int G, int A, int bn)
{
    if (const char* env = std::getenv("LDM_SCAN_NT")) {          // tuning only
        const int v = std::atoi(env);
        if ((v == 256 || v == 512) && v >= bn) return (v / bn) * bn;
    }
    if (bn <= 64 && scan_layout(G, A, bn, 256, piece_bound(G, 256)).end <= 48 * 1024) return (256 / bn) * bn;
    return (kScanThreadsMax / bn) * bn;                    // bn = 32, 64, 96, 128 -> 512, 512, 480, 512
}

int scan_pick_bn(int G, int A)
{
    auto fits = [&](int bn) {
        const int nt = scan_pick_threads(G, A, bn);
        return scan_smem_bytes(G, A, bn, nt, piece_bound(G, nt)) <= kSmemBudget;
    };
    if (const char* env = std::getenv("LDM_SCAN_BN")) {
        const int v = std::atoi(env);
        if (v >= 32 && v <= 128 && v % 32 == 0 && fits(v)) return v;
    }
    // widest chunk (a multiple of 32, at most 128 columns) whose H tile fits; then the evenest split of the row
    int bn = 128;
    while (bn > 32 && !fits(bn)) bn -= 32;
    if (!fits(bn)) return 0;                                     // table too tall for one CTA: dense contraction instead
    const int chunks = (G + bn - 1) / bn;
    const int even = (((G + chunks - 1) / chunks) + 31) & ~31;
    return even < bn ? even : bn;
}

template <bool BATCH, bool SMALL, int U>
static void launch_scan_shape(const StepArgs* one, const StepArgs* d_batch, dim3 grid, int nt, size_t smem, cudaStream_t stream)
{
    static bool configured[64] = {};
    int device = 0;
    cudaGetDevice(&device);
    configure_once((const void*)scan_kernel<BATCH, SMALL, U>, configured[device & 63]);
    static const StepArgs none{};
    scan_kernel<BATCH, SMALL, U><<<grid, nt, smem, stream>>>(one ? *one : none, d_batch);
}

// unroll: 4 or 5 for a single problem (scan_unroll of its shape), 0 for a batch (decided per member inside the kernel)
template <bool BATCH>
static void launch_scan_variant(const StepArgs* one, const StepArgs* d_batch, dim3 grid, int nt, int unroll, size_t smem, cudaStream_t stream)
{
    if (nt <= 256) launch_scan_shape<BATCH, true, 0>(one, d_batch, grid, nt, smem, stream);
    else if (BATCH || unroll == 0) launch_scan_shape<BATCH, false, 0>(one, d_batch, grid, nt, smem, stream);
    else if (unroll == 5) launch_scan_shape<false, false, 5>(one, d_batch, grid, nt, smem, stream);
    else launch_scan_shape<false, false, 4>(one, d_batch, grid, nt, smem, stream);
}

static int scan_unroll_of(const DevProblem& P)
{
    const int ng = P.scan_nt / P.scan_bn > 1 ? P.scan_nt / P.scan_bn : 1;
    return scan_unroll(P.scan_nt, P.A, ng);
}

void launch_scan(const StepArgs& one, cudaStream_t stream)
{
    const DevProblem& P = one.P;
    launch_scan_variant<false>(&one, nullptr, dim3(local_rows(P), P.scan_chunks, 1), P.scan_nt, scan_unroll_of(P),
                               scan_smem_bytes(P.G, P.A, P.scan_bn, P.scan_nt, P.scan_max_pieces), stream);
}

// every member of a batch launch shares the CTA shape nt (= ng * bn); shared memory is sized for the largest member
void launch_scan_batch(const StepArgs* d_batch, int n, int nt, int max_chunks, int max_rows, size_t max_smem, cudaStream_t stream)
{
    launch_scan_variant<true>(nullptr, d_batch, dim3(max_rows, max_chunks, n), nt, 0, max_smem, stream);
}

// the same CTA shape as the one-launch-per-step kernel of this problem: the summation order follows the shape, and the
// persistent, per-step, batched and sharded paths of one handle must agree bit for bit
void launch_scan_exchange(const StepArgs& one, cudaStream_t stream)
{
    const DevProblem& P = one.P;
    const size_t smem = scan_smem_bytes(P.G, P.A, P.scan_bn, P.scan_nt, P.scan_max_pieces);
    const dim3 grid(local_rows(P) * (1 + P.scan_chunks), 1, 1);
    static bool configured[64][2] = {};
    int device = 0;
    cudaGetDevice(&device);
    if (P.scan_nt <= 256) {
        configure_once((const void*)scan_exchange_kernel<true>, configured[device & 63][0]);
        scan_exchange_kernel<true><<<grid, P.scan_nt, smem, stream>>>(one);
    } else {
        configure_once((const void*)scan_exchange_kernel<false>, configured[device & 63][1]);
        scan_exchange_kernel<false><<<grid, P.scan_nt, smem, stream>>>(one);
    }
}

cudaError_t launch_persistent_scan(const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid)
{
    const DevProblem& P = args.step.P;
    const int G = P.G, n_off = P.n_off, nt = P.scan_nt;
    const void* kernel = nt <= 256 ? (const void*)persistent_scan_kernel<true, 0>
                       : scan_unroll_of(P) == 5 ? (const void*)persistent_scan_kernel<false, 5> : (const void*)persistent_scan_kernel<false, 4>;
    // shared memory: the step's working set, then the 16-bit y-side index table, then as many weight rows as still fit
    const size_t base = scan_smem_bytes(G, P.A, P.scan_bn, nt, P.scan_max_pieces);
    const size_t u_bytes = align32(sizeof(unsigned short) * (size_t)n_off * G);
    PersistentArgs with_cache = args;
    with_cache.cached_w_rows = -1;
    with_cache.ring_cache_offset = -1;
    // the CTA's ring share (one CSR row per warp, at most 8 entries per lane): 4-byte node indices, then 8-byte weights
    size_t ring_bytes = 0;
    if (P.ring_share_rows <= nt / 32 && P.ring_share_row_len <= 32 * 8 && std::getenv("LDM_SCAN_NO_RING_CACHE") == nullptr)
        ring_bytes = align32(sizeof(int) * (size_t)((P.ring_share_entries + 1) & ~1) + sizeof(double) * (size_t)P.ring_share_entries + 8);
    size_t smem = base;
    if (base + ring_bytes <= kSmemBudget && ring_bytes > 0) {
        with_cache.ring_cache_offset = (long long)base;
        smem = base + ring_bytes;
    } else {
        ring_bytes = 0;
    }
    if (base + ring_bytes + u_bytes <= kSmemBudget && std::getenv("LDM_SCAN_NO_CACHE") == nullptr) {
        const int rows = (int)std::min<size_t>((size_t)n_off, (kSmemBudget - base - ring_bytes - u_bytes) / ((size_t)G * sizeof(double)));
        with_cache.cached_w_rows = rows;
        with_cache.ring_cache_offset = ring_bytes > 0 ? (long long)(base + u_bytes + (size_t)rows * G * sizeof(double)) : -1;
        smem = base + ring_bytes + u_bytes + (size_t)rows * G * sizeof(double);
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    if (e != cudaSuccess) return e;
    int device = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem);
    if (e != cudaSuccess) return e;
    if (max_grid) *max_grid = sms * per_sm;
    if (grid <= 0) return cudaSuccess;                  // query only
    if (grid > sms * per_sm) return cudaErrorCooperativeLaunchTooLarge;
    void* params[] = {&with_cache};
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(nt), params, smem, stream);
}

}  // namespace ldm
