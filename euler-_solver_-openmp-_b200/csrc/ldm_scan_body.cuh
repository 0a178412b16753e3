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
//      (ldm_model.h ScanPieces); lanes = 32 consecutive nodes y read the node-contiguous tables coalesced -- two consecutive
//      offsets per load (y_table_index) --, gather the two H columns of their interpolation pair and accumulate; per-piece
//      partial sums are added up in list order;
//   4. per-chunk partial sums to global memory; the closing code (finish_factored_kernel, or the closing phase of a
//      persistent kernel) adds the chunks, subtracts the ring correction and closes the step.
//
// This header holds the device code that the one-launch-per-step kernel (ldm_scan.cu) and the persistent multi-step
// kernels (ldm_persist.cu) share: the same code, so the same summation order and the same bits on every path.
#pragma once
#include <algorithm>

#include "ldm_device.h"
#include "ldm_pointwise.cuh"

#ifndef LDM_Y_UNROLL
#define LDM_Y_UNROLL 4
#endif
#ifndef LDM_Y_UNROLL_WIDE
#define LDM_Y_UNROLL_WIDE 8
#endif

namespace ldm {

// Optional phase timeline (debug build `make phase-clocks`): thread 0 of every CTA stores clock64() at the phase
// boundaries; read back by tools/scan_clocks.py.  Off in the product build.
#ifdef LDM_PHASE_CLOCKS
static __device__ long long g_scan_clock[148 * 8][12];   // one copy per translation unit (ldm_debug_scan_clocks / ldm_debug_persist_clocks)
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

constexpr size_t kSmemBudget = 226 * 1024; // dynamic shared memory a CTA of this path may ask for

struct ScanLayout {                       // byte offsets into dynamic shared memory, identical on host and device
    int ld, a_pad, n_yblk, part_stride, ng;
    size_t hs, part, rowsum, rowbuf, bsum, tap_w, qv, tap_row, sel, work, end;
};

__host__ __device__ inline size_t align32(size_t v) { return (v + 31) & ~(size_t)31; }

// `pieces`: the largest piece count of a chunk's y-side work list (one partial sum per piece and lane)
__host__ __device__ inline ScanLayout scan_layout(int G, int A, int bn, int NT, int pieces)
{
    ScanLayout L;
    L.ld = bn + kScanLead + (G > bn ? 1 : 0);             // several chunks per row: one more zero column behind the last (y_side_run)
    L.n_yblk = (G + 31) / 32;
    L.part_stride = L.n_yblk * 32;
    L.ng = NT / bn > 1 ? NT / bn : 1;                     // thread groups of bn columns sharing the alphas of a super-chunk
    const int span = L.ng * scan_unroll(NT, A, L.ng);
    L.a_pad = ((A + span - 1) / span) * span;             // whole super-chunks (zero-weight rows behind A)
    size_t at = 0;
    L.hs = at;      at = align32(at + sizeof(double) * ((size_t)A * L.ld + 1));             // + the zero behind the last row's right zero column
    L.part = at;    at = align32(at + sizeof(double) * 32 * (size_t)pieces);
    L.rowsum = at;  at = align32(at + sizeof(double) * (size_t)L.part_stride);
    L.rowbuf = at;  at = align32(at + sizeof(double) * 2 * (size_t)L.part_stride);   // persistent kernel: the new row of C, the row of the WC functional
    L.bsum = at;    at = align32(at + sizeof(double) * 2 * (size_t)L.ng * bn);
    L.tap_w = at;   at = align32(at + sizeof(double) * 4 * (size_t)L.a_pad);
    L.qv = at;      at = align32(at + sizeof(double) * (size_t)L.a_pad);
    L.tap_row = at; at = align32(at + sizeof(int) * 4 * (size_t)L.a_pad);
    L.sel = at;     at = align32(at + sizeof(int4) * (size_t)L.a_pad);                       // rows closed by alpha + q of the first of them
    // the chunk's y-side work list (pieces, first piece of every warp and of every y block): read once per table build instead
    // of from global memory in front of every piece
    L.work = at;    at = align32(at + sizeof(int4) * (size_t)pieces + sizeof(int) * (size_t)(NT / 32 + 1 + L.n_yblk + 1));
    L.end = at;
    return L;
}

inline int scan_unroll_of(const DevProblem& P)
{
    const int ng = P.scan_nt / P.scan_bn > 1 ? P.scan_nt / P.scan_bn : 1;
    return scan_unroll(P.scan_nt, P.A, ng);
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
// (64 registers), LDM_Y_UNROLL_WIDE in the 512-thread ones (measured: 4 and, with the offsets fetched two by two, 8: 6 and 12 are slower)
constexpr int kYUnroll = LDM_Y_UNROLL;

// One (node y, offset) pair of the y side: hrow[u - 1] * (1 - w) + hrow[u] * w added to s, the roundings spelled out (product,
// complement, fma, sum).  MASKED (rows of several chunks): the tile holds columns [q0, q0 + bn) only, and `u` is the pair's
// POSITION in the tile's row -- kScanLead - 1 + (upper - q0), clamped on the host to [0, kScanLead + bn] (DevProblem
// upper_c16, one table per chunk) -- so that a tap outside the chunk's columns reads one of the zero columns that flank the tile
// (two on either side: a pair that lies outside altogether reads two of them) and a pair that straddles a chunk boundary is
// taken tap by tap in both chunks, without a clamp or a select on the device; hrow then points at the row's first lead column.
template <bool MASKED>
__device__ __forceinline__ double y_side_pair(double s, const double* hrow, int u, double w)
{
    if (MASKED) return __dadd_rn(s, __fma_rn(hrow[u], __dadd_rn(1.0, -w), __dmul_rn(hrow[u + 1], w)));
    return __dadd_rn(s, __fma_rn(hrow[u - 1], __dadd_rn(1.0, -w), __dmul_rn(hrow[u], w)));
}

// The y-side tables (16-bit interval indices, FP64 weights) hold the offsets two by two: entry (o, y) at
// ((o >> 1) * G + y) * 2 + (o & 1), so one 4-byte and one 16-byte load fetch a node's entries of two consecutive offsets --
// three load/store instructions per pair and lane instead of four where the phase is bound by the shared-memory queue.
__host__ __device__ inline size_t y_table_index(int o, int G, int y) { return ((size_t)(o >> 1) * G + y) * 2 + (o & 1); }

// One run of consecutive offsets [o, o + n) of the y side in which nothing changes but the position: the H row moves by dh
// doubles per offset (-ld on the minus side, where beta falls as the offset grows, +ld on the plus side, 0 for the single
// offset of m == 2).  Offsets in order, the same accumulation whichever way the entries were fetched.  U_SHARED / W_SHARED:
// the table lives in shared memory (the persistent kernel's copies) or in global memory, so every load knows its space.
template <bool MASKED, bool U_SHARED, bool W_SHARED, int YU>
__device__ __forceinline__ double y_side_run(double s, const double* hrow, int dh, const unsigned short* u_table, const double* w_table,
                                             int G, int y, int o, int n)
{
    __builtin_assume(U_SHARED ? __isShared(u_table) : __isGlobal(u_table));
    __builtin_assume(W_SHARED ? __isShared(w_table) : __isGlobal(w_table));
    if ((o & 1) && n > 0) {                                    // the second entry of a pair of offsets first
        const size_t at = y_table_index(o, G, y);
        s = y_side_pair<MASKED>(s, hrow, (int)u_table[at], w_table[at]);
        hrow += dh; o++; n--;
    }
    const size_t at = y_table_index(o, G, y);                  // o is even
    const unsigned int* up = reinterpret_cast<const unsigned int*>(u_table + at);
    const double2* wp = reinterpret_cast<const double2*>(w_table + at);
    #pragma unroll (YU / 2)
    for (int i = 0; i < (n >> 1); i++) {
        const unsigned int uu = *up;
        const double2 ww = *wp;
        s = y_side_pair<MASKED>(s, hrow, (int)(uu & 0xffffu), ww.x);
        s = y_side_pair<MASKED>(s, hrow + dh, (int)(uu >> 16), ww.y);
        hrow += 2 * dh; up += G; wp += G;
    }
    if (n & 1) {                                               // and the first entry of the last pair
        const size_t last = y_table_index(o + n - 1, G, y);
        s = y_side_pair<MASKED>(s, hrow, (int)u_table[last], w_table[last]);
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
        // MASKED: positions in the tile's row from this chunk's table; otherwise interval indices, the row shifted to column 0
        const double* hrow = Hs + beta * ld + (MASKED ? 0 : kScanLead - q0);
        const unsigned short* u_global = MASKED ? P.upper_c16 + (size_t)(q0 / bn) * ((P.n_off + 1) & ~1) * G : P.upper_p16;
        if (w_shared) s = y_side_run<MASKED, true, true, YU>(s, hrow, dh, u_cache, w_cache, G, y, o, end - o);
        else if (CACHED) s = y_side_run<MASKED, true, false, YU>(s, hrow, dh, u_cache, P.weight_p, G, y, o, end - o);
        else s = y_side_run<MASKED, false, false, YU>(s, hrow, dh, u_global, P.weight_p, G, y, o, end - o);
        o = end;
    }
    return s;
}

// The work of one CTA: row `lx` of this rank, column chunk `chunk`.
// fused (persistent kernel): the per-node sums stay in shared memory (row_sum) for the closing phase of the same CTA, nothing is written to global memory and the WC rows are left to the caller.
// `mid` runs between the two phases (the persistent kernel requests there what its closing phase will need, so that
// those loads fly under the y side).
struct NoMid { __device__ void operator()() const {} };
// `gate(sc)` runs in every thread before the rows of C that super-chunk sc of the E phase reads are requested (the work-list
// kernel waits there, warp by warp, until those rows of the state have arrived: persistent_chunks_kernel); gate.bypass():
// request them past L1 (rows that arrive while the item runs were never covered by an L1 invalidation on this SM).
struct NoGate {
    __device__ void operator()(int) const {}
    __device__ bool bypass() const { return false; }
};

// STAGED_WORK: the chunk's y-side work list is copied to shared memory when the tables are built (kernels that build them once
// per launch); otherwise it is read where it lies.
template <int kScanUnroll, typename Mid = NoMid, typename Gate = NoGate, bool STAGED_WORK = false>
__device__ __forceinline__ void scan_body(const StepArgs& S, const int chunk, const int lx, const bool build_tables,
                                          const int cached_w_rows, const bool fused = false, Mid mid = Mid(), Gate gate = Gate())
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
    int4* sel = reinterpret_cast<int4*>(scan_smem + L.sel);              // [a_pad] rows beta closed by alpha: [x, y), and q[x] (its bits in z, w)
    int4* work_piece = reinterpret_cast<int4*>(scan_smem + L.work);      // [pieces] this chunk's y-side pieces (y block, first offset, end offset)
    int* work_warp_begin = reinterpret_cast<int*>(work_piece + P.scan_max_pieces);   // [warps + 1] first piece of every warp
    int* work_block_begin = work_warp_begin + NT / 32 + 1;               // [n_yblk + 1] first piece of every y block
    unsigned short* u_cache = reinterpret_cast<unsigned short*>(scan_smem + L.end);                    // as DevProblem::upper_p16 (offsets two by two)
    double* w_cache = reinterpret_cast<double*>(scan_smem + align32(L.end + sizeof(unsigned short) * (size_t)((n_off + 1) & ~1) * G));

    SCAN_MARK(0);
    if (build_tables) {
        if (cached_w_rows >= 0 && !STAGED_WORK) {                             // (the one-chunk persistent kernel fills them itself: bulk copies)
            const int n_even = (n_off + 1) & ~1;                              // the tables hold the offsets two by two
            for (int i = tid; i < n_even * G; i += NT) u_cache[i] = P.upper_p16[i];
            for (int i = tid; i < cached_w_rows * G; i += NT) w_cache[i] = P.weight_p[i];   // cached_w_rows is even
        }
        for (int t = tid; t < L.a_pad * 2; t += NT) {
            const int alpha = t >> 1, f = t & 1;
            double w_lo = 0.0, w_hi = 0.0;
            int r_lo = 0, r_hi = 0;
            {
                // the three table words requested together (a load behind `u != 0` would wait for u first: one more round trip
                // to L2 in front of every item of the work-list kernel)
                const bool valid = alpha < A && f < n_fold;
                const size_t at = (size_t)x * n_off + (valid ? folded_offset(A, n_fold, alpha, f) : 0);
                const int u = P.upper[at];
                const double w = P.weight[at], qa = P.qvec[valid ? alpha : 0];
                if (valid && u != 0) {
                    w_lo = qa * (1.0 - w); w_hi = qa * w;
                    r_lo = row_slot(P, u - 1) * G; r_hi = row_slot(P, u) * G;
                }
            }
            tap_w[alpha * 4 + f * 2] = w_lo; tap_w[alpha * 4 + f * 2 + 1] = w_hi;
            tap_row[alpha * 4 + f * 2] = r_lo; tap_row[alpha * 4 + f * 2 + 1] = r_hi;
        }
        if (STAGED_WORK) {                                                    // tables built once per launch (the one-chunk persistent kernel)
            const int first = P.scan_chunk_begin[chunk], n_pieces = P.scan_chunk_begin[chunk + 1] - first, warps = NT / 32;
            for (int i = tid; i < n_pieces; i += NT) work_piece[i] = P.scan_piece[first + i];
            for (int i = tid; i <= warps; i += NT) work_warp_begin[i] = P.scan_warp_begin[chunk * (warps + 1) + i];
            for (int i = tid; i <= L.n_yblk; i += NT) work_block_begin[i] = P.scan_block_begin[chunk * (L.n_yblk + 1) + i];
        }
        for (int a = tid; a < L.a_pad; a += NT) {
            qv[a] = a < A ? P.qvec[a] : 0.0;
            sel[a] = a < A ? P.beta_record[a] : make_int4(0, 0, 0, 0);           // rows [x, y) and q[x]: most alphas close one row
        }
        for (int b = tid; b < A; b += NT) {
            Hs[b * ld + kScanLead - 1] = 0.0;                                 // the zero column (never written again)
            if (P.scan_chunks > 1) {                                          // several chunks: two zero columns on either side (y_side_pair)
                Hs[b * ld] = 0.0;
                Hs[b * ld + kScanLead + bn] = 0.0;                            // its right neighbour is the next row's first lead column
            }
        }
        if (tid == 0 && P.scan_chunks > 1) Hs[A * ld] = 0.0;                 // ... or, behind the last row, this
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
        auto issue_past_l1 = [&](int sc) {
            #pragma unroll
            for (int k = 0; k < kScanUnroll; k++) {
                const int4 ra = *reinterpret_cast<const int4*>(tap_row + (sc * span + grp * kScanUnroll + k) * 4);
                raw[k][0] = __ldcg(c + (unsigned)ra.x); raw[k][1] = __ldcg(c + (unsigned)ra.y); raw[k][2] = __ldcg(c + (unsigned)ra.z); raw[k][3] = __ldcg(c + (unsigned)ra.w);
            }
        };
        const bool past_l1 = gate.bypass();
        if (active) {
            gate(0);
            if (past_l1) issue_past_l1(0); else issue(0);
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
                if (sc + 1 < n_sc) {
                    gate(sc + 1);
                    if (past_l1) issue_past_l1(sc + 1); else issue(sc + 1);
                }
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
                    const int4 rows = sel[sc * span + grp * kScanUnroll + k];       // empty beyond A
                    const double v = live ? start + prefix[k] : 0.0;
                    double* hp = hcol + rows.x * ld;
                    if (rows.x < rows.y) {
                        *hp = v * __hiloint2double(rows.w, rows.z);               // = v * qv[rows.x], without the dependent load
                        hp += ld;
                        #pragma unroll 1                       // one row per alpha on average: no unrolled body, no remainder logic
                        for (int r = rows.x + 1; r < rows.y; r++, hp += ld) *hp = v * qv[r];
                    }
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
        const int warp = tid >> 5, lane = tid & 31;
        const bool masked = P.scan_chunks > 1;
        // (node, offset) pairs in flight per lane: the 256-thread CTAs live within 64 registers, the 512-thread ones can afford more
        constexpr int kYU = kScanUnroll == 2 ? kYUnroll : LDM_Y_UNROLL_WIDE;
        // the work list: from shared memory where the tables are built once per launch (no global round trip in front of every
        // piece); a kernel that rebuilds its tables per item reads it where it lies (measured: staging it per item gains nothing)
        const int4* pieces = STAGED_WORK ? work_piece : P.scan_piece + P.scan_chunk_begin[chunk];
        const int* warp_begin = STAGED_WORK ? work_warp_begin : P.scan_warp_begin + chunk * (NT / 32 + 1);
        const int* block_begin = STAGED_WORK ? work_block_begin : P.scan_block_begin + chunk * (L.n_yblk + 1);
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
// launches, whose members may differ)
template <bool SMALL, int U, typename Mid = NoMid, typename Gate = NoGate, bool STAGED_WORK = false>
__device__ __forceinline__ void scan_body_of(const StepArgs& S, const int chunk, const int lx, const bool build_tables,
                                             const int cached_w_rows, const bool fused = false, Mid mid = Mid(), Gate gate = Gate())
{
    if (SMALL) { scan_body<2, Mid, Gate, STAGED_WORK>(S, chunk, lx, build_tables, cached_w_rows, fused, mid, gate); return; }
    if (U != 0) { scan_body<(U == 0 ? 4 : U), Mid, Gate, STAGED_WORK>(S, chunk, lx, build_tables, cached_w_rows, fused, mid, gate); return; }
    const int ng = (int)blockDim.x / S.P.scan_bn > 1 ? (int)blockDim.x / S.P.scan_bn : 1;
    if (scan_unroll((int)blockDim.x, S.P.A, ng) == 5) scan_body<5, Mid, Gate, STAGED_WORK>(S, chunk, lx, build_tables, cached_w_rows, fused, mid, gate);
    else scan_body<4, Mid, Gate, STAGED_WORK>(S, chunk, lx, build_tables, cached_w_rows, fused, mid, gate);
}

// ---- bulk asynchronous copies into shared memory (TMA, cp.async.bulk): one thread issues, completion is counted in bytes on an
// mbarrier, nobody spends registers or instructions on the bytes.  Used for the resident y-side tables of the one-chunk
// persistent kernel: they land under the kernel's prologue and the first E phase. ----
__device__ __forceinline__ unsigned int shared_address(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_barrier_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(shared_address(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_barrier_expect(unsigned long long* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(shared_address(bar)), "r"(bytes) : "memory");
}
// bytes: a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_copy_to_shared(void* dst, const void* src, unsigned int bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(shared_address(dst)), "l"(src), "r"(bytes), "r"(shared_address(bar)) : "memory");
}
__device__ __forceinline__ void bulk_barrier_wait(unsigned long long* bar, unsigned int parity)
{
    unsigned int done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(shared_address(bar)), "r"(parity) : "memory");
    } while (!done);
}

// ---- hand-overs between CTAs (persistent kernels) ----
// Every wait is bounded: a partner that never shows up (a lost peer, a kernel that died) becomes a sticky code in the
// handle's device status word -- the host turns it into LDM_ERR_CUDA after the stream has drained (ldm_context.h
// check_device_status) -- instead of a hung GPU or garbage handed on as data.
constexpr long long kSpinCycles = 1LL << 34;          // ~8.7 s at 1.97 GHz
enum : int { kStatusOk = 0, kStatusBarrierTimeout = 1, kStatusRingTimeout = 2, kStatusPeerTimeout = 3 };

__device__ __forceinline__ int load_status(const int* status)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(status) : "memory");
    return v;
}

struct Spin {                                         // every 1024 polls: has anybody reported an error, has the time run out
    long long t0 = 0;
    unsigned int polls = 0;
    __device__ __forceinline__ bool give_up(const int* status)
    {
        if ((++polls & 0x3ffu) != 0) return false;
        if (load_status(status) != kStatusOk) return true;
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        return now - t0 > kSpinCycles;
    }
};

// Spin until *counter >= target: relaxed polls (an acquire load in the loop would pay for its fence on every poll), then
// ONE acquire load orders everything that follows behind the writers' releases (measured: cheaper than a fence here).
// false: gave up (status set).
__device__ __forceinline__ bool wait_for_count(const unsigned int* counter, unsigned int target, int* status)
{
    unsigned int seen;
    Spin spin;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (seen >= target) break;
        if (spin.give_up(status)) { atomicCAS(status, kStatusOk, kStatusBarrierTimeout); return false; }
    }
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    return true;
}

// Hand-over of one ring value between CTAs without a fence (see persistent_scan_kernel): the slot holds kRingEmpty until
// the value is published.  A computed NaN never has this bit pattern, so a state that has gone NaN does not stall the poll.
constexpr unsigned long long kRingEmpty = ~0ULL;

__device__ __forceinline__ void publish_ring(double* slot, double value)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(slot), "d"(value) : "memory");
}

__device__ __forceinline__ double take_ring(double* slot, int* status)
{
    unsigned long long bits;
    Spin spin;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(slot) : "memory");
        if (bits != kRingEmpty) break;
        if (spin.give_up(status)) { atomicCAS(status, kStatusOk, kStatusRingTimeout); return 0.0; }
    }
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(kRingEmpty) : "memory");
    return __longlong_as_double((long long)bits);
}

// Grid-wide barrier of a persistent kernel: a monotone arrival counter in global memory (zeroed before the launch; the
// cooperative launch guarantees that every CTA is resident).  After the CTA barrier thread 0 arrives with a releasing
// reduction (cumulative over the CTA's writes, and nothing waits for a returned value) and polls until all CTAs of
// barrier number `index` have arrived.  Returns false in every thread of the CTA once the status word is set -- by this
// wait running out of time or by anybody else -- and the caller leaves its step loop.
__device__ __forceinline__ bool grid_barrier(unsigned int* counter, unsigned int index, int* status)
{
    __shared__ int barrier_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        const bool arrived = wait_for_count(counter, (index + 1u) * gridDim.x, status);
        barrier_ok = arrived && load_status(status) == kStatusOk;
    }
    __syncthreads();
    return barrier_ok != 0;
}

}  // namespace

}  // namespace ldm
