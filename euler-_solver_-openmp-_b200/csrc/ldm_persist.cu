// Persistent multi-step kernels of the prefix-sum path for sm_100a: all Euler steps of a solve (EulerSolver::Solve,
// Solver.cpp:557-592) inside ONE cooperative launch.
//
//   persistent_scan_kernel     grids whose rows are one chunk each (G <= 128), one rank: one CTA per moment-grid row, one
//                              grid barrier per step, tables resident in shared memory across steps;
//   persistent_chunks_kernel   everything else on the prefix-sum path -- rows of several chunks (G > 128) and the xi-sharded
//                              step of SURVEY 8e on 1..8 ranks: one CTA per SM pulls (row, chunk) items from a list ordered by
//                              cost, the CTA that completes a row closes it and stores it straight into every rank's next
//                              replica over NVLink, and a step ends with per-rank epoch flags instead of a barrier.
//
// Both run scan_body (ldm_scan_body.cuh) unchanged and close a node with the arithmetic of finish_factored_kernel
// (ldm_kernels.cu), so every path of a handle -- per-step launches, persistent, sharded -- produces the same bits.
#include "../../include/ldm.h"
#include "ldm_scan_body.cuh"

namespace ldm {

namespace {

// Optional event trace of the work-list kernel (debug build `make phase-clocks`; tools/chunk_trace.py): thread 0 of every CTA
// logs (globaltimer, event, step, local row, chunk).  Off in the product build.
#ifdef LDM_PHASE_CLOCKS
constexpr int kTraceCtas = 148, kTraceEvents = 4096;
__device__ ulonglong2 g_trace[kTraceCtas][kTraceEvents];
__device__ unsigned int g_trace_n[kTraceCtas];
__device__ __forceinline__ void trace_event(int type, int k, int lx, int chunk)
{
    if (threadIdx.x != 0 || blockIdx.x >= kTraceCtas) return;
    const unsigned int n = g_trace_n[blockIdx.x];
    if (n >= kTraceEvents) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[blockIdx.x][n] = make_ulonglong2(t, ((unsigned long long)type << 48) | ((unsigned long long)(k & 0xffff) << 32) | ((unsigned long long)(lx & 0xffff) << 16) | (unsigned long long)(chunk & 0xffff));
    g_trace_n[blockIdx.x] = n + 1;
}
#define TRACE(type, k, lx, chunk) trace_event(type, k, lx, chunk)
#else
#define TRACE(type, k, lx, chunk) do { } while (0)
#endif

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
    unsigned int barrier_index = A.barriers_done;      // the arrival counter runs on from launch to launch
    // The resident y-side tables -- all 16-bit interval indices, the first cached_w_rows rows of weights (behind the step's
    // working set: scan_body) -- by two bulk asynchronous copies (TMA): requested now, waited for in front of the first y side.
    __shared__ __align__(8) unsigned long long tables_landed;
    const bool resident_tables = A.cached_w_rows >= 0;
    if (resident_tables && tid == 0) {
        const size_t n_even = (size_t)((P.n_off + 1) & ~1);
        unsigned char* u_cache = scan_smem + L.end;
        unsigned char* w_cache = scan_smem + align32(L.end + sizeof(unsigned short) * n_even * G);
        const unsigned int u_bytes = (unsigned int)((sizeof(unsigned short) * n_even * G + 15) & ~(size_t)15);   // (both ends are padded)
        const unsigned int w_bytes = (unsigned int)(sizeof(double) * (size_t)A.cached_w_rows * G);              // an even number of rows
        bulk_barrier_init(&tables_landed);
        bulk_barrier_expect(&tables_landed, u_bytes + w_bytes);
        bulk_copy_to_shared(u_cache, P.upper_p16, u_bytes, &tables_landed);
        if (w_bytes) bulk_copy_to_shared(w_cache, P.weight_p, w_bytes, &tables_landed);
    }

    // prologue: the row dots of WC for the initial state
    if (tid < 32) {
        const double s = wc_row_dot(P.wc_operator + (size_t)x * G, A.C[parity] + (size_t)x * G, G, tid);
        if (tid == 0) A.wc_rows[parity * G + x] = s;
    }
    if (!grid_barrier(A.barrier, barrier_index++, A.status)) return;

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

#ifdef LDM_PHASE_CLOCKS
    if (tid == 0) {                                  // which SM, and the shape of this CTA's ring share: CSR rows << 32 | entries
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_scan_clock[blockIdx.x % (148 * 8)][7] = smid;
        g_scan_clock[blockIdx.x % (148 * 8)][2] = ((long long)(share_r1 - share_r0) << 32) | (long long)(P.ring_begin[share_r1] - share_e0);
    }
#endif
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
            if (k == 0 && resident_tables) bulk_barrier_wait(&tables_landed, 0);     // issued at the kernel's first instruction: long there
            if (closer) {
                if (has_ring) ring_v = take_ring(ring_slot + q, A.status);
                c_old = S.C[q];
            }
            N = S.N_cur[0];
            if (warp == last_warp) {
                #pragma unroll
                for (int j = 0; j < 4; j++) wc_part[j] = lane + 32 * j < G ? wc_rows_now[lane + 32 * j] : 0.0;
            }
        };
        scan_body_of<SMALL, U, decltype(mid), NoGate, true>(S, 0, x, k == 0, A.cached_w_rows, true, mid);   // ends with a CTA barrier
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
        const bool rhs_only = A.last_is_rhs_only && k == A.n_steps - 1;      // the right-hand side at the last state: nothing is advanced
        if (x == 0 && tid == 0) {
            first_moment_update(P, N, WC, S.dt, S.scalars, rhs_only ? nullptr : S.N_next);
            if (values_next && !rhs_only) values_next[0] = S.N_next[0];
            if (derivs_here) derivs_here[0] = S.scalars[1];
        }
        if (closer) {
            const double inside = (0.0 + row_sum[tid]) - ring_v;
            const double mc = __dmul_rn(add_outside(P, inside, N, S.C, S.scalars, ow, 0.0), P.hh);
            const double dc = second_moment_rhs(P, mc, c_old, N, WC, mn, wn);
            double v = __dadd_rn(c_old, __dmul_rn(dc, S.dt));                      // Solver.cpp:577
            v = v < 0.0 ? 0.0 : v;                                                 // Solver.cpp:578-580
            if (!rhs_only) {
                S.C_next[q] = v;
                c_new[tid] = v;
                if (values_next) values_next[1 + q] = v;
            }
            if (derivs_here) derivs_here[1 + q] = dc;
        }
        if (rhs_only) break;                                                       // (uniform: the launch ends here)
        __syncthreads();
        if (warp == 0) {     // this row's dot of the WC functional for the next state
            const double s = wc_row_dot(op_row, c_new, G, lane);
            if (lane == 0) A.wc_rows[(parity ^ 1) * G + x] = s;
        }
        SCAN_MARK(8);
        if (!grid_barrier(A.barrier, barrier_index++, A.status)) return;     // a lost partner: stop stepping (the host reports the status word)
        SCAN_MARK(9);
        parity ^= 1;
    }
}

// Spin until every rank has published `need` steps in this rank's flag array (lane r of one warp polls rank r; the
// flags are written by the peers over NVLink with a system-scope release).  Relaxed polls of local memory, then one
// system-scope acquire per lane; the CTA barrier that follows hands the ordering on to the other threads.
__device__ __forceinline__ bool wait_for_ranks(const unsigned long long* flags, int world, unsigned long long need, int* status)
{
    const int r = threadIdx.x;
    bool ok = true;
    if (r < world) {
        unsigned long long seen;
        Spin spin;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + r) : "memory");
            if (seen >= need) break;
            if (spin.give_up(status)) { atomicCAS(status, kStatusOk, kStatusPeerTimeout); ok = false; break; }
        }
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + r) : "memory");
    }
    return ok;
}

// Non-blocking form of wait_for_ranks (one warp): true when this rank's flag array already shows `need` steps for every rank.
__device__ __forceinline__ bool ranks_have_published(const unsigned long long* flags, int world, unsigned long long need)
{
    const int r = threadIdx.x;
    bool here = true;
    if (r < world) {
        unsigned long long seen;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + r) : "memory");
        here = seen >= need;
    }
    const bool all = __all_sync(0xffffffffu, here);
    if (all && r < world) {
        unsigned long long again;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(again) : "l"(flags + r) : "memory");
        here = again >= need;
    }
    return __all_sync(0xffffffffu, here);
}

// Spin (one warp) until rows lo..hi of state `need` have arrived in this rank's replica: the closing code of a row stores
// the state's number into the row's flag on every rank behind a system fence.  Relaxed polls of local memory, one
// acquiring fence per lane at the end; the CTA barrier that follows hands the ordering on to the other threads.
__device__ __forceinline__ bool wait_for_rows(const unsigned long long* row_flags, int lo, int hi, unsigned long long need, int* status)
{
    bool ok = true;
    Spin spin;
    for (int r = lo + (int)threadIdx.x; r <= hi && ok; r += 32) {
        unsigned long long seen;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(row_flags + r) : "memory");
            if (seen >= need) break;
            if (spin.give_up(status)) { atomicCAS(status, kStatusOk, kStatusPeerTimeout); ok = false; break; }
        }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return !__any_sync(0xffffffffu, !ok);
}

// The gate of an item that starts before its state is complete (scan_body's Gate): before a thread group requests the rows of
// C that its next block of alphas reads, its warps wait -- each for itself, lanes striding the rows -- until those rows have
// arrived.  The item's E phase then follows the closings of the step before it row by row instead of waiting for the last
// row of its window, and the rows are requested past L1 (no L1 invalidation on this SM lies between their arrival and the
// loads; the relaxed poll of a flag that the writer stored behind its system fence orders the L2 reads that follow it).
struct RowGate {
    const unsigned long long* row_flags;   // this rank's
    const int4* blocks;                    // this row's windows, [super-chunk * groups + group]: rows lo..hi of the minus side, of the plus side
    unsigned long long need;
    int* status;
    int groups, bn, super_chunks;
    bool active;
    // the flags of the next block are requested one super-chunk ahead, so that a block whose rows are already there costs
    // no round trip to L2 (lane l holds the flags of row lo + l on either side)
    unsigned long long ahead_minus = 0, ahead_plus = 0;
    int ahead_of = -1;

    __device__ bool bypass() const { return active; }
    __device__ static unsigned long long peek(const unsigned long long* flag)
    {
        unsigned long long seen;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
        return seen;
    }
    __device__ void request(int sc)
    {
        const int lane = threadIdx.x & 31;
        const int4 w = blocks[sc * groups + (int)threadIdx.x / bn];
        ahead_minus = ~0ULL; ahead_plus = ~0ULL;
        if (w.x + lane <= w.y) ahead_minus = peek(row_flags + w.x + lane);
        if (w.z + lane <= w.w) ahead_plus = peek(row_flags + w.z + lane);
        ahead_of = sc;
    }
    __device__ void operator()(int sc)
    {
        if (!active) return;
        const int lane = threadIdx.x & 31;
        if (ahead_of != sc) request(sc);
        const int4 w = blocks[sc * groups + (int)threadIdx.x / bn];
        Spin spin;
        bool ok = true;
        #pragma unroll 1
        for (int side = 0; side < 2 && ok; side++) {
            const int lo = side == 0 ? w.x : w.z, hi = side == 0 ? w.y : w.w;
            unsigned long long seen = side == 0 ? ahead_minus : ahead_plus;    // of row lo + lane (all ones: none)
            for (int r = lo + lane; r <= hi && ok; r += 32) {
                while (seen < need) {
                    if (spin.give_up(status)) { atomicCAS(status, kStatusOk, kStatusPeerTimeout); ok = false; break; }
                    seen = peek(row_flags + r);
                }
                if (r + 32 <= hi) seen = peek(row_flags + r + 32);
            }
        }
        __syncwarp();
        if (sc + 1 < super_chunks) request(sc + 1);
    }
};

// Persistent work-list kernel (cooperative launch, one CTA per SM or fewer): see ChunkedArgs in ldm_device.h.
// The CTAs pull (step, local row, chunk) items off ONE list that runs through all steps of the launch -- rows in ascending
// order, the chunks of a row by descending cost, one atomic per item, fetched under the item before -- and nothing but data
// dependencies orders the steps:
//   * an item of step k (state e = epoch + k -> e + 1) starts when the rows of state e that its E phase reads -- the window
//     of moment-grid rows around x that the kernel's support reaches (ChunkedArgs::row_window) -- have arrived in this rank's
//     replica (per-row flags), so the first rows of step k + 1 are under way while the last rows of step k are still being
//     computed, closed and shipped, here and on the other ranks;
//   * the chunk sums go to global memory, the row's arrival counter is bumped, and the CTA that brings the last chunk of a
//     row CLOSES it.  A closing reads the whole state (WC, the ring correction), so it waits until every rank has
//     published state e; then chunk sums in chunk order, ring correction, closure, Euler update, clamp -- the
//     arithmetic of finish_factored_kernel -- and the new row and its dot with the WC functional are stored into the next
//     replica of every rank (peer stores over NVLink: the all-gather of SURVEY 8e, row by row), one system fence, the row's
//     flag on every rank; whoever closes the rank's last row of a step publishes the step in every rank's flag array;
//   * the CTA's share of the sparse ring correction of state e (CSR rows dealt out by entry count; handed over without
//     fences through ping-pong slots holding an all-ones NaN until written, as in persistent_scan_kernel) is computed at
//     the first item boundary at which state e is complete, and at the latest before the CTA closes a row of step k or
//     turns to an item of step k + 1 -- so every share of step k is out before anybody can wait for anything of state
//     e + 1, and no wait can close a cycle.
// A buffer is only ever rewritten by closings, a closing of step k + 1 needs state e + 1 complete, and every read of state
// e (E phases, ring shares, closings of step k) precedes the flag that completes state e + 1: the two replicas, the one
// set of chunk sums (row x lies in its own window) and the two halves of the ring hand-over suffice.
template <bool SMALL, int U>
__global__ void __launch_bounds__(SMALL ? 256 : kScanThreadsMax, SMALL ? 4 : 1)
persistent_chunks_kernel(const ChunkedArgs A)
{
    const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, warps = NT / 32;
    __shared__ double wc_shared;
    __shared__ unsigned int next_item[2];
    __shared__ int4 next_record[2];
    __shared__ int gate;
    StepArgs S = A.step;
    const DevProblem& P = S.P;
    const int G = P.G, n_local = local_rows(P), chunks = P.scan_chunks, n = n_local * G;
    const ScanLayout L = scan_layout(G, P.A, P.scan_bn, NT, P.scan_max_pieces);
    extern __shared__ __align__(32) unsigned char scan_smem[];
    double* c_new = reinterpret_cast<double*>(scan_smem + L.rowbuf);              // [G] the row being closed, at the new time
    unsigned int* fetch_counter = A.counters;
    unsigned int* rows_closed = A.counters + 1;
    unsigned int* row_arrivals = A.counters + 3;
    const unsigned long long* own_flags = A.flags[A.rank];                         // [kMaxPeers] steps published by every rank ...
    const unsigned long long* own_row_flags = own_flags + kMaxPeers;               // ... then [G] newest state of every row that has arrived here

    // prologue: the state this launch starts from must be complete in this rank's replica; then the row dots of the WC
    // functional for it, all G rows on every rank (no exchange), dealt out over the warps of the grid
    if (tid < 32) {
        const bool ok = wait_for_ranks(own_flags, A.world, A.epoch, A.status);
        if (__any_sync(0xffffffffu, !ok) && tid == 0) atomicCAS(A.status, kStatusOk, kStatusPeerTimeout);
    }
    __syncthreads();
    {
        const double* C0 = A.replica[A.parity][A.rank];
        double* tail = A.replica[A.parity][A.rank] + A.words;
        for (int xr = blockIdx.x * warps + warp; xr < G; xr += gridDim.x * warps) {
            const double s = wc_row_dot(P.wc_operator + (size_t)xr * G, C0 + (size_t)row_slot(P, xr) * G, G, lane);
            if (lane == 0) tail[xr] = s;
        }
    }
    if (tid == 0) {                                                                // this CTA's first item
        next_item[0] = atomicAdd(fetch_counter, 1u);
        next_record[0] = A.items[next_item[0] % (unsigned int)A.n_items];
    }
    if (!grid_barrier(A.counters + 2, 0, A.status)) return;
    unsigned int item = next_item[0];
    unsigned int turn = 0;                                                         // items this CTA has taken so far
    const unsigned int total = (unsigned int)A.n_steps * (unsigned int)A.n_items;
    int k = 0;                                                                     // the step `item` belongs to
    unsigned int step_begin = 0;                                                   // its first item
    int complete = 0;            // states epoch .. epoch + complete are known to be complete in this rank's replica
    int ring_done = -1;          // newest step whose ring share this CTA has handed over

    // this CTA's share of the ring correction of step kk (state epoch + kk must be complete here)
    auto ring_share = [&](int kk) {
        const double* C_of = A.replica[A.parity ^ (kk & 1)][A.rank];
        double* slot = A.ring_pingpong + (size_t)(kk & 1) * n;
        for (int row = A.ring_cta_begin[blockIdx.x] + warp; row < A.ring_cta_begin[blockIdx.x + 1]; row += warps) {
            const double acc = ring_row_dot(P, C_of, row, lane);
            if (lane == 0) publish_ring(slot + P.ring_point[row], acc);
        }
        ring_done = kk;
        TRACE(7, kk, 0, 0);
    };
    // all threads: wait until every rank has published state epoch + kk; false: a lost partner (status word set)
    auto await_state = [&](int kk) -> bool {
        if (complete >= kk) return true;
        if (tid < 32) {
            const bool ok = wait_for_ranks(own_flags, A.world, A.epoch + (unsigned long long)kk, A.status);
            const bool all = !__any_sync(0xffffffffu, !ok);
            if (tid == 0) gate = all && load_status(A.status) == kStatusOk;
        }
        __syncthreads();
        const bool ok = gate != 0;
        __syncthreads();                                                           // `gate` is rewritten by the next wait
        if (ok) complete = kk;
        return ok;
    };

    while (item < total) {
        while (item >= step_begin + (unsigned int)A.n_items) { k++; step_begin += (unsigned int)A.n_items; }
        SCAN_MARK(10);
        TRACE(1, k, 0, 0);
        // shares of earlier steps that this CTA still owes (it took no item boundary at which their state was complete)
        while (ring_done < k - 1) {
            if (!await_state(ring_done + 1)) return;
            ring_share(ring_done + 1);
        }
        const int4 it = next_record[turn & 1];                                     // (local row, chunk, first and last row of the row's window)
        const int lx = it.x;
        const int x = P.rank + P.world * lx;
        bool gated = false;
        if (complete < k) {
            // Is state epoch + k complete, or at least every row of the item's window here?  (One round of independent polls.)
            // Then the item runs like any other.  Otherwise it goes ahead behind the gate (RowGate) once the step before has
            // closed this very row -- its chunk sums and arrival counter are reused step after step (a shape whose block
            // windows do not fit the row buffers waits for the whole window instead).
            if (tid < 32) {
                const unsigned long long need = A.epoch + (unsigned long long)k;
                bool published = true, here = true;
                if (tid < A.world) published = RowGate::peek(own_flags + tid) >= need;
                for (int r = it.z + tid; r <= it.w; r += 32) here = here && RowGate::peek(own_row_flags + r) >= need;
                published = __all_sync(0xffffffffu, published);
                here = __all_sync(0xffffffffu, here);
                const bool can_gate = A.blocks_per_row <= L.part_stride;
                int g;
                if (published || here) {
                    asm volatile("fence.acq_rel.sys;" ::: "memory");
                    g = published ? 2 : 3;
                } else if (can_gate) {
                    g = wait_for_rows(own_row_flags, x, x, need, A.status) ? 1 : 0;
                } else {
                    g = wait_for_rows(own_row_flags, it.z, it.w, need, A.status) ? 3 : 0;
                }
                if (load_status(A.status) != kStatusOk) g = 0;
                if (tid == 0) gate = g;
            } else if (A.blocks_per_row <= L.part_stride) {
                // meanwhile: the row's block windows into shared memory (the closing's row buffers are idle while an item runs)
                int4* const windows = reinterpret_cast<int4*>(scan_smem + L.rowbuf);
                for (int b = tid - 32; b < A.blocks_per_row; b += NT - 32) windows[b] = A.block_window[(size_t)x * A.blocks_per_row + b];
            }
            __syncthreads();
            const int g = gate;
            __syncthreads();
            if (g == 0) return;                                                    // a lost partner: the host reports the status word
            if (g == 2) complete = k;
            gated = g == 1;
        }
        if (complete >= k && ring_done < k) ring_share(k);
        // the same check between the two phases of the item and behind it: a share that comes an item late keeps the
        // closings of this step waiting (take_ring), and with them everything that depends on the rows they publish
        auto ring_if_complete = [&]() {
            if (ring_done >= k) return;
            if (complete < k) {
                if (tid < 32) {
                    const bool all = ranks_have_published(own_flags, A.world, A.epoch + (unsigned long long)k);
                    if (tid == 0) gate = all ? 1 : 0;
                }
                __syncthreads();
                const int g = gate;
                __syncthreads();
                if (g == 0) return;
                complete = k;
            }
            ring_share(k);
        };
        const int parity = A.parity ^ (k & 1);
        double* const C_cur = A.replica[parity][A.rank];
        S.C = C_cur;
        S.N_cur = A.N + parity;
        S.N_next = A.N + (parity ^ 1);
        SCAN_MARK(11);
        TRACE(2, k, lx, it.y);

        if (tid == 0) {                                                              // the next item, fetched under this one's work
            const unsigned int next = atomicAdd(fetch_counter, 1u);
            next_item[(turn + 1) & 1] = next;
            next_record[(turn + 1) & 1] = A.items[next % (unsigned int)A.n_items];
        }
        const int4* const windows = reinterpret_cast<const int4*>(scan_smem + L.rowbuf);
        const RowGate row_gate{own_row_flags, windows, A.epoch + (unsigned long long)k, A.status, L.ng, P.scan_bn, A.blocks_per_row / L.ng, gated};
        scan_body_of<SMALL, U>(S, it.y, lx, true, -1, false, ring_if_complete, row_gate);   // chunk sums -> S.partial; ends with a CTA barrier
        ring_if_complete();
        TRACE(3, k, lx, it.y);
        // Arrival: ONE read-modify-write that releases this CTA's chunk sums and acquires the others' (behind the CTA barrier at
        // the end of scan_body, in front of the one below); the CTA that brings a row's last chunk closes the row.  (A closer
        // fixed in advance -- the row's last chunk in the list -- was measured: it waits for chunks that are themselves held
        // up behind gates, and the convoy costs a fifth of the step.)
        if (tid == 0) {
            unsigned int before;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(before) : "l"(row_arrivals + lx), "r"(1u) : "memory");
            gate = before + 1u == (unsigned int)(k + 1) * (unsigned int)chunks;
        }
        __syncthreads();
        const bool closes_row = gate != 0;
        __syncthreads();                                 // `gate` is rewritten by the next wait
        turn++;
        item = next_item[turn & 1];
        if (!closes_row) continue;

        // ---- close local row lx: everything finish_factored_kernel does for it, same arithmetic ----
        // What does not need the complete state comes first, so that it flies while the state completes: the ring correction
        // of a light row (its samples read rows of the row's own window, which the E phases have waited for) and the
        // operands of the closure.  Behind the wait there is one round of loads (WC row dots, N) between the CTA and its stores.
        TRACE(4, k, lx, 0);
        double* const scratch = reinterpret_cast<double*>(scan_smem + L.hs);         // the H tile is free between items
        double* const ring_buf = c_new + L.part_stride;                              // [G] ring correction of a light row
        const bool self_ring = lx < A.self_ring_rows;
        const int cr0 = P.ring_row_begin[lx], cr1 = P.ring_row_begin[lx + 1];
        const int e0 = P.ring_begin[cr0], n_entries = self_ring ? P.ring_begin[cr1] - e0 : 0;
        const bool staged = A.self_ring_capacity >= 0;                                // the tile has room for the staging (tiny tables: no)
        const int capacity = staged ? A.self_ring_capacity : 0;
        double* const gathered = scratch;                                             // [n_entries] C at the entries' nodes
        double* const weights = scratch + capacity;                                   // [n_entries]
        double* const op_stage = scratch + 2 * capacity;                              // [G] this row of the WC functional
        const double* const op_row = staged ? op_stage : P.wc_operator + (size_t)x * G;
        {
            constexpr int kStage = 8;                                                 // entries per thread in flight (capacity 4096 = 8 x 512)
            for (int first = 0; first < n_entries; first += NT * kStage) {
                int node[kStage];
                double value[kStage], weight[kStage];
                #pragma unroll
                for (int j = 0; j < kStage; j++) {
                    const int e = first + tid + j * NT;
                    node[j] = e < n_entries ? P.ring_node[e0 + e] : -1;
                    weight[j] = e < n_entries ? P.ring_weight[e0 + e] : 0.0;
                }
                #pragma unroll
                for (int j = 0; j < kStage; j++) value[j] = node[j] >= 0 ? C_cur[node[j]] : 0.0;
                #pragma unroll
                for (int j = 0; j < kStage; j++) {
                    const int e = first + tid + j * NT;
                    if (e < n_entries) { gathered[e] = value[j]; weights[e] = weight[j]; }
                }
            }
        }
        for (int y = tid; y < G; y += NT) {
            if (staged) op_stage[y] = P.wc_operator[(size_t)x * G + y];
            if (self_ring) ring_buf[y] = 0.0;
        }
        constexpr int kPre = 2;                                                       // nodes per thread whose operands are requested ahead
        const bool ahead = G <= kPre * NT;
        double pre_sum[kPre], pre_c[kPre], pre_ow[kPre], pre_mn[kPre], pre_wn[kPre];
        #pragma unroll
        for (int j = 0; j < kPre; j++) {
            const int y = tid + j * NT;
            pre_sum[j] = 0.0; pre_c[j] = 0.0; pre_ow[j] = 0.0; pre_mn[j] = 0.0; pre_wn[j] = 0.0;
            if (ahead && y < G) {
                #pragma unroll 8
                for (int t = 0; t < chunks; t++) pre_sum[j] += __ldcg(S.partial + (size_t)t * n + lx * G + y);
                pre_c[j] = C_cur[(size_t)(P.rank * P.rows_per_rank + lx) * G + y];
                pre_ow[j] = P.out_weight[(size_t)x * G + y]; pre_mn[j] = P.m_nodes[(size_t)x * G + y]; pre_wn[j] = P.w_nodes[(size_t)x * G + y];
            }
        }
        __syncthreads();
        if (self_ring) {
            // the dots in the order of ring_row_dot / ring_kernel (lanes stride the entries, fma in entry order, butterfly)
            for (int row = cr0 + warp; row < cr1; row += warps) {
                const int b = P.ring_begin[row] - e0, e = P.ring_begin[row + 1] - e0;
                double acc = 0.0;
                for (int en = b + lane; en < e; en += 32) acc = fma(weights[en], gathered[en], acc);
                acc = warp_sum(acc);
                if (lane == 0) ring_buf[P.ring_point[row] - lx * G] = acc;
            }
        }
        if (!await_state(k)) return;
        if (ring_done < k) ring_share(k);
        TRACE(5, k, lx, 0);
        const double* wc_tail = C_cur + A.words;
        const long long point = A.first_point + k;
        double* values_next = A.values ? A.values + (point + 1) * A.stride : nullptr;
        double* derivs_here = A.derivs ? A.derivs + point * A.stride : nullptr;
        double* ring_slot = A.ring_pingpong + (size_t)(k & 1) * n;
        if (warp == 0) {
            const double t = wc_total(wc_tail, G, lane);
            if (lane == 0) wc_shared = t;
        }
        __syncthreads();                                                             // also: ring_buf is complete
        const double N = S.N_cur[0];
        const double WC = __dmul_rn(add_outside(P, wc_shared, N, C_cur, S.scalars, P.wc_outside, 0.0), P.hh);   // Solver.cpp:311, 314
        if (lx == 0 && tid == 0) {
            first_moment_update(P, N, WC, S.dt, S.scalars, S.N_next);
            if (values_next) values_next[0] = S.N_next[0];
            if (derivs_here) derivs_here[0] = S.scalars[1];
            if (A.observables && x == 0) {
                A.observables[2 * point] = N;                                       // Values[k][0]       (Solver.cpp:1011)
                A.observables[2 * point + 1] = C_cur[G - 1];                        // Values[k][G]       (Solver.cpp:1014); row 0 is slot 0
            }
        }
        auto close_node = [&](int y, double inside, double c, double ow, double mn, double wn) {
            const int ql = lx * G + y;
            const int q = x * G + y;
            const double ring_v = self_ring ? ring_buf[y] : (P.ring_present[ql] != 0 ? take_ring(ring_slot + ql, A.status) : 0.0);
            inside -= ring_v;
            const size_t slot = (size_t)(P.rank * P.rows_per_rank + lx) * G + y;   // this node in the gathered layout
            const double mc = __dmul_rn(add_outside(P, inside, N, C_cur, S.scalars, ow, 0.0), P.hh);
            const double dc = second_moment_rhs(P, mc, c, N, WC, mn, wn);
            double v = __dadd_rn(c, __dmul_rn(dc, S.dt));                           // Solver.cpp:577
            v = v < 0.0 ? 0.0 : v;                                                  // Solver.cpp:578-580
            for (int r = 0; r < A.world; r++) A.replica[parity ^ 1][r][slot] = v;   // the all-gather: NVLink peer stores
            c_new[y] = v;
            if (values_next) values_next[1 + q] = v;
            if (derivs_here) derivs_here[1 + q] = dc;
        };
        if (ahead) {
            #pragma unroll
            for (int j = 0; j < kPre; j++) {
                const int y = tid + j * NT;
                if (y < G) close_node(y, pre_sum[j], pre_c[j], pre_ow[j], pre_mn[j], pre_wn[j]);
            }
        } else {
            for (int y = tid; y < G; y += NT) {
                double inside = 0.0;
                for (int t = 0; t < chunks; t++) inside += __ldcg(S.partial + (size_t)t * n + lx * G + y);
                close_node(y, inside, C_cur[(size_t)(P.rank * P.rows_per_rank + lx) * G + y], P.out_weight[(size_t)x * G + y],
                           P.m_nodes[(size_t)x * G + y], P.w_nodes[(size_t)x * G + y]);
            }
        }
        __syncthreads();
        if (warp == 0) {     // this row's dot with the WC functional for the next state, to every rank's tail
            const double s = wc_row_dot(op_row, c_new, G, lane);
            if (lane == 0) {
                for (int r = 0; r < A.world; r++) A.replica[parity ^ 1][r][A.words + x] = s;
            }
        }
        __syncthreads();
        if (tid == 0) {
            // release: ONE system fence per closing CTA, behind the CTA barrier (cumulative over what the barrier ordered
            // before it), then the row's flag on every rank; the CTA that closes the rank's last row of the step publishes
            // the step in every rank's flag array
            // (a system fence takes ~4 us on this machine -- measured with tools/chunk_trace.py: eight releasing stores, each
            // with its own fence, kept every step's completion waiting for 30 us -- so the flags are relaxed stores behind ONE
            // fence, and one rank alone fences at device scope)
            const unsigned long long state = A.epoch + (unsigned long long)k + 1ULL;
            if (A.world == 1) __threadfence(); else __threadfence_system();
            for (int r = 0; r < A.world; r++)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(A.flags[r] + kMaxPeers + x), "l"(state) : "memory");
            const unsigned int closed = atomicAdd(rows_closed, 1u) + 1u;
            if (closed == (unsigned int)(k + 1) * (unsigned int)n_local) {
                if (A.world == 1) __threadfence(); else __threadfence_system();   // behind the other closers' fences and arrivals
                for (int r = 0; r < A.world; r++)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(A.flags[r] + A.rank), "l"(state) : "memory");
            }
        }
        SCAN_MARK(8);
        TRACE(6, k, lx, 0);
    }
    // the shares this CTA still owes: the closings of the last steps wait for them
    while (ring_done < A.n_steps - 1) {
        if (!await_state(ring_done + 1)) return;
        ring_share(ring_done + 1);
    }
}


// Barrier among the T CTAs of one team (see grid_barrier): a monotone arrival counter per team.  T == 1: the CTA barrier alone.
__device__ __forceinline__ bool team_barrier(unsigned int* counter, unsigned int index, int T, int* status)
{
    __shared__ int team_ok;
    __syncthreads();
    if (T == 1) return true;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        const bool arrived = wait_for_count(counter, (index + 1u) * (unsigned int)T, status);
        team_ok = arrived && load_status(status) == kStatusOk;
    }
    __syncthreads();
    return team_ok != 0;
}

// Persistent batch kernel (cooperative launch): see BatchArgs in ldm_device.h.  The sets' rows are one chunk each (G <= 128).
// Per step a CTA sums the G row dots of the WC functional of its set (written by the team's closing phases of the step
// before), then for each of its rows runs scan_body (per-node sums stay in shared memory), applies the ring correction
// of that row's nodes itself (ring_row_dot: the bits of ring_kernel), closes the row -- closure, Euler update, clamp, the
// arithmetic of finish_factored_kernel -- and forms the row's dot with the WC functional for the next state.  One team
// barrier per step; a team of one needs none.
template <bool SMALL>
__global__ void __launch_bounds__(SMALL ? 256 : kScanThreadsMax, SMALL ? 4 : 1)
persistent_batch_kernel(const BatchArgs A)
{
    const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, warps = NT / 32;
    __shared__ double wc_shared;
    const int4 me = A.cta[blockIdx.x];
    const int set = me.x, r = me.y, T = me.z;
    const StepArgs* const args[2] = {A.args[0] + set, A.args[1] + set};
    const DevProblem& P = args[0]->P;                                   // the same problem in both
    const int G = P.G;
    const ScanLayout L = scan_layout(G, P.A, P.scan_bn, NT, P.scan_max_pieces);
    extern __shared__ __align__(32) unsigned char scan_smem[];
    double* row_sum = reinterpret_cast<double*>(scan_smem + L.rowsum);  // [G] after scan_body(fused)
    double* c_new = reinterpret_cast<double*>(scan_smem + L.rowbuf);    // [G] the row being closed, at the new time
    double* ring_buf = c_new + L.part_stride;                           // [G] its ring correction
    double* wc_rows = A.wc_rows[set];
    unsigned int* barrier = A.barriers + set;
    int parity = A.parity;
    unsigned int barrier_index = 0;

    // prologue: the row dots of WC for the initial state, this CTA's rows
    if (warp == 0) {
        for (int lx = r; lx < G; lx += T) {
            const double s = wc_row_dot(P.wc_operator + (size_t)lx * G, args[parity]->C + (size_t)lx * G, G, lane);
            if (lane == 0) wc_rows[parity * G + lx] = s;
        }
    }
    if (!team_barrier(barrier, barrier_index++, T, A.status)) return;

    for (int k = 0; k < A.n_steps; k++) {
        const StepArgs& S = *args[parity];
        if (warp == 0) {
            const double t = wc_total(wc_rows + parity * G, G, lane);
            if (lane == 0) wc_shared = t;
        }
        __syncthreads();
        const double N = S.N_cur[0];
        const double WC = __dmul_rn(add_outside(P, wc_shared, N, S.C, S.scalars, P.wc_outside, 0.0), P.hh);   // Solver.cpp:311, 314
        const bool one_row = r + T >= G;                                // this CTA owns a single row: its tap tables stay in shared memory
        for (int lx = r; lx < G; lx += T) {
            scan_body_of<SMALL, 0>(S, 0, lx, !one_row || k == 0, -1, true);   // per-node sums -> row_sum; ends with a CTA barrier
            for (int y = tid; y < G; y += NT) ring_buf[y] = 0.0;
            __syncthreads();
            for (int row = P.ring_row_begin[lx] + warp; row < P.ring_row_begin[lx + 1]; row += warps) {
                const double acc = ring_row_dot(P, S.C, row, lane);
                if (lane == 0) ring_buf[P.ring_point[row] - lx * G] = acc;
            }
            __syncthreads();
            for (int y = tid; y < G; y += NT) {
                const int q = lx * G + y;
                const double inside = (0.0 + row_sum[y]) - ring_buf[y];
                const double c = S.C[q];
                const double mc = __dmul_rn(add_outside(P, inside, N, S.C, S.scalars, P.out_weight[q], 0.0), P.hh);
                const double dc = second_moment_rhs(P, mc, c, N, WC, P.m_nodes[q], P.w_nodes[q]);
                double v = __dadd_rn(c, __dmul_rn(dc, S.dt));           // Solver.cpp:577
                v = v < 0.0 ? 0.0 : v;                                  // Solver.cpp:578-580
                S.C_next[q] = v;
                c_new[y] = v;
            }
            __syncthreads();
            if (warp == 0) {     // this row's dot with the WC functional for the next state
                const double s = wc_row_dot(P.wc_operator + (size_t)lx * G, c_new, G, lane);
                if (lane == 0) wc_rows[(parity ^ 1) * G + lx] = s;
            }
        }
        if (r == 0 && tid == 0) first_moment_update(P, N, WC, S.dt, S.scalars, S.N_next);
        if (!team_barrier(barrier, barrier_index++, T, A.status)) return;    // a lost partner: the host reports the status word
        parity ^= 1;
    }
}

}  // namespace

cudaError_t launch_persistent_scan(const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid)
{
    const DevProblem& P = args.step.P;
    const int G = P.G, n_off = P.n_off, nt = P.scan_nt;
    const void* kernel = nt <= 256 ? (const void*)persistent_scan_kernel<true, 0>
                       : scan_unroll_of(P) == 5 ? (const void*)persistent_scan_kernel<false, 5> : (const void*)persistent_scan_kernel<false, 4>;
    // shared memory: the step's working set, then the 16-bit y-side index table, then as many weight rows as still fit
    const size_t base = scan_smem_bytes(G, P.A, P.scan_bn, nt, P.scan_max_pieces);
    const size_t u_bytes = align32(sizeof(unsigned short) * (size_t)((n_off + 1) & ~1) * G);   // offsets two by two
    PersistentArgs with_cache = args;
    with_cache.cached_w_rows = -1;
    with_cache.ring_cache_offset = -1;
    // the CTA's ring share (one CSR row per warp, at most 8 entries per lane): 4-byte node indices, then 8-byte weights
    size_t ring_bytes = 0;
    if (P.ring_share_rows <= nt / 32 && P.ring_share_row_len <= 32 * 8)
        ring_bytes = align32(sizeof(int) * (size_t)((P.ring_share_entries + 1) & ~1) + sizeof(double) * (size_t)P.ring_share_entries + 8);
    size_t smem = base;
    if (base + ring_bytes <= kSmemBudget && ring_bytes > 0) {
        with_cache.ring_cache_offset = (long long)base;
        smem = base + ring_bytes;
    } else {
        ring_bytes = 0;
    }
    // The weight rows do not get all that is left: shared memory and L1 share the SM's 256 KB, and the E phase lives on L1 hits
    // for the tap rows that neighbouring alphas share.  Measured at 16 384 points (us per step against the budget of this
    // launch's shared memory): 226 KB 12.01, 196 KB 11.73, 164 KB 11.77, 132 KB and less 12.30 (the index table no longer fits).
    constexpr size_t kCachedRowsBudget = 196 * 1024;
    const size_t budget = std::min(kSmemBudget, kCachedRowsBudget);
    if (base + ring_bytes + u_bytes <= budget) {
        const int rows = (int)std::min<size_t>((size_t)((n_off + 1) & ~1), (budget - base - ring_bytes - u_bytes) / ((size_t)G * sizeof(double))) & ~1;
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

void chunk_block_shape(const DevProblem& P, int* alphas_per_block, int* groups, int* padded_alphas)
{
    const ScanLayout L = scan_layout(P.G, P.A, P.scan_bn, P.scan_nt, P.scan_max_pieces);
    *alphas_per_block = scan_unroll(P.scan_nt, P.A, L.ng);
    *groups = L.ng;
    *padded_alphas = L.a_pad;
}

cudaError_t launch_persistent_chunks(const ChunkedArgs& args, int grid, cudaStream_t stream, int* max_grid)
{
    const DevProblem& P = args.step.P;
    const int nt = P.scan_nt;
    const void* kernel = nt <= 256 ? (const void*)persistent_chunks_kernel<true, 0>
                       : scan_unroll_of(P) == 5 ? (const void*)persistent_chunks_kernel<false, 5> : (const void*)persistent_chunks_kernel<false, 4>;
    const size_t smem = scan_smem_bytes(P.G, P.A, P.scan_bn, nt, P.scan_max_pieces);
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
    ChunkedArgs a = args;
    void* params[] = {&a};
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(nt), params, smem, stream);
}

cudaError_t launch_persistent_batch(const BatchArgs& args, int grid, int nt, size_t smem, cudaStream_t stream, int* max_grid)
{
    const void* kernel = nt <= 256 ? (const void*)persistent_batch_kernel<true> : (const void*)persistent_batch_kernel<false>;
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
    BatchArgs a = args;
    void* params[] = {&a};
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(nt), params, smem, stream);
}

#ifdef LDM_PHASE_CLOCKS
extern "C" LDM_API int ldm_debug_trace_reset()
{
    cudaDeviceSynchronize();
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_trace_n) != cudaSuccess) return 1;
    return (int)cudaMemset(p, 0, sizeof(unsigned int) * kTraceCtas);
}
// events: [148][4096][2] 64-bit words (globaltimer ns; type << 48 | step << 32 | local row << 16 | chunk), counts: [148]
extern "C" LDM_API int ldm_debug_trace_read(unsigned long long* events, unsigned int* counts)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(unsigned int) * kTraceCtas) != cudaSuccess) return 1;
    return (int)cudaMemcpyFromSymbol(events, g_trace, sizeof(ulonglong2) * kTraceCtas * kTraceEvents);
}
extern "C" LDM_API int ldm_debug_persist_clocks(long long* out, int rows)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_scan_clock, sizeof(long long) * 12 * (size_t)rows);
}
#endif

}  // namespace ldm
