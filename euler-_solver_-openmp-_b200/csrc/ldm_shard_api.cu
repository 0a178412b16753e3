// The xi-sharded entry points of include/ldm.h: one large problem, moment-grid rows dealt out cyclically over
// the ranks, every rank holding a full replica of C.  Two ways to run a step:
//   * ldm_shard_step: this rank's rows into a caller-owned buffer; the caller all-gathers (NCCL) -- the reference variant;
//   * ldm_shard_peer_*: the all-gather fused into the update.  On the prefix-sum path that is the persistent work-list
//     kernel (ldm_persist.cu persistent_chunks_kernel): any number of Euler steps in ONE launch per rank, rows stored into
//     every rank's next replica over NVLink as they close, per-rank epoch flags instead of barriers.  The same kernel on
//     one rank carries ldm_advance / ldm_solve of grids whose rows span several chunks (advance_chunked).
#include <cuda_runtime.h>

#include <cstring>
#include <numeric>

#include "ldm_context.h"

namespace ldm {

namespace {

// (local row, chunk) items of this rank in the order they are handed out: rows in ascending order -- so that rows close, and
// ship, progressively and the next step's first rows find their windows complete while this step's last rows are still
// under way -- and within a row the chunk with the most y-side work first (on a stretched grid the far chunks own twice the
// pairs of the near ones).  Also the window of rows each row's E phase reads: the interpolation pairs of all its offsets.
void ensure_items(ldm_context* c)
{
    const DevProblem& d = c->dev;
    if (!c->d_row_window) {
        const std::vector<int>& upper = c->tables.axisM.upper;                    // [G][n_off]; 0: the sample lies beyond the grid
        std::vector<int> window((size_t)2 * d.G);
        for (int x = 0; x < d.G; x++) {
            int lo = x, hi = x;                                                   // a row always waits for itself: its chunk sums and
                                                                                  // arrival counter are reused step after step
            for (int o = 0; o < d.n_off; o++) {
                const int u = upper[(size_t)x * d.n_off + o];
                if (u != 0) { lo = std::min(lo, u - 1); hi = std::max(hi, u); }
            }
            window[2 * (size_t)x] = lo; window[2 * (size_t)x + 1] = std::min(hi, d.G - 1);
        }
        c->d_row_window = reinterpret_cast<int2*>(c->upload(window));
        // the same per block of alphas: an item that starts before the state is complete waits block by block
        int U = 0, ng = 0, a_pad = 0;
        chunk_block_shape(d, &U, &ng, &a_pad);
        const int blocks = a_pad / U, n_fold = d.n_off / d.A;
        std::vector<int> bw((size_t)4 * d.G * blocks);
        for (int x = 0; x < d.G; x++)
            for (int b = 0; b < blocks; b++) {
                int range[2][2] = {{d.G, -1}, {d.G, -1}};                         // empty: lo > hi
                for (int alpha = b * U; alpha < std::min((b + 1) * U, d.A); alpha++)
                    for (int f = 0; f < n_fold; f++) {
                        const int o = n_fold == 1 ? 0 : (f == 0 ? d.A - 1 - alpha : d.A + alpha);
                        const int u = upper[(size_t)x * d.n_off + o];
                        if (u != 0) { range[f][0] = std::min(range[f][0], u - 1); range[f][1] = std::max(range[f][1], u); }
                    }
                int* out = &bw[((size_t)x * blocks + b) * 4];
                out[0] = range[0][0]; out[1] = range[0][1]; out[2] = range[1][0]; out[3] = range[1][1];
            }
        c->d_block_window = reinterpret_cast<int4*>(c->upload(bw));
        c->blocks_per_row = blocks;
    }
    if (c->d_items && c->items_rank == d.rank && c->items_world == d.world) return;
    const int n_local = local_rows(d);
    const std::vector<int>& upper = c->tables.axisM.upper;
    std::vector<int> items;
    items.reserve((size_t)4 * n_local * d.scan_chunks);
    auto push = [&](int lx, int chunk) {
        const int x = d.rank + d.world * lx;
        int lo = x, hi = x;
        for (int o = 0; o < d.n_off; o++) {
            const int u = upper[(size_t)x * d.n_off + o];
            if (u != 0) { lo = std::min(lo, u - 1); hi = std::max(hi, u); }
        }
        items.push_back(lx); items.push_back(chunk); items.push_back(lo); items.push_back(std::min(hi, d.G - 1));
    };
    const int order = tuning().item_order;
    if (order == 1) {
        for (int k = 0; k < d.scan_chunks; k++)
            for (int lx = 0; lx < n_local; lx++) push(lx, c->chunk_order_host[(size_t)k]);
    } else {
        for (int i = 0; i < n_local; i++)
            for (int k = 0; k < d.scan_chunks; k++) push(order == 2 ? n_local - 1 - i : i, c->chunk_order_host[(size_t)k]);
    }
    if (c->d_items) cudaFree(c->d_items);
    c->d_items = nullptr;
    LDM_CUDA(cudaMalloc((void**)&c->d_items, items.size() * sizeof(int)));
    LDM_CUDA(cudaMemcpy(c->d_items, items.data(), items.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->n_items = n_local * d.scan_chunks;
    c->items_rank = d.rank;
    c->items_world = d.world;
}

// The ring correction inside the work-list kernel.  Light grid rows -- a prefix of this rank's rows, as long as a row's
// entries fit the staging area in the idle H tile -- are corrected by the CTA that closes them; the CSR rows of the heavy
// rows behind them (on a stretched grid the last few rows own most samples beyond the circle) are dealt out over the `grid`
// CTAs of a launch in contiguous shares of about equal entry count.
void ensure_ring_cut(ldm_context* c, int grid)
{
    const DevProblem& d = c->dev;
    if (c->d_ring_cut && c->cut_grid == grid && c->cut_rank == d.rank && c->cut_world == d.world) return;
    const std::vector<int>& begin = c->ring_row_begin_host;
    const std::vector<int>& grid_row = c->ring_grid_row_begin_host;
    const int rows = (int)begin.size() - 1, n_local = (int)grid_row.size() - 1;
    int U = 0, ng = 0, a_pad = 0;
    chunk_block_shape(d, &U, &ng, &a_pad);
    const long long tile = (long long)d.A * (d.scan_bn + 2);                       // doubles of the H tile (ldm_scan_body.cuh ScanLayout::hs)
    c->self_ring_capacity = tile >= d.G ? (int)std::min<long long>(4096, (tile - d.G) / 2) : -1;
    c->self_ring_rows = 0;
    while (c->self_ring_rows < n_local && c->self_ring_capacity >= 0 &&
           begin[(size_t)grid_row[(size_t)c->self_ring_rows + 1]] - begin[(size_t)grid_row[(size_t)c->self_ring_rows]] <= c->self_ring_capacity)
        c->self_ring_rows++;
    const int first = n_local > 0 ? grid_row[(size_t)c->self_ring_rows] : 0;    // first CSR row that goes through the shares
    const long long base = rows > 0 ? begin[(size_t)first] : 0, total = rows > 0 ? begin[(size_t)rows] - base : 0;
    std::vector<int> cut((size_t)grid + 1, std::max(rows, 0));
    cut[0] = first;
    int row = first;
    for (int s = 1; s < grid; s++) {
        const long long want = base + total * s / grid;
        while (row < rows && begin[(size_t)row] < want) row++;
        cut[(size_t)s] = row;
    }
    if (c->d_ring_cut) cudaFree(c->d_ring_cut);
    c->d_ring_cut = nullptr;
    LDM_CUDA(cudaMalloc((void**)&c->d_ring_cut, cut.size() * sizeof(int)));
    LDM_CUDA(cudaMemcpy(c->d_ring_cut, cut.data(), cut.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->cut_grid = grid;
    c->cut_rank = d.rank;
    c->cut_world = d.world;
}

void peer_buffers_alloc(ldm_context* h)
{
    const size_t words = (size_t)h->dev.world * h->dev.rows_per_rank * h->dev.G + (size_t)h->dev.G;   // the replica and its tail of WC row dots
    for (int b = 0; b < 2; b++) {
        if (h->peer_own[b]) cudaFree(h->peer_own[b]);
        h->peer_own[b] = nullptr;
        LDM_CUDA(cudaMalloc((void**)&h->peer_own[b], words * sizeof(double)));
        LDM_CUDA(cudaMemset(h->peer_own[b], 0, words * sizeof(double)));
    }
    if (!h->peer_own_flags) LDM_CUDA(cudaMalloc((void**)&h->peer_own_flags, kFlagWords(h->dev.G) * sizeof(unsigned long long)));
    if (!h->peer_done) LDM_CUDA(cudaMalloc((void**)&h->peer_done, sizeof(unsigned int)));
    if (!h->peer_status) LDM_CUDA(cudaMalloc((void**)&h->peer_status, sizeof(int)));
    LDM_CUDA(cudaMemset(h->peer_own_flags, 0, kFlagWords(h->dev.G) * sizeof(unsigned long long)));
    LDM_CUDA(cudaMemset(h->peer_done, 0, sizeof(unsigned int)));
    LDM_CUDA(cudaMemset(h->peer_status, 0, sizeof(int)));
}

}  // namespace

// The work-list kernel applies on the prefix-sum path (the closing phase is built on its chunk sums), SquaredN / Zero / Alpha
// extrapolation, when its grid is co-resident.
bool chunked_possible(ldm_context* c)
{
    if (!c->persistent_enabled || c->species != 1 || c->algo != LDM_ALGO_PREFIX || !c->plan.scan) return false;
    if (c->dev.extrapolation == LDM_EXTRAPOLATION_EXPONENTIAL) return false;        // needs its own launches every step
    if (c->closure != LDM_CLOSURE_POWER1) return false;
    if (c->chunk_grid == 0) {
        ChunkedArgs probe{};
        probe.step = make_step_args(c, c->d_C[0], c->d_N, c->d_N + 1, 0.0, nullptr, nullptr, nullptr);
        int max_grid = 0;
        const cudaError_t e = launch_persistent_chunks(probe, 0, c->stream, &max_grid);
        c->chunk_grid = (e == cudaSuccess && max_grid > 0) ? max_grid : -1;
        if (e != cudaSuccess) cudaGetLastError();
    }
    return c->chunk_grid > 0;
}

// Everything a launch of the work-list kernel needs besides the state, built once per (rank, world): ring operator, work
// list, ring shares, counters.  Some of it synchronises the device (cudaFree, the ring operator's upload), so a peer
// group sets it up BEFORE anybody steps (ldm_shard_peer_upload) -- several ranks may share one device in one process,
// and a rank that synchronises the device while another rank's kernel waits for it would never come back.
int prepare_chunked(ldm_context* c)
{
    const DevProblem& d = c->dev;
    ensure_ring_operator(c);
    ensure_items(c);
    int grid = std::min(c->chunk_grid, c->n_items);
    if (c->chunk_grid_limit > 0) grid = std::min(grid, c->chunk_grid_limit);
    ensure_ring_cut(c, grid);
    if (!c->d_counters) c->d_counters = c->alloc<unsigned int>(3 + (size_t)d.G);
    if (!c->d_ring2) c->d_ring2 = c->alloc<double>(2 * c->P());
    return grid;
}

void advance_chunked(ldm_context* c, int n_steps, double dt, cudaStream_t s, long long first_point, double* d_obs, double* d_values,
                     double* d_derivs)
{
    if (n_steps <= 0) return;
    const DevProblem& d = c->dev;
    const bool peers = c->peer_connected;
    if (!peers && d.world != 1) throw std::invalid_argument("a sharded handle steps through its peer group (ldm_shard_peer_connect)");
    const int grid = prepare_chunked(c);
    const int n_local = local_rows(d);
    const size_t P_local = (size_t)n_local * d.G;

    ChunkedArgs a{};
    a.world = d.world; a.rank = d.rank;
    a.N = c->d_N;
    a.items = c->d_items; a.n_items = c->n_items;
    a.row_window = c->d_row_window;
    a.block_window = c->d_block_window; a.blocks_per_row = c->blocks_per_row;
    a.counters = c->d_counters;
    a.ring_pingpong = c->d_ring2;
    a.ring_cta_begin = c->d_ring_cut;
    a.self_ring_rows = c->self_ring_rows; a.self_ring_capacity = c->self_ring_capacity;
    a.status = c->d_status;
    a.stride = (long long)(1 + c->P());
    if (peers) {
        a.words = (long long)d.world * d.rows_per_rank * d.G;
        for (int r = 0; r < d.world; r++) {
            a.replica[0][r] = c->peer_C[0][r]; a.replica[1][r] = c->peer_C[1][r];
            a.flags[r] = c->peer_flags[r];
        }
    } else {
        a.words = (long long)c->P();
        a.replica[0][0] = c->d_C[0]; a.replica[1][0] = c->d_C[1];
        a.flags[0] = c->d_flags;
    }
    // the item counter is monotone over the steps of a launch: keep it well inside 32 bits
    const int per_launch = std::max(1, (int)std::min<long long>(n_steps, ((1LL << 30) - grid) / std::max(1, c->n_items)));
    for (int done = 0; done < n_steps; done += per_launch) {
        const int steps = std::min(per_launch, n_steps - done);
        const int parity = peers ? c->peer_parity : c->cur;
        a.step = make_step_args(c, a.replica[parity][d.rank], c->d_N + parity, c->d_N + (parity ^ 1), dt, nullptr, nullptr, nullptr);
        a.step.wc_rows = nullptr;                       // the WC row dots live in the replicas' tails here
        a.n_steps = steps;
        a.parity = parity;
        a.epoch = peers ? c->peer_epoch : c->local_epoch;
        a.first_point = first_point + done;
        a.observables = d_obs; a.values = d_values; a.derivs = d_derivs;
        LDM_CUDA(cudaMemsetAsync(c->d_counters, 0, (3 + (size_t)d.G) * sizeof(unsigned int), s));
        LDM_CUDA(cudaMemsetAsync(c->d_ring2, 0xFF, 2 * P_local * sizeof(double), s));      // every entry: not yet published
        LDM_CUDA(launch_persistent_chunks(a, grid, s, nullptr));
        c->launches += 1;
        if (peers) { c->peer_epoch += (unsigned long long)steps; c->peer_parity ^= (steps & 1); c->cur = c->peer_parity; }
        else { c->local_epoch += (unsigned long long)steps; c->cur ^= (steps & 1); }
    }
}

}  // namespace ldm

using namespace ldm;

extern "C" {

int ldm_shard_configure(ldm_handle h, int rank, int world)
{
    if (!h || world < 1 || rank < 0 || rank >= world) return fail(LDM_ERR_ARGUMENT, "bad rank/world");
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (world > h->dev.G) return fail(LDM_ERR_ARGUMENT, "more ranks than moment-grid rows");
    h->dev.world = world;
    h->dev.rank = rank;
    h->dev.rows_per_rank = (h->dev.G + world - 1) / world;
    h->peer_connected = false;          // a peer group belongs to one (rank, world)
    return LDM_OK;
}

int ldm_shard_rows_per_rank(ldm_handle h) { return h ? h->dev.rows_per_rank : 0; }

int ldm_shard_set_N(ldm_handle h, double N)
{
    if (h && h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        h->cur = 0;
        LDM_CUDA(cudaMemcpy(h->d_N, &N, sizeof(double), cudaMemcpyHostToDevice));
    });
}

int ldm_shard_get_N(ldm_handle h, double* N)
{
    if (h && h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (!h || !N) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        LDM_CUDA(cudaDeviceSynchronize());
        LDM_CUDA(cudaMemcpy(N, h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToHost));
    });
}

int ldm_shard_step(ldm_handle h, const double* d_C_gathered, double* d_C_shard_out, double dt, void* cuda_stream)
{
    if (h && h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (!h || !d_C_gathered || !d_C_shard_out) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
        const int cur = h->cur, nxt = cur ^ 1;
        rhs_and_update(h, h->dev, d_C_gathered, h->d_N + cur, h->d_N + nxt, dt, nullptr, nullptr, d_C_shard_out, s);
        h->cur = nxt;
        LDM_CUDA(cudaGetLastError());
    });
}

int ldm_shard_peer_alloc(ldm_handle h, unsigned char* handles)
{
    if (!h || !handles) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (h->dev.world > ldm::kMaxPeers) return fail(LDM_ERR_ARGUMENT, "at most 8 ranks");
    if (h->closure != LDM_CLOSURE_POWER1) return fail(LDM_ERR_UNSUPPORTED, "the fused exchange carries the reference's closure only: use ldm_shard_step (ldm_set_closure)");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        peer_buffers_alloc(h);
        static_assert(sizeof(cudaIpcMemHandle_t) == LDM_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t ipc;
        void* exported[3] = {h->peer_own[0], h->peer_own[1], h->peer_own_flags};
        for (int i = 0; i < 3; i++) {
            LDM_CUDA(cudaIpcGetMemHandle(&ipc, exported[i]));
            std::memcpy(handles + (size_t)i * LDM_IPC_HANDLE_BYTES, &ipc, LDM_IPC_HANDLE_BYTES);
        }
        LDM_CUDA(cudaDeviceSynchronize());
    });
}

int ldm_shard_peer_connect(ldm_handle h, const unsigned char* all_handles)
{
    if (!h || !all_handles) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (!h->peer_own[0]) return fail(LDM_ERR_ARGUMENT, "call ldm_shard_peer_alloc first");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        for (void* p : h->peer_opened) cudaIpcCloseMemHandle(p);
        h->peer_opened.clear();
        for (int r = 0; r < h->dev.world; r++) {
            void* mapped[3] = {h->peer_own[0], h->peer_own[1], h->peer_own_flags};
            if (r != h->dev.rank) {
                for (int i = 0; i < 3; i++) {
                    cudaIpcMemHandle_t ipc;
                    std::memcpy(&ipc, all_handles + ((size_t)r * 3 + i) * LDM_IPC_HANDLE_BYTES, LDM_IPC_HANDLE_BYTES);
                    LDM_CUDA(cudaIpcOpenMemHandle(&mapped[i], ipc, cudaIpcMemLazyEnablePeerAccess));
                    h->peer_opened.push_back(mapped[i]);
                }
            }
            h->peer_C[0][r] = static_cast<double*>(mapped[0]);
            h->peer_C[1][r] = static_cast<double*>(mapped[1]);
            h->peer_flags[r] = static_cast<unsigned long long*>(mapped[2]);
        }
        h->peer_connected = true;
    });
}

int ldm_shard_peer_connect_local(int world, const ldm_handle* handles)
{
    if (world < 1 || world > ldm::kMaxPeers || !handles) return fail(LDM_ERR_ARGUMENT, "1 .. 8 handles");
    for (int r = 0; r < world; r++) {
        if (!handles[r]) return fail(LDM_ERR_ARGUMENT, "NULL handle");
        if (handles[r]->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
        if (handles[r]->closure != LDM_CLOSURE_POWER1) return fail(LDM_ERR_UNSUPPORTED, "the fused exchange carries the reference's closure only: use ldm_shard_step (ldm_set_closure)");
        if (handles[r]->dev.world != world || handles[r]->dev.rank != r || handles[r]->dev.G != handles[0]->dev.G)
            return fail(LDM_ERR_ARGUMENT, "handles[r] must be configured as rank r of `world` on the same grid (ldm_shard_configure)");
    }
    return guarded([&] {
        for (int r = 0; r < world; r++) {
            ldm_context* h = handles[r];
            LDM_CUDA(cudaSetDevice(h->device));
            peer_buffers_alloc(h);
            for (int q = 0; q < world; q++) {
                if (handles[q]->device == h->device) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else LDM_CUDA(e);
            }
            LDM_CUDA(cudaDeviceSynchronize());
        }
        for (int r = 0; r < world; r++) {
            ldm_context* h = handles[r];
            for (void* p : h->peer_opened) cudaIpcCloseMemHandle(p);
            h->peer_opened.clear();
            for (int q = 0; q < world; q++) {
                h->peer_C[0][q] = handles[q]->peer_own[0];
                h->peer_C[1][q] = handles[q]->peer_own[1];
                h->peer_flags[q] = handles[q]->peer_own_flags;
            }
            h->peer_connected = true;
        }
    });
}

#ifdef LDM_PHASE_CLOCKS
// Measurement aid of the debug build (tools/shard_rank_timing.py): rank `rank` of `world` steps ALONE -- every peer's
// replica is this rank's own, the other ranks' flags count as published -- so that one GPU shows what a rank of a
// W-rank run computes and how its items and closings schedule, without the exchange.  The other ranks' rows stay stale.
LDM_API int ldm_debug_shard_solo(ldm_handle h)
{
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        peer_buffers_alloc(h);
        unsigned long long* scratch = h->alloc<unsigned long long>((size_t)ldm::kFlagWords(h->dev.G));
        for (int q = 0; q < h->dev.world; q++) {
            h->peer_C[0][q] = h->peer_own[0];
            h->peer_C[1][q] = h->peer_own[1];
            h->peer_flags[q] = q == h->dev.rank ? h->peer_own_flags : scratch;
        }
        h->peer_connected = true;
        h->peer_solo = true;
        LDM_CUDA(cudaDeviceSynchronize());
    });
}
#endif

int ldm_shard_peer_upload(ldm_handle h, const double* C_gathered, double N)
{
    if (!h || !C_gathered) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (!h->peer_connected) return fail(LDM_ERR_ARGUMENT, "peers are not connected");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        const size_t words = (size_t)h->dev.world * h->dev.rows_per_rank * h->dev.G;
        LDM_CUDA(cudaDeviceSynchronize());
        LDM_CUDA(cudaMemcpy(h->peer_own[0], C_gathered, words * sizeof(double), cudaMemcpyHostToDevice));
        LDM_CUDA(cudaMemset(h->peer_own_flags, 0, ldm::kFlagWords(h->dev.G) * sizeof(unsigned long long)));
        LDM_CUDA(cudaMemset(h->peer_done, 0, sizeof(unsigned int)));
        LDM_CUDA(cudaMemset(h->peer_status, 0, sizeof(int)));
        LDM_CUDA(cudaMemset(h->d_status, 0, sizeof(int)));
        h->cur = 0;
        LDM_CUDA(cudaMemcpy(h->d_N, &N, sizeof(double), cudaMemcpyHostToDevice));
        h->peer_epoch = 0;
        h->peer_parity = 0;
        if (h->peer_solo) {                                 // the other ranks and their rows count as published, for ever
            LDM_CUDA(cudaMemcpy(h->peer_own[1], C_gathered, words * sizeof(double), cudaMemcpyHostToDevice));
            std::vector<unsigned long long> flags((size_t)ldm::kFlagWords(h->dev.G), 0ULL);
            for (int q = 0; q < h->dev.world; q++) if (q != h->dev.rank) flags[(size_t)q] = 1ULL << 62;
            for (int x = 0; x < h->dev.G; x++) if (x % h->dev.world != h->dev.rank) flags[(size_t)ldm::kMaxPeers + x] = 1ULL << 62;
            LDM_CUDA(cudaMemcpy(h->peer_own_flags, flags.data(), flags.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        }
        if (chunked_possible(h)) prepare_chunked(h);       // before anybody steps (see prepare_chunked)
        else ensure_ring_operator(h);
        LDM_CUDA(cudaDeviceSynchronize());
    });
}

int ldm_shard_peer_advance(ldm_handle h, int n_steps, double dt, void* cuda_stream)
{
    if (!h || n_steps < 0) return fail(LDM_ERR_ARGUMENT, "bad argument");
    if (!h->peer_connected) return fail(LDM_ERR_ARGUMENT, "peers are not connected");
    if (h->algo == LDM_ALGO_DIRECT) return fail(LDM_ERR_UNSUPPORTED, "the fused exchange rides on the factored path");
    if (h->dev.extrapolation == LDM_EXTRAPOLATION_EXPONENTIAL) return fail(LDM_ERR_UNSUPPORTED, "Exponential extrapolation: use ldm_shard_step");
    if (h->closure != LDM_CLOSURE_POWER1) return fail(LDM_ERR_UNSUPPORTED, "a closure over the w quadrature: use ldm_shard_step");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
        ldm_context* c = h;
        if (chunked_possible(c)) {
            // prefix-sum path: all n_steps in one launch of the persistent work-list kernel
            advance_chunked(c, n_steps, dt, s, 0, nullptr, nullptr, nullptr);
            LDM_CUDA(cudaGetLastError());
            return;
        }
        // dense contraction (a table that is not separable): per step a one-warp wait for the peers' previous step, the
        // integral kernels, and the closing kernel that stores the updated rows into every rank's next replica
        ensure_ring_operator(c);
        for (int k = 0; k < n_steps; k++) {
            const int cur = c->cur, nxt = cur ^ 1;
            const int pb = c->peer_parity;
            ldm::StepArgs args = make_step_args(c, c->peer_own[pb], c->d_N + cur, c->d_N + nxt, dt, nullptr, nullptr, nullptr);
            args.peer_count = c->dev.world;
            for (int r = 0; r < c->dev.world; r++) {
                args.peer_C_next[r] = c->peer_C[pb ^ 1][r];
                args.peer_flags[r] = c->peer_flags[r];
            }
            args.done_counter = c->peer_done;
            args.epoch = c->peer_epoch + 1;
            args.peer_status = c->peer_status;
            ldm::launch_peer_wait(c->peer_own_flags, c->dev.world, c->peer_epoch, c->peer_status, s);
            const bool side = args.P.n_ring > 0;
            if (side) {
                LDM_CUDA(cudaEventRecord(c->ev_fork, s));
                LDM_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
                ldm::launch_ring(args, c->aux_stream);
                LDM_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
            }
            if (c->plan.scan) ldm::launch_scan(args, s);
            else ldm::launch_factored(c->plan, args, s);
            if (side) LDM_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
            ldm::launch_finish_factored(args, s);
            c->launches += 3 + (side ? 1 : 0);
            c->cur = nxt;
            c->peer_parity = pb ^ 1;
            c->peer_epoch += 1;
        }
        LDM_CUDA(cudaGetLastError());
    });
}

int ldm_shard_peer_step(ldm_handle h, double dt, void* cuda_stream) { return ldm_shard_peer_advance(h, 1, dt, cuda_stream); }

int ldm_shard_peer_download(ldm_handle h, double* C_gathered, double* N)
{
    if (!h || !C_gathered) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (!h->peer_connected) return fail(LDM_ERR_ARGUMENT, "peers are not connected");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        const size_t words = (size_t)h->dev.world * h->dev.rows_per_rank * h->dev.G;
        LDM_CUDA(cudaDeviceSynchronize());
        check_device_status(h);              // this rank's own kernels gave up on a partner: say so at once
        // the current replica is complete once every rank has published the last step
        ldm::launch_peer_wait(h->peer_own_flags, h->dev.world, h->peer_epoch, h->peer_status, h->stream);
        LDM_CUDA(cudaDeviceSynchronize());
        int status = 0;
        LDM_CUDA(cudaMemcpy(&status, h->peer_status, sizeof(int), cudaMemcpyDeviceToHost));
        if (status != 0) throw CudaError("rank " + std::to_string(status - 1) + " never published its step (peer exchange timed out)");
        LDM_CUDA(cudaMemcpy(C_gathered, h->peer_own[h->peer_parity], words * sizeof(double), cudaMemcpyDeviceToHost));
        if (N) LDM_CUDA(cudaMemcpy(N, h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
