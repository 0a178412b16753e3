// extern "C" ABI of libldm_b200.so (include/ldm.h): handles, host<->device plumbing, the stepping loop.
// S:n = "/root/reference/Euler Solver ver.5.3/Solver.cpp" line n.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <type_traits>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "ldm_context.h"

namespace {

thread_local std::string g_error;

// |Q - q q^T| <= tolerance * q q^T decides whether a kernel table takes the prefix-sum path: libm's exp leaves ~2e-15 on
// the reference's tables; anything beyond 1e-13 is structure, not rounding
constexpr double kSeparableTolerance = 1e-13;

}  // namespace

namespace ldm {

int fail(int code, const std::string& what)
{
    g_error = what;
    return code;
}

Tuning& tuning()
{
    static Tuning t;
    return t;
}


// The ring-correction operator of the factored path for this rank's rows (time-invariant; rebuilt when the
// sharding changes).  Built lazily: the direct algorithm never needs it.
void ensure_ring_operator(ldm_context* c)
{
    DevProblem& d = c->dev;
    if (c->ring_rank == d.rank && c->ring_world == d.world) return;
    ldm::RingOperator op;
    ldm::build_ring_operator(c->tables, d.rank, d.world, op);
    for (void* p : c->ring_allocations) cudaFree(p);
    c->ring_allocations.clear();
    auto up = [&](const auto& v) {
        using T = typename std::decay_t<decltype(v)>::value_type;
        void* p = nullptr;
        LDM_CUDA(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
        c->ring_allocations.push_back(p);
        if (!v.empty()) LDM_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return static_cast<const T*>(p);
    };
    d.n_ring = (int)op.point.size();
    d.ring_point = up(op.point);
    d.ring_begin = up(op.row_begin);
    d.ring_row_begin = up(op.grid_row_begin);
    {
        std::vector<unsigned char> present((size_t)ldm::local_rows(d) * d.G, 0);
        for (int point : op.point) present[(size_t)point] = 1;
        d.ring_present = up(present);
    }
    {   // the CSR rows in contiguous shares of about equal entry count, one share per local grid row (= CTA of the persistent
        // kernel) -- and of at most one CSR row per warp of that CTA, the shape its shared-memory copy of the share needs (cut by
        // entries alone, the shares in front of the far rows collect up to 23 short CSR rows at 16 384 points and the copy was
        // off for every CTA): the smallest entry budget with which the rows fit the shares under both limits
        const int shares = (int)op.grid_row_begin.size() - 1, rows = (int)op.point.size();
        const long long total = (long long)op.node.size();
        const int cap = d.scan_bn > 0 ? std::max(1, d.scan_nt / 32) : 16;
        auto length = [&](int r) { return (long long)(op.row_begin[(size_t)r + 1] - op.row_begin[(size_t)r]); };
        std::vector<int> cut((size_t)shares + 1, rows);
        cut[0] = 0;
        auto deal = [&](long long budget, bool keep) {              // shares the rows take with this entry budget
            int count = 0, r = 0;
            while (r < rows) {
                int n = 0;
                long long e = 0;
                while (r < rows && n < cap && (n == 0 || e + length(r) <= budget)) { e += length(r); n++; r++; }
                count++;
                if (keep && count < shares) cut[(size_t)count] = r;
            }
            return count;
        };
        if (rows > 0 && deal(total, false) <= shares) {
            long long lo = 1, hi = total;                               // smallest budget that needs no more than `shares` shares
            while (lo < hi) {
                const long long mid = lo + (hi - lo) / 2;
                if (deal(mid, false) <= shares) hi = mid; else lo = mid + 1;
            }
            deal(lo, true);
        } else {                                                        // more CSR rows than warps in all CTAs together: by entries alone
            int row = 0;
            for (int s = 1; s < shares; s++) {
                const long long want = total * s / shares;
                while (row < rows && op.row_begin[(size_t)row] < want) row++;
                cut[(size_t)s] = row;
            }
        }
        d.ring_cta_begin = up(cut);
        // shape of the shares, for the persistent kernel's shared-memory copy of its share: the most entries, CSR rows and
        // entries per row any share holds
        d.ring_share_entries = 0; d.ring_share_rows = 0; d.ring_share_row_len = 0;
        for (int s = 0; s < shares; s++) {
            const int r0 = cut[(size_t)s], r1 = cut[(size_t)s + 1];
            d.ring_share_entries = std::max(d.ring_share_entries, op.row_begin[(size_t)r1] - op.row_begin[(size_t)r0]);
            d.ring_share_rows = std::max(d.ring_share_rows, r1 - r0);
            for (int r = r0; r < r1; r++) d.ring_share_row_len = std::max(d.ring_share_row_len, op.row_begin[(size_t)r + 1] - op.row_begin[(size_t)r]);
        }
    }
#ifdef LDM_PHASE_CLOCKS
    std::fprintf(stderr, "ring operator: G %d, %d CSR rows, %zu entries; shares: most entries %d, most rows %d, longest row %d\n", d.G,
                 (int)op.point.size(), op.node.size(), d.ring_share_entries, d.ring_share_rows, d.ring_share_row_len);
#endif
    d.ring_node = up(op.node);
    d.ring_weight = up(op.weight);
    LDM_CUDA(cudaMemset(c->d_ring, 0, c->P() * sizeof(double)));
    c->ring_rank = d.rank;
    c->ring_world = d.world;
    c->ring_entries = (long long)op.node.size();
    c->ring_samples = op.samples;
    c->ring_row_begin_host = op.row_begin;
    c->ring_grid_row_begin_host = op.grid_row_begin;
    LDM_CUDA(cudaDeviceSynchronize());      // one-time setup: everything above is visible to every stream
}

// One right-hand side at the state (N_cur, C) for this rank's rows: scalars <- WC, dN; optional outputs:
// sum_out (inside-sample sums), deriv_out (dC/dt), N_next / C_next (explicit-Euler update with clamp).
ldm::StepArgs make_step_args(ldm_context* c, const double* C, const double* N_cur, double* N_next, double dt,
                             double* sum_out, double* deriv_out, double* C_next)
{
    ldm::StepArgs a{};
    a.P = c->dev;
    a.C = C; a.N_cur = N_cur; a.N_next = N_next;
    a.partial = c->d_partial; a.ring = c->d_ring; a.wc_rows = c->d_wc_rows; a.scalars = c->d_scalars;
    a.sum_out = sum_out; a.deriv_out = deriv_out; a.C_next = C_next;
    a.out_exp = nullptr;
    a.tiles = c->plan.tiles; a.chunks = c->plan.chunks;
    a.peer_count = 0; a.done_counter = nullptr; a.epoch = 0;
    a.dt = dt;
    return a;
}

// the m*C kernel of the active plan: prefix sums for a separable table, otherwise the dense contraction
void launch_integral(ldm_context* c, const ldm::StepArgs& args, cudaStream_t s)
{
    if (c->plan.scan) ldm::launch_scan(args, s);
    else ldm::launch_factored(c->plan, args, s);
}

void rhs_and_update(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, double* N_next, double dt,
                    double* sum_out, double* deriv_out, double* C_next, cudaStream_t s)
{
    if (c->closure != LDM_CLOSURE_POWER1) {
        // a closure that does not factorise (ldm_closure.cu): the m*C sums, WC and the first-moment equation as always --
        // this function once more, as the reference's closure would run it, closing nothing -- then the integrals over the
        // w quadrature and the closing with the exchanged term
        const int closure = c->closure;
        c->closure = LDM_CLOSURE_POWER1;
        try { rhs_and_update(c, P, C, N_cur, N_next, dt, c->d_sum, nullptr, nullptr, s); }
        catch (...) { c->closure = closure; throw; }
        c->closure = closure;
        if (sum_out && sum_out != c->d_sum)
            LDM_CUDA(cudaMemcpyAsync(sum_out, c->d_sum, c->P() * sizeof(double), cudaMemcpyDeviceToDevice, s));
        closure_close(c, P, C, N_cur, dt, deriv_out, C_next, s);
        return;
    }
    const double* out_exp = nullptr;
    if (P.extrapolation == LDM_EXTRAPOLATION_EXPONENTIAL) {
        // the one type whose outside value depends on each sample's own distance, through a decay rate that changes every
        // step (S:392-408): the per-step scalars, then the outside samples of every node enumerated one by one
        if (!c->d_out_exp) c->d_out_exp = c->alloc<double>(c->P());
        ldm::launch_exponential_prepare(P, C, N_cur, c->d_scalars, s);
        ldm::launch_sample_sum(P, ldm::SAMPLES_OUT_EXP, C, c->d_out_exp, s, c->d_scalars);
        c->launches += 2;
        out_exp = c->d_out_exp;
    }
    if (c->algo == LDM_ALGO_DIRECT) {
        ldm::launch_sample_sum(P, ldm::SAMPLES_INSIDE, C, c->d_sum, s);
        ldm::launch_wc(P, C, N_cur, N_next, c->d_scalars, dt, s);
        c->launches += 2;
        if (!deriv_out && !C_next) return;             // integrals only (ldm_integrals, a closure that closes elsewhere)
        ldm::launch_pointwise(P, C, c->d_sum, N_cur, c->d_scalars, dt, deriv_out, C_next, s, out_exp);
        c->launches += 1;
        return;
    }
    ensure_ring_operator(c);
    ldm::StepArgs args = make_step_args(c, C, N_cur, N_next, dt, sum_out, deriv_out, C_next);
    args.out_exp = out_exp;
    // the sparse ring correction only reads C: it runs on a side stream beside the (register-bound, one CTA per
    // SM) factored kernel and joins before the closing kernel
    const bool side = args.P.n_ring > 0;
    if (side) {
        LDM_CUDA(cudaEventRecord(c->ev_fork, s));
        LDM_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        ldm::launch_ring(args, c->aux_stream);
        LDM_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
    }
    launch_integral(c, args, s);
    if (side) LDM_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    ldm::launch_finish_factored(args, s);
    c->launches += 2 + (side ? 1 : 0);
}

// ---- several species ----
size_t state_width(const ldm_context* c) { return c->species == 1 ? 1 + c->P() : (size_t)c->species + (size_t)c->species * c->species * c->P(); }

// the integrals of all pair densities (no closing): partial sums, ring correction, WC row dots -- the pairs' one-species
// handles are grouped by kernel shape and every group advances in one launch per kernel (blockIdx.z = pair)
void multi_build_groups(ldm_context* parent)
{
    for (ldm_context::PairGroup& g : parent->pair_groups) { cudaFree(g.d_args[0]); cudaFree(g.d_args[1]); }
    parent->pair_groups.clear();
    if (parent->multi_graph) { cudaGraphExecDestroy(parent->multi_graph); parent->multi_graph = nullptr; }   // it holds the old argument arrays
    std::vector<std::vector<ldm_context*>> members;
    for (ldm_context* h : parent->sub) {
        ensure_ring_operator(h);
        size_t at = parent->pair_groups.size();
        for (size_t k = 0; k < parent->pair_groups.size(); k++) {
            const ldm_context::PairGroup& cand = parent->pair_groups[k];
            if ((cand.shape.scan && h->plan.scan && cand.scan_nt == h->dev.scan_nt) ||
                (!cand.shape.scan && !h->plan.scan && cand.shape.bm == h->plan.bm && cand.shape.bn == h->plan.bn &&
                 cand.shape.tn == h->plan.tn)) at = k;
        }
        if (at == parent->pair_groups.size()) {
            parent->pair_groups.emplace_back();
            parent->pair_groups.back().shape = h->plan;
            parent->pair_groups.back().scan_nt = h->dev.scan_nt;
            members.emplace_back();
        }
        ldm_context::PairGroup& g = parent->pair_groups[at];
        members[at].push_back(h);
        g.members++;
        g.max_slabs = std::max(g.max_slabs, h->plan.slabs());
        g.max_rows = std::max(g.max_rows, h->dev.G);
        g.max_G = std::max(g.max_G, h->dev.G);
        g.max_A = std::max(g.max_A, h->dev.A);
        g.max_ring = std::max(g.max_ring, h->dev.n_ring);
        if (h->plan.scan) g.max_smem = std::max(g.max_smem, ldm::scan_smem_bytes(h->dev.G, h->dev.A, h->dev.scan_bn, h->dev.scan_nt, h->dev.scan_max_pieces));
    }
    for (size_t k = 0; k < parent->pair_groups.size(); k++)
        for (int parity = 0; parity < 2; parity++) {
            std::vector<ldm::StepArgs> host;
            for (ldm_context* h : members[k]) host.push_back(make_step_args(h, h->d_C[parity], nullptr, nullptr, 0.0, nullptr, nullptr, nullptr));
            void* p = nullptr;
            LDM_CUDA(cudaMalloc(&p, host.size() * sizeof(ldm::StepArgs)));
            LDM_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(ldm::StepArgs), cudaMemcpyHostToDevice));
            parent->pair_groups[k].d_args[parity] = static_cast<ldm::StepArgs*>(p);
        }
}

void multi_integrals(ldm_context* parent, cudaStream_t s)
{
    if (parent->pair_groups.empty()) multi_build_groups(parent);
    for (ldm_context::PairGroup& g : parent->pair_groups) {
        const ldm::StepArgs* args = g.d_args[parent->cur];
        const bool side = g.max_ring > 0;
        if (side) {
            LDM_CUDA(cudaEventRecord(parent->ev_fork, s));
            LDM_CUDA(cudaStreamWaitEvent(parent->aux_stream, parent->ev_fork, 0));
            ldm::launch_ring_batch(args, g.members, g.max_ring, parent->aux_stream);
            LDM_CUDA(cudaEventRecord(parent->ev_join, parent->aux_stream));
        }
        if (g.shape.scan) ldm::launch_scan_batch(args, g.members, g.scan_nt, g.max_slabs, g.max_rows, g.max_smem, s);
        else ldm::launch_factored_batch(g.shape, args, g.members, g.max_slabs, g.max_rows, g.max_G, g.max_A, s);
        if (side) LDM_CUDA(cudaStreamWaitEvent(s, parent->ev_join, 0));
        parent->launches += 1 + (side ? 1 : 0);
    }
}

// Equation::operator() for several species at the resident state, optionally with the clamped Euler update
void multi_rhs_and_update(ldm_context* c, double dt, bool want_derivs, bool update, cudaStream_t s)
{
    multi_integrals(c, s);
    const int n = c->species, cur = c->cur, nxt = cur ^ 1;
    ldm::MultiArgs a{};
    a.species = n; a.G = c->dev.G; a.extrapolation = c->dev.extrapolation;
    a.hh = c->dev.hh; a.dt = dt;
    for (int i = 0; i < n; i++) { a.b[i] = c->rate_b[i]; a.d[i] = c->rate_d[i]; }
    for (int i = 0; i < n * n; i++) a.dd[i] = c->rate_dd[i];
    a.N_cur = c->d_Nm + (size_t)cur * n;
    a.N_next = update ? c->d_Nm + (size_t)nxt * n : nullptr;
    a.dN_out = c->d_dN;
    for (int p = 0; p < n * n; p++) {
        ldm_context* h = c->sub[(size_t)p];
        ldm::MultiPair& m = a.pair[p];
        m.C = h->d_C[cur];
        m.C_next = update ? h->d_C[nxt] : nullptr;
        m.deriv_out = want_derivs ? h->d_deriv : nullptr;
        m.partial = h->d_partial; m.slabs = h->plan.slabs();
        m.ring = h->d_ring; m.wc_rows = h->d_wc_rows;
        m.out_weight = h->dev.out_weight; m.m_nodes = h->dev.m_nodes; m.w_nodes = h->dev.w_nodes;
        m.wc_outside = h->dev.wc_outside;
    }
    ldm::launch_finish_multi(a, s);
    c->launches += 1;
    if (update) c->cur = nxt;
}

// n_steps clamped Euler steps (Solver.cpp:577-582) of a multi-species handle at its resident state.  A step is three to
// five small launches on two streams; the launches of TWO consecutive steps -- both parities of the state ping-pong -- are
// captured once into a CUDA graph and replayed, the odd steps at either end run as ordinary launches (same kernels, same
// arguments: the same bits either way).
void multi_advance(ldm_context* c, int n_steps, double dt, cudaStream_t s)
{
    if (n_steps <= 0) return;
    if (c->pair_groups.empty()) multi_build_groups(c);                       // allocates: never inside a capture
    auto step = [&] { multi_rhs_and_update(c, dt, false, true, s); n_steps--; };
    if (c->cur & 1) step();                                                  // the graph starts at parity 0
    const bool fresh = c->multi_graph && c->multi_graph_dt == dt && c->multi_graph_extrapolation == c->dev.extrapolation;
    if (!fresh && n_steps >= 8 && tuning().multi_graph != 0) {
        if (c->multi_graph) { cudaGraphExecDestroy(c->multi_graph); c->multi_graph = nullptr; }
        step(); step();                                                      // first use of every kernel outside the capture
        const long long before = c->launches;
        cudaGraph_t graph = nullptr;
        LDM_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            multi_rhs_and_update(c, dt, false, true, s);
            multi_rhs_and_update(c, dt, false, true, s);
        } catch (...) {
            cudaStreamEndCapture(s, &graph);
            if (graph) cudaGraphDestroy(graph);
            c->cur = 0;
            throw;
        }
        LDM_CUDA(cudaStreamEndCapture(s, &graph));
        c->multi_graph_launches = c->launches - before;
        c->launches = before;
        const cudaError_t e = cudaGraphInstantiate(&c->multi_graph, graph, 0);
        cudaGraphDestroy(graph);
        LDM_CUDA(e);
        c->multi_graph_dt = dt;
        c->multi_graph_extrapolation = c->dev.extrapolation;
    }
    if (c->multi_graph && c->multi_graph_dt == dt && c->multi_graph_extrapolation == c->dev.extrapolation && tuning().multi_graph != 0)
        for (; n_steps >= 2; n_steps -= 2) {
            LDM_CUDA(cudaGraphLaunch(c->multi_graph, s));
            c->launches += c->multi_graph_launches;
        }
    while (n_steps > 0) step();
}

void multi_upload(ldm_context* c, const double* state, cudaStream_t s)
{
    const size_t P = c->P();
    const int n = c->species;
    c->cur = 0;
    LDM_CUDA(cudaMemcpyAsync(c->d_Nm, state, n * sizeof(double), cudaMemcpyHostToDevice, s));
    for (int p = 0; p < n * n; p++)
        LDM_CUDA(cudaMemcpyAsync(c->sub[(size_t)p]->d_C[0], state + n + (size_t)p * P, P * sizeof(double), cudaMemcpyHostToDevice, s));
}

void multi_download(ldm_context* c, double* state, cudaStream_t s)
{
    const size_t P = c->P();
    const int n = c->species;
    LDM_CUDA(cudaMemcpyAsync(state, c->d_Nm + (size_t)c->cur * n, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    for (int p = 0; p < n * n; p++)
        LDM_CUDA(cudaMemcpyAsync(state + n + (size_t)p * P, c->sub[(size_t)p]->d_C[c->cur], P * sizeof(double), cudaMemcpyDeviceToHost, s));
}

void multi_download_derivs(ldm_context* c, double* deriv, cudaStream_t s)
{
    const size_t P = c->P();
    const int n = c->species;
    LDM_CUDA(cudaMemcpyAsync(deriv, c->d_dN, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    for (int p = 0; p < n * n; p++)
        LDM_CUDA(cudaMemcpyAsync(deriv + n + (size_t)p * P, c->sub[(size_t)p]->d_deriv, P * sizeof(double), cudaMemcpyDeviceToHost, s));
}

// All Euler steps of a solve inside one cooperative launch (ldm_factored.cu persistent_kernel): possible when every row
// is one slab, on one rank, and the grid (G row CTAs + at least one spare CTA for the ring correction) is co-resident.
bool persistent_possible(ldm_context* c)
{
    if (!c->persistent_enabled || c->algo == LDM_ALGO_DIRECT || c->dev.world != 1) return false;
    if (c->dev.extrapolation == LDM_EXTRAPOLATION_EXPONENTIAL) return false;        // needs its own launches every step
    if (c->closure != LDM_CLOSURE_POWER1) return false;                             // likewise (ldm_closure.cu)
    const ldm::FactoredPlan& p = c->plan;
    if (p.slabs() != 1) return false;
    if (c->persistent_grid == 0) {
        ldm::PersistentArgs probe{};
        probe.step = make_step_args(c, c->d_C[0], c->d_N, c->d_N + 1, 0.0, nullptr, nullptr, c->d_C[1]);
        int max_grid = 0;
        const cudaError_t e = p.scan ? ldm::launch_persistent_scan(probe, 0, c->stream, &max_grid)
                                     : ldm::launch_persistent(p, probe, 0, c->stream, &max_grid);
        c->persistent_grid = (e == cudaSuccess && max_grid > 0) ? max_grid : -1;
        if (e != cudaSuccess) cudaGetLastError();
    }
    return c->persistent_grid >= c->dev.G + (p.scan ? 0 : 1);
}

// n_steps Euler steps from the resident state; history pointers (device) may be NULL.
// rhs_only_last (prefix-sum path): the last of the n_steps only evaluates the right-hand side at its state -- the derivative of
// the last point of a solve (S:582-583) inside the same launch -- and advances nothing.
void advance_persistent(ldm_context* c, int n_steps, double dt, long long first_point, double* d_obs, double* d_values, double* d_derivs,
                        bool rhs_only_last = false)
{
    if (n_steps <= 0) return;
    ensure_ring_operator(c);
    ldm::PersistentArgs args{};
    args.step = make_step_args(c, c->d_C[c->cur], c->d_N + c->cur, c->d_N + (c->cur ^ 1), dt, nullptr, nullptr, c->d_C[c->cur ^ 1]);
    args.C[0] = c->d_C[0]; args.C[1] = c->d_C[1];
    args.N = c->d_N;
    args.parity = c->cur;
    args.n_steps = n_steps;
    args.first_point = first_point;
    args.stride = (long long)(1 + c->P());
    args.observables = d_obs; args.values = d_values; args.derivs = d_derivs;
    args.wc_rows = c->d_wc_rows2; args.barrier = c->d_barrier;
    args.status = c->d_status;
    const int advanced = n_steps - (rhs_only_last && c->plan.scan ? 1 : 0);
    if (c->plan.scan) {
        // The arrival counter of the grid barrier is monotone and the ring hand-over slots are back at "not yet published" when a
        // launch has run to its end, so neither is reset per launch (two memsets in front of every launch cost 5 us of a
        // 20-step solve): only before the first launch, after one that timed out, and long before the counter could wrap.
        if (!c->d_ring2) { c->d_ring2 = c->alloc<double>(2 * c->P()); c->persistent_dirty = true; }
        if (c->persistent_dirty || c->barriers_done > (1u << 24)) {
            LDM_CUDA(cudaMemsetAsync(c->d_ring2, 0xFF, 2 * c->P() * sizeof(double), c->stream));      // every entry: not yet published
            LDM_CUDA(cudaMemsetAsync(c->d_barrier, 0, (1 + (size_t)c->dev.G) * sizeof(unsigned int), c->stream));
            c->barriers_done = 0;
            c->persistent_dirty = false;
        }
        args.barriers_done = c->barriers_done;
        args.last_is_rhs_only = rhs_only_last ? 1 : 0;
    }
    args.ring_pingpong = c->d_ring2;
    int device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (c->plan.scan) {
        // one CTA per moment-grid row, each closing its own row: no spare CTAs, one grid barrier per step
        LDM_CUDA(ldm::launch_persistent_scan(args, c->dev.G, c->stream, nullptr));
        c->barriers_done += 1u + (unsigned int)advanced;          // the prologue's, then one per step that advanced
    } else {
        const int grid = std::min(c->persistent_grid, std::max(sms, c->dev.G + 1));
        LDM_CUDA(ldm::launch_persistent(c->plan, args, grid, c->stream, nullptr));
    }
    c->launches += 1;
    c->cur ^= (advanced & 1);
}

// Which multi-step kernel, if any, carries the Euler loop of this handle: one CTA per row (rows of one chunk, G <= 128),
// the work-list kernel (prefix-sum path with rows of several chunks), or neither (one launch per kernel and step).
enum class Resident { None, Rows, Chunks };

Resident resident_mode(ldm_context* c)
{
    if (persistent_possible(c)) return Resident::Rows;
    // Rows of several chunks on ONE rank: one launch per kernel and step.  The work-list kernel (what the xi-sharded ranks run) pays
    // about 2 us per item for what kernel boundaries give for free -- an arrival with its fence per item, closings and ring shares
    // in line, waits for rows of the step before -- and measured 10 % slower at every size (G = 200: 58.2 against 51.6 us per
    // step, 256: 81.3 / 73.8, 333: 172 / 154, 512: 609 / 555; tools/chunked_modes.py); it is taken only where it is asked for.
    if (c->dev.world == 1 && c->work_list_on_one_rank && chunked_possible(c)) return Resident::Chunks;
    return Resident::None;
}

void advance_resident(ldm_context* c, Resident mode, int n_steps, double dt, double* d_obs, double* d_values, double* d_derivs,
                      bool rhs_only_last = false)
{
    if (mode == Resident::Rows) advance_persistent(c, n_steps, dt, 0, d_obs, d_values, d_derivs, rhs_only_last);
    else advance_chunked(c, n_steps, dt, c->stream, 0, d_obs, d_values, d_derivs);
}

void check_device_status(ldm_context* c)
{
    int status = 0;
    LDM_CUDA(cudaMemcpy(&status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (status == 0) return;
    cudaMemset(c->d_status, 0, sizeof(int));
    c->persistent_dirty = true;                     // arrival counter and hand-over slots of the persistent kernel: start over
    static const char* what[] = {"", "a grid barrier of the persistent kernel timed out", "a ring value was never handed over",
                                 "a rank never published its step (peer exchange timed out)"};
    throw CudaError(std::string("device status ") + std::to_string(status) + ": " + (status > 0 && status < 4 ? what[status] : "unknown") +
                    " -- the results of this call are not valid");
}

// Device-side history of ldm_solve: two segment buffers per handle, kept between calls (grown when a call needs more), so
// that a solve neither allocates inside its timed region nor holds the whole trajectory on the device -- one segment is
// copied out on the copy stream while the next one is being computed.
constexpr size_t kHistorySegmentBytes = 16u << 20;      // small enough that a 200-step solve of the 16 384-point grid overlaps (4 segments)

void ensure_history(ldm_context* c, size_t doubles)
{
    if (!c->copy_stream) {
        LDM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            LDM_CUDA(cudaEventCreateWithFlags(&c->ev_segment[b], cudaEventDisableTiming));
            LDM_CUDA(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
        }
    }
    if (doubles <= c->hist_capacity) return;
    LDM_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (int b = 0; b < 2; b++) {
        if (c->d_hist[b]) cudaFree(c->d_hist[b]);
        c->d_hist[b] = nullptr;
        LDM_CUDA(cudaMalloc((void**)&c->d_hist[b], doubles * sizeof(double)));
    }
    c->hist_capacity = doubles;
}

// The device's address of a host buffer that it can write (pinned and mapped memory), else NULL (pageable memory).
double* device_view_of_host(double* host)
{
    if (!host) return nullptr;
    cudaPointerAttributes attributes{};
    if (cudaPointerGetAttributes(&attributes, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attributes.type != cudaMemoryTypeHost || !attributes.devicePointer) return nullptr;
    return static_cast<double*>(attributes.devicePointer);
}

void step_resident(ldm_context* c, double dt, double* deriv_out)
{
    const int cur = c->cur, nxt = cur ^ 1;
    rhs_and_update(c, c->dev, c->d_C[cur], c->d_N + cur, c->d_N + nxt, dt, nullptr, deriv_out, c->d_C[nxt], c->stream);
    c->cur = nxt;
}

void upload_state(ldm_context* c, const double* state)
{
    c->cur = 0;
    LDM_CUDA(cudaMemcpyAsync(c->d_N, state, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LDM_CUDA(cudaMemcpyAsync(c->d_C[0], state + 1, c->P() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
}

}  // namespace ldm

using namespace ldm;

// What the reference asserts on its own tables (S:77 grid end = cutoff, S:88 K odd, S:303-304 quadrature extent within the
// table) plus what the index arithmetic of build_tables needs from a table that did not come from ldm_problem_generate.
// Empty string: fine.
static std::string problem_defect(const ldm_problem* p)
{
    auto finite = [](double v) { return std::isfinite(v); };
    if (p->K % 2 == 0) return "K (kernel-grid nodes) must be odd (Solver.cpp:88)";
    if (!(p->h > 0) || !finite(p->h)) return "KernelsGridStep must be positive";
    if (!(p->cutoff > 0) || !finite(p->cutoff)) return "CutoffDistance must be positive";
    auto extent_ok = [&](int m) { return m >= 2 && m <= p->K && (m == 2 || m % 2 == 1); };
    if (!extent_ok(p->mM)) return "MaxNumForM must be 2 or an odd number in [3, K] (Solver.cpp:125-132)";
    if (!extent_ok(p->mW)) return "MaxNumForW must be 2 or an odd number in [3, K] (Solver.cpp:125-132)";
    if (p->species > 1) {
        for (int i = 0; i < p->species; i++) if (!extent_ok(p->mM_of[i])) return "an entry of MaxNumForM is not 2 or an odd number in [3, K]";
        for (int i = 0; i < p->species * p->species; i++) if (!extent_ok(p->mW_of[i])) return "an entry of MaxNumForW is not 2 or an odd number in [3, K]";
    }
    if (p->moments_grid[0] != 0.0) return "MomentsGrid must start at 0";
    for (int k = 1; k < p->G; k++)
        if (!(p->moments_grid[k] > p->moments_grid[k - 1]) || !finite(p->moments_grid[k])) return "MomentsGrid must be strictly increasing";
    if (!(std::fabs(p->moments_grid[p->G - 1] - p->cutoff) < 1e-8)) return "MomentsGrid must end at the cutoff (Solver.cpp:77)";
    if (p->kernels_grid[0] != 0.0) return "KernelsGrid must start at 0";
    for (int k = 1; k < p->K; k++)
        if (!(p->kernels_grid[k] > p->kernels_grid[k - 1]) || !finite(p->kernels_grid[k])) return "KernelsGrid must be strictly increasing";
    return "";
}

extern "C" {

const char* ldm_last_error(void) { return g_error.c_str(); }
const char* ldm_version(void) { return "ldm_b200 0.1 (sm_100a)"; }

int ldm_problem_generate(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G, ldm_problem** out)
{
    if (!out) return fail(LDM_ERR_ARGUMENT, "out is NULL");
    return guarded([&] {
        auto owned = std::make_unique<ldm::OwnedProblem>();
        ldm::generate_problem(sigma_w, sigma_m, b, d, d_prime, forced_G, *owned);
        *out = &owned.release()->view;      // view is the first member: the pointer doubles as the owner
    });
}

int ldm_problem_generate_multi(int species, const double* sigma_w, const double* sigma_m, const double* b, const double* d,
                               const double* d_prime, int forced_G, ldm_problem** out)
{
    if (!out || !sigma_w || !sigma_m || !b || !d || !d_prime) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    return guarded([&] {
        auto owned = std::make_unique<ldm::OwnedProblem>();
        ldm::generate_problem_multi(species, sigma_w, sigma_m, b, d, d_prime, forced_G, *owned);
        *out = &owned.release()->view;
    });
}

void ldm_problem_free(ldm_problem* p)
{
    if (p) delete reinterpret_cast<ldm::OwnedProblem*>(p);
}

int ldm_problem_check_kernels(const ldm_problem* p, double* integral_w, double* integral_m)
{
    if (!p || !integral_w || !integral_m) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    ldm::check_kernels(*p, *integral_w, *integral_m);
    return LDM_OK;
}

int ldm_problem_kernel_structure(const ldm_problem* p, int* separable, double* max_rel_dev)
{
    if (!p || !separable || !max_rel_dev) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (p->species != 1 || p->G < 2 || p->K < 3 || !p->moments_grid || !p->kernels_grid || !p->M || !p->W)
        return fail(LDM_ERR_ARGUMENT, "incomplete problem description");
    if (const std::string defect = problem_defect(p); !defect.empty()) return fail(LDM_ERR_ARGUMENT, defect);
    return guarded([&] {
        ldm::Tables t;
        ldm::build_tables(*p, t);
        ldm::SeparableKernel k;
        ldm::detect_separable(t.offM, kSeparableTolerance, k);
        *separable = k.ok ? 1 : 0;
        *max_rel_dev = k.max_rel_dev;
    });
}

int ldm_tuning_set(const char* name, int value)
{
    if (!name) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    ldm::Tuning& t = tuning();
    const std::string key(name);
    if (key == "reset") { t = ldm::Tuning(); ldm::closure_force_tiles(-1); }
    else if (key == "closure_tiles") ldm::closure_force_tiles(value);
    else if (key == "scan_bn") t.scan_bn = value;
    else if (key == "scan_nt") t.scan_nt = value;
    else if (key == "factored_bm") t.factored_bm = value;
    else if (key == "factored_bn") t.factored_bn = value;
    else if (key == "factored_tn") t.factored_tn = value;
    else if (key == "persistent") t.persistent = value;
    else if (key == "scan") t.scan = value;
    else if (key == "chunk_grid") t.chunk_grid = value;
    else if (key == "batch_capacity") t.batch_capacity = value;
    else if (key == "item_order") t.item_order = value;
    else if (key == "multi_graph") t.multi_graph = value;
    else return fail(LDM_ERR_ARGUMENT, "unknown tuning key");
    return LDM_OK;
}

int ldm_create(const ldm_problem* p, int device, ldm_handle* out)
{
    if (!p || !out) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (p->species < 1 || p->species > LDM_MAX_SPECIES) return fail(LDM_ERR_UNSUPPORTED, "1 .. LDM_MAX_SPECIES species");
    if (p->G < 2 || p->K < 3 || !p->moments_grid || !p->kernels_grid || !p->M || !p->W)
        return fail(LDM_ERR_ARGUMENT, "incomplete problem description");
    if (p->species > 1 && (!p->mM_of || !p->mW_of || !p->sigma_m_of || !p->sigma_w_of || !p->b_of || !p->d_of || !p->d_prime_of))
        return fail(LDM_ERR_ARGUMENT, "a multi-species problem needs its per-species arrays (ldm_problem_generate_multi)");
    if (const std::string defect = problem_defect(p); !defect.empty()) return fail(LDM_ERR_ARGUMENT, defect);
    if (p->species > 1) {
        if (!p->mM_of || !p->mW_of || !p->sigma_m_of || !p->sigma_w_of || !p->b_of || !p->d_of || !p->d_prime_of)
            return fail(LDM_ERR_ARGUMENT, "a multi-species problem needs its per-species arrays (ldm_problem_generate_multi)");
        return guarded([&] {
            const int n = p->species;
            auto c = std::make_unique<ldm_context>();
            c->device = device;
            c->species = n;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    const ldm_problem view = ldm::species_view(*p, i, j);
                    ldm_handle h = nullptr;
                    if (ldm_create(&view, device, &h) != LDM_OK) throw CudaError(g_error);
                    c->sub.push_back(h);
                }
            LDM_CUDA(cudaSetDevice(device));
            LDM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            LDM_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
            LDM_CUDA(cudaEventCreate(&c->ev_begin));
            LDM_CUDA(cudaEventCreate(&c->ev_end));
            LDM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            LDM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
            c->dev = c->sub[0]->dev;                 // G, hh, extrapolation (the tables stay the pairs')
            c->algo = c->sub[0]->algo;
            for (ldm_context* h : c->sub)
                if (h->algo != c->algo) { h->algo = LDM_ALGO_FACTORED; h->plan = h->plan_dense; c->algo = LDM_ALGO_FACTORED; }
            if (c->algo == LDM_ALGO_FACTORED)
                for (ldm_context* h : c->sub) { h->algo = LDM_ALGO_FACTORED; h->plan = h->plan_dense; }
            for (int i = 0; i < n; i++) { c->rate_b[i] = p->b_of[i]; c->rate_d[i] = p->d_of[i]; }
            for (int i = 0; i < n * n; i++) c->rate_dd[i] = p->d_prime_of[i];
            c->d_Nm = c->alloc<double>(2 * (size_t)n);
            c->d_dN = c->alloc<double>((size_t)n);
            *out = c.release();
        });
    }
    return guarded([&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw CudaError("no CUDA device available -- libldm_b200 has no CPU fallback");
        if (device < 0 || device >= count) throw std::invalid_argument("device index out of range");
        auto c = std::make_unique<ldm_context>();
        c->device = device;
        LDM_CUDA(cudaSetDevice(device));
        LDM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        LDM_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        LDM_CUDA(cudaEventCreate(&c->ev_begin));
        LDM_CUDA(cudaEventCreate(&c->ev_end));
        LDM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        LDM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));

        ldm::build_tables(*p, c->tables);
        {   // what a closure over the w quadrature would be built from (ldm_set_closure): the caller's problem may be gone by then
            ldm::OwnedProblem& seed = c->w_seed;
            seed.moments_grid.assign(p->moments_grid, p->moments_grid + p->G);
            seed.kernels_grid.assign(p->kernels_grid, p->kernels_grid + p->K);
            seed.W.assign(p->W, p->W + (size_t)p->K * p->K);
            seed.view = *p;
            seed.view.species = 1;
            seed.view.moments_grid = seed.moments_grid.data();
            seed.view.kernels_grid = seed.kernels_grid.data();
            seed.view.W = seed.W.data();
            seed.view.M = seed.W.data();               // the w table in the m table's place: build_tables lays out ITS quadrature
            seed.view.mM = p->mW;
            seed.view.sigma_m = p->sigma_w;
            seed.view.mM_of = nullptr; seed.view.mW_of = nullptr; seed.view.sigma_m_of = nullptr; seed.view.sigma_w_of = nullptr;
            seed.view.b_of = nullptr; seed.view.d_of = nullptr; seed.view.d_prime_of = nullptr;
        }
        const ldm::Tables& t = c->tables;
        DevProblem& d = c->dev;
        d.G = t.G; d.n_off = t.offM.n_off; d.A = t.offM.A;
        d.world = 1; d.rank = 0; d.rows_per_rank = t.G;
        d.extrapolation = LDM_EXTRAPOLATION_SQUARED_N;
        d.cutoff = p->cutoff;
        for (int k = 1; k < 5; k++) d.tail_gap[k - 1] = p->G > k ? p->cutoff - p->moments_grid[p->G - 1 - k] : 1.0;   // S:397
        d.wc_out_w = c->upload(t.wc_out_w);
        d.wc_out_distance = c->upload(t.wc_out_distance);
        d.wc_out_n = (int)t.wc_out_w.size();
        d.hh = t.hh; d.b = t.b; d.d = t.d; d.dd = t.dd; d.s_max = t.s_max; d.wc_outside = t.wc_outside;
        d.s = c->upload(t.axisM.s);
        d.limit = c->upload(t.axisM.limit);
        d.weight = c->upload(t.axisM.weight);
        d.upper = c->upload(t.axisM.upper);
        {
            const int n_off = d.n_off, G = d.G;
            std::vector<int> ut((size_t)G * n_off);
            std::vector<double> wt((size_t)G * n_off);
            for (int n = 0; n < G; n++)
                for (int o = 0; o < n_off; o++) {
                    ut[(size_t)o * G + n] = t.axisM.upper[(size_t)n * n_off + o];
                    wt[(size_t)o * G + n] = t.axisM.weight[(size_t)n * n_off + o];
                }
            d.upper_t = c->upload(ut);
            d.weight_t = c->upload(wt);
        }
        d.alpha = c->upload(t.offM.alpha);
        d.quadrant = c->upload(t.offM.quadrant);
        d.m_nodes = c->upload(t.m_nodes);
        d.w_nodes = c->upload(t.w_nodes);
        d.wc_operator = c->upload(t.wc_operator);
        const size_t P = c->P();
        double* out_weight = c->alloc<double>(P);
        d.out_weight = out_weight;
        c->d_C[0] = c->alloc<double>(P + (size_t)d.G);     // + the tail of WC row dots (persistent_chunks_kernel)
        c->d_C[1] = c->alloc<double>(P + (size_t)d.G);
        c->d_N = c->alloc<double>(2);
        c->d_status = c->alloc<int>(1);
        c->d_flags = c->alloc<unsigned long long>(ldm::kFlagWords(d.G));
        LDM_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
        LDM_CUDA(cudaMemsetAsync(c->d_flags, 0, ldm::kFlagWords(d.G) * sizeof(unsigned long long), c->stream));
        c->d_scalars = c->alloc<double>(ldm::kScalars);
        c->d_sum = c->alloc<double>(P);
        c->d_deriv = c->alloc<double>(P);
        // structure of the m table: separable inside its disc (every table Properties::Generate builds) -> prefix sums
        d.qvec = nullptr; d.a_star = nullptr; d.beta_record = nullptr; d.scan_piece = nullptr; d.scan_chunk_begin = nullptr; d.scan_warp_begin = nullptr; d.scan_block_begin = nullptr; d.scan_chunk_order = nullptr; d.upper_c16 = nullptr; d.scan_max_pieces = 0; d.scan_bn = 0; d.scan_chunks = 0; d.scan_nt = 512; d.scan_live_rows = 0;
        {   // the y-side tables of the prefix-sum path, offsets two by two (ldm_scan_body.cuh y_table_index)
            const int n_even = (d.n_off + 1) & ~1;
            std::vector<unsigned short> u16((size_t)n_even * d.G, 0);
            std::vector<double> wp((size_t)n_even * d.G, 0.0);
            for (int n = 0; n < d.G; n++)
                for (int o = 0; o < d.n_off; o++) {
                    const size_t at = ((size_t)(o >> 1) * d.G + n) * 2 + (o & 1);
                    u16[at] = (unsigned short)t.axisM.upper[(size_t)n * d.n_off + o];
                    wp[at] = t.axisM.weight[(size_t)n * d.n_off + o];
                }
            d.upper_p16 = c->upload(u16);
            d.weight_p = c->upload(wp);
        }
        ldm::detect_separable(t.offM, kSeparableTolerance, c->separable);
        const int scan_bn = (c->separable.ok && d.G <= 65535) ? ldm::scan_pick_bn(d.G, d.A) : 0;
        c->plan_dense = ldm::plan_factored(d);
        if (scan_bn > 0 && tuning().scan != 0) {
            d.qvec = c->upload(c->separable.q);
            d.a_star = c->upload(c->separable.a_star);
            std::vector<int> records((size_t)d.A * 4, 0);
            for (int a = 0; a < d.A; a++) {
                const int lo = c->separable.beta_lo[a], hi = c->separable.beta_hi[a];
                const double q_first = lo < hi ? c->separable.q[(size_t)lo] : 0.0;
                records[4 * (size_t)a] = lo; records[4 * (size_t)a + 1] = hi;
                std::memcpy(&records[4 * (size_t)a + 2], &q_first, sizeof(double));         // low word, high word (little endian)
            }
            d.beta_record = reinterpret_cast<const int4*>(c->upload(records));
            d.scan_bn = scan_bn;
            d.scan_chunks = (d.G + scan_bn - 1) / scan_bn;
            d.scan_nt = ldm::scan_pick_threads(d.G, d.A, scan_bn);
            ldm::ScanPieces pieces;
            ldm::build_scan_pieces(t, scan_bn, d.scan_nt / 32, pieces);
            d.scan_piece = reinterpret_cast<const int4*>(c->upload(pieces.piece));
            d.scan_chunk_begin = c->upload(pieces.chunk_begin);
            d.scan_warp_begin = c->upload(pieces.warp_begin);
            d.scan_block_begin = c->upload(pieces.block_begin);
            d.scan_max_pieces = pieces.max_pieces;
            if (d.scan_chunks > 1) {
                // per chunk: where a pair lies in a row of the chunk's H tile (lead columns included), clamped to the zero columns
                const int n_even = (d.n_off + 1) & ~1, lead = ldm::scan_lead_columns();
                std::vector<unsigned short> pos((size_t)d.scan_chunks * n_even * d.G, 0);
                for (int ch = 0; ch < d.scan_chunks; ch++)
                    for (int n = 0; n < d.G; n++)
                        for (int o = 0; o < d.n_off; o++) {
                            const int at_column = t.axisM.upper[(size_t)n * d.n_off + o] - ch * scan_bn;     // the pair reads columns at_column - 1, at_column
                            const int clamped = std::min(std::max(at_column, -1), scan_bn + 1);
                            pos[(size_t)ch * n_even * d.G + ((size_t)(o >> 1) * d.G + n) * 2 + (o & 1)] = (unsigned short)(lead - 1 + clamped);
                        }
                d.upper_c16 = c->upload(pos);
            }
            {
                std::vector<long long> work((size_t)d.scan_chunks, 0);
                for (int ch = 0; ch < d.scan_chunks; ch++)
                    for (int pi = pieces.chunk_begin[(size_t)ch]; pi < pieces.chunk_begin[(size_t)ch + 1]; pi++)
                        work[(size_t)ch] += pieces.piece[4 * (size_t)pi + 2] - pieces.piece[4 * (size_t)pi + 1];
                std::vector<int> order((size_t)d.scan_chunks);
                for (int ch = 0; ch < d.scan_chunks; ch++) order[(size_t)ch] = ch;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return work[(size_t)a] > work[(size_t)b]; });
                d.scan_chunk_order = c->upload(order);
                c->chunk_order_host = order;
            }
            d.scan_live_rows = 0;
            while (d.scan_live_rows < d.A && c->separable.a_star[(size_t)d.scan_live_rows] >= 0) d.scan_live_rows++;
            c->plan_scan = c->plan_dense;
            c->plan_scan.scan = true;
            c->plan_scan.tiles = 1; c->plan_scan.chunks = d.scan_chunks;
            c->scan_available = true;
        }
        c->plan = c->scan_available ? c->plan_scan : c->plan_dense;
        c->algo = c->scan_available ? LDM_ALGO_PREFIX : LDM_ALGO_FACTORED;
        c->d_partial = c->alloc<double>(P * (size_t)std::max(c->plan_dense.slabs(), c->scan_available ? c->plan_scan.slabs() : 1));
        c->d_ring = c->alloc<double>(P);
        c->d_wc_rows = c->alloc<double>((size_t)d.G);
        c->d_wc_rows2 = c->alloc<double>(2 * (size_t)d.G);
        c->d_barrier = c->alloc<unsigned int>(1 + (size_t)d.G);
        if (tuning().persistent >= 0) c->persistent_enabled = tuning().persistent != 0;
        c->work_list_on_one_rank = tuning().persistent == 1;
        c->chunk_grid_limit = tuning().chunk_grid;
        d.n_ring = 0;
        LDM_CUDA(cudaMemsetAsync(c->d_C[0], 0, P * sizeof(double), c->stream));
        LDM_CUDA(cudaMemsetAsync(c->d_C[1], 0, P * sizeof(double), c->stream));
        LDM_CUDA(cudaMemsetAsync(c->d_N, 0, 2 * sizeof(double), c->stream));
        // time-invariant: the m-weight of every node's samples that fall beyond the cutoff (they see N^2, S:369-373)
        ldm::launch_sample_sum(d, ldm::SAMPLES_OUT_WEIGHT, c->d_C[0], out_weight, c->stream);
        c->launches += 1;
        LDM_CUDA(cudaGetLastError());
        // the ring-correction operator of the unsharded problem, here rather than at the first step: callers create the
        // entries of a sweep from several host threads, and this host-side build is most of an entry's set-up cost
        // (ldm_shard_configure rebuilds it for the rank's rows)
        ensure_ring_operator(c.get());
        LDM_CUDA(cudaStreamSynchronize(c->stream));
        *out = c.release();
    });
}

void ldm_destroy(ldm_handle h) { delete h; }

int ldm_set_algorithm(ldm_handle h, int algo)
{
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    if (algo == LDM_ALGO_AUTO) {
        bool all = h->species == 1 ? h->scan_available : true;
        for (ldm_context* c : h->sub) all = all && c->scan_available;
        algo = all ? LDM_ALGO_PREFIX : LDM_ALGO_FACTORED;
    }
    if (algo != LDM_ALGO_DIRECT && algo != LDM_ALGO_FACTORED && algo != LDM_ALGO_PREFIX) return fail(LDM_ERR_ARGUMENT, "unknown algorithm");
    if (h->species != 1) {
        if (algo == LDM_ALGO_DIRECT) return fail(LDM_ERR_UNSUPPORTED, "several species run on the factored paths only");
        for (ldm_context* c : h->sub)
            if (algo == LDM_ALGO_PREFIX && !c->scan_available)
                return fail(LDM_ERR_UNSUPPORTED, "an m table is not separable inside its disc: prefix sums do not apply");
        for (ldm_context* c : h->sub) { c->algo = algo; c->plan = algo == LDM_ALGO_PREFIX ? c->plan_scan : c->plan_dense; }
        h->algo = algo;
        cudaSetDevice(h->device);
        for (ldm_context::PairGroup& g : h->pair_groups) { cudaFree(g.d_args[0]); cudaFree(g.d_args[1]); }
        h->pair_groups.clear();            // rebuilt for the new kernel shapes at the next right-hand side
        return LDM_OK;
    }
    if (algo == LDM_ALGO_PREFIX && !h->scan_available)
        return fail(LDM_ERR_UNSUPPORTED, "the m table is not separable inside its disc: prefix sums do not apply");
    h->algo = algo;
    if (algo != LDM_ALGO_DIRECT) {
        h->plan = algo == LDM_ALGO_PREFIX ? h->plan_scan : h->plan_dense;
        h->persistent_grid = 0;          // co-residency differs between the two kernels: query again
    }
    return LDM_OK;
}

int ldm_get_algorithm(ldm_handle h) { return h ? h->algo : LDM_ERR_ARGUMENT; }

int ldm_set_extrapolation(ldm_handle h, int type)
{
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    if (type == LDM_EXTRAPOLATION_EXPONENTIAL && (h->species != 1 || h->dev.G < 5))
        return fail(LDM_ERR_UNSUPPORTED, "Exponential extrapolation is built for one species on grids of at least 5 nodes (it reads the last 5 of row 0, Solver.cpp:394-401)");
    if (type < LDM_EXTRAPOLATION_ZERO || type > LDM_EXTRAPOLATION_EXPONENTIAL) return fail(LDM_ERR_ARGUMENT, "unknown extrapolation type");
    if (type != LDM_EXTRAPOLATION_SQUARED_N && h->closure != LDM_CLOSURE_POWER1)
        return fail(LDM_ERR_UNSUPPORTED, "the closures over the w quadrature are built for SquaredN extrapolation (ldm_set_closure)");
    h->dev.extrapolation = type;
    return LDM_OK;
}

int ldm_count_points_rk4(double t_max, double dt)
{
    if (!(dt > 0)) return 1;
    int n = 1;
    double time = 0;
    while (time < t_max) { time += dt / 2; time += dt / 2; n++; }      // S:627-635
    return n;
}

int ldm_solve_rk4(ldm_handle h, const double* init, double dt, double t_max, double* times, double* values, double* derivs, int* n_points)
{
    if (!h || !init) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (!(dt > 0)) return fail(LDM_ERR_ARGUMENT, "step must be positive");
    if (h->species == 1 && h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the Runge-Kutta solver runs on one rank");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        const int n = ldm_count_points_rk4(t_max, dt);
        const size_t P = h->P(), width = state_width(h);
        const int species = h->species;
        if (n_points) *n_points = n;
        if (times) {
            double time = 0;
            for (int k = 0; k < n; k++) { times[k] = time; if (k + 1 < n) { time += dt / 2; time += dt / 2; } }
        }
        // device: the value [N | C], one stage state, four derivatives -- each laid out like the reference's Vector
        std::vector<void*> scratch;
        auto scratch_alloc = [&](size_t count) {
            void* p = nullptr;
            LDM_CUDA(cudaMalloc(&p, count * sizeof(double)));
            scratch.push_back(p);
            return static_cast<double*>(p);
        };
        try {
            double* value[2] = {scratch_alloc(width), scratch_alloc(width)};
            double* stage = scratch_alloc(width);
            double* c[4] = {scratch_alloc(width), scratch_alloc(width), scratch_alloc(width), scratch_alloc(width)};
            double* d_values = values ? scratch_alloc((size_t)n * width) : nullptr;
            double* d_derivs = derivs ? scratch_alloc((size_t)n * width) : nullptr;
            cudaStream_t s = h->stream;
            LDM_CUDA(cudaMemcpyAsync(value[0], init, width * sizeof(double), cudaMemcpyHostToDevice, s));
            // Equation::operator() at a device state [N | C]: dC straight into out + 1, dN from the step scalars.  Several
            // species: the state is dealt out to the pairs' handles, evaluated there, and the derivative gathered again
            auto rhs = [&](const double* state, double* out) {
                if (species == 1) {
                    rhs_and_update(h, h->dev, state + 1, state, nullptr, 0.0, nullptr, out + 1, nullptr, s);
                    LDM_CUDA(cudaMemcpyAsync(out, h->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
                    return;
                }
                h->cur = 0;
                LDM_CUDA(cudaMemcpyAsync(h->d_Nm, state, species * sizeof(double), cudaMemcpyDeviceToDevice, s));
                for (int p = 0; p < species * species; p++)
                    LDM_CUDA(cudaMemcpyAsync(h->sub[(size_t)p]->d_C[0], state + species + (size_t)p * P, P * sizeof(double), cudaMemcpyDeviceToDevice, s));
                multi_rhs_and_update(h, 0.0, true, false, s);
                LDM_CUDA(cudaMemcpyAsync(out, h->d_dN, species * sizeof(double), cudaMemcpyDeviceToDevice, s));
                for (int p = 0; p < species * species; p++)
                    LDM_CUDA(cudaMemcpyAsync(out + species + (size_t)p * P, h->sub[(size_t)p]->d_deriv, P * sizeof(double), cudaMemcpyDeviceToDevice, s));
            };
            LDM_CUDA(cudaEventRecord(h->ev_begin, s));
            int cur = 0;
            for (int k = 0; k + 1 < n; k++) {
                if (d_values) LDM_CUDA(cudaMemcpyAsync(d_values + (size_t)k * width, value[cur], width * sizeof(double), cudaMemcpyDeviceToDevice, s));
                rhs(value[cur], c[0]);                                                               // S:629, S:631
                if (d_derivs) LDM_CUDA(cudaMemcpyAsync(d_derivs + (size_t)k * width, c[0], width * sizeof(double), cudaMemcpyDeviceToDevice, s));
                ldm::launch_rk4_stage((int)width, value[cur], c[0], dt, 2.0, stage, s);              // S:633
                rhs(stage, c[1]);
                ldm::launch_rk4_stage((int)width, value[cur], c[1], dt, 2.0, stage, s);              // S:634
                rhs(stage, c[2]);
                ldm::launch_rk4_stage((int)width, value[cur], c[2], dt, 1.0, stage, s);              // S:636
                rhs(stage, c[3]);
                ldm::launch_rk4_close((int)width, value[cur], c[0], c[1], c[2], c[3], dt, value[cur ^ 1], s);   // S:637-645
                h->launches += 4;
                cur ^= 1;
            }
            if (d_values) LDM_CUDA(cudaMemcpyAsync(d_values + (size_t)(n - 1) * width, value[cur], width * sizeof(double), cudaMemcpyDeviceToDevice, s));
            if (d_derivs) {
                rhs(value[cur], c[0]);                                                               // S:652
                LDM_CUDA(cudaMemcpyAsync(d_derivs + (size_t)(n - 1) * width, c[0], width * sizeof(double), cudaMemcpyDeviceToDevice, s));
            }
            LDM_CUDA(cudaEventRecord(h->ev_end, s));
            LDM_CUDA(cudaGetLastError());
            // the handle's resident state follows the solve
            h->cur = 0;
            if (species == 1) {
                LDM_CUDA(cudaMemcpyAsync(h->d_N, value[cur], sizeof(double), cudaMemcpyDeviceToDevice, s));
                LDM_CUDA(cudaMemcpyAsync(h->d_C[0], value[cur] + 1, P * sizeof(double), cudaMemcpyDeviceToDevice, s));
            } else {
                LDM_CUDA(cudaMemcpyAsync(h->d_Nm, value[cur], species * sizeof(double), cudaMemcpyDeviceToDevice, s));
                for (int p = 0; p < species * species; p++)
                    LDM_CUDA(cudaMemcpyAsync(h->sub[(size_t)p]->d_C[0], value[cur] + species + (size_t)p * P, P * sizeof(double), cudaMemcpyDeviceToDevice, s));
            }
            if (values) LDM_CUDA(cudaMemcpyAsync(values, d_values, (size_t)n * width * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (derivs) LDM_CUDA(cudaMemcpyAsync(derivs, d_derivs, (size_t)n * width * sizeof(double), cudaMemcpyDeviceToHost, s));
            LDM_CUDA(cudaStreamSynchronize(s));
            LDM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev_begin, h->ev_end));
        } catch (...) {
            for (void* p : scratch) cudaFree(p);
            throw;
        }
        for (void* p : scratch) cudaFree(p);
    });
}

int ldm_rhs(ldm_handle h, const double* state, double* deriv)
{
    if (!h || !state || !deriv) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1)
        return guarded([&] {
            LDM_CUDA(cudaSetDevice(h->device));
            multi_upload(h, state, h->stream);
            multi_rhs_and_update(h, 0.0, true, false, h->stream);
            LDM_CUDA(cudaGetLastError());
            multi_download_derivs(h, deriv, h->stream);
            LDM_CUDA(cudaStreamSynchronize(h->stream));
        });
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        upload_state(h, state);
        rhs_and_update(h, h->dev, h->d_C[0], h->d_N, nullptr, 0.0, nullptr, h->d_deriv, nullptr, h->stream);
        LDM_CUDA(cudaGetLastError());
        LDM_CUDA(cudaMemcpyAsync(deriv, h->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        LDM_CUDA(cudaMemcpyAsync(deriv + 1, h->d_deriv, h->P() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        LDM_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int ldm_integrals(ldm_handle h, const double* state, double* WC, double* MC)
{
    if (!h || !state || !WC) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        upload_state(h, state);
        rhs_and_update(h, h->dev, h->d_C[0], h->d_N, nullptr, 0.0, h->d_sum, nullptr, nullptr, h->stream);
        LDM_CUDA(cudaGetLastError());
        LDM_CUDA(cudaMemcpyAsync(WC, h->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        std::vector<double> sum;
        if (MC) {
            sum.resize(h->P());
            LDM_CUDA(cudaMemcpyAsync(sum.data(), h->d_sum, h->P() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        LDM_CUDA(cudaStreamSynchronize(h->stream));
        if (MC) {   // the same closing expression as the pointwise kernel
            std::vector<double> ow(h->P());
            LDM_CUDA(cudaMemcpy(ow.data(), h->dev.out_weight, h->P() * sizeof(double), cudaMemcpyDeviceToHost));
            const double N = state[0], asymptote = N * N, last = state[h->dev.G];        // outside_value() / outside_sum() of ldm_pointwise.cuh
            if (h->dev.extrapolation == LDM_EXTRAPOLATION_EXPONENTIAL) {
                std::vector<double> ex(h->P());
                double sc[ldm::kScalars];
                LDM_CUDA(cudaMemcpy(ex.data(), h->d_out_exp, h->P() * sizeof(double), cudaMemcpyDeviceToHost));
                LDM_CUDA(cudaMemcpy(sc, h->d_scalars, sizeof sc, cudaMemcpyDeviceToHost));
                for (size_t q = 0; q < h->P(); q++) MC[q] = (sum[q] + (asymptote * ow[q] + sc[3] * ex[q])) * h->dev.hh;
            } else {
                const double beyond = (h->dev.extrapolation == LDM_EXTRAPOLATION_SQUARED_N || std::fabs(last - asymptote) < 1e-8) ? asymptote
                                      : h->dev.extrapolation == LDM_EXTRAPOLATION_ZERO ? 0.0 : state[1];
                for (size_t q = 0; q < h->P(); q++) MC[q] = (sum[q] + beyond * ow[q]) * h->dev.hh;
            }
        }
    });
}

int ldm_kernels_at_nodes(ldm_handle h, double* m_at_nodes, double* w_at_nodes)
{
    if (!h || !m_at_nodes || !w_at_nodes) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1) return fail(LDM_ERR_UNSUPPORTED, "not built for several species (ldm.h: ldm_problem_generate_multi)");
    std::memcpy(m_at_nodes, h->tables.m_nodes.data(), h->P() * sizeof(double));
    std::memcpy(w_at_nodes, h->tables.w_nodes.data(), h->P() * sizeof(double));
    return LDM_OK;
}

int ldm_count_points(double t_max, double dt)
{
    if (!(dt > 0)) return 1;
    int n = 1;
    double time = 0;
    time += dt;                                   // S:574
    while (time < t_max) { n++; time += dt; }     // S:575, S:588
    return n;
}

int ldm_solve(ldm_handle h, const double* init, double dt, double t_max, int record, double* times, double* values,
              double* derivs, int* n_points)
{
    if (!h || !init) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (record < LDM_RECORD_LAST || record > LDM_RECORD_ALL) return fail(LDM_ERR_ARGUMENT, "unknown record mode");
    if (!(dt > 0)) return fail(LDM_ERR_ARGUMENT, "step must be positive");
    if (record != LDM_RECORD_LAST && !values) return fail(LDM_ERR_ARGUMENT, "values is NULL");
    if (record == LDM_RECORD_ALL && !derivs) return fail(LDM_ERR_ARGUMENT, "derivs is NULL");
    if (h->species != 1) {
        if (record == LDM_RECORD_OBSERVABLES)
            return fail(LDM_ERR_UNSUPPORTED, "the observables of Experiment_14_4 (N and C[0][G-1]) are defined for one species");
        return guarded([&] {
            LDM_CUDA(cudaSetDevice(h->device));
            const int n = ldm_count_points(t_max, dt);
            const size_t width = state_width(h);
            if (n_points) *n_points = n;
            if (times) {
                double time = 0;
                for (int k = 0; k < n; k++) { times[k] = time; time += dt; }
            }
            cudaStream_t s = h->stream;
            multi_upload(h, init, s);
            LDM_CUDA(cudaEventRecord(h->ev_begin, s));
            if (record == LDM_RECORD_LAST) {                                 // nothing leaves the device on the way: replayed as a graph
                multi_advance(h, n - 1, dt, s);
                if (derivs) multi_rhs_and_update(h, dt, true, false, s);
            }
            for (int k = 0; k < n && record != LDM_RECORD_LAST; k++) {
                if (record >= LDM_RECORD_VALUES) multi_download(h, values + (size_t)k * width, s);
                const bool last = (k == n - 1);
                const bool want_deriv = (record == LDM_RECORD_ALL) || (last && derivs);
                if (last && !want_deriv) break;
                multi_rhs_and_update(h, dt, want_deriv, !last, s);       // S:567 / S:577-582
                if (record == LDM_RECORD_ALL) multi_download_derivs(h, derivs + (size_t)k * width, s);
            }
            LDM_CUDA(cudaEventRecord(h->ev_end, s));
            LDM_CUDA(cudaGetLastError());
            if (record == LDM_RECORD_LAST) {
                if (values) multi_download(h, values, s);
                if (derivs) multi_download_derivs(h, derivs, s);
            }
            LDM_CUDA(cudaStreamSynchronize(s));
            LDM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev_begin, h->ev_end));
        });
    }
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        const int n = ldm_count_points(t_max, dt);
        const size_t P = h->P(), width = 1 + P;
        if (n_points) *n_points = n;
        if (times) {
            double time = 0;
            for (int k = 0; k < n; k++) { times[k] = time; time += dt; }   // the accumulated clock of S:565-588
        }
        cudaStream_t s = h->stream;
        upload_state(h, init);
        const Resident mode = resident_mode(h);
        // History leaves the device segment by segment: `seg_len` steps per segment, two buffers, the copy of one segment
        // under the computation of the next.  Per step a segment holds the observables recorded when the step begins
        // (OBSERVABLES), or the state behind the step (VALUES) and the derivative that produced it (ALL).
        const size_t per_step = record == LDM_RECORD_OBSERVABLES ? 2 : record == LDM_RECORD_VALUES ? width : record == LDM_RECORD_ALL ? 2 * width : 0;
        int seg_len = std::max(1, n - 1);
        // Host buffers that the device can address (pinned and mapped: cudaHostAlloc, cudaHostRegister -- what a caller who cares
        // about transfer time hands over anyway): the persistent kernel stores every point of the history straight into them, in
        // ONE launch for the whole solve -- no device-side history, no copies behind the last step, nothing to wait for but the
        // kernel.  262 KB per step at 16 384 points are 22 GB/s of posted writes, under what the link carries.
        double* const direct_values = per_step != 0 && mode != Resident::None ? device_view_of_host(values) : nullptr;
        double* const direct_derivs = record == LDM_RECORD_ALL && direct_values ? device_view_of_host(derivs) : nullptr;
        const bool direct = direct_values && (record != LDM_RECORD_ALL || direct_derivs) && n > 1;
        if (per_step != 0 && !direct) {
            const size_t longest = std::max<size_t>(1, kHistorySegmentBytes / (per_step * sizeof(double)));   // steps of a full segment
            seg_len = (int)std::min<size_t>((size_t)seg_len, longest);
            // sized for the longest segment this record mode can ask for (at most 4096 steps' worth), not for this call: a short
            // solve followed by a long one must not re-allocate (measured: a 20-step warm-up in front of a 200-step solve put 4 ms
            // of cudaFree / cudaMalloc into the second call -- 36 instead of 15 us per step end to end)
            ensure_history(h, std::max((size_t)seg_len, std::min<size_t>(longest, 4096)) * per_step);
        }
        LDM_CUDA(cudaEventRecord(h->ev_begin, s));
        // ... and the derivative at the last point (S:582-583) as one more step of the same launch that advances nothing
        const bool fold_last = direct && record == LDM_RECORD_ALL && mode == Resident::Rows && h->plan.scan;
        if (direct)
            advance_resident(h, mode, fold_last ? n : n - 1, dt, record == LDM_RECORD_OBSERVABLES ? direct_values : nullptr,
                             record >= LDM_RECORD_VALUES ? direct_values : nullptr, direct_derivs, fold_last);
        // Values[0] (S:568): a host copy, made while the device is already at work (the kernels write from Values[1] on)
        if (record >= LDM_RECORD_VALUES && direct) std::memcpy(values, init, width * sizeof(double));
        for (int k0 = 0, seg = 0; k0 < n - 1 && !direct; k0 += seg_len, seg++) {
            const int len = std::min(seg_len, n - 1 - k0), b = seg & 1;
            double* buf = per_step != 0 ? h->d_hist[b] : nullptr;
            if (buf && seg >= 2) LDM_CUDA(cudaStreamWaitEvent(s, h->ev_copied[b], 0));          // the buffer's previous segment has left
            double* d_obs = record == LDM_RECORD_OBSERVABLES ? buf : nullptr;
            double* d_val = record >= LDM_RECORD_VALUES ? buf : nullptr;                         // [len][width]: the states behind the steps
            double* d_der = record == LDM_RECORD_ALL ? buf + (size_t)len * width : nullptr;      // [len][width]
            if (mode != Resident::None) {
                // all steps of the segment inside ONE launch; the kernel records the history itself (it stores the state
                // behind step j at values + (j + 1) * width, hence the shifted base)
                advance_resident(h, mode, len, dt, d_obs, d_val ? d_val - width : nullptr, d_der);
            } else {
                for (int j = 0; j < len; j++) {
                    if (d_obs) {
                        ldm::launch_record_observables(h->dev, h->d_N + h->cur, h->d_C[h->cur], d_obs + (size_t)j * 2, s);
                        h->launches += 1;
                    }
                    step_resident(h, dt, d_der ? h->d_deriv : nullptr);                          // S:567 / S:577-582
                    if (d_val) {
                        LDM_CUDA(cudaMemcpyAsync(d_val + (size_t)j * width, h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToDevice, s));
                        LDM_CUDA(cudaMemcpyAsync(d_val + (size_t)j * width + 1, h->d_C[h->cur], P * sizeof(double), cudaMemcpyDeviceToDevice, s));
                    }
                    if (d_der) {
                        LDM_CUDA(cudaMemcpyAsync(d_der + (size_t)j * width, h->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
                        LDM_CUDA(cudaMemcpyAsync(d_der + (size_t)j * width + 1, h->d_deriv, P * sizeof(double), cudaMemcpyDeviceToDevice, s));
                    }
                }
            }
            if (buf) {
                LDM_CUDA(cudaEventRecord(h->ev_segment[b], s));
                LDM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_segment[b], 0));
                if (d_obs) LDM_CUDA(cudaMemcpyAsync(values + (size_t)k0 * 2, d_obs, (size_t)len * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
                if (d_val) LDM_CUDA(cudaMemcpyAsync(values + (size_t)(k0 + 1) * width, d_val, (size_t)len * width * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
                if (d_der) LDM_CUDA(cudaMemcpyAsync(derivs + (size_t)k0 * width, d_der, (size_t)len * width * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
                LDM_CUDA(cudaEventRecord(h->ev_copied[b], h->copy_stream));
            }
        }
        if (record >= LDM_RECORD_VALUES && !direct) std::memcpy(values, init, width * sizeof(double));   // Values[0] (S:568)
        // the last point: its observables; and the reference evaluates the equation once more there (S:582-583)
        if (record == LDM_RECORD_OBSERVABLES) {
            ldm::launch_record_observables(h->dev, h->d_N + h->cur, h->d_C[h->cur], h->d_scalars + 5, s);
            h->launches += 1;
            LDM_CUDA(cudaMemcpyAsync(values + (size_t)(n - 1) * 2, h->d_scalars + 5, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if ((record == LDM_RECORD_ALL || derivs) && !fold_last) {
            rhs_and_update(h, h->dev, h->d_C[h->cur], h->d_N + h->cur, nullptr, dt, nullptr, h->d_deriv, nullptr, s);
            double* out = record == LDM_RECORD_ALL ? derivs + (size_t)(n - 1) * width : derivs;
            LDM_CUDA(cudaMemcpyAsync(out, h->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
            LDM_CUDA(cudaMemcpyAsync(out + 1, h->d_deriv, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        LDM_CUDA(cudaEventRecord(h->ev_end, s));
        LDM_CUDA(cudaGetLastError());
        if (record == LDM_RECORD_LAST && values) {
            LDM_CUDA(cudaMemcpyAsync(values, h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToHost, s));
            LDM_CUDA(cudaMemcpyAsync(values + 1, h->d_C[h->cur], P * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if (!h->h_status) LDM_CUDA(cudaHostAlloc((void**)&h->h_status, sizeof(int), cudaHostAllocDefault));
        LDM_CUDA(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, s));     // the status word, behind the work
        LDM_CUDA(cudaStreamSynchronize(s));
        if (h->copy_stream) LDM_CUDA(cudaStreamSynchronize(h->copy_stream));
        if (*h->h_status != 0) check_device_status(h);
        LDM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev_begin, h->ev_end));
    });
}

int ldm_state_upload(ldm_handle h, const double* state)
{
    if (!h || !state) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species == 1 && h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        if (h->species != 1) multi_upload(h, state, h->stream);
        else upload_state(h, state);
        LDM_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int ldm_state_download(ldm_handle h, double* state)
{
    if (!h || !state) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (h->species != 1)
        return guarded([&] {
            LDM_CUDA(cudaSetDevice(h->device));
            multi_download(h, state, h->stream);
            LDM_CUDA(cudaStreamSynchronize(h->stream));
        });
    if (h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        LDM_CUDA(cudaMemcpyAsync(state, h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        LDM_CUDA(cudaMemcpyAsync(state + 1, h->d_C[h->cur], h->P() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        LDM_CUDA(cudaStreamSynchronize(h->stream));
        check_device_status(h);
    });
}

int ldm_advance(ldm_handle h, int n_steps, double dt)
{
    if (!h || n_steps < 0) return fail(LDM_ERR_ARGUMENT, "bad argument");
    if (h->species == 1 && h->dev.world != 1) return fail(LDM_ERR_UNSUPPORTED, "the handle is configured as one rank of several (ldm_shard_configure): use the ldm_shard_* entry points, or ldm_shard_configure(h, 0, 1) to go back");
    if (h->species != 1)
        return guarded([&] {
            LDM_CUDA(cudaSetDevice(h->device));
            LDM_CUDA(cudaEventRecord(h->ev_begin, h->stream));
            multi_advance(h, n_steps, dt, h->stream);
            LDM_CUDA(cudaEventRecord(h->ev_end, h->stream));
            LDM_CUDA(cudaGetLastError());
        });
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        const Resident mode = resident_mode(h);
        LDM_CUDA(cudaEventRecord(h->ev_begin, h->stream));
        if (mode != Resident::None) advance_resident(h, mode, n_steps, dt, nullptr, nullptr, nullptr);
        else for (int k = 0; k < n_steps; k++) step_resident(h, dt, nullptr);
        LDM_CUDA(cudaEventRecord(h->ev_end, h->stream));
        LDM_CUDA(cudaGetLastError());
    });
}

int ldm_synchronize(ldm_handle h)
{
    if (!h) return fail(LDM_ERR_ARGUMENT, "NULL handle");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        LDM_CUDA(cudaStreamSynchronize(h->stream));
        if (h->species == 1) check_device_status(h);
    });
}

int ldm_last_elapsed_ms(ldm_handle h, float* ms)
{
    if (!h || !ms) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    return guarded([&] {
        LDM_CUDA(cudaSetDevice(h->device));
        LDM_CUDA(cudaEventSynchronize(h->ev_end));
        LDM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev_begin, h->ev_end));
        *ms = h->last_ms;
    });
}

long long ldm_launch_count(ldm_handle h) { return h ? h->launches : 0; }

}  // extern "C"

namespace {

// The members of a batch that live on ONE device: all parameter sets of one CTA shape advance in one launch per kernel
// (blockIdx.z is the set).  Runs (points - 1) Euler steps and returns when the device is done.
// inits / lasts (may be NULL): the members' initial states in, their last states out -- through ONE pinned staging buffer and
// one host <-> device copy each way for the whole slice (two small copies per member from pageable memory cost a batch of
// 128 sets 6 ms, an eighth of its 200 steps)
void advance_batch_on_device(const std::vector<ldm_context*>& handles, int points, double dt, const double* const* inits = nullptr,
                             double* const* lasts = nullptr)
{
    const int n = (int)handles.size();
    if (n == 0) return;
        // All parameter sets of one tile shape advance in ONE launch per kernel: blockIdx.z is the set.
        ldm_context* lead = handles[0];   // all on one device
        LDM_CUDA(cudaSetDevice(lead->device));
        std::vector<long long> offset((size_t)n);
        size_t total = 0;
        for (int i = 0; i < n; i++) { offset[(size_t)i] = (long long)total; total += 1 + handles[i]->P(); }
        double* stage_host = nullptr;
        double* stage_device = nullptr;
        ldm::StateMove* moves_device = nullptr;
        if (inits || lasts) {
            const size_t bytes = total * sizeof(double) + 2 * (size_t)n * sizeof(ldm::StateMove);
            if (lead->batch_stage_bytes < bytes) {
                if (lead->batch_stage_host) cudaFreeHost(lead->batch_stage_host);
                if (lead->batch_stage_device) cudaFree(lead->batch_stage_device);
                lead->batch_stage_host = lead->batch_stage_device = nullptr;
                lead->batch_stage_bytes = 0;
                LDM_CUDA(cudaMallocHost(&lead->batch_stage_host, bytes));
                LDM_CUDA(cudaMalloc(&lead->batch_stage_device, bytes));
                lead->batch_stage_bytes = bytes;
            }
            LDM_CUDA(cudaStreamSynchronize(lead->stream));                    // the buffer's last use has left
            stage_host = static_cast<double*>(lead->batch_stage_host);
            stage_device = static_cast<double*>(lead->batch_stage_device);
            ldm::StateMove* moves_host = reinterpret_cast<ldm::StateMove*>(stage_host + total);
            moves_device = reinterpret_cast<ldm::StateMove*>(stage_device + total);
            const int last_parity = (points - 1) & 1;
            for (int i = 0; i < n; i++) {
                ldm_context* h = handles[i];
                if (inits) std::memcpy(stage_host + offset[(size_t)i], inits[i], (1 + h->P()) * sizeof(double));
                moves_host[i] = ldm::StateMove{offset[(size_t)i], (long long)h->P(), h->d_N, h->d_C[0]};                         // in: parity 0
                moves_host[n + i] = ldm::StateMove{offset[(size_t)i], (long long)h->P(), h->d_N + last_parity, h->d_C[last_parity]};   // out
                if (inits) h->cur = 0;
            }
            if (inits) {
                LDM_CUDA(cudaMemcpyAsync(stage_device, stage_host, bytes, cudaMemcpyHostToDevice, lead->stream));
                ldm::launch_state_moves(moves_device, n, stage_device, true, lead->stream);
            } else {
                LDM_CUDA(cudaMemcpyAsync(moves_device, moves_host, 2 * (size_t)n * sizeof(ldm::StateMove), cudaMemcpyHostToDevice, lead->stream));
            }
        }
        struct Group {
            ldm::FactoredPlan shape;
            std::vector<int> members;
            int max_slabs = 0, max_rows = 0, max_G = 0, max_A = 0, max_ring = 0, max_points = 0;
            size_t max_smem = 0;     // prefix-sum groups: the largest member's shared memory
            int scan_nt = 512;
            bool persistent = false;   // all steps inside one launch (persistent_batch_kernel)
            ldm::StepArgs* d_args[2] = {nullptr, nullptr};
        };
        std::vector<Group> groups;
        for (int i = 0; i < n; i++) {
            ldm_context* h = handles[i];
            ensure_ring_operator(h);
                Group* g = nullptr;
            for (Group& cand : groups)
                if ((cand.shape.scan && h->plan.scan && cand.scan_nt == h->dev.scan_nt) ||   // the prefix-sum kernel takes its tile shape at run time
                    (!cand.shape.scan && !h->plan.scan && cand.shape.bm == h->plan.bm && cand.shape.bn == h->plan.bn &&
                     cand.shape.tn == h->plan.tn)) g = &cand;
            if (!g) { groups.emplace_back(); g = &groups.back(); g->shape = h->plan; g->scan_nt = h->dev.scan_nt; }
            g->members.push_back(i);
            g->max_slabs = std::max(g->max_slabs, h->plan.slabs());
            g->max_rows = std::max(g->max_rows, h->dev.G);
            g->max_G = std::max(g->max_G, h->dev.G);
            g->max_A = std::max(g->max_A, h->dev.A);
            g->max_ring = std::max(g->max_ring, h->dev.n_ring);
            g->max_points = std::max(g->max_points, h->dev.G * h->dev.G);
            if (h->plan.scan) g->max_smem = std::max(g->max_smem, ldm::scan_smem_bytes(h->dev.G, h->dev.A, h->dev.scan_bn, h->dev.scan_nt, h->dev.scan_max_pieces));
        }
        std::vector<void*> scratch;
        try {
            for (Group& g : groups)
                for (int parity = 0; parity < 2; parity++) {
                    std::vector<ldm::StepArgs> host;
                    for (int i : g.members) {
                        ldm_context* h = handles[i];
                        host.push_back(make_step_args(h, h->d_C[parity], h->d_N + parity, h->d_N + (parity ^ 1), dt, nullptr, nullptr,
                                                      h->d_C[parity ^ 1]));
                    }
                    void* p = nullptr;
                    LDM_CUDA(cudaMalloc(&p, host.size() * sizeof(ldm::StepArgs)));
                    scratch.push_back(p);
                    LDM_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(ldm::StepArgs), cudaMemcpyHostToDevice));
                    g.d_args[parity] = static_cast<ldm::StepArgs*>(p);
                }
            cudaStream_t s = lead->stream;
            LDM_CUDA(cudaEventRecord(lead->ev_begin, s));
            // Prefix-sum groups whose rows are one chunk each: ALL steps of all their sets in one launch per wave of the
            // persistent batch kernel, every set with a team of CTAs sized by its cost.  Everything else: one launch per
            // kernel and step, below.
            for (Group& g : groups) {
                g.persistent = lead->persistent_enabled && g.shape.scan && points > 1;
                for (int i : g.members)
                    g.persistent = g.persistent && handles[i]->dev.scan_chunks == 1 && handles[i]->persistent_enabled;
                if (!g.persistent) continue;
                int capacity = 0;
                ldm::BatchArgs probe{};
                if (ldm::launch_persistent_batch(probe, 0, g.scan_nt, g.max_smem, s, &capacity) != cudaSuccess || capacity < 1) {
                    cudaGetLastError();
                    g.persistent = false;
                    continue;
                }
                if (tuning().batch_capacity > 0) capacity = std::min(capacity, tuning().batch_capacity);
                const int members = (int)g.members.size();
                // cost model (SM cycles): per row ~3 k of tables, closing and barriers plus ~6.4 per (quadrature node x column)
                // of E and y-side work; a step of a set with a team of T costs ceil(G / T) rows (+ the team barrier)
                std::vector<double> row_cost((size_t)members);
                std::vector<int> rows((size_t)members);
                for (int m = 0; m < members; m++) {
                    const ldm::DevProblem& d = handles[g.members[(size_t)m]]->dev;
                    rows[(size_t)m] = d.G;
                    row_cost[(size_t)m] = 3000.0 + 6.4 * d.A * d.G;
                }
                std::vector<int> order((size_t)members);
                for (int m = 0; m < members; m++) order[(size_t)m] = m;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return row_cost[(size_t)a] * rows[(size_t)a] > row_cost[(size_t)b] * rows[(size_t)b]; });
                // one wave holds as many sets as the launch has CTAs.  (Measured: waves of WHOLE teams -- one CTA per row, tables
                // resident for all steps -- are slower, 336 against 228 us per step for 128 sets of 4096 points: a team's step then
                // waits for the row that owns most of the ring correction, which small teams average out over their rows.)
                const int waves = (members + capacity - 1) / capacity;
                std::vector<double*> wc_ptr((size_t)members);
                for (int m = 0; m < members; m++) wc_ptr[(size_t)m] = handles[g.members[(size_t)m]]->d_wc_rows2;
                auto device_copy = [&](const void* host, size_t bytes) {
                    void* p = nullptr;
                    LDM_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 8)));
                    scratch.push_back(p);
                    LDM_CUDA(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s));
                    return p;
                };
                double* const* d_wc_ptr = static_cast<double* const*>(device_copy(wc_ptr.data(), wc_ptr.size() * sizeof(double*)));
                unsigned int* d_barriers = nullptr;
                LDM_CUDA(cudaMalloc((void**)&d_barriers, (size_t)members * sizeof(unsigned int)));
                scratch.push_back(d_barriers);
                LDM_CUDA(cudaMemsetAsync(d_barriers, 0, (size_t)members * sizeof(unsigned int), s));
                for (int w = 0; w < waves; w++) {
                    std::vector<int> mine;                        // members of this wave: dealt out round-robin by cost
                    for (int at = w; at < members; at += waves) mine.push_back(order[(size_t)at]);
                    std::vector<int> team(mine.size(), 1);
                    int spare = capacity - (int)mine.size();
                    auto step_time = [&](size_t j) {
                        const int m = mine[j], T = team[j];
                        return ((rows[(size_t)m] + T - 1) / T) * row_cost[(size_t)m] + (T > 1 ? 2000.0 : 0.0);
                    };
                    while (spare > 0) {                           // one more CTA to the set whose step is the longest
                        size_t worst = mine.size();
                        double worst_time = -1;
                        for (size_t j = 0; j < mine.size(); j++)
                            if (team[j] < rows[(size_t)mine[j]] && step_time(j) > worst_time) { worst_time = step_time(j); worst = j; }
                        if (worst == mine.size()) break;
                        team[worst]++;
                        spare--;
                    }
                    std::vector<int> cta;                         // int4 per CTA
                    for (size_t j = 0; j < mine.size(); j++)
                        for (int r = 0; r < team[j]; r++) { cta.push_back(mine[j]); cta.push_back(r); cta.push_back(team[j]); cta.push_back(0); }
                    ldm::BatchArgs a{};
                    a.args[0] = g.d_args[0]; a.args[1] = g.d_args[1];
                    a.cta = static_cast<const int4*>(device_copy(cta.data(), cta.size() * sizeof(int)));
                    a.wc_rows = d_wc_ptr;
                    a.barriers = d_barriers;
                    a.n_steps = points - 1;
                    a.parity = 0;
                    a.status = lead->d_status;
                    LDM_CUDA(ldm::launch_persistent_batch(a, (int)(cta.size() / 4), g.scan_nt, g.max_smem, s, nullptr));
                    lead->launches += 1;
                }
            }
            int parity = 0;
            for (int k = 0; k + 1 < points; k++) {
                for (Group& g : groups) {
                    if (g.persistent) continue;
                    const int members = (int)g.members.size();
                    const bool side = g.max_ring > 0;
                    if (side) {
                        LDM_CUDA(cudaEventRecord(lead->ev_fork, s));
                        LDM_CUDA(cudaStreamWaitEvent(lead->aux_stream, lead->ev_fork, 0));
                        ldm::launch_ring_batch(g.d_args[parity], members, g.max_ring, lead->aux_stream);
                        LDM_CUDA(cudaEventRecord(lead->ev_join, lead->aux_stream));
                    }
                    if (g.shape.scan) ldm::launch_scan_batch(g.d_args[parity], members, g.scan_nt, g.max_slabs, g.max_rows, g.max_smem, s);
                    else ldm::launch_factored_batch(g.shape, g.d_args[parity], members, g.max_slabs, g.max_rows, g.max_G, g.max_A, s);
                    if (side) LDM_CUDA(cudaStreamWaitEvent(s, lead->ev_join, 0));
                    ldm::launch_finish_factored_batch(g.d_args[parity], members, g.max_points, s);
                    lead->launches += 2 + (side ? 1 : 0);
                }
                parity ^= 1;
            }
            LDM_CUDA(cudaEventRecord(lead->ev_end, s));
            LDM_CUDA(cudaGetLastError());
            if (lasts) {
                ldm::launch_state_moves(moves_device + n, n, stage_device, false, s);
                LDM_CUDA(cudaMemcpyAsync(stage_host, stage_device, total * sizeof(double), cudaMemcpyDeviceToHost, s));
            }
            LDM_CUDA(cudaStreamSynchronize(s));
            check_device_status(lead);
            for (int i = 0; i < n; i++) handles[i]->cur = (points - 1) & 1;
            if (lasts)
                for (int i = 0; i < n; i++) std::memcpy(lasts[i], stage_host + offset[(size_t)i], (1 + handles[i]->P()) * sizeof(double));
        } catch (...) {
            for (void* p : scratch) cudaFree(p);
            throw;
        }
        for (void* p : scratch) cudaFree(p);
}

}  // namespace

extern "C" {

int ldm_solve_batch(int n, const ldm_handle* handles, const double* const* inits, double dt, double t_max,
                    double* const* last_states)
{
    if (n < 0 || (n > 0 && (!handles || !inits || !last_states))) return fail(LDM_ERR_ARGUMENT, "NULL argument");
    if (!(dt > 0)) return fail(LDM_ERR_ARGUMENT, "step must be positive");
    return guarded([&] {
        const int points = ldm_count_points(t_max, dt);
        // entries of several species advance beside the batch, each on its own stream (its step is a replayed graph of a
        // few launches over all its pairs: multi_advance); the one-species entries form the batch proper
        std::vector<ldm_context*> several;
        std::vector<int> several_slot;
        std::vector<ldm_handle> one_species;
        std::vector<const double*> one_inits;
        std::vector<double*> one_last;
        for (int i = 0; i < n; i++) {
            if (!handles[i] || !inits[i] || !last_states[i]) throw std::invalid_argument("NULL entry in batch");
            if (handles[i]->species != 1) { several.push_back(handles[i]); several_slot.push_back(i); }
            else { one_species.push_back(handles[i]); one_inits.push_back(inits[i]); one_last.push_back(last_states[i]); }
        }
        for (size_t j = 0; j < several.size(); j++) {
            ldm_context* h = several[j];
            LDM_CUDA(cudaSetDevice(h->device));
            multi_upload(h, inits[several_slot[j]], h->stream);
            multi_advance(h, points - 1, dt, h->stream);
            multi_download(h, last_states[several_slot[j]], h->stream);
        }
        struct Drain {                                   // whatever happens to the batch: the entries beside it are waited for
            std::vector<ldm_context*>& entries;
            ~Drain() { for (ldm_context* h : entries) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); } }
        } drain{several};
        n = (int)one_species.size();
        handles = one_species.data();
        const double* const* inits_of = one_inits.data();
        double* const* last_of = one_last.data();
        bool batched = n > 0;
        for (int i = 0; i < n; i++) {
            if (handles[i]->dev.world != 1) throw std::invalid_argument("a batch entry is configured as one rank of several (ldm_shard_configure)");
            ldm_context* h = handles[i];
            batched = batched && h->algo != LDM_ALGO_DIRECT && h->dev.world == 1 && h->dev.extrapolation != LDM_EXTRAPOLATION_EXPONENTIAL &&
                      h->closure == LDM_CLOSURE_POWER1;
        }
        if (!batched) {
            // sample-by-sample algorithm, Exponential extrapolation or a closure over the w quadrature: every parameter set
            // advances on its own stream
            for (int i = 0; i < n; i++) {
                LDM_CUDA(cudaSetDevice(handles[i]->device));
                upload_state(handles[i], inits_of[i]);
            }
            for (int k = 0; k + 1 < points; k++)
                for (int i = 0; i < n; i++) {
                    LDM_CUDA(cudaSetDevice(handles[i]->device));
                    step_resident(handles[i], dt, nullptr);
                }
            for (int i = 0; i < n; i++) {
                ldm_context* h = handles[i];
                LDM_CUDA(cudaSetDevice(h->device));
                LDM_CUDA(cudaGetLastError());
                LDM_CUDA(cudaMemcpyAsync(last_of[i], h->d_N + h->cur, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
                LDM_CUDA(cudaMemcpyAsync(last_of[i] + 1, h->d_C[h->cur], h->P() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            }
            for (int i = 0; i < n; i++) {
                LDM_CUDA(cudaSetDevice(handles[i]->device));
                LDM_CUDA(cudaStreamSynchronize(handles[i]->stream));
            }
        } else {
            // one slice of the batch per device (the caller decides which set lives where, ldm_create's device argument);
            // the devices advance concurrently, one host thread each; the states travel through one staging buffer per device
            struct Slice { int device; std::vector<ldm_context*> members; std::vector<const double*> inits; std::vector<double*> lasts; };
            std::vector<Slice> per_device;
            for (int i = 0; i < n; i++) {
                auto it = std::find_if(per_device.begin(), per_device.end(), [&](const Slice& e) { return e.device == handles[i]->device; });
                if (it == per_device.end()) { per_device.push_back(Slice{handles[i]->device, {}, {}, {}}); it = per_device.end() - 1; }
                it->members.push_back(handles[i]); it->inits.push_back(inits_of[i]); it->lasts.push_back(last_of[i]);
            }
            for (Slice& slice : per_device)                                    // pending work of the members' own streams first
                for (ldm_context* h : slice.members) {
                    LDM_CUDA(cudaSetDevice(h->device));
                    if (h != slice.members[0]) LDM_CUDA(cudaStreamSynchronize(h->stream));
                }
            if (per_device.size() == 1) {
                advance_batch_on_device(per_device[0].members, points, dt, per_device[0].inits.data(), per_device[0].lasts.data());
            } else {
                std::vector<std::string> errors(per_device.size());
                std::vector<std::thread> workers;
                for (size_t d = 0; d < per_device.size(); d++)
                    workers.emplace_back([&, d] {
                        try { advance_batch_on_device(per_device[d].members, points, dt, per_device[d].inits.data(), per_device[d].lasts.data()); }
                        catch (const std::exception& e) { errors[d] = e.what(); if (errors[d].empty()) errors[d] = "error"; }
                    });
                for (std::thread& w : workers) w.join();
                for (const std::string& e : errors) if (!e.empty()) throw CudaError(e);
            }
        }
    });
}

int ldm_device_count(void)
{
    int count = 0;
    return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}

}  // extern "C"
