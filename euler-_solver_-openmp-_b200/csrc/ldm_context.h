// Host-side state behind an ldm_handle and the helpers the entry points of include/ldm.h share (ldm_api.cu: create, solve,
// batch; ldm_shard_api.cu: the xi-sharded entry points).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ldm.h"
#include "ldm_device.h"
#include "ldm_model.h"

namespace ldm {

int fail(int code, const std::string& what);          // sets ldm_last_error() of the calling thread, returns code

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

}  // namespace ldm

#define LDM_CUDA(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            throw ldm::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_));                     \
    } while (0)

struct ldm_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;     // the ring correction runs beside the factored kernel
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_fork = nullptr, ev_join = nullptr;
    // history of ldm_solve: two segment buffers kept between calls, copied out on their own stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_segment[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    double* d_hist[2] = {nullptr, nullptr};
    size_t hist_capacity = 0;              // doubles per buffer
    ldm::Tables tables;
    ldm::DevProblem dev{};
    ldm::FactoredPlan plan;                // the active plan: plan_scan or plan_dense
    ldm::FactoredPlan plan_dense, plan_scan;
    ldm::SeparableKernel separable;        // structure of the m table (prefix-sum path when .ok)
    bool scan_available = false;
    std::vector<void*> allocations;
    double* d_C[2] = {nullptr, nullptr};   // ping-pong pair density, natural layout; each P doubles plus a tail of G (the WC row dots
                                           // of persistent_chunks_kernel)
    double* d_N = nullptr;                 // [2] ping-pong first moment
    double* d_scalars = nullptr;           // [kScalars]
    double* d_sum = nullptr;               // [P] per-node inside-sample sum
    double* d_deriv = nullptr;             // [P] dC/dt
    double* d_partial = nullptr;           // [tiles][P] factored path: per beta-tile sums
    double* d_ring = nullptr;              // [P] factored path: samples between the cutoff circle and the square edge
    double* d_out_exp = nullptr;           // [P] Exponential extrapolation: per-node sum over the outside samples, this step's decay
    double* d_wc_rows = nullptr;           // [G] factored path: per-row dot products of the WC operator (3-kernel path)
    double* d_wc_rows2 = nullptr;          // [2][G] the same inside the prefix-sum persistent kernel (ping-pong with the state)
    double* d_ring2 = nullptr;             // [2][P] ring values of even / odd steps inside that kernel (PersistentArgs::ring_pingpong)
    unsigned int* d_barrier = nullptr;     // [1 + G] arrival counter of that kernel's grid barrier, then its per-row ring counts
    // fused update + all-gather (ldm_shard_peer_*)
    double* peer_own[2] = {nullptr, nullptr};                 // this rank's two replicas of C (gathered layout)
    unsigned long long* peer_own_flags = nullptr;             // [kFlagWords(G)] last step published by every rank, newest state of every row
    unsigned int* peer_done = nullptr;
    int* peer_status = nullptr;
    double* peer_C[2][ldm::kMaxPeers] = {};                   // every rank's replicas, peer-mapped
    unsigned long long* peer_flags[ldm::kMaxPeers] = {};
    std::vector<void*> peer_opened;
    unsigned long long peer_epoch = 0;
    int peer_parity = 0;
    bool peer_connected = false;
    bool peer_solo = false;                // debug build only (ldm_debug_shard_solo): this rank steps alone, the others count as done
    int persistent_grid = 0;               // co-resident CTAs of the persistent kernel (0: not queried yet, -1: unavailable)
    bool persistent_enabled = true;
    bool work_list_on_one_rank = false;    // ldm_tuning_set("persistent", 1): the work-list kernel also for unsharded chunked grids (tests, measurements)
    // persistent work-list kernel (ldm_persist.cu persistent_chunks_kernel): rows of several chunks and the xi-sharded step
    int chunk_grid = 0;                    // co-resident CTAs of that kernel (0: not queried yet, -1: unavailable)
    int chunk_grid_limit = 0;              // > 0: launch at most this many CTAs (ldm_tuning_set chunk_grid; two ranks on one device in tests)
    int4* d_items = nullptr;               // (local row, chunk) work list of the current (rank, world): rows ascending
    int2* d_row_window = nullptr;          // [G] rows that the E phase of row x reads (ChunkedArgs::row_window)
    int4* d_block_window = nullptr;        // [G][blocks_per_row] ... per block of alphas (ChunkedArgs::block_window)
    int blocks_per_row = 0;
    int n_items = 0;
    int items_rank = -1, items_world = -1;
    int* d_ring_cut = nullptr;             // [cut_grid + 1] ring CSR rows dealt out over the CTAs of a launch
    int cut_grid = 0, cut_rank = -1, cut_world = -1;
    std::vector<int> chunk_order_host;     // the column chunks by descending y-side work (DevProblem::scan_chunk_order)
    std::vector<int> ring_row_begin_host;  // CSR row starts of the current ring operator (to deal the rows out by entry count)
    std::vector<int> ring_grid_row_begin_host;   // first CSR row of each local moment-grid row
    int self_ring_rows = 0, self_ring_capacity = -1;   // ChunkedArgs: the leading light rows close with their own ring correction
    unsigned int* d_counters = nullptr;    // [3 + rows] ChunkedArgs::counters
    unsigned long long* d_flags = nullptr; // [kFlagWords(G)] one rank without peers: its own flag array
    unsigned long long local_epoch = 0;    // steps that kernel has taken on this handle outside a peer group
    int* d_status = nullptr;               // sticky device status word of the bounded waits (ldm_scan_body.cuh); 0 = fine
    std::vector<void*> ring_allocations;   // the ring-correction CSR of the current (rank, world)
    int ring_rank = -1, ring_world = -1;
    long long ring_entries = 0, ring_samples = 0;
    // several species (ldm_problem_generate_multi): species*species one-species handles -- pair (i, j) carries the tables of
    // M(i) and W(i,j) and the pair density C_ij in its own d_C ping-pong; the first moments live here
    int species = 1;
    std::vector<ldm_context*> sub;
    double* d_Nm = nullptr;                // [2][species]
    double* d_dN = nullptr;                // [species]
    double rate_b[ldm::kMaxSpecies] = {}, rate_d[ldm::kMaxSpecies] = {}, rate_dd[ldm::kMaxSpecies * ldm::kMaxSpecies] = {};
    // the pairs grouped by kernel shape: one batched launch per group and kernel (blockIdx.z = pair), argument arrays for
    // both parities of the state ping-pong kept on the device; rebuilt when the algorithm changes
    struct PairGroup {
        ldm::FactoredPlan shape;
        int scan_nt = 512, members = 0, max_slabs = 0, max_rows = 0, max_G = 0, max_A = 0, max_ring = 0;
        size_t max_smem = 0;
        ldm::StepArgs* d_args[2] = {nullptr, nullptr};
    };
    std::vector<PairGroup> pair_groups;
    // the launches of two consecutive Euler steps (both parities of the state ping-pong) as an instantiated CUDA graph: a
    // multi-species step is three to five small launches, so the loop of ldm_advance / ldm_solve / ldm_solve_batch replays
    // it; rebuilt when the step size, the extrapolation type or the pair groups change
    // ldm_solve_batch, on the first member of a device's slice: pinned host and device staging for the states of all members
    // (+ their move descriptors), kept between calls
    void* batch_stage_host = nullptr;
    void* batch_stage_device = nullptr;
    size_t batch_stage_bytes = 0;
    unsigned int barriers_done = 0;        // grid barriers the arrival counter of the one-chunk persistent kernel has seen (monotone across launches)
    bool persistent_dirty = true;          // counter and ring hand-over slots need their initial values (first launch, or a launch that timed out)
    int* h_status = nullptr;               // pinned: the status word read back behind the work of a call, without a round trip of its own
    cudaGraphExec_t multi_graph = nullptr;
    double multi_graph_dt = 0;
    int multi_graph_extrapolation = -1;
    long long multi_graph_launches = 0;
    // closures whose triple integral does not factorise (ldm_closure.cu; ldm.h ldm_set_closure).  w_seed: what the tables
    // of the w quadrature are built from when such a closure is first asked for (grids and the W table of ldm_create's
    // problem); closure_w: those tables on the device, laid out like the m side's so that the same enumeration serves
    int closure = LDM_CLOSURE_POWER1;
    ldm::OwnedProblem w_seed;
    bool closure_ready = false;
    ldm::DevProblem closure_w{};
    double closure_w_sum = 0;              // sum of w over all quadrature samples (times hh: the midpoint integral of w)
    double* d_own_density = nullptr;       // [n_off_w][n_off_w] C~ at the samples of the WC integral, this step
    double* d_w_conv = nullptr;            // [P] (w*C)(xi)
    double* d_k_int = nullptr;             // [P] K(xi)
    int cur = 0;
    int algo = LDM_ALGO_FACTORED;
    long long launches = 0;
    float last_ms = 0.f;
    double shard_N[2] = {0, 0};

    size_t P() const { return (size_t)dev.G * dev.G; }

    template <typename T>
    T* alloc(size_t count)
    {
        void* p = nullptr;
        LDM_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        allocations.push_back(p);
        return static_cast<T*>(p);
    }
    template <typename T>
    T* upload(const std::vector<T>& v)
    {
        T* p = alloc<T>(v.size());
        if (!v.empty()) LDM_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }
    ~ldm_context()
    {
        for (ldm_context* c : sub) delete c;
        cudaSetDevice(device);
        for (PairGroup& g : pair_groups) { cudaFree(g.d_args[0]); cudaFree(g.d_args[1]); }
        if (multi_graph) cudaGraphExecDestroy(multi_graph);
        if (h_status) cudaFreeHost(h_status);
        if (batch_stage_host) cudaFreeHost(batch_stage_host);
        if (batch_stage_device) cudaFree(batch_stage_device);
        for (void* p : allocations) cudaFree(p);
        for (void* p : ring_allocations) cudaFree(p);
        for (void* p : peer_opened) cudaIpcCloseMemHandle(p);
        for (void* p : {(void*)peer_own[0], (void*)peer_own[1], (void*)peer_own_flags, (void*)peer_done, (void*)peer_status})
            if (p) cudaFree(p);
        for (int b = 0; b < 2; b++) {
            if (d_hist[b]) cudaFree(d_hist[b]);
            if (ev_segment[b]) cudaEventDestroy(ev_segment[b]);
            if (ev_copied[b]) cudaEventDestroy(ev_copied[b]);
        }
        if (d_items) cudaFree(d_items);
        if (d_ring_cut) cudaFree(d_ring_cut);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace ldm {

template <typename F>
int guarded(F&& body)
{
    try {
        body();
        return LDM_OK;
    } catch (const CudaError& e) {
        return fail(LDM_ERR_CUDA, e.what());
    } catch (const std::invalid_argument& e) {
        return fail(LDM_ERR_ARGUMENT, e.what());
    } catch (const std::exception& e) {
        return fail(LDM_ERR_ARGUMENT, e.what());
    }
}

// defined in ldm_api.cu
void ensure_ring_operator(ldm_context* c);
StepArgs make_step_args(ldm_context* c, const double* C, const double* N_cur, double* N_next, double dt, double* sum_out,
                        double* deriv_out, double* C_next);
void rhs_and_update(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, double* N_next, double dt,
                    double* sum_out, double* deriv_out, double* C_next, cudaStream_t s);
// throws CudaError when a bounded wait of a persistent kernel ran out (the stream must have drained); clears the word
void check_device_status(ldm_context* c);

// defined in ldm_closure.cu: what follows the m*C integral and the first-moment equation when the handle's closure is not
// the reference's -- the integrals over the w quadrature and the closing with the exchanged term.  inside_sum: the m*C
// sums of this right-hand side (c->d_sum); scalars[0] = WC.
void closure_close(ldm_context* c, const DevProblem& P, const double* C, const double* N_cur, double dt, double* deriv_out,
                   double* C_next, cudaStream_t s);
// test hook (ldm_tuning_set "closure_tiles"): which variant of the sums kernel runs -- a lane's offsets held in registers
// (2 tiles of 32) or read per sample (0); anything else: chosen by the width of the quadrature
void closure_force_tiles(int tiles);

// defined in ldm_shard_api.cu: the persistent work-list kernel on this handle's current sharding.  replicas: the
// handle's own d_C (one rank, natural layout) or the peer group's.  History pointers (device) may be NULL.
bool chunked_possible(ldm_context* c);
void advance_chunked(ldm_context* c, int n_steps, double dt, cudaStream_t s, long long first_point, double* d_obs, double* d_values,
                     double* d_derivs);

}  // namespace ldm
