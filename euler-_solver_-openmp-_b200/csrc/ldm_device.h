// Device-side view of one Law-Dieckmann problem and the kernel launchers (sm_100a).
#pragma once
#include <cuda_runtime.h>

namespace ldm {

// Everything a kernel needs about one parameter set; all pointers are device pointers.
// Index conventions: moment node x, y in [0,G); signed quadrature offset o in [0,n_off);
// alpha in [0,A) = compact index of a kernel node in the odd-node quadrant table.
struct DevProblem {
    int G, n_off, A;
    // xi-sharding: this rank computes rows x = rank, rank+world, ...; the full C is held in
    // "gathered" layout, row x at slot (x % world) * rows_per_rank + x / world.  world == 1: natural layout.
    int world, rank, rows_per_rank;
    int extrapolation;         // ExtrapolationType of C beyond the cutoff (Solver.cpp:28): 0 Zero, 1 SquaredN (default), 2 Alpha, 3 Exponential
    // Exponential only (Solver.cpp:392-408): the cutoff, its distance to the four moment nodes before the last, and the
    // WC samples beyond the cutoff one by one
    double cutoff, tail_gap[4];
    const double* wc_out_w;
    const double* wc_out_distance;
    int wc_out_n;
    double hh;                 // KernelsGridStep^2 (Solver.cpp:314)
    double b, d, dd;
    double s_max;              // largest sample abscissa that is still inside the moment grid
    double wc_outside;         // sum of w over the samples of WC that lie beyond the cutoff
    const double* s;           // [G][n_off] sample abscissa |g_x + offset_o|              (Solver.cpp:307)
    const double* limit;       // [G][n_off] inside  <=>  partner abscissa <= limit       (Solver.cpp:329-330)
    const double* weight;      // [G][n_off] interpolation weight along this axis          (Solver.cpp:344)
    const int* upper;          // [G][n_off] upper node of the interval                    (Solver.cpp:335)
    const int* upper_t;        // [n_off][G] the same two tables transposed (node-contiguous)
    const double* weight_t;
    const int* alpha;          // [n_off]
    const double* quadrant;    // [A][A] odd-node quadrant of the m table
    // prefix-sum path (ldm_scan.cu; ldm_model.h SeparableKernel): Q[beta][alpha] = qvec[beta] qvec[alpha] for
    // alpha <= a_star[beta], 0 beyond.  NULL qvec: the table is not separable, the dense contraction is used.
    const double* qvec;        // [A]
    const int* a_star;         // [A]
    const int4* beta_record;   // [A] rows beta with a_star[beta] == alpha: [x, y), and the bits of qvec[x] in z (low) and w (high)
    const unsigned short* upper_p16;   // upper_t in 16 bits and the weights with the offsets two by two: entry (o, y) at
    const double* weight_p;            // ((o >> 1) * G + y) * 2 + (o & 1)  (ldm_scan_body.cuh y_table_index); n_off rounded up to even
    const unsigned short* upper_c16;   // rows of several chunks: [scan_chunks] tables like upper_p16 holding the pair's position in the
                                       // chunk's H tile row, clamped to its zero columns (ldm_scan_body.cuh y_side_pair); else NULL
    // y-side work list (ldm_model.h ScanPieces): pieces (y block, first offset, end offset) dealt out to the warps
    const int4* scan_piece;        // all chunks' pieces
    const int* scan_chunk_begin;   // [scan_chunks + 1] first piece of each chunk
    const int* scan_warp_begin;    // [scan_chunks][scan_nt / 32 + 1] relative to the chunk's first piece
    const int* scan_block_begin;   // [scan_chunks][ceil(G/32) + 1] relative to the chunk's first piece
    const int* scan_chunk_order;   // [scan_chunks] chunks by descending y-side work (the order the CTAs are issued in)
    int scan_max_pieces;
    int scan_bn, scan_chunks;  // column chunk width (multiple of 32) and count
    int scan_live_rows;        // rows beta < scan_live_rows have a_star >= 0; the rest of the table is empty
    int scan_nt;               // threads per CTA = ng * scan_bn (ng lanes of a warp share one column)
    const double* m_nodes;     // [G][G]  m(xi) at the moment nodes                        (Solver.cpp:469)
    const double* w_nodes;     // [G][G]  w(xi)                                            (Solver.cpp:490)
    const double* out_weight;  // [G][G]  sum of m over this node's samples beyond the cutoff
    const double* wc_operator; // [G][G]  WC = hh * (<wc_operator, C> + N^2 wc_outside)   (Solver.cpp:265-274)
    // ring correction of the factored path (ldm_model.h RingOperator): CSR over the nodes of this rank's rows that
    // own samples inside the grid square but beyond the cutoff circle
    int n_ring;                // CSR rows
    const int* ring_point;     // [n_ring] local_row * G + y
    const int* ring_begin;     // [n_ring + 1]
    const int* ring_row_begin; // [local_rows + 1] first CSR row of each local moment-grid row
    const unsigned char* ring_present;   // [local_rows * G] 1 where the node owns ring samples (has a CSR row)
    const int* ring_cta_begin; // [local_rows + 1] the CSR rows dealt out evenly by entry count over local_rows CTAs (persistent kernel)
    int ring_share_entries, ring_share_rows, ring_share_row_len;   // the largest share: entries, CSR rows, entries of one row
    const int* ring_node;      // C index (gathered layout)
    const double* ring_weight;
};

// Everything one Euler step (or one right-hand side) of one problem touches.  Kernels take it by value for a single
// problem, or index an array of them by blockIdx.z when a batch of independent parameter sets shares one launch.
constexpr int kMaxPeers = 8;

struct StepArgs {
    DevProblem P;
    const double* C;           // pair density at the old time, gathered layout (full replica)
    const double* N_cur;       // first moment at the old time
    double* N_next;            // may be NULL (right-hand side only)
    double* partial;           // [slabs][local_rows][G] factored path
    double* ring;              // [local_rows][G] ring correction (zero outside the listed nodes)
    double* wc_rows;           // [G] per-row dot products of the WC operator
    double* scalars;           // [kScalars] WC, dN/dt
    double* sum_out;           // may be NULL: inside-sample sums
    double* deriv_out;         // may be NULL: dC/dt of this rank's rows
    double* C_next;            // may be NULL: updated rows
    const double* out_exp;     // Exponential extrapolation: [local_rows][G] sum of m * exp(-decay (distance - cutoff)) over the
                               // node's samples beyond the cutoff, this step's decay rate
    int tiles, chunks;         // factored plan: slab = tile * chunks + chunk
    // fused update + all-gather over NVLink (xi-sharded runs, ldm_shard_peer_*): when peer_count > 0 the closing kernel
    // stores every updated value straight into the next-step replica of EVERY rank (peer-mapped memory, own rank
    // included), then the last CTA to finish releases this rank's epoch into every rank's flag array
    int peer_count;
    double* peer_C_next[kMaxPeers];
    unsigned long long* peer_flags[kMaxPeers];
    unsigned int* done_counter;
    unsigned long long epoch;
    int* peer_status;          // the status word a lost peer is reported in
    double dt;
};

// Which samples sample_sum accumulates (all use the same (x,o_i) x (y,o_j) enumeration):
enum SampleMode {
    SAMPLES_INSIDE = 0,        // sum m * bilinear(C) over samples inside the cutoff circle  -> the direct algorithm
    SAMPLES_RING = 1,          // ... over samples inside the grid square but outside the circle -> correction of the factored path
    SAMPLES_OUT_WEIGHT = 2,    // sum m over every sample outside the circle (time-invariant, run once at create)
    SAMPLES_OUT_EXP = 3        // sum m * exp(-decay * (hypot(sx, sy) - cutoff)) over them, decay = scalars[2] (Exponential extrapolation)
};

// Scalars of one step, device resident: [0] WC, [1] dN/dt; Exponential extrapolation: [2] decay rate of this step,
// [3] C[0][G-1] - N^2 (0 when the asymptote test of S:357 holds), [4] sum of w * exp(...) over the WC samples beyond the cutoff
constexpr int kScalars = 8;

__host__ __device__ inline int row_slot(const DevProblem& P, int x)
{
    return P.world == 1 ? x : (x % P.world) * P.rows_per_rank + x / P.world;
}
__host__ __device__ inline int local_rows(const DevProblem& P) { return (P.G - P.rank + P.world - 1) / P.world; }

// C is always the full pair density in gathered layout; `out`, inside_sum, deriv_out, C_next are
// shard-local [local_rows][G].
void launch_record_observables(const DevProblem& P, const double* N_cur, const double* C, double* obs_row, cudaStream_t stream);

// RungeKutta4thSolver (Solver.cpp:597-654) on [N | C] state vectors of n doubles: out = value + (c * step / divisor);
// out = clamp(value + ((c0 + c1*2 + c2*2 + c3) * step / 6))
void launch_rk4_stage(int n, const double* value, const double* c, double step, double divisor, double* out, cudaStream_t stream);
void launch_rk4_close(int n, const double* value, const double* c0, const double* c1, const double* c2, const double* c3, double step,
                      double* out, cudaStream_t stream);

void launch_sample_sum(const DevProblem& P, SampleMode mode, const double* C, double* out, cudaStream_t stream,
                       const double* scalars = nullptr);
// Exponential extrapolation, once per right-hand side: decay rate from row 0 of C (S:392-405), the factor C[0][G-1] - N^2
// and the WC sum over the outside samples -> scalars[2..4]
void launch_exponential_prepare(const DevProblem& P, const double* C, const double* N_cur, double* scalars, cudaStream_t stream);

void launch_wc(const DevProblem& P, const double* C, const double* N_cur, double* N_next, double* scalars,
               double dt, cudaStream_t stream);

// Pointwise assembly (Solver.cpp:459-508) + optional explicit-Euler update and clamp (Solver.cpp:577-580).
// inside_sum: per node sum over inside samples of m*C~ (unscaled).  deriv_out / C_next may be NULL.
void launch_pointwise(const DevProblem& P, const double* C, const double* inside_sum, const double* N_cur,
                      const double* scalars, double dt, double* deriv_out, double* C_next, cudaStream_t stream,
                      const double* out_exp = nullptr);

// Closing kernel of the factored path: inside_sum = sum_tiles partial - ring, WC from the per-row dot products,
// the first-moment equation with its update (block 0), then the same pointwise assembly/update as above.
// sum_out (may be NULL) receives inside_sum for ldm_integrals.
void launch_finish_factored(const StepArgs& one, cudaStream_t stream);
void launch_finish_factored_batch(const StepArgs* d_batch, int n, int max_points, cudaStream_t stream);

// Closing kernel for several species (Solver.cpp:446-511 with SpeciesNumber > 1): every pair (i, j) has had its integrals
// computed by the one-species kernels (partial sums, ring correction, WC row dots); this kernel forms WC_ij, MC_ij, the
// first-moment equations and dC_ij in the reference's operation order, and optionally the clamped Euler update.
constexpr int kMaxSpecies = 4;
struct MultiPair {
    const double* C;           // [P] pair density C_ij at the old time
    double* C_next;            // may be NULL
    double* deriv_out;         // may be NULL: dC_ij/dt
    const double* partial;     // [slabs][P]
    const double* ring;        // [P]
    const double* wc_rows;     // [G]
    const double* out_weight;  // [P] of species i
    const double* m_nodes;     // [P] m_i at the nodes
    const double* w_nodes;     // [P] w_ij at the nodes
    double wc_outside;
    int slabs;
};
struct MultiArgs {
    int species, G, extrapolation;
    double hh, dt;
    double b[kMaxSpecies], d[kMaxSpecies], dd[kMaxSpecies * kMaxSpecies];
    const double* N_cur;       // [species]
    double* N_next;            // may be NULL
    double* dN_out;            // [species]
    MultiPair pair[kMaxSpecies * kMaxSpecies];
};
void launch_finish_multi(const MultiArgs& args, cudaStream_t stream);

// blocks until every rank's flag has reached `epoch` (or a spin budget runs out: *status becomes non-zero)
void launch_peer_wait(const unsigned long long* flags, int world, unsigned long long epoch, int* status, cudaStream_t stream);

// ring[lx][y] for the nodes of the ring CSR (other entries of `ring` stay zero)
void launch_ring(const StepArgs& one, cudaStream_t stream);
void launch_ring_batch(const StepArgs* d_batch, int n, int max_ring_rows, cudaStream_t stream);
// The states of a batch between ONE staging buffer and the members' own buffers (ldm_solve_batch: one host <-> device copy
// for all members instead of two per member).  Member i's state [N | C] lies at stage + offset doubles.
struct StateMove {
    long long offset, count;   // doubles in front of the member's state; doubles of its C
    double* N;                 // the member's first moment (this parity)
    double* C;                 // its pair density (this parity)
};
void launch_state_moves(const StateMove* d_moves, int n, double* stage, bool to_members, cudaStream_t stream);

// Factored path (ldm_factored.cu): partial[slab][local_rows][G] <- per-slab sums over all samples inside the grid
// square; the closing kernel forms inside_sum = sum_slabs partial - ring.
// Arguments of the persistent multi-step kernel (ldm_factored.cu): `step` carries the problem and the scratch
// buffers; the state ping-pongs between C[0]/C[1] and N[0]/N[1] starting at `parity`.
struct PersistentArgs {
    StepArgs step;
    double* C[2];
    double* N;                 // [2]
    int parity;
    int n_steps;
    long long first_point;     // index of the stored point the first step starts from
    long long stride;          // doubles per stored point (1 + G*G)
    double* observables;       // may be NULL: [points][2]
    double* values;            // may be NULL: [points][stride]
    double* derivs;            // may be NULL: [points][stride]
    unsigned int barriers_done; // prefix-sum persistent kernel: grid barriers the counter has seen in earlier launches (it is monotone: no reset per launch)
    int last_is_rhs_only;      // the last of the n_steps only evaluates the right-hand side at its state (derivs, scalars): nothing is advanced
    int cached_w_rows;         // set by the launcher: offsets whose y-side weights stay in shared memory (-1: no cache)
    // prefix-sum persistent kernel: [2][G] row dots of the WC functional (ping-pong with the state) and the arrival
    // counter of its grid barrier (zeroed before every launch)
    double* wc_rows;
    unsigned int* barrier;
    double* ring_pingpong;     // [2][G*G] ring values of even / odd steps; every byte 0xFF (an all-ones NaN) = not yet published
    int* status;               // the handle's device status word (ldm_scan_body.cuh: bounded waits)
    long long ring_cache_offset;   // set by the launcher: byte offset in shared memory of this CTA's copy of its ring share (nodes, then
                                   // weights), or -1 when the shares do not have the one-row-per-warp shape
};

struct FactoredPlan {
    int bm = 80, bn = 64;      // CTA tile: beta rows x columns
    int tn = 8;                // columns per thread (8: 8-warp tiles, 4: twice the threads); the contraction runs on the FP64
                               // tensor cores (mma.sync m8n8k4)
    bool scan = false;         // prefix-sum path (separable kernel table): tiles = 1, chunks = P.scan_chunks, runtime tile shape
    int tiles = 1, chunks = 1; // beta tiles x column chunks; slabs = tiles * chunks partial sums per node
    int slabs() const { return tiles * chunks; }
};
FactoredPlan plan_factored(const DevProblem& P);

// Prefix-sum path (ldm_scan.cu).  scan_pick_bn: widest column chunk whose working set fits one CTA's shared memory.
int scan_pick_bn(int G, int A);
int scan_pick_threads(int G, int A, int bn);
void launch_scan(const StepArgs& one, cudaStream_t stream);
void launch_scan_batch(const StepArgs* d_batch, int n, int nt, int max_chunks, int max_rows, size_t max_smem, cudaStream_t stream);
size_t scan_smem_bytes(int G, int A, int bn, int nt, int max_pieces);
int scan_lead_columns();                       // doubles in front of column 0 of a row of the H tile
cudaError_t launch_persistent_scan(const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid);
// single problem: `one` by value; batch: d_batch[n] in device memory, all entries sharing the plan's tile shape
void launch_factored(const FactoredPlan& plan, const StepArgs& one, cudaStream_t stream);
void launch_factored_batch(const FactoredPlan& shape, const StepArgs* d_batch, int n, int max_slabs, int max_rows, int max_G, int max_A,
                           cudaStream_t stream);

// grid <= 0: only report the largest co-resident grid in *max_grid
cudaError_t launch_persistent(const FactoredPlan& plan, const PersistentArgs& args, int grid, cudaStream_t stream, int* max_grid);

// Arguments of the persistent work-list kernel (ldm_persist.cu persistent_chunks_kernel): n_steps Euler steps of one
// problem on `world` ranks in ONE launch per rank.  The pair density lives in two replicas per rank (ping-pong), each
// `words` doubles in gathered layout followed by a tail of G doubles: the row dots of the WC functional for that state.
// Rank r's closing code stores its updated rows -- and their row dots -- straight into replica[parity ^ 1][*] of EVERY
// rank (peer-mapped memory over NVLink, own rank included), then the number of the new state into the row's flag
// flags[*][kMaxPeers + x] and, once all its rows are out, the number of steps it has completed into flags[*][r].  An item of
// the next step starts when the rows its E phase reads have arrived (row_window), a closing when every rank has published.
constexpr int kFlagWords(int G) { return kMaxPeers + G; }     // a rank's flag array: [kMaxPeers] steps per rank, [G] states per row
struct ChunkedArgs {
    StepArgs step;             // the problem and its scratch (partial sums, scalars); C and N are set per step by the kernel
    int n_steps, parity, world, rank;
    unsigned long long epoch;  // steps taken since the upload: what every rank has published when this launch may begin
    long long words;           // doubles of a replica in front of its WC tail
    double* replica[2][kMaxPeers];
    unsigned long long* flags[kMaxPeers];
    double* N;                 // [2] first moment, ping-pong with the replicas (every rank advances its own copy: no all-reduce)
    const int4* items;         // (local row, chunk, first and last moment-grid row that the row's E phase reads) in the order they are
                               // handed out: rows ascending, a row's chunks by descending cost
    int n_items;
    const int2* row_window;    // [G] first and last moment-grid row that the E phase of row x reads
    const int4* block_window;  // [G][blocks_per_row] the same per block of alphas that one thread group takes in one super-chunk of the E
                               // phase (block = super-chunk * groups + group): rows lo..hi of the minus side, lo..hi of the plus side
    int blocks_per_row;
    unsigned int* counters;    // zeroed before the launch: [0] item fetch counter, [1] rows closed, [2] prologue barrier,
                               // [3 + lx] chunk arrivals of local row lx; all monotone over the steps of the launch
    double* ring_pingpong;     // [2][local_rows * G] ring values of even / odd steps; all-ones NaN = not yet published
    const int* ring_cta_begin; // [grid + 1] the ring CSR rows of the HEAVY grid rows dealt out evenly by entry count over the CTAs of this launch
    int self_ring_rows;        // local rows lx < self_ring_rows are light: their closing CTA forms their ring correction itself (at most
    int self_ring_capacity;    // self_ring_capacity entries per grid row, staged in the idle H tile); the rest goes through the shares
    int* status;
    // history, one rank only (may be NULL): as PersistentArgs
    long long first_point, stride;
    double* observables;
    double* values;
    double* derivs;
};
// the E phase's block structure for this problem's CTA shape: alphas per block, thread groups, padded alpha count
void chunk_block_shape(const DevProblem& P, int* alphas_per_block, int* groups, int* padded_alphas);
// grid <= 0: only report the largest co-resident grid in *max_grid
cudaError_t launch_persistent_chunks(const ChunkedArgs& args, int grid, cudaStream_t stream, int* max_grid);

// Arguments of the persistent batch kernel (ldm_persist.cu persistent_batch_kernel): n_steps Euler steps of MANY independent
// parameter sets -- the loop of Experiments::CalculateWithVaryingSigma (Solver.cpp:940-963) -- in ONE launch.  Every set
// owns a team of consecutive CTAs (sized by its cost on the host); the CTAs of a team deal the set's rows out among
// themselves and meet at a per-set barrier once per step.  No grid-wide barrier: the sets never wait for each other.
struct BatchArgs {
    const StepArgs* args[2];   // [parity][sets] the sets' arguments for both parities of their state ping-pong
    const int4* cta;           // [grid] (set, rank within its team, team size, 0)
    double* const* wc_rows;    // [sets] -> [2][G] row dots of the WC functional, ping-pong with the state
    unsigned int* barriers;    // [sets] arrival counters of the team barriers (zeroed before the launch)
    int n_steps, parity;
    int* status;
};
// grid <= 0: only report the largest co-resident grid in *max_grid (for CTAs of `nt` threads with `smem` bytes)
cudaError_t launch_persistent_batch(const BatchArgs& args, int grid, int nt, size_t smem, cudaStream_t stream, int* max_grid);

// Shape overrides for tests (ldm_tuning_set): 0 / -1 = the library decides.  Read when a handle is created.
struct Tuning {
    int scan_bn = 0, scan_nt = 0;              // prefix-sum path: chunk width, threads per CTA
    int factored_bm = 0, factored_bn = 0, factored_tn = 0;   // dense contraction: tile shape
    int persistent = -1;                       // 0: one launch per step even where a persistent kernel applies
    int scan = -1;                             // 0: never take the prefix-sum path
    int chunk_grid = 0;                        // persistent_chunks_kernel: CTAs per launch (tests run two ranks on one device)
    int batch_capacity = 0;                    // persistent_batch_kernel: CTA budget of a launch (tests: several waves of sets)
    int multi_graph = -1;                      // 0: the multi-species loop as ordinary launches (timing comparisons)
    int item_order = 0;                        // persistent_chunks_kernel, experiments: 0 rows ascending (default), 1 chunk by chunk, 2 rows descending
};
Tuning& tuning();

}  // namespace ldm
