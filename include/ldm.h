/* ldm.h -- C ABI of the B200-native Law-Dieckmann spatial-moment step (libldm_b200.so).
 *
 * The reference (Kintobor1370/Euler-_Solver_-OpenMP-, "Euler Solver ver.5.3") has no FFI: its seams
 * are two by-value C++ calls inside one translation unit.  Each entry point below names the
 * reference interface it replaces (S:n = Solver.cpp line n).  Plain pointers and sizes only; all
 * state vectors use the reference's layout [N | C row-major G x G] (S:255-262), one species.
 *
 * Every function returns LDM_OK (0) or a negative code; ldm_last_error() gives the text.  There is no
 * CPU fallback: without a CUDA device every compute entry point fails with LDM_ERR_CUDA.
 */
#ifndef LDM_H
#define LDM_H

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only what this header declares is exported */
#if defined(__GNUC__)
#define LDM_API __attribute__((visibility("default")))
#else
#define LDM_API
#endif

#define LDM_OK              0
#define LDM_ERR_ARGUMENT   -1   /* the reference would assert (S:260, S:327, ...) */
#define LDM_ERR_CUDA       -2   /* no device / CUDA runtime failure */
#define LDM_ERR_UNSUPPORTED -3  /* a combination that is not built (see the entry point's comment) */
#define LDM_MAX_SPECIES     4

/* What the reference keeps in `struct Properties` (S:40-60) for one species, flattened.
 * Tables are K x K row-major, exactly the reference's M(0) and W(0,0). */
typedef struct ldm_problem {
    int species;                  /* SpeciesNumber (S:42); 1 ... LDM_MAX_SPECIES */
    int G;                        /* MomentsGrid.size() (S:57) */
    int K;                        /* KernelsGrid.size() (S:59), odd */
    int mM;                       /* MaxNumForM[0] (S:52) */
    int mW;                       /* MaxNumForW[0][0] (S:51) */
    double h;                     /* KernelsGridStep (S:60) */
    double cutoff;                /* CutoffDistance (S:58) */
    double sigma_m, sigma_w;      /* SigmaM[0], SigmaW[0][0] (S:54-55) */
    double b, d, d_prime;         /* S:44-46 */
    const double* moments_grid;   /* G */
    const double* kernels_grid;   /* K */
    const double* M;              /* K*K; species > 1: species tables M(0), M(1), ... one after the other */
    const double* W;              /* K*K; species > 1: species*species tables W(0,0), W(0,1), ... */
    /* species > 1 (ldm_problem_generate_multi; NULL otherwise): the per-species / per-pair values of S:44-55 */
    const int* mM_of;             /* species:          MaxNumForM[i] */
    const int* mW_of;             /* species*species:  MaxNumForW[i][j] */
    const double* sigma_m_of;     /* species */
    const double* sigma_w_of;     /* species*species */
    const double* b_of;           /* species */
    const double* d_of;           /* species */
    const double* d_prime_of;     /* species*species */
} ldm_problem;

typedef struct ldm_context* ldm_handle;

/* algorithm selector for ldm_set_algorithm */
#define LDM_ALGO_AUTO      0
#define LDM_ALGO_DIRECT    1   /* per-sample gather, the reference's sum re-ordered only */
#define LDM_ALGO_FACTORED  2   /* interpolate - contract with the kernel table - interpolate, plus an exact cutoff correction */
#define LDM_ALGO_PREFIX    3   /* the same with the contraction collapsed to a running sum: needs a table that is separable
                                  inside its disc (every table of Properties::GenerateKernel, S:98-145, is); AUTO picks it
                                  when the table passes the check at ldm_create, LDM_ALGO_FACTORED otherwise */

/* history selector of ldm_solve (SolverResults, S:525-537) */
#define LDM_RECORD_LAST        0   /* last state only (what Experiment_14_6 reads, S:1050-1055) */
#define LDM_RECORD_OBSERVABLES 1   /* N and C[0][G-1] at every point (what Experiment_14_4 reads, S:1011-1015) */
#define LDM_RECORD_VALUES      2   /* SolverResults::Values: every state */
#define LDM_RECORD_ALL         3   /* Values and Derivatives */

/* ---- model setup: replaces Properties::Generate / GenerateGrids / GenerateKernel (S:63-241), host only ----
 * forced_G > 0 pins the moment-grid node count (the synthetic grids of BASELINE.json); <= 0 = the
 * reference's search (S:68-74).  The returned problem owns its arrays; free with ldm_problem_free. */
LDM_API int  ldm_problem_generate(double sigma_w, double sigma_m, double b, double d, double d_prime,
                          int forced_G, ldm_problem** out);
LDM_API void ldm_problem_free(ldm_problem* p);
/* The same for several species (S:186-241 with num > 1; the reference ships such inputs as "(UNSUPPORTED)" examples, but
 * its Equation handles them, S:454-511).  sigma_w, d_prime: species*species row-major; sigma_m, b, d: species.  Grids
 * come from the largest sigmas (S:203).  State vectors are [N_0..N_{s-1} | C_00 | C_01 | ... | C_{s-1,s-1}], each C_ij
 * G*G row-major (S:255-262, T:740-806).  A handle on such a problem supports ldm_rhs, ldm_solve (LDM_RECORD_LAST /
 * VALUES / ALL), ldm_set_algorithm and ldm_set_extrapolation; the other entry points return LDM_ERR_UNSUPPORTED. */
LDM_API int  ldm_problem_generate_multi(int species, const double* sigma_w, const double* sigma_m, const double* b, const double* d,
                                const double* d_prime, int forced_G, ldm_problem** out);
/* Properties::CheckKernels (S:147-183): the two midpoint-rule integrals (each must be 1 +- 1e-8) */
LDM_API int  ldm_problem_check_kernels(const ldm_problem* p, double* integral_w, double* integral_m);

/* Structure of the m table that LDM_ALGO_PREFIX needs (host only): separable inside its disc, i.e.
 * M[2b+1][2a+1] = q[b] q[a] on a staircase region and 0 beyond, as every table of GenerateKernel (S:113-123) is.
 * *separable = 1/0; *max_rel_dev = largest |M - q q^T| / (q q^T) over the non-zero quadrature entries. */
LDM_API int  ldm_problem_kernel_structure(const ldm_problem* p, int* separable, double* max_rel_dev);

/* ---- Equation(Properties) (S:413): upload tables, build the time-invariant index/weight tables ---- */
LDM_API int  ldm_create(const ldm_problem* p, int device, ldm_handle* out);
LDM_API void ldm_destroy(ldm_handle h);
LDM_API int  ldm_set_algorithm(ldm_handle h, int algo);
LDM_API int  ldm_get_algorithm(ldm_handle h);                                /* the algorithm in force (never LDM_ALGO_AUTO) */

/* ExtrapolationType (S:28) that Equation::operator() and the solvers are called with: what the pair density is taken
 * to be beyond the cutoff (S:355-389).  Every experiment of the reference passes SquaredN (the default).  Zero and
 * Alpha follow the reference to the letter -- including Alpha's integer division (S:377), which makes it C[0][0] -- and
 * so does the asymptote test of S:357 that precedes them.  Exponential (S:392-408) depends on every sample's own
 * distance through a decay rate that changes each step, so it costs one enumeration of the outside samples per
 * right-hand side (three more launches per step, no persistent kernel, no batching); one species, G >= 5. */
#define LDM_EXTRAPOLATION_ZERO        0
#define LDM_EXTRAPOLATION_SQUARED_N   1
#define LDM_EXTRAPOLATION_ALPHA       2
#define LDM_EXTRAPOLATION_EXPONENTIAL 3
LDM_API int  ldm_set_extrapolation(ldm_handle h, int type);

/* ---- EXTENSION (no counterpart in the reference): the closures whose triple integral does NOT factorise.
 * The reference closes the third moment with the power-1 closure T(xi, xi') = C(xi) C(xi') / N (Law & Dieckmann 2000,
 * eq. 14.10), so its integral of w(xi') T(xi, xi') is C(xi) * WC / N with one scalar WC (S:265-274, S:483-488).  The
 * other closures of the same paper need, at EVERY node xi, the two integrals BASELINE.json's north_star names:
 *     (w*C)(xi) = integral of w(xi') C(xi + xi') dxi'          K(xi) = integral of w(xi') C(xi') C(xi + xi') dxi'
 *   LDM_CLOSURE_POWER1            C(xi) WC / N                                                   (the reference; default)
 *   LDM_CLOSURE_POWER2_SYMMETRIC  1/2 [ C(xi) WC / N + C(xi) (w*C)(xi) / N + K(xi) / N - N^3 integral of w ]
 *   LDM_CLOSURE_POWER3            C(xi) K(xi) / N^3                                              (Kirkwood superposition)
 * replace the term of S:486-487; everything else of Equation::operator() stays.  Same quadrature as the reference's WC
 * (MakeIntegralWithPairDensity, S:297-316, extent MaxNumForW, bilinear C, N^2 beyond the cutoff).  K(xi) is a
 * convolution with the time-varying table w(xi') C(xi'), which is not separable: it is evaluated sample by sample
 * (G*G*(mW-1)^2 gathers per right-hand side), whatever ldm_set_algorithm says about the m*C integral.  One species,
 * SquaredN extrapolation; such a handle steps with one launch per kernel and step (no persistent kernel) in ldm_rhs,
 * ldm_solve, ldm_solve_rk4, ldm_advance, ldm_solve_batch and -- xi-sharded: the rows of both integrals are independent
 * given C, as for m*C -- ldm_shard_step (the caller all-gathers); the fused exchange (ldm_shard_peer_*) carries the
 * reference's closure only.  At C == N^2 all three closures give the
 * reference's right-hand side; away from it there is no reference to compare with (oracle: ldm_oracle_closure.c). */
#define LDM_CLOSURE_POWER1            1
#define LDM_CLOSURE_POWER2_SYMMETRIC  2
#define LDM_CLOSURE_POWER3            3
LDM_API int  ldm_set_closure(ldm_handle h, int closure);
LDM_API int  ldm_get_closure(ldm_handle h);
/* the two integrals on their own at `state` (1+G*G doubles): G*G each, either may be NULL; works under any closure */
LDM_API int  ldm_closure_integrals(ldm_handle h, const double* state, double* w_conv_C, double* K);

/* ---- Equation::operator()(t, vals, type) (S:431-513): host state in, host derivative out ---- */
LDM_API int  ldm_rhs(ldm_handle h, const double* state, double* deriv);
/* the two integrals on their own (S:265-295): WC scalar, MC G*G (MC may be NULL) */
LDM_API int  ldm_integrals(ldm_handle h, const double* state, double* WC, double* MC);
/* the time-invariant kernel look-ups of S:469-476 / S:490-497 at the moment nodes (G*G each) */
LDM_API int  ldm_kernels_at_nodes(ldm_handle h, double* m_at_nodes, double* w_at_nodes);

/* ---- EulerSolver::Solve (S:557-592) ----
 * ldm_count_points reproduces `time += step; while (time < time_max)` (S:574-589).
 * ldm_solve: init/1+G*G doubles.  times: n_points.  values/derivs by `record`:
 *   LAST: values = 1+G*G (derivs, if non-NULL, = derivative at the last point);
 *   OBSERVABLES: values = n_points*2; VALUES: values = n_points*(1+G*G); ALL: derivs likewise.
 * Host buffers; the copies in and out are part of the call.  values / derivs in pinned, mapped host memory (cudaHostAlloc,
 * cudaHostRegister) are written by the kernel itself, point by point, in one launch for the whole solve; pageable buffers
 * receive the history through two device-side segments that are copied out under the computation. */
LDM_API int  ldm_count_points(double t_max, double dt);
LDM_API int  ldm_solve(ldm_handle h, const double* init, double dt, double t_max, int record,
               double* times, double* values, double* derivs, int* n_points);

/* ---- RungeKutta4thSolver::Solve (S:597-654; present in the reference, instantiated by none of its experiments) on
 * the same right-hand-side kernels.  ldm_count_points_rk4 reproduces its `while (time < time_max)` with two half
 * steps of the clock per iteration (S:627-635).  Full history: times n_points, values and derivs (either may be NULL)
 * n_points*(1+G*G).  Intermediate stages are not clamped, the closed step is (S:643-645). ---- */
LDM_API int  ldm_count_points_rk4(double t_max, double dt);
LDM_API int  ldm_solve_rk4(ldm_handle h, const double* init, double dt, double t_max, double* times, double* values, double* derivs,
                   int* n_points);

/* ---- device-resident stepping (state stays in HBM between calls) ---- */
LDM_API int  ldm_state_upload(ldm_handle h, const double* state);          /* 1+G*G doubles */
LDM_API int  ldm_state_download(ldm_handle h, double* state);
LDM_API int  ldm_advance(ldm_handle h, int n_steps, double dt);            /* n Euler steps, asynchronous */
LDM_API int  ldm_synchronize(ldm_handle h);
/* device time of the last ldm_advance / ldm_solve stepping loop in ms (CUDA events on the stream) */
LDM_API int  ldm_last_elapsed_ms(ldm_handle h, float* ms);
/* number of kernels launched by this handle since creation */
LDM_API long long ldm_launch_count(ldm_handle h);

/* ---- the sigma-sweep loop Experiments::CalculateWithVaryingSigma (S:940-963): n independent solves,
 * results slot-indexed (no shared push_back).  last_states[i] receives 1+G_i*G_i doubles (an entry of s species: its
 * whole state vector, s + s*s*G_i*G_i).  The handles may live on several devices (ldm_create's device argument): every
 * device advances its slice of the batch concurrently -- the one-species entries whose rows are one chunk inside ONE
 * persistent launch (a team of CTAs per set), the others with one launch per kernel and step for all of the device's
 * sets, entries of several species beside them as replayed CUDA graphs on their own streams;
 * ldm_device_count() = visible devices (0 without CUDA). ---- */
LDM_API int  ldm_device_count(void);
LDM_API int  ldm_solve_batch(int n, const ldm_handle* handles, const double* const* inits, double dt, double t_max,
                     double* const* last_states);

/* ---- xi-sharded stepping for one large problem: this rank owns moment-grid rows
 * x = rank, rank+world, ... and produces their updated C rows; the caller all-gathers.
 * ldm_shard_layout: rows_per_rank = ceil(G/world); the gather buffer holds world*rows_per_rank*G doubles,
 * rank-major (row x lives at slot (x%world)*rows_per_rank + x/world).
 * ldm_shard_step: d_C_gathered (device) is the full C in gathered layout at the old time; on return
 * (stream-ordered) d_C_shard_out (device, rows_per_rank*G) holds this rank's new rows and the handle's N is
 * advanced (every rank recomputes N redundantly, so no all-reduce is needed). */
LDM_API int  ldm_shard_configure(ldm_handle h, int rank, int world);
LDM_API int  ldm_shard_rows_per_rank(ldm_handle h);
LDM_API int  ldm_shard_step(ldm_handle h, const double* d_C_gathered, double* d_C_shard_out, double dt, void* cuda_stream);
LDM_API int  ldm_shard_set_N(ldm_handle h, double N);
LDM_API int  ldm_shard_get_N(ldm_handle h, double* N);

/* ---- xi-sharded stepping with the all-gather fused into the update (NVLink peer stores, no NCCL on the data path).
 * One process per GPU.  ldm_shard_peer_alloc: after ldm_shard_configure; allocates this rank's two replicas of C
 * (gathered layout) and its flag array, returns three 64-byte CUDA IPC handles.  Exchange the handles of all ranks
 * out of band (e.g. torch.distributed.all_gather), then ldm_shard_peer_connect maps every peer's buffers.
 * ldm_shard_peer_upload: the full C (gathered layout, host) and N into this rank's replica 0; follow with a
 * host-side barrier across ranks.  ldm_shard_peer_step: waits (on the device) until every rank has published the
 * previous step, updates this rank's rows and stores them into EVERY rank's next replica, publishes the step.
 * ldm_shard_peer_download: this rank's current replica.  Returns LDM_ERR_CUDA if a peer never showed up. ---- */
#define LDM_IPC_HANDLE_BYTES 64
LDM_API int  ldm_shard_peer_alloc(ldm_handle h, unsigned char* handles /* 3 * LDM_IPC_HANDLE_BYTES */);
LDM_API int  ldm_shard_peer_connect(ldm_handle h, const unsigned char* all_handles /* world * 3 * LDM_IPC_HANDLE_BYTES, rank-major */);
LDM_API int  ldm_shard_peer_upload(ldm_handle h, const double* C_gathered, double N);
LDM_API int  ldm_shard_peer_step(ldm_handle h, double dt, void* cuda_stream);          /* = ldm_shard_peer_advance(h, 1, ...) */
/* n_steps Euler steps, asynchronous on cuda_stream.  On the prefix-sum path all of them run inside ONE launch per rank
 * (a persistent kernel: rows are stored into every rank's next replica as they close, a step ends with per-rank epoch
 * flags instead of a barrier); every rank must make the same call.  A rank that never shows up turns into LDM_ERR_CUDA
 * at the next ldm_shard_peer_download / ldm_synchronize (bounded waits on the device, ~9 s), not into a hang. */
LDM_API int  ldm_shard_peer_advance(ldm_handle h, int n_steps, double dt, void* cuda_stream);
/* The same peer group inside ONE process that drives several devices (or, for tests, several ranks on one device):
 * handles[r] is configured as rank r of `world`; allocates the replicas and wires them up through peer access --
 * replaces ldm_shard_peer_alloc + the exchange of IPC handles + ldm_shard_peer_connect. */
LDM_API int  ldm_shard_peer_connect_local(int world, const ldm_handle* handles);
LDM_API int  ldm_shard_peer_download(ldm_handle h, double* C_gathered, double* N);

/* ---- test hook: forces shapes the library would otherwise choose itself, so that tests reach the chunked and tiled code
 * paths on small grids.  Read when a handle is created.  Keys: "scan_bn", "scan_nt" (prefix-sum path: chunk width, threads),
 * "factored_bm", "factored_bn", "factored_tn" (dense contraction tiles), "persistent" (0: one launch per kernel and
 * step), "scan" (0: never the prefix-sum path), "chunk_grid" (CTAs of the persistent work-list kernel), "batch_capacity" (CTA budget of
 * the persistent batch kernel), "item_order" (work-list kernel, measurements: 0 rows ascending, 1 chunk by chunk, 2 rows descending),
 * "closure_tiles" (closure integrals: 2 = a lane's offsets held in registers, two tiles of 32; 0 = read per sample),
 * "reset". ---- */
LDM_API int  ldm_tuning_set(const char* name, int value);

/* ---- measurement support: FP64 (DFMA) peak of `device` in TFLOP/s, measured live by a register-resident
 * FMA chain; MEASURED_PEAKS.json has no FP64 figure, this is the roofline denominator bench.py reports. ---- */
LDM_API int  ldm_measure_fp64_peak(int device, double* tflops, double* reserved);
/* the two resources that bound the prefix-sum path instead, measured live: the load/store data pipe in
 * wavefronts per second over the whole device (conflict-free shared-memory loads; one wavefront = up to 128 bytes per
 * clock and SM), and the L2 -> SM read bandwidth in GB/s (16-byte loads bypassing L1 over an L2-resident buffer) */
LDM_API int  ldm_measure_memory_pipes(int device, double* wavefronts_per_s, double* l2_gbs);

LDM_API const char* ldm_last_error(void);
LDM_API const char* ldm_version(void);

#ifdef __cplusplus
}
#endif
#endif
