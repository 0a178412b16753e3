// The prefix-sum evaluation of the m*C integral (Solver.cpp:277-316) for sm_100a, one launch per step: scan_kernel and its
// batched form (blockIdx.z = parameter set or species pair), the choice of the CTA shape, and the launchers.  The device
// code itself -- scan_body -- lives in ldm_scan_body.cuh, shared with the persistent kernels of ldm_persist.cu.
#include "../../include/ldm.h"
#include "ldm_scan_body.cuh"

namespace ldm {

namespace {

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

void configure_once(const void* kernel, bool& done)
{
    if (!done) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
        done = true;
    }
}

}  // namespace

#ifdef LDM_PHASE_CLOCKS
extern "C" LDM_API int ldm_debug_scan_clocks(long long* out, int rows)
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
    if (const int v = tuning().scan_nt; (v == 256 || v == 512) && v >= bn) return (v / bn) * bn;      // ldm_tuning_set: tests
    if (bn <= 64 && scan_layout(G, A, bn, 256, piece_bound(G, 256)).end <= 48 * 1024) return (256 / bn) * bn;
    return (kScanThreadsMax / bn) * bn;                    // bn = 32, 64, 96, 128 -> 512, 512, 480, 512
}

int scan_pick_bn(int G, int A)
{
    auto fits = [&](int bn) {
        const int nt = scan_pick_threads(G, A, bn);
        return scan_smem_bytes(G, A, bn, nt, piece_bound(G, nt)) <= kSmemBudget;
    };
    if (const int v = tuning().scan_bn; v >= 32 && v <= 128 && v % 32 == 0 && fits(v)) return v;         // ldm_tuning_set: tests
    // widest chunk (a multiple of 32, at most 128 columns) whose H tile fits; then the evenest split of the row
    int bn = 128;
    while (bn > 32 && !fits(bn)) bn -= 32;
    if (!fits(bn)) return 0;                                     // table too tall for one CTA: dense contraction instead
    const int chunks = (G + bn - 1) / bn;
    const int even = (((G + chunks - 1) / chunks) + 31) & ~31;
    return even < bn ? even : bn;
}

int scan_lead_columns() { return kScanLead; }

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

}  // namespace ldm
