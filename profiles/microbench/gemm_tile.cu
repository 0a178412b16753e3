// scratch: FP64 smem->register outer-product microkernel, operands from shared memory like factored_kernel's K loop
#include <cstdio>
#include <cuda_runtime.h>
template <int BM, int BN, int TM, int TN, int BK>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_kernel(double* sink, int chunks)
{
    constexpr int TYN = BM / TM, TXN = BN / TN, NT = TYN * TXN;
    __shared__ __align__(16) double Qs[2][BK * BM];
    __shared__ __align__(16) double Es[2][BK * BN];
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    for (int e = tid; e < 2 * BK * BM; e += NT) (&Qs[0][0])[e] = 1.0 + e * 1e-6;
    for (int e = tid; e < 2 * BK * BN; e += NT) (&Es[0][0])[e] = 0.5 + e * 1e-6;
    __syncthreads();
    double acc[TM][TN];
    for (int i = 0; i < TM; i++) for (int j = 0; j < TN; j++) acc[i][j] = 0;
    for (int c = 0; c < chunks; c++) {
        const double* qs = Qs[c & 1];
        const double* es = Es[c & 1];
        #pragma unroll
        for (int k = 0; k < BK; k++) {
            double a[TM], b[TN];
            #pragma unroll
            for (int i = 0; i < TM; i++) a[i] = qs[k * BM + i * TYN + ty];
            #pragma unroll
            for (int j = 0; j < TN; j += 2) {
                const double2 v = *reinterpret_cast<const double2*>(es + k * BN + (j / 2) * (2 * TXN) + tx * 2);
                b[j] = v.x; b[j + 1] = v.y;
            }
            #pragma unroll
            for (int i = 0; i < TM; i++)
                #pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double t = 0;
    for (int i = 0; i < TM; i++) for (int j = 0; j < TN; j++) t += acc[i][j];
    if (t == 1.2345) sink[0] = t;
}
template <int BM, int BN, int TM, int TN, int BK>
void run(const char* name, int ctas_per_sm)
{
    constexpr int NT = (BM / TM) * (BN / TN);
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int chunks = 400, blocks = 148 * ctas_per_sm;
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        gemm_kernel<BM, BN, TM, TN, BK><<<blocks, NT>>>(sink, chunks);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    double fma = (double)blocks * chunks * BK * BM * BN;
    printf("%-22s threads %3d, CTAs/SM %d: %.2f TFLOP/s (%s)\n", name, NT, ctas_per_sm, 2 * fma / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    run<80, 128, 5, 8, 8>("80x128 5x8 bk8", 1);
    run<80, 128, 5, 4, 8>("80x128 5x4 bk8", 1);
    run<80, 128, 8, 8, 8>("80x128 8x8 bk8", 1);
    run<80, 128, 10, 8, 8>("80x128 10x8 bk8", 1);
    run<80, 128, 10, 4, 8>("80x128 10x4 bk8", 1);
    run<80, 64, 5, 8, 16>("80x64 5x8", 2);
    run<80, 64, 10, 8, 8>("80x64 10x8", 2);
    run<40, 128, 5, 8, 16>("40x128 5x8", 2);
    run<80, 128, 8, 16, 8>("80x128 8x16 bk8", 1);
    return 0;
}
