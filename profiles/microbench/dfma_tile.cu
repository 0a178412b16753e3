// scratch micro-benchmark: DFMA throughput of register-tile outer products (no memory traffic)
#include <cstdio>
#include <cuda_runtime.h>
template <int TM, int TN>
__global__ void __launch_bounds__(256) tile_kernel(double* sink, double s, int iters)
{
    double acc[TM][TN], a[TM], b[TN];
    for (int i = 0; i < TM; i++) a[i] = s + threadIdx.x * 1e-9 + i;
    for (int j = 0; j < TN; j++) b[j] = s * 0.5 + j;
    for (int i = 0; i < TM; i++) for (int j = 0; j < TN; j++) acc[i][j] = 0;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            #pragma unroll
            for (int i = 0; i < TM; i++)
                #pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            #pragma unroll
            for (int i = 0; i < TM; i++) a[i] += 1e-9;      // cheap operand change, keeps the compiler honest
            #pragma unroll
            for (int j = 0; j < TN; j++) b[j] -= 1e-9;
        }
    }
    double t = 0;
    for (int i = 0; i < TM; i++) for (int j = 0; j < TN; j++) t += acc[i][j];
    if (t == 1.2345) sink[0] = t;
}
template <int TM, int TN>
void run(const char* name, int warps_per_sm)
{
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, threads = 256, blocks = 148 * warps_per_sm / 8;
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        tile_kernel<TM, TN><<<blocks, threads>>>(sink, 1.0, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    double fma = (double)blocks * threads * iters * 8 * (TM * TN);
    double extra = (double)blocks * threads * iters * 8 * (TM + TN);
    printf("%-10s warps/SM %2d: %.2f TFLOP/s FMA-only (%.2f incl. the DADD operand updates)\n", name, warps_per_sm, 2 * fma / best / 1e9, (2 * fma + extra) / best / 1e9);
}
int main()
{
    run<5, 8>("5x8", 8); run<5, 8>("5x8", 16); run<5, 4>("5x4", 16); run<8, 8>("8x8", 8); run<4, 4>("4x4", 32); run<2, 8>("2x8", 32); run<10, 4>("10x4", 8);
    return 0;
}
