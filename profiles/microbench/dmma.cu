// scratch: FP64 tensor-core (mma.sync m8n8k4 f64) throughput on B200
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* sink, int iters)
{
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-6, b = 0.5 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double t = 0;
    for (int i = 0; i < NACC; i++) t += c[i][0] + c[i][1];
    if (t == 1.2345) sink[0] = t;
}
template <int NACC>
void run(int warps_per_sm)
{
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148 * warps_per_sm / 8;
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        dmma_kernel<NACC><<<blocks, 256>>>(sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
    }
    double fma = (double)blocks * 8 * iters * NACC * 256.0;
    printf("DMMA m8n8k4: %d accumulators/warp, %2d warps/SM: %.2f TFLOP/s (%s)\n", NACC, warps_per_sm, 2 * fma / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<4>(8); run<8>(8); run<8>(16); run<16>(16); run<8>(32); return 0; }
