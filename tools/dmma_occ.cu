// DMMA throughput vs warps/SM and independent accumulator chains per warp
#include <cstdio>
#include <cuda_runtime.h>
template <int NCHAIN>
__global__ void dmma(double* out, int iters) {
    double c[NCHAIN][2];
    for (int i = 0; i < NCHAIN; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}
template <int NCHAIN>
double run(int blocks, int threads) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    dmma<NCHAIN><<<blocks, threads>>>(d, 64);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); dmma<NCHAIN><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 16.0 * NCHAIN * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaFree(d); return best;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps : {4, 8, 12, 16}) {
        printf("1 CTA/SM x %2d warps: chains 2: %.2f | 5: %.2f | 8: %.2f | 16: %.2f TF\n", warps,
               run<2>(sms, warps * 32), run<5>(sms, warps * 32), run<8>(sms, warps * 32), run<16>(sms, warps * 32));
    }
    return 0;
}
