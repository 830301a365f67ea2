// micro-benchmark: fp64 tensor-core (DMMA, mma.sync f64) vs DFMA throughput on this GPU.
// build+run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_peak.cu -o /tmp/dmma && /tmp/dmma
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_884(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}

__global__ void dmma_16816(double* out, int iters) {
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - (threadIdx.x + i) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1.2345) out[0] = s;
}

__global__ void dfma(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}

template <typename K>
double run(K k, int blocks, int threads, int iters, double flop_per_thread_iter) {
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, threads>>>(d, 64);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = flop_per_thread_iter * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaFree(d);
    return best;
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32, blocks = sms * 4;
        // m8n8k4: 8*8*4*2 flop per warp-instr = 512 flop/warp = 16/thread ; x8 per iter
        printf("warps/CTA %2d x4 CTA/SM: DFMA %.2f TF | DMMA m8n8k4 %.2f TF | DMMA m16n8k16 %.2f TF\n", warps,
               run(dfma, blocks, threads, 2048, 2.0 * 64),
               run(dmma_884, blocks, threads, 2048, 8 * 16.0),
               run(dmma_16816, blocks, threads, 2048, 4 * (16.0 * 8 * 16 * 2 / 32)));
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
