// micro-benchmark: DMMA issue rate of ONE warp per scheduler when its operands come from shared
// memory (the situation of a half of hf_main_mma_kernel running alone).  Cycles per DMMA, per variant.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_lds.cu -o /tmp/dmma_lds && /tmp/dmma_lds
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int CSH = 68, KC = 40, NCELL = 64;
// V = 0: backward loop of the kernel (1 A + 5 B LDS.64 per 5 DMMA)
// V = 1: the same with 16-byte loads (cells permuted: a lane owns two consecutive cells): 1 A + 5 B LDS.128 per 10 DMMA
// V = 2: register operands only
// V = 3: forward-like loop: 4 A + 2 B LDS.64 per 8 DMMA
// V = 4: forward-like with 16-byte loads: 2 A (transposed, 2 points each) + 1 B LDS.128 per 8 DMMA
template <int V>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    double* tb = sm;                 // [KC][CSH]
    double* wf = sm + KC * CSH;      // [32][136]
    for (int i = threadIdx.x; i < KC * CSH + 32 * 136 + 64 * 40; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3, g = lane >> 2, t = lane & 3;
    double acc[10][2];
    for (int i = 0; i < 10; ++i) acc[i][0] = acc[i][1] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {
            const double* tbp = tb + g * CSH + t;
            const double* ar = wf + (8 * warp + g) * 136 + t;
#pragma unroll
            for (int c4 = 0; c4 < NCELL / 4; ++c4) {
                const double a = ar[4 * c4];
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) dmma(acc[ni][0], acc[ni][1], a, tbp[8 * ni * CSH + 4 * c4]);
            }
        } else if (V == 1) {
            const double* tbp = tb + g * CSH + 2 * t;
            const double* ar = wf + (8 * warp + g) * 136 + 2 * t;
#pragma unroll
            for (int j = 0; j < NCELL / 8; ++j) {
                const double2 a = *reinterpret_cast<const double2*>(ar + 8 * j);
                double2 b[5];
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) b[ni] = *reinterpret_cast<const double2*>(tbp + 8 * ni * CSH + 8 * j);
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) dmma(acc[ni][0], acc[ni][1], a.x, b[ni].x);
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) dmma(acc[ni][0], acc[ni][1], a.y, b[ni].y);
            }
        } else if (V == 2) {
            double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
#pragma unroll
            for (int c4 = 0; c4 < NCELL / 4; ++c4)
#pragma unroll
                for (int ni = 0; ni < 5; ++ni) dmma(acc[ni][0], acc[ni][1], a, b);
        } else if (V == 3) {
            const double* tbp = tb + 16 * warp + g;
            const double* ca = wf + g * 136 + t;
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                double a[4], b[2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = ca[8 * mi * 136 + 4 * k4];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) b[ni] = tbp[(4 * k4 + t) * CSH + 8 * ni];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma(acc[2 * mi + ni][0], acc[2 * mi + ni][1], a[mi], b[ni]);
            }
        } else {
            const double* tbp = tb + 16 * warp + 2 * g;
            const double* ca = wf + 32 * 136 + t * 40 + 2 * g;  // [k][36+] transposed coefficients
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                const double2 a0 = *reinterpret_cast<const double2*>(ca + 4 * k4 * 40);
                const double2 a1 = *reinterpret_cast<const double2*>(ca + 4 * k4 * 40 + 16);
                const double2 b = *reinterpret_cast<const double2*>(tbp + (4 * k4 + t) * CSH);
                dmma(acc[0][0], acc[0][1], a0.x, b.x);
                dmma(acc[1][0], acc[1][1], a0.x, b.y);
                dmma(acc[2][0], acc[2][1], a0.y, b.x);
                dmma(acc[3][0], acc[3][1], a0.y, b.y);
                dmma(acc[4][0], acc[4][1], a1.x, b.x);
                dmma(acc[5][0], acc[5][1], a1.x, b.y);
                dmma(acc[6][0], acc[6][1], a1.y, b.x);
                dmma(acc[7][0], acc[7][1], a1.y, b.y);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 10; ++i) s += acc[i][0] + acc[i][1];
    if (s == 1.2345) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V>
void run(const char* name, int dmma_per_iter, int warps) {
    double* d; long long* c;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int smem = sizeof(double) * (KC * CSH + 32 * 136 + 64 * 40);
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200;
    k<V><<<148, warps * 32, smem>>>(d, c, 10);
    k<V><<<148, warps * 32, smem>>>(d, c, iters);
    long long h = 0;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %d warps/SM: %.2f cycles per DMMA per warp (%s)\n", name, warps, (double)h / ((double)iters * dmma_per_iter),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(c);
}
int main() {
    for (int warps : {4, 8}) {
        run<2>("registers only, 5 chains", 80, warps);
        run<0>("backward: 1 A + 5 B LDS.64 per 5 DMMA", 80, warps);
        run<1>("backward: 1 A + 5 B LDS.128 per 10 DMMA", 80, warps);
        run<3>("forward: 4 A + 2 B LDS.64 per 8 DMMA", 80, warps);
        run<4>("forward: 2 A + 1 B LDS.128 per 8 DMMA", 80, warps);
    }
    return 0;
}
