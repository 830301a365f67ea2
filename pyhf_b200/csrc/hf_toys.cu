// K4 Philox toy sampler, K5 test-statistic epilogue, p-value counting and the fp64 pipe
// micro-benchmark used as the roofline denominator.
//
// K4 replaces Simultaneous.sample (probability.py:263-274) / scipy.stats.poisson(rate).rvs,
// norm(loc, scale).rvs (tensor/numpy_backend.py:28-31, :43-47).  Counter-based Philox-4x32-10:
// element (toy i, column j, draw k) owns the counter (i_lo, i_hi, j, k) under key = seed, so the
// toys of a job do not depend on how they are sharded over GPUs.
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include "hf_internal.cuh"

namespace {

struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

__host__ __device__ inline void philox_round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c[0] = hi1 ^ c[1] ^ k[0];
    c[1] = lo1;
    c[2] = hi0 ^ c[3] ^ k[1];
    c[3] = lo0;
}

__host__ __device__ inline void philox4x32_10(uint64_t seed, uint64_t i, uint32_t j, uint32_t d,
                                              uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), j, d};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// two uniforms in (0, 1) with 53 random bits each
__host__ __device__ inline void uniform2(uint64_t seed, uint64_t i, uint32_t j, uint32_t d,
                                         double& u0, double& u1) {
    uint32_t r[4];
    philox4x32_10(seed, i, j, d, r);
    const uint64_t a = ((uint64_t)r[0] << 32) | r[1];
    const uint64_t b = ((uint64_t)r[2] << 32) | r[3];
    u0 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    u1 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ double sample_normal(uint64_t seed, uint64_t i, uint32_t j) {
    double u0, u1;
    uniform2(seed, i, j, 0u, u0, u1);
    return sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);  // Box-Muller
}

// Poisson: product-of-uniforms for small rates, Hoermann's PTRS transformed rejection otherwise
__device__ double sample_poisson(double lam, uint64_t seed, uint64_t i, uint32_t j) {
    if (!(lam > 0.0)) return (lam == 0.0) ? 0.0 : NAN;
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        double prod = 1.0, x = 0.0;
        for (uint32_t d = 0;; ++d) {
            double u0, u1;
            uniform2(seed, i, j, d, u0, u1);
            prod *= u0;
            if (!(prod > enlam)) return x;
            x += 1.0;
            prod *= u1;
            if (!(prod > enlam)) return x;
            x += 1.0;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t d = 0;; ++d) {
        double U, V;
        uniform2(seed, i, j, d, U, V);
        U -= 0.5;
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0))
            return k;
    }
}

// column kinds of the data vector for one parameter point: 0 Poisson(mean), 1 Normal(mean, sigma)
__global__ void toy_columns_kernel(DevPlan pl, int* __restrict__ kind, double* __restrict__ sigma) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pl.D) return;
    int k = 0;
    double sg = 0.0;
    if (j >= pl.NB) {
        const int a = j - pl.NB;
        k = 2;  // aux column without a constraint term: reproduced as-is (mean)
        for (int i = 0; i < pl.NG; ++i)
            if (pl.g_aux[i] == a) { k = 1; sg = pl.g_sigma[i]; }
        for (int i = 0; i < pl.NPO; ++i)
            if (pl.p_aux[i] == a) k = 0;
    }
    kind[j] = k;
    sigma[j] = sg;
}

__global__ void toy_sample_kernel(int D, const int* __restrict__ kind, const double* __restrict__ mean,
                                  const double* __restrict__ sigma, uint64_t seed, uint64_t offset,
                                  int64_t N, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * D) return;
    const int64_t i = idx / D;
    const int j = (int)(idx - i * D);
    const double m = mean[j];
    const int k = kind[j];
    double v;
    if (k == 0) v = sample_poisson(m, seed, offset + (uint64_t)i, (uint32_t)j);
    else if (k == 1) v = m + sigma[j] * sample_normal(seed, offset + (uint64_t)i, (uint32_t)j);
    else v = m;
    out[idx] = v;
}

__global__ void sample_poisson_kernel(const double* __restrict__ rate, int64_t M, int64_t N,
                                      uint64_t seed, uint64_t offset, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * M) return;
    const int64_t i = idx / M;
    const int64_t j = idx - i * M;
    out[idx] = sample_poisson(rate[j], seed, offset + (uint64_t)i, (uint32_t)j);
}

__global__ void sample_normal_kernel(const double* __restrict__ loc, const double* __restrict__ scale,
                                     int64_t M, int64_t N, uint64_t seed, uint64_t offset,
                                     double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * M) return;
    const int64_t i = idx / M;
    const int64_t j = idx - i * M;
    out[idx] = loc[j] + scale[j] * sample_normal(seed, offset + (uint64_t)i, (uint32_t)j);
}

// K5: infer/test_statistics.py:16-60, :507-520
__global__ void teststat_kernel(const double* __restrict__ ffix, const double* __restrict__ ffree,
                                const double* __restrict__ muhat, double mu, int mode, int64_t N,
                                double* __restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = fmax(ffix[i] - ffree[i], 0.0);  // clip(..., 0, None), test_statistics.py:56
    if (mode == 0 && muhat[i] > mu) v = 0.0;   // :30-32
    if (mode == 1 && muhat[i] < 0.0) v = 0.0;  // :516-520
    q[i] = v;
}

// counts[j] = #{i : samples[i] >= values[j]}  (calculators.py:632-640)
__global__ void pvalue_counts_kernel(const double* __restrict__ samples, int64_t N,
                                     const double* __restrict__ values, int64_t M,
                                     unsigned long long* __restrict__ counts) {
    __shared__ double tile[1024];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double v = (j < M) ? values[j] : 0.0;
    unsigned long long c = 0;
    for (int64_t base = (int64_t)blockIdx.y * 1024; base < N; base += (int64_t)gridDim.y * 1024) {
        const int n = (int)min((int64_t)1024, N - base);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x) tile[t] = samples[base + t];
        __syncthreads();
        for (int t = 0; t < n; ++t) c += (tile[t] >= v) ? 1ull : 0ull;
    }
    if (j < M && c) atomicAdd(&counts[j], c);
}

// same counts against samples sorted ascending: N - lower_bound(value); O(M log N) instead of the
// O(N M) sweep above (the reference is O(N^2) when every b-only toy gets its own p-values,
// calculators.py:955-973)
__global__ void pvalue_counts_sorted_kernel(const double* __restrict__ sorted, int64_t N,
                                            const double* __restrict__ values, int64_t M,
                                            long long* __restrict__ counts) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const double v = values[j];
    int64_t lo = 0, hi = N;  // first index with sorted[i] >= v
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] >= v) hi = mid; else lo = mid + 1;
    }
    counts[j] = isnan(v) ? 0 : (N - lo);
}

// fp64 FMA pipe peak: 16 independent chains per thread, no memory traffic
__global__ void fp64_peak_kernel(double* __restrict__ out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + 1e-3 * (threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

int hf_sample_toys_launch(hf_plan* plan, const double* pars, uint64_t seed, uint64_t offset,
                          int64_t N, double* out, cudaStream_t st) {
    if (N <= 0) return HF_OK;
    const DevPlan& pl = plan->d;
    const int D = pl.D;
    // scratch: mean[D] | sigma[D] | kind[D] in the staging block
    const size_t bytes = sizeof(double) * 2 * (size_t)D + sizeof(int) * (size_t)D + 512;
    if (plan->stage_bytes < bytes) {
        if (plan->stage_block) cudaFree(plan->stage_block);
        plan->stage_block = nullptr;
        plan->stage_bytes = 0;
        cudaError_t e = cudaMalloc(&plan->stage_block, bytes);
        if (e != cudaSuccess) {
            hf_set_error("cudaMalloc for the sampler scratch: %s", cudaGetErrorString(e));
            return HF_ERR_NOMEM;
        }
        plan->stage_bytes = bytes;
    }
    double* mean = (double*)plan->stage_block;
    double* sigma = mean + D;
    int* kind = (int*)(sigma + D);
    int rc = hf_expected_launch(plan, pars, 1, mean, st);
    if (rc) return rc;
    toy_columns_kernel<<<(D + 127) / 128, 128, 0, st>>>(pl, kind, sigma);
    const int64_t n = N * D;
    toy_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D, kind, mean, sigma, seed, offset, N, out);
    plan->launches += 2;
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

extern "C" {

int hf_sample_poisson(const double* rate, int64_t M, int64_t N, uint64_t seed, uint64_t offset,
                      double* out, void* stream) {
    if (!rate || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    const int64_t n = N * M;
    if (n <= 0) return HF_OK;
    sample_poisson_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rate, M, N, seed, offset, out);
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_sample_normal(const double* loc, const double* scale, int64_t M, int64_t N, uint64_t seed,
                     uint64_t offset, double* out, void* stream) {
    if (!loc || !scale || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    const int64_t n = N * M;
    if (n <= 0) return HF_OK;
    sample_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(loc, scale, M, N, seed, offset, out);
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_teststat(const double* fun_fixed, const double* fun_free, const double* muhat, double mu,
                int32_t mode, int64_t N, double* q, void* stream) {
    HfRange nvtx_range("K5 hf_teststat");
    if (!fun_fixed || !fun_free || !muhat || !q) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (N <= 0) return HF_OK;
    teststat_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fun_fixed, fun_free, muhat, mu, mode, N, q);
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_pvalue_counts(const double* samples, int64_t N, const double* values, int64_t M,
                     int64_t* counts, void* stream) {
    if (!samples || !values || !counts) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (M <= 0) return HF_OK;
    HF_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * M, st));
    if (N <= 0) return HF_OK;
    const unsigned gx = (unsigned)((M + 255) / 256);
    unsigned gy = (unsigned)std::min<int64_t>((N + 1023) / 1024, std::max<int64_t>(1, 2368 / gx));
    pvalue_counts_kernel<<<dim3(gx, gy), 256, 0, st>>>(samples, N, values, M, (unsigned long long*)counts);
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_pvalue_counts_sorted(const double* sorted_samples, int64_t N, const double* values, int64_t M,
                            int64_t* counts, void* stream) {
    if (!sorted_samples || !values || !counts) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (M <= 0) return HF_OK;
    pvalue_counts_sorted_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        sorted_samples, N, values, M, (long long*)counts);
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_fp64_peak(double* tflops_out, void* stream) {
    if (!tflops_out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d_out = nullptr;
    const int threads = 256, blocks = sms * 8, iters = 4096;
    HF_CUDA_OK(cudaMalloc(&d_out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(d_out, 256, 1.0);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fp64_peak_kernel<<<blocks, threads, 0, st>>>(d_out, iters, 1.0 + rep);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * threads * (double)blocks;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    HF_CUDA_OK(cudaGetLastError());
    *tflops_out = best;
    return HF_OK;
}

}  // extern "C"
