// Hessian of 2*NLL at a batch of points: central differences of the ANALYTIC gradient.
//
// parameter uncertainties / correlations.  In the reference they
// exist only behind iminuit (optimize/opt_minuit.py:136-152: covariance from MIGRAD/HESSE,
// errordef = 1 on the objective 2*NLL, i.e. covariance = 2 * inverse(Hessian of 2*NLL));
// optimize/mixins.py:83-120 only reshapes them.  Here every column of the Hessian is one more row
// of a K2 batch: 2 * P gradient evaluations per point, all points of a chunk in one launch.
//
//   H[i][j] = d(grad_i 2NLL)/dx_j ~ (g_i(x + h_j e_j) - g_i(x - h_j e_j)) / (x+ - x-)_j ,
//   h_j = 1e-4 * max(1, |x_j|); one-sided where the step would leave [lo_j, hi_j]
//   (gamma >= 1e-10 and the POI on its bound after a q~ fit); symmetrised; rows and columns of
//   fixed parameters are those of the identity.
#include <algorithm>

#include "hf_internal.cuh"

namespace {

constexpr double kHessStep = 1e-4;

// points of entry b: b * 2P + 2j (plus), b * 2P + 2j + 1 (minus)
__global__ void hess_points_kernel(int P, const double* __restrict__ pars, int64_t nb,
                                   const double* __restrict__ lo, const double* __restrict__ hi,
                                   const uint8_t* __restrict__ fixed, int64_t b0, int per_point,
                                   double* __restrict__ XP, int* __restrict__ row_idx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npts = nb * 2 * P;
    if (idx >= npts * P) return;
    const int64_t pt = idx / P;
    const int p = (int)(idx - pt * P);
    const int64_t b = pt / (2 * P);
    const int c = (int)(pt - b * 2 * P);
    const int j = c >> 1;
    double v = pars[(b0 + b) * P + p];
    if (p == j && !fixed[j]) {
        const double h = kHessStep * fmax(1.0, fabs(v));
        const double t = (c & 1) ? v - h : v + h;
        if (t >= lo[j] && t <= hi[j]) v = t;  // else: stay -> one-sided difference
    }
    XP[idx] = v;
    if (p == 0) row_idx[pt] = per_point ? (int)(b0 + b) : 0;
}

__global__ void hess_assemble_kernel(int P, int64_t nb, const double* __restrict__ XP,
                                     const double* __restrict__ GP, const uint8_t* __restrict__ fixed,
                                     int64_t b0, double* __restrict__ H) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t PP = (int64_t)P * P;
    if (idx >= nb * PP) return;
    const int64_t b = idx / PP;
    const int64_t e = idx - b * PP;
    const int i = (int)(e / P), j = (int)(e - (int64_t)i * P);
    double v;
    if (fixed[i] || fixed[j]) {
        v = (i == j) ? 1.0 : 0.0;
    } else {
        const double* xp = XP + b * 2 * PP;  // [2P][P]
        const double* gp = GP + b * 2 * PP;
        auto entry = [&](int r, int c) {   // d g_r / d x_c   (GP holds grad of ln L: factor -2)
            const double dx = xp[(int64_t)(2 * c) * P + c] - xp[(int64_t)(2 * c + 1) * P + c];
            return -2.0 * (gp[(int64_t)(2 * c) * P + r] - gp[(int64_t)(2 * c + 1) * P + r]) / dx;
        };
        v = 0.5 * (entry(i, j) + entry(j, i));
    }
    H[(b0 + b) * PP + e] = v;
}

}  // namespace

extern "C" int hf_hessian(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                          int64_t B, const double* lo, const double* hi, const uint8_t* fixed,
                          double* hess, void* stream) {
    HfRange nvtx_range("hf_hessian");
    if (!plan || !pars || !data || !lo || !hi || !fixed || !hess) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (B <= 0) return HF_OK;
    if (data_rows != 1 && data_rows != B) {
        hf_set_error("data_rows must be 1 or B");
        return HF_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int P = plan->P;
    // chunk so that one K2 launch sees at most ~64k points
    const int64_t per = 2 * (int64_t)P;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)1 << 16) / per));
    const size_t npts = (size_t)chunk * per;
    // workspace kept in the plan (grown on demand, freed with the plan): XP | GP | VP | ridx
    const size_t need = sizeof(double) * (2 * npts * P + npts) + sizeof(int) * npts + 1024;
    if (need > plan->hess_bytes) {
        if (plan->hess_block) cudaFree(plan->hess_block);
        plan->hess_block = nullptr;
        plan->hess_bytes = 0;
        if (cudaMalloc(&plan->hess_block, need) != cudaSuccess) {
            hf_set_error("cudaMalloc(%zu) for the Hessian workspace failed", need);
            return HF_ERR_NOMEM;
        }
        plan->hess_bytes = need;
    }
    double* XP = reinterpret_cast<double*>(plan->hess_block);
    double* GP = XP + npts * P;
    double* VP = GP + npts * P;
    int* ridx = reinterpret_cast<int*>(VP + npts);
    int rc = HF_OK;
    for (int64_t b0 = 0; b0 < B && rc == HF_OK; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        const int64_t np = nb * per;
        hess_points_kernel<<<(unsigned)((np * P + 255) / 256), 256, 0, st>>>(
            P, pars, nb, lo, hi, fixed, b0, data_rows == 1 ? 0 : 1, XP, ridx);
        plan->launches++;
        rc = hf_eval_launch(plan, XP, data, data_rows, np, VP, GP, st, ridx);
        if (rc) break;
        hess_assemble_kernel<<<(unsigned)((nb * (int64_t)P * P + 255) / 256), 256, 0, st>>>(
            P, nb, XP, GP, fixed, b0, hess);
        plan->launches++;
    }
    if (rc) return rc;
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

// Closed-form Hessian of 2*NLL (hf_hess_analytic.cu) at B points: mode 2 = exact (on the given data),
// mode 1 = Fisher (on the Asimov data of each point).  HF_ERR_UNSUPPORTED when the plan does not
// qualify (clipping, > 8 samples, > 120 histosys, per-bin parameters acting on several bins).
extern "C" int hf_hessian_analytic(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                                   int64_t B, const uint8_t* fixed, int32_t mode, double* hess, void* stream) {
    if (!plan || !pars || !data || !hess) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (B <= 0) return HF_OK;
    if (data_rows != 1 && data_rows != B) {
        hf_set_error("data_rows must be 1 or B");
        return HF_ERR_INVALID;
    }
    if (mode != 1 && mode != 2) {
        hf_set_error("mode must be 1 (Fisher) or 2 (exact)");
        return HF_ERR_INVALID;
    }
    if (!plan->d.ah_ok) {
        hf_set_error("plan does not qualify for the analytic Hessian");
        return HF_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = hf_hess_analytic_scratch_doubles(plan);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(((size_t)1 << 29) / (sizeof(double) * stride))));
    const size_t need = sizeof(double) * stride * (size_t)chunk + 1024;
    if (need > plan->hess_bytes) {
        if (plan->hess_block) cudaFree(plan->hess_block);
        plan->hess_block = nullptr;
        plan->hess_bytes = 0;
        if (cudaMalloc(&plan->hess_block, need) != cudaSuccess) {
            hf_set_error("cudaMalloc(%zu) for the Hessian workspace failed", need);
            return HF_ERR_NOMEM;
        }
        plan->hess_bytes = need;
    }
    const int P = plan->P;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        int rc = hf_hess_analytic_launch(plan, pars + b0 * P, nullptr, nullptr, (int)nb, nullptr, mode,
                                         data, data_rows, nullptr, b0, hess + b0 * (int64_t)P * P, 0,
                                         reinterpret_cast<double*>(plan->hess_block), fixed, st);
        if (rc) return rc;
    }
    return HF_OK;
}
