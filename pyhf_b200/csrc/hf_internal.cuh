// Internal definitions shared by the kernels of libhf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/hf_b200.h"

#define HF_MAX_PT 16

// Device-resident plan tables (all pointers device memory, immutable after hf_plan_create).
struct DevPlan {
    int P, NB, S, NAUX, NSEG, H, R, hs_code, NSF, K, NG, NPO, D, Hp;
    int has_clip_sample, has_clip_bin;
    double clip_sample, clip_bin;
    const double* nominal;       // [S][NB]
    const int* bin_seg;          // [NB]
    const int* hs_par;           // [H]
    const int* hs_row_sample;    // [R]
    const int* sample_hs_row;    // [S]  (-1: sample carries no histosys)
    const double* t1;            // [H][R][NB]
    const double* t2;            // [H][R][NB]
    const double* t1t;           // [R][NB][Hp]  (h fastest: backward sweep)
    const double* t2t;           // [R][NB][Hp]
    const int* sf_ptr;           // [S*NSEG+1]
    const int* sf_par;           // [NSF]
    const int* sf_kind;          // [NSF]
    const double* sf_coef;       // [NSF][8]
    const int* sfp_ptr;          // [P+1]   by-parameter CSR of scalar-factor entries
    const int* sfp_entry;        // [NSF]   entry index
    const int* sfp_q;            // [NSF]   q = sample*NSEG+seg of that entry
    const int* sfp_kind;         // [NSF]   kind, in by-parameter order
    const double* sfp_coef;      // [NSF][8] coefficients, in by-parameter order
    const int* bf_par;           // [K][S][NB]
    const int* g_par;            // [NG]
    const int* g_aux;            // [NG]
    const double* g_sigma;       // [NG]
    const int* p_par;            // [NPO]
    const int* p_aux;            // [NPO]
    const double* p_factor;      // [NPO]
    const int* par_cons_kind;    // [P] 0 none, 1 gauss, 2 poisson
    const int* par_cons_idx;     // [P] index into g_* / p_*
    // ---- DMMA path (hf_eval_mma.cu): tables re-tiled for bulk copies into shared memory
    const int* qg_first;         // [n_qg] first q of a group of consecutive (sample, segment) products
    const int* qg_n;             // [n_qg] 1..4 members with identical (parameter, kind) entry lists
    int n_qg;
    int max_sf_q, max_sf_par;    // longest scalar-factor entry list per (sample, segment) / per parameter
    int mma_ok;                  // plan fits the tensor-core kernel (0 < H, R <= 4, <= 2 other samples)
    int mma_ntiles;              // bin tiles of HF_MMA_BT bins
    int mma_nchunks;             // K chunks of HF_MMA_KC columns (column 2h = T1_h, 2h+1 = T2_h)
    int mma_hs_unique;           // every histosys modifier has its own parameter (plain gradient stores)
    int mma_bf_unique;           // every per-bin parameter acts on exactly one bin (plain gradient stores)
    int mma_nother;              // samples without histosys
    int mma_other[2];            // their sample indices
    const double* mma_tt;        // [ntiles][2 halves][nchunks][HF_MMA_KC rows at HF_MMA_ROW][64 cells] (cell = (bin_local % 16)*4 + row)
    // ---- analytic Hessian (hf_hess_analytic.cu)
    int ah_ok;                   // plan qualifies (else the minimiser differences the analytic gradient)
    int ah_nsfp;                 // distinct parameters among the scalar-factor entries
    int ah_nbl;                  // (bin, per-bin parameter) pairs
    const int* ah_sfp_list;      // [ah_nsfp] parameter index
    const int* ah_sf_j;          // [NSF] position of the entry's parameter in ah_sfp_list
    const int* ah_bl_ptr;        // [NB+1] CSR over bins of the per-bin parameters acting there
    const int* ah_bl_par;        // [ah_nbl]
    const unsigned* ah_bl_mask;  // [ah_nbl] bit s: the parameter multiplies sample s on that bin
    const int* ah_seg_first;     // [NSEG+1] first bin of a segment (bins of a segment are contiguous)
    const unsigned char* ah_par_mult;  // [P] the parameter multiplies rates (normfactor / lumi / per-bin factor)
};

#define HF_AH_SMAX 8    // samples the analytic-Hessian kernel keeps in registers
#define HF_AH_KLMAX 4   // per-bin parameters on one bin

#define HF_MMA_BT 32    // bins per tile
#define HF_MMA_NC 128   // cells per tile (4 rows per bin)
#define HF_MMA_CS 136   // row stride of the Delta / w*F tile [point][128 cells]; row p starts at HF_MMA_WFROW(p)
// Row p of the Delta / w*F tile starts at p*136 + 2*((p>>1)&3) doubles: eight consecutive rows then sit at
// the eight different 2-double slots mod 16 (epilogue: lane = point, 16-byte accesses at one column), rows
// 2m, 2m+1 eight doubles apart (A' operand of the backward product: lanes (g, t) read row 8hw+g, cells 8j+2t)
// and rows p, p+2 two slots apart (Delta store: lanes (g, t) write rows 2g+m at column 4t).
#define HF_MMA_WFROW(p) ((p) * HF_MMA_CS + 2 * (((p) >> 1) & 3))
#define HF_MMA_NCH 64   // cells of a half tile (16 bins): the unit one half of the CTA (4 warps) works on
#define HF_MMA_CSH 72   // row stride of a half-tile table chunk (== 8 mod 16 doubles); row r starts at HF_MMA_ROW(r)
// Row r of a table chunk starts at r*72 + 4*((r>>1)&1) doubles: the four rows k4*4 + t of a forward
// k-step then sit at bank offsets {0, 8, 4, 12} and the two rows 8ni + g, g in {0, 1}, of a backward
// quarter warp at {0, 8}, so the 16-byte fragment loads of BOTH products are conflict-free.
#define HF_MMA_ROW(r) ((r) * HF_MMA_CSH + 4 * (((r) >> 1) & 1))
#define HF_MMA_KC 40    // K columns per chunk (20 modifiers)
#define HF_MMA_M 32     // parameter points per CTA

// Per-call scratch, batch-major ("parameter-batch-major": point index fastest) so that warps
// whose lanes are points read/write coalesced.  ld = padded number of points.
struct Work {
    double* parsT;  // [P][ld]
    double* c1T;    // [H][ld]   histosys coefficient of T1
    double* c2T;    // [H][ld]   histosys coefficient of T2
    double* dc1T;   // [H][ld]
    double* dc2T;   // [H][ld]
    double* FsegT;  // [S*NSEG][ld]  product of scalar factors per (sample, segment)
    double* GsegT;  // [S*NSEG][ld]  sum_{b in seg} w_b r_{s,b}
    double* gradT;  // [P][ld]
    double* mainv;  // [ld]          main-term partial sums
    double* dconst; // [data_rows]   -sum lgamma(n+1) - sum ln(sigma sqrt(2pi)) ... (theta-independent)
};

// ---------------------------------------------------------------- interpolation primitives
// histosys: delta = c1*T1 + c2*T2.  code4p: interpolators/code4p.py:63-117; code0: code0.py:66-88.
__host__ __device__ inline void hf_hs_coeffs(int code, double a, double& c1, double& c2,
                                             double& dc1, double& dc2) {
    if (code == HF_HS_CODE4P) {
        c1 = a;
        dc1 = 1.0;
        if (a > 1.0) {
            c2 = 8.0 * a;
            dc2 = 8.0;
        } else if (a < -1.0) {
            c2 = -8.0 * a;
            dc2 = -8.0;
        } else {
            const double a2 = a * a;
            c2 = a2 * (a2 * (a2 * 3.0 - 10.0) + 15.0);
            dc2 = a * (a2 * (a2 * 18.0 - 40.0) + 30.0);
        }
    } else {
        const bool pos = a > 0.0;
        c1 = pos ? a : 0.0;
        c2 = pos ? 0.0 : a;
        // the derivative AT the kink alpha == 0 is the right-sided one: that is what the forward
        // differences of the reference's SLSQP (optimize/opt_scipy.py:96-105, jac=None) see from
        // suggested_init, and it decides which side of the kink a fit starts into (the reference's
        // validation model 2bin_2channel_coupledhistosys has a cusp maximum there)
        const bool dpos = a >= 0.0;
        dc1 = dpos ? 1.0 : 0.0;
        dc2 = dpos ? 0.0 : 1.0;
    }
}

// scalar factor f(theta) and dlog f/dtheta.  code4: interpolators/code4.py:178-239 (alpha>=1 up,
// alpha<=-1 down, 6th-order polynomial between); code1: code1.py:84-104 (alpha>0 up).
__host__ __device__ inline void hf_sf_eval(int kind, const double* __restrict__ cf, double t,
                                           double& f, double& dl) {
    if (kind == HF_SF_THETA) {
        f = t;
        dl = 1.0 / t;
        return;
    }
    const double lnu = cf[0], lnd = cf[1];
    if (kind == HF_SF_CODE4 && t > -1.0 && t < 1.0) {
        double poly = cf[7], dpoly = 0.0;
#pragma unroll
        for (int i = 6; i >= 2; --i) {
            dpoly = dpoly * t + poly;
            poly = poly * t + cf[i];
        }
        dpoly = dpoly * t + poly;
        poly = poly * t + 1.0;
        f = poly;
        dl = dpoly / poly;
        return;
    }
    // (code1 at its kink t == 0: f = 1 from either branch, derivative right-sided -- see hf_hs_coeffs)
    const bool up = (kind == HF_SF_CODE4) ? (t >= 1.0) : (t >= 0.0);
    if (up) {
        f = exp(t * lnu);
        dl = lnu;
    } else {
        f = exp(-t * lnd);
        dl = -lnd;
    }
}

// Branch-free accumulation of one scalar factor into (prod, logsum): F = prod * exp(logsum).
// Lanes are parameter points, so per-entry branches on alpha would diverge in every warp and
// each entry would pay for both the polynomial and the exp(); here the outside-range factors
// exp(+-alpha ln hi/lo) are summed in the log domain and ONE exp() per (sample, channel) is taken.
__device__ __forceinline__ void hf_sf_accum(int kind, const double* __restrict__ cf, double t,
                                            double& prod, double& logsum) {
    if (kind == HF_SF_THETA) {  // kind is uniform across the warp (same entry, different points)
        prod *= t;
        return;
    }
    const double lnu = cf[0], lnd = cf[1];
    if (kind == HF_SF_CODE4) {
        double poly = cf[7];
#pragma unroll
        for (int i = 6; i >= 2; --i) poly = fma(poly, t, cf[i]);
        poly = fma(poly, t, 1.0);
        const bool inside = (t > -1.0) && (t < 1.0);
        prod *= inside ? poly : 1.0;
        logsum += inside ? 0.0 : ((t >= 1.0) ? t * lnu : -t * lnd);
    } else {
        logsum += (t > 0.0) ? t * lnu : -t * lnd;
    }
}

// reciprocal of a positive normal number: MUFU.RCP64H seed + the refinement steps of the IEEE
// sequence, without its special-case slow path (a CALL inside the hot loop) -- full precision to
// the last ulp or two, which is what a gradient term needs
__device__ __forceinline__ double hf_rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// p ? a : b as ONE select instruction.  Written in PTX because the compiler otherwise turns a
// ternary whose arms are expensive into a branch, and lanes are parameter points: ~15 % of the
// alphas lie outside (-1, 1), so nearly every warp would run both arms one after the other.
__device__ __forceinline__ double hf_selp(bool p, double a, double b) {
    double r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.f64 %0, %1, %2, q;\n\t}"
        : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
    return r;
}

// dlog f / dtheta of a code4 normsys entry, branch-free (interpolators/code4.py:178-239)
__device__ __forceinline__ double hf_sf_dlog_code4(const double* __restrict__ cf, double t) {
    double poly = cf[7], dpoly = 0.0;
#pragma unroll
    for (int i = 6; i >= 2; --i) {
        dpoly = fma(dpoly, t, poly);
        poly = fma(poly, t, cf[i]);
    }
    dpoly = fma(dpoly, t, poly);
    poly = fma(poly, t, 1.0);
    const double in = dpoly * hf_rcp_pos(poly);
    const double out = hf_selp(t >= 1.0, cf[0], -cf[1]);
    return hf_selp((t > -1.0) && (t < 1.0), in, out);
}

// dlog f / dtheta of a normsys entry (kind != THETA)
__device__ __forceinline__ double hf_sf_dlog(int kind, const double* __restrict__ cf, double t) {
    if (kind == HF_SF_CODE4) return hf_sf_dlog_code4(cf, t);  // kind is uniform across the warp
    return (t >= 0.0) ? cf[0] : -cf[1];  // right-sided at the kink of code1
}

// main / constraint Poisson term without the theta-independent lgamma(n+1):
// xlogy(n, lam) - lam   (tensor/numpy_backend.py:435-436)
__host__ __device__ inline double hf_pois_term(double n, double lam) {
    const double xl = (n == 0.0 && !(lam != lam)) ? 0.0 : n * log(lam);
    return xl - lam;
}

// ---------------------------------------------------------------- host-side helpers
// NVTX range around a C-ABI entry point (K1..K5 of DESIGN.md): shows up in Nsight timelines and
// lets `ncu --nvtx --nvtx-include` select the kernels of one path
#include <nvtx3/nvToolsExt.h>
struct HfRange {
    explicit HfRange(const char* name) { nvtxRangePushA(name); }
    ~HfRange() { nvtxRangePop(); }
    HfRange(const HfRange&) = delete;
    HfRange& operator=(const HfRange&) = delete;
};
void hf_set_error(const char* fmt, ...);
#define HF_CUDA_OK(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            hf_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,         \
                         cudaGetErrorString(e__));                                    \
            return HF_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

struct hf_plan {
    DevPlan d;
    void* table_block = nullptr;  // one allocation holding every table
    // workspace
    Work w;
    void* work_block = nullptr;
    int64_t work_ld = 0;       // points capacity
    int64_t dconst_cap = 0;
    double* dconst = nullptr;
    // data constants computed once by a caller that evaluates the same data array many times
    // (the fit loop): hf_eval_launch skips hf_data_const_kernel when (data, rows) match
    const double* dconst_held_data = nullptr;
    int64_t dconst_held_rows = 0;
    const double* dconst_held = nullptr;
    int32_t* h_counters = nullptr;  // pinned host mirror of the fit loop's device counters
    // fit workspace
    void* fit_block = nullptr;
    size_t fit_bytes = 0;
    // per-fit ("solo") minimiser, hf_fit_solo.cu: gpos [P] | lpos [P] | glist [ng] | llist [nl] | queue head [1]
    int* solo_dev = nullptr;
    int solo_ng = 0, solo_nl = 0;
    int ah_max_lc = 0;  // most (bin, per-bin parameter) pairs inside one segment
    void* solo_scratch = nullptr;
    size_t solo_scratch_bytes = 0;
    int solo_last_grid = 0, solo_last_occ = 0, solo_last_smem = 0, solo_last_nt = 0, solo_last_lowrank = 0;  // launch shape of the last batch (bench.py)
    // workspace of hf_hessian
    void* hess_block = nullptr;
    size_t hess_bytes = 0;
    // staging for the *_host entry points
    void* stage_block = nullptr;
    size_t stage_bytes = 0;
    int64_t launches = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host-buffer entry points
    bool timing = false;
    std::vector<cudaEvent_t> ev;  // start/stop pairs of hf_main_kernel launches
    int device = 0;
    int sm_count = 148;
    // host copies of small tables the host logic needs
    int P, NB, S, NAUX, NSEG, H, R, NSF, K, NG, NPO, D;
    std::vector<int32_t> h_bf_par, h_hs_par, h_sf_par;
    // fit: finite-difference column structure (built lazily by hf_fit.cu)
    bool fd_ready = false;
    int fd_ncolors = 0;
    std::vector<int32_t> fd_color;    // [P]  colour of a per-bin ("local") parameter, -1 = global
    std::vector<int32_t> fd_partner;  // [P][ncolors] local parameter of that colour on the same bin
};

int hf_ensure_work(hf_plan* plan, int64_t B, int64_t data_rows);
// data row of point pt: row_idx[pt] if row_idx, else pt when data_rows == B, else 0
int hf_eval_launch(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                   int64_t B, double* out, double* grad, cudaStream_t st,
                   const int* row_idx = nullptr);
int hf_expected_launch(hf_plan* plan, const double* pars, int64_t B, double* out, cudaStream_t st);
int hf_data_const_launch(hf_plan* plan, const double* data, int64_t rows, double* out, cudaStream_t st);
// DMMA main kernel (hf_eval_mma.cu); returns HF_ERR_UNSUPPORTED when the plan does not qualify
int hf_main_mma_launch(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
                       const int* row_idx, bool grad, cudaStream_t st);
// analytic Hessian of 2*NLL (hf_hess_analytic.cu); request i: point X[list ? list[i] : i], matrix
// H[same row] (or the one shared matrix), mode 1 = Fisher / 2 = exact (per fit, or mode_all)
size_t hf_hess_analytic_scratch_doubles(const hf_plan* plan);
int hf_hess_analytic_launch(hf_plan* plan, const double* X, const int* list, const int* n_dev, int nreq,
                            const int* mode, int mode_all, const double* data, int64_t data_rows,
                            const int* data_map, int64_t toy0, double* Hout, int to_shared, double* scratch,
                            const uint8_t* fixed_struct, cudaStream_t st);
// per-fit minimiser (hf_fit_solo.cu): one CTA per fit, the whole Newton iteration on the device.
// *handled = 0 when the plan / model size is outside its scope (the lock-step driver runs instead)
bool hf_fit_solo_supported(hf_plan* plan);
int hf_fit_solo_launch(hf_plan* plan, const double* data, int64_t data_rows, const double* dconst,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* fixed, const uint8_t* fixed_alt, const uint8_t* use_alt, int n_alt,
                       const int32_t* data_map, int64_t N, const hf_fit_opts& opts, double* x_out, double* fun_out,
                       int32_t* status_out, int32_t* niter_out, cudaStream_t st, int* handled);
static inline int hf_mma_ks(int H) {  // padded coefficient row stride, == 12 mod 16 doubles
    int ks = 2 * H;
    ks = (ks + HF_MMA_KC - 1) / HF_MMA_KC * HF_MMA_KC;
    while (ks % 16 != 12) ++ks;
    return ks;
}
