/*
 * hf_b200.h -- C-ABI of libhf_b200.so: the B200-native HistFactory hot path.
 *
 * The reference (scikit-hep/pyhf) is pure Python and has NO native interface for this path;
 * each entry point below names the reference function(s) it replaces (paths relative to
 * /root/reference/src/pyhf).  INTEGRATION.md shows the ctypes binding a pyhf maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every `dev` pointer is device memory owned by the caller (PyTorch: tensor.data_ptr());
 *     every `host` pointer is host memory (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all functions return 0 on success or a negative hf_status; hf_last_error() gives the text.
 *     Nothing throws across the ABI.
 *   - a plan owns its device tables and a scratch workspace: use a plan from one stream at a time
 *     (one process per GPU, one Python thread per process is the intended deployment).
 *   - all arithmetic is IEEE fp64.
 */
#ifndef HF_B200_H
#define HF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HF_OK = 0,
    HF_ERR_INVALID = -1, /* bad argument / inconsistent plan description */
    HF_ERR_CUDA = -2,    /* CUDA runtime error (text in hf_last_error) */
    HF_ERR_NOMEM = -3,
    HF_ERR_UNSUPPORTED = -4
} hf_status;

/* scalar-factor entry kinds (hf_plan_desc.sf_kind) */
#define HF_SF_THETA 0 /* normfactor / lumi : f = theta        (modifiers/normfactor.py:94-115, lumi.py:92-107) */
#define HF_SF_CODE4 1 /* normsys code4     : interpolators/code4.py:172-239 */
#define HF_SF_CODE1 2 /* normsys code1     : interpolators/code1.py:81-104 */
/* histosys interpolation (hf_plan_desc.hs_code) */
#define HF_HS_CODE0 0  /* interpolators/code0.py:66-88   : T1 = hi-nom, T2 = nom-lo */
#define HF_HS_CODE4P 4 /* interpolators/code4p.py:63-117 : T1 = S, T2 = A */

/*
 * Host-side description of one HistFactory model (what pyhf_b200.plan.extract_plan reads off a
 * live pyhf.Model).  All pointers are HOST memory, copied by
 * hf_plan_create.  Replaces the dense (n_mod, n_sample, batch, n_bin) constant tensors built by
 * pdf.py:566-611, modifiers/*.py `*_combined.__init__`, constraints.py:14-83,144-202.
 */
typedef struct {
    int32_t n_pars;    /* P  : len(config.par_order flattened), pdf.py:296 */
    int32_t n_bins;    /* NB : config.nmaindata */
    int32_t n_samples; /* S  : len(config.samples) */
    int32_t n_aux;     /* config.nauxdata; data vector = [NB main | n_aux aux] (pdf.py:818-825) */
    int32_t n_segs;    /* NSEG: channels; scalar factors are constant inside a segment */
    int32_t n_hs;      /* H  : histosys modifiers */
    int32_t n_hs_rows; /* R  : samples that carry any histosys */
    int32_t hs_code;   /* HF_HS_CODE0 | HF_HS_CODE4P */
    int32_t n_sf;      /* scalar-factor entries (one per modifier x sample x segment) */
    int32_t n_bf_slots;/* K  : max per-bin factors on one (sample, bin) cell */
    int32_t n_gauss;   /* Gaussian-constrained parameters (constraints.py:14-141) */
    int32_t n_pois;    /* Poisson-constrained parameters  (constraints.py:144-262) */
    int32_t has_clip_sample, has_clip_bin; /* pdf.py:696-699, :708-709 */
    double clip_sample, clip_bin;
    const double* nominal;        /* [S][NB]      pdf.py:598-600 */
    const int32_t* bin_seg;       /* [NB]         segment (channel) of each bin */
    const int32_t* hs_par;        /* [H]          flat parameter index of each histosys alpha */
    const int32_t* hs_row_sample; /* [R]          sample index of each histosys row */
    const double* hs_t1;          /* [H][R][NB]   zero where masked (histosys.py:172) */
    const double* hs_t2;          /* [H][R][NB] */
    const int32_t* sf_ptr;        /* [S*NSEG+1]   CSR over q = sample*NSEG + seg */
    const int32_t* sf_par;        /* [n_sf] */
    const int32_t* sf_kind;       /* [n_sf]       HF_SF_* */
    const double* sf_coef;        /* [n_sf][8]    ln(hi), ln(lo), a1..a6 (code4.py:127-129) */
    const int32_t* bf_par;        /* [K][S][NB]   parameter index or -1 (shapesys.py:116-171) */
    const int32_t* g_par;         /* [n_gauss] */
    const int32_t* g_aux;         /* [n_gauss]    index into the aux block of the data vector */
    const double* g_sigma;        /* [n_gauss] */
    const int32_t* p_par;         /* [n_pois] */
    const int32_t* p_aux;         /* [n_pois] */
    const double* p_factor;       /* [n_pois]     rate = theta * factor (constraints.py:231-236) */
} hf_plan_desc;

typedef struct hf_plan hf_plan;

/* library */
int hf_version(void);
const char* hf_last_error(void);
/* measured peak of the fp64 FMA pipe (TFLOP/s) and of a device copy (GB/s): the roofline
 * denominators bench.py reports next to MEASURED_PEAKS.json */
int hf_fp64_peak(double* tflops_out, void* stream);

/* plan lifetime: tables are uploaded to the CURRENT device */
int hf_plan_create(const hf_plan_desc* desc, hf_plan** out);
void hf_plan_destroy(hf_plan* plan);
/* kernels launched by this plan since creation (for bench.py's gpu_launches claim) */
int64_t hf_plan_launch_count(const hf_plan* plan);
/* per-kernel timing of the dominant kernel (hf_main_kernel): when enabled, every launch is
 * bracketed by CUDA events on the launching stream; hf_plan_main_kernel_ms synchronises those
 * events, returns the summed milliseconds and the number of launches since the last reset */
int hf_plan_set_timing(hf_plan* plan, int enable);
int hf_plan_main_kernel_ms(hf_plan* plan, double* ms_out, int64_t* launches_out);

/*
 * K1 fused forward: out[b] = ln L(pars[b] | data[b or 0])
 * replaces Model.logpdf (pdf.py:957-996) = _MainModel._expected_data_unchecked (pdf.py:675-713)
 * + 7 x *_combined.apply (modifiers/*.py) + interpolators code0/1/4/4p + _ConstraintModel
 * (pdf.py:482-563) + probability.*.log_prob (probability.py:52-299).
 *   pars dev [B][P]; data dev [data_rows][NB+n_aux] with data_rows == 1 (shared) or == B.
 */
int hf_logpdf(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
              int64_t B, double* out, void* stream);

/* K2 fused forward + analytic gradient; grad dev [B][P].
 * No reference counterpart on numpy (finite differences, optimize/opt_scipy.py:96-105). */
int hf_logpdf_grad(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                   int64_t B, double* out, double* grad, void* stream);

/* expected data [B][NB+n_aux]: Model.expected_data (pdf.py:886-906) */
int hf_expected_data(hf_plan* plan, const double* pars, int64_t B, double* out, void* stream);

/* Host-buffer variants (the end-to-end call a reference-side binding would make):
 * H2D of pars/data, the kernels, D2H of the results, all on `stream`, then a stream sync. */
int hf_logpdf_grad_host(hf_plan* plan, const double* pars_host, const double* data_host,
                        int64_t data_rows, int64_t B, double* out_host, double* grad_host,
                        void* stream);

/*
 * K3 batched bounded minimiser: N independent fits of 2*NLL = -2 ln L
 * replaces scipy.optimize.minimize(method="SLSQP") driven by optimize/opt_scipy.py:46-105 /
 * optimize/mixins.py:124-209 under infer/mle.py:69-205.
 *   data  dev [data_rows][D], data_rows in {1, N}
 *   init  dev [init_rows][P], init_rows in {1, N}; FIXED parameters stay at their init value
 *   lo/hi dev [P]; fixed dev [P] (uint8)
 *   x dev [N][P]; fun dev [N] (= 2*NLL at x, all constant terms included); status dev [N]
 *   (0 converged, 1 iteration limit, 2 numerical failure); niter dev [N] (may be NULL)
 */
typedef struct {
    int32_t max_iter;   /* default 200 */
    int32_t reserved0;  /* (unused; keeps the layout of earlier builds) */
    double tol_edm;     /* stop when the estimated 2*NLL distance to the minimum < tol_edm (default 1e-11) */
    double tol_step;    /* and the preconditioned step max-norm < tol_step (default 1e-8) */
    int32_t verbose;
    int32_t reserved;
} hf_fit_opts;

void hf_fit_default_opts(hf_fit_opts* opts);
int hf_fit_batch(hf_plan* plan, const double* data, int64_t data_rows, const double* init,
                 int64_t init_rows, const double* lo, const double* hi, const uint8_t* fixed,
                 int64_t N, const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                 int32_t* niter, void* stream);
/* One lock-step batch over fits that differ in their fixed mask and share data rows: fit i uses
 * data row data_map[i] (any value in [0, data_rows); NULL = row i, or row 0 when data_rows == 1)
 * and the mask fixed_alt when use_alt[i] != 0, else fixed (both NULL = hf_fit_batch).  This is how
 * the constrained (POI fixed, infer/mle.py:138-205) and unconstrained fits of every toy of BOTH
 * hypotheses (infer/calculators.py:838-869) run as one batch: the slowest fit is waited for once. */
int hf_fit_batch_mixed(hf_plan* plan, const double* data, int64_t data_rows, const int32_t* data_map,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* fixed, const uint8_t* fixed_alt, const uint8_t* use_alt,
                       int64_t N, const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                       int32_t* niter, void* stream);
/* The same lock-step batch with a TABLE of masks: masks [n_masks][P] (1..256 rows), fit i uses row
 * mask_id[i] (ids must be < n_masks -- larger ones are clamped to the last row; NULL only when
 * n_masks == 1).  A parameter gets a finite-difference Hessian column when it is free under at
 * least one row.  This is what lets ONE batch serve many signal hypotheses that share a background
 * model (pyhf_b200/patchset.py: one normfactor per signal point, all but one fixed at zero in
 * every fit; reference workflow workspace.py:392-442 + patchset.py:300, one Model and five fits
 * per patch). */
int hf_fit_batch_masks(hf_plan* plan, const double* data, int64_t data_rows, const int32_t* data_map,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* masks, int32_t n_masks, const uint8_t* mask_id, int64_t N,
                       const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                       int32_t* niter, void* stream);
/* Launch shape of the most recent hf_fit_* call on this plan (bench.py reports it next to fits/s):
 * info[0] = 1 per-fit kernel (one CTA per fit, hf_fit_solo.cu) / 0 lock-step driver (hf_fit.cu),
 * [1] grid (persistent CTAs), [2] threads per CTA, [3] resident CTAs per SM, [4] dynamic shared memory
 * per CTA (bytes), [5] SMs of the device, [6] 1 when the per-bin block is eliminated in low-rank form,
 * [7] reserved. */
int hf_plan_fit_info(const hf_plan* plan, int32_t info[8]);
int hf_fit_batch_host(hf_plan* plan, const double* data_host, int64_t data_rows,
                      const double* init_host, int64_t init_rows, const double* lo_host,
                      const double* hi_host, const uint8_t* fixed_host, int64_t N,
                      const hf_fit_opts* opts, double* x_host, double* fun_host,
                      int32_t* status_host, int32_t* niter_host, void* stream);

/*
 * Hessian of 2*NLL at B points: central differences of the analytic gradient (one-sided where a
 * step would leave [lo, hi]), symmetrised; rows/columns of fixed parameters are the identity's.
 * The covariance the reference gets from iminuit (optimize/opt_minuit.py:136-152, errordef = 1 on
 * 2*NLL) is 2 * inverse(hess) on the free parameters; optimize/mixins.py:83-120 consumes it for
 * return_uncertainties / return_correlations.
 *   pars dev [B][P]; data dev [data_rows][D], data_rows in {1, B}; lo/hi dev [P]; fixed dev [P];
 *   hess dev [B][P][P]
 */
int hf_hessian(hf_plan* plan, const double* pars, const double* data, int64_t data_rows, int64_t B,
               const double* lo, const double* hi, const uint8_t* fixed, double* hess, void* stream);

/* The same matrix in closed form (one CTA per point; the curvature the minimiser K3 uses): mode 2 =
 * exact Hessian on the given data, mode 1 = Fisher matrix (Hessian on the Asimov data of each point,
 * positive semi-definite).  Returns HF_ERR_UNSUPPORTED for plans outside its scope (clipping, more
 * than 8 samples or 120 histosys modifiers, per-bin parameters acting on several bins): callers fall
 * back to hf_hessian.  fixed may be NULL. */
int hf_hessian_analytic(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                        int64_t B, const uint8_t* fixed, int32_t mode, double* hess, void* stream);

/*
 * K4 Philox-4x32-10 toy sampler: out[i] ~ Pois(lambda(pars)) (x) Normal(theta, sigma) (x)
 * Pois(theta*factor), i = 0..N-1, counter = (seed, offset + i) so results do not depend on how
 * toys are sharded over ranks.  Replaces Simultaneous.sample (probability.py:263-274) and
 * scipy.stats.poisson/norm(...).rvs (tensor/numpy_backend.py:28-31, :43-47).
 *   pars dev [P]; out dev [N][NB+n_aux]
 */
int hf_sample_toys(hf_plan* plan, const double* pars, uint64_t seed, uint64_t offset, int64_t N,
                   double* out, void* stream);
/* generic element samplers behind backend.poisson_dist(rate).sample / normal_dist(mu, sigma).sample
 *   rate/loc/scale dev [M]; out dev [N][M]; element (i, j) uses counter (seed, offset + i, j) */
int hf_sample_poisson(const double* rate, int64_t M, int64_t N, uint64_t seed, uint64_t offset,
                      double* out, void* stream);
int hf_sample_normal(const double* loc, const double* scale, int64_t M, int64_t N, uint64_t seed,
                     uint64_t offset, double* out, void* stream);

/*
 * K5 test-statistic epilogue: q[i] = max(0, fun_fixed[i] - fun_free[i]), zeroed where
 * muhat[i] > mu (mode 0: q_mu / q~_mu, infer/test_statistics.py:16-35) or muhat[i] < 0
 * (mode 1: q0, :507-520); mode 2: t_mu (no zeroing, :38-60).  All dev [N].
 */
int hf_teststat(const double* fun_fixed, const double* fun_free, const double* muhat, double mu,
                int32_t mode, int64_t N, double* q, void* stream);
/* p-value counts of EmpiricalDistribution.pvalue (infer/calculators.py:632-640):
 * counts[j] = #{ i : samples[i] >= values[j] }.  samples dev [N], values dev [M], counts dev [M] */
int hf_pvalue_counts(const double* samples, int64_t N, const double* values, int64_t M,
                     int64_t* counts, void* stream);
/* The same counts against samples SORTED ascending (binary search per value, O(M log N)): the
 * path for ToyCalculator.expected_pvalues (infer/calculators.py:955-973), which asks for the
 * p-values of every background-only toy -- O(N^2) in the reference. */
int hf_pvalue_counts_sorted(const double* sorted_samples, int64_t N, const double* values, int64_t M,
                            int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HF_B200_H */
