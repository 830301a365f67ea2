// K1 / K2: fused HistFactory log-likelihood (and analytic gradient) for a batch of parameter
// points.  sm_100a, fp64 vector pipe (no tensor cores: see DESIGN.md).
//
// Pipeline per batch (all on one stream):
//   hf_prep_kernel      per (point, parameter) transposition to batch-major + histosys
//                       coefficients c1,c2,dc1,dc2 + per-(sample,segment) scalar-factor products
//   hf_main_kernel      the hot kernel: one CTA per PT parameter points x all bins.
//                         forward : thread = bin, register tile acc[RC rows][PT points], histosys
//                                   tables streamed coalesced from L2 ([H][R][NB]), per-point
//                                   coefficients broadcast from shared memory; sample sum,
//                                   Poisson main term, warp-shuffle reduction over bins
//                         backward: thread = histosys modifier, w*F tile staged in shared memory,
//                                   transposed tables ([R][NB][H]) streamed coalesced
//   hf_finish_kernel    scalar-factor / constraint gradient pieces, constraint terms, constants,
//                       transposition back to row-major [B][P]
//
// Reference being replaced: pdf.py:645-713, modifiers/*.py apply(), interpolators/code{0,1,4,4p}.py,
// constraints.py:100-141,219-262, tensor/numpy_backend.py:435-436,482-490, probability.py:276-299.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "hf_internal.cuh"

// ------------------------------------------------------------------------------------------
// prep: grid-stride over points; lanes = consecutive points so batch-major stores coalesce.
// ------------------------------------------------------------------------------------------
__global__ void hf_transpose_pars_kernel(DevPlan pl, Work w, const double* __restrict__ pars,
                                         int64_t B, int64_t ld) {
    // tile transpose [B][P] -> [P][ld]; padded points replicate point B-1
    __shared__ double tile[32][33];
    const int64_t pt0 = (int64_t)blockIdx.x * 32;
    const int p0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int64_t pt = pt0 + j;
        if (pt >= B) pt = B - 1;
        const int p = p0 + tx;
        tile[j][tx] = (p < pl.P) ? pars[pt * pl.P + p] : 0.0;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j;
        const int64_t pt = pt0 + tx;
        if (p < pl.P && pt < ld) w.parsT[(int64_t)p * ld + pt] = tile[tx][j];
    }
}

// points per thread of hf_prep_kernel.  4 amortises one coefficient fetch over 4 evaluations, but
// the block then walks 512 points x 99 normsys parameter rows = 400 KB, past L1, and every
// parameter load becomes an L2 round trip (ncu: 14.5 long-scoreboard stalls per issue, fp64 pipe
// 29 %); with 1 the rows of a block (99 KB) stay in L1: 1.25 -> 0.87 ms on C4.
#ifndef HF_PREP_TB
#define HF_PREP_TB 128
#endif
constexpr int PREP_TB = HF_PREP_TB;    // threads per block

#ifndef HF_FIN_CAP
#define HF_FIN_CAP 12
#endif
constexpr int FIN_CAP = HF_FIN_CAP;  // entries of a parameter staged per pass (per warp)

// F_q of NQ consecutive (sample, segment) products with identical (parameter, kind) lists: the
// parameter value, its range predicates and the clamped log-domain argument are fetched /
// computed once per entry and serve all NQ polynomials.  Everything per lane is select-based
// (lanes are parameter points: a branch on alpha would diverge in nearly every warp); the
// outside-range factors exp(+-alpha ln hi/lo) are summed in the log domain, one exp() per product.
template <int NQ, int PPT>
__device__ __forceinline__ void hf_prep_group(const DevPlan& pl, const Work& w, int64_t ld, int q0,
                                              const int64_t* pt, const bool* ok, double* __restrict__ cf_s,
                                              int* __restrict__ kind_s, int* __restrict__ par_s, int cap) {
    double prod[NQ][PPT], logsum[NQ][PPT];
#pragma unroll
    for (int j = 0; j < NQ; ++j)
#pragma unroll
        for (int i = 0; i < PPT; ++i) { prod[j][i] = 1.0; logsum[j][i] = 0.0; }
    const int len = pl.sf_ptr[q0 + 1] - pl.sf_ptr[q0];
    // the entry tables go through shared memory: the coefficient stream (0.5 MB on C4) does not
    // survive in L1 next to the parameter rows.  A uniform 16-byte shared load still costs two
    // wavefronts, so the 64 bytes of an entry are the L1 cost of the kernel: every fetched
    // coefficient serves PPT points.
    for (int c0 = 0; c0 < len; c0 += cap) {
        const int ne = min(cap, len - c0);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const double* src = pl.sf_coef + 8 * ((int64_t)pl.sf_ptr[q0 + j] + c0);
            for (int i = threadIdx.x; i < ne * 8; i += blockDim.x) cf_s[(size_t)j * cap * 8 + i] = src[i];
        }
        for (int i = threadIdx.x; i < ne; i += blockDim.x) {
            kind_s[i] = pl.sf_kind[pl.sf_ptr[q0] + c0 + i];
            par_s[i] = pl.sf_par[pl.sf_ptr[q0] + c0 + i];
        }
        __syncthreads();
        for (int e = 0; e < ne; ++e) {
            const int kind = kind_s[e];  // uniform across the block
            const int64_t prow = (int64_t)par_s[e] * ld;
            double t[PPT];
#pragma unroll
            for (int i = 0; i < PPT; ++i) t[i] = w.parsT[prow + pt[i]];
            if (kind == HF_SF_THETA) {
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < PPT; ++i) prod[j][i] *= t[i];
            } else if (kind == HF_SF_CODE4) {
                bool inside[PPT], up[PPT];
                double tm[PPT];
#pragma unroll
                for (int i = 0; i < PPT; ++i) {
                    inside[i] = (t[i] > -1.0) && (t[i] < 1.0);
                    up[i] = t[i] >= 1.0;
                    tm[i] = hf_selp(inside[i], 0.0, t[i]);
                }
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    double cf[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) cf[c] = cf_s[((size_t)j * cap + e) * 8 + c];
                    const double nlnd = -cf[1];
#pragma unroll
                    for (int i = 0; i < PPT; ++i) {
                        double poly = cf[7];
#pragma unroll
                        for (int c = 6; c >= 2; --c) poly = fma(poly, t[i], cf[c]);
                        poly = fma(poly, t[i], 1.0);
                        prod[j][i] *= hf_selp(inside[i], poly, 1.0);
                        logsum[j][i] = fma(tm[i], hf_selp(up[i], cf[0], nlnd), logsum[j][i]);
                    }
                }
            } else {  // code1: exp(alpha ln hi) above zero, exp(-alpha ln lo) below
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const double lnu = cf_s[((size_t)j * cap + e) * 8], nlnd = -cf_s[((size_t)j * cap + e) * 8 + 1];
#pragma unroll
                    for (int i = 0; i < PPT; ++i)
                        logsum[j][i] = fma(t[i], hf_selp(t[i] > 0.0, lnu, nlnd), logsum[j][i]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j)
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            if (!ok[i]) continue;
            double v = prod[j][i];
            if (logsum[j][i] != 0.0) v *= exp(logsum[j][i]);
            w.FsegT[(int64_t)(q0 + j) * ld + pt[i]] = v;
        }
}

#ifndef HF_PREP_PPT
#define HF_PREP_PPT 2
#endif
#ifndef HF_PREP_CHUNK
#define HF_PREP_CHUNK 64
#endif
constexpr int PREP_PPT = HF_PREP_PPT;      // points per thread
constexpr int PREP_CHUNK = HF_PREP_CHUNK;  // entries of a product staged per pass (x up to 4 products)

__global__ void __launch_bounds__(PREP_TB)
hf_prep_kernel(DevPlan pl, Work w, int64_t ld, int do_grad, int skip_hs) {
    extern __shared__ __align__(16) double prep_smem[];
    const int cap = min(max(pl.max_sf_q, 1), PREP_CHUNK);
    double* cf_s = prep_smem;                                       // [4][cap][8]
    int* kind_s = reinterpret_cast<int*>(cf_s + 4 * 8 * (size_t)cap);   // [cap]
    int* par_s = kind_s + cap;                                      // [cap]
    int64_t pt[PREP_PPT];
    bool ok[PREP_PPT];
#pragma unroll
    for (int i = 0; i < PREP_PPT; ++i) {
        pt[i] = ((int64_t)blockIdx.x * PREP_PPT + i) * blockDim.x + threadIdx.x;
        ok[i] = pt[i] < ld;
        if (!ok[i]) pt[i] = ld - 1;
    }
    // blockIdx.y strides over the work items: first H histosys modifiers, then the groups of
    // (sample, segment) products
    // (skip_hs: the DMMA kernel computes the histosys coefficients itself, their batch-major scratch --
    //  4 x H x ld doubles, 210 MB on C4 at B = 65536 -- is neither written nor read)
    for (int item = blockIdx.y + (skip_hs ? pl.H : 0); item < pl.H + pl.n_qg; item += gridDim.y) {
        if (item < pl.H) {
            const int h = item;
#pragma unroll
            for (int i = 0; i < PREP_PPT; ++i) {
                if (!ok[i]) continue;
                const double a = w.parsT[(int64_t)pl.hs_par[h] * ld + pt[i]];
                double c1, c2, d1, d2;
                hf_hs_coeffs(pl.hs_code, a, c1, c2, d1, d2);
                w.c1T[(int64_t)h * ld + pt[i]] = c1;
                w.c2T[(int64_t)h * ld + pt[i]] = c2;
                if (do_grad) {
                    w.dc1T[(int64_t)h * ld + pt[i]] = d1;
                    w.dc2T[(int64_t)h * ld + pt[i]] = d2;
                }
            }
        } else {
            const int g = item - pl.H;
            const int q0 = pl.qg_first[g];
            switch (pl.qg_n[g]) {  // uniform across the block
                case 1: hf_prep_group<1, PREP_PPT>(pl, w, ld, q0, pt, ok, cf_s, kind_s, par_s, cap); break;
                case 2: hf_prep_group<2, PREP_PPT>(pl, w, ld, q0, pt, ok, cf_s, kind_s, par_s, cap); break;
                case 3: hf_prep_group<3, PREP_PPT>(pl, w, ld, q0, pt, ok, cf_s, kind_s, par_s, cap); break;
                default: hf_prep_group<4, PREP_PPT>(pl, w, ld, q0, pt, ok, cf_s, kind_s, par_s, cap); break;
            }
        }
    }
}

// theta-independent terms of one data row:  - sum_b lgamma(n_b+1) - sum_pois lgamma(a+1)
//                                           - sum_gauss ln(sigma sqrt(2 pi))
__global__ void hf_data_const_kernel(DevPlan pl, const double* __restrict__ data, int64_t rows,
                                     double* __restrict__ out) {
    const int64_t row = blockIdx.x;
    if (row >= rows) return;
    const double* d = data + row * pl.D;
    double acc = 0.0;
    for (int i = threadIdx.x; i < pl.NB; i += blockDim.x) acc -= lgamma(d[i] + 1.0);
    for (int i = threadIdx.x; i < pl.NPO; i += blockDim.x) acc -= lgamma(d[pl.NB + pl.p_aux[i]] + 1.0);
    for (int i = threadIdx.x; i < pl.NG; i += blockDim.x)
        acc -= log(pl.g_sigma[i] * 2.5066282746310002);  // sqrt(2 pi), numpy_backend.py:486-488
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[row] = v;
    }
}

// ------------------------------------------------------------------------------------------
// the hot kernel
// ------------------------------------------------------------------------------------------
// F = Fseg[s][seg] (scalar factors) and bf = product of per-bin factors of cell (s, bb)
template <int PT>
__device__ __forceinline__ void load_fseg(const DevPlan& pl, const Work& w, int64_t ld,
                                          int64_t pt0, int s, int seg, double (&F)[PT]) {
    const double2* fp =
        reinterpret_cast<const double2*>(w.FsegT + ((int64_t)s * pl.NSEG + seg) * ld + pt0);
#pragma unroll
    for (int i = 0; i < PT / 2; ++i) {
        const double2 v = fp[i];
        F[2 * i] = v.x;
        F[2 * i + 1] = v.y;
    }
}

template <int PT>
__device__ __forceinline__ void mul_binfactors(const DevPlan& pl, const Work& w, int64_t ld,
                                               int64_t pt0, int s, int bb, double (&F)[PT]) {
    for (int k = 0; k < pl.K; ++k) {
        const int p = pl.bf_par[((int64_t)k * pl.S + s) * pl.NB + bb];
        if (p >= 0) {
            const double2* gp = reinterpret_cast<const double2*>(w.parsT + (int64_t)p * ld + pt0);
#pragma unroll
            for (int i = 0; i < PT / 2; ++i) {
                const double2 v = gp[i];
                F[2 * i] *= v.x;
                F[2 * i + 1] *= v.y;
            }
        }
    }
}

template <int PT>
__device__ __forceinline__ void load_factor(const DevPlan& pl, const Work& w, int64_t ld,
                                            int64_t pt0, int s, int bb, int seg,
                                            double (&F)[PT]) {
    load_fseg<PT>(pl, w, ld, pt0, s, seg, F);
    mul_binfactors<PT>(pl, w, ld, pt0, s, bb, F);
}

template <int PT, int RC, int TB, bool GRAD, int MINB>
__global__ void __launch_bounds__(TB, MINB)
hf_main_kernel(DevPlan pl, Work w, int64_t B, int64_t ld, const double* __restrict__ data,
               int per_point, const int* __restrict__ row_idx) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int64_t row_off[PT];  // element offset of each point's data row
    constexpr int WPAD = PT + 2;  // 16-byte aligned rows, conflict-free vector stores
    double* c1s = smem;                      // [H][PT]
    double* c2s = c1s + (int64_t)pl.H * PT;  // [H][PT]
    double* wfs = c2s + (int64_t)pl.H * PT;  // GRAD: [R*TB][WPAD]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int64_t pt0 = (int64_t)blockIdx.x * PT;
    const int H = pl.H, R = pl.R, NB = pl.NB, S = pl.S;

    for (int i = tid; i < H * PT; i += TB) {
        const int h = i / PT, j = i - h * PT;
        c1s[i] = w.c1T[(int64_t)h * ld + pt0 + j];
        c2s[i] = w.c2T[(int64_t)h * ld + pt0 + j];
    }
    if (tid < PT) {
        int64_t pt = pt0 + tid;
        if (pt >= B) pt = B - 1;
        const int64_t row = row_idx ? (int64_t)row_idx[pt] : (per_point ? pt : 0);
        row_off[tid] = row * pl.D;
    }
    __syncthreads();

    const int ntiles = (NB + TB - 1) / TB;
    const int64_t hstride = (int64_t)R * NB;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int b = tile * TB + tid;
        const bool valid = b < NB;
        const int bb = valid ? b : NB - 1;
        const int seg = pl.bin_seg[bb];

        double lam[PT];
#pragma unroll
        for (int i = 0; i < PT; ++i) lam[i] = 0.0;

        // ---- (1a) samples that carry histosys: register-tiled contraction over modifiers
        for (int rc0 = 0; rc0 < R; rc0 += RC) {
            double acc[RC][PT];
#pragma unroll
            for (int j = 0; j < RC; ++j)
#pragma unroll
                for (int i = 0; i < PT; ++i) acc[j][i] = 0.0;
            const double* t1p[RC];
            const double* t2p[RC];
#pragma unroll
            for (int j = 0; j < RC; ++j) {
                const int row = (rc0 + j < R) ? rc0 + j : R - 1;
                t1p[j] = pl.t1 + (int64_t)row * NB + bb;
                t2p[j] = pl.t2 + (int64_t)row * NB + bb;
            }
#pragma unroll 2
            for (int h = 0; h < H; ++h) {
                double t1v[RC], t2v[RC];
#pragma unroll
                for (int j = 0; j < RC; ++j) {
                    t1v[j] = __ldg(t1p[j] + h * hstride);
                    t2v[j] = __ldg(t2p[j] + h * hstride);
                }
                const double2* a2 = reinterpret_cast<const double2*>(c1s + h * PT);
                const double2* g2 = reinterpret_cast<const double2*>(c2s + h * PT);
#pragma unroll
                for (int i = 0; i < PT / 2; ++i) {
                    const double2 a = a2[i];
                    const double2 g = g2[i];
#pragma unroll
                    for (int j = 0; j < RC; ++j) {
                        acc[j][2 * i] = fma(a.x, t1v[j], fma(g.x, t2v[j], acc[j][2 * i]));
                        acc[j][2 * i + 1] = fma(a.y, t1v[j], fma(g.y, t2v[j], acc[j][2 * i + 1]));
                    }
                }
            }
            // fold the chunk into lambda; stash base = nominal + delta for the backward sweep
#pragma unroll
            for (int j = 0; j < RC; ++j) {
                const int row = rc0 + j;
                if (row < R) {
                    const int s = pl.hs_row_sample[row];
                    const double nom = pl.nominal[(int64_t)s * NB + bb];
                    double F[PT];
                    load_factor<PT>(pl, w, ld, pt0, s, bb, seg, F);
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        const double base = nom + acc[j][i];
                        double r = F[i] * base;
                        if (pl.has_clip_sample) r = fmax(r, pl.clip_sample);
                        lam[i] += r;
                        if (GRAD) wfs[((int64_t)row * TB + tid) * WPAD + i] = base;
                    }
                }
            }
        }
        // ---- (1b) samples without histosys
        for (int s = 0; s < S; ++s) {
            if (pl.sample_hs_row[s] >= 0) continue;
            const double nom = pl.nominal[(int64_t)s * NB + bb];
            double F[PT];
            load_factor<PT>(pl, w, ld, pt0, s, bb, seg, F);
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                double r = F[i] * nom;
                if (pl.has_clip_sample) r = fmax(r, pl.clip_sample);
                lam[i] += r;
            }
        }

        // ---- (2) Poisson main term + d/dlambda
        double wgt[PT];
        {
            double term[PT];
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                const double n = data[row_off[i] + bb];
                double l = lam[i];
                bool live = true;
                if (pl.has_clip_bin && l < pl.clip_bin) {
                    l = pl.clip_bin;
                    live = false;
                }
                term[i] = valid ? hf_pois_term(n, l) : 0.0;
                // d/dl [xlogy(n, l) - l] is exactly -1 for n == 0 (also at l == 0, where n / l would be NaN)
                wgt[i] = (valid && live) ? ((n == 0.0) ? -1.0 : n / l - 1.0) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                double v = term[i];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) atomicAdd(&w.mainv[pt0 + i], v);
            }
        }

        if (GRAD) {
            // ---- (3) w*F per cell for the backward sweep; G_{s,seg}; per-bin factor gradients
            const int seg0 = __shfl_sync(0xffffffffu, seg, 0);
            const bool uniform = __all_sync(0xffffffffu, seg == seg0);
            for (int s = 0; s < S; ++s) {
                const int row = pl.sample_hs_row[s];
                const double nom = pl.nominal[(int64_t)s * NB + bb];
                // wr0 = w * (prod of per-bin factors) * base  : G_{s,seg} excludes the scalar
                // factors so that d/dtheta of a normfactor stays finite at theta == 0
                double wr0[PT], wfs_[PT];
#pragma unroll
                for (int i = 0; i < PT; ++i) wr0[i] = 1.0;
                mul_binfactors<PT>(pl, w, ld, pt0, s, bb, wr0);
                {
                    double F[PT];
                    load_fseg<PT>(pl, w, ld, pt0, s, seg, F);
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        const double base =
                            (row >= 0) ? wfs[((int64_t)row * TB + tid) * WPAD + i] : nom;
                        const double fb = F[i] * wr0[i];
                        const double r = fb * base;
                        const bool live = !(pl.has_clip_sample && r < pl.clip_sample);
                        const double wl = live ? wgt[i] : 0.0;
                        wfs_[i] = wl * F[i] * base;   // w * (scalar factors) * base
                        wr0[i] = wl * wr0[i] * base;
                        if (row >= 0) wfs[((int64_t)row * TB + tid) * WPAD + i] = wl * fb;
                    }
                }
                // per-bin factors: d/dgamma_k = w * Fseg * base * prod_{k' != k} gamma_k'
                // (no division: stays finite for a shapefactor sitting on its bound 0)
                for (int k = 0; k < pl.K; ++k) {
                    const int p = pl.bf_par[((int64_t)k * S + s) * NB + bb];
                    if (p >= 0 && valid) {
                        double oth[PT];
#pragma unroll
                        for (int i = 0; i < PT; ++i) oth[i] = wfs_[i];
                        for (int k2 = 0; k2 < pl.K; ++k2) {
                            if (k2 == k) continue;
                            const int p2 = pl.bf_par[((int64_t)k2 * S + s) * NB + bb];
                            if (p2 < 0) continue;
#pragma unroll
                            for (int i = 0; i < PT; ++i) oth[i] *= w.parsT[(int64_t)p2 * ld + pt0 + i];
                        }
#pragma unroll
                        for (int i = 0; i < PT; ++i)
                            atomicAdd(&w.gradT[(int64_t)p * ld + pt0 + i], oth[i]);
                    }
                }
                // G_{s,seg} += w * bf * base  (reduced over the bins of a segment)
                double* gp = w.GsegT + ((int64_t)s * pl.NSEG + seg) * ld + pt0;
                if (uniform) {
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        double v = wr0[i];
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                        if (lane == 0) atomicAdd(gp + i, v);
                    }
                } else if (valid) {
#pragma unroll
                    for (int i = 0; i < PT; ++i) atomicAdd(gp + i, wr0[i]);
                }
            }
            __syncthreads();

            // ---- (4) backward sweep: thread = histosys modifier, reduce over the tile's cells
            //          grad_h += dc1 * sum wF*T1 + dc2 * sum wF*T2
            const int nb_tile = min(TB, NB - tile * TB);
            for (int h = tid; h < pl.Hp; h += TB) {
                double u[PT], v[PT];
#pragma unroll
                for (int i = 0; i < PT; ++i) u[i] = v[i] = 0.0;
                for (int row = 0; row < R; ++row) {
                    const double* q1 = pl.t1t + ((int64_t)row * NB + tile * TB) * pl.Hp + h;
                    const double* q2 = pl.t2t + ((int64_t)row * NB + tile * TB) * pl.Hp + h;
                    const double* wrow = wfs + (int64_t)row * TB * WPAD;
#pragma unroll 4
                    for (int c = 0; c < nb_tile; ++c) {
                        const double a = __ldg(q1 + (int64_t)c * pl.Hp);
                        const double g = __ldg(q2 + (int64_t)c * pl.Hp);
                        const double2* wf2 = reinterpret_cast<const double2*>(wrow + c * WPAD);
#pragma unroll
                        for (int i = 0; i < PT / 2; ++i) {
                            const double2 x = wf2[i];
                            u[2 * i] = fma(x.x, a, u[2 * i]);
                            u[2 * i + 1] = fma(x.y, a, u[2 * i + 1]);
                            v[2 * i] = fma(x.x, g, v[2 * i]);
                            v[2 * i + 1] = fma(x.y, g, v[2 * i + 1]);
                        }
                    }
                }
                if (h < H) {
                    const int p = pl.hs_par[h];
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        const double d1 = w.dc1T[(int64_t)h * ld + pt0 + i];
                        const double d2 = w.dc2T[(int64_t)h * ld + pt0 + i];
                        atomicAdd(&w.gradT[(int64_t)p * ld + pt0 + i], d1 * u[i] + d2 * v[i]);
                    }
                }
            }
            __syncthreads();
        }
    }
}


// ------------------------------------------------------------------------------------------
// point-major form of the hot kernel for plans WITHOUT histosys (C3-class models)
// ------------------------------------------------------------------------------------------
// hf_main_kernel maps thread = bin with a register tile of points, which is what the histosys
// contraction wants; without histosys a (point, bin) costs ~S multiplies, one log and one divide,
// and that mapping leaves most of a warp's work in index arithmetic and shared-memory broadcasts
// (1.1e8 evals/s on C3, ~20 cycles per (point, bin) and SM).  Here thread = parameter point (lanes =
// consecutive points: every access to the batch-major scratch parsT / FsegT / GsegT / gradT is
// coalesced), a block sweeps one tile of PM_BT bins whose tables (nominal, per-bin parameter
// indices, segment ids, shared observed counts) sit in shared memory and are read as broadcasts,
// the scalar factors of the current segment live in registers, and the sums leave through
// fire-and-forget fp64 reductions (RED), one per (point, quantity) and tile.
constexpr int PM_BT = 32;    // bins per block
constexpr int PM_TB = 128;   // points per block
constexpr int PM_SMAX = 8;   // samples held in registers
constexpr int PM_KMAX = 4;   // per-bin factors on one cell

template <bool GRAD>
__global__ void __launch_bounds__(PM_TB)
hf_main_pm_kernel(DevPlan pl, Work w, int64_t B, int64_t ld, const double* __restrict__ data,
                  int per_point, const int* __restrict__ row_idx) {
    __shared__ double s_nom[PM_SMAX][PM_BT];
    __shared__ double s_data[PM_BT];
    __shared__ int s_bf[PM_KMAX][PM_SMAX][PM_BT];
    __shared__ int s_seg[PM_BT];
    const int tid = threadIdx.x;
    const int S = pl.S, K = pl.K, NB = pl.NB, NSEG = pl.NSEG;
    const int b0 = blockIdx.y * PM_BT;
    const int nb = min(PM_BT, NB - b0);
    for (int i = tid; i < S * PM_BT; i += PM_TB) {
        const int s_ = i / PM_BT, bl = i - s_ * PM_BT;
        const int bb = min(b0 + bl, NB - 1);
        s_nom[s_][bl] = pl.nominal[(int64_t)s_ * NB + bb];
        for (int k = 0; k < K; ++k) s_bf[k][s_][bl] = pl.bf_par[((int64_t)k * S + s_) * NB + bb];
    }
    if (tid < PM_BT) {
        const int bb = min(b0 + tid, NB - 1);
        s_seg[tid] = pl.bin_seg[bb];
        s_data[tid] = (!per_point && !row_idx) ? data[bb] : 0.0;
    }
    __syncthreads();
    const int64_t pt = (int64_t)blockIdx.x * PM_TB + tid;
    if (pt >= ld) return;
    const int64_t ptc = (pt < B) ? pt : B - 1;  // padded points repeat the last one
    const bool own_row = per_point || row_idx;
    const double* drow = data + (row_idx ? (int64_t)row_idx[ptc] : (per_point ? ptc : 0)) * pl.D;

    double F[PM_SMAX], G[PM_SMAX];
#pragma unroll
    for (int s_ = 0; s_ < PM_SMAX; ++s_) F[s_] = G[s_] = 0.0;
    int cur = -1;
    double macc = 0.0;
    for (int bl = 0; bl < nb; ++bl) {
        const int seg = s_seg[bl];
        if (seg != cur) {  // (uniform across the block)
            if (GRAD && cur >= 0) {
#pragma unroll
                for (int s_ = 0; s_ < PM_SMAX; ++s_)
                    if (s_ < S) atomicAdd(&w.GsegT[((int64_t)s_ * NSEG + cur) * ld + pt], G[s_]);
            }
#pragma unroll
            for (int s_ = 0; s_ < PM_SMAX; ++s_) {
                G[s_] = 0.0;
                if (s_ < S) F[s_] = w.FsegT[((int64_t)s_ * NSEG + seg) * ld + pt];
            }
            cur = seg;
        }
        double lam = 0.0, bfs[PM_SMAX];
#pragma unroll
        for (int s_ = 0; s_ < PM_SMAX; ++s_) {
            bfs[s_] = 1.0;
            if (s_ < S) {
                for (int k = 0; k < K; ++k) {
                    const int p = s_bf[k][s_][bl];
                    if (p >= 0) bfs[s_] *= w.parsT[(int64_t)p * ld + pt];
                }
                double r = F[s_] * bfs[s_] * s_nom[s_][bl];
                if (pl.has_clip_sample) r = fmax(r, pl.clip_sample);
                lam += r;
            }
        }
        const double n = own_row ? drow[b0 + bl] : s_data[bl];
        double l = lam;
        bool live = true;
        if (pl.has_clip_bin && l < pl.clip_bin) {
            l = pl.clip_bin;
            live = false;
        }
        macc += hf_pois_term(n, l);
        if (GRAD) {
            // d/dl [xlogy(n, l) - l] is exactly -1 for n == 0 (also at l == 0, where n / l would be NaN)
            const double wgt = live ? ((n == 0.0) ? -1.0 : n / l - 1.0) : 0.0;
#pragma unroll
            for (int s_ = 0; s_ < PM_SMAX; ++s_) {
                if (s_ < S) {
                    const double nom = s_nom[s_][bl];
                    const bool lv = !(pl.has_clip_sample && F[s_] * bfs[s_] * nom < pl.clip_sample);
                    const double wl = lv ? wgt : 0.0;
                    G[s_] = fma(wl * bfs[s_], nom, G[s_]);  // excludes the scalar factors (finite at theta == 0)
                    for (int k = 0; k < K; ++k) {
                        const int p = s_bf[k][s_][bl];
                        if (p < 0) continue;
                        double oth = wl * F[s_] * nom;  // no division: finite for a factor sitting on 0
                        for (int k2 = 0; k2 < K; ++k2) {
                            const int p2 = s_bf[k2][s_][bl];
                            if (k2 != k && p2 >= 0) oth *= w.parsT[(int64_t)p2 * ld + pt];
                        }
                        atomicAdd(&w.gradT[(int64_t)p * ld + pt], oth);
                    }
                }
            }
        }
    }
    if (GRAD && cur >= 0) {
#pragma unroll
        for (int s_ = 0; s_ < PM_SMAX; ++s_)
            if (s_ < S) atomicAdd(&w.GsegT[((int64_t)s_ * NSEG + cur) * ld + pt], G[s_]);
    }
    atomicAdd(&w.mainv[pt], macc);
}

static int hf_main_pm_launch(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
                             const int* row_idx, bool grad, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    const dim3 grid((unsigned)((ld + PM_TB - 1) / PM_TB), (unsigned)((pl.NB + PM_BT - 1) / PM_BT));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (plan->timing) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    if (grad) hf_main_pm_kernel<true><<<grid, PM_TB, 0, st>>>(pl, plan->w, B, ld, data, per_point, row_idx);
    else hf_main_pm_kernel<false><<<grid, PM_TB, 0, st>>>(pl, plan->w, B, ld, data, per_point, row_idx);
    if (plan->timing) {
        cudaEventRecord(e1, st);
        plan->ev.push_back(e0);
        plan->ev.push_back(e1);
    }
    HF_CUDA_OK(cudaGetLastError());
    plan->launches++;
    return HF_OK;
}

// ------------------------------------------------------------------------------------------
// finish: constraints, scalar-factor gradients, constants, transposition to row-major
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t hf_row_of(const int* row_idx, int per_point, int64_t pt) {
    return row_idx ? (int64_t)row_idx[pt] : (per_point ? pt : 0);
}

// 32 points x 8 term lanes per block: lane ty adds the constraint terms i = ty, ty + 8, ... of its
// point, so that a small batch (a handful of straggling fits) does not walk all NG + NPO terms
// serially (C3: 119 dependent loads + 100 logs = 83 us per launch with one thread per point)
__global__ void __launch_bounds__(256)
hf_finish_value_kernel(DevPlan pl, Work w, int64_t B, int64_t ld, const double* __restrict__ data,
                       int per_point, const int* __restrict__ row_idx, double* __restrict__ out) {
    __shared__ double part[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t pt = (int64_t)blockIdx.x * 32 + tx;
    double acc = 0.0;
    int64_t row = 0;
    if (pt < B) {
        row = hf_row_of(row_idx, per_point, pt);
        const double* aux = data + row * pl.D + pl.NB;
#pragma unroll 4
        for (int i = ty; i < pl.NG; i += 8) {
            const double th = w.parsT[(int64_t)pl.g_par[i] * ld + pt];
            const double z = (aux[pl.g_aux[i]] - th) / (1.4142135623730951 * pl.g_sigma[i]);
            acc -= z * z;  // numpy_backend.py:489
        }
#pragma unroll 4
        for (int i = ty; i < pl.NPO; i += 8) {
            const double th = w.parsT[(int64_t)pl.p_par[i] * ld + pt];
            acc += hf_pois_term(aux[pl.p_aux[i]], th * pl.p_factor[i]);
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && pt < B) {
        double v = w.mainv[pt] + w.dconst[row];
#pragma unroll
        for (int k = 0; k < 8; ++k) v += part[k][tx];
        out[pt] = v;
    }
}

#ifndef HF_FIN_PPT
#define HF_FIN_PPT 2
#endif
constexpr int FIN_PPT = HF_FIN_PPT;  // points per thread in hf_finish_grad_kernel
#ifndef HF_FIN_UNROLL
#define HF_FIN_UNROLL 2
#endif
constexpr int FIN_UNROLL = HF_FIN_UNROLL;  // entries of a parameter evaluated side by side

#ifndef HF_FIN_MINB
#define HF_FIN_MINB 3
#endif
__global__ void __launch_bounds__(256, HF_FIN_MINB)
hf_finish_grad_kernel(DevPlan pl, Work w, int64_t B, int64_t ld,
                                      const double* __restrict__ data, int per_point,
                                      const int* __restrict__ row_idx, double* __restrict__ grad,
                                      int x_in_smem) {
    // 128 points x 32 parameters per block; every thread carries 4 points so that one fetch of an
    // entry's coefficients serves 4 evaluations; transposed through shared memory so that both
    // the batch-major reads and the row-major writes coalesce
    extern __shared__ __align__(16) double fin_smem[];
    // per-warp staging of one parameter's entry table: [8 warps][max_sf_par] x (8 coef, kind, q)
    const int cap = min(max(pl.max_sf_par, 1), FIN_CAP);
    double* cf_w = fin_smem + (size_t)threadIdx.y * cap * 8;
    int* kind_w = reinterpret_cast<int*>(fin_smem + (size_t)8 * cap * 8) + (size_t)threadIdx.y * cap * 2;
    int* q_w = kind_w + cap;
    const int64_t pt0 = (int64_t)blockIdx.x * 32 * FIN_PPT;
    const int p0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    // X_q = F_q * G_q of the block's points, staged once: every parameter of the block walks the
    // same S*NSEG rows, which would otherwise be re-fetched from L2 for each of its entries
    constexpr int NPB = 32 * FIN_PPT;
    double* xs = fin_smem + (size_t)8 * cap * 9;  // behind the per-warp entry tables
    if (x_in_smem) {
        const int nq = pl.S * pl.NSEG;
        for (int q = ty; q < nq; q += 8) {
#pragma unroll
            for (int i = 0; i < FIN_PPT; ++i) {
                int64_t pt = pt0 + tx + 32 * i;
                if (pt >= B) pt = B - 1;
                xs[(size_t)q * NPB + tx + 32 * i] = w.FsegT[(int64_t)q * ld + pt] * w.GsegT[(int64_t)q * ld + pt];
            }
        }
        __syncthreads();
    }
    // the warp's four parameters keep their sums in registers until every warp is done with xs:
    // the transposition tile then reuses that shared memory (more resident blocks per SM)
    double gout[4][FIN_PPT];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = ty + 8 * jj;
        const int p = p0 + j;
        double g[FIN_PPT], th[FIN_PPT];
        int64_t pt[FIN_PPT];
        bool ok[FIN_PPT];
#pragma unroll
        for (int i = 0; i < FIN_PPT; ++i) {
            pt[i] = pt0 + tx + 32 * i;
            ok[i] = p < pl.P && pt[i] < B;
            if (pt[i] >= B) pt[i] = B - 1;
            g[i] = ok[i] ? w.gradT[(int64_t)p * ld + pt[i]] : 0.0;
            th[i] = (p < pl.P) ? w.parsT[(int64_t)p * ld + pt[i]] : 1.0;
        }
        if (p < pl.P) {
            // scalar factors (normsys / normfactor / lumi): sum_e (dF_q/dtheta) * G_q with
            // dF_q/dtheta = dlog f_e * F_q, or the product of the *other* factors of q for a
            // plain theta factor (finite at theta == 0: the POI sits on its bound in q~ fits)
            const int k0 = pl.sfp_ptr[p], e1 = pl.sfp_ptr[p + 1];
            if (k0 < e1 && pl.sfp_kind[k0] != HF_SF_THETA) {
                // the chunk after the current one travels through registers while the current one
                // is evaluated: its L2 latency (7 chunks per parameter on C4) is off the warp's
                // critical path
                constexpr int NPF = (FIN_CAP * 8 + 31) / 32, NPI = (FIN_CAP + 31) / 32;
                double pf[NPF];
                int pk[NPI], pq[NPI];
                auto fetch = [&](int kc) {
                    const int ne = min(cap, e1 - kc);
#pragma unroll
                    for (int r = 0; r < NPF; ++r) {
                        const int i = tx + 32 * r;
                        pf[r] = (i < ne * 8) ? pl.sfp_coef[8 * (int64_t)kc + i] : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < NPI; ++r) {
                        const int i = tx + 32 * r;
                        pk[r] = (i < ne) ? pl.sfp_kind[kc + i] : 0;
                        pq[r] = (i < ne) ? pl.sfp_q[kc + i] : 0;
                    }
                };
                fetch(k0);
                for (int kc = k0; kc < e1; kc += cap) {
                const int ne = min(cap, e1 - kc);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < NPF; ++r) {
                    const int i = tx + 32 * r;
                    if (i < ne * 8) cf_w[i] = pf[r];
                }
#pragma unroll
                for (int r = 0; r < NPI; ++r) {
                    const int i = tx + 32 * r;
                    if (i < ne) { kind_w[i] = pk[r]; q_w[i] = pq[r]; }
                }
                __syncwarp();
                if (kc + cap < e1) fetch(kc + cap);
#pragma unroll FIN_UNROLL
                for (int k = 0; k < ne; ++k) {
                    const int kind = kind_w[k];
                    const int q = q_w[k];
                    double cf[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) cf[c] = cf_w[8 * k + c];
                    // one (warp-uniform) branch on the kind around ALL points of the thread, so
                    // that their dependent Horner / reciprocal chains interleave
                    double dl[FIN_PPT];
                    if (kind == HF_SF_CODE4) {
#pragma unroll
                        for (int i = 0; i < FIN_PPT; ++i) dl[i] = hf_sf_dlog_code4(cf, th[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < FIN_PPT; ++i) dl[i] = (th[i] >= 0.0) ? cf[0] : -cf[1];  // right-sided at the kink
                    }
                    if (x_in_smem) {
#pragma unroll
                        for (int i = 0; i < FIN_PPT; ++i)
                            g[i] = fma(dl[i], xs[(size_t)q * NPB + tx + 32 * i], g[i]);
                    } else {
                        const int64_t qrow = (int64_t)q * ld;
#pragma unroll
                        for (int i = 0; i < FIN_PPT; ++i)
                            g[i] = fma(dl[i], w.FsegT[qrow + pt[i]] * w.GsegT[qrow + pt[i]], g[i]);
                    }
                }
                }
            } else {
                for (int k = k0; k < e1; ++k) {
                    const int e = pl.sfp_entry[k];
                    const int q = pl.sfp_q[k];
#pragma unroll
                    for (int i = 0; i < FIN_PPT; ++i) {
                        double coef;
                        if (pl.sf_kind[e] == HF_SF_THETA) {
                            double logsum = 0.0;
                            coef = 1.0;
                            for (int e2 = pl.sf_ptr[q]; e2 < pl.sf_ptr[q + 1]; ++e2) {
                                if (e2 == e) continue;
                                hf_sf_accum(pl.sf_kind[e2], pl.sf_coef + 8 * (int64_t)e2,
                                            w.parsT[(int64_t)pl.sf_par[e2] * ld + pt[i]], coef, logsum);
                            }
                            if (logsum != 0.0) coef *= exp(logsum);
                        } else {
                            coef = hf_sf_dlog(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, th[i]) *
                                   w.FsegT[(int64_t)q * ld + pt[i]];
                        }
                        g[i] = fma(coef, w.GsegT[(int64_t)q * ld + pt[i]], g[i]);
                    }
                }
            }
            const int ck = pl.par_cons_kind[p];
            if (ck != 0) {
                const int ci = pl.par_cons_idx[p];
#pragma unroll
                for (int i = 0; i < FIN_PPT; ++i) {
                    const int64_t row = hf_row_of(row_idx, per_point, pt[i]);
                    if (ck == 1) {
                        const double sg = pl.g_sigma[ci];
                        g[i] += (data[row * pl.D + pl.NB + pl.g_aux[ci]] - th[i]) / (sg * sg);
                    } else {
                        {
                            const double a = data[row * pl.D + pl.NB + pl.p_aux[ci]];
                            g[i] += ((a == 0.0) ? 0.0 : a / th[i]) - pl.p_factor[ci];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < FIN_PPT; ++i) gout[jj][i] = g[i];
    }
    __syncthreads();
    double (*tile)[NPB + 1] = reinterpret_cast<double (*)[NPB + 1]>(xs);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < FIN_PPT; ++i) tile[ty + 8 * jj][tx + 32 * i] = gout[jj][i];
    __syncthreads();
    for (int r = ty; r < 32 * FIN_PPT; r += 8) {
        const int64_t pt = pt0 + r;
        const int p = p0 + tx;
        if (p < pl.P && pt < B) grad[pt * pl.P + p] = tile[tx][r];
    }
}

// expected data [B][D]: lambda via the forward kernel pieces is overkill for an off-path call;
// one thread per (point, bin).  Model.expected_data, pdf.py:886-906.
__global__ void hf_expected_kernel(DevPlan pl, Work w, int64_t B, int64_t ld,
                                   double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * pl.D) return;
    const int64_t pt = idx / pl.D;
    const int j = (int)(idx - pt * pl.D);
    if (j >= pl.NB) {
        const int a = j - pl.NB;
        double v = 0.0;
        for (int i = 0; i < pl.NG; ++i)
            if (pl.g_aux[i] == a) v = w.parsT[(int64_t)pl.g_par[i] * ld + pt];
        for (int i = 0; i < pl.NPO; ++i)
            if (pl.p_aux[i] == a) v = w.parsT[(int64_t)pl.p_par[i] * ld + pt] * pl.p_factor[i];
        out[idx] = v;
        return;
    }
    const int b = j;
    const int seg = pl.bin_seg[b];
    double lam = 0.0;
    for (int s = 0; s < pl.S; ++s) {
        double F = w.FsegT[((int64_t)s * pl.NSEG + seg) * ld + pt];
        for (int k = 0; k < pl.K; ++k) {
            const int p = pl.bf_par[((int64_t)k * pl.S + s) * pl.NB + b];
            if (p >= 0) F *= w.parsT[(int64_t)p * ld + pt];
        }
        double base = pl.nominal[(int64_t)s * pl.NB + b];
        const int row = pl.sample_hs_row[s];
        if (row >= 0) {
            for (int h = 0; h < pl.H; ++h) {
                const int64_t o = ((int64_t)h * pl.R + row) * pl.NB + b;
                base = fma(w.c1T[(int64_t)h * ld + pt], pl.t1[o],
                           fma(w.c2T[(int64_t)h * ld + pt], pl.t2[o], base));
            }
        }
        double r = F * base;
        if (pl.has_clip_sample) r = fmax(r, pl.clip_sample);
        lam += r;
    }
    if (pl.has_clip_bin) lam = fmax(lam, pl.clip_bin);
    out[idx] = lam;
}

static void hf_prep_launch(const DevPlan& pl, const Work& w, int64_t ld, int do_grad, cudaStream_t st, int skip_hs = 0) {
    const int items = (skip_hs ? 0 : pl.H) + pl.n_qg;
    const int gy = items < 1 ? 1 : (items > 64 ? 64 : items);
    const dim3 grid((unsigned)((ld + PREP_TB * PREP_PPT - 1) / (PREP_TB * PREP_PPT)), (unsigned)gy);
    const size_t smem = (size_t)std::min(std::max(pl.max_sf_q, 1), PREP_CHUNK) * (4 * 64 + 8) + 16;
    hf_prep_kernel<<<grid, PREP_TB, smem, st>>>(pl, w, ld, do_grad, skip_hs);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <int PT, int RC, int TB, int MINB>
static int launch_main(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
                       const int* row_idx, bool grad, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    size_t smem = sizeof(double) * (size_t)(2 * pl.H * PT);
    if (grad) smem += sizeof(double) * (size_t)pl.R * TB * (PT + 2);
    const dim3 grid((unsigned)(ld / PT));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (plan->timing) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    if (grad) {
        auto k = hf_main_kernel<PT, RC, TB, true, MINB>;
        HF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TB, smem, st>>>(pl, plan->w, B, ld, data, per_point, row_idx);
    } else {
        auto k = hf_main_kernel<PT, RC, TB, false, MINB>;
        HF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TB, smem, st>>>(pl, plan->w, B, ld, data, per_point, row_idx);
    }
    if (plan->timing) {
        cudaEventRecord(e1, st);
        plan->ev.push_back(e0);
        plan->ev.push_back(e1);
    }
    HF_CUDA_OK(cudaGetLastError());
    plan->launches++;
    return HF_OK;
}

// smem budget decides the point tile: PT=16 needs (2H*16 + R*128*18)*8 bytes for the gradient.
static int pick_pt(const hf_plan* plan, bool grad, int64_t B) {
    const DevPlan& pl = plan->d;
    if (const char* e = getenv("HF_PT")) {  // tuning knob (tools/, not the product default)
        const int v = atoi(e);
        if (v == 2 || v == 4 || v == 8 || v == 16) return v;
    }
    const size_t lim = 200 * 1024;
    auto need = [&](int pt) {
        size_t s = sizeof(double) * (size_t)(2 * pl.H * pt);
        if (grad) s += sizeof(double) * (size_t)pl.R * 128 * (pt + 2);
        return s;
    };
    if (B >= 16 * 148 && need(16) <= lim / 2) return 16;
    if (B >= 8 * 64 && need(8) <= lim) return 8;
    if (need(4) <= lim) return 4;
    if (need(2) <= lim) return 2;
    return 0;
}

int hf_eval_launch(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                   int64_t B, double* out, double* grad, cudaStream_t st, const int* row_idx) {
    if (B <= 0) return HF_OK;
    const DevPlan& pl = plan->d;
    if (!row_idx && data_rows != 1 && data_rows != B) {
        hf_set_error("data_rows must be 1 or B (got %lld, B=%lld)", (long long)data_rows, (long long)B);
        return HF_ERR_INVALID;
    }
    const bool do_grad = grad != nullptr;
    const int PT = pick_pt(plan, do_grad, B);
    if (PT == 0) {
        hf_set_error("model too large for the shared-memory tile (H=%d, R=%d)", pl.H, pl.R);
        return HF_ERR_UNSUPPORTED;
    }
    const int64_t ld = round_up(B, 32);
    int rc = hf_ensure_work(plan, ld, data_rows);
    if (rc) return rc;
    Work& w = plan->w;
    const int per_point = (data_rows == 1) ? 0 : 1;

    {
        dim3 grid((unsigned)(ld / 32), (unsigned)((pl.P + 31) / 32));
        hf_transpose_pars_kernel<<<grid, dim3(32, 8), 0, st>>>(pl, w, pars, B, ld);
        plan->launches++;
    }
    // tensor-core (DMMA) formulation for large batches of plans that qualify
    bool use_mma = pl.mma_ok && B >= 256;
    if (const char* e = getenv("HF_MMA")) use_mma = pl.mma_ok && atoi(e) != 0;
    {
        hf_prep_launch(pl, w, ld, do_grad ? 1 : 0, st, use_mma ? 1 : 0);
        plan->launches++;
    }
    // accumulators: GsegT | gradT | mainv are carved back to back (hf_ensure_work): one memset
    if (do_grad) {
        HF_CUDA_OK(cudaMemsetAsync(w.GsegT, 0, sizeof(double) * ld * ((size_t)pl.S * pl.NSEG + pl.P + 1), st));
    } else {
        HF_CUDA_OK(cudaMemsetAsync(w.mainv, 0, sizeof(double) * ld, st));
    }
    if (plan->dconst_held && plan->dconst_held_data == data && plan->dconst_held_rows == data_rows) {
        w.dconst = const_cast<double*>(plan->dconst_held);
    } else {
        hf_data_const_kernel<<<(unsigned)data_rows, 128, 0, st>>>(pl, data, data_rows, plan->dconst);
        plan->launches++;
        w.dconst = plan->dconst;
    }

    if (use_mma) {
        rc = hf_main_mma_launch(plan, B, ld, data, per_point, row_idx, do_grad, st);
        if (rc != HF_ERR_UNSUPPORTED) {
            if (rc) return rc;
            goto finish;
        }
        // (did not take the batch after all: the bin-major kernel below needs the coefficient scratch)
        hf_prep_launch(pl, w, ld, do_grad ? 1 : 0, st, 0);
        plan->launches++;
    }
    {
        // histosys-free plans in batches: point-major kernel (single fits keep the bin-major one below)
        bool use_pm = pl.H == 0 && pl.S <= PM_SMAX && pl.K <= PM_KMAX && B >= 32;
        if (const char* e = getenv("HF_PM")) use_pm = use_pm && atoi(e) != 0;
        if (use_pm) {
            rc = hf_main_pm_launch(plan, B, ld, data, per_point, row_idx, do_grad, st);
            if (rc) return rc;
            goto finish;
        }
    }
    {
    // point tile: RC = rows per register tile
    int minb = 3;
    if (const char* e = getenv("HF_MINB")) minb = atoi(e);
    if (PT == 16) {
        if (pl.R >= 3 || pl.R == 0) rc = launch_main<16, 4, 128, 2>(plan, B, ld, data, per_point, row_idx, do_grad, st);
        else if (pl.R == 2) rc = launch_main<16, 2, 128, 2>(plan, B, ld, data, per_point, row_idx, do_grad, st);
        else rc = launch_main<16, 1, 128, 2>(plan, B, ld, data, per_point, row_idx, do_grad, st);
    } else if (PT == 8) {
        if (pl.R >= 3 || pl.R == 0) {
            if (minb >= 4) rc = launch_main<8, 4, 128, 4>(plan, B, ld, data, per_point, row_idx, do_grad, st);
            else rc = launch_main<8, 4, 128, 3>(plan, B, ld, data, per_point, row_idx, do_grad, st);
        } else rc = launch_main<8, 2, 128, 3>(plan, B, ld, data, per_point, row_idx, do_grad, st);
    } else if (PT == 4) {
        rc = launch_main<4, 4, 128, 4>(plan, B, ld, data, per_point, row_idx, do_grad, st);
    } else {
        rc = launch_main<2, 4, 128, 4>(plan, B, ld, data, per_point, row_idx, do_grad, st);
    }
    if (rc) return rc;
    }
finish:
    hf_finish_value_kernel<<<(unsigned)((B + 31) / 32), dim3(32, 8), 0, st>>>(pl, w, B, ld, data, per_point,
                                                                              row_idx, out);
    plan->launches++;
    if (do_grad) {
        dim3 grid((unsigned)((B + 32 * FIN_PPT - 1) / (32 * FIN_PPT)), (unsigned)((pl.P + 31) / 32));
        size_t fin_smem = (size_t)std::min(std::max(pl.max_sf_par, 1), FIN_CAP) * (8 * 72) + 16;
        const size_t xs_bytes = sizeof(double) * (size_t)pl.S * pl.NSEG * 32 * FIN_PPT;
        const size_t tile_bytes = sizeof(double) * 32 * (32 * FIN_PPT + 1);  // reuses the xs region
        const int x_in_smem = (fin_smem + xs_bytes <= 150 * 1024) ? 1 : 0;
        fin_smem += std::max(x_in_smem ? xs_bytes : (size_t)0, tile_bytes);
        HF_CUDA_OK(cudaFuncSetAttribute(hf_finish_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
        hf_finish_grad_kernel<<<grid, dim3(32, 8), fin_smem, st>>>(pl, w, B, ld, data, per_point, row_idx, grad, x_in_smem);
        plan->launches++;
    }
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_data_const_launch(hf_plan* plan, const double* data, int64_t rows, double* out, cudaStream_t st) {
    if (rows <= 0) return HF_OK;
    hf_data_const_kernel<<<(unsigned)rows, 128, 0, st>>>(plan->d, data, rows, out);
    plan->launches++;
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}

int hf_expected_launch(hf_plan* plan, const double* pars, int64_t B, double* out, cudaStream_t st) {
    if (B <= 0) return HF_OK;
    const DevPlan& pl = plan->d;
    const int64_t ld = round_up(B, 32);
    int rc = hf_ensure_work(plan, ld, 1);
    if (rc) return rc;
    Work& w = plan->w;
    dim3 grid((unsigned)(ld / 32), (unsigned)((pl.P + 31) / 32));
    hf_transpose_pars_kernel<<<grid, dim3(32, 8), 0, st>>>(pl, w, pars, B, ld);
    hf_prep_launch(pl, w, ld, 0, st);
    const int64_t n = B * pl.D;
    hf_expected_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pl, w, B, ld, out);
    plan->launches += 3;
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}
