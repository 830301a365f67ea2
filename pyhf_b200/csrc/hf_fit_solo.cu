// K3, per-fit form: ONE CTA runs one bounded maximum-likelihood fit from start to finish on the device.
//
// Replaces scipy.optimize.minimize(method="SLSQP") as driven by optimize/opt_scipy.py:46-105 under
// infer/mle.py:69-205, for plans inside the scope of the closed-form Hessian (hf_hess_analytic.cu)
// whose per-bin parameters never share a bin.  The lock-step driver of hf_fit.cu (one batched K2
// launch per round for all fits, host-side round control) stays for every other plan.
//
// Why a second form: the lock-step loop pays ~30 stream operations and 2-3 host synchronisations per
// round, ~100 rounds per batch, and every fit waits for the slowest one -- a 22 ms floor per toy
// hypotest that capped the 1 -> 8 GPU strong scaling.  Fits are independent, so here a CTA takes a fit
// from a device-side queue and iterates alone; nothing returns to the host until the batch is done.
//
// One iteration of a fit (damped projected Newton, the algorithm of oracle/fit_proto.py):
//   sweep   one pass over the bins at the current point gives 2NLL, its gradient AND its Hessian in
//           closed form (exact, or the Fisher matrix = Hessian on the Asimov data of the point, far
//           from the minimum), stored structured: S_gg (global x global, packed, shared memory),
//           C (per-bin parameter x global) and the diagonal of the per-bin block -- the full P x P
//           matrix is never formed.  Histosys x histosys runs on the fp64 tensor-core path (DMMA).
//   solve   the per-bin block is diagonal, so it is eliminated analytically and only the Schur
//           complement on the global parameters is factorised (Cholesky in shared memory);
//           Levenberg-Marquardt shift tau, active-set projection on the box bounds.
//   trial   value-only sweep at clip(x + d); Armijo test; a rejection raises tau and re-solves.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "hf_internal.cuh"

namespace {

constexpr int TB = 32;             // bins per tile
constexpr int NT = 256;            // threads per CTA
constexpr int SMX = HF_AH_SMAX;    // samples
constexpr double TINY = 7.888609052210118e-31;  // 2^-100
constexpr double kArmijo = 1e-4;
constexpr int kMaxKicks = 2;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i
__device__ __forceinline__ int trisym(int i, int j) { return i >= j ? tri(i, j) : tri(j, i); }
__device__ __forceinline__ void tri_decode(int e, int& a, int& b) {  // e = a (a + 1) / 2 + b, b <= a
    a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > e) --a;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    b = e - a * (a + 1) / 2;
}

// dlog f / dtheta and d2log f / dtheta2 of one scalar-factor entry (interpolators/code4.py:178-239,
// code1.py:81-104; normfactor / lumi: f = theta)
__device__ __forceinline__ void sf_dlogs(int kind, const double* __restrict__ cf, double t, double& dl,
                                         double& d2) {
    if (kind == HF_SF_THETA) {
        dl = 1.0 / t;
        d2 = -(dl * dl);  // dl * dl + d2 == 0 exactly: f is linear in theta
        return;
    }
    if (kind == HF_SF_CODE4 && t > -1.0 && t < 1.0) {
        double poly = cf[7], dp = 0.0, ddp = 0.0;
#pragma unroll
        for (int i = 6; i >= 2; --i) {
            ddp = fma(ddp, t, 2.0 * dp);
            dp = fma(dp, t, poly);
            poly = fma(poly, t, cf[i]);
        }
        ddp = fma(ddp, t, 2.0 * dp);
        dp = fma(dp, t, poly);
        poly = fma(poly, t, 1.0);
        dl = dp / poly;
        d2 = ddp / poly - dl * dl;
        return;
    }
    const bool up = (kind == HF_SF_CODE4) ? (t >= 1.0) : (t >= 0.0);
    dl = up ? cf[0] : -cf[1];
    d2 = 0.0;
}

// offsets (in doubles) of the shared-memory carve; built on the host (solo_layout)
struct SoloLayout {
    int xs, xc, xt, g, d, hc, Fq, base, rS, FB, Wt, wgt, w2t, sqW, lwj, lw2, gpart, Y, ywords;
    int S, r, z, linv, dinv, gl, hl, red;
    int ints;       // start of the int region (in doubles)
    int total;      // doubles
    // "small" placement: these live in shared memory too (else in the per-CTA global scratch)
    int small;
    int Sorig, Sred, C, SegM, SegG, SegN, LOCw;   // offsets in doubles (shared memory or scratch)
    int scratch_doubles;                         // per-CTA global scratch (big placement)
    int ltile;                                   // rows of C staged through the Y region (big placement)
};

struct SoloArgs {
    int64_t N;
    const double* data; int64_t data_rows; const int* data_map; const double* dconst;
    const double* init; int64_t init_rows;
    const double *lo, *hi;
    const uint8_t *fixed, *fixed_alt, *use_alt; int n_alt;
    double tol_edm, tol_step, exact_below, damp0;
    int max_iter, trace;
    int64_t trace_fit;  // HF_FIT_TRACE_FIT: index of one more fit to trace (-1: none)
    int ng, nl, npk, Hpad, nj, Hc, npart;
    const int *gpos, *lpos, *glist, *llist;   // [P], [P], [ng], [nl]
    double* scratch; int64_t scratch_stride;
    int* next;                                // device-side queue head
    double *x_out, *fun_out; int32_t *status_out, *niter_out;
    SoloLayout L;
};

// pointers of one CTA
struct Ctx {
    double *xs, *xc, *xt, *g, *d, *hc, *Fq, *base, *rS, *FB, *Wt, *wgt, *w2t, *sqW, *lwj, *lw2, *gpart, *Y;
    double *S, *r, *z, *linv, *dinv, *gl, *hl, *red;
    double *Sorig, *Sred, *C, *SegM, *SegG, *SegN, *LOCw;
    int *segt, *lgp, *actg, *actl;
    unsigned* lmk;
    int* flag;  // [4] small flags
    long long* prof;  // [8] clock64 sums per phase (thread 0; printed with the trace)
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) s += red[k];
    return s;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) s = fmax(s, red[k]);
    return s;
}

// ---------------------------------------------------------------------------------------------
// One pass over the bins at the point xin [P] (shared memory).  FULL: value, gradient (c.g) and the
// structured Hessian (c.S -> copied to c.Sorig, c.C, c.hl); else the value only.  Returns 2NLL.
// ---------------------------------------------------------------------------------------------
template <int MAXT, bool FULL>
__device__ double solo_sweep(const DevPlan& pl, const SoloArgs& a, const Ctx& c, const double* xin, bool fisher,
                             const double* __restrict__ drow, double dconst) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int P = pl.P, NB = pl.NB, S = pl.S, NSEG = pl.NSEG, H = pl.H, R = pl.R, K = pl.K;
    const int Hpad = a.Hpad, nj = a.nj, Hc = a.Hc, npart = a.npart;
    const int ng = a.ng, nl = a.nl, npk = a.npk;
    long long t_mark = clock64();
    auto lap = [&](int slot) {
        if (tid == 0) { const long long now = clock64(); c.prof[slot] += now - t_mark; t_mark = now; }
    };

    for (int p = tid; p < P; p += NT) {
        double v = xin[p];
        if (pl.ah_par_mult[p] && fabs(v) < TINY) v = (v < 0.0) ? -TINY : TINY;
        c.xs[p] = v;
        if (FULL) c.g[p] = 0.0;
    }
    if (FULL) {
        for (int i = tid; i < npk; i += NT) c.S[i] = 0.0;
        for (int i = tid; i < nl * ng; i += NT) c.C[i] = 0.0;
        for (int i = tid; i < nl; i += NT) c.hl[i] = 0.0;
        const int nseg_words = NSEG * S * S + NSEG * S + NSEG * S * npart * Hc + pl.ah_nbl * S;
        for (int i = tid; i < nseg_words; i += NT) c.SegM[i] = 0.0;  // SegM | SegG | SegN | LOCw are contiguous
        if (MAXT > 0)
            for (int i = tid; i < TB * Hpad; i += NT) c.Y[i] = 0.0;
    }
    __syncthreads();
    if (MAXT > 0) {
        for (int h = tid; h < H; h += NT) {
            const double al = c.xs[pl.hs_par[h]];
            double c1, c2, d1, d2;
            hf_hs_coeffs(pl.hs_code, al, c1, c2, d1, d2);
            double dd = 0.0;
            if (pl.hs_code == HF_HS_CODE4P && !(al > 1.0) && !(al < -1.0)) {
                const double a2 = al * al;
                dd = a2 * (90.0 * a2 - 120.0) + 30.0;
            }
            c.hc[5 * h] = c1; c.hc[5 * h + 1] = c2; c.hc[5 * h + 2] = d1; c.hc[5 * h + 3] = d2; c.hc[5 * h + 4] = dd;
        }
    }
    // product of the scalar factors of every (sample, segment): one warp per product, lanes = entries
    for (int q = warp; q < S * NSEG; q += NT / 32) {
        double prod = 1.0, logsum = 0.0;
        for (int e = pl.sf_ptr[q] + lane; e < pl.sf_ptr[q + 1]; e += 32)
            hf_sf_accum(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, c.xs[pl.sf_par[e]], prod, logsum);
        for (int o = 16; o > 0; o >>= 1) {
            prod *= __shfl_xor_sync(0xffffffffu, prod, o);
            logsum += __shfl_xor_sync(0xffffffffu, logsum, o);
        }
        if (lane == 0) c.Fq[q] = (logsum != 0.0) ? prod * exp(logsum) : prod;
    }
    __syncthreads();

    lap(FULL ? 0 : 5);
    // ---------------------------------------------------------------- sweep over the bins
    const int my_h = (MAXT > 0) ? (tid % Hc) : 0, my_part = (MAXT > 0) ? (tid / Hc) : 0;
    const bool h_thread = MAXT > 0 && my_h < H && my_part < npart;
    double Nacc[SMX];
#pragma unroll
    for (int s = 0; s < SMX; ++s) Nacc[s] = 0.0;
    double ddacc = 0.0, ghacc = 0.0;
    int ncur = -1;
    double macc = 0.0;
    int mcur = -1;
    const bool pair_thread = tid < S * S, g_thread = tid >= SMX * SMX && tid < SMX * SMX + S;
    const int ps = tid / S, ps2 = tid - (tid / S) * S, gs = tid - SMX * SMX;
    double acc[MAXT > 0 ? MAXT : 1][2];
#pragma unroll
    for (int k = 0; k < (MAXT > 0 ? MAXT : 1); ++k) acc[k][0] = acc[k][1] = 0.0;
    const int nt8 = (H + 7) / 8, ntile = nt8 * (nt8 + 1) / 2;
    double dc1 = 0.0, dc2 = 0.0, ddc2 = 0.0;
    if (h_thread) { dc1 = c.hc[5 * my_h + 2]; dc2 = c.hc[5 * my_h + 3]; ddc2 = c.hc[5 * my_h + 4]; }
    double vacc = 0.0;

    for (int b0 = 0; b0 < NB; b0 += TB) {
        // ---- (1) histosys delta of every (row, bin) of the tile
        if (MAXT > 0) {
            if (warp < R) {
                const int b = min(b0 + lane, NB - 1);
                const double* p1 = pl.t1 + (int64_t)warp * NB + b;
                const double* p2 = pl.t2 + (int64_t)warp * NB + b;
                const int64_t hstride = (int64_t)R * NB;
                double s0 = 0.0, s1 = 0.0;
                int h = 0;
                for (; h + 1 < H; h += 2) {
                    s0 = fma(c.hc[5 * h], __ldg(p1 + h * hstride), fma(c.hc[5 * h + 1], __ldg(p2 + h * hstride), s0));
                    s1 = fma(c.hc[5 * h + 5], __ldg(p1 + (h + 1) * hstride), fma(c.hc[5 * h + 6], __ldg(p2 + (h + 1) * hstride), s1));
                }
                if (h < H) s0 = fma(c.hc[5 * h], __ldg(p1 + h * hstride), fma(c.hc[5 * h + 1], __ldg(p2 + h * hstride), s0));
                c.base[warp * TB + lane] = s0 + s1;
            }
            __syncthreads();
        }
        // ---- (2a) rates, value, weights, per-bin parameter
        if (tid < TB) {
            const int bl = tid, b = b0 + bl;
            const bool valid = b < NB;
            const int bb = valid ? b : NB - 1;
            const int seg = pl.bin_seg[bb];
            double lam = 0.0, rr[SMX];
#pragma unroll
            for (int s = 0; s < SMX; ++s) {
                rr[s] = 0.0;
                if (s < S) {
                    const int hr = pl.sample_hs_row[s];
                    double bs = pl.nominal[(int64_t)s * NB + bb];
                    if (MAXT > 0 && hr >= 0) bs += c.base[hr * TB + bl];
                    double bf = 1.0;
                    for (int k = 0; k < K; ++k) {
                        const int p = pl.bf_par[((int64_t)k * S + s) * NB + bb];
                        if (p >= 0) bf *= c.xs[p];
                    }
                    const double fb = c.Fq[s * NSEG + seg] * bf;
                    rr[s] = valid ? fb * bs : 0.0;
                    if (FULL) {
                        c.FB[s * TB + bl] = valid ? fb : 0.0;
                        c.rS[s * TB + bl] = rr[s];
                    }
                    lam += rr[s];
                }
            }
            const double n = drow[bb];
            if (valid) {
                // main term lambda - xlogy(n, lambda) (tensor/numpy_backend.py:435-436); a rate that is
                // positive only through the 2^-100 stand-in of a parameter sitting on 0 counts as 0
                const bool zero = lam >= 0.0 && lam < 1e-28;
                if (zero) vacc += (n == 0.0) ? 0.0 : INFINITY;
                else vacc += (n == 0.0) ? lam : lam - n * log(lam);
            }
            if (FULL) {
                double W = 0.0, wg = 0.0;
                if (valid) {
                    if (lam > 0.0) {
                        const double il = 1.0 / lam;
                        W = fisher ? 2.0 * il : 2.0 * n * il * il;
                        wg = (n == 0.0) ? 2.0 : 2.0 * (1.0 - n * il);
                    } else {
                        wg = 2.0;
                    }
                }
                const double w2 = fisher ? 0.0 : wg;
                c.Wt[bl] = W; c.wgt[bl] = wg; c.w2t[bl] = w2; c.sqW[bl] = sqrt(fmax(W, 0.0));
                c.segt[bl] = seg;
                int gp = -1;
                if (valid) {
                    const int i0 = pl.ah_bl_ptr[bb];
                    if (pl.ah_bl_ptr[bb + 1] > i0) {
                        gp = pl.ah_bl_par[i0];
                        const unsigned mk = pl.ah_bl_mask[i0];
                        const double ig = 1.0 / c.xs[gp];
                        double sum = 0.0;
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s < S && ((mk >> s) & 1u)) sum += rr[s];
                        const double Jg = sum * ig;
                        c.lmk[bl] = mk;
                        c.lwj[bl] = W * Jg;
                        c.lw2[bl] = w2 * ig;
                        const int li = a.lpos[gp];
                        c.hl[li] = W * Jg * Jg;
                        c.g[gp] = wg * Jg;
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s < S) c.LOCw[(int64_t)i0 * S + s] = rr[s] * (W * Jg + (((mk >> s) & 1u) ? w2 * ig : 0.0));
                    }
                }
                c.lgp[bl] = gp;
            }
        }
        if (!FULL) {
            if (MAXT > 0) __syncthreads();  // base is rewritten by the next tile
            continue;
        }
        __syncthreads();
        // ---- (2b) per-segment S x S sums  M_c = sum_b W r r^T ,  G_c = sum_b wg r
        if (pair_thread || g_thread) {
            for (int bl = 0; bl < TB && b0 + bl < NB; ++bl) {
                const int seg = c.segt[bl];
                if (seg != mcur) {
                    if (mcur >= 0) {
                        if (pair_thread) c.SegM[(int64_t)mcur * S * S + tid] += macc;
                        else c.SegG[(int64_t)mcur * S + gs] += macc;
                    }
                    macc = 0.0;
                    mcur = seg;
                }
                if (pair_thread) macc = fma(c.Wt[bl] * c.rS[ps * TB + bl], c.rS[ps2 * TB + bl], macc);
                else macc = fma(c.wgt[bl], c.rS[gs * TB + bl], macc);
            }
        }
        // ---- (3) histosys columns of the tile: a_b[h], the S x H sums N_c, per-bin parameter rows
        if (MAXT > 0) {
            if (h_thread) {
                for (int bl = my_part; bl < TB; bl += npart) {
                    const int b = b0 + bl;
                    if (b >= NB) { c.Y[bl * Hpad + my_h] = 0.0; continue; }
                    const int seg = c.segt[bl];
                    if (seg != ncur) {
                        if (ncur >= 0) {
#pragma unroll
                            for (int s = 0; s < SMX; ++s)
                                if (s < S && Nacc[s] != 0.0) c.SegN[(((int64_t)ncur * S + s) * npart + my_part) * Hc + my_h] += Nacc[s];
                        }
#pragma unroll
                        for (int s = 0; s < SMX; ++s) Nacc[s] = 0.0;
                        ncur = seg;
                    }
                    double A[SMX], asum = 0.0, ddl = 0.0;
#pragma unroll
                    for (int r = 0; r < SMX; ++r) {
                        A[r] = 0.0;
                        if (r < R) {
                            const int64_t o = ((int64_t)r * NB + b) * pl.Hp + my_h;
                            const double v1 = __ldg(pl.t1t + o), v2 = __ldg(pl.t2t + o);
                            const double fb = c.FB[pl.hs_row_sample[r] * TB + bl];
                            A[r] = fb * fma(dc1, v1, dc2 * v2);
                            asum += A[r];
                            ddl = fma(fb * ddc2, v2, ddl);
                        }
                    }
                    c.Y[bl * Hpad + my_h] = c.sqW[bl] * asum;
                    const double W = c.Wt[bl], w2 = c.w2t[bl];
                    ghacc = fma(c.wgt[bl], asum, ghacc);
#pragma unroll
                    for (int s = 0; s < SMX; ++s)
                        if (s < S) Nacc[s] = fma(W * c.rS[s * TB + bl], asum, Nacc[s]);
#pragma unroll
                    for (int r = 0; r < SMX; ++r)
                        if (r < R) {
                            const int sr = pl.hs_row_sample[r];
#pragma unroll
                            for (int s = 0; s < SMX; ++s)
                                if (s == sr) Nacc[s] = fma(w2, A[r], Nacc[s]);
                        }
                    ddacc = fma(w2, ddl, ddacc);
                    const int gp = c.lgp[bl];
                    if (gp >= 0) {
                        const unsigned mk = c.lmk[bl];
                        double sumA = 0.0;
#pragma unroll
                        for (int r = 0; r < SMX; ++r)
                            if (r < R && ((mk >> pl.hs_row_sample[r]) & 1u)) sumA += A[r];
                        c.C[(int64_t)a.lpos[gp] * ng + a.gpos[pl.hs_par[my_h]]] = c.lwj[bl] * asum + c.lw2[bl] * sumA;
                    }
                }
            }
            __syncthreads();
            // ---- (4) histosys x histosys block: acc += Y^T Y on the fp64 tensor-core path
#pragma unroll
            for (int k = 0; k < (MAXT > 0 ? MAXT : 1); ++k) {
                const int tile = warp + (NT / 32) * k;
                if (tile < ntile) {
                    int ti, tj;
                    tri_decode(tile, tj, ti);  // ti <= tj : row tile, column tile
                    const double* pa = c.Y + t4 * Hpad + 8 * ti + g8;
                    const double* pb = c.Y + t4 * Hpad + 8 * tj + g8;
#pragma unroll
                    for (int k0 = 0; k0 < TB; k0 += 4) dmma884(acc[k][0], acc[k][1], pa[k0 * Hpad], pb[k0 * Hpad]);
                }
            }
        }
        __syncthreads();
    }

    lap(FULL ? 1 : 6);
    // ---- value: main term + constraint terms (constraints.py:100-141, :219-262)
    double cacc = 0.0;
    {
        const double* aux = drow + NB;
        for (int i = tid; i < pl.NG; i += NT) {
            const double z = (aux[pl.g_aux[i]] - xin[pl.g_par[i]]) / pl.g_sigma[i];
            cacc = fma(0.5 * z, z, cacc);
        }
        for (int i = tid; i < pl.NPO; i += NT) cacc -= hf_pois_term(aux[pl.p_aux[i]], xin[pl.p_par[i]] * pl.p_factor[i]);
    }
    const double total = block_sum(vacc + cacc, c.red);
    const double fval = 2.0 * (total - dconst);
    if (!FULL) { lap(6); return fval; }

    // ---- flush what the threads still carry
    if (mcur >= 0) {
        if (pair_thread) c.SegM[(int64_t)mcur * S * S + tid] += macc;
        else if (g_thread) c.SegG[(int64_t)mcur * S + gs] += macc;
    }
    if (MAXT > 0) {
        if (h_thread && ncur >= 0) {
#pragma unroll
            for (int s = 0; s < SMX; ++s)
                if (s < S && Nacc[s] != 0.0) c.SegN[(((int64_t)ncur * S + s) * npart + my_part) * Hc + my_h] += Nacc[s];
        }
        c.gpart[tid] = h_thread ? ghacc : 0.0;
        c.gpart[NT + tid] = h_thread ? ddacc : 0.0;
        __syncthreads();
        if (tid < H) {
            double gsum = 0.0, dsum = 0.0;
            for (int k = 0; k < npart; ++k) { gsum += c.gpart[k * Hc + tid]; dsum += c.gpart[NT + k * Hc + tid]; }
            const int p = pl.hs_par[tid];
            c.g[p] += gsum;
            const int gi = a.gpos[p];
            c.S[tri(gi, gi)] += dsum;
        }
        __syncthreads();
        // histosys block out of the registers (each unordered pair once)
#pragma unroll
        for (int k = 0; k < (MAXT > 0 ? MAXT : 1); ++k) {
            const int tile = warp + (NT / 32) * k;
            if (tile < ntile) {
                int ti, tj;
                tri_decode(tile, tj, ti);
                const int hi = 8 * ti + g8;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int hj = 8 * tj + 2 * t4 + cc;
                    if (hi < H && hj < H && hi <= hj) c.S[trisym(a.gpos[pl.hs_par[hi]], a.gpos[pl.hs_par[hj]])] += acc[k][cc];
                }
            }
        }
    }
    __syncthreads();
    // ---------------------------------------------------------------- scalar-factor blocks, channel by channel
    double* DL = c.Y;                  // [S][nj]
    double* D2 = DL + SMX * nj;        // [S][nj]
    double* U0 = D2 + SMX * nj;        // [S][nj]   M_c DL
    double* Msm = U0 + SMX * nj;       // [S][S]
    double* G2sm = Msm + SMX * SMX;    // [S]
    const double hw = fisher ? 0.0 : 1.0;  // second-order terms enter the exact Hessian only
    lap(2);
    if (nj > 0) {
        const int npair = nj * (nj + 1) / 2;
        for (int ch = 0; ch < NSEG; ++ch) {
            for (int i = tid; i < 2 * SMX * nj; i += NT) DL[i] = 0.0;  // DL and D2
            if (tid < S * S) Msm[tid] = c.SegM[(int64_t)ch * S * S + tid];
            if (tid < S) G2sm[tid] = c.SegG[(int64_t)ch * S + tid];
            __syncthreads();
            for (int s = 0; s < S; ++s) {
                const int q = s * NSEG + ch;
                for (int e = pl.sf_ptr[q] + tid; e < pl.sf_ptr[q + 1]; e += NT) {
                    double dl, d2;
                    sf_dlogs(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, c.xs[pl.sf_par[e]], dl, d2);
                    const int j = pl.ah_sf_j[e];
                    DL[s * nj + j] = dl;
                    D2[s * nj + j] = d2;
                }
            }
            __syncthreads();
            for (int i = tid; i < S * nj; i += NT) {
                const int s = i / nj, j = i - s * nj;
                double u = 0.0;
                for (int s2 = 0; s2 < S; ++s2) u = fma(Msm[s * S + s2], DL[s2 * nj + j], u);
                U0[i] = u;
            }
            // gradient of the scalar-factor parameters
            for (int j = tid; j < nj; j += NT) {
                double gj = 0.0;
                for (int s = 0; s < S; ++s) gj = fma(DL[s * nj + j], G2sm[s], gj);
                if (gj != 0.0) c.g[pl.ah_sfp_list[j]] += gj;
            }
            __syncthreads();
            // scalar x scalar (each unordered pair once)
            for (int e = tid; e < npair; e += NT) {
                int j2, j1;
                tri_decode(e, j2, j1);  // j1 <= j2
                double v = 0.0;
                if (j1 != j2) {
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j1], fma(hw * G2sm[s], DL[s * nj + j2], U0[s * nj + j2]), v);
                } else {
                    for (int s = 0; s < S; ++s) {
                        const double dl = DL[s * nj + j1];
                        v = fma(dl, U0[s * nj + j1], v);
                        v = fma(hw * G2sm[s], dl * dl + D2[s * nj + j1], v);
                    }
                }
                if (v != 0.0) c.S[trisym(a.gpos[pl.ah_sfp_list[j1]], a.gpos[pl.ah_sfp_list[j2]])] += v;
            }
            __syncthreads();
            // histosys x scalar (a parameter may be both: shared-memory atomics keep coinciding targets exact)
            if (MAXT > 0) {
                for (int e = tid; e < H * nj; e += NT) {
                    const int j = e / H, h = e - j * H;
                    double v = 0.0;
                    for (int s = 0; s < S; ++s) {
                        double nsum = 0.0;
                        for (int k = 0; k < npart; ++k) nsum += c.SegN[(((int64_t)ch * S + s) * npart + k) * Hc + h];
                        v = fma(DL[s * nj + j], nsum, v);
                    }
                    if (v != 0.0) {
                        const int p1 = pl.hs_par[h], p2 = pl.ah_sfp_list[j];
                        atomicAdd(&c.S[trisym(a.gpos[p1], a.gpos[p2])], (p1 == p2) ? 2.0 * v : v);
                    }
                }
            }
            // per-bin parameters of this channel x scalar
            {
                const int i0 = pl.ah_bl_ptr[pl.ah_seg_first[ch]], i1 = pl.ah_bl_ptr[pl.ah_seg_first[ch + 1]];
                for (int e = tid; e < (i1 - i0) * nj; e += NT) {
                    const int il = e / nj, j = e - il * nj;
                    double v = 0.0;
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j], c.LOCw[(int64_t)(i0 + il) * S + s], v);
                    if (v != 0.0) c.C[(int64_t)a.lpos[pl.ah_bl_par[i0 + il]] * ng + a.gpos[pl.ah_sfp_list[j]]] += v;
                }
            }
            __syncthreads();
        }
    }
    // ---- constraint terms: gradient and curvature
    {
        const double* aux = drow + NB;
        for (int p = tid; p < P; p += NT) {
            const int ck = pl.par_cons_kind[p];
            if (ck == 0) continue;
            double gadd, hadd;
            if (ck == 1) {
                const int ci = pl.par_cons_idx[p];
                const double sg = pl.g_sigma[ci];
                gadd = 2.0 * (xin[p] - aux[pl.g_aux[ci]]) / (sg * sg);
                hadd = 2.0 / (sg * sg);
            } else {
                const int ci = pl.par_cons_idx[p];
                const double th = c.xs[p], av = aux[pl.p_aux[ci]];
                gadd = 2.0 * (pl.p_factor[ci] - av / th);
                hadd = fisher ? 2.0 * pl.p_factor[ci] / th : 2.0 * av / (th * th);
            }
            c.g[p] += gadd;
            const int li = a.lpos[p];
            if (li >= 0) c.hl[li] += hadd;
            else { const int gi = a.gpos[p]; c.S[tri(gi, gi)] += hadd; }
        }
    }
    __syncthreads();
    for (int i = tid; i < npk; i += NT) c.Sorig[i] = c.S[i];
    __syncthreads();
    lap(3);
    return fval;
}

// ---------------------------------------------------------------------------------------------
// Damped Newton direction from the structured Hessian: d in c.d, returns false when no shift made the
// reduced matrix positive definite.
// ---------------------------------------------------------------------------------------------
__device__ bool solo_solve(const DevPlan& pl, const SoloArgs& a, const Ctx& c, const uint8_t* __restrict__ mask,
                           double damp, int& attempts_out, double& edm_out, double& step_out) {
    const int tid = threadIdx.x;
    const int P = pl.P, ng = a.ng, nl = a.nl, npk = a.npk;
    const int TW = 16, TH = NT / TW;
    const int tx = tid % TW, ty = tid / TW;
    for (int i = tid; i < ng; i += NT) {
        const int p = a.glist[i];
        const double xv = c.xc[p], gv = c.g[p];
        c.actg[i] = (mask[p] || (xv <= a.lo[p] && gv > 0.0) || (xv >= a.hi[p] && gv < 0.0)) ? 1 : 0;
    }
    for (int i = tid; i < nl; i += NT) {
        const int p = a.llist[i];
        const double xv = c.xc[p], gv = c.g[p];
        c.actl[i] = (mask[p] || (xv <= a.lo[p] && gv > 0.0) || (xv >= a.hi[p] && gv < 0.0)) ? 1 : 0;
    }
    if (tid == 0) c.flag[0] = 0;
    __syncthreads();

    double tau = damp, tau_base = damp, maxdiag = -1.0;
    bool ok = false, have_copy = false;
    int attempts = 0;
    for (int attempt = 0; attempt < 40 && !ok; ++attempt) {
        ++attempts;
        bool bad = false;
        if (have_copy) {
            for (int e = tid; e < npk; e += NT) c.S[e] = c.Sred[e];
            __syncthreads();
            for (int i = tid; i < ng; i += NT)
                if (!c.actg[i]) c.S[tri(i, i)] += tau - tau_base;
            __syncthreads();
        } else {
            tau_base = tau;
            for (int i = tid; i < nl; i += NT) {
                double di = 0.0, gv = 0.0;
                if (!c.actl[i]) {
                    const double h = c.hl[i] + tau;
                    if (!(h > 0.0) || !isfinite(h)) c.flag[0] = 1;
                    di = 1.0 / h;
                    gv = c.g[a.llist[i]] * di;
                }
                c.dinv[i] = di;
                c.gl[i] = gv;
            }
            for (int e = tid; e < npk; e += NT) c.S[e] = c.Sorig[e];
            for (int i = tid; i < ng; i += NT) c.r[i] = -c.g[a.glist[i]];
            __syncthreads();
            for (int i = tid; i < ng; i += NT) c.S[tri(i, i)] += tau;
            bad = c.flag[0] != 0;
            if (!bad && nl > 0) {
                if (a.L.small) {
                    for (int e = tid; e < npk; e += NT) {
                        int i, j;
                        tri_decode(e, i, j);
                        double accv = 0.0;
                        for (int l = 0; l < nl; ++l) accv = fma(c.C[l * ng + i] * c.dinv[l], c.C[l * ng + j], accv);
                        c.S[e] -= accv;
                    }
                    for (int i = tid; i < ng; i += NT) {
                        double accv = 0.0;
                        for (int l = 0; l < nl; ++l) accv = fma(c.C[l * ng + i], c.gl[l], accv);
                        c.r[i] += accv;
                    }
                } else {
                    double* W = c.Y;  // staged rows of C
                    const int ltile = a.L.ltile;
                    for (int l0 = 0; l0 < nl; l0 += ltile) {
                        const int nt = min(ltile, nl - l0);
                        __syncthreads();
                        for (int e = tid; e < nt * ng; e += NT) W[e] = c.C[(int64_t)l0 * ng + e];
                        __syncthreads();
                        for (int e = tid; e < npk; e += NT) {
                            int i, j;
                            tri_decode(e, i, j);
                            double accv = 0.0;
                            for (int t = 0; t < nt; ++t) accv = fma(W[t * ng + i] * c.dinv[l0 + t], W[t * ng + j], accv);
                            c.S[e] -= accv;
                        }
                        for (int i = tid; i < ng; i += NT) {
                            double accv = 0.0;
                            for (int t = 0; t < nt; ++t) accv = fma(W[t * ng + i], c.gl[l0 + t], accv);
                            c.r[i] += accv;
                        }
                    }
                }
            }
            __syncthreads();
            // parameters of the active set (fixed, or on a bound with the gradient pushing outward)
            // keep their value: identity rows with a zero right-hand side
            for (int e = tid; e < npk; e += NT) {
                int i, j;
                tri_decode(e, i, j);
                if (c.actg[i] || c.actg[j]) c.S[e] = (i == j) ? 1.0 : 0.0;
            }
            for (int i = tid; i < ng; i += NT)
                if (c.actg[i]) c.r[i] = 0.0;
            __syncthreads();
            if (!bad) {
                for (int e = tid; e < npk; e += NT) c.Sred[e] = c.S[e];
                for (int i = tid; i < ng; i += NT) c.z[i] = c.r[i];  // (the right-hand side survives retries in z ... restored below)
                have_copy = true;
            }
            __syncthreads();
        }
        // right-looking Cholesky of S; linv holds 1 / L_kk
        for (int j = 0; j < ng && !bad; ++j) {
            const double piv = c.S[tri(j, j)];
            if (!(piv > 0.0) || !isfinite(piv)) { bad = true; break; }
            const double inv = 1.0 / sqrt(piv);
            for (int i = j + 1 + tid; i < ng; i += NT) c.S[tri(i, j)] *= inv;
            if (tid == 0) c.linv[j] = inv;
            __syncthreads();
            for (int i = j + 1 + ty; i < ng; i += TH) {
                const double lij = c.S[tri(i, j)];
                for (int k = j + 1 + tx; k <= i; k += TW) c.S[tri(i, k)] = fma(-lij, c.S[tri(k, j)], c.S[tri(i, k)]);
            }
            __syncthreads();
        }
        ok = !bad;
        if (!ok) {
            if (maxdiag < 0.0) {
                double m = 0.0;
                for (int i = tid; i < nl; i += NT)
                    if (!c.actl[i]) m = fmax(m, fabs(c.hl[i]));
                for (int i = tid; i < ng; i += NT)
                    if (!c.actg[i]) m = fmax(m, fabs(c.Sorig[tri(i, i)]));
                maxdiag = block_max(m, c.red);
            }
            __syncthreads();
            if (tid == 0) c.flag[0] = 0;
            tau = (tau < 1e-8 * maxdiag) ? 1e-8 * fmax(maxdiag, 1e-300) : tau * 10.0;
            __syncthreads();
        }
    }
    attempts_out = attempts;
    if (!ok) return false;
    // right-hand side of the reduced system (kept in z by the reduction pass)
    for (int i = tid; i < ng; i += NT) c.r[i] = c.z[i];
    __syncthreads();
    // L y = r
    for (int k = 0; k < ng; ++k) {
        const double zk = c.r[k] * c.linv[k];
        for (int i = k + 1 + tid; i < ng; i += NT) c.r[i] = fma(-c.S[tri(i, k)], zk, c.r[i]);
        __syncthreads();
        if (tid == 0) c.z[k] = zk;
    }
    __syncthreads();
    // L^T d_g = y   (d_g is left in r)
    for (int k = ng - 1; k >= 0; --k) {
        const double dk = c.z[k] * c.linv[k];
        for (int i = tid; i < k; i += NT) c.z[i] = fma(-c.S[tri(k, i)], dk, c.z[i]);
        if (tid == 0) c.r[k] = dk;
        __syncthreads();
    }
    // d_l = -(g_l + C_l . d_g) / (H_ll + tau)
    for (int p = tid; p < P; p += NT) c.d[p] = 0.0;
    __syncthreads();
    for (int i = tid; i < ng; i += NT) c.d[a.glist[i]] = c.r[i];
    for (int i = tid >> 5; i < nl; i += NT / 32) {  // one warp per per-bin parameter, lanes over the globals
        double accv = 0.0;
        if (!c.actl[i]) {
            const double* crow = c.C + (int64_t)i * ng;
            for (int k = tid & 31; k < ng; k += 32) accv = fma(crow[k], c.r[k], accv);
        }
        for (int o = 16; o > 0; o >>= 1) accv += __shfl_xor_sync(0xffffffffu, accv, o);
        if ((tid & 31) == 0) c.d[a.llist[i]] = c.actl[i] ? 0.0 : -c.gl[i] - c.dinv[i] * accv;
    }
    __syncthreads();
    double edm = 0.0, step = 0.0;
    for (int p = tid; p < P; p += NT) {
        edm = fma(-c.g[p], c.d[p], edm);
        step = fmax(step, fabs(c.d[p]));
    }
    edm_out = block_sum(edm, c.red);
    step_out = block_max(step, c.red);
    return true;
}

// mask / data row of fit i
__device__ __forceinline__ const uint8_t* solo_mask(const SoloArgs& a, int P, int64_t i) {
    const int k = a.use_alt ? min((int)a.use_alt[i], a.n_alt) : 0;
    return k ? a.fixed_alt + (int64_t)(k - 1) * P : a.fixed;
}

template <int MAXT>
__global__ void __launch_bounds__(NT, (MAXT <= 1) ? 3 : ((MAXT <= 4) ? 2 : 1))
fit_solo_kernel(DevPlan pl, SoloArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const int P = pl.P;
    const SoloLayout& L = a.L;
    Ctx c;
    c.xs = sm + L.xs; c.xc = sm + L.xc; c.xt = sm + L.xt; c.g = sm + L.g; c.d = sm + L.d; c.hc = sm + L.hc;
    c.Fq = sm + L.Fq; c.base = sm + L.base; c.rS = sm + L.rS; c.FB = sm + L.FB; c.Wt = sm + L.Wt;
    c.wgt = sm + L.wgt; c.w2t = sm + L.w2t; c.sqW = sm + L.sqW; c.lwj = sm + L.lwj; c.lw2 = sm + L.lw2;
    c.gpart = sm + L.gpart; c.Y = sm + L.Y; c.S = sm + L.S; c.r = sm + L.r; c.z = sm + L.z;
    c.linv = sm + L.linv; c.dinv = sm + L.dinv; c.gl = sm + L.gl; c.hl = sm + L.hl; c.red = sm + L.red;
    double* ext = L.small ? sm : (a.scratch + (int64_t)blockIdx.x * a.scratch_stride);
    c.Sorig = ext + L.Sorig; c.Sred = ext + L.Sred; c.C = ext + L.C; c.SegM = ext + L.SegM; c.SegG = ext + L.SegG;
    c.SegN = ext + L.SegN; c.LOCw = ext + L.LOCw;
    int* ip = reinterpret_cast<int*>(sm + L.ints);
    c.segt = ip; c.lgp = ip + TB; c.lmk = reinterpret_cast<unsigned*>(ip + 2 * TB);
    c.actg = ip + 3 * TB; c.actl = c.actg + a.ng; c.flag = c.actl + a.nl;
    __shared__ long long s_prof[8];
    c.prof = s_prof;
    __shared__ int s_fit;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_fit = atomicAdd(a.next, 1);
        __syncthreads();
        const int64_t fit = s_fit;
        if (fit >= a.N) return;
        const uint8_t* mask = solo_mask(a, P, fit);
        const int64_t row = a.data_map ? (int64_t)a.data_map[fit] : ((a.data_rows == 1) ? 0 : fit);
        const double* drow = a.data + row * pl.D;
        const double dconst = a.dconst[row];
        const double* x0 = a.init + (a.init_rows == 1 ? 0 : fit) * P;
        for (int p = tid; p < P; p += NT) {
            double v = x0[p];
            if (!mask[p]) v = fmin(fmax(v, a.lo[p]), a.hi[p]);
            c.xc[p] = v;
        }
        __syncthreads();
        bool fisher = true;
        if (tid < 8) c.prof[tid] = 0;
        __syncthreads();
        double f = solo_sweep<MAXT, true>(pl, a, c, c.xc, fisher, drow, dconst);
        int status = 1, niter = 0, nfail = 0, nkick = 0, nindef = 0;
        double damp = a.damp0, edm_prev = INFINITY;
        bool kick_next = false, chk = false, exact_tried = false;
        if (!isfinite(f)) status = 2;
        const int max_rounds = a.max_iter * 4 + 40;
        for (int round = 0; round < max_rounds && status == 1; ++round) {
            bool kicked = false;
            double edm = 0.0, step = 0.0;
            int attempts = 1;
            if (kick_next) {
                // displacement that breaks an exact symmetry of the current point (a saddle reached
                // along it): small, deterministic, taken unconditionally
                const unsigned salt = 40503u * (unsigned)(nkick + 1);
                for (int p = tid; p < P; p += NT) {
                    const unsigned hsh = ((unsigned)(p + 1) * 2654435761u + salt) >> 8;
                    const double u = (double)(hsh & 0xffffu) / 32768.0 - 1.0;
                    c.d[p] = mask[p] ? 0.0 : 1e-3 * fmax(1.0, fabs(c.xc[p])) * u;
                }
                __syncthreads();
                kick_next = false;
                kicked = true;
                nkick += 1;
                damp = 0.0;
            } else {
                const long long t_s = clock64();
                const bool ok = solo_solve(pl, a, c, mask, damp, attempts, edm, step);
                if (tid == 0) { c.prof[4] += clock64() - t_s; c.prof[7] += 1; }
                if (!ok || !isfinite(edm)) {
                    if (((a.trace && fit < 2) || fit == a.trace_fit) && tid == 0)
                        printf("[hf_fit solo] fit %lld it %d: no direction (factorised %d, attempts %d, edm %.3e, f %.12g)\n",
                               (long long)fit, niter, (int)ok, attempts, edm, f);
                    status = 2; break;
                }
                // a nearly singular exact Hessian shows up as an exploding step: Fisher matrix instead
                if (!fisher && (!(edm <= 100.0 * edm_prev + 1.0) || step > 100.0)) {
                    fisher = true;
                    exact_tried = true;
                    f = solo_sweep<MAXT, true>(pl, a, c, c.xc, fisher, drow, dconst);
                    continue;
                }
                const bool converged = edm < a.tol_edm && step < a.tol_step;
                if (converged && damp > 0.0) { damp = 0.0; continue; }  // look again without the shift: it hides an indefinite Hessian
                if (converged && fisher && !exact_tried) {
                    // the stopping rule is met on the Fisher matrix, which is positive semi-definite by
                    // construction: the exact Hessian of this point decides whether it is a minimum
                    exact_tried = true;
                    fisher = false;
                    f = solo_sweep<MAXT, true>(pl, a, c, c.xc, fisher, drow, dconst);
                    continue;
                }
                const int ncreep = (attempts > 1 && edm < 1e-3) ? nindef + 1 : 0;
                const bool saddle = ((converged && attempts > 1) || ncreep >= 4) && nkick < kMaxKicks;
                nindef = saddle ? 0 : ncreep;
                if (converged && !saddle) { status = 0; break; }
                if (saddle) { kick_next = true; continue; }
            }
            // ---- trial point, value only
            for (int p = tid; p < P; p += NT) {
                double v = c.xc[p];
                if (!mask[p]) v = fmin(fmax(v + c.d[p], a.lo[p]), a.hi[p]);
                c.xt[p] = v;
            }
            __syncthreads();
            const double ft = solo_sweep<MAXT, false>(pl, a, c, c.xt, fisher, drow, dconst);
            double dec = 0.0;
            for (int p = tid; p < P; p += NT) dec = fma(c.g[p], c.xt[p] - c.xc[p], dec);
            dec = block_sum(dec, c.red);
            const double slack = 4e-15 * fmax(1.0, fabs(f));
            const bool ok = isfinite(ft) && (kicked || ft <= f + kArmijo * dec + slack);
            if (((a.trace && fit < 2) || fit == a.trace_fit) && tid == 0)
                printf("[hf_fit solo] fit %lld it %d f %.12g ft-f %.3e dec %.3e edm %.3e step %.3e damp %.3e %s attempts %d %s\n",
                       (long long)fit, niter, f, ft - f, dec, edm, step, damp, fisher ? "fisher" : "exact", attempts,
                       ok ? "accept" : "reject");
            if (kicked && !ok) { status = 0; break; }  // (not even finite: stay where the fit was)
            if (ok) {
                __syncthreads();
                for (int p = tid; p < P; p += NT) c.xc[p] = c.xt[p];
                __syncthreads();
                nfail = 0;
                chk = false;
                exact_tried = false;
                const double dmp = 0.2 * damp;
                damp = (dmp < 1e-6) ? 0.0 : dmp;
                niter += 1;
                // exact Hessian close to the minimum, Fisher matrix (positive semi-definite) far from it
                fisher = !(edm < a.exact_below);  // (a kick carries edm = 0)
                if (!kicked) edm_prev = edm;
                f = solo_sweep<MAXT, true>(pl, a, c, c.xc, fisher, drow, dconst);
                if (!isfinite(f)) { status = 2; break; }
                if (niter >= a.max_iter) { status = 1; break; }
            } else {
                const int nr = nfail + 1;
                if (edm < 1e-7 && damp > 0.0) {
                    damp = 0.0;  // (same test once more without the damping: it would hide an indefinite Hessian)
                    chk = true;
                } else if (edm < 1e-7 || chk) {
                    // predicted decrease below the rounding floor of f -- or, at a kink, an undamped step
                    // that cannot be realised where the damped one had nothing left to gain: converged
                    if (attempts > 1 && nkick < kMaxKicks) kick_next = true;  // ... on a saddle
                    else { status = 0; break; }
                } else if (nr > 60) {
                    status = 2; break;
                } else {
                    damp = fmax(8.0 * damp, 1e-2);
                    nfail = nr;
                }
            }
        }
        __syncthreads();
        if (((a.trace && fit < 2) || fit == a.trace_fit) && tid == 0)
            printf("[hf_fit solo] fit %lld done: status %d niter %d f %.12g | cycles: full sweep prolog %lld tiles %lld flush %lld "
                   "channels %lld | solve %lld (%lld calls) | value sweep prolog %lld tiles+sum %lld\n",
                   (long long)fit, status, niter, f, c.prof[0], c.prof[1], c.prof[2], c.prof[3], c.prof[4], c.prof[7],
                   c.prof[5], c.prof[6]);
        for (int p = tid; p < P; p += NT) a.x_out[fit * P + p] = c.xc[p];
        if (tid == 0) {
            a.fun_out[fit] = f;
            a.status_out[fit] = status;
            if (a.niter_out) a.niter_out[fit] = niter;
        }
    }
}

SoloLayout solo_layout(const hf_plan* plan, int ng, int nl, int Hpad, int nj, int Hc, int npart, bool small) {
    const DevPlan& pl = plan->d;
    SoloLayout L;
    memset(&L, 0, sizeof(L));
    int o = 0;
    auto take = [&](int n) { const int at = o; o += (n + 1) & ~1; return at; };
    const int P = pl.P, H = pl.H, S = pl.S, NSEG = pl.NSEG;
    L.xs = take(P); L.xc = take(P); L.xt = take(P); L.g = take(P); L.d = take(P);
    L.hc = take(5 * std::max(H, 1)); L.Fq = take(S * NSEG);
    L.base = take(SMX * TB); L.rS = take(SMX * TB); L.FB = take(SMX * TB);
    L.Wt = take(TB); L.wgt = take(TB); L.w2t = take(TB); L.sqW = take(TB); L.lwj = take(TB); L.lw2 = take(TB);
    L.gpart = take(2 * NT);
    const int npk = ng * (ng + 1) / 2;
    // the Y region doubles as DL | D2 | U0 | Msm | G2sm after the sweep and as the staging tile of C
    L.ltile = small ? 0 : std::max(1, std::min(nl, 16));
    L.ywords = std::max(std::max(H > 0 ? TB * Hpad : 0, 3 * SMX * nj + SMX * SMX + SMX), L.ltile * ng);
    L.Y = take(L.ywords);
    L.S = take(npk); L.r = take(ng); L.z = take(ng); L.linv = take(ng);
    L.dinv = take(std::max(nl, 1)); L.gl = take(std::max(nl, 1)); L.hl = take(std::max(nl, 1));
    L.red = take(32);
    L.ints = take((3 * TB + ng + nl + 8 + 1) / 2 + 1);
    L.small = small ? 1 : 0;
    // external block (shared memory when small, per-CTA global scratch else); SegM | SegG | SegN | LOCw contiguous
    int e = small ? o : 0;
    auto take_e = [&](int n) { const int at = e; e += (n + 1) & ~1; return at; };
    L.Sorig = take_e(npk); L.Sred = take_e(npk); L.C = take_e(std::max(nl * ng, 1));
    {
        const int nseg_words = NSEG * S * S + NSEG * S + NSEG * S * npart * Hc + pl.ah_nbl * S;
        L.SegM = take_e(nseg_words);
        L.SegG = L.SegM + NSEG * S * S;
        L.SegN = L.SegG + NSEG * S;
        L.LOCw = L.SegN + NSEG * S * npart * Hc;
    }
    if (small) { o = e; L.scratch_doubles = 0; }
    else L.scratch_doubles = e + 8;
    L.total = o;
    return L;
}

__global__ void solo_zero_kernel(int* p) { *p = 0; }

}  // namespace

bool hf_fit_solo_supported(hf_plan* plan) {
    const DevPlan& pl = plan->d;
    if (!pl.ah_ok) return false;
    if (plan->fd_ncolors > 1) return false;  // two per-bin parameters on one bin: their block is not diagonal
    if (const char* e = getenv("HF_FIT_SOLO"))
        if (e[0] == '0') return false;
    return true;
}

int hf_fit_solo_launch(hf_plan* plan, const double* data, int64_t data_rows, const double* dconst,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* fixed, const uint8_t* fixed_alt, const uint8_t* use_alt, int n_alt,
                       const int32_t* data_map, int64_t N, const hf_fit_opts& opts, double* x_out, double* fun_out,
                       int32_t* status_out, int32_t* niter_out, cudaStream_t st, int* handled) {
    *handled = 0;
    const DevPlan& pl = plan->d;
    const int P = pl.P;
    // global / per-bin split (plan->fd_color: >= 0 for a parameter that acts on exactly one bin)
    if (!plan->solo_dev) {
        std::vector<int32_t> tab(2 * (size_t)P, -1), glist, llist;
        for (int p = 0; p < P; ++p) {
            if (plan->fd_color[p] >= 0) { tab[P + p] = (int)llist.size(); llist.push_back(p); }
            else { tab[p] = (int)glist.size(); glist.push_back(p); }
        }
        plan->solo_ng = (int)glist.size();
        plan->solo_nl = (int)llist.size();
        tab.insert(tab.end(), glist.begin(), glist.end());
        tab.insert(tab.end(), llist.begin(), llist.end());
        tab.push_back(0);
        HF_CUDA_OK(cudaMalloc(&plan->solo_dev, sizeof(int32_t) * tab.size()));
        HF_CUDA_OK(cudaMemcpy(plan->solo_dev, tab.data(), sizeof(int32_t) * tab.size(), cudaMemcpyHostToDevice));
    }
    const int ng = plan->solo_ng, nl = plan->solo_nl;
    const int H = pl.H;
    int Hpad = (H + 7) / 8 * 8;
    while (Hpad % 16 != 8) Hpad += 8;  // row stride == 8 mod 16 doubles: conflict-free DMMA fragment loads
    const int Hc = (H + 31) & ~31;
    const int npart = (H > 0) ? std::max(1, NT / Hc) : 1;
    const int nj = pl.ah_nsfp;
    const int nt8 = (H + 7) / 8, ntile = nt8 * (nt8 + 1) / 2;
    const int maxt = (H == 0) ? 0 : (ntile + NT / 32 - 1) / (NT / 32);
    if (maxt > 16) return HF_OK;
    SoloLayout L = solo_layout(plan, ng, nl, Hpad, nj, Hc, npart, true);
    const size_t small_limit = 100 * 1024;  // two or more CTAs per SM
    if (sizeof(double) * (size_t)L.total > small_limit) L = solo_layout(plan, ng, nl, Hpad, nj, Hc, npart, false);
    const size_t smem = sizeof(double) * (size_t)L.total + 16;
    if (smem > 225 * 1024) return HF_OK;  // the lock-step driver takes the fit

    SoloArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.data = data; a.data_rows = data_rows; a.data_map = data_map; a.dconst = dconst;
    a.init = init; a.init_rows = init_rows; a.lo = lo; a.hi = hi;
    a.fixed = fixed; a.fixed_alt = fixed_alt; a.use_alt = use_alt; a.n_alt = fixed_alt ? n_alt : 0;
    a.tol_edm = opts.tol_edm; a.tol_step = opts.tol_step; a.exact_below = 10.0; a.damp0 = 0.0;
    if (const char* e = getenv("HF_FIT_EXACT_BELOW")) a.exact_below = atof(e);
    if (const char* e = getenv("HF_FIT_DAMP0")) a.damp0 = atof(e);
    a.max_iter = opts.max_iter; a.trace = opts.verbose >= 3 ? 1 : 0;
    a.trace_fit = -1;
    if (const char* e = getenv("HF_FIT_TRACE_FIT")) a.trace_fit = atoll(e);
    a.ng = ng; a.nl = nl; a.npk = ng * (ng + 1) / 2; a.Hpad = Hpad; a.nj = nj; a.Hc = Hc; a.npart = npart;
    a.gpos = plan->solo_dev; a.lpos = plan->solo_dev + P; a.glist = plan->solo_dev + 2 * P;
    a.llist = plan->solo_dev + 2 * P + ng; a.next = plan->solo_dev + 2 * P + ng + nl;
    a.x_out = x_out; a.fun_out = fun_out; a.status_out = status_out; a.niter_out = niter_out;
    a.L = L;

    void (*kern)(DevPlan, SoloArgs) = nullptr;
    if (maxt == 0) kern = fit_solo_kernel<0>;
    else if (maxt == 1) kern = fit_solo_kernel<1>;
    else if (maxt <= 4) kern = fit_solo_kernel<4>;
    else kern = fit_solo_kernel<16>;
    HF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    HF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    occ = std::max(occ, 1);
    const int64_t grid = std::min<int64_t>(N, (int64_t)plan->sm_count * occ);
    if (!L.small) {
        const size_t need = sizeof(double) * (size_t)L.scratch_doubles * (size_t)grid;
        if (need > plan->solo_scratch_bytes) {
            if (plan->solo_scratch) cudaFree(plan->solo_scratch);
            plan->solo_scratch = nullptr;
            plan->solo_scratch_bytes = 0;
            cudaError_t e = cudaMalloc(&plan->solo_scratch, need);
            if (e != cudaSuccess) {
                hf_set_error("cudaMalloc(%zu) for the per-fit scratch: %s", need, cudaGetErrorString(e));
                return HF_ERR_NOMEM;
            }
            plan->solo_scratch_bytes = need;
        }
        a.scratch = reinterpret_cast<double*>(plan->solo_scratch);
        a.scratch_stride = L.scratch_doubles;
    }
    solo_zero_kernel<<<1, 1, 0, st>>>(a.next);
    kern<<<(unsigned)grid, NT, smem, st>>>(pl, a);
    plan->launches += 2;
    HF_CUDA_OK(cudaGetLastError());
    plan->solo_last_grid = (int)grid;
    plan->solo_last_occ = occ;
    plan->solo_last_smem = (int)smem;
    *handled = 1;
    return HF_OK;
}
