// Analytic Hessian of 2*NLL for the batched minimiser (K3): one CTA per (fit, current point).
//
// Replaces the finite-difference curvature of hf_fit.cu (n_global + n_colours gradient
// evaluations per matrix: 201 K2 evaluations = 7e8 flop on C4) by the closed form, ~1.5e7 flop:
//
//   lambda_b = sum_s F_{s,c(b)}(theta) * B_{s,b}(gamma) * base_{s,b}(alpha)          (pdf.py:675-713)
//   H = sum_b [ W_b J_b J_b^T + w2_b Hess(lambda_b) ] + constraint curvature
//   exact  : W_b = 2 n_b / lambda_b^2 , w2_b = 2 (1 - n_b / lambda_b)
//   Fisher : n_b := lambda_b (the Asimov data of the point), i.e. W_b = 2 / lambda_b , w2_b = 0
//
// and the structure of a HistFactory rate does the rest:
//   * histosys x histosys  = Y^T Y with Y[b][h] = sqrt(W_b) a_b[h],  a_b[h] = sum_r FB_{r,b} dT_{h,r,b}:
//     the only dense product (H x H x bins) -- on the fp64 tensor-core path (DMMA m8n8k4), accumulators
//     in registers across the whole bin sweep;
//   * scalar factors (normsys / normfactor / lumi) are constant inside a channel c, so their columns
//     are J_b = DL_c^T r_b with r_b the S per-sample rates: every block they take part in collapses to
//     S x S (M_c = sum_b W r r^T, G_c = sum_b w2 r) and S x H (N_c) sums per channel, contracted with
//     the per-channel table of dlog f / dtheta afterwards;
//   * per-bin parameters (staterror / shapesys / shapefactor) act on one bin each: one row per parameter.
// A multiplicative parameter that sits exactly at 0 (the POI on its bound in q~ fits) is evaluated at
// 2^-100: every quantity is then an exact power-of-two rescaling of its limit, no 0/0.
//
// Reference being replaced: none (the reference's SLSQP builds a BFGS estimate from forward-difference
// gradients, optimize/opt_scipy.py:96-105); checked against central differences of the analytic
// gradient (hf_hessian) in tests/test_gpu_hessian.py.
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include "hf_internal.cuh"

namespace {

constexpr int TB = 32;      // bins per tile
constexpr int NTHR = 256;
constexpr int SMX = HF_AH_SMAX;   // samples
constexpr int MAXT = 16;    // 8x8 tiles of the histosys block per warp (upper triangle): H <= 120
constexpr int KLX = HF_AH_KLMAX;  // per-bin parameters on one bin
constexpr double TINY = 7.888609052210118e-31;  // 2^-100

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

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

struct AhArgs {
    const double* X;      // [.][P] points; request i reads row list[i] (or i)
    const int* list;      // request -> row of X / fit index, or null
    const int* n_dev;     // device-side request count (or null: nreq)
    int nreq;
    const int* mode;      // per fit: 1 = Fisher (Asimov data of the point), 2 = exact; null = Fisher
    int mode_all;         // used when mode == null: 1 or 2
    const double* data;   // [data_rows][D]
    int64_t data_rows;
    const int* data_map;  // fit -> data row (indexed toy0 + fit), or null
    int64_t toy0;
    double* H;            // [.][P][P] (row = fit index) or one matrix when to_shared
    int to_shared;
    double* scratch;      // per request
    int64_t scratch_stride;
    const uint8_t* fixed_struct;  // [P] rows / columns forced to the identity's, or null
};

__global__ void __launch_bounds__(NTHR, 2)
hf_hess_analytic_kernel(DevPlan pl, AhArgs a, int Hpad, int nj) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int nreq = a.n_dev ? min(*a.n_dev, a.nreq) : a.nreq;
    if ((int)blockIdx.x >= nreq) return;
    const int64_t toy = a.list ? (int64_t)a.list[blockIdx.x] : (int64_t)blockIdx.x;
    const int P = pl.P, NB = pl.NB, S = pl.S, NSEG = pl.NSEG, H = pl.H, R = pl.R, K = pl.K;
    const int mode = a.mode ? a.mode[toy] : a.mode_all;
    const bool fisher = (mode != 2);
    const int64_t row = a.data_map ? (int64_t)a.data_map[a.toy0 + toy]
                                   : ((a.data_rows == 1) ? 0 : a.toy0 + toy);
    const double* __restrict__ drow = a.data + row * pl.D;
    const double* __restrict__ x = a.X + toy * P;
    double* __restrict__ Hm = a.to_shared ? a.H : (a.H + toy * (int64_t)P * P);
    // scratch of this request (global, L2 resident)
    double* SegM = a.scratch + (int64_t)blockIdx.x * a.scratch_stride;  // [NSEG][S*S]
    double* SegG = SegM + (int64_t)NSEG * S * S;                         // [NSEG][S]
    double* SegN = SegG + (int64_t)NSEG * S;                             // [NSEG][S][H]
    double* LOCw = SegN + (int64_t)NSEG * S * H;                         // [n_bl][S]
    const int64_t nscr = (int64_t)NSEG * S * S + (int64_t)NSEG * S + (int64_t)NSEG * S * H + (int64_t)pl.ah_nbl * S;

    // shared memory
    double* xs = sm;                        // [P]
    double* hc = xs + P;                    // [H][5]  c1 c2 dc1 dc2 ddc2
    double* Fq = hc + 5 * H;                // [S*NSEG]
    double* base = Fq + S * NSEG;           // [SMX][TB]  histosys delta of a row
    double* rS = base + SMX * TB;           // [SMX][TB]  rate of a sample
    double* FB = rS + SMX * TB;             // [SMX][TB]  F * B of a sample
    double* Wt = FB + SMX * TB;             // [TB]
    double* w2t = Wt + TB;
    double* sqW = w2t + TB;
    double* lwj = sqW + TB;                 // [TB][KLX]  W * J_g
    double* lw2 = lwj + TB * KLX;           // [TB][KLX]  w2 / gamma
    double* Y = lw2 + TB * KLX;             // [TB][Hpad]; after the sweep: DL | D2 | U0 | Msm | G2sm
    const int ywords = max(TB * Hpad, 3 * SMX * nj + SMX * SMX + SMX);
    int* segt = reinterpret_cast<int*>(Y + ywords);  // [TB]
    int* lng = segt + TB;                   // [TB]       locals on the bin
    int* lgp = lng + TB;                    // [TB][KLX]  parameter
    unsigned* lmk = reinterpret_cast<unsigned*>(lgp + TB * KLX);  // [TB][KLX] sample mask

    for (int64_t i = tid; i < (int64_t)P * P; i += NTHR) Hm[i] = 0.0;
    for (int64_t i = tid; i < nscr; i += NTHR) SegM[i] = 0.0;
    for (int i = tid; i < TB * Hpad; i += NTHR) Y[i] = 0.0;
    for (int p = tid; p < P; p += NTHR) {
        double v = x[p];
        if (pl.ah_par_mult[p] && fabs(v) < TINY) v = (v < 0.0) ? -TINY : TINY;
        xs[p] = v;
    }
    __syncthreads();
    for (int h = tid; h < H; h += NTHR) {
        const double al = xs[pl.hs_par[h]];
        double c1, c2, d1, d2;
        hf_hs_coeffs(pl.hs_code, al, c1, c2, d1, d2);
        double dd = 0.0;
        if (pl.hs_code == HF_HS_CODE4P && !(al > 1.0) && !(al < -1.0)) {
            const double a2 = al * al;
            dd = a2 * (90.0 * a2 - 120.0) + 30.0;
        }
        hc[5 * h] = c1; hc[5 * h + 1] = c2; hc[5 * h + 2] = d1; hc[5 * h + 3] = d2; hc[5 * h + 4] = dd;
    }
    for (int q = tid; q < S * NSEG; q += NTHR) {
        double prod = 1.0, logsum = 0.0;
        for (int e = pl.sf_ptr[q]; e < pl.sf_ptr[q + 1]; ++e)
            hf_sf_accum(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, xs[pl.sf_par[e]], prod, logsum);
        if (logsum != 0.0) prod *= exp(logsum);
        Fq[q] = prod;
    }
    __syncthreads();

    // ---------------------------------------------------------------- sweep over the bins
    // thread roles that carry state across tiles
    const int Hc = (H + 31) & ~31;                       // histosys columns rounded to warps
    const int npart = (H > 0) ? max(1, NTHR / Hc) : 1;   // bins of a tile are split over `npart` thread groups
    const int my_h = (H > 0) ? (tid % Hc) : 0, my_part = (H > 0) ? (tid / Hc) : 0;
    const bool h_thread = H > 0 && my_h < H && my_part < npart;
    double Nacc[SMX];  // N_{c,s}[h] of the current segment
#pragma unroll
    for (int s = 0; s < SMX; ++s) Nacc[s] = 0.0;
    double ddacc = 0.0;
    int ncur = -1;
    double macc = 0.0;  // pair thread (s, s') or G thread s
    int mcur = -1;
    const bool pair_thread = tid < S * S, g_thread = tid >= SMX * SMX && tid < SMX * SMX + S;
    const int ps = tid / S, ps2 = tid - (tid / S) * S, gs = tid - SMX * SMX;
    double acc[MAXT][2];
#pragma unroll
    for (int k = 0; k < MAXT; ++k) acc[k][0] = acc[k][1] = 0.0;
    const int nt8 = (H + 7) / 8, ntile = nt8 * (nt8 + 1) / 2;
    double dc1 = 0.0, dc2 = 0.0, ddc2 = 0.0;
    if (h_thread) { dc1 = hc[5 * my_h + 2]; dc2 = hc[5 * my_h + 3]; ddc2 = hc[5 * my_h + 4]; }

    for (int b0 = 0; b0 < NB; b0 += TB) {
        // ---- (1) histosys delta of every (row, bin) of the tile
        if (warp < R) {
            const int b = min(b0 + lane, NB - 1);
            const double* p1 = pl.t1 + (int64_t)warp * NB + b;
            const double* p2 = pl.t2 + (int64_t)warp * NB + b;
            const int64_t hstride = (int64_t)R * NB;
            double s0 = 0.0, s1 = 0.0;
            int h = 0;
            for (; h + 1 < H; h += 2) {
                s0 = fma(hc[5 * h], __ldg(p1 + h * hstride), fma(hc[5 * h + 1], __ldg(p2 + h * hstride), s0));
                s1 = fma(hc[5 * h + 5], __ldg(p1 + (h + 1) * hstride), fma(hc[5 * h + 6], __ldg(p2 + (h + 1) * hstride), s1));
            }
            if (h < H) s0 = fma(hc[5 * h], __ldg(p1 + h * hstride), fma(hc[5 * h + 1], __ldg(p2 + h * hstride), s0));
            base[warp * TB + lane] = s0 + s1;
        }
        __syncthreads();
        // ---- (2a) rates, weights, per-bin parameters
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
                    if (hr >= 0) bs += base[hr * TB + bl];
                    double bf = 1.0;
                    for (int k = 0; k < K; ++k) {
                        const int p = pl.bf_par[((int64_t)k * S + s) * NB + bb];
                        if (p >= 0) bf *= xs[p];
                    }
                    const double fb = Fq[s * NSEG + seg] * bf;
                    rr[s] = valid ? fb * bs : 0.0;
                    FB[s * TB + bl] = valid ? fb : 0.0;
                    rS[s * TB + bl] = rr[s];
                    lam += rr[s];
                }
            }
            const double n = fisher ? lam : drow[bb];
            double W = 0.0, w2 = 0.0;
            if (valid) {
                const double il = 1.0 / lam;
                W = 2.0 * n * il * il;
                w2 = 2.0 * (1.0 - n * il);
            }
            Wt[bl] = W; w2t[bl] = w2; sqW[bl] = sqrt(fmax(W, 0.0));
            segt[bl] = seg;
            int nl = 0;
            if (valid) {
                const int i0 = pl.ah_bl_ptr[bb], i1 = pl.ah_bl_ptr[bb + 1];
                nl = i1 - i0;
                double Jg[KLX], ig[KLX];
#pragma unroll
                for (int k = 0; k < KLX; ++k) {
                    Jg[k] = ig[k] = 0.0;
                    if (k < nl) {
                        const int gp = pl.ah_bl_par[i0 + k];
                        const unsigned mk = pl.ah_bl_mask[i0 + k];
                        ig[k] = 1.0 / xs[gp];
                        double sum = 0.0;
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s < S && ((mk >> s) & 1u)) sum += rr[s];
                        Jg[k] = sum * ig[k];
                        lgp[bl * KLX + k] = gp;
                        lmk[bl * KLX + k] = mk;
                        lwj[bl * KLX + k] = W * Jg[k];
                        lw2[bl * KLX + k] = w2 * ig[k];
                        Hm[(int64_t)gp * P + gp] += W * Jg[k] * Jg[k];
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s < S) LOCw[(int64_t)(i0 + k) * S + s] = rr[s] * (W * Jg[k] + (((mk >> s) & 1u) ? w2 * ig[k] : 0.0));
                    }
                }
                // two per-bin parameters on the same bin
#pragma unroll
                for (int k = 0; k < KLX; ++k)
#pragma unroll
                    for (int k2 = k + 1; k2 < KLX; ++k2)
                        if (k2 < nl) {
                            const unsigned both = pl.ah_bl_mask[i0 + k] & pl.ah_bl_mask[i0 + k2];
                            double sum = 0.0;
#pragma unroll
                            for (int s = 0; s < SMX; ++s)
                                if (s < S && ((both >> s) & 1u)) sum += rr[s];
                            Hm[(int64_t)pl.ah_bl_par[i0 + k] * P + pl.ah_bl_par[i0 + k2]] +=
                                W * Jg[k] * Jg[k2] + w2 * sum * ig[k] * ig[k2];
                        }
            }
            lng[bl] = nl;
        }
        __syncthreads();
        // ---- (2b) per-segment S x S sums  M_c = sum_b W r r^T ,  G_c = sum_b w2 r
        if (pair_thread || g_thread) {
            for (int bl = 0; bl < TB && b0 + bl < NB; ++bl) {
                const int seg = segt[bl];
                if (seg != mcur) {
                    if (mcur >= 0) {
                        if (pair_thread) SegM[(int64_t)mcur * S * S + tid] += macc;
                        else SegG[(int64_t)mcur * S + gs] += macc;
                    }
                    macc = 0.0;
                    mcur = seg;
                }
                if (pair_thread) macc = fma(Wt[bl] * rS[ps * TB + bl], rS[ps2 * TB + bl], macc);
                else macc = fma(w2t[bl], rS[gs * TB + bl], macc);
            }
        }
        // ---- (3) histosys columns of the tile: a_b[h], the S x H sums N_c, per-bin parameter rows
        if (h_thread) {
            for (int bl = my_part; bl < TB; bl += npart) {
                const int b = b0 + bl;
                if (b >= NB) { Y[bl * Hpad + my_h] = 0.0; continue; }
                const int seg = segt[bl];
                if (seg != ncur) {
                    if (ncur >= 0) {
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s < S && Nacc[s] != 0.0) atomicAdd(&SegN[((int64_t)ncur * S + s) * H + my_h], Nacc[s]);
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
                        const double fb = FB[pl.hs_row_sample[r] * TB + bl];
                        A[r] = fb * fma(dc1, v1, dc2 * v2);
                        asum += A[r];
                        ddl = fma(fb * ddc2, v2, ddl);
                    }
                }
                Y[bl * Hpad + my_h] = sqW[bl] * asum;
                const double W = Wt[bl], w2 = w2t[bl];
#pragma unroll
                for (int s = 0; s < SMX; ++s)
                    if (s < S) Nacc[s] = fma(W * rS[s * TB + bl], asum, Nacc[s]);
#pragma unroll
                for (int r = 0; r < SMX; ++r)
                    if (r < R) {
                        // (static register index: scan the samples for the row's one)
                        const int sr = pl.hs_row_sample[r];
#pragma unroll
                        for (int s = 0; s < SMX; ++s)
                            if (s == sr) Nacc[s] = fma(w2, A[r], Nacc[s]);
                    }
                ddacc = fma(w2, ddl, ddacc);
                const int nl = lng[bl];
                for (int k = 0; k < nl; ++k) {
                    const unsigned mk = lmk[bl * KLX + k];
                    double sumA = 0.0;
#pragma unroll
                    for (int r = 0; r < SMX; ++r)
                        if (r < R && ((mk >> pl.hs_row_sample[r]) & 1u)) sumA += A[r];
                    Hm[(int64_t)lgp[bl * KLX + k] * P + pl.hs_par[my_h]] += lwj[bl * KLX + k] * asum + lw2[bl * KLX + k] * sumA;
                }
            }
        }
        __syncthreads();
        // ---- (4) histosys x histosys block: acc += Y^T Y on the fp64 tensor-core path
        if (H > 0) {
#pragma unroll
            for (int k = 0; k < MAXT; ++k) {
                const int tile = warp + 8 * k;
                if (tile < ntile) {
                    int ti, tj;
                    tri_decode(tile, tj, ti);  // ti <= tj : row tile, column tile
                    const double* pa = Y + t4 * Hpad + 8 * ti + g8;
                    const double* pb = Y + t4 * Hpad + 8 * tj + g8;
#pragma unroll
                    for (int k0 = 0; k0 < TB; k0 += 4) dmma884(acc[k][0], acc[k][1], pa[k0 * Hpad], pb[k0 * Hpad]);
                }
            }
        }
        __syncthreads();
    }
    // ---- flush the per-segment sums still in registers
    if (mcur >= 0) {
        if (pair_thread) SegM[(int64_t)mcur * S * S + tid] += macc;
        else if (g_thread) SegG[(int64_t)mcur * S + gs] += macc;
    }
    if (h_thread) {
        if (ncur >= 0) {
#pragma unroll
            for (int s = 0; s < SMX; ++s)
                if (s < S && Nacc[s] != 0.0) atomicAdd(&SegN[((int64_t)ncur * S + s) * H + my_h], Nacc[s]);
        }
        if (ddacc != 0.0) atomicAdd(&Hm[(int64_t)pl.hs_par[my_h] * P + pl.hs_par[my_h]], ddacc);
    }
    __syncthreads();
    // ---- (5) histosys block out of the registers (upper triangle; the other one comes from the closing symmetrisation)
    if (H > 0) {
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
            const int tile = warp + 8 * k;
            if (tile < ntile) {
                int ti, tj;
                tri_decode(tile, tj, ti);
                const int hi = 8 * ti + g8;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int hj = 8 * tj + 2 * t4 + c;
                    if (hi < H && hj < H && hi <= hj) Hm[(int64_t)pl.hs_par[hi] * P + pl.hs_par[hj]] += acc[k][c];
                }
            }
        }
    }
    __syncthreads();
    // ---------------------------------------------------------------- scalar-factor blocks, channel by channel
    double* DL = Y;                    // [S][nj]
    double* D2 = DL + SMX * nj;        // [S][nj]
    double* U0 = D2 + SMX * nj;        // [S][nj]   M_c DL
    double* Msm = U0 + SMX * nj;       // [S][S]
    double* G2sm = Msm + SMX * SMX;    // [S]
    if (nj > 0) {
        const int npair = nj * (nj + 1) / 2;
        for (int c = 0; c < NSEG; ++c) {
            for (int i = tid; i < 2 * SMX * nj; i += NTHR) DL[i] = 0.0;  // DL and D2
            if (tid < S * S) Msm[tid] = SegM[(int64_t)c * S * S + tid];
            if (tid < S) G2sm[tid] = SegG[(int64_t)c * S + tid];
            __syncthreads();
            for (int s = 0; s < S; ++s) {
                const int q = s * NSEG + c;
                for (int e = pl.sf_ptr[q] + tid; e < pl.sf_ptr[q + 1]; e += NTHR) {
                    double dl, d2;
                    sf_dlogs(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, xs[pl.sf_par[e]], dl, d2);
                    const int j = pl.ah_sf_j[e];
                    DL[s * nj + j] = dl;
                    D2[s * nj + j] = d2;
                }
            }
            __syncthreads();
            for (int i = tid; i < S * nj; i += NTHR) {
                const int s = i / nj, j = i - s * nj;
                double u = 0.0;
                for (int s2 = 0; s2 < S; ++s2) u = fma(Msm[s * S + s2], DL[s2 * nj + j], u);
                U0[i] = u;
            }
            __syncthreads();
            // scalar x scalar (each unordered pair once)
            for (int e = tid; e < npair; e += NTHR) {
                int j2, j1;
                tri_decode(e, j2, j1);  // j1 <= j2
                double v = 0.0;
                if (j1 != j2) {
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j1], fma(G2sm[s], DL[s * nj + j2], U0[s * nj + j2]), v);
                } else {
                    for (int s = 0; s < S; ++s) {
                        const double dl = DL[s * nj + j1];
                        v = fma(dl, U0[s * nj + j1], v);
                        v = fma(G2sm[s], dl * dl + D2[s * nj + j1], v);
                    }
                }
                if (v != 0.0) Hm[(int64_t)pl.ah_sfp_list[j1] * P + pl.ah_sfp_list[j2]] += v;
            }
            __syncthreads();
            // histosys x scalar
            for (int e = tid; e < H * nj; e += NTHR) {
                const int j = e / H, h = e - j * H;
                double v = 0.0;
                for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j], SegN[((int64_t)c * S + s) * H + h], v);
                if (v != 0.0) {
                    const int p1 = pl.hs_par[h], p2 = pl.ah_sfp_list[j];
                    Hm[(int64_t)p1 * P + p2] += (p1 == p2) ? 2.0 * v : v;
                }
            }
            // per-bin parameters of this channel x scalar
            {
                const int i0 = pl.ah_bl_ptr[pl.ah_seg_first[c]], i1 = pl.ah_bl_ptr[pl.ah_seg_first[c + 1]];
                for (int e = tid; e < (i1 - i0) * nj; e += NTHR) {
                    const int il = e / nj, j = e - il * nj;
                    double v = 0.0;
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j], LOCw[(int64_t)(i0 + il) * S + s], v);
                    if (v != 0.0) Hm[(int64_t)pl.ah_bl_par[i0 + il] * P + pl.ah_sfp_list[j]] += v;
                }
            }
            __syncthreads();
        }
    }
    // ---- constraint curvature (constraints.py:100-141, :219-262)
    for (int p = tid; p < P; p += NTHR) {
        const int ck = pl.par_cons_kind[p];
        if (ck == 1) {
            const double sg = pl.g_sigma[pl.par_cons_idx[p]];
            Hm[(int64_t)p * P + p] += 2.0 / (sg * sg);
        } else if (ck == 2) {
            const int ci = pl.par_cons_idx[p];
            const double th = xs[p];
            Hm[(int64_t)p * P + p] += fisher ? 2.0 * pl.p_factor[ci] / th : 2.0 * drow[NB + pl.p_aux[ci]] / (th * th);
        }
    }
    __syncthreads();
    // ---- symmetrise (every unordered pair was written once, in one of the two triangles)
    const uint8_t* fx = a.fixed_struct;
    for (int64_t e = tid; e < (int64_t)P * (P + 1) / 2; e += NTHR) {
        int i, j;
        tri_decode((int)e, i, j);  // j <= i
        if (i == j) {
            if (fx && fx[i]) Hm[(int64_t)i * P + i] = 1.0;
        } else {
            double v = Hm[(int64_t)i * P + j] + Hm[(int64_t)j * P + i];
            if (fx && (fx[i] || fx[j])) v = 0.0;
            Hm[(int64_t)i * P + j] = v;
            Hm[(int64_t)j * P + i] = v;
        }
    }
}

}  // namespace

size_t hf_hess_analytic_scratch_doubles(const hf_plan* plan) {
    const DevPlan& pl = plan->d;
    return (size_t)pl.NSEG * pl.S * pl.S + (size_t)pl.NSEG * pl.S + (size_t)pl.NSEG * pl.S * pl.H +
           (size_t)pl.ah_nbl * pl.S + 8;
}

int hf_hess_analytic_launch(hf_plan* plan, const double* X, const int* list, const int* n_dev, int nreq,
                            const int* mode, int mode_all, const double* data, int64_t data_rows,
                            const int* data_map, int64_t toy0, double* Hout, int to_shared, double* scratch,
                            const uint8_t* fixed_struct, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    if (!pl.ah_ok) {
        hf_set_error("plan does not qualify for the analytic Hessian");
        return HF_ERR_UNSUPPORTED;
    }
    if (nreq <= 0) return HF_OK;
    int Hpad = (pl.H + 7) / 8 * 8;
    while (Hpad % 16 != 8) Hpad += 8;  // row stride == 8 mod 16 doubles: conflict-free DMMA fragment loads
    const int nj = pl.ah_nsfp;
    const size_t ywords = std::max<size_t>((size_t)TB * Hpad, (size_t)3 * SMX * nj + SMX * SMX + SMX);
    const size_t smem = sizeof(double) * ((size_t)pl.P + 5 * (size_t)pl.H + (size_t)pl.S * pl.NSEG + 3 * SMX * TB +
                                          3 * TB + 2 * TB * KLX + ywords) +
                        sizeof(int) * (2 * TB + 2 * TB * KLX) + 64;
    if (smem > 200 * 1024) {
        hf_set_error("model too large for the analytic-Hessian kernel (%zu bytes of shared memory)", smem);
        return HF_ERR_UNSUPPORTED;
    }
    HF_CUDA_OK(cudaFuncSetAttribute(hf_hess_analytic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AhArgs a;
    a.X = X; a.list = list; a.n_dev = n_dev; a.nreq = nreq; a.mode = mode; a.mode_all = mode_all;
    a.data = data; a.data_rows = data_rows; a.data_map = data_map; a.toy0 = toy0;
    a.H = Hout; a.to_shared = to_shared; a.scratch = scratch;
    a.scratch_stride = (int64_t)hf_hess_analytic_scratch_doubles(plan);
    a.fixed_struct = fixed_struct;
    hf_hess_analytic_kernel<<<(unsigned)nreq, NTHR, smem, st>>>(pl, a, Hpad, nj);
    plan->launches++;
    HF_CUDA_OK(cudaGetLastError());
    return HF_OK;
}
