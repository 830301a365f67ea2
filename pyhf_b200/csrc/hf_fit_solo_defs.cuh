// Definitions shared by the translation units of the per-fit minimiser (hf_fit_solo.cu holds the host
// side and the 256-thread kernel instances, hf_fit_solo128.cu the 128-thread ones).
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "hf_internal.cuh"

namespace solo {


constexpr int TB = 32;             // bins per tile
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
    int wstride;                                 // row stride of the staged rows (== 8 mod 16 doubles)
    int ngs;                                     // row stride of C
    int lowrank, DLall, Qc, vq, Tq, uq, ppos;    // low-rank form of the per-bin block: flag and offsets
    int tdec;                                    // tile decode table
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
    int ngs;   // row stride of C
    int ntdec;  // entries of the tile decode table
    int bigch, njs, max_lc;  // tensor-core form of the scalar-factor blocks: flag, row stride of DL / U1, locals per channel
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
    // low-rank form of the per-bin block (plans without histosys)
    double *DLall, *Qc, *vq, *Tq, *uq;
    int *bl2l, *blseg, *gposj;          // [n_bl] position among the per-bin parameters / channel; [nj] position among the globals
    unsigned char *pj1, *pj2;           // [nj (nj + 1) / 2] the pair of scalar-factor parameters of a packed entry
    unsigned short* ppos;               // ... and its place in the packed S_gg
    unsigned short* tdec;               // [ntdec] (row << 8 | column) of the t-th tile of a lower triangle
    long long* prof;  // [16] clock64 sums per phase (thread 0; printed with the trace)
};


typedef void (*SoloKernel)(DevPlan, SoloArgs);
struct SoloVariant { int nt, maxt; SoloKernel kern; };
// kernel instances compiled in hf_fit_solo128.cu
int solo128_variants(const SoloVariant** out);

}  // namespace solo
