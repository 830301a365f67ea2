// K3: batched bounded minimiser of 2*NLL -- lock-step projected Newton on the device.
//
// Replaces scipy.optimize.minimize(method="SLSQP") as driven by optimize/opt_scipy.py:46-105 under
// infer/mle.py:69-205.  The algorithm is the one prototyped (and pinned against the reference's
// fits) in oracle/fit_proto.py:
//
//   * every fit advances one trial point per round; all trial points of a round go through ONE
//     batched K2 launch (hf_eval_launch), so histosys tables are shared by all toys of a CTA tile;
//   * curvature = forward differences of the ANALYTIC gradient: the perturbed points of every fit
//     that needs a Hessian form another K2 batch.  Per-bin ("local") parameters that never share
//     a bin are perturbed together (graph colouring), so a Hessian costs n_global + n_colours
//     gradient evaluations instead of P;
//   * far from the minimum the exact Hessian of a HistFactory NLL is often indefinite; there the
//     Fisher matrix (= Hessian on the Asimov dataset of the current point, always PSD) is used;
//     once the estimated distance to the minimum is small the exact Hessian takes over;
//   * one CTA per fit solves the reduced Newton system (active-set projection on the box bounds,
//     fixed parameters frozen) by an in-shared-memory packed Cholesky factorisation with
//     diagonal loading when the matrix is not positive definite;
//   * projected backtracking line search with an Armijo test (one trial per round).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "hf_internal.cuh"

namespace {

constexpr double kFdStep = 1e-5;
constexpr double kArmijo = 1e-4;
constexpr int kDirThreads = 128;

struct FitDev {
    int P, D, ncol, nglob, ncolors;
    int64_t NC;
    // per-fit state
    double *X, *G, *Dd, *XT, *GT, *VT, *F, *T, *EDM, *EDMPREV, *H, *Hshared;
    int *done, *need_dir, *own, *status, *niter, *refresh, *rlist, *counters;
    int *nfail, *force, *rlist2;  // consecutive rejections, (unused), second request list
    int* pending;  // fit waits for its Fisher fallback (computed at the top of the next round)
    int* ownstart; // start point differs from fit 0's: the fit gets its own start-up curvature
    // saddle escape: a fit that meets the stopping rule at a point whose Hessian is NOT positive
    // definite (the factorisation needed a shift) sits on a saddle or a ridge, typically by an exact
    // symmetry of the start point (two systematics with identical templates): it is displaced by a
    // small deterministic perturbation and carries on (at most twice)
    int *indef, *nkick, *kick;  // last factorisation needed a shift | kicks so far | 1 = kick trial in flight, 2 = requested
    int* nindef;                // consecutive near-stationary directions (edm < 1e-3) whose Hessian needed a shift
    int* chk;                   // the stopping rule is being re-checked without damping at this point
    int* hfresh;                // the fit's matrix is the exact Hessian AT its current point
    double* kscale;             // relative size of the next saddle kick (shrinks when a kick lands outside the domain)
    double* damp;                 // Levenberg-Marquardt damping added to the diagonal (0 = pure Newton)
    int *live, *live_next;   // compacted list of fits still running (order is arbitrary)
    const int* islocal;      // [P] 1 when the parameter only couples to the global block (diagonal local block)
    int use_local;           // structured Cholesky: local parameters first, their block is diagonal
    // structure
    const int* colpar;   // [nglob]  global free parameters, one FD column each
    const int* pcolor;   // [P]      colour of a free local parameter, -1 global, -2 fixed
    const int* partner;  // [P][ncolors]
    const int* pcol;     // [P]      FD column index of a global free parameter, else -1
    const double *lo, *hi;
    const uint8_t* fixed;         // [P] mask of the batch
    const uint8_t* fixed_alt;     // [n_alt][P] further masks (fit i uses row use_alt[i] - 1), or null
    const uint8_t* use_alt;       // [N] per fit (global fit index): 0 = `fixed`, k > 0 = row k - 1 of fixed_alt; or null
    int n_alt;                    // rows of fixed_alt (ids beyond it are clamped, never read out of bounds)
    const uint8_t* fixed_struct;  // [P] fixed under every mask of the batch (no FD column)
    const int* data_map;          // [N] data row of each fit (global fit index), or null
    int64_t toy0, data_rows;      // first fit of the current chunk; rows of the data array
    double* chol_scratch;  // global-memory factor storage for n > smem capacity
    int chol_in_smem;
    double tol_edm, tol_step, refresh_ratio, exact_below, damp0;
    int max_iter;
    int trace;  // verbose >= 3: print the state of the last few live fits every round
};

__device__ __forceinline__ double fd_step(double x) { return kFdStep * fmax(1.0, fabs(x)); }

// mask / data row of chunk-local fit `toy`
__device__ __forceinline__ const uint8_t* fit_mask(const FitDev& s, int64_t toy) {
    const int k = s.use_alt ? min((int)s.use_alt[s.toy0 + toy], s.n_alt) : 0;
    return k ? s.fixed_alt + (int64_t)(k - 1) * s.P : s.fixed;
}
__device__ __forceinline__ int64_t fit_data_row(const FitDev& s, int64_t toy) {
    if (s.data_map) return (int64_t)s.data_map[s.toy0 + toy];
    return (s.data_rows == 1) ? 0 : s.toy0 + toy;
}

// ---------------------------------------------------------------- state initialisation
__global__ void fit_init_kernel(FitDev s, const double* __restrict__ init, int64_t init_rows) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.NC * s.P) return;
    const int64_t i = idx / s.P;
    const int p = (int)(idx - i * s.P);
    double v = init[(init_rows == 1 ? 0 : i) * s.P + p];
    if (!fit_mask(s, i)[p]) v = fmin(fmax(v, s.lo[p]), s.hi[p]);
    s.X[idx] = v;
    s.XT[idx] = v;
    if (init_rows != 1 && i > 0) {
        double v0 = init[p];
        if (!fit_mask(s, 0)[p]) v0 = fmin(fmax(v0, s.lo[p]), s.hi[p]);
        if (v != v0) s.ownstart[i] = 1;  // (benign race: every writer stores 1)
    }
    if (p == 0) {
        s.done[i] = 0; s.need_dir[i] = 1; s.own[i] = 0; s.status[i] = 1; s.niter[i] = 0;
        s.refresh[i] = 0; s.T[i] = 1.0; s.EDM[i] = INFINITY; s.EDMPREV[i] = INFINITY;
        s.nfail[i] = 0; s.force[i] = 0; s.damp[i] = s.damp0; s.pending[i] = 0;
        s.indef[i] = 0; s.nkick[i] = 0; s.kick[i] = 0; s.nindef[i] = 0; s.chk[i] = 0; s.hfresh[i] = 0; s.kscale[i] = 1e-3;
    }
}

// F = -2 * logpdf, G = -2 * grad for the start point
__global__ void fit_adopt_kernel(FitDev s) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.NC * s.P) return;
    const int64_t i = idx / s.P;
    s.G[idx] = -2.0 * s.GT[idx];
    if (idx - i * s.P == 0) {
        const double f = -2.0 * s.VT[i];
        s.F[i] = f;
        if (!isfinite(f)) { s.done[i] = 1; s.status[i] = 2; }
    }
}

// ---------------------------------------------------------------- Hessian by finite differences
// points of refresh entry r: r*(ncol+1) + c, c < ncol perturbed columns, c == ncol base point
// (second part of the same grid: the data row of every entry, see fit_fd_rows)
__device__ __forceinline__ void fit_fd_rows(const FitDev& s, const int* __restrict__ list, int64_t nlist,
                                            const int* __restrict__ mode, const double* __restrict__ data,
                                            const double* __restrict__ asimov, double* __restrict__ DR,
                                            int64_t idx);

__global__ void fit_fd_points_kernel(FitDev s, const int* __restrict__ list, int64_t nlist,
                                     double* __restrict__ XP, int* __restrict__ row_idx,
                                     const int* __restrict__ mode, const double* __restrict__ data,
                                     const double* __restrict__ asimov, double* __restrict__ DR) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npts = nlist * (s.ncol + 1);
    if (idx >= npts * s.P) {
        fit_fd_rows(s, list, nlist, mode, data, asimov, DR, idx - npts * s.P);
        return;
    }
    const int64_t pt = idx / s.P;
    const int p = (int)(idx - pt * s.P);
    const int64_t r = pt / (s.ncol + 1);
    const int c = (int)(pt - r * (s.ncol + 1));
    const int64_t toy = list[r];
    double v = s.X[toy * s.P + p];
    if (c < s.nglob) {
        if (s.colpar[c] == p) v += fd_step(v);
    } else if (c < s.ncol) {
        if (s.pcolor[p] == c - s.nglob) v += fd_step(v);
    }
    XP[idx] = v;
    if (p == 0) row_idx[pt] = (int)r;
}

// data row per refresh entry: Asimov (Fisher mode) or the fit's own data (exact mode)
__device__ __forceinline__ void fit_fd_rows(const FitDev& s, const int* __restrict__ list, int64_t nlist,
                                            const int* __restrict__ mode, const double* __restrict__ data,
                                            const double* __restrict__ asimov, double* __restrict__ DR,
                                            int64_t idx) {
    if (idx >= nlist * s.D) return;
    const int64_t r = idx / s.D;
    const int j = (int)(idx - r * s.D);
    const int64_t toy = list[r];
    const int m = mode ? mode[toy] : 1;
    DR[idx] = (m == 1) ? asimov[idx] : data[fit_data_row(s, toy) * s.D + j];
}

__global__ void fit_gather_x_kernel(FitDev s, const int* __restrict__ list, int64_t nlist,
                                    double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nlist * s.P) return;
    const int64_t r = idx / s.P;
    out[idx] = s.X[(int64_t)list[r] * s.P + (idx - r * s.P)];
}

__device__ __forceinline__ double fd_entry(const FitDev& s, const double* __restrict__ gp,
                                           const double* __restrict__ xrow, int i, int j) {
    // d(grad 2NLL)_i / d x_j for free i, j (unsymmetrised); gp = gradients of the entry's points
    const int P = s.P;
    const double* base = gp + (int64_t)s.ncol * P;
    const int cj = s.pcol[j];
    if (cj >= 0) return -2.0 * (gp[(int64_t)cj * P + i] - base[i]) / fd_step(xrow[j]);
    // j local
    const int ci = s.pcol[i];
    if (ci >= 0) return -2.0 * (gp[(int64_t)ci * P + j] - base[j]) / fd_step(xrow[i]);  // symmetry
    const int col = s.pcolor[j];
    if (s.partner[(int64_t)i * s.ncolors + col] != j) return 0.0;
    return -2.0 * (gp[(int64_t)(s.nglob + col) * P + i] - base[i]) / fd_step(xrow[j]);
}

__global__ void fit_fd_assemble_kernel(FitDev s, const int* __restrict__ list, int64_t nlist,
                                       const double* __restrict__ GP, double* __restrict__ Hdst,
                                       int to_shared) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t PP = (int64_t)s.P * s.P;
    if (idx >= nlist * PP) return;
    const int64_t r = idx / PP;
    const int64_t e = idx - r * PP;
    const int i = (int)(e / s.P), j = (int)(e - (int64_t)i * s.P);
    const int64_t toy = list[r];
    double v;
    if (s.fixed_struct[i] || s.fixed_struct[j]) {
        v = (i == j) ? 1.0 : 0.0;
    } else {
        const double* gp = GP + r * (int64_t)(s.ncol + 1) * s.P;
        const double* xrow = s.X + toy * s.P;
        v = 0.5 * (fd_entry(s, gp, xrow, i, j) + fd_entry(s, gp, xrow, j, i));
    }
    double* dst = to_shared ? Hdst : (Hdst + toy * PP);
    dst[e] = v;
}

// ---------------------------------------------------------------- Newton direction (one CTA per fit)
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i

// packed lower-triangular Cholesky (Crout by columns), 2 barriers per column.
// The first nloc variables are per-bin ("local") parameters that never share a bin: their block of
// the Hessian is diagonal, so their columns need no dot products and their rows no cross terms
// (the factor of a C3 fit costs ~25k FMA instead of 288k).
__device__ bool chol_factor(double* A, int n, int nloc, int tid, int nth) {
    // local block: all its columns are independent of each other -> two parallel steps
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j = tid; j < nloc; j += nth) {
        const double piv = A[tri(j, j)];
        if (!(piv > 0.0) || !isfinite(piv)) s_bad = 1;
        else A[tri(j, j)] = sqrt(piv);
    }
    __syncthreads();
    if (s_bad) return false;
    {
        const int ng = n - nloc;
        for (int e = tid; e < ng * nloc; e += nth) {
            const int i = nloc + e / nloc, j = e - (e / nloc) * nloc;
            A[tri(i, j)] /= A[tri(j, j)];
        }
    }
    __syncthreads();
    // dense (global) block: Crout by columns
    for (int j = nloc; j < n; ++j) {
        for (int i = j + tid; i < n; i += nth) {
            const double* ri = A + tri(i, 0);
            const double* rj = A + tri(j, 0);
            double acc = ri[j];
            for (int k = 0; k < j; ++k) acc = fma(-ri[k], rj[k], acc);
            A[tri(i, j)] = acc;
        }
        __syncthreads();
        const double piv = A[tri(j, j)];
        if (!(piv > 0.0) || !isfinite(piv)) {
            __syncthreads();
            return false;
        }
        const double l = sqrt(piv);
        __syncthreads();
        for (int i = j + tid; i < n; i += nth) A[tri(i, j)] = (i == j) ? l : A[tri(i, j)] / l;
        __syncthreads();
    }
    __syncthreads();
    return true;
}

constexpr int kMaxKicks = 2;

// displacement that breaks an exact symmetry of the current point (see FitDev::indef)
__device__ void fit_write_kick(const FitDev& s, int64_t toy, int tid, int nth) {
    const int P = s.P;
    double* d = s.Dd + toy * P;
    const double* x = s.X + toy * P;
    const uint8_t* mask = fit_mask(s, toy);
    const unsigned salt = 40503u * (unsigned)(s.nkick[toy] + 1);
    for (int p = tid; p < P; p += nth) {
        const unsigned hsh = ((unsigned)(p + 1) * 2654435761u + salt) >> 8;
        const double u = (double)(hsh & 0xffffu) / 32768.0 - 1.0;
        d[p] = mask[p] ? 0.0 : s.kscale[toy] * fmax(1.0, fabs(x[p])) * u;
    }
    __syncthreads();
    if (tid == 0) {
        s.T[toy] = 1.0;
        s.need_dir[toy] = 0;
        s.pending[toy] = 0;
        s.kick[toy] = 1;
        s.nkick[toy] += 1;
        s.own[toy] = 0;  // the next direction pass asks for the exact Hessian of the displaced point
        s.damp[toy] = 0.0;
    }
}

__global__ void __launch_bounds__(kDirThreads)
fit_direction_kernel(FitDev s, const int* __restrict__ list, int64_t nlist, int allow_refresh) {
    extern __shared__ __align__(16) double sm[];
    const int P = s.P;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t toy = list ? (int64_t)list[blockIdx.x] : (int64_t)blockIdx.x;
    if (toy >= s.NC) return;
    if (s.done[toy] || !s.need_dir[toy]) return;
    if (s.kick[toy] == 2) { fit_write_kick(s, toy, tid, nth); return; }

    // shared layout: b[P] | map[P] (int) | packed A (if it fits)
    double* bvec = sm;
    int* fidx = reinterpret_cast<int*>(bvec + P);
    double* A = s.chol_in_smem ? (bvec + P + (P + 1) / 2 + 1)
                               : (s.chol_scratch + (int64_t)blockIdx.x * ((int64_t)P * (P + 1) / 2));
    __shared__ int s_n;
    __shared__ double s_red[kDirThreads / 32];
    __shared__ int s_flag;

    const double* x = s.X + toy * P;
    const double* g = s.G + toy * P;
    // active set; local parameters go first so that the factorisation can exploit their diagonal
    // block.  Flags are evaluated in parallel, the ordered compaction is a short serial pass over
    // shared memory (P is small).
    __shared__ int s_nloc;
    int* flag = reinterpret_cast<int*>(bvec);  // reuse: bvec is filled after the compaction
    const uint8_t* mask = fit_mask(s, toy);
    for (int p = tid; p < P; p += nth) {
        const double xv = x[p], gv = g[p];
        const bool act = mask[p] || (xv <= s.lo[p] && gv > 0.0) || (xv >= s.hi[p] && gv < 0.0);
        flag[p] = act ? 0 : ((s.use_local && s.islocal[p]) ? 1 : 2);
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int p = 0; p < P; ++p)
            if (flag[p] == 1) fidx[n++] = p;
        s_nloc = n;
        for (int p = 0; p < P; ++p)
            if (flag[p] == 2) fidx[n++] = p;
        s_n = n;
    }
    __syncthreads();
    const int n = s_n, nloc = s_nloc;
    // (the pass after a refresh, allow_refresh != 1, is what makes the fit own its matrix)
    const double* Hsrc = (s.own[toy] || allow_refresh != 1) ? (s.H + toy * (int64_t)P * P) : s.Hshared;
    if (tid == 0 && allow_refresh != 1) s.own[toy] = 1;

    double maxdiag = 0.0;
    for (int i = tid; i < n; i += nth) maxdiag = fmax(maxdiag, fabs(Hsrc[(int64_t)fidx[i] * P + fidx[i]]));
    for (int o = 16; o > 0; o >>= 1) maxdiag = fmax(maxdiag, __shfl_xor_sync(0xffffffffu, maxdiag, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = maxdiag;
    __syncthreads();
    maxdiag = 0.0;
    for (int k = 0; k < nth / 32; ++k) maxdiag = fmax(maxdiag, s_red[k]);
    __syncthreads();

    double tau = s.damp[toy];  // Levenberg-Marquardt damping carried by the fit (0 = pure Newton)
    bool ok = false;
    int attempts = 0;
    for (int attempt = 0; attempt < 40 && !ok; ++attempt) {
        ++attempts;
        // rows of the packed lower triangle; off-diagonal entries of the local block are never
        // read (they are structurally zero) and are not loaded
        for (int i = tid; i < nloc; i += nth) A[tri(i, i)] = Hsrc[(int64_t)fidx[i] * P + fidx[i]] + tau;
        for (int i = nloc; i < n; ++i) {
            const double* hrow = Hsrc + (int64_t)fidx[i] * P;
            double* arow = A + tri(i, 0);
            for (int j = tid; j <= i; j += nth) arow[j] = hrow[fidx[j]] + ((j == i) ? tau : 0.0);
        }
        __syncthreads();
        ok = chol_factor(A, n, nloc, tid, nth);
        if (!ok) tau = (tau < 1e-8 * maxdiag) ? 1e-8 * fmax(maxdiag, 1e-300) : tau * 10.0;
    }
    if (!ok) {
        if (tid == 0) { s.done[toy] = 1; s.status[toy] = 2; }
        return;
    }
    // solve L z = -g_F, then L^T d = z, using the structure: local columns only reach global rows
    for (int i = tid; i < n; i += nth) bvec[i] = -g[fidx[i]];
    __syncthreads();
    for (int i = tid; i < nloc; i += nth) bvec[i] /= A[tri(i, i)];  // z of the local block
    __syncthreads();
    for (int i = nloc + tid; i < n; i += nth) {  // b_glob -= L_gl z_loc
        const double* ri = A + tri(i, 0);
        double acc = bvec[i];
        for (int k = 0; k < nloc; ++k) acc = fma(-ri[k], bvec[k], acc);
        bvec[i] = acc;
    }
    __syncthreads();
    for (int k = nloc; k < n; ++k) {  // dense global block, column oriented
        if (tid == 0) bvec[k] /= A[tri(k, k)];
        __syncthreads();
        const double zk = bvec[k];
        for (int i = k + 1 + tid; i < n; i += nth) bvec[i] = fma(-A[tri(i, k)], zk, bvec[i]);
        __syncthreads();
    }
    for (int k = n - 1; k >= nloc; --k) {  // L_gg^T d_glob = z_glob
        if (tid == 0) bvec[k] /= A[tri(k, k)];
        __syncthreads();
        const double dk = bvec[k];
        const double* rk = A + tri(k, 0);
        for (int i = nloc + tid; i < k; i += nth) bvec[i] = fma(-rk[i], dk, bvec[i]);
        __syncthreads();
    }
    for (int i = tid; i < nloc; i += nth) {  // d_loc = (z_loc - L_gl^T d_glob) / L_ll
        double acc = bvec[i];
        for (int k = nloc; k < n; ++k) acc = fma(-A[tri(k, i)], bvec[k], acc);
        bvec[i] = acc / A[tri(i, i)];
    }
    __syncthreads();
    // edm = -g_F . d_F ; step = max |d|
    double edm = 0.0, step = 0.0;
    for (int i = tid; i < n; i += nth) {
        edm = fma(-g[fidx[i]], bvec[i], edm);
        step = fmax(step, fabs(bvec[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        edm += __shfl_xor_sync(0xffffffffu, edm, o);
        step = fmax(step, __shfl_xor_sync(0xffffffffu, step, o));
    }
    __shared__ double s_red2[kDirThreads / 32];
    if ((tid & 31) == 0) { s_red[tid >> 5] = edm; s_red2[tid >> 5] = step; }
    __syncthreads();
    edm = 0.0; step = 0.0;
    for (int k = 0; k < nth / 32; ++k) { edm += s_red[k]; step = fmax(step, s_red2[k]); }

    if (allow_refresh == 1) {
        const bool slow = edm > s.refresh_ratio * s.EDMPREV[toy];
        const bool forced = s.force[toy] != 0;
        const bool fresh = s.need_dir[toy] == 1;  // 2 = same point, damping raised after a rejection
        // a fit about to meet the stopping rule on a matrix computed at an earlier point gets the exact
        // Hessian of THIS point first: a stale positive definite matrix would hide a saddle
        const bool stale_end = isfinite(edm) && edm < s.tol_edm && step < s.tol_step && !s.hfresh[toy];
        const bool want = stale_end || (fresh && (forced || (s.niter[toy] >= 1 && (!s.own[toy] || slow) && edm > s.tol_edm)));
        if (want) {
            if (tid == 0) {
                // exact Hessian close to the minimum, Fisher matrix (PSD) far from it or after a
                // failed line search
                const int kind = (!forced && edm < s.exact_below) ? 2 : 1;
                s.refresh[toy] = kind;
                s.hfresh[toy] = (kind == 2) ? 1 : 0;
                if (kind == 1) atomicAdd(&s.counters[5], 1);  // Asimov rows needed this round
                s.force[toy] = 0;
                s.EDM[toy] = edm;  // estimate before the refresh (checked after it)
                const int k = atomicAdd(&s.counters[0], 1);
                s.rlist[k] = (int)toy;
            }
            return;  // need_dir stays set: the direction is recomputed after the refresh
        }
    } else if (allow_refresh == 0) {
        // after an EXACT refresh: a nearly singular / indefinite exact Hessian shows up as an
        // exploding step; fall back to the Fisher matrix for this iteration
        if (s.refresh[toy] == 2 && (!(edm <= 100.0 * s.EDM[toy] + 1.0) || step > 100.0)) {
            if (tid == 0) {
                s.refresh[toy] = 1;
                s.hfresh[toy] = 2;   // (Fisher stand-in at this point: the exact matrix was unusable, do not ask again)
                s.pending[toy] = 1;  // sits out this round's trial; served at the top of the next
                const int k = atomicAdd(&s.counters[2], 1);
                s.rlist2[k] = (int)toy;
            }
            return;
        }
    }
    double* d = s.Dd + toy * P;
    for (int p = tid; p < P; p += nth) d[p] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += nth) d[fidx[i]] = bvec[i];
    const bool converged = isfinite(edm) && edm < s.tol_edm && step < s.tol_step;
    if (converged && s.damp[toy] > 0.0) {
        // the stopping rule is met under a Levenberg-Marquardt shift, which would hide an
        // indefinite Hessian (a saddle reached along an exact symmetry): look again at the same
        // point without damping before the fit is allowed to end
        __syncthreads();
        if (tid == 0) { s.damp[toy] = 0.0; s.need_dir[toy] = 2; s.pending[toy] = 1; }
        return;
    }
    // creeping along a ridge: nearly stationary, Hessian not positive definite, several times in a row
    const int ncreep = (attempts > 1 && isfinite(edm) && edm < 1e-3) ? s.nindef[toy] + 1 : 0;
    const bool saddle = ((converged && attempts > 1) || ncreep >= 4) && s.nkick[toy] < kMaxKicks;
    __syncthreads();
    if (tid == 0) {
        s.nindef[toy] = saddle ? 0 : ncreep;
        s.EDM[toy] = edm;
        s.EDMPREV[toy] = edm;
        s.T[toy] = 1.0;
        s.need_dir[toy] = 0;
        s.pending[toy] = 0;
        s.indef[toy] = attempts > 1 ? 1 : 0;
        if (!isfinite(edm)) { s.done[toy] = 1; s.status[toy] = 2; }
        else if (converged && !saddle) { s.done[toy] = 1; s.status[toy] = 0; }
    }
    if (saddle) { __syncthreads(); fit_write_kick(s, toy, tid, nth); }
}

// ---------------------------------------------------------------- Newton direction, Schur form
// Same contract as fit_direction_kernel, for models whose local (per-bin) parameters never share a
// bin (use_local): the local block of the Hessian is diagonal, so the system is reduced to the
// global parameters,
//     S = H_gg + tau I - H_gl (H_ll + tau I)^-1 H_lg ,   S d_g = -g_g + H_gl (H_ll + tau I)^-1 g_l ,
//     d_l = -(H_ll + tau I)^-1 (g_l + H_lg d_g) ,
// and only S (n_glob x n_glob, packed) ever lives in shared memory: ~2 KB for C3 (20 globals),
// 160 KB for C4 (200).  H_gl is streamed through a [ltile][n_glob] staging tile.
__device__ __forceinline__ void tri_decode(int e, int& a, int& b) {
    a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > e) --a;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    b = e - a * (a + 1) / 2;
}

__global__ void fit_direction_schur_kernel(FitDev s, const int* __restrict__ list, int64_t nlist,
                                           int allow_refresh, int ngmax, int nlmax, int ltile,
                                           int backup) {
    extern __shared__ __align__(16) double sm[];
    const int P = s.P;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
    const int64_t toy = (int64_t)list[blockIdx.x];
    if (toy >= s.NC) return;
    if (s.done[toy] || !s.need_dir[toy]) return;
    if (s.kick[toy] == 2) { fit_write_kick(s, toy, tid, nth); return; }

    const int ngp = ngmax | 1;
    double* S = sm;                                  // packed lower triangle of the reduced matrix
    double* r = S + ngmax * (ngmax + 1) / 2;         // rhs of the global block, then its solution
    double* z = r + ngmax;
    double* linv = z + ngmax;                        // 1 / L_kk
    double* dinv = linv + ngmax;                     // 1 / (H_ll + tau)
    double* gl = dinv + nlmax;                       // g_l / (H_ll + tau), then d_l
    double* W = gl + nlmax;                          // [ltile][ngp] staged H_gl
    int* glb = reinterpret_cast<int*>(W + (size_t)ltile * ngp);
    int* loc = glb + ngmax;
    // copy of the reduced matrix before factorisation (when it fits): a retry with a larger shift
    // then costs one small Cholesky instead of the whole reduction
    double* S2 = reinterpret_cast<double*>(loc + ((nlmax + ngmax + 1) & ~1) - ngmax);
    __shared__ int s_ng, s_nl, s_bad;
    __shared__ double s_red[32], s_red2[32];

    const double* x = s.X + toy * P;
    const double* g = s.G + toy * P;
    // (the pass after a refresh, allow_refresh != 1, is what makes the fit own its matrix)
    const double* Hsrc = (s.own[toy] || allow_refresh != 1) ? (s.H + toy * (int64_t)P * P) : s.Hshared;
    if (tid == 0 && allow_refresh != 1) s.own[toy] = 1;

    // active set (box bounds, fixed parameters), compacted by one warp with ballots
    const uint8_t* mask = fit_mask(s, toy);
    if (tid < 32) {
        int nl_ = 0, ng_ = 0;
#pragma unroll 4
        for (int p0 = 0; p0 < P; p0 += 32) {
            const int p = p0 + lane;
            int f = 0;
            if (p < P) {
                const double xv = x[p], gv = g[p];
                const bool act = mask[p] || (xv <= s.lo[p] && gv > 0.0) || (xv >= s.hi[p] && gv < 0.0);
                f = act ? 0 : (s.islocal[p] ? 1 : 2);
            }
            const unsigned ml = __ballot_sync(0xffffffffu, f == 1), mg = __ballot_sync(0xffffffffu, f == 2);
            const unsigned lt = (1u << lane) - 1u;
            if (f == 1) loc[nl_ + __popc(ml & lt)] = p;
            if (f == 2) glb[ng_ + __popc(mg & lt)] = p;
            nl_ += __popc(ml);
            ng_ += __popc(mg);
        }
        if (lane == 0) { s_nl = nl_; s_ng = ng_; s_bad = 0; }
    }
    __syncthreads();
    const int ng = s_ng, nl = s_nl, npk = ng * (ng + 1) / 2;
    const int TW = (nth >= 256) ? 16 : 8, TH = nth / TW;
    const int tx = tid % TW, ty = tid / TW;

    double tau = s.damp[toy];  // Levenberg-Marquardt damping carried by the fit (0 = pure Newton)
    double maxdiag = -1.0;
    bool ok = false;
    int attempts = 0;
    bool have_copy = false;
    double tau_base = tau;  // shift inside the stored copy (and on the local block)
    for (int attempt = 0; attempt < 40 && !ok; ++attempt) {
        ++attempts;
        bool bad = false;
        if (have_copy) {
            // only the global block takes the extra shift: still a positive definite modification
            for (int e = tid; e < npk; e += nth) S[e] = S2[e];
            __syncthreads();
            for (int a = tid; a < ng; a += nth) S[tri(a, a)] += tau - tau_base;
            __syncthreads();
        } else {
        tau_base = tau;
        for (int i = tid; i < nl; i += nth) {
            const int l = loc[i];
            const double h = Hsrc[(int64_t)l * P + l] + tau;
            if (!(h > 0.0) || !isfinite(h)) s_bad = 1;
            const double di = 1.0 / h;
            dinv[i] = di;
            gl[i] = g[l] * di;
        }
        for (int e = tid; e < npk; e += nth) {
            int a, b;
            tri_decode(e, a, b);
            S[e] = Hsrc[(int64_t)glb[a] * P + glb[b]] + ((a == b) ? tau : 0.0);
        }
        for (int a = tid; a < ng; a += nth) r[a] = -g[glb[a]];
        __syncthreads();
        bad = s_bad != 0;
        for (int l0 = 0; l0 < nl && !bad; l0 += ltile) {
            const int nt = min(ltile, nl - l0);
#pragma unroll 4
            for (int e = tid; e < ng * nt; e += nth) {
                const int a = e / nt, t = e - a * nt;
                W[t * ngp + a] = Hsrc[(int64_t)glb[a] * P + loc[l0 + t]];
            }
            __syncthreads();
            for (int e = tid; e < npk; e += nth) {
                int a, b;
                tri_decode(e, a, b);
                double acc = 0.0;
                for (int t = 0; t < nt; ++t) acc = fma(W[t * ngp + a] * dinv[l0 + t], W[t * ngp + b], acc);
                S[e] -= acc;
            }
            for (int a = tid; a < ng; a += nth) {
                double acc = 0.0;
                for (int t = 0; t < nt; ++t) acc = fma(W[t * ngp + a], gl[l0 + t], acc);
                r[a] += acc;
            }
            __syncthreads();
        }
        if (backup && !bad) {
            for (int e = tid; e < npk; e += nth) S2[e] = S[e];
            have_copy = true;
        }
        }
        // right-looking Cholesky of S; the diagonal keeps the pivots, linv holds 1/L_kk
        for (int j = 0; j < ng && !bad; ++j) {
            const double piv = S[tri(j, j)];
            if (!(piv > 0.0) || !isfinite(piv)) { bad = true; break; }
            const double inv = 1.0 / sqrt(piv);
            for (int i = j + 1 + tid; i < ng; i += nth) S[tri(i, j)] *= inv;
            if (tid == 0) linv[j] = inv;
            __syncthreads();
            for (int i = j + 1 + ty; i < ng; i += TH) {
                const double lij = S[tri(i, j)];
                for (int k = j + 1 + tx; k <= i; k += TW) S[tri(i, k)] = fma(-lij, S[tri(k, j)], S[tri(i, k)]);
            }
            __syncthreads();
        }
        ok = !bad;
        if (!ok) {
            if (maxdiag < 0.0) {
                double m = 0.0;
                for (int i = tid; i < nl; i += nth) m = fmax(m, fabs(Hsrc[(int64_t)loc[i] * P + loc[i]]));
                for (int a = tid; a < ng; a += nth) m = fmax(m, fabs(Hsrc[(int64_t)glb[a] * P + glb[a]]));
                for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) s_red[tid >> 5] = m;
                __syncthreads();
                m = 0.0;
                for (int k = 0; k < (nth + 31) / 32; ++k) m = fmax(m, s_red[k]);
                maxdiag = m;
            }
            __syncthreads();
            if (tid == 0) s_bad = 0;
            tau = (tau < 1e-8 * maxdiag) ? 1e-8 * fmax(maxdiag, 1e-300) : tau * 10.0;
            __syncthreads();
        }
    }
    if (!ok) {
        if (tid == 0) { s.done[toy] = 1; s.status[toy] = 2; }
        return;
    }
    if (tid == 0 && attempts > 1) atomicAdd(&s.counters[4], attempts - 1);
    // L z = r
    for (int k = 0; k < ng; ++k) {
        const double zk = r[k] * linv[k];
        for (int i = k + 1 + tid; i < ng; i += nth) r[i] = fma(-S[tri(i, k)], zk, r[i]);
        if (tid == 0) z[k] = zk;
        __syncthreads();
    }
    // L^T d_g = z   (d_g is left in r)
    for (int k = ng - 1; k >= 0; --k) {
        const double dk = z[k] * linv[k];
        for (int i = tid; i < k; i += nth) z[i] = fma(-S[tri(k, i)], dk, z[i]);
        if (tid == 0) r[k] = dk;
        __syncthreads();
    }
    // d_l = -(g_l + H_lg d_g) / (H_ll + tau)   (left in gl)
    {
        // nsub threads share one local parameter (interleaved over a) when threads outnumber locals
        int nsub = 1;
        while (nsub < 8 && nl * nsub * 2 <= nth) nsub *= 2;
        const int i = tid / nsub, sub = tid - (tid / nsub) * nsub;
        for (int i0 = 0; i0 < nl; i0 += nth / nsub) {
            const int ii = i0 + i;
            double acc = 0.0;
            if (ii < nl) {
                const int l = loc[ii];
#pragma unroll 8
                for (int a = sub; a < ng; a += nsub) acc = fma(Hsrc[(int64_t)glb[a] * P + l], r[a], acc);
            }
            for (int o = nsub >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (ii < nl && sub == 0) gl[ii] = -gl[ii] - dinv[ii] * acc;
        }
    }
    __syncthreads();
    // edm = -g_F . d_F ; step = max |d|
    double edm = 0.0, step = 0.0;
    for (int i = tid; i < nl; i += nth) {
        edm = fma(-g[loc[i]], gl[i], edm);
        step = fmax(step, fabs(gl[i]));
    }
    for (int a = tid; a < ng; a += nth) {
        edm = fma(-g[glb[a]], r[a], edm);
        step = fmax(step, fabs(r[a]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        edm += __shfl_xor_sync(0xffffffffu, edm, o);
        step = fmax(step, __shfl_xor_sync(0xffffffffu, step, o));
    }
    if (lane == 0) { s_red[tid >> 5] = edm; s_red2[tid >> 5] = step; }
    __syncthreads();
    edm = 0.0; step = 0.0;
    for (int k = 0; k < (nth + 31) / 32; ++k) { edm += s_red[k]; step = fmax(step, s_red2[k]); }

    if (allow_refresh == 1) {
        const bool slow = edm > s.refresh_ratio * s.EDMPREV[toy];
        const bool forced = s.force[toy] != 0;
        const bool fresh = s.need_dir[toy] == 1;  // 2 = same point, damping raised after a rejection
        // a fit about to meet the stopping rule on a matrix computed at an earlier point gets the exact
        // Hessian of THIS point first: a stale positive definite matrix would hide a saddle
        const bool stale_end = isfinite(edm) && edm < s.tol_edm && step < s.tol_step && !s.hfresh[toy];
        const bool want = stale_end || (fresh && (forced || (s.niter[toy] >= 1 && (!s.own[toy] || slow) && edm > s.tol_edm)));
        if (want) {
            if (tid == 0) {
                const int kind = (!forced && edm < s.exact_below) ? 2 : 1;
                s.refresh[toy] = kind;
                s.hfresh[toy] = (kind == 2) ? 1 : 0;
                if (kind == 1) atomicAdd(&s.counters[5], 1);  // Asimov rows needed this round
                s.force[toy] = 0;
                s.EDM[toy] = edm;
                const int k = atomicAdd(&s.counters[0], 1);
                s.rlist[k] = (int)toy;
            }
            return;
        }
    } else if (allow_refresh == 0) {
        if (s.refresh[toy] == 2 && (!(edm <= 100.0 * s.EDM[toy] + 1.0) || step > 100.0)) {
            if (tid == 0) {
                s.refresh[toy] = 1;
                s.hfresh[toy] = 2;   // (Fisher stand-in at this point: the exact matrix was unusable, do not ask again)
                s.pending[toy] = 1;  // sits out this round's trial; served at the top of the next
                const int k = atomicAdd(&s.counters[2], 1);
                s.rlist2[k] = (int)toy;
            }
            return;
        }
    }
    double* d = s.Dd + toy * P;
    for (int p = tid; p < P; p += nth) d[p] = 0.0;
    __syncthreads();
    for (int i = tid; i < nl; i += nth) d[loc[i]] = gl[i];
    for (int a = tid; a < ng; a += nth) d[glb[a]] = r[a];
    const bool converged = isfinite(edm) && edm < s.tol_edm && step < s.tol_step;
    if (converged && s.damp[toy] > 0.0) {
        // the stopping rule is met under a Levenberg-Marquardt shift, which would hide an
        // indefinite Hessian (a saddle reached along an exact symmetry): look again at the same
        // point without damping before the fit is allowed to end
        __syncthreads();
        if (tid == 0) { s.damp[toy] = 0.0; s.need_dir[toy] = 2; s.pending[toy] = 1; }
        return;
    }
    // creeping along a ridge: nearly stationary, Hessian not positive definite, several times in a row
    const int ncreep = (attempts > 1 && isfinite(edm) && edm < 1e-3) ? s.nindef[toy] + 1 : 0;
    const bool saddle = ((converged && attempts > 1) || ncreep >= 4) && s.nkick[toy] < kMaxKicks;
    __syncthreads();
    if (tid == 0) {
        s.nindef[toy] = saddle ? 0 : ncreep;
        s.EDM[toy] = edm;
        s.EDMPREV[toy] = edm;
        s.T[toy] = 1.0;
        s.need_dir[toy] = 0;
        s.pending[toy] = 0;
        s.indef[toy] = attempts > 1 ? 1 : 0;
        if (!isfinite(edm)) { s.done[toy] = 1; s.status[toy] = 2; }
        else if (converged && !saddle) { s.done[toy] = 1; s.status[toy] = 0; }
    }
    if (saddle) { __syncthreads(); fit_write_kick(s, toy, tid, nth); }
}

// ---------------------------------------------------------------- trial point / acceptance
// compact trial points of the live fits: XTc[k] = clip(X + T*D) of fit live[k]
__global__ void fit_trial_kernel(FitDev s, const int* __restrict__ live, int64_t nl,
                                 double* __restrict__ XTc, int* __restrict__ ridx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nl * s.P) return;
    const int64_t k = idx / s.P;
    const int p = (int)(idx - k * s.P);
    const int64_t toy = live[k];
    double v = s.X[toy * s.P + p];
    if (!s.done[toy] && !s.pending[toy] && !fit_mask(s, toy)[p]) v = fmin(fmax(v + s.T[toy] * s.Dd[toy * s.P + p], s.lo[p]), s.hi[p]);
    XTc[idx] = v;
    if (p == 0) ridx[k] = (int)fit_data_row(s, toy);
}

// one warp per live fit; fits that stay alive are appended to live_next
__global__ void fit_accept_kernel(FitDev s, const int* __restrict__ live, int64_t nl,
                                  const double* __restrict__ XTc, const double* __restrict__ VTc,
                                  const double* __restrict__ GTc, int* __restrict__ live_next) {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= nl) return;
    const int64_t toy = live[k];
    if (s.done[toy]) return;
    if (s.pending[toy]) {  // no trial this round: stays alive, nothing else changes
        if (lane == 0) live_next[atomicAdd(&s.counters[1], 1)] = (int)toy;
        return;
    }
    const int P = s.P;
    const double* x = s.X + toy * P;
    const double* xt = XTc + k * P;
    const double* g = s.G + toy * P;
    double dec = 0.0;
    for (int p = lane; p < P; p += 32) dec = fma(g[p], xt[p] - x[p], dec);
    for (int o = 16; o > 0; o >>= 1) dec += __shfl_xor_sync(0xffffffffu, dec, o);
    const double f = s.F[toy];
    const double ft = -2.0 * VTc[k];
    const double slack = 4e-15 * fmax(1.0, fabs(f));
    const bool kicked = s.kick[toy] == 1;  // displaced off a saddle: taken unconditionally
    // a kick is taken unconditionally as long as it stays inside the domain and costs less than one unit
    // of 2NLL; else it is tried again at a tenth of the size (narrow valleys: curvatures of 1e7 and a
    // clipped rate next to the saddle in the reference's clipping model)
    const bool ok = isfinite(ft) && (kicked ? ft <= f + 1.0 : ft <= f + kArmijo * dec + slack);
    bool alive = true;
    if (kicked && !ok) {
        if (lane == 0) {
            if (s.kscale[toy] > 1e-9) {
                s.kscale[toy] *= 0.1;
                s.kick[toy] = 2;
                s.need_dir[toy] = 1;
                s.nkick[toy] -= 1;
                live_next[atomicAdd(&s.counters[1], 1)] = (int)toy;
            } else {  // stay where the fit was
                s.done[toy] = 1; s.status[toy] = 0; s.kick[toy] = 0;
            }
        }
        return;
    }
    if (s.trace && (nl <= 3 || toy < 2) && lane == 0)
        printf("[hf_fit trace] fit %lld it %d f %.12g ft-f %.3e dec %.3e edm %.3e damp %.3e own %d refresh %d nfail %d indef %d nkick %d %s\n",
               (long long)toy, s.niter[toy], f, ft - f, dec, s.EDM[toy], s.damp[toy], s.own[toy], s.refresh[toy],
               s.nfail[toy], s.indef[toy], s.nkick[toy], ok ? "accept" : "reject");
    if (ok) {
        if (lane == 0) atomicAdd(&s.counters[3], 1);
        for (int p = lane; p < P; p += 32) {
            s.X[toy * P + p] = xt[p];
            s.G[toy * P + p] = -2.0 * GTc[k * P + p];
        }
        if (lane == 0) {
            s.F[toy] = ft;
            s.need_dir[toy] = 1;
            s.nfail[toy] = 0;
            s.kick[toy] = 0;
            s.chk[toy] = 0;
            s.hfresh[toy] = 0;
            const double dmp = 0.2 * s.damp[toy];
            s.damp[toy] = (dmp < 1e-6) ? 0.0 : dmp;
            const int it = s.niter[toy] + 1;
            s.niter[toy] = it;
            if (it >= s.max_iter) { s.done[toy] = 1; s.status[toy] = 1; alive = false; }
        }
    } else if (lane == 0) {
        // rejected: raise the damping and recompute the direction at the same point (trust-region
        // style) instead of shrinking the step along a direction a nearly singular Hessian made bad
        const int nr = s.nfail[toy] + 1;
        if (s.EDM[toy] < 1e-7 && s.damp[toy] > 0.0) {
            // (same test once more without the damping: it would hide an indefinite Hessian)
            s.damp[toy] = 0.0;
            s.need_dir[toy] = 2;
            s.chk[toy] = 1;
        } else if (s.EDM[toy] < 1e-7 || s.chk[toy]) {
            // predicted decrease below the rounding floor of f -- or, at a kink, an undamped step that
            // cannot be realised where the damped one had nothing left to gain: converged
            if (s.indef[toy] && s.nkick[toy] < kMaxKicks) {
                s.kick[toy] = 2;  // ... on a saddle: the next direction pass displaces the fit
                s.need_dir[toy] = 1;
            } else {
                s.done[toy] = 1; s.status[toy] = 0; alive = false;
            }
        } else if (nr > 60) {
            s.done[toy] = 1; s.status[toy] = 2; alive = false;
        } else {
            const double dmp = s.damp[toy];
            s.damp[toy] = fmax(8.0 * dmp, 1e-2);
            s.nfail[toy] = nr;
            s.need_dir[toy] = 2;
        }
    }
    if (lane == 0 && alive) live_next[atomicAdd(&s.counters[1], 1)] = (int)toy;
}

// fits whose start point differs from fit 0's (their start-up curvature cannot be the shared matrix)
__global__ void fit_ownstart_list_kernel(FitDev s, int* __restrict__ list, int* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.NC || !s.ownstart[i] || s.done[i]) return;
    list[atomicAdd(count, 1)] = (int)i;
    s.own[i] = 1;
}

__global__ void fit_iota_kernel(int* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)i;
}

__global__ void fit_finish_kernel(FitDev s, int64_t toy0, double* __restrict__ x,
                                  double* __restrict__ fun, int32_t* __restrict__ status,
                                  int32_t* __restrict__ niter) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.NC * s.P) return;
    const int64_t i = idx / s.P;
    x[(toy0 + i) * s.P + (idx - i * s.P)] = s.X[idx];
    if (idx - i * s.P == 0) {
        fun[toy0 + i] = s.F[i];
        status[toy0 + i] = s.status[i];
        if (niter) niter[toy0 + i] = s.niter[i];
    }
}

// ---------------------------------------------------------------- host side
struct Carver {
    char* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) / 256 * 256;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += sizeof(T) * n;
        return p;
    }
};

void build_fd_structure(hf_plan* plan) {
    if (plan->fd_ready) return;
    const int P = plan->P, NB = plan->NB, S = plan->S, K = plan->K;
    // a parameter is "local" when it only ever acts as a per-bin factor, on exactly one bin
    std::vector<int> bin_of(P, -1);
    std::vector<char> global(P, 0), seen(P, 0);
    for (int p : plan->h_hs_par) global[p] = 1;
    for (int p : plan->h_sf_par) global[p] = 1;
    for (int k = 0; k < K; ++k)
        for (int s = 0; s < S; ++s)
            for (int b = 0; b < NB; ++b) {
                const int p = plan->h_bf_par[((size_t)k * S + s) * NB + b];
                if (p < 0) continue;
                seen[p] = 1;
                if (bin_of[p] == -1) bin_of[p] = b;
                else if (bin_of[p] != b) global[p] = 1;
            }
    for (int p = 0; p < P; ++p)
        if (!seen[p]) global[p] = 1;  // parameters without any effect still need a column
    // colour = rank among the local parameters of the same bin
    std::vector<std::vector<int>> per_bin(NB);
    for (int p = 0; p < P; ++p)
        if (!global[p]) per_bin[bin_of[p]].push_back(p);
    int ncolors = 0;
    for (auto& v : per_bin) ncolors = std::max<int>(ncolors, (int)v.size());
    plan->fd_ncolors = ncolors;
    plan->fd_color.assign(P, -1);
    plan->fd_partner.assign((size_t)P * std::max(ncolors, 1), -1);
    for (auto& v : per_bin)
        for (size_t a = 0; a < v.size(); ++a) {
            plan->fd_color[v[a]] = (int)a;
            for (size_t c = 0; c < v.size(); ++c) plan->fd_partner[(size_t)v[a] * ncolors + c] = v[c];
        }
    plan->fd_ready = true;
}

}  // namespace

int hf_fit_launch(hf_plan* plan, const double* data, int64_t data_rows, const double* init,
                  int64_t init_rows, const double* lo, const double* hi, const uint8_t* fixed,
                  int64_t N, const hf_fit_opts* opts_in, double* x_out, double* fun_out,
                  int32_t* status_out, int32_t* niter_out, cudaStream_t st,
                  const uint8_t* fixed_alt, const uint8_t* use_alt, const int32_t* data_map, int n_alt) {
    HfRange nvtx_range("K3 hf_fit_batch");
    if (N <= 0) return HF_OK;
    if (n_alt < 1 || n_alt > 255) {
        hf_set_error("the mask table holds 1..255 alternative masks (got %d)", n_alt);
        return HF_ERR_INVALID;
    }
    if ((!data_map && data_rows != 1 && data_rows != N) || data_rows < 1 ||
        (init_rows != 1 && init_rows != N)) {
        hf_set_error("data_rows/init_rows must be 1 or N (data_rows is free when a data map is given)");
        return HF_ERR_INVALID;
    }
    if ((fixed_alt == nullptr) != (use_alt == nullptr)) {
        hf_set_error("fixed_alt and use_alt must be given together");
        return HF_ERR_INVALID;
    }
    hf_fit_opts opts;
    hf_fit_default_opts(&opts);
    if (opts_in) opts = *opts_in;
    const int P = plan->P, D = plan->D;
    build_fd_structure(plan);

    plan->solo_last_nt = 0;  // (set again by the per-fit kernel when it takes the batch)
    // plans inside the scope of the per-fit kernel run there: one CTA per fit, no host in the loop
    // (hf_fit_solo.cu); everything else takes the lock-step driver below
    if (hf_fit_solo_supported(plan)) {
        double* DC = nullptr;
        const size_t need_dc = sizeof(double) * (size_t)data_rows + 256;
        if (need_dc > plan->fit_bytes) {
            if (plan->fit_block) cudaFree(plan->fit_block);
            plan->fit_block = nullptr;
            plan->fit_bytes = 0;
            cudaError_t e = cudaMalloc(&plan->fit_block, need_dc);
            if (e != cudaSuccess) {
                hf_set_error("cudaMalloc(%zu) for the fit workspace: %s", need_dc, cudaGetErrorString(e));
                return HF_ERR_NOMEM;
            }
            plan->fit_bytes = need_dc;
        }
        DC = reinterpret_cast<double*>(plan->fit_block);
        int rc0 = hf_data_const_launch(plan, data, data_rows, DC, st);
        if (rc0) return rc0;
        int handled = 0;
        rc0 = hf_fit_solo_launch(plan, data, data_rows, DC, init, init_rows, lo, hi, fixed, fixed_alt, use_alt, n_alt,
                                 data_map, N, opts, x_out, fun_out, status_out, niter_out, st, &handled);
        if (rc0) return rc0;
        if (handled) return HF_OK;
    }
    // ---- column structure for this fixed mask (host: P is small)
    // (a parameter gets no FD column only when it is fixed under every mask of the batch)
    std::vector<uint8_t> h_fixed(P), h_alt;
    HF_CUDA_OK(cudaMemcpyAsync(h_fixed.data(), fixed, P, cudaMemcpyDeviceToHost, st));
    if (fixed_alt) {
        h_alt.resize((size_t)n_alt * P);
        HF_CUDA_OK(cudaMemcpyAsync(h_alt.data(), fixed_alt, (size_t)n_alt * P, cudaMemcpyDeviceToHost, st));
    }
    HF_CUDA_OK(cudaStreamSynchronize(st));
    if (fixed_alt)
        for (int k = 0; k < n_alt; ++k)
            for (int p = 0; p < P; ++p) h_fixed[p] = (h_fixed[p] && h_alt[(size_t)k * P + p]) ? 1 : 0;
    std::vector<int32_t> colpar, pcolor(P), pcol(P, -1);
    std::vector<char> color_used(std::max(plan->fd_ncolors, 1), 0);
    for (int p = 0; p < P; ++p) {
        if (h_fixed[p]) { pcolor[p] = -2; continue; }
        pcolor[p] = plan->fd_color[p];
        if (pcolor[p] < 0) { pcol[p] = (int)colpar.size(); colpar.push_back(p); }
        else color_used[pcolor[p]] = 1;
    }
    const int nglob = (int)colpar.size();
    const int ncolors = plan->fd_ncolors;
    const int ncol = nglob + ncolors;
    int nfree = 0;
    for (int p = 0; p < P; ++p) nfree += h_fixed[p] ? 0 : 1;

    // curvature: closed form (hf_hess_analytic.cu) when the plan qualifies, else forward differences
    // of the analytic gradient
    bool use_ah = plan->d.ah_ok != 0;
    if (const char* e = getenv("HF_FIT_HESS")) use_ah = use_ah && e[0] != 'f';  // "fd": tuning / test knob
    const size_t ah_stride = hf_hess_analytic_scratch_doubles(plan);
    // ---- chunking: per-fit Hessians dominate the memory
    size_t free_b = 0, total_b = 0;
    HF_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_fit = sizeof(double) * ((size_t)P * P + 6 * (size_t)P + 8) + 64;
    const size_t budget = std::min<size_t>(free_b / 3, (size_t)48 << 30);
    int64_t NC = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)(budget / per_fit)));
    // FD batch: points per K2 launch
    const int64_t max_pts = use_ah ? (int64_t)(ncol + 1)
                                   : std::max<int64_t>((int64_t)(ncol + 1), std::min<int64_t>((int64_t)1 << 18,
                                (int64_t)(budget / (sizeof(double) * (2 * (size_t)P + 2) + 8))));
    const int64_t ah_cap = std::max<int64_t>(1, std::min<int64_t>(NC, (int64_t)(((size_t)2 << 30) / (sizeof(double) * ah_stride))));
    const int64_t rl_cap = std::max<int64_t>(1, max_pts / (ncol + 1));  // refresh entries per sub-batch

    const size_t tri_bytes = sizeof(double) * (size_t)P * (P + 1) / 2;
    const size_t dir_smem_small = sizeof(double) * (P + (P + 1) / 2 + 2);
    const bool chol_in_smem = dir_smem_small + tri_bytes <= 200 * 1024;
    const int64_t dir_grid_cap = chol_in_smem ? 0 : std::min<int64_t>(NC, 4096);

    Carver cv{nullptr};
    double* AHS = nullptr;  // scratch of the analytic-Hessian kernel, per request
    auto carve = [&](Carver& c, FitDev& s, double*& XP, double*& GP, double*& VP, double*& DR,
                     double*& AS, double*& XG, int*& ridx, int*& d_colpar, int*& d_pcolor,
                     int*& d_partner, int*& d_pcol, int*& d_islocal, uint8_t*& d_struct, double*& DC) {
        s.X = c.take<double>(NC * P); s.G = c.take<double>(NC * P); s.Dd = c.take<double>(NC * P);
        s.XT = c.take<double>(NC * P); s.GT = c.take<double>(NC * P); s.VT = c.take<double>(NC);
        s.F = c.take<double>(NC); s.T = c.take<double>(NC); s.EDM = c.take<double>(NC);
        s.EDMPREV = c.take<double>(NC);
        s.H = c.take<double>((size_t)NC * P * P); s.Hshared = c.take<double>((size_t)P * P);
        s.done = c.take<int>(NC); s.need_dir = c.take<int>(NC); s.own = c.take<int>(NC);
        s.status = c.take<int>(NC); s.niter = c.take<int>(NC); s.refresh = c.take<int>(NC);
        s.rlist = c.take<int>(NC + 1); s.counters = c.take<int>(8);
        s.live = c.take<int>(NC + 1); s.live_next = c.take<int>(NC + 1);
        s.nfail = c.take<int>(NC); s.force = c.take<int>(NC); s.rlist2 = c.take<int>(NC + 1);
        s.pending = c.take<int>(NC);
        s.ownstart = c.take<int>(NC);
        s.indef = c.take<int>(NC); s.nkick = c.take<int>(NC); s.kick = c.take<int>(NC);
        s.nindef = c.take<int>(NC);
        s.chk = c.take<int>(NC);
        s.hfresh = c.take<int>(NC);
        s.kscale = c.take<double>(NC);
        s.damp = c.take<double>(NC);
        AHS = use_ah ? c.take<double>((size_t)ah_cap * ah_stride) : nullptr;
        d_islocal = c.take<int>(P);
        d_struct = c.take<uint8_t>(P);
        DC = c.take<double>(data_rows);
        XP = c.take<double>((size_t)max_pts * P); GP = c.take<double>((size_t)max_pts * P);
        VP = c.take<double>(max_pts); DR = c.take<double>((size_t)rl_cap * D);
        AS = c.take<double>((size_t)rl_cap * D);
        XG = c.take<double>((size_t)rl_cap * P); ridx = c.take<int>(std::max<int64_t>(max_pts, NC) + 1);
        d_colpar = c.take<int>(std::max(nglob, 1)); d_pcolor = c.take<int>(P);
        d_partner = c.take<int>((size_t)P * std::max(ncolors, 1)); d_pcol = c.take<int>(P);
        s.chol_scratch = chol_in_smem ? nullptr : c.take<double>((size_t)dir_grid_cap * P * (P + 1) / 2);
    };
    FitDev s;
    memset(&s, 0, sizeof(s));
    double *XP, *GP, *VP, *DR, *AS, *XG;
    int *ridx, *d_colpar, *d_pcolor, *d_partner, *d_pcol, *d_islocal;
    uint8_t* d_struct;
    double* DC;
    carve(cv, s, XP, GP, VP, DR, AS, XG, ridx, d_colpar, d_pcolor, d_partner, d_pcol, d_islocal, d_struct, DC);
    const size_t need = cv.off + 4096;
    if (need > plan->fit_bytes) {
        if (plan->fit_block) cudaFree(plan->fit_block);
        plan->fit_block = nullptr;
        plan->fit_bytes = 0;
        cudaError_t e = cudaMalloc(&plan->fit_block, need);
        if (e != cudaSuccess) {
            hf_set_error("cudaMalloc(%zu) for the fit workspace: %s", need, cudaGetErrorString(e));
            return HF_ERR_NOMEM;
        }
        plan->fit_bytes = need;
    }
    Carver cv2{(char*)plan->fit_block};
    carve(cv2, s, XP, GP, VP, DR, AS, XG, ridx, d_colpar, d_pcolor, d_partner, d_pcol, d_islocal, d_struct, DC);
    // theta-independent terms of every data row, once for the whole fit (every round evaluates
    // trial points against this same array)
    struct HeldConst {
        hf_plan* p;
        ~HeldConst() { p->dconst_held = nullptr; p->dconst_held_data = nullptr; p->dconst_held_rows = 0; }
    } held{plan};
    {
        int rc0 = hf_data_const_launch(plan, data, data_rows, DC, st);
        if (rc0) return rc0;
        plan->dconst_held = DC; plan->dconst_held_data = data; plan->dconst_held_rows = data_rows;
    }
    if (!plan->h_counters) HF_CUDA_OK(cudaHostAlloc((void**)&plan->h_counters, sizeof(int32_t) * 16, cudaHostAllocDefault));
    s.P = P; s.D = D; s.ncol = ncol; s.nglob = nglob; s.ncolors = std::max(ncolors, 1);
    s.colpar = d_colpar; s.pcolor = d_pcolor; s.partner = d_partner; s.pcol = d_pcol;
    s.lo = lo; s.hi = hi; s.fixed = fixed;
    s.fixed_alt = fixed_alt; s.use_alt = use_alt; s.fixed_struct = d_struct; s.data_map = data_map;
    s.n_alt = fixed_alt ? n_alt : 0;
    s.data_rows = data_rows; s.toy0 = 0;
    HF_CUDA_OK(cudaMemcpyAsync(d_struct, h_fixed.data(), P, cudaMemcpyHostToDevice, st));
    s.chol_in_smem = chol_in_smem ? 1 : 0;
    s.islocal = d_islocal;
    s.use_local = (plan->fd_ncolors == 1) ? 1 : 0;  // locals never share a bin: their block is diagonal
    {
        std::vector<int32_t> isl(P, 0);
        for (int p = 0; p < P; ++p) isl[p] = (plan->fd_color[p] >= 0) ? 1 : 0;
        HF_CUDA_OK(cudaMemcpyAsync(d_islocal, isl.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
        HF_CUDA_OK(cudaStreamSynchronize(st));  // isl goes out of scope
    }
    s.tol_edm = opts.tol_edm; s.tol_step = opts.tol_step; s.refresh_ratio = 0.05; s.exact_below = 10.0;
    if (const char* e = getenv("HF_FIT_REFRESH")) s.refresh_ratio = atof(e);    // tuning knobs (tools/)
    if (const char* e = getenv("HF_FIT_EXACT_BELOW")) s.exact_below = atof(e);
    s.damp0 = 0.0;
    if (const char* e = getenv("HF_FIT_DAMP0")) s.damp0 = atof(e);
    s.max_iter = opts.max_iter;
    s.trace = opts.verbose >= 3 ? 1 : 0;
    HF_CUDA_OK(cudaMemcpyAsync(d_colpar, colpar.data(), sizeof(int) * nglob, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_pcolor, pcolor.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_pcol, pcol.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
    if (ncolors > 0)
        HF_CUDA_OK(cudaMemcpyAsync(d_partner, plan->fd_partner.data(), sizeof(int) * (size_t)P * ncolors,
                                   cudaMemcpyHostToDevice, st));

    const size_t dir_smem = dir_smem_small + (chol_in_smem ? tri_bytes : 0);
    HF_CUDA_OK(cudaFuncSetAttribute(fit_direction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)dir_smem));
    // Schur-form solver: only the global block is factorised (needs a diagonal local block)
    const int sch_ng = std::max(nglob, 1), sch_nl = std::max(nfree - nglob, 1);
    // two launch shapes: "bulk" (thousands of fits: small CTAs, small staging tile, many CTAs per
    // SM) and "latency" (the lock-step tail: a handful of fits, everything in flight at once)
    struct SchurCfg { int threads, ltile, backup; size_t smem; };
    auto schur_cfg = [&](bool bulk) {
        SchurCfg c;
        c.ltile = std::min(sch_nl, sch_ng > 64 ? 16 : 32);
        if (!bulk && sizeof(double) * (size_t)sch_nl * (sch_ng | 1) <= 24 * 1024) c.ltile = sch_nl;
        c.threads = bulk ? (sch_ng <= 32 ? 64 : (sch_ng <= 96 ? 128 : 512)) : (sch_ng <= 96 ? 256 : 512);
        const size_t base = sizeof(double) * ((size_t)sch_ng * (sch_ng + 1) / 2 + 3 * (size_t)sch_ng +
                                              2 * (size_t)sch_nl + (size_t)c.ltile * (sch_ng | 1)) +
                            sizeof(int) * ((size_t)sch_ng + sch_nl + 2) + 16;
        const size_t copy = sizeof(double) * ((size_t)sch_ng * (sch_ng + 1) / 2);
        c.backup = (base + copy <= 64 * 1024) ? 1 : 0;
        c.smem = base + (c.backup ? copy : 0);
        return c;
    };
    const SchurCfg cfg_bulk = schur_cfg(true), cfg_lat = schur_cfg(false);
    const size_t sch_smem = std::max(cfg_bulk.smem, cfg_lat.smem);
    const char* sch_env = getenv("HF_FIT_SCHUR");
    const bool use_schur = plan->fd_ncolors <= 1 && sch_smem <= 220 * 1024 && !(sch_env && sch_env[0] == '0');
    if (use_schur)
        HF_CUDA_OK(cudaFuncSetAttribute(fit_direction_schur_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sch_smem));
    auto launch_dir = [&](const int* lst, int64_t n, int mode) -> int {
        if (use_schur) {
            const SchurCfg& c = (n >= 4096) ? cfg_bulk : cfg_lat;
            fit_direction_schur_kernel<<<(unsigned)n, c.threads, c.smem, st>>>(s, lst, n, mode, sch_ng, sch_nl,
                                                                                c.ltile, c.backup);
            plan->launches++;
            return HF_OK;
        }
        const int64_t dstep = chol_in_smem ? n : dir_grid_cap;
        for (int64_t i0 = 0; i0 < n; i0 += dstep) {
            const int64_t nb = std::min<int64_t>(dstep, n - i0);
            fit_direction_kernel<<<(unsigned)nb, kDirThreads, dir_smem, st>>>(s, lst + i0, nb, mode);
            plan->launches++;
            if (!chol_in_smem) HF_CUDA_OK(cudaStreamSynchronize(st));
        }
        return HF_OK;
    };

    auto blocks = [](int64_t n, int t) { return (unsigned)((n + t - 1) / t); };

    auto run_hessians = [&](const int* d_list, int64_t nlist, const int* mode, int64_t toy0,
                            int to_shared, bool need_asimov = true) -> int {
        if (use_ah) {
            for (int64_t r0 = 0; r0 < nlist; r0 += ah_cap) {
                const int64_t nl = std::min<int64_t>(ah_cap, nlist - r0);
                int rc = hf_hess_analytic_launch(plan, s.X, d_list + r0, nullptr, (int)nl, mode, 1, data, data_rows,
                                                 data_map, toy0, to_shared ? s.Hshared : s.H, to_shared, AHS,
                                                 d_struct, st);
                if (rc) return rc;
            }
            return HF_OK;
        }
        for (int64_t r0 = 0; r0 < nlist; r0 += rl_cap) {
            const int64_t nl = std::min<int64_t>(rl_cap, nlist - r0);
            const int* lst = d_list + r0;
            int rc = HF_OK;
            if (need_asimov) {  // Fisher entries: Asimov data of the entry's current point
                fit_gather_x_kernel<<<blocks(nl * P, 256), 256, 0, st>>>(s, lst, nl, XG);
                plan->launches++;
                rc = hf_expected_launch(plan, XG, nl, AS, st);
                if (rc) return rc;
            }
            const int64_t npts = nl * (ncol + 1);
            fit_fd_points_kernel<<<blocks(npts * P + nl * D, 256), 256, 0, st>>>(s, lst, nl, XP, ridx, mode,
                                                                                  data, AS, DR);
            plan->launches++;
            rc = hf_eval_launch(plan, XP, DR, nl, npts, VP, GP, st, ridx);
            if (rc) return rc;
            fit_fd_assemble_kernel<<<blocks(nl * (int64_t)P * P, 256), 256, 0, st>>>(
                s, lst, nl, GP, to_shared ? s.Hshared : s.H, to_shared);
            plan->launches++;
        }
        HF_CUDA_OK(cudaGetLastError());
        return HF_OK;
    };

    std::vector<int> h_list;
    int32_t* h_counters = plan->h_counters;
    const int max_rounds = opts.max_iter * 4 + 40;

    for (int64_t toy0 = 0; toy0 < N; toy0 += NC) {
        const int64_t nc = std::min<int64_t>(NC, N - toy0);
        s.NC = nc;
        s.toy0 = toy0;
        const double* init_c = init + (init_rows == 1 ? 0 : toy0 * P);
        const double* data_c = data + ((data_rows == 1 || data_map) ? 0 : toy0 * D);
        HF_CUDA_OK(cudaMemsetAsync(s.ownstart, 0, sizeof(int) * nc, st));
        fit_init_kernel<<<blocks(nc * P, 256), 256, 0, st>>>(s, init_c, init_rows);
        plan->launches++;
        int rc = hf_eval_launch(plan, s.X, data_c, data_map ? data_rows : (data_rows == 1 ? 1 : nc), nc, s.VT,
                                s.GT, st, data_map ? data_map + toy0 : nullptr);
        if (rc) return rc;
        fit_adopt_kernel<<<blocks(nc * P, 256), 256, 0, st>>>(s);
        plan->launches++;
        // shared start-up curvature: Fisher matrix at the first fit's start point
        {
            const int zero = 0;
            HF_CUDA_OK(cudaMemcpyAsync(s.rlist, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
            rc = run_hessians(s.rlist, 1, nullptr, toy0, 1);
            if (rc) return rc;
        }
        if (use_ah && init_rows != 1) {
            // fits that do not start where fit 0 starts get the Fisher matrix of their own start
            // point (multistart, scans with the POI fixed at different values); device-side count
            HF_CUDA_OK(cudaMemsetAsync(s.counters, 0, sizeof(int) * 8, st));
            fit_ownstart_list_kernel<<<blocks(nc, 256), 256, 0, st>>>(s, s.rlist, s.counters + 6);
            plan->launches++;
            for (int64_t r0 = 0; r0 < nc; r0 += ah_cap) {
                // (request r0 + i; the kernel leaves when r0 + i >= the device count)
                rc = hf_hess_analytic_launch(plan, s.X, s.rlist + r0, s.counters + 6, (int)std::min<int64_t>(ah_cap, nc - r0),
                                             nullptr, 1, data, data_rows, data_map, toy0, s.H, 0, AHS, d_struct, st);
                if (rc) return rc;
                if (r0 + ah_cap < nc) {  // next chunk: remaining count = max(count - ah_cap, 0)
                    hf_set_error("own-start curvature for more than %lld fits per chunk is not supported", (long long)ah_cap);
                    return HF_ERR_UNSUPPORTED;
                }
            }
        }
        fit_iota_kernel<<<blocks(nc, 256), 256, 0, st>>>(s.live, nc);
        plan->launches++;
        int64_t nl = nc;  // live fits
        int nfall = 0;    // Fisher fallbacks requested in the previous round
        double t_dir = 0, t_hess = 0, t_dir2 = 0, t_eval = 0, t_acc = 0;
        int n_rounds = 0; long long n_dir = 0, n_ref = 0, n_evals = 0, n_retry = 0;
        auto tick = [&]() { cudaStreamSynchronize(st); return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const bool prof = opts.verbose >= 2;
        for (int round = 0; round < max_rounds && nl > 0; ++round) {
            double t0 = prof ? tick() : 0.0;
            n_rounds++; n_dir += nl; n_evals += nl;
            // fits whose exact Hessian turned out unusable last round: Fisher matrix instead
            // (requested by the post-refresh direction pass, read back with last round's counts)
            if (nfall > 0) {
                rc = run_hessians(s.rlist2, nfall, s.refresh, toy0, 0);
                if (rc) return rc;
                rc = launch_dir(s.rlist2, nfall, -1);
                if (rc) return rc;
                nfall = 0;
            }
            HF_CUDA_OK(cudaMemsetAsync(s.counters, 0, sizeof(int) * 8, st));
            // ---- Newton directions of the live fits that accepted a step
            rc = launch_dir(s.live, nl, 1);
            if (rc) return rc;
            HF_CUDA_OK(cudaMemcpyAsync(h_counters, s.counters, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
            HF_CUDA_OK(cudaStreamSynchronize(st));
            const int nrefresh = h_counters[0];
            n_ref += nrefresh;
            if (prof) { double t1 = tick(); t_dir += t1 - t0; t0 = t1; }
            if (nrefresh > 0) {
                rc = run_hessians(s.rlist, nrefresh, s.refresh, toy0, 0, h_counters[5] > 0);
                if (rc) return rc;
                if (prof) { double t1 = tick(); t_hess += t1 - t0; t0 = t1; }
                // recompute the directions of the refreshed fits (this pass marks ownership)
                rc = launch_dir(s.rlist, nrefresh, 0);
                if (rc) return rc;
            }
            if (prof) { double t1 = tick(); t_dir2 += t1 - t0; t0 = t1; }
            // ---- one trial point per live fit, evaluated as one compact K2 batch
            fit_trial_kernel<<<blocks(nl * P, 256), 256, 0, st>>>(s, s.live, nl, s.XT, ridx);
            plan->launches++;
            rc = hf_eval_launch(plan, s.XT, data, data_rows, nl, s.VT, s.GT, st, ridx);
            if (rc) return rc;
            if (prof) { double t1 = tick(); t_eval += t1 - t0; t0 = t1; }
            // (counters 1 and 3 are still zero from the memset at the top of the round)
            fit_accept_kernel<<<blocks(nl * 32, 256), 256, 0, st>>>(s, s.live, nl, s.XT, s.VT, s.GT, s.live_next);
            plan->launches++;
            HF_CUDA_OK(cudaMemcpyAsync(h_counters, s.counters, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
            HF_CUDA_OK(cudaStreamSynchronize(st));
            if (opts.verbose) fprintf(stderr, "[hf_fit] round %d: live %lld refresh %d accepted %d -> live %d\n", round, (long long)nl, nrefresh, h_counters[3], h_counters[1]);
            nl = h_counters[1];
            nfall = h_counters[2];
            n_retry += h_counters[4];
            std::swap(s.live, s.live_next);
            if (prof) { double t1 = tick(); t_acc += t1 - t0; t0 = t1; }
        }
        if (prof)
            fprintf(stderr, "[hf_fit] chunk of %lld fits: %d rounds, dir1 %.1f ms (%lld), hessians %.1f ms (%lld), dir2 %.1f ms, "
                            "trial+eval %.1f ms (%lld), accept %.1f ms, factorisation retries %lld\n", (long long)nc, n_rounds, 1e3 * t_dir, n_dir,
                    1e3 * t_hess, n_ref, 1e3 * t_dir2, 1e3 * t_eval, n_evals, 1e3 * t_acc, n_retry);
        fit_finish_kernel<<<blocks(nc * P, 256), 256, 0, st>>>(s, toy0, x_out, fun_out, status_out, niter_out);
        plan->launches++;
        HF_CUDA_OK(cudaGetLastError());
    }
    return HF_OK;
}
