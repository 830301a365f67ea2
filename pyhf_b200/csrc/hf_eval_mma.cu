// hf_main_mma_kernel: the dominant kernel of K1/K2 with the histosys contraction on the fp64
// tensor-core path (DMMA, mma.sync.m8n8k4.f64) and the tables staged in shared memory by
// bulk asynchronous copies (cp.async.bulk + mbarrier, the TMA engine).
//
// Why: ncu on the first kernel (profiles/r1_hf_main_kernel_full_v1.csv) showed the vector-FMA
// formulation limited by LSU wavefronts, not by the fp64 pipe: a broadcast LDS.128 delivers 16
// useful bytes per 2 wavefronts, the backward sweep needed 20 wavefronts per 32 DFMA.  DMMA has
// the same measured peak as DFMA on B200 (tools/dmma_peak.cu: 37.0 vs 36.8 TFLOP/s) but takes
// its operands as lane-distributed fragments: 2 doubles per lane per 256 FMAs, read from shared
// memory conflict-free (row strides == 4 mod 16 doubles).
//
// One CTA (4 warps) owns HF_MMA_M = 32 parameter points and sweeps the bins in tiles of 32
// (128 cells = 32 bins x 4 histosys rows):
//   forward  Delta[pt][cell] = sum_k coef[pt][k] T[k][cell]      k = (modifier, T1|T2), K = 2H
//            warp = 32 cells, 4x4 DMMA tiles, K streamed in chunks of 40 columns
//   epilogue per (point, bin): factors, sample sum, Poisson term, w = n/lambda - 1, w*F -> smem
//   backward acc[pt][k] += sum_cell wF[pt][cell] T[k][cell]       warp = 8 points, all k
//            accumulators stay in registers over all bin tiles, flushed once per CTA
// Table chunks [40][132] (42 KB) are double buffered; chunk i+1 is in flight while chunk i is
// consumed.  Reference being replaced: see hf_eval.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hf_internal.cuh"

namespace {

constexpr int BT = HF_MMA_BT, NC = HF_MMA_NC, CS = HF_MMA_CS, KC = HF_MMA_KC, MM = HF_MMA_M;
constexpr int NT_BWD = KC / 8;  // n-tiles per chunk in the backward product
constexpr uint32_t CHUNK_BYTES = KC * CS * sizeof(double);

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double cell_factor(const DevPlan& pl, const Work& w, int64_t ld, int64_t p,
                                              int s, int bb, int seg, double& bfprod) {
    bfprod = 1.0;
    for (int k = 0; k < pl.K; ++k) {
        const int q = pl.bf_par[((int64_t)k * pl.S + s) * pl.NB + bb];
        if (q >= 0) bfprod *= w.parsT[(int64_t)q * ld + p];
    }
    return w.FsegT[((int64_t)s * pl.NSEG + seg) * ld + p];
}

// d/dgamma_k of cell (s, bb): wfb = w * Fseg * base ; times the product of the other per-bin factors
__device__ __forceinline__ void cell_bf_grad(const DevPlan& pl, const Work& w, int64_t ld, int64_t p,
                                             int s, int bb, double wfb) {
    for (int k = 0; k < pl.K; ++k) {
        const int q = pl.bf_par[((int64_t)k * pl.S + s) * pl.NB + bb];
        if (q < 0) continue;
        double oth = wfb;
        for (int k2 = 0; k2 < pl.K; ++k2) {
            if (k2 == k) continue;
            const int q2 = pl.bf_par[((int64_t)k2 * pl.S + s) * pl.NB + bb];
            if (q2 >= 0) oth *= w.parsT[(int64_t)q2 * ld + p];
        }
        atomicAdd(&w.gradT[(int64_t)q * ld + p], oth);
    }
}

constexpr int SMAX = 8;  // samples handled by the register-resident per-(sample, segment) sums
constexpr int NSTAGE = 2;  // table-chunk buffers in flight (42 KB each); 3 measured no faster (also with mbarrier release)
constexpr int NTHR = 256;  // 8 warps: forward = 16 cells each, backward = (8 points) x (64 cells) each

template <int NCH, bool GRAD, bool HASBF>
__global__ void __launch_bounds__(NTHR, 1)
hf_main_mma_kernel(DevPlan pl, Work w, int64_t B, int64_t ld, const double* __restrict__ data,
                   int per_point, const int* __restrict__ row_idx, int KS) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* coefs = reinterpret_cast<double*>(smem_raw);            // [MM][KS]
    double* tbuf = coefs + (size_t)MM * KS;                          // [NSTAGE][KC*CS]
    double* wfs2 = tbuf + NSTAGE * KC * CS;                          // [2][MM][CS]  Delta, then w*F
    uint64_t* bars = reinterpret_cast<uint64_t*>(wfs2 + 2 * MM * CS); // full[NSTAGE] | coefs | empty[NSTAGE]
    int64_t* row_off = reinterpret_cast<int64_t*>(bars + 8);         // [MM]
    // tile constants, double buffered: tile i+1 is fetched (cp.async) while tile i is worked on
    double* nd_s2 = reinterpret_cast<double*>(row_off + MM);         // [2][MM][BT+1] observed counts
    double* nom_s2 = nd_s2 + 2 * MM * (BT + 1);                      // [2][S][BT]    nominal of the tile
    int* seg_s2 = reinterpret_cast<int*>(nom_s2 + 2 * SMAX * BT);    // [2][BT]
    int* srow_s = seg_s2 + 2 * BT;                                   // [SMAX]     histosys row of a sample
    int* bf_s2 = srow_s + SMAX;                                      // [2][K][S][BT] per-bin factor indices
    const int bf_stride = pl.K * pl.S * BT;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int64_t pt0 = (int64_t)blockIdx.x * MM;
    const int NB = pl.NB, S = pl.S, ntiles = pl.mma_ntiles;
    const double* tt = pl.mma_tt;

    if (tid == 0) {
        for (int i = 0; i <= NSTAGE; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < NSTAGE; ++i) mbar_init(&bars[NSTAGE + 1 + i], NTHR / 32);  // one arrival per warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < MM) {
        int64_t pt = pt0 + tid;
        if (pt >= B) pt = B - 1;
        const int64_t row = row_idx ? (int64_t)row_idx[pt] : (per_point ? pt : 0);
        row_off[tid] = row * pl.D;
    }
    // sample positions: 0..3 = histosys rows (sample hs_row_sample[r]), 4..7 = samples without histosys
    if (tid < SMAX) {
        int smp = -1;
        if (tid < 4) {
            if (tid < pl.R) smp = pl.hs_row_sample[tid];
        } else {
            int cnt = 0;
            for (int s2 = 0; s2 < pl.S; ++s2)
                if (pl.sample_hs_row[s2] < 0) {
                    if (cnt == tid - 4) smp = s2;
                    ++cnt;
                }
        }
        srow_s[tid] = smp;
    }
    __syncthreads();
    int pos_smp[SMAX], noff[SMAX], nother = 0;
#pragma unroll
    for (int q = 0; q < SMAX; ++q) {
        pos_smp[q] = srow_s[q];
        noff[q] = max(pos_smp[q], 0) * BT;          // row of the tile's nominal table
        if (q >= 4 && pos_smp[q] >= 0) nother = q - 3;  // samples without histosys fill 4, 5, ...
    }
    auto prefetch_tile = [&](int tile) {  // all threads; completion: cp_async_wait_all + barrier
        const int b0 = tile * BT, buf = tile & 1;
        double* nom_d = nom_s2 + buf * SMAX * BT;
        int* bf_d = bf_s2 + buf * bf_stride;
        for (int i = tid; i < pl.S * BT; i += NTHR) {
            const int s = i / BT, bl = i - s * BT;
            cp_async8(&nom_d[i], &pl.nominal[(int64_t)s * pl.NB + min(b0 + bl, pl.NB - 1)]);
        }
        for (int i = tid; i < bf_stride; i += NTHR) {
            const int ks = i / BT, bl = i - ks * BT;
            cp_async4(&bf_d[i], &pl.bf_par[(int64_t)ks * pl.NB + min(b0 + bl, pl.NB - 1)]);
        }
        if (tid < BT) cp_async4(&seg_s2[buf * BT + tid], &pl.bin_seg[min(b0 + tid, pl.NB - 1)]);
        // observed counts of the 32 points: a warp reads one point's 32 bins per pass
        for (int q = warp; q < MM; q += NTHR / 32)
            cp_async8(&nd_s2[(buf * MM + q) * (BT + 1) + lane], &data[row_off[q] + min(b0 + lane, pl.NB - 1)]);
    };
    prefetch_tile(0);
    int issued = 0;    // table chunks requested so far (thread 0)
    int consumed = 0;  // table chunks consumed so far (uniform)
    const int per_tile = (GRAD ? 2 : 1) * NCH;
    const int n_items = ntiles * per_tile;
    auto issue_next = [&]() {  // thread 0 only
        if (issued < n_items) {
            const int tile = issued / per_tile, chunk = (issued % per_tile) % NCH;
            // the buffer is free once every warp has arrived for the item that used it last
            if (issued >= NSTAGE) mbar_wait(&bars[NSTAGE + 1 + issued % NSTAGE], ((issued / NSTAGE) - 1) & 1);
            uint64_t* bar = &bars[issued % NSTAGE];
            mbar_expect_tx(bar, CHUNK_BYTES);
            bulk_g2s(tbuf + (size_t)(issued % NSTAGE) * KC * CS,
                     tt + ((size_t)tile * NCH + chunk) * KC * CS, CHUNK_BYTES, bar);
        }
        ++issued;
    };
    if (tid == 0) {
        mbar_expect_tx(&bars[NSTAGE], (uint32_t)(MM * KS * sizeof(double)));
        bulk_g2s(coefs, w.coefK + pt0 * KS, (uint32_t)(MM * KS * sizeof(double)), &bars[NSTAGE]);
        for (int i = 0; i < NSTAGE - 1; ++i) issue_next();  // prologue: NSTAGE-1 chunks in flight
    }
    mbar_wait(&bars[NSTAGE], 0);

    // backward accumulators: this warp's 8 points x all K columns
    double bacc[NCH][NT_BWD][2];
    if (GRAD) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int n = 0; n < NT_BWD; ++n) bacc[c][n][0] = bacc[c][n][1] = 0.0;
    }
    // epilogue state: thread = point `lane`, warp = 8 consecutive bins of every tile
    const int64_t p = pt0 + lane;
    double val = 0.0;      // main-term partial sum of this point over this warp's bins
    double gacc[SMAX];     // per-sample sums of w*bf*base inside the current segment
    double fseg[SMAX];     // scalar-factor products of this point in the current segment
    int gseg = -1;
#pragma unroll
    for (int s = 0; s < SMAX; ++s) gacc[s] = fseg[s] = 0.0;
    const int ra = 2 * (t & 1);

    for (int tile = 0; tile < ntiles; ++tile) {
        // Delta / w*F of consecutive tiles alternate between two buffers: a warp that is ahead
        // writes the next tile's Delta while a slow one still reads this tile's w*F in the
        // backward product (the buffer it writes was last read two tiles ago, and two CTA
        // barriers lie in between)
        double* wfs = wfs2 + (size_t)(tile & 1) * MM * CS;
        // ---------------------------------------------------------------- forward contraction
        {
            double facc[4][2][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) facc[mi][ni][0] = facc[mi][ni][1] = 0.0;
#pragma unroll 1
            for (int chunk = 0; chunk < NCH; ++chunk) {
                if (tid == 0) issue_next();
                mbar_wait(&bars[consumed % NSTAGE], (consumed / NSTAGE) & 1);
                const double* tb = tbuf + (size_t)(consumed % NSTAGE) * KC * CS + 16 * warp + g;
                const double* ca = coefs + (size_t)g * KS + chunk * KC + t;
#pragma unroll 5
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    double a[4], b[2];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = ca[(size_t)(8 * mi) * KS + k4 * 4];
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) b[ni] = tb[(k4 * 4 + t) * CS + 8 * ni];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni)
                            dmma884(facc[mi][ni][0], facc[mi][ni][1], a[mi], b[ni]);
                }
                // this warp is done with the buffer (no CTA-wide barrier: warps may run one chunk apart)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[NSTAGE + 1 + consumed % NSTAGE]);
                ++consumed;
            }
            // Delta fragments -> shared [point][cell]; lane owns point 8mi+g, bin 4*warp+2ni+(t>>1),
            // rows ra, ra+1 of every (mi, ni) tile
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    double2 st;
                    st.x = facc[mi][ni][0];
                    st.y = facc[mi][ni][1];
                    const int bl = 4 * warp + 2 * ni + (t >> 1);
                    *reinterpret_cast<double2*>(&wfs[(size_t)(8 * mi + g) * CS + 4 * bl + ra]) = st;
                }
        }
        // ---- the tile's model constants were requested one tile ago (prefetch_tile): make them
        //      visible, then request the next tile's into the other buffer (its last readers, the
        //      epilogue of tile - 1, are past the barrier above)
        cp_async_wait_all();
        __syncthreads();
        if (tile + 1 < ntiles) prefetch_tile(tile + 1);
        const double* nd_s = nd_s2 + (tile & 1) * MM * (BT + 1);
        const double* nom_s = nom_s2 + (tile & 1) * SMAX * BT;
        const int* seg_s = seg_s2 + (tile & 1) * BT;
        const int* bf_s = bf_s2 + (tile & 1) * bf_stride;

        // ---------------------------------------------------------------- epilogue
        // positions 0..3 are the histosys rows (Delta in cell[0..3]), 4..7 the other samples.
        // Fast path (the warp's 4 bins lie in one segment): the 4 bins are processed as 4
        // interleaved dependency chains (the log / divide chains are ~1000 cycles each and only
        // two warps share a scheduler); slow path: one bin at a time with a segment switch.
        auto switch_segment = [&](int seg) {
            if (GRAD && gseg >= 0) {
#pragma unroll
                for (int q = 0; q < SMAX; ++q)
                    if (pos_smp[q] >= 0 && gacc[q] != 0.0)
                        atomicAdd(&w.GsegT[((int64_t)pos_smp[q] * pl.NSEG + gseg) * ld + p], gacc[q]);
            }
#pragma unroll
            for (int q = 0; q < SMAX; ++q) {
                gacc[q] = 0.0;
                // scalar factors of this point in the new segment stay in registers
                fseg[q] = (pos_smp[q] >= 0) ? w.FsegT[((int64_t)pos_smp[q] * pl.NSEG + seg) * ld + p] : 0.0;
            }
            gseg = seg;
        };
        const int seg_first = seg_s[4 * warp], seg_last = seg_s[4 * warp + 3];
        bool plain = true;  // none of the warp's 4 bins carries a per-bin factor
        if (HASBF) {
            for (int ks = 0; ks < pl.K * S; ++ks) {
                const int* r4 = bf_s + ks * BT + 4 * warp;
                plain = plain && ((r4[0] & r4[1] & r4[2] & r4[3]) < 0);
            }
        }
        // bins are ordered, so the warp's 4 bins span at most two segments iff the inner ones
        // belong to the first or the last one; each segment gets one (masked) pass
        const bool two_seg = (seg_s[4 * warp + 1] == seg_first || seg_s[4 * warp + 1] == seg_last) &&
                             (seg_s[4 * warp + 2] == seg_first || seg_s[4 * warp + 2] == seg_last);
        if (plain && two_seg) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int seg_cur = pass ? seg_last : seg_first;
                if (pass && seg_last == seg_first) break;
                if (seg_cur != gseg) switch_segment(seg_cur);
                double lam[4], wgt[4];
                bool act[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int bl = 4 * warp + j;
                    act[j] = seg_s[bl] == seg_cur;
                    const double4 dlt = *reinterpret_cast<const double4*>(wfs + (size_t)lane * CS + 4 * bl);
                    const double dl[4] = {dlt.x, dlt.y, dlt.z, dlt.w};
                    // positions without a sample carry fseg = 0 and read nominal row 0: no predicates
                    // (only the rarely used positions 5..7 keep a uniform guard)
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < SMAX; ++q) {
                        if (q < 5 || q < 4 + nother) {
                            double base = nom_s[noff[q] + bl];
                            if (q < 4) base += dl[q];
                            acc = fma(fseg[q], base, acc);
                        }
                    }
                    lam[j] = acc;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int bl = 4 * warp + j;
                    const bool valid = (tile * BT + bl < NB) && act[j];
                    const double n = nd_s[lane * (BT + 1) + bl];
                    if (valid) val += hf_pois_term(n, lam[j]);
                    // n * rcp(lambda) (correctly rounded reciprocal + one FMA) instead of the IEEE divide:
                    // <= 1.5 ulp on n/lambda, same inf / NaN behaviour at lambda = 0
                    // (n == 0: the derivative of xlogy(0, l) - l is exactly -1, also at l == 0)
                    wgt[j] = valid ? ((n == 0.0) ? -1.0 : fma(n, __drcp_rn(lam[j]), -1.0)) : 0.0;
                }
                if (GRAD) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (!act[j]) continue;  // warp-uniform
                        const int bl = 4 * warp + j;
                        double* cell = wfs + (size_t)lane * CS + 4 * bl;
                        const double4 dlt = *reinterpret_cast<const double4*>(cell);
                        const double dl[4] = {dlt.x, dlt.y, dlt.z, dlt.w};
                        double wf[4];
#pragma unroll
                        for (int q = 0; q < SMAX; ++q) {
                            if (q < 5 || q < 4 + nother) {
                                double base = nom_s[noff[q] + bl];
                                if (q < 4) base += dl[q];
                                gacc[q] = fma(wgt[j], base, gacc[q]);  // (never flushed for empty positions)
                                if (q < 4) wf[q] = wgt[j] * fseg[q];
                            }
                        }
                        double4 st;  // A' operand of the backward product
                        st.x = wf[0]; st.y = wf[1]; st.z = wf[2]; st.w = wf[3];
                        *reinterpret_cast<double4*>(cell) = st;
                    }
                }
            }
        } else {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int bl = 4 * warp + j;
            const int b = tile * BT + bl;
            const bool valid = b < NB;
            const int bb = valid ? b : NB - 1;
            const int seg = seg_s[bl];
            double* cell = wfs + (size_t)lane * CS + 4 * bl;
            if (seg != gseg) switch_segment(seg);  // warp-uniform: all lanes look at the same bin
            const double4 dlt = *reinterpret_cast<const double4*>(cell);
            double basev[SMAX], bfv[SMAX];
            basev[0] = dlt.x; basev[1] = dlt.y; basev[2] = dlt.z; basev[3] = dlt.w;
#pragma unroll
            for (int q = 4; q < SMAX; ++q) basev[q] = 0.0;
            double lam = 0.0;
#pragma unroll
            for (int q = 0; q < SMAX; ++q) {
                bfv[q] = 1.0;
                if (pos_smp[q] >= 0) {
                    const int s = pos_smp[q];
                    if (HASBF) {
                        for (int k = 0; k < pl.K; ++k) {
                            const int par = bf_s[(k * S + s) * BT + bl];
                            if (par >= 0) bfv[q] *= w.parsT[(int64_t)par * ld + p];
                        }
                    }
                    basev[q] += nom_s[s * BT + bl];
                    lam = fma(fseg[q] * bfv[q], basev[q], lam);
                }
            }
            const double n = nd_s[lane * (BT + 1) + bl];
            if (valid) val += hf_pois_term(n, lam);
            if (GRAD) {
                const double wgt = valid ? ((n == 0.0) ? -1.0 : n / lam - 1.0) : 0.0;
                double wf[4];
#pragma unroll
                for (int q = 0; q < SMAX; ++q) {
                    if (pos_smp[q] >= 0) {
                        const double wb = wgt * bfv[q];
                        gacc[q] = fma(wb, basev[q], gacc[q]);
                        if (q < 4) wf[q] = wb * fseg[q];
                        if (HASBF && wgt != 0.0)
                            cell_bf_grad(pl, w, ld, p, pos_smp[q], bb, wgt * fseg[q] * basev[q]);
                    } else if (q < 4) {
                        wf[q] = 0.0;
                    }
                }
                double4 st;  // A' operand of the backward product
                st.x = wf[0]; st.y = wf[1]; st.z = wf[2]; st.w = wf[3];
                *reinterpret_cast<double4*>(cell) = st;
            }
        }
        }

        // ---------------------------------------------------------------- backward contraction
        if (GRAD) {
            __syncthreads();  // w*F tile complete
#pragma unroll
            for (int chunk = 0; chunk < NCH; ++chunk) {
                if (tid == 0) issue_next();
                mbar_wait(&bars[consumed % NSTAGE], (consumed / NSTAGE) & 1);
                // warp = (m-tile: points 8*(warp&3)+g) x (cell half: 64*(warp>>2) ..)
                const int c0 = 64 * (warp >> 2);
                const double* tb = tbuf + (size_t)(consumed % NSTAGE) * KC * CS + (size_t)g * CS + t + c0;
                const double* arow = wfs + (size_t)(8 * (warp & 3) + g) * CS + t + c0;
#pragma unroll 4
                for (int c4 = 0; c4 < NC / 8; ++c4) {
                    const double a = arow[4 * c4];
#pragma unroll
                    for (int ni = 0; ni < NT_BWD; ++ni) {
                        const double b = tb[(size_t)(8 * ni) * CS + 4 * c4];
                        dmma884(bacc[chunk][ni][0], bacc[chunk][ni][1], a, b);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[NSTAGE + 1 + consumed % NSTAGE]);
                ++consumed;
            }
        }
    }

    // ---------------------------------------------------------------- flush
    atomicAdd(&w.mainv[p], val);  // the eight warps hold different bins of the same points
    if (GRAD) {
        if (gseg >= 0) {
#pragma unroll
            for (int q = 0; q < SMAX; ++q)
                if (pos_smp[q] >= 0 && gacc[q] != 0.0)
                    atomicAdd(&w.GsegT[((int64_t)pos_smp[q] * pl.NSEG + gseg) * ld + p], gacc[q]);
        }
        // bacc[chunk][ni] = C'[point 8*warp+g][k = chunk*KC + 8ni + 2t, +1] = (T1, T2) sums of one modifier
        const int64_t pw = pt0 + 8 * (warp & 3) + g;  // the two cell halves add their partial sums
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk)
#pragma unroll
            for (int ni = 0; ni < NT_BWD; ++ni) {
                const int h = (chunk * KC + 8 * ni + 2 * t) >> 1;
                if (h < pl.H) {
                    const int par = pl.hs_par[h];
                    double c1, c2, d1, d2;
                    hf_hs_coeffs(pl.hs_code, w.parsT[(int64_t)par * ld + pw], c1, c2, d1, d2);
                    atomicAdd(&w.gradT[(int64_t)par * ld + pw],
                              d1 * bacc[chunk][ni][0] + d2 * bacc[chunk][ni][1]);
                }
            }
    }
}

// point-major coefficient rows for the A operand: coefK[pt][2h] = c1_h, [2h+1] = c2_h, zero padded
__global__ void hf_prep_mma_kernel(DevPlan pl, Work w, int64_t ld, int KS) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int half = KS / 2;
    if (idx >= ld * half) return;
    const int64_t pt = idx / half;
    const int h = (int)(idx - pt * half);
    double2 v;
    v.x = 0.0;
    v.y = 0.0;
    if (h < pl.H) {
        double d1, d2;
        hf_hs_coeffs(pl.hs_code, w.parsT[(int64_t)pl.hs_par[h] * ld + pt], v.x, v.y, d1, d2);
    }
    *reinterpret_cast<double2*>(&w.coefK[pt * KS + 2 * h]) = v;
}

template <int NCH>
int launch(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
           const int* row_idx, bool grad, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    const int KS = hf_mma_ks(pl.H);
    size_t smem = sizeof(double) * ((size_t)MM * KS + NSTAGE * KC * CS + 2 * MM * CS) + 64 + 8 * MM + 128 +
                  2 * sizeof(double) * (MM * (BT + 1) + SMAX * BT) +
                  sizeof(int) * (2 * BT + SMAX + 2 * (size_t)pl.K * pl.S * BT) + 64;
    const int64_t half = KS / 2;
    hf_prep_mma_kernel<<<(unsigned)((ld * half + 255) / 256), 256, 0, st>>>(pl, plan->w, ld, KS);
    plan->launches++;
    const dim3 grid((unsigned)(ld / MM));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (plan->timing) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
#define HF_MMA_GO(G_, BF_)                                                                          \
    {                                                                                               \
        auto k = hf_main_mma_kernel<NCH, G_, BF_>;                                                  \
        HF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<grid, NTHR, smem, st>>>(pl, plan->w, B, ld, data, per_point, row_idx, KS);              \
    }
    if (grad) {
        if (pl.K > 0) HF_MMA_GO(true, true) else HF_MMA_GO(true, false)
    } else {
        if (pl.K > 0) HF_MMA_GO(false, true) else HF_MMA_GO(false, false)
    }
#undef HF_MMA_GO
    if (plan->timing) {
        cudaEventRecord(e1, st);
        plan->ev.push_back(e0);
        plan->ev.push_back(e1);
    }
    HF_CUDA_OK(cudaGetLastError());
    plan->launches++;
    return HF_OK;
}

}  // namespace

int hf_main_mma_launch(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
                       const int* row_idx, bool grad, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    if (!pl.mma_ok || (ld % MM) != 0) return HF_ERR_UNSUPPORTED;
    switch (pl.mma_nchunks) {
        case 1: return launch<1>(plan, B, ld, data, per_point, row_idx, grad, st);
        case 2: return launch<2>(plan, B, ld, data, per_point, row_idx, grad, st);
        case 3: return launch<3>(plan, B, ld, data, per_point, row_idx, grad, st);
        case 4: return launch<4>(plan, B, ld, data, per_point, row_idx, grad, st);
        case 5: return launch<5>(plan, B, ld, data, per_point, row_idx, grad, st);
        case 6: return launch<6>(plan, B, ld, data, per_point, row_idx, grad, st);
        default: return HF_ERR_UNSUPPORTED;
    }
}
