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
// One CTA (8 warps) owns HF_MMA_M = 32 parameter points and sweeps the bins in tiles of 32
// (128 cells = 32 bins x 4 histosys rows):
//   forward  Delta[pt][cell] = sum_k coef[pt][k] T[k][cell]      k = (modifier, T1|T2), K = 2H
//            warp = 16 cells x 32 points, K streamed in chunks of 40 columns
//   epilogue per (point, bin): factors, sample sum, Poisson term, w = n/lambda - 1, w*F -> smem
//   backward acc[pt][k] += sum_cell wF[pt][cell] T[k][cell]       warp = 8 points x 64 cells, all k
//            accumulators stay in registers over all bin tiles, flushed once per CTA
// The CTA is two independent HALVES of 4 warps: half h owns bins 16h .. 16h+15 of every tile
// (cells 64h .. 64h+63) through all three phases, with its own stream of table chunks
// of 40 rows (22.5 KB, three buffers, bulk copies + mbarriers), its own tile constants and its own
// named barrier.  The epilogue is a latency chain (log, reciprocal) that leaves the fp64 pipe idle;
// the halves are free to drift apart, so that one can be in its DMMA phases while the other is in
// its epilogue (HF_MMA_TRACE=1 prints the phase time stamps of CTA 0).  Reference being replaced:
// see hf_eval.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "hf_internal.cuh"

namespace {

constexpr int BT = HF_MMA_BT, CS = HF_MMA_CS, KC = HF_MMA_KC, MM = HF_MMA_M;
constexpr int NCH_CELLS = HF_MMA_NCH, CSH = HF_MMA_CSH, HB = BT / 2;  // half tile: 64 cells, 16 bins
constexpr int NT_BWD = KC / 8;  // n-tiles per chunk in the backward product
constexpr uint32_t CHUNK_BYTES = KC * CSH * sizeof(double);

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
__device__ __forceinline__ void half_barrier(int half) {  // the 128 threads of one half of the CTA
    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ln(x) for the Poisson term of the epilogue, where it is the longest dependency chain: table of
// 128 intervals of the mantissa in [0.75, 1.5) (1/c_i and -ln(1/c_i) of the interval centres), then
// ln(m) = ln(c_i) + log1p(r), r = m/c_i - 1, |r| < 2^-8, log1p as a degree-7 polynomial (next term
// < 7e-21 absolute).  About 1 ulp; zero, denormal, negative, inf and NaN arguments go to log().
constexpr int LOGTAB_N = 128;
__device__ double2 g_logtab[LOGTAB_N];
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(x);
    int e = (hi >> 20) - 1023;
    int mhi = (hi & 0x000fffff) | 0x3ff00000;
    const int up = mhi >= 0x3ff80000;  // mantissa in [1.5, 2) -> [0.75, 1)
    mhi -= up << 20;
    e += up;
    const double m = __hiloint2double(mhi, __double2loint(x));
    const double2 t = tab[(mhi - 0x3fe80000) >> 13];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(r, p, 1.0 / 5.0);
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    const double ed = (double)e;
    const double small = fma(ed, 1.90821492927058770002e-10, fma(r * r, p, r));
    return fma(ed, 6.93147180369123816490e-01, t.y) + small;
}
__device__ __forceinline__ double pois_term_fast(double n, double lam, const double2* __restrict__ tab) {
    const double xl = (n == 0.0 && !(lam != lam)) ? 0.0 : n * fast_log(lam, tab);
    return xl - lam;
}

// A operand of the forward product in shared memory: element (column k, point p) of the CTA's
// coefficient block.  Column-major with the 16 point PAIRS of a column XOR-swizzled by the column, so
// that one 16-byte load gives a lane the A fragments of two m-tiles (points 2g, 2g+1) and the four
// columns k4*4 + t of a quarter warp fall into different banks.
__device__ __forceinline__ int coef_at(int k, int p) { return k * MM + ((((p >> 1) ^ (2 * (k & 3))) << 1) | (p & 1)); }

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
// table-chunk buffers per half (22.5 KB each).  A half (one warp per scheduler) consumes a chunk in
// >= 1 280 cycles, a bulk copy takes ~2 500 from request to completion (HF_MMA_TRACE: with two
// buffers the DMMA phases of a half ran at half the pipe rate, waiting for chunks), so two
// chunks have to be in flight while the third is consumed
constexpr int NSTAGE = 3;
constexpr int NTHR = 256;  // 8 warps: forward = 16 cells each, backward = (8 points) x (64 cells) each
constexpr int HTHR = NTHR / 2;  // threads of a half

// NPOS: sample positions kept in registers (0..3 the histosys rows, 4.. the samples without histosys);
// 5 covers plans with at most one histosys-free sample and keeps 18 registers free for the products
template <int NCH, bool GRAD, bool HASBF, int NPOS>
__global__ void __launch_bounds__(NTHR, 1)
hf_main_mma_kernel(DevPlan pl, Work w, int64_t B, int64_t ld, const double* __restrict__ data,
                   int per_point, const int* __restrict__ row_idx, int KS, long long* trace, int ltab_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* coefs = reinterpret_cast<double*>(smem_raw);            // [KS][MM] column-major, pairs of points swizzled (coef_at)
    double* tbuf_all = coefs + (size_t)MM * KS;                      // [2 halves][NSTAGE][KC*CSH]
    double* wfs = tbuf_all + 2 * NSTAGE * KC * CSH;                  // [MM][CS]  Delta, then w*F
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(wfs + MM * CS); // per half: full[NSTAGE] | empty[NSTAGE]
    int64_t* row_off = reinterpret_cast<int64_t*>(bars_all + 16);    // [MM]
    // tile constants: those of tile i+1 are requested (cp.async) when the epilogue of tile i is done
    // and land during the two products that follow
    double* nom_s = reinterpret_cast<double*>(row_off + MM);         // [S][BT]    nominal of the tile
    int* seg_s = reinterpret_cast<int*>(nom_s + SMAX * BT);          // [BT]
    int* srow_s = seg_s + BT;                                        // [SMAX]     histosys row of a sample
    int* bf_s = srow_s + SMAX;                                       // [K][S][BT] per-bin factor indices
    double2* ltab = reinterpret_cast<double2*>(smem_raw + ltab_off); // [LOGTAB_N] table of fast_log

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int half = warp >> 2, hw = warp & 3, htid = tid & (HTHR - 1);
    double* tbuf = tbuf_all + (size_t)half * NSTAGE * KC * CSH;
    uint64_t* bars = bars_all + half * 2 * NSTAGE;  // full[NSTAGE] | empty[NSTAGE] of this half
    const int64_t pt0 = (int64_t)blockIdx.x * MM;
    const int NB = pl.NB, S = pl.S, ntiles = pl.mma_ntiles;
    const double* tt = pl.mma_tt;

    if (tid == 0) {
        for (int hh = 0; hh < 2; ++hh) {
            for (int i = 0; i < NSTAGE; ++i) mbar_init(&bars_all[hh * 2 * NSTAGE + i], 1);
            for (int i = 0; i < NSTAGE; ++i)
                mbar_init(&bars_all[hh * 2 * NSTAGE + NSTAGE + i], HTHR / 32);  // one arrival per warp of the half
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < LOGTAB_N) ltab[tid] = g_logtab[tid];
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
    int pos_smp[NPOS], noff[NPOS], nother = 0;
#pragma unroll
    for (int q = 0; q < NPOS; ++q) {
        pos_smp[q] = srow_s[q];
        noff[q] = max(pos_smp[q], 0) * BT;          // row of the tile's nominal table
        if (q >= 4 && pos_smp[q] >= 0) nother = q - 3;  // samples without histosys fill 4, 5, ...
    }
    // the threads of a half fetch the constants of the half's 16 bins; completion:
    // cp_async_wait_all + half_barrier
    auto prefetch_tile = [&](int tile) {
        const int b0 = tile * BT, hb0 = HB * half;
        for (int i = htid; i < pl.S * HB; i += HTHR) {
            const int s = i / HB, bl = hb0 + (i - s * HB);
            cp_async8(&nom_s[s * BT + bl], &pl.nominal[(int64_t)s * pl.NB + min(b0 + bl, pl.NB - 1)]);
        }
        for (int i = htid; i < pl.K * pl.S * HB; i += HTHR) {
            const int ks = i / HB, bl = hb0 + (i - ks * HB);
            cp_async4(&bf_s[ks * BT + bl], &pl.bf_par[(int64_t)ks * pl.NB + min(b0 + bl, pl.NB - 1)]);
        }
        if (htid < HB) cp_async4(&seg_s[hb0 + htid], &pl.bin_seg[min(b0 + hb0 + htid, pl.NB - 1)]);
    };
    const double* drow = data + row_off[lane];  // observed counts of this lane's point (epilogue: lane = point)
    prefetch_tile(0);
    int issued = 0;    // table chunks of this half requested so far (uniform within the half)
    int consumed = 0;  // table chunks of this half consumed so far (uniform within the half)
    const int per_tile = (GRAD ? 2 : 1) * NCH;
    const int n_items = ntiles * per_tile;
    // the warps of a half take turns as the producer (lane 0 of warp `issued % 4`): a fixed producer
    // warp falls behind the others by its waits for free buffers and the half pays the skew at its barriers
    auto issue_next = [&]() {  // every thread of the half calls it; one lane does the work
        if (issued < n_items && lane == 0 && hw == (issued & 3)) {
            const int tile = issued / per_tile, chunk = (issued % per_tile) % NCH;
            // the buffer is free once every warp of the half has arrived for the item that used it last
            if (issued >= NSTAGE) mbar_wait(&bars[NSTAGE + issued % NSTAGE], ((issued / NSTAGE) - 1) & 1);
            uint64_t* bar = &bars[issued % NSTAGE];
            mbar_expect_tx(bar, CHUNK_BYTES);
            bulk_g2s(tbuf + (size_t)(issued % NSTAGE) * KC * CSH,
                     tt + (((size_t)tile * 2 + half) * NCH + chunk) * KC * CSH, CHUNK_BYTES, bar);
        }
        ++issued;
    };
    for (int i = 0; i < NSTAGE - 1; ++i) issue_next();  // prologue: NSTAGE-1 chunks in flight
    // A operand of the forward product: the histosys coefficients (c1_h, c2_h) of the CTA's 32 points,
    // KS columns (zero padded) in the layout of coef_at, computed here from the batch-major parameters
    // (lanes = consecutive points: coalesced reads) instead of travelling through a global scratch
    for (int i = tid; i < (KS / 2) * MM; i += NTHR) {
        const int h = i / MM, pt = i - h * MM;
        double2 v;
        v.x = v.y = 0.0;
        if (h < pl.H) {
            double d1, d2;
            hf_hs_coeffs(pl.hs_code, w.parsT[(int64_t)pl.hs_par[h] * ld + pt0 + pt], v.x, v.y, d1, d2);
        }
        coefs[coef_at(2 * h, pt)] = v.x;
        coefs[coef_at(2 * h + 1, pt)] = v.y;
    }
    __syncthreads();

    // backward accumulators: this warp's 8 points x all K columns
    double bacc[NCH][NT_BWD][2];
    if (GRAD) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int n = 0; n < NT_BWD; ++n) bacc[c][n][0] = bacc[c][n][1] = 0.0;
    }
    // epilogue state: thread = point `lane`, warp = 4 consecutive bins of every tile
    const int64_t p = pt0 + lane;
    double val = 0.0;      // main-term partial sum of this point over this warp's bins
    double gacc[NPOS];     // per-sample sums of w*bf*base inside the current segment
    double fseg[NPOS];     // scalar-factor products of this point in the current segment
    int gseg = -1;
#pragma unroll
    for (int s = 0; s < NPOS; ++s) gacc[s] = fseg[s] = 0.0;

    for (int tile = 0; tile < ntiles; ++tile) {
        // HF_MMA_TRACE: phase time stamps of the first CTA, [half][tile][forward, epilogue, backward, end]
        const bool tr = trace != nullptr && blockIdx.x == 0 && hw == 0 && lane == 0;
        if (tr) trace[(half * ntiles + tile) * 4 + 0] = clock64();
        // ---------------------------------------------------------------- forward contraction
        {
            double facc[4][2][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) facc[mi][ni][0] = facc[mi][ni][1] = 0.0;
#pragma unroll 1
            for (int chunk = 0; chunk < NCH; ++chunk) {
                issue_next();
                mbar_wait(&bars[consumed % NSTAGE], (consumed / NSTAGE) & 1);
                // m-tile mi, row g = point 16*(mi>>1) + 2g + (mi&1); n-tile ni, column n = cell 2n + ni of
                // the warp's 16: three 16-byte loads feed the eight DMMA of a k-step
                // (row k4*4 + t of the chunk: HF_MMA_ROW puts a lane-constant 4*(t>>1) on top of the stride)
                const double* tb = tbuf + (size_t)(consumed % NSTAGE) * KC * CSH + 16 * hw + 2 * g + 4 * (t >> 1);
                const double* ca = coefs + (size_t)(chunk * KC + t) * MM;  // column chunk*KC + k4*4 + t: (k & 3) == t
                const int sw0 = ((g ^ (2 * t)) << 1), sw1 = (((8 + g) ^ (2 * t)) << 1);
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    const double2 a01 = *reinterpret_cast<const double2*>(ca + k4 * 4 * MM + sw0);
                    const double2 a23 = *reinterpret_cast<const double2*>(ca + k4 * 4 * MM + sw1);
                    const double2 b01 = *reinterpret_cast<const double2*>(tb + (k4 * 4 + t) * CSH);
                    const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double b[2] = {b01.x, b01.y};
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni)
                            dmma884(facc[mi][ni][0], facc[mi][ni][1], a[mi], b[ni]);
                }
                // this warp is done with the buffer (no CTA-wide barrier: warps may run one chunk apart)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[NSTAGE + consumed % NSTAGE]);
                ++consumed;
            }
            // Delta fragments -> shared [point][cell]; with the tile mapping above a lane owns, per
            // m-tile, all four rows of bin 4*warp + t of one point (row = 2j + ni).  The cells still hold
            // w*F of the previous tile until every warp of the half has left its backward product.
            half_barrier(half);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double4 st;
                st.x = facc[mi][0][0];
                st.y = facc[mi][1][0];
                st.z = facc[mi][0][1];
                st.w = facc[mi][1][1];
                const int pt = 16 * (mi >> 1) + 2 * g + (mi & 1);
                *reinterpret_cast<double4*>(&wfs[HF_MMA_WFROW(pt) + 4 * (4 * warp + t)]) = st;
            }
        }
        // ---- the tile's model constants were requested after the previous epilogue: make them and
        //      the Delta tile visible
        cp_async_wait_all();
        half_barrier(half);
        if (tr) trace[(half * ntiles + tile) * 4 + 1] = clock64();
        // observed counts of this point on the warp's 4 bins: straight from global memory (L2), in
        // flight while the rates are formed -- staging them cost 8 KB of shared memory
        double nd4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) nd4[j] = drow[min(tile * BT + 4 * warp + j, NB - 1)];

        // ---------------------------------------------------------------- epilogue
        // positions 0..3 are the histosys rows (Delta in cell[0..3]), 4..7 the other samples.
        // Fast path (the warp's 4 bins lie in one segment): the 4 bins are processed as 4
        // interleaved dependency chains (the log / divide chains are ~1000 cycles each and only
        // two warps share a scheduler); slow path: one bin at a time with a segment switch.
        auto switch_segment = [&](int seg) {
            if (GRAD && gseg >= 0) {
#pragma unroll
                for (int q = 0; q < NPOS; ++q)
                    if (pos_smp[q] >= 0 && gacc[q] != 0.0)
                        atomicAdd(&w.GsegT[((int64_t)pos_smp[q] * pl.NSEG + gseg) * ld + p], gacc[q]);
            }
#pragma unroll
            for (int q = 0; q < NPOS; ++q) {
                gacc[q] = 0.0;
                // scalar factors of this point in the new segment stay in registers
                fseg[q] = (pos_smp[q] >= 0) ? w.FsegT[((int64_t)pos_smp[q] * pl.NSEG + seg) * ld + p] : 0.0;
            }
            gseg = seg;
        };
        const int seg_first = seg_s[4 * warp], seg_last = seg_s[4 * warp + 3];
        // per-bin factors of the warp's 4 bins: none, or (K == 1: staterror / shapesys / shapefactor alone)
        // ONE parameter per bin, shared by the samples that carry it (gmask: bit = sample position);
        // anything else takes the general path below
        bool simple = true;
        int gpar[4] = {-1, -1, -1, -1};
        unsigned gmask[4] = {0u, 0u, 0u, 0u};
        if (HASBF) {
            if (pl.K == 1) {
#pragma unroll
                for (int q = 0; q < NPOS; ++q) {
                    if (pos_smp[q] < 0) continue;
                    const int* r4 = bf_s + pos_smp[q] * BT + 4 * warp;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int par = r4[j];
                        if (par < 0) continue;
                        if (gpar[j] >= 0 && gpar[j] != par) simple = false;
                        gpar[j] = par;
                        gmask[j] |= 1u << q;
                    }
                }
            } else {
                for (int ks = 0; ks < pl.K * S; ++ks) {
                    const int* r4 = bf_s + ks * BT + 4 * warp;
                    simple = simple && ((r4[0] & r4[1] & r4[2] & r4[3]) < 0);
                }
            }
        }
        const bool hasg = HASBF && (gpar[0] >= 0 || gpar[1] >= 0 || gpar[2] >= 0 || gpar[3] >= 0);
        // bins are ordered, so the warp's 4 bins span at most two segments iff the inner ones
        // belong to the first or the last one; each segment gets one (masked) pass
        const bool two_seg = (seg_s[4 * warp + 1] == seg_first || seg_s[4 * warp + 1] == seg_last) &&
                             (seg_s[4 * warp + 2] == seg_first || seg_s[4 * warp + 2] == seg_last);
        // the 4 bins as 4 interleaved dependency chains; WG: some of them carry a per-bin parameter
        auto fast_bins = [&](auto wg_tag) {
            constexpr bool WG = decltype(wg_tag)::value;
            double gam[4];
            if (WG) {  // the four loads are in flight together
#pragma unroll
                for (int j = 0; j < 4; ++j) gam[j] = (gpar[j] >= 0) ? w.parsT[(int64_t)gpar[j] * ld + p] : 1.0;
            }
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
                    const double4 dlt = *reinterpret_cast<const double4*>(wfs + HF_MMA_WFROW(lane) + 4 * bl);
                    const double dl[4] = {dlt.x, dlt.y, dlt.z, dlt.w};
                    // positions without a sample carry fseg = 0 and read nominal row 0: no predicates
                    // (only the rarely used positions 5..7 keep a uniform guard)
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < NPOS; ++q) {
                        if (q < 5 || q < 4 + nother) {
                            double base = nom_s[noff[q] + bl];
                            if (q < 4) base += dl[q];
                            double f = fseg[q];
                            if (WG) f *= ((gmask[j] >> q) & 1u) ? gam[j] : 1.0;
                            acc = fma(f, base, acc);
                        }
                    }
                    lam[j] = acc;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int bl = 4 * warp + j;
                    const bool valid = (tile * BT + bl < NB) && act[j];
                    const double n = nd4[j];
                    if (valid) val += pois_term_fast(n, lam[j], ltab);
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
                        double* cell = wfs + HF_MMA_WFROW(lane) + 4 * bl;
                        const double4 dlt = *reinterpret_cast<const double4*>(cell);
                        const double dl[4] = {dlt.x, dlt.y, dlt.z, dlt.w};
                        double wf[4];
                        double gsum = 0.0;  // d lambda / d gamma of the bin's parameter
#pragma unroll
                        for (int q = 0; q < NPOS; ++q) {
                            if (q < 5 || q < 4 + nother) {
                                double base = nom_s[noff[q] + bl];
                                if (q < 4) base += dl[q];
                                double wq = wgt[j];
                                if (WG) {
                                    const bool on = (gmask[j] >> q) & 1u;
                                    wq *= on ? gam[j] : 1.0;
                                    gsum = fma(on ? fseg[q] : 0.0, base, gsum);
                                }
                                gacc[q] = fma(wq, base, gacc[q]);  // (never flushed for empty positions)
                                if (q < 4) wf[q] = wq * fseg[q];
                            }
                        }
                        double4 st;  // A' operand of the backward product
                        st.x = wf[0]; st.y = wf[1]; st.z = wf[2]; st.w = wf[3];
                        *reinterpret_cast<double4*>(cell) = st;
                        if (WG && gpar[j] >= 0 && wgt[j] != 0.0) {
                            // (a parameter that acts on this bin only: this is the row's one contribution)
                            if (pl.mma_bf_unique) w.gradT[(int64_t)gpar[j] * ld + p] = wgt[j] * gsum;
                            else atomicAdd(&w.gradT[(int64_t)gpar[j] * ld + p], wgt[j] * gsum);
                        }
                    }
                }
            }
        };
        if (simple && two_seg) {
            if (hasg) fast_bins(std::true_type{});
            else fast_bins(std::false_type{});
        } else {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int bl = 4 * warp + j;
            const int b = tile * BT + bl;
            const bool valid = b < NB;
            const int bb = valid ? b : NB - 1;
            const int seg = seg_s[bl];
            double* cell = wfs + HF_MMA_WFROW(lane) + 4 * bl;
            if (seg != gseg) switch_segment(seg);  // warp-uniform: all lanes look at the same bin
            const double4 dlt = *reinterpret_cast<const double4*>(cell);
            double basev[NPOS], bfv[NPOS];
            basev[0] = dlt.x; basev[1] = dlt.y; basev[2] = dlt.z; basev[3] = dlt.w;
#pragma unroll
            for (int q = 4; q < NPOS; ++q) basev[q] = 0.0;
            double lam = 0.0;
#pragma unroll
            for (int q = 0; q < NPOS; ++q) {
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
            const double n = drow[bb];
            if (valid) val += hf_pois_term(n, lam);
            if (GRAD) {
                const double wgt = valid ? ((n == 0.0) ? -1.0 : n / lam - 1.0) : 0.0;
                double wf[4];
#pragma unroll
                for (int q = 0; q < NPOS; ++q) {
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
        half_barrier(half);  // w*F of the half's 16 bins complete, tile constants free
        if (tr) trace[(half * ntiles + tile) * 4 + 2] = clock64();
        if (tile + 1 < ntiles) prefetch_tile(tile + 1);
        if (GRAD) {
#pragma unroll
            for (int chunk = 0; chunk < NCH; ++chunk) {
                issue_next();
                mbar_wait(&bars[consumed % NSTAGE], (consumed / NSTAGE) & 1);
                // warp = (m-tile: points 8*hw+g) x (the half's 64 cells).  The k dimension of a DMMA is
                // four cells in any order both operands agree on: k-step 2j + e takes cells 8j + 2t + e, so
                // one 16-byte load per operand feeds two k-steps (6 loads per 10 DMMA instead of 12)
                const double* tb = tbuf + (size_t)(consumed % NSTAGE) * KC * CSH + HF_MMA_ROW(g) + 2 * t;
                const double* arow = wfs + HF_MMA_WFROW(8 * hw + g) + 2 * t + NCH_CELLS * half;
#pragma unroll
                for (int j = 0; j < NCH_CELLS / 8; ++j) {
                    const double2 a = *reinterpret_cast<const double2*>(arow + 8 * j);
                    double2 b[NT_BWD];
#pragma unroll
                    for (int ni = 0; ni < NT_BWD; ++ni)
                        b[ni] = *reinterpret_cast<const double2*>(tb + (size_t)(8 * ni) * CSH + 8 * j);
#pragma unroll
                    for (int ni = 0; ni < NT_BWD; ++ni) dmma884(bacc[chunk][ni][0], bacc[chunk][ni][1], a.x, b[ni].x);
#pragma unroll
                    for (int ni = 0; ni < NT_BWD; ++ni) dmma884(bacc[chunk][ni][0], bacc[chunk][ni][1], a.y, b[ni].y);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[NSTAGE + consumed % NSTAGE]);
                ++consumed;
            }
        }
        if (tr) trace[(half * ntiles + tile) * 4 + 3] = clock64();
    }

    // ---------------------------------------------------------------- flush
    // The eight warps hold different bins of the same points, the two halves partial sums over
    // different cells of the same (point, column): both are combined through the table buffers of
    // the half (free now: every chunk requested has been consumed) instead of atomics on global
    // memory, so the main term and the histosys gradients leave as plain stores.
    double* xch = tbuf;                                  // [NCH][NT_BWD][2][4 warps][32 lanes] of this half
    double* vals = tbuf + NCH * NT_BWD * 2 * HTHR;       // [4 warps][32 lanes]
    const double* xch_other = tbuf_all + (size_t)(1 - half) * NSTAGE * KC * CSH;
    half_barrier(half);  // every warp of the half is done with the last table chunk
    vals[hw * 32 + lane] = val;
    if (GRAD) {
        if (gseg >= 0) {
#pragma unroll
            for (int q = 0; q < NPOS; ++q)
                if (pos_smp[q] >= 0 && gacc[q] != 0.0)
                    atomicAdd(&w.GsegT[((int64_t)pos_smp[q] * pl.NSEG + gseg) * ld + p], gacc[q]);
        }
        // chunk c is finished by half (c & 1): the other half hands its partial sums over
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk)
            if ((chunk & 1) != half) {
#pragma unroll
                for (int ni = 0; ni < NT_BWD; ++ni) {
                    xch[((chunk * NT_BWD + ni) * 2 + 0) * HTHR + htid] = bacc[chunk][ni][0];
                    xch[((chunk * NT_BWD + ni) * 2 + 1) * HTHR + htid] = bacc[chunk][ni][1];
                }
            }
    }
    __syncthreads();
    if (warp == 0) {  // (constraint terms and constants: hf_finish_value_kernel, at full occupancy --
                      //  folded into this kernel they sat on every CTA's critical path: +0.2 ms per step)
        double v = 0.0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int ww = 0; ww < HTHR / 32; ++ww)
                v += (tbuf_all + (size_t)hh * NSTAGE * KC * CSH + NCH * NT_BWD * 2 * HTHR)[ww * 32 + lane];
        w.mainv[p] = v;
    }
    if (GRAD) {
        // bacc[chunk][ni] = C'[point 8*hw+g][k = chunk*KC + 8ni + 2t, +1] = (T1, T2) sums of one modifier
        const int64_t pw = pt0 + 8 * hw + g;
#pragma unroll
        for (int chunk = 0; chunk < NCH; ++chunk)
            if ((chunk & 1) == half) {
#pragma unroll
                for (int ni = 0; ni < NT_BWD; ++ni) {
                    const int h = (chunk * KC + 8 * ni + 2 * t) >> 1;
                    if (h < pl.H) {
                        const int par = pl.hs_par[h];
                        // alpha of (point, modifier) from the coefficient block: c1 (code4p), c1 + c2 (code0)
                        const double c1s = coefs[coef_at(2 * h, 8 * hw + g)], c2s = coefs[coef_at(2 * h + 1, 8 * hw + g)];
                        const double alpha = (pl.hs_code == HF_HS_CODE4P) ? c1s : c1s + c2s;
                        double c1, c2, d1, d2;
                        hf_hs_coeffs(pl.hs_code, alpha, c1, c2, d1, d2);
                        const double s0 = bacc[chunk][ni][0] + xch_other[((chunk * NT_BWD + ni) * 2 + 0) * HTHR + htid];
                        const double s1 = bacc[chunk][ni][1] + xch_other[((chunk * NT_BWD + ni) * 2 + 1) * HTHR + htid];
                        const double gh = d1 * s0 + d2 * s1;
                        if (pl.mma_hs_unique) w.gradT[(int64_t)par * ld + pw] = gh;  // the row's only writer
                        else atomicAdd(&w.gradT[(int64_t)par * ld + pw], gh);
                    }
                }
            }
    }
}

// table of fast_log: (1/c_i rounded, -ln of that rounded value) -- uploaded once per device
int upload_logtab() {
    static bool done[64] = {};
    int dev = 0;
    HF_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return HF_OK;
    double2 h[LOGTAB_N];
    for (int i = 0; i < LOGTAB_N; ++i) {
        const long double c = i < 64 ? 0.75L + (i + 0.5L) / 256.0L : 1.0L + (i - 64 + 0.5L) / 128.0L;
        h[i].x = (double)(1.0L / c);
        h[i].y = (double)(-logl((long double)h[i].x));
    }
    HF_CUDA_OK(cudaMemcpyToSymbol(g_logtab, h, sizeof(h)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return HF_OK;
}

template <int NCH>
int launch(hf_plan* plan, int64_t B, int64_t ld, const double* data, int per_point,
           const int* row_idx, bool grad, cudaStream_t st) {
    const DevPlan& pl = plan->d;
    const int KS = hf_mma_ks(pl.H);
    size_t smem = sizeof(double) * ((size_t)MM * KS + 2 * NSTAGE * KC * CSH + MM * CS) + 128 + 8 * MM + 128 +
                  sizeof(double) * (SMAX * BT) +
                  sizeof(int) * (BT + SMAX + (size_t)pl.K * pl.S * BT) + 64;
    smem = (smem + 15) / 16 * 16;
    const int ltab_off = (int)smem;
    smem += sizeof(double2) * LOGTAB_N;
    if (smem > 227 * 1024) return HF_ERR_UNSUPPORTED;  // (many histosys / per-bin factor kinds: bin-major kernel)
    const dim3 grid((unsigned)(ld / MM));
    if (const int rc_tab = upload_logtab()) return rc_tab;
    long long* trace = nullptr;
    static const bool want_trace = getenv("HF_MMA_TRACE") && atoi(getenv("HF_MMA_TRACE")) != 0;
    if (want_trace) {
        HF_CUDA_OK(cudaMalloc(&trace, sizeof(long long) * 8 * pl.mma_ntiles));
        HF_CUDA_OK(cudaMemsetAsync(trace, 0, sizeof(long long) * 8 * pl.mma_ntiles, st));
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (plan->timing) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
#define HF_MMA_GO2(G_, BF_, NP_)                                                                       \
    {                                                                                               \
        auto k = hf_main_mma_kernel<NCH, G_, BF_, NP_>;                                               \
        HF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<grid, NTHR, smem, st>>>(pl, plan->w, B, ld, data, per_point, row_idx, KS, trace, ltab_off); \
    }
#define HF_MMA_GO(G_, BF_) \
    { if (pl.S - pl.R <= 1) HF_MMA_GO2(G_, BF_, 5) else HF_MMA_GO2(G_, BF_, SMAX) }
    if (grad) {
        if (pl.K > 0) HF_MMA_GO(true, true) else HF_MMA_GO(true, false)
    } else {
        if (pl.K > 0) HF_MMA_GO(false, true) else HF_MMA_GO(false, false)
    }
#undef HF_MMA_GO
#undef HF_MMA_GO2
    if (plan->timing) {
        cudaEventRecord(e1, st);
        plan->ev.push_back(e0);
        plan->ev.push_back(e1);
    }
    HF_CUDA_OK(cudaGetLastError());
    plan->launches++;
    if (trace) {  // diagnostic: cycles of the phases of CTA 0, relative to the first stamp
        std::vector<long long> h(8 * (size_t)pl.mma_ntiles);
        HF_CUDA_OK(cudaStreamSynchronize(st));
        HF_CUDA_OK(cudaMemcpy(h.data(), trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        long long t0 = h[0];
        for (long long v : h) if (v && v < t0) t0 = v;
        for (int tile = 0; tile < pl.mma_ntiles; ++tile) {
            fprintf(stderr, "mma-trace tile %2d", tile);
            for (int hh = 0; hh < 2; ++hh) {
                const long long* r = &h[((size_t)hh * pl.mma_ntiles + tile) * 4];
                fprintf(stderr, " | half %d F@%7lld E@%7lld B@%7lld end@%7lld (F %5lld E %5lld B %5lld)", hh, r[0] - t0,
                        r[1] - t0, r[2] - t0, r[3] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2]);
            }
            fprintf(stderr, "\n");
        }
    }
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
