// Device code of the per-fit minimiser (see hf_fit_solo.cu), compiled once per CTA size: the including
// file defines SOLO_NT (threads per CTA) and SOLO_NS (the namespace of this instance).
namespace SOLO_NS {
constexpr int NT = SOLO_NT;

// (a, b) of the t-th entry of a lower triangle enumerated row by row, b <= a, from the table the CTA
// built at start-up (tri_decode costs a square root and two loops; the DMMA tile loops ask ~1500 times
// per Newton iteration)
__device__ __forceinline__ void tile_decode(const Ctx& c, int t, int& a, int& b) {
    const unsigned v = c.tdec[t];
    a = (int)(v >> 8);
    b = (int)(v & 255u);
}

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
__device__ __forceinline__ double solo_sweep(const DevPlan& pl, const SoloArgs& a, const Ctx& c, const double* xin, bool fisher,
                             const double* __restrict__ drow, double dconst) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int P = pl.P, NB = pl.NB, S = pl.S, NSEG = pl.NSEG, H = pl.H, R = pl.R, K = pl.K;
    const int Hpad = a.Hpad, nj = a.nj, Hc = a.Hc, npart = a.npart;
    const int ng = a.ng, nl = a.nl, npk = a.npk, ngs = a.ngs;
    // bins per tile: 32 where the histosys phases own the other warps; the whole CTA (one bin per thread)
    // for plans without histosys -- C3's 100 bins are ONE tile, 3 barriers per sweep instead of 12
    constexpr int TBv = (MAXT == 0) ? NT : TB;
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
        if (!(MAXT == 0 && a.L.lowrank))
            for (int i = tid; i < nl * ngs; i += NT) c.C[i] = 0.0;
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
            // (stride 6: the pair (c1, c2) is one 16-byte shared-memory load in the delta loop)
            c.hc[6 * h] = c1; c.hc[6 * h + 1] = c2; c.hc[6 * h + 2] = d1; c.hc[6 * h + 3] = d2; c.hc[6 * h + 4] = dd;
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
    double Nacc[SMX], NaccR[SMX];
    int srow[SMX];  // sample of every histosys row
#pragma unroll
    for (int s = 0; s < SMX; ++s) {
        Nacc[s] = NaccR[s] = 0.0;
        srow[s] = (MAXT > 0 && s < R) ? pl.hs_row_sample[s] : -1;
    }
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
    if (h_thread) { dc1 = c.hc[6 * my_h + 2]; dc2 = c.hc[6 * my_h + 3]; ddc2 = c.hc[6 * my_h + 4]; }
    double vacc = 0.0;
    const int nsplit = (MAXT > 0) ? max(1, (NT / 32) / max(R, 1)) : 1;

    for (int b0 = 0; b0 < NB; b0 += TBv) {
        // ---- (1) histosys delta of every (row, bin) of the tile
        if (MAXT > 0) {
            // every warp takes one (row, slice of the modifiers): many independent loads in flight
            if (warp < R * nsplit) {
                const int row = warp % R, part = warp / R;
                const int hchunk = (H + nsplit - 1) / nsplit;
                const int h0 = part * hchunk, h1 = min(H, h0 + hchunk);
                const int b = min(b0 + lane, NB - 1);
                const double* p1 = pl.t1 + (int64_t)row * NB + b;
                const double* p2 = pl.t2 + (int64_t)row * NB + b;
                const int64_t hstride = (int64_t)R * NB;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
                for (int h = h0; h < h1; ++h) {
                    const double2 cc = *reinterpret_cast<const double2*>(c.hc + 6 * h);
                    s0 = fma(cc.x, __ldg(p1 + h * hstride), s0);
                    s1 = fma(cc.y, __ldg(p2 + h * hstride), s1);
                }
                c.base[warp * TB + lane] = s0 + s1;
            }
            __syncthreads();
        }
        // ---- (2a) rates, value, weights, per-bin parameter
        if (tid < TBv) {
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
                    if (MAXT > 0 && hr >= 0)
                        for (int part = 0; part < nsplit; ++part) bs += c.base[(part * R + hr) * TB + bl];
                    double bf = 1.0;
                    for (int k = 0; k < K; ++k) {
                        const int p = pl.bf_par[((int64_t)k * S + s) * NB + bb];
                        if (p >= 0) bf *= c.xs[p];
                    }
                    const double fb = c.Fq[s * NSEG + seg] * bf;
                    rr[s] = valid ? fb * bs : 0.0;
                    if (FULL) {
                        if (MAXT > 0) c.FB[s * TBv + bl] = valid ? fb : 0.0;
                        c.rS[s * TBv + bl] = rr[s];
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
            for (int bl = 0; bl < TBv && b0 + bl < NB; ++bl) {
                const int seg = c.segt[bl];
                if (seg != mcur) {
                    if (mcur >= 0) {
                        if (pair_thread) c.SegM[(int64_t)mcur * S * S + tid] += macc;
                        else c.SegG[(int64_t)mcur * S + gs] += macc;
                    }
                    macc = 0.0;
                    mcur = seg;
                }
                if (pair_thread) macc = fma(c.Wt[bl] * c.rS[ps * TBv + bl], c.rS[ps2 * TBv + bl], macc);
                else macc = fma(c.wgt[bl], c.rS[gs * TBv + bl], macc);
            }
        }
        // ---- (3) histosys columns of the tile: a_b[h], the S x H sums N_c, per-bin parameter rows
        if (MAXT > 0) {
            if (h_thread) {
                double nv1[SMX], nv2[SMX];  // table values of the next bin, fetched while this one is worked on
                auto fetch = [&](int bl_) {
                    const int b_ = b0 + bl_;
#pragma unroll
                    for (int r = 0; r < SMX; ++r) {
                        nv1[r] = nv2[r] = 0.0;
                        if (r < R && bl_ < TB && b_ < NB) {
                            const int64_t o = ((int64_t)r * NB + b_) * pl.Hp + my_h;
                            nv1[r] = __ldg(pl.t1t + o);
                            nv2[r] = __ldg(pl.t2t + o);
                        }
                    }
                };
                fetch(my_part);
                for (int bl = my_part; bl < TB; bl += npart) {
                    const int b = b0 + bl;
                    double cv1[SMX], cv2[SMX];
#pragma unroll
                    for (int r = 0; r < SMX; ++r) { cv1[r] = nv1[r]; cv2[r] = nv2[r]; }
                    fetch(bl + npart);
                    if (b >= NB) { c.Y[bl * Hpad + my_h] = 0.0; continue; }
                    const int seg = c.segt[bl];
                    if (seg != ncur) {
                        if (ncur >= 0) {
#pragma unroll
                            for (int s = 0; s < SMX; ++s)
                                if (s < S) {
                                    double v = Nacc[s];
#pragma unroll
                                    for (int r = 0; r < SMX; ++r)
                                        if (r < R && srow[r] == s) v += NaccR[r];
                                    if (v != 0.0) c.SegN[(((int64_t)ncur * S + s) * npart + my_part) * Hc + my_h] += v;
                                }
                        }
#pragma unroll
                        for (int s = 0; s < SMX; ++s) Nacc[s] = NaccR[s] = 0.0;
                        ncur = seg;
                    }
                    double A[SMX], asum = 0.0, ddl = 0.0;
#pragma unroll
                    for (int r = 0; r < SMX; ++r) {
                        A[r] = 0.0;
                        if (r < R) {
                            const double v1 = cv1[r], v2 = cv2[r];
                            const double fb = c.FB[srow[r] * TB + bl];
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
                        if (r < R) NaccR[r] = fma(w2, A[r], NaccR[r]);  // (joins the row's sample when the sums are flushed)
                    ddacc = fma(w2, ddl, ddacc);
                    const int gp = c.lgp[bl];
                    if (gp >= 0) {
                        const unsigned mk = c.lmk[bl];
                        double sumA = 0.0;
#pragma unroll
                        for (int r = 0; r < SMX; ++r)
                            if (r < R && ((mk >> srow[r]) & 1u)) sumA += A[r];
                        c.C[(int64_t)a.lpos[gp] * ngs + a.gpos[pl.hs_par[my_h]]] = c.lwj[bl] * asum + c.lw2[bl] * sumA;
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
                    tile_decode(c, tile, tj, ti);  // ti <= tj : row tile, column tile
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
                if (s < S) {
                    double v = Nacc[s];
#pragma unroll
                    for (int r = 0; r < SMX; ++r)
                        if (r < R && srow[r] == s) v += NaccR[r];
                    if (v != 0.0) c.SegN[(((int64_t)ncur * S + s) * npart + my_part) * Hc + my_h] += v;
                }
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
                tile_decode(c, tile, tj, ti);
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
    double* D2 = DL + S * nj;          // [S][nj]
    double* U0 = D2 + S * nj;          // [S][nj]   M_c DL
    double* Msm = U0 + ((S * nj + 1) & ~1);  // [S][S]
    double* G2sm = Msm + SMX * SMX;    // [S]
    double* Ns = G2sm + SMX;           // [S][H]   N_c of the channel
    double* Lw = Ns + ((S * H + 1) & ~1);  // [locals of the channel][S]
    const double hw = fisher ? 0.0 : 1.0;  // second-order terms enter the exact Hessian only
    lap(2);
    if (MAXT >= 8 && a.bigch) {
        // Many scalar-factor parameters: the two sums over (channel, sample),
        //     scalar x scalar    sum_{c,s} DL_c[s][j1] U1_c[s][j2] ,  U1 = (M_c + diag G_c) DL_c
        //     histosys x scalar  sum_{c,s} N_c[s][h] DL_c[s][j]
        // are products with inner dimension NSEG * S: 8 x 8 output tiles on the fp64 tensor-core path,
        // accumulators in registers across the channels, written to S_gg once at the end.
        constexpr int NSCW = (120 + NT / 32 - 1) / (NT / 32), NHSW = (225 + NT / 32 - 1) / (NT / 32);
        const int njs = a.njs;
        double* DLb = c.Y;                  // [8][njs]  rows >= S and columns >= nj stay zero
        double* U1 = DLb + 8 * njs;         // [8][njs]
        double* D2b = U1 + 8 * njs;         // [S][nj]
        double* Mb = D2b + ((S * nj + 1) & ~1);
        double* Gb = Mb + SMX * SMX;
        double* Nb = Gb + SMX;              // [8][Hpad]
        double* Lb = Nb + 8 * Hpad;         // [locals of the channel][S]
        double* d2sum = Lb + ((a.max_lc * S + 1) & ~1);  // [nj]
        double accsc[NSCW][2], acchs[NHSW][2];
#pragma unroll
        for (int q = 0; q < NSCW; ++q) accsc[q][0] = accsc[q][1] = 0.0;
#pragma unroll
        for (int q = 0; q < NHSW; ++q) acchs[q][0] = acchs[q][1] = 0.0;
        const int nt8j = (nj + 7) / 8, nt8h = (H + 7) / 8;
        const int ntsc = nt8j * (nt8j + 1) / 2, nths = nt8h * nt8j;
        for (int i = tid; i < 16 * njs; i += NT) DLb[i] = 0.0;   // DL and U1
        for (int i = tid; i < 8 * Hpad; i += NT) Nb[i] = 0.0;
        for (int i = tid; i < nj; i += NT) d2sum[i] = 0.0;
        __syncthreads();
        for (int ch = 0; ch < NSEG; ++ch) {
            for (int i = tid; i < S * njs; i += NT) DLb[i] = 0.0;
            for (int i = tid; i < S * nj; i += NT) D2b[i] = 0.0;
            if (tid < S * S) Mb[tid] = c.SegM[(int64_t)ch * S * S + tid];
            if (tid < S) Gb[tid] = c.SegG[(int64_t)ch * S + tid];
            const int li0 = pl.ah_bl_ptr[pl.ah_seg_first[ch]], li1 = pl.ah_bl_ptr[pl.ah_seg_first[ch + 1]];
            for (int i = tid; i < S * H; i += NT) {  // N_c summed over the thread groups that built it
                const int s_ = i / H, h = i - s_ * H;
                double nsum = 0.0;
                for (int k = 0; k < npart; ++k) nsum += c.SegN[(((int64_t)ch * S + s_) * npart + k) * Hc + h];
                Nb[s_ * Hpad + h] = nsum;
            }
            for (int i = tid; i < (li1 - li0) * S; i += NT) Lb[i] = c.LOCw[(int64_t)li0 * S + i];
            __syncthreads();
            for (int s_ = 0; s_ < S; ++s_) {
                const int q = s_ * NSEG + ch;
                for (int e = pl.sf_ptr[q] + tid; e < pl.sf_ptr[q + 1]; e += NT) {
                    double dl, d2;
                    sf_dlogs(pl.sf_kind[e], pl.sf_coef + 8 * (int64_t)e, c.xs[pl.sf_par[e]], dl, d2);
                    const int j = pl.ah_sf_j[e];
                    DLb[s_ * njs + j] = dl;
                    D2b[s_ * nj + j] = d2;
                }
            }
            __syncthreads();
            for (int i = tid; i < S * nj; i += NT) {
                const int s_ = i / nj, j = i - s_ * nj;
                double u = hw * Gb[s_] * DLb[s_ * njs + j];
                for (int s2 = 0; s2 < S; ++s2) u = fma(Mb[s_ * S + s2], DLb[s2 * njs + j], u);
                U1[s_ * njs + j] = u;
            }
            for (int j = tid; j < nj; j += NT) {
                double gj = 0.0, dj = 0.0;
                for (int s_ = 0; s_ < S; ++s_) {
                    gj = fma(DLb[s_ * njs + j], Gb[s_], gj);
                    dj = fma(hw * Gb[s_], D2b[s_ * nj + j], dj);
                }
                if (gj != 0.0) c.g[pl.ah_sfp_list[j]] += gj;
                d2sum[j] += dj;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < NSCW; ++q) {
                const int tile = warp + (NT / 32) * q;
                if (tile < ntsc) {
                    int ti, tj;
                    tile_decode(c, tile, ti, tj);  // tj <= ti
                    const double* pa = DLb + t4 * njs + 8 * ti + g8;
                    const double* pb = U1 + t4 * njs + 8 * tj + g8;
                    dmma884(accsc[q][0], accsc[q][1], pa[0], pb[0]);
                    dmma884(accsc[q][0], accsc[q][1], pa[4 * njs], pb[4 * njs]);
                }
            }
#pragma unroll
            for (int q = 0; q < NHSW; ++q) {
                const int tile = warp + (NT / 32) * q;
                if (tile < nths) {
                    const int th = tile / nt8j, tj = tile - th * nt8j;
                    const double* pa = Nb + t4 * Hpad + 8 * th + g8;
                    const double* pb = DLb + t4 * njs + 8 * tj + g8;
                    dmma884(acchs[q][0], acchs[q][1], pa[0], pb[0]);
                    dmma884(acchs[q][0], acchs[q][1], pa[4 * Hpad], pb[4 * njs]);
                }
            }
            // per-bin parameters of this channel x scalar
            for (int e = tid; e < (li1 - li0) * nj; e += NT) {
                const int il = e / nj, j = e - il * nj;
                double v = 0.0;
                for (int s_ = 0; s_ < S; ++s_) v = fma(DLb[s_ * njs + j], Lb[il * S + s_], v);
                if (v != 0.0) atomicAdd(&c.C[(int64_t)a.lpos[pl.ah_bl_par[li0 + il]] * ngs + a.gpos[pl.ah_sfp_list[j]]], v);  // (RED: no wait for the old value)
            }
            __syncthreads();
        }
        // accumulators -> S_gg (each unordered pair of scalar-factor parameters once)
#pragma unroll
        for (int q = 0; q < NSCW; ++q) {
            const int tile = warp + (NT / 32) * q;
            if (tile < ntsc) {
                int ti, tj;
                tile_decode(c, tile, ti, tj);
                const int j1 = 8 * ti + g8;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int j2 = 8 * tj + 2 * t4 + cc;
                    if (j1 < nj && j2 <= j1 && accsc[q][cc] != 0.0)
                        c.S[trisym(a.gpos[pl.ah_sfp_list[j1]], a.gpos[pl.ah_sfp_list[j2]])] += accsc[q][cc];
                }
            }
        }
        __syncthreads();
        for (int j = tid; j < nj; j += NT) {
            const int gi = a.gpos[pl.ah_sfp_list[j]];
            c.S[tri(gi, gi)] += d2sum[j];
        }
        __syncthreads();
        // (a parameter may be histosys and scalar factor at once: atomics keep coinciding targets exact)
#pragma unroll
        for (int q = 0; q < NHSW; ++q) {
            const int tile = warp + (NT / 32) * q;
            if (tile < nths) {
                const int th = tile / nt8j, tj = tile - th * nt8j;
                const int h = 8 * th + g8;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int j = 8 * tj + 2 * t4 + cc;
                    if (h < H && j < nj && acchs[q][cc] != 0.0) {
                        const int p1 = pl.hs_par[h], p2 = pl.ah_sfp_list[j];
                        atomicAdd(&c.S[trisym(a.gpos[p1], a.gpos[p2])], (p1 == p2) ? 2.0 * acchs[q][cc] : acchs[q][cc]);
                    }
                }
            }
        }
        __syncthreads();
    } else if (nj > 0) {
        const int npair = nj * (nj + 1) / 2;
        const bool lowrank = MAXT == 0 && a.L.lowrank;
        for (int ch = 0; ch < NSEG; ++ch) {
            // (low-rank form of the per-bin block: the channel's dlog f table is kept for the solve)
            if (lowrank) DL = c.DLall + ch * S * nj;
            for (int i = tid; i < S * nj; i += NT) DL[i] = 0.0;
            for (int i = tid; i < S * nj; i += NT) D2[i] = 0.0;
            if (tid < S * S) Msm[tid] = c.SegM[(int64_t)ch * S * S + tid];
            if (tid < S) G2sm[tid] = c.SegG[(int64_t)ch * S + tid];
            const int li0 = pl.ah_bl_ptr[pl.ah_seg_first[ch]], li1 = pl.ah_bl_ptr[pl.ah_seg_first[ch + 1]];
            if (MAXT > 0)
                for (int i = tid; i < S * H; i += NT) {  // N_c summed over the thread groups that built it
                    const int s_ = i / H, h = i - s_ * H;
                    double nsum = 0.0;
                    for (int k = 0; k < npart; ++k) nsum += c.SegN[(((int64_t)ch * S + s_) * npart + k) * Hc + h];
                    Ns[i] = nsum;
                }
            if (a.L.small) Lw = c.LOCw + (int64_t)li0 * S;  // (already in shared memory)
            else
                for (int i = tid; i < (li1 - li0) * S; i += NT) Lw[i] = c.LOCw[(int64_t)li0 * S + i];
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
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j], Ns[s * H + h], v);
                    if (v != 0.0) {
                        const int p1 = pl.hs_par[h], p2 = pl.ah_sfp_list[j];
                        atomicAdd(&c.S[trisym(a.gpos[p1], a.gpos[p2])], (p1 == p2) ? 2.0 * v : v);
                    }
                }
            }
            // per-bin parameters of this channel x scalar
            if (!lowrank) {
                for (int e = tid; e < (li1 - li0) * nj; e += NT) {
                    const int il = e / nj, j = e - il * nj;
                    double v = 0.0;
                    for (int s = 0; s < S; ++s) v = fma(DL[s * nj + j], Lw[il * S + s], v);
                    if (v != 0.0) c.C[(int64_t)a.lpos[pl.ah_bl_par[li0 + il]] * ngs + a.gpos[pl.ah_sfp_list[j]]] += v;
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
// Cholesky of the packed lower triangle S (n x n) in shared memory; the diagonal keeps the pivots,
// linv holds 1 / L_kk, the strict lower triangle the factor.  Returns false on a non-positive pivot.
// ---------------------------------------------------------------------------------------------
// small systems: right-looking, column by column
__device__ __forceinline__ bool chol_columns(double* S, double* linv, int n) {
    const int tid = threadIdx.x;
    const int TW = 16, TH = NT / TW;
    const int tx = tid % TW, ty = tid / TW;
    for (int j = 0; j < n; ++j) {
        const double piv = S[tri(j, j)];
        if (!(piv > 0.0) || !isfinite(piv)) return false;
        const double inv = 1.0 / sqrt(piv);
        for (int i = j + 1 + tid; i < n; i += NT) S[tri(i, j)] *= inv;
        if (tid == 0) linv[j] = inv;
        __syncthreads();
        for (int i = j + 1 + ty; i < n; i += TH) {
            const double lij = S[tri(i, j)];
            for (int k = j + 1 + tx; k <= i; k += TW) S[tri(i, k)] = fma(-lij, S[tri(k, j)], S[tri(i, k)]);
        }
        __syncthreads();
    }
    return true;
}

// systems of up to 32 unknowns: ONE warp, lane = row, values in registers, no CTA barrier inside
// L y = r, L^T d = y for the same systems; the solution replaces r
__device__ __forceinline__ void subst_warp(const double* S, const double* linv, double* r, int n) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) {
        double ri = (lane < n) ? r[lane] : 0.0, zi = 0.0, di = 0.0;
        for (int k = 0; k < n; ++k) {
            const double zk = __shfl_sync(0xffffffffu, ri, k) * linv[k];
            if (lane == k) zi = zk;
            if (lane > k && lane < n) ri = fma(-S[tri(lane, k)], zk, ri);
        }
        for (int k = n - 1; k >= 0; --k) {
            const double dk = __shfl_sync(0xffffffffu, zi, k) * linv[k];
            if (lane == k) di = dk;
            if (lane < k) zi = fma(-S[tri(k, lane)], dk, zi);
        }
        if (lane < n) r[lane] = di;
    }
    __syncthreads();
}

// large systems: panels of 8 columns; inside a panel one thread per row updates the <= 7 remaining
// panel columns of its row, the trailing matrix takes the whole panel at once as 8 x 8 tiles on the
// fp64 tensor-core path (two DMMA.8x8x4 per tile)
__device__ __forceinline__ bool chol_blocked(const Ctx& c, double* S, double* linv, int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int nt = (n + 7) / 8;
    for (int p = 0; p < nt; ++p) {
        const int j0 = 8 * p, j1 = min(n, j0 + 8);
        for (int j = j0; j < j1; ++j) {
            const double piv = S[tri(j, j)];
            if (!(piv > 0.0) || !isfinite(piv)) return false;
            const double inv2 = 1.0 / piv;
            if (tid == 0) linv[j] = 1.0 / sqrt(piv);
            // (column j stays unscaled until the panel is done: rows read each other's entries of it)
            for (int i = j + 1 + tid; i < n; i += NT) {
                double* row = S + tri(i, 0);
                const double m = row[j] * inv2;
                const int kend = min(i, j1 - 1);
                for (int k = j + 1; k <= kend; ++k) row[k] = fma(-m, S[tri(k, j)], row[k]);
            }
            __syncthreads();
        }
        for (int i = j0 + 1 + tid; i < n; i += NT) {
            double* row = S + tri(i, 0);
            const int kend = min(i - 1, j1 - 1);
            for (int k = j0; k <= kend; ++k) row[k] *= linv[k];
        }
        __syncthreads();
        // trailing matrix: tile (ti, tj), ti >= tj > p:  S -= L[ti rows][panel] L[tj rows][panel]^T
        const int ntr = nt - p - 1, ntiles = ntr * (ntr + 1) / 2;
        for (int t = warp; t < ntiles; t += NT / 32) {
            int ta, tb;
            tile_decode(c, t, ta, tb);
            const int ti = p + 1 + ta, tj = p + 1 + tb;
            const int i = 8 * ti + g8, k0 = 8 * tj + 2 * t4, ib = 8 * tj + g8;
            const bool ok0 = i < n && k0 <= i, ok1 = i < n && k0 + 1 <= i;
            double d0 = ok0 ? S[tri(i, k0)] : 0.0, d1 = ok1 ? S[tri(i, k0 + 1)] : 0.0;
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const int kk = j0 + 4 * st + t4;
                const double av = (i < n && kk < j1) ? -S[tri(i, kk)] : 0.0;
                const double bv = (ib < n && kk < j1) ? S[tri(ib, kk)] : 0.0;
                dmma884(d0, d1, av, bv);
            }
            if (ok0) S[tri(i, k0)] = d0;
            if (ok1) S[tri(i, k0 + 1)] = d1;
        }
        __syncthreads();
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Damped Newton direction from the structured Hessian: d in c.d, returns false when no shift made the
// reduced matrix positive definite.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool solo_solve(const DevPlan& pl, const SoloArgs& a, const Ctx& c, const uint8_t* __restrict__ mask,
                           double damp, int& attempts_out, double& edm_out, double& step_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int P = pl.P, ng = a.ng, nl = a.nl, npk = a.npk, ngs = a.ngs;
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
    long long t_mark = clock64();
    auto lap = [&](int slot) {
        if (tid == 0) { const long long now = clock64(); c.prof[slot] += now - t_mark; t_mark = now; }
    };
    for (int attempt = 0; attempt < 40 && !ok; ++attempt) {
        ++attempts;
        bool bad = false;
        if (tid == 0) c.prof[12] += 1;
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
            if (!bad && nl > 0 && a.L.lowrank) {
                // No histosys: row l of C is DL_c^T w_l (w_l: the S weights of the parameter's bin), so the
                // eliminated block is sum_c DL_c^T Q_c DL_c with Q_c = sum_{l in c} w_l w_l^T / (H_ll + tau):
                // an S x S matrix per channel instead of an n_local x n_global one
                const int S = pl.S, NSEG = pl.NSEG, nj = a.nj, nbl = pl.ah_nbl;
                const int per = S * S + S;
                for (int e = tid; e < NSEG * per; e += NT) {
                    const int ch = e / per, k = e - ch * per;
                    const int i0 = pl.ah_bl_ptr[pl.ah_seg_first[ch]], i1 = pl.ah_bl_ptr[pl.ah_seg_first[ch + 1]];
                    double accv = 0.0;
                    if (k < S * S) {
                        const int s1 = k / S, s2 = k - s1 * S;
                        for (int i = i0; i < i1; ++i) accv = fma(c.dinv[c.bl2l[i]] * c.LOCw[i * S + s1], c.LOCw[i * S + s2], accv);
                        c.Qc[ch * S * S + k] = accv;
                    } else {
                        const int s1 = k - S * S;
                        for (int i = i0; i < i1; ++i) accv = fma(c.gl[c.bl2l[i]], c.LOCw[i * S + s1], accv);
                        c.vq[ch * S + s1] = accv;
                    }
                }
                __syncthreads();
                for (int e = tid; e < NSEG * S * nj; e += NT) {
                    const int cs = e / nj, j = e - cs * nj, ch = cs / S;
                    double accv = 0.0;
                    for (int s2 = 0; s2 < S; ++s2) accv = fma(c.Qc[cs * S + s2], c.DLall[(ch * S + s2) * nj + j], accv);
                    c.Tq[e] = accv;
                }
                __syncthreads();
                for (int e = tid; e < nj * (nj + 1) / 2; e += NT) {
                    const int j2 = c.pj2[e], j1 = c.pj1[e];  // j1 <= j2
                    double accv = 0.0;
                    for (int cs = 0; cs < NSEG * S; ++cs) accv = fma(c.DLall[cs * nj + j1], c.Tq[cs * nj + j2], accv);
                    c.S[c.ppos[e]] -= accv;
                }
                for (int j = tid; j < nj; j += NT) {
                    double accv = 0.0;
                    for (int cs = 0; cs < NSEG * S; ++cs) accv = fma(c.DLall[cs * nj + j], c.vq[cs], accv);
                    c.r[c.gposj[j]] += accv;
                }
                (void)nbl;
            } else if (!bad && nl > 0) {
                if (a.L.small) {
                    // scalar form: every thread owns entries of the packed triangle
                    for (int e = tid; e < npk; e += NT) {
                        int i, j;
                        tri_decode(e, i, j);
                        double accv = 0.0;
#pragma unroll 4
                        for (int l = 0; l < nl; ++l) accv = fma(c.C[l * ngs + i] * c.dinv[l], c.C[l * ngs + j], accv);
                        c.S[e] -= accv;
                    }
                    for (int i = tid; i < ng; i += NT) {
                        double accv = 0.0;
                        for (int l = 0; l < nl; ++l) accv = fma(c.C[l * ngs + i], c.gl[l], accv);
                        c.r[i] += accv;
                    }
                } else {
                    // S -= sum_l (C_l sqrt(d_l)) (C_l sqrt(d_l))^T as 8 x 8 tiles on the fp64 tensor-core
                    // path; rows of C are staged (scaled, zero padded) through the Y region
                    double* W = c.Y;
                    double* sq = c.gpart;  // [ltile] sqrt(d_l) g_l of the staged rows
                    const int ltile = a.L.ltile, wst = a.L.wstride;
                    const int nt8 = (ng + 7) / 8, ntiles = nt8 * (nt8 + 1) / 2;
                    for (int l0 = 0; l0 < nl; l0 += ltile) {
                        const int nt = min(ltile, nl - l0);
                        __syncthreads();
                        for (int e = tid; e < ltile * wst; e += NT) {
                            const int t = e / wst, k = e - t * wst;
                            double v = 0.0;
                            if (t < nt && k < ng) v = c.C[(int64_t)(l0 + t) * ngs + k] * sqrt(c.dinv[l0 + t]);
                            W[e] = v;
                        }
                        if (tid < ltile) sq[tid] = (tid < nt) ? sqrt(c.dinv[l0 + tid]) * c.g[a.llist[l0 + tid]] : 0.0;
                        __syncthreads();
                        for (int t = warp; t < ntiles; t += NT / 32) {
                            int ti, tj;
                            tile_decode(c, t, ti, tj);  // tj <= ti
                            const int i = 8 * ti + g8, k0 = 8 * tj + 2 * t4;
                            const bool ok0 = i < ng && k0 <= i, ok1 = i < ng && k0 + 1 <= i;
                            double d0 = ok0 ? c.S[tri(i, k0)] : 0.0, d1 = ok1 ? c.S[tri(i, k0 + 1)] : 0.0;
                            const double* pa = W + t4 * wst + 8 * ti + g8;
                            const double* pb = W + t4 * wst + 8 * tj + g8;
                            for (int ks = 0; ks < ltile; ks += 4) dmma884(d0, d1, -pa[ks * wst], pb[ks * wst]);
                            if (ok0) c.S[tri(i, k0)] = d0;
                            if (ok1) c.S[tri(i, k0 + 1)] = d1;
                        }
                        for (int i = tid; i < ng; i += NT) {
                            double accv = 0.0;
                            for (int t = 0; t < nt; ++t) accv = fma(W[t * wst + i], sq[t], accv);
                            c.r[i] += accv;
                        }
                    }
                }
            }
            __syncthreads();
            // parameters of the active set (fixed, or on a bound with the gradient pushing outward)
            // keep their value: identity rows with a zero right-hand side
            for (int i = 0; i < ng; ++i)
                if (c.actg[i])
                    for (int k = tid; k < ng; k += NT) c.S[trisym(i, k)] = (i == k) ? 1.0 : 0.0;
            for (int i = tid; i < ng; i += NT)
                if (c.actg[i]) c.r[i] = 0.0;
            __syncthreads();
            if (!bad) {
                for (int e = tid; e < npk; e += NT) c.Sred[e] = c.S[e];
                have_copy = true;
            }
            __syncthreads();
        }
        lap(8);
        // (a one-warp Cholesky, lane = row, was measured slower than the 2-barriers-per-column form on
        //  C3's 20 x 20 system -- its trailing update is a serial chain per lane; the substitutions do gain)
        if (!bad) bad = !((ng > 48) ? chol_blocked(c, c.S, c.linv, ng) : chol_columns(c.S, c.linv, ng));
        lap(9);
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
    __syncthreads();
    if (ng <= 32) {
        subst_warp(c.S, c.linv, c.r, ng);
    } else {
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
    }
    lap(10);
    // d_l = -(g_l + C_l . d_g) / (H_ll + tau)
    for (int p = tid; p < P; p += NT) c.d[p] = 0.0;
    __syncthreads();
    for (int i = tid; i < ng; i += NT) c.d[a.glist[i]] = c.r[i];
    if (a.L.lowrank) {
        // d_l = -gl_l - (w_l . u_c) / (H_ll + tau),  u_c = DL_c d_g
        const int S = pl.S, NSEG = pl.NSEG, nj = a.nj;
        for (int cs = tid; cs < NSEG * S; cs += NT) {
            double accv = 0.0;
            for (int j = 0; j < nj; ++j) accv = fma(c.DLall[cs * nj + j], c.r[c.gposj[j]], accv);
            c.uq[cs] = accv;
        }
        __syncthreads();
        for (int i = tid; i < pl.ah_nbl; i += NT) {
            const int li = c.bl2l[i], ch = c.blseg[i];
            double accv = 0.0;
            for (int s_ = 0; s_ < S; ++s_) accv = fma(c.LOCw[i * S + s_], c.uq[ch * S + s_], accv);
            c.d[a.llist[li]] = c.actl[li] ? 0.0 : -c.gl[li] - c.dinv[li] * accv;
        }
    } else if (a.L.small) {
        // C in shared memory, odd row stride: one thread per per-bin parameter reads its row conflict-free
        for (int i = tid; i < nl; i += NT) {
            double dl = 0.0;
            if (!c.actl[i]) {
                const double* crow = c.C + i * ngs;
                double accv = 0.0;
#pragma unroll 4
                for (int k = 0; k < ng; ++k) accv = fma(crow[k], c.r[k], accv);
                dl = -c.gl[i] - c.dinv[i] * accv;
            }
            c.d[a.llist[i]] = dl;
        }
    } else {
        for (int i = warp; i < nl; i += NT / 32) {  // C in global memory: one warp per row, coalesced
            double accv = 0.0;
            if (!c.actl[i]) {
                const double* crow = c.C + (int64_t)i * ngs;
                for (int k = lane; k < ng; k += 32) accv = fma(crow[k], c.r[k], accv);
            }
            for (int o = 16; o > 0; o >>= 1) accv += __shfl_xor_sync(0xffffffffu, accv, o);
            if (lane == 0) c.d[a.llist[i]] = c.actl[i] ? 0.0 : -c.gl[i] - c.dinv[i] * accv;
        }
    }
    __syncthreads();
    double edm = 0.0, step = 0.0;
    for (int p = tid; p < P; p += NT) {
        edm = fma(-c.g[p], c.d[p], edm);
        step = fmax(step, fabs(c.d[p]));
    }
    edm_out = block_sum(edm, c.red);
    step_out = block_max(step, c.red);
    lap(11);
    return true;
}

// mask / data row of fit i
__device__ __forceinline__ const uint8_t* solo_mask(const SoloArgs& a, int P, int64_t i) {
    const int k = a.use_alt ? min((int)a.use_alt[i], a.n_alt) : 0;
    return k ? a.fixed_alt + (int64_t)(k - 1) * P : a.fixed;
}

template <int MAXT>
__global__ void __launch_bounds__(NT, (NT == 128) ? ((MAXT <= 1) ? 6 : 3) : ((MAXT == 0) ? 3 : ((MAXT <= 4) ? 2 : 1)))
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
    constexpr int TBk = (MAXT == 0) ? NT : TB;  // bins per tile (see solo_sweep)
    c.segt = ip; c.lgp = ip + TBk; c.lmk = reinterpret_cast<unsigned*>(ip + 2 * TBk);
    c.actg = ip + 3 * TBk; c.actl = c.actg + a.ng; c.flag = c.actl + a.nl;
    __shared__ long long s_prof[16];
    c.prof = s_prof;
    c.tdec = reinterpret_cast<unsigned short*>(sm + L.tdec);
    for (int t = tid; t < a.ntdec; t += NT) {
        int ta, tb;
        tri_decode(t, ta, tb);
        c.tdec[t] = (unsigned short)((ta << 8) | tb);
    }
    __syncthreads();
    c.DLall = sm + L.DLall; c.Qc = sm + L.Qc; c.vq = sm + L.vq; c.Tq = sm + L.Tq; c.uq = sm + L.uq;
    c.bl2l = c.flag + 4; c.blseg = c.bl2l + pl.ah_nbl; c.gposj = c.blseg + pl.ah_nbl;
    c.pj1 = reinterpret_cast<unsigned char*>(c.gposj + a.nj);
    c.pj2 = c.pj1 + a.nj * (a.nj + 1) / 2;
    c.ppos = reinterpret_cast<unsigned short*>(sm + L.ppos);
    if (MAXT == 0 && L.lowrank) {
        // index tables of the low-rank solve, once per CTA
        for (int ch = 0; ch < pl.NSEG; ++ch) {
            const int i0 = pl.ah_bl_ptr[pl.ah_seg_first[ch]], i1 = pl.ah_bl_ptr[pl.ah_seg_first[ch + 1]];
            for (int i = i0 + tid; i < i1; i += NT) { c.bl2l[i] = a.lpos[pl.ah_bl_par[i]]; c.blseg[i] = ch; }
        }
        for (int j = tid; j < a.nj; j += NT) c.gposj[j] = a.gpos[pl.ah_sfp_list[j]];
        for (int e = tid; e < a.nj * (a.nj + 1) / 2; e += NT) {
            int j2, j1;
            tri_decode(e, j2, j1);
            c.pj1[e] = (unsigned char)j1;
            c.pj2[e] = (unsigned char)j2;
            c.ppos[e] = (unsigned short)trisym(a.gpos[pl.ah_sfp_list[j1]], a.gpos[pl.ah_sfp_list[j2]]);
        }
        __syncthreads();
    }
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
        if (tid < 16) c.prof[tid] = 0;
        __syncthreads();
        double f = INFINITY;
        int status = 1, niter = 0, nfail = 0, nkick = 0, nindef = 0;
        double damp = a.damp0, edm_prev = INFINITY, kscale = 1e-3;
        bool kick_next = false, chk = false, exact_tried = false;
        bool need_sweep = true;  // value, gradient and Hessian of the current point are due (ONE call site: the sweep is inlined)
        const int max_rounds = a.max_iter * 4 + 40;
        for (int round = 0; round < max_rounds && status == 1; ++round) {
            if (need_sweep) {
                f = solo_sweep<MAXT, true>(pl, a, c, c.xc, fisher, drow, dconst);
                need_sweep = false;
                if (!isfinite(f)) { status = 2; break; }
                if (niter >= a.max_iter) { status = 1; break; }
            }
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
                    c.d[p] = mask[p] ? 0.0 : kscale * fmax(1.0, fabs(c.xc[p])) * u;
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
                    need_sweep = true;
                    continue;
                }
                const bool converged = edm < a.tol_edm && step < a.tol_step;
                if (converged && damp > 0.0) { damp = 0.0; continue; }  // look again without the shift: it hides an indefinite Hessian
                if (converged && fisher && !exact_tried) {
                    // the stopping rule is met on the Fisher matrix, which is positive semi-definite by
                    // construction: the exact Hessian of this point decides whether it is a minimum
                    exact_tried = true;
                    fisher = false;
                    need_sweep = true;
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
            // (a kick is taken unconditionally as long as it stays inside the domain and costs less than one
            //  unit of 2NLL; else it is tried again at a tenth of the size)
            const bool ok = isfinite(ft) && (kicked ? ft <= f + 1.0 : ft <= f + kArmijo * dec + slack);
            if (((a.trace && fit < 2) || fit == a.trace_fit) && tid == 0)
                printf("[hf_fit solo] fit %lld it %d f %.12g ft-f %.3e dec %.3e edm %.3e step %.3e damp %.3e %s attempts %d %s\n",
                       (long long)fit, niter, f, ft - f, dec, edm, step, damp, fisher ? "fisher" : "exact", attempts,
                       ok ? "accept" : "reject");
            if (kicked && !ok) {
                if (kscale > 1e-9) { kscale *= 0.1; kick_next = true; nkick -= 1; continue; }
                status = 0;  // stay where the fit was
                break;
            }
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
                f = ft;
                need_sweep = true;
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
                   "channels %lld | solve %lld (%lld calls, %lld factorisations: reduce %lld factor %lld substitute %lld rest %lld) | "
                   "value sweep prolog %lld tiles+sum %lld\n",
                   (long long)fit, status, niter, f, c.prof[0], c.prof[1], c.prof[2], c.prof[3], c.prof[4], c.prof[7],
                   c.prof[12], c.prof[8], c.prof[9], c.prof[10], c.prof[11], c.prof[5], c.prof[6]);
        for (int p = tid; p < P; p += NT) a.x_out[fit * P + p] = c.xc[p];
        if (tid == 0) {
            a.fun_out[fit] = f;
            a.status_out[fit] = status;
            if (a.niter_out) a.niter_out[fit] = niter;
        }
    }
}


}  // namespace SOLO_NS
