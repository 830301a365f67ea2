// C-ABI of libhf_b200.so (include/hf_b200.h): plan lifetime, workspace, thin entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "hf_internal.cuh"

static thread_local char g_err[1024] = "";

void hf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hf_expected_launch(hf_plan* plan, const double* pars, int64_t B, double* out, cudaStream_t st);
int hf_fit_launch(hf_plan* plan, const double* data, int64_t data_rows, const double* init,
                  int64_t init_rows, const double* lo, const double* hi, const uint8_t* fixed,
                  int64_t N, const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                  int32_t* niter, cudaStream_t st, const uint8_t* fixed_alt = nullptr,
                  const uint8_t* use_alt = nullptr, const int32_t* data_map = nullptr, int n_alt = 1);
int hf_sample_toys_launch(hf_plan* plan, const double* pars, uint64_t seed, uint64_t offset,
                          int64_t N, double* out, cudaStream_t st);

extern "C" {

int hf_version(void) { return 100; }
const char* hf_last_error(void) { return g_err; }

// ------------------------------------------------------------------ plan
namespace {
struct Packer {
    // lays out every table in one device allocation, 256-byte aligned
    std::vector<char> host;
    size_t add(const void* src, size_t bytes) {
        size_t off = (host.size() + 255) / 256 * 256;
        host.resize(off + bytes, 0);
        if (src && bytes) memcpy(host.data() + off, src, bytes);
        return off;
    }
};
}  // namespace

int hf_plan_create(const hf_plan_desc* d, hf_plan** out) {
    if (!d || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (d->n_pars <= 0 || d->n_bins <= 0 || d->n_samples <= 0 || d->n_segs <= 0 || d->n_hs < 0 ||
        d->n_sf < 0 || d->n_bf_slots < 0 || d->n_gauss < 0 || d->n_pois < 0 || d->n_aux < 0 ||
        (d->n_hs > 0 && d->n_hs_rows <= 0)) {
        hf_set_error("inconsistent plan sizes");
        return HF_ERR_INVALID;
    }
    const int P = d->n_pars, NB = d->n_bins, S = d->n_samples, NSEG = d->n_segs, H = d->n_hs,
              R = d->n_hs ? d->n_hs_rows : 0, NSF = d->n_sf, K = d->n_bf_slots;
    // ---- validate indices on the host (a bad plan must not become an illegal address)
    auto in_range = [](const int32_t* a, size_t n, int lo, int hi) {
        for (size_t i = 0; i < n; ++i)
            if (a[i] < lo || a[i] >= hi) return false;
        return true;
    };
    bool ok = in_range(d->bin_seg, NB, 0, NSEG) && in_range(d->hs_par, H, 0, P) &&
              in_range(d->hs_row_sample, R, 0, S) && in_range(d->sf_par, NSF, 0, P) &&
              in_range(d->sf_kind, NSF, 0, 3) && in_range(d->bf_par, (size_t)K * S * NB, -1, P) &&
              in_range(d->g_par, d->n_gauss, 0, P) && in_range(d->g_aux, d->n_gauss, 0, std::max(d->n_aux, 1)) &&
              in_range(d->p_par, d->n_pois, 0, P) && in_range(d->p_aux, d->n_pois, 0, std::max(d->n_aux, 1));
    if (ok) {
        if (d->sf_ptr[0] != 0 || d->sf_ptr[S * NSEG] != NSF) ok = false;
        for (int q = 0; ok && q < S * NSEG; ++q)
            if (d->sf_ptr[q + 1] < d->sf_ptr[q]) ok = false;
    }
    if (d->hs_code != HF_HS_CODE0 && d->hs_code != HF_HS_CODE4P) ok = false;
    if (!ok) {
        hf_set_error("plan tables hold out-of-range indices");
        return HF_ERR_INVALID;
    }

    hf_plan* plan = new (std::nothrow) hf_plan();
    if (!plan) return HF_ERR_NOMEM;
    DevPlan& dp = plan->d;
    memset(&dp, 0, sizeof(dp));
    dp.P = P; dp.NB = NB; dp.S = S; dp.NAUX = d->n_aux; dp.NSEG = NSEG; dp.H = H; dp.R = R;
    dp.hs_code = d->hs_code; dp.NSF = NSF; dp.K = K; dp.NG = d->n_gauss; dp.NPO = d->n_pois;
    dp.D = NB + d->n_aux;
    dp.Hp = (H + 31) / 32 * 32;
    dp.has_clip_sample = d->has_clip_sample; dp.has_clip_bin = d->has_clip_bin;
    dp.clip_sample = d->clip_sample; dp.clip_bin = d->clip_bin;
    plan->P = P; plan->NB = NB; plan->S = S; plan->NAUX = d->n_aux; plan->NSEG = NSEG; plan->H = H;
    plan->R = R; plan->NSF = NSF; plan->K = K; plan->NG = d->n_gauss; plan->NPO = d->n_pois;
    plan->D = dp.D;

    // derived host tables
    std::vector<int32_t> sample_row(S, -1);
    for (int r = 0; r < R; ++r) sample_row[d->hs_row_sample[r]] = r;
    // transposed histosys tables [R][NB][Hp]
    const size_t ncell = (size_t)R * NB;
    std::vector<double> t1t(ncell * dp.Hp, 0.0), t2t(ncell * dp.Hp, 0.0);
    for (int h = 0; h < H; ++h)
        for (size_t c = 0; c < ncell; ++c) {
            t1t[c * dp.Hp + h] = d->hs_t1[(size_t)h * ncell + c];
            t2t[c * dp.Hp + h] = d->hs_t2[(size_t)h * ncell + c];
        }
    // by-parameter CSR of the scalar-factor entries
    std::vector<int32_t> sfp_ptr(P + 1, 0), sfp_entry(NSF), sfp_q(NSF);
    for (int e = 0; e < NSF; ++e) sfp_ptr[d->sf_par[e] + 1]++;
    for (int p = 0; p < P; ++p) sfp_ptr[p + 1] += sfp_ptr[p];
    {
        std::vector<int32_t> fill(sfp_ptr.begin(), sfp_ptr.end() - 1);
        for (int q = 0; q < S * NSEG; ++q)
            for (int e = d->sf_ptr[q]; e < d->sf_ptr[q + 1]; ++e) {
                const int k = fill[d->sf_par[e]]++;
                sfp_entry[k] = e;
                sfp_q[k] = q;
            }
    }
    std::vector<int32_t> sfp_kind(NSF);
    std::vector<double> sfp_coef((size_t)NSF * 8);
    for (int k = 0; k < NSF; ++k) {
        sfp_kind[k] = d->sf_kind[sfp_entry[k]];
        for (int i = 0; i < 8; ++i) sfp_coef[(size_t)k * 8 + i] = d->sf_coef[(size_t)sfp_entry[k] * 8 + i];
    }
    std::vector<int32_t> cons_kind(P, 0), cons_idx(P, 0);
    for (int i = 0; i < d->n_gauss; ++i) { cons_kind[d->g_par[i]] = 1; cons_idx[d->g_par[i]] = i; }
    for (int i = 0; i < d->n_pois; ++i) { cons_kind[d->p_par[i]] = 2; cons_idx[d->p_par[i]] = i; }

    // DMMA path tables: [tile][half][chunk][k_local][cell], cell = (bin_local % 16) * 4 + row, zero
    // padded: one (tile, half, chunk) block is what one half of a CTA fetches with one bulk copy
    int n_other = 0, other[2] = {-1, -1};
    for (int s = 0; s < S; ++s)
        if (sample_row[s] < 0) {
            if (n_other < 2) other[n_other] = s;
            ++n_other;
        }
    const bool mma_ok = H > 0 && R <= 4 && n_other <= 4 && !d->has_clip_sample && !d->has_clip_bin;
    const int mma_ntiles = (NB + HF_MMA_BT - 1) / HF_MMA_BT;
    const int mma_nchunks = (2 * H + HF_MMA_KC - 1) / HF_MMA_KC;
    std::vector<double> tt;
    if (mma_ok) {
        tt.assign((size_t)mma_ntiles * 2 * mma_nchunks * HF_MMA_KC * HF_MMA_CSH, 0.0);
        for (int h = 0; h < H; ++h)
            for (int r = 0; r < R; ++r)
                for (int b = 0; b < NB; ++b) {
                    const int tile = b / HF_MMA_BT, bl = b % HF_MMA_BT, half = bl / 16, cell = (bl % 16) * 4 + r;
                    for (int which = 0; which < 2; ++which) {
                        const int k = 2 * h + which, chunk = k / HF_MMA_KC, kl = k % HF_MMA_KC;
                        const double v = (which ? d->hs_t2 : d->hs_t1)[((size_t)h * R + r) * NB + b];
                        tt[(((size_t)tile * 2 + half) * mma_nchunks + chunk) * HF_MMA_KC * HF_MMA_CSH + HF_MMA_ROW(kl) + cell] = v;
                    }
                }
    }

    Packer pk;
    const size_t o_tt = pk.add(tt.data(), sizeof(double) * tt.size());
    const size_t o_nom = pk.add(d->nominal, sizeof(double) * S * NB);
    const size_t o_seg = pk.add(d->bin_seg, sizeof(int32_t) * NB);
    const size_t o_hpar = pk.add(d->hs_par, sizeof(int32_t) * H);
    const size_t o_hrow = pk.add(d->hs_row_sample, sizeof(int32_t) * R);
    const size_t o_srow = pk.add(sample_row.data(), sizeof(int32_t) * S);
    const size_t o_t1 = pk.add(d->hs_t1, sizeof(double) * H * ncell);
    const size_t o_t2 = pk.add(d->hs_t2, sizeof(double) * H * ncell);
    const size_t o_t1t = pk.add(t1t.data(), sizeof(double) * t1t.size());
    const size_t o_t2t = pk.add(t2t.data(), sizeof(double) * t2t.size());
    const size_t o_sfptr = pk.add(d->sf_ptr, sizeof(int32_t) * (S * NSEG + 1));
    const size_t o_sfpar = pk.add(d->sf_par, sizeof(int32_t) * NSF);
    const size_t o_sfkind = pk.add(d->sf_kind, sizeof(int32_t) * NSF);
    const size_t o_sfcoef = pk.add(d->sf_coef, sizeof(double) * 8 * NSF);
    const size_t o_sfpptr = pk.add(sfp_ptr.data(), sizeof(int32_t) * (P + 1));
    const size_t o_sfpent = pk.add(sfp_entry.data(), sizeof(int32_t) * NSF);
    const size_t o_sfpq = pk.add(sfp_q.data(), sizeof(int32_t) * NSF);
    const size_t o_sfpkind = pk.add(sfp_kind.data(), sizeof(int32_t) * NSF);
    const size_t o_sfpcoef = pk.add(sfp_coef.data(), sizeof(double) * 8 * NSF);
    const size_t o_bf = pk.add(d->bf_par, sizeof(int32_t) * (size_t)K * S * NB);
    const size_t o_gpar = pk.add(d->g_par, sizeof(int32_t) * d->n_gauss);
    const size_t o_gaux = pk.add(d->g_aux, sizeof(int32_t) * d->n_gauss);
    const size_t o_gsig = pk.add(d->g_sigma, sizeof(double) * d->n_gauss);
    const size_t o_ppar = pk.add(d->p_par, sizeof(int32_t) * d->n_pois);
    const size_t o_paux = pk.add(d->p_aux, sizeof(int32_t) * d->n_pois);
    const size_t o_pfac = pk.add(d->p_factor, sizeof(double) * d->n_pois);
    // groups of up to 4 consecutive (sample, segment) products whose entry lists name the same
    // parameters with the same kinds in the same order (normally: the samples of a channel share
    // their normalisation systematics): hf_prep_kernel loads a parameter once per group
    std::vector<int32_t> qg_first, qg_n;
    for (int q = 0; q < S * NSEG;) {
        int n = 1;
        const int e0 = d->sf_ptr[q], len = d->sf_ptr[q + 1] - e0;
        while (n < 4 && q + n < S * NSEG && len > 0) {
            const int f0 = d->sf_ptr[q + n];
            bool same = (d->sf_ptr[q + n + 1] - f0) == len;
            for (int i = 0; same && i < len; ++i)
                same = d->sf_par[e0 + i] == d->sf_par[f0 + i] && d->sf_kind[e0 + i] == d->sf_kind[f0 + i];
            if (!same) break;
            ++n;
        }
        qg_first.push_back(q);
        qg_n.push_back(n);
        q += n;
    }
    const size_t o_qgf = pk.add(qg_first.data(), sizeof(int32_t) * qg_first.size());
    const size_t o_qgn = pk.add(qg_n.data(), sizeof(int32_t) * qg_n.size());
    const size_t o_ck = pk.add(cons_kind.data(), sizeof(int32_t) * P);
    const size_t o_ci = pk.add(cons_idx.data(), sizeof(int32_t) * P);
    // ---- analytic-Hessian tables and qualification (hf_hess_analytic.cu)
    bool ah_ok = !d->has_clip_sample && !d->has_clip_bin && S <= HF_AH_SMAX && H <= 120 && R <= HF_AH_SMAX;
    std::vector<int32_t> sfp_list, sf_j(NSF, 0), par_j(P, -1), bl_ptr(NB + 1, 0), bl_par, seg_first(NSEG + 1, NB);
    std::vector<uint32_t> bl_mask;
    std::vector<uint8_t> par_mult(P, 0);
    {
        for (int e = 0; e < NSF; ++e) {
            const int p = d->sf_par[e];
            if (par_j[p] < 0) { par_j[p] = (int)sfp_list.size(); sfp_list.push_back(p); }
            sf_j[e] = par_j[p];
            if (d->sf_kind[e] == HF_SF_THETA) par_mult[p] = 1;
        }
        // one entry per (sample, segment, parameter)
        for (int q = 0; ah_ok && q < S * NSEG; ++q)
            for (int e = d->sf_ptr[q]; ah_ok && e < d->sf_ptr[q + 1]; ++e)
                for (int e2 = e + 1; e2 < d->sf_ptr[q + 1]; ++e2)
                    if (d->sf_par[e] == d->sf_par[e2]) ah_ok = false;
        // bins of a segment are contiguous
        for (int b = 0; b < NB; ++b) {
            if (b > 0 && d->bin_seg[b] < d->bin_seg[b - 1]) ah_ok = false;
            seg_first[d->bin_seg[b]] = std::min(seg_first[d->bin_seg[b]], b);
        }
        for (int c = NSEG - 1; c >= 0; --c)
            if (seg_first[c] == NB) seg_first[c] = seg_first[c + 1];  // empty segment
        // per-bin parameters: each acts on exactly one bin, is no scalar-factor / histosys parameter,
        // and multiplies a cell at most once
        std::vector<int32_t> bin_of(P, -1);
        std::vector<char> is_glob(P, 0);
        for (int h = 0; h < H; ++h) is_glob[d->hs_par[h]] = 1;
        for (int e = 0; e < NSF; ++e) is_glob[d->sf_par[e]] = 1;
        for (int b = 0; b < NB; ++b) {
            std::vector<int32_t> here;
            std::vector<uint32_t> mk;
            for (int s2 = 0; s2 < S; ++s2)
                for (int k = 0; k < K; ++k) {
                    const int p = d->bf_par[((size_t)k * S + s2) * NB + b];
                    if (p < 0) continue;
                    par_mult[p] = 1;
                    if (is_glob[p] || (bin_of[p] >= 0 && bin_of[p] != b)) ah_ok = false;
                    bin_of[p] = b;
                    size_t i = 0;
                    while (i < here.size() && here[i] != p) ++i;
                    if (i == here.size()) { here.push_back(p); mk.push_back(0u); }
                    if (s2 < 32) {
                        if (mk[i] & (1u << s2)) ah_ok = false;  // the same parameter twice on one cell
                        mk[i] |= 1u << s2;
                    }
                }
            if ((int)here.size() > HF_AH_KLMAX) ah_ok = false;
            bl_ptr[b + 1] = bl_ptr[b] + (int)here.size();
            bl_par.insert(bl_par.end(), here.begin(), here.end());
            bl_mask.insert(bl_mask.end(), mk.begin(), mk.end());
        }
    }
    for (int c = 0; c < NSEG; ++c) plan->ah_max_lc = std::max(plan->ah_max_lc, bl_ptr[seg_first[c + 1]] - bl_ptr[seg_first[c]]);
    const size_t o_ahl = pk.add(sfp_list.data(), sizeof(int32_t) * sfp_list.size());
    const size_t o_ahj = pk.add(sf_j.data(), sizeof(int32_t) * NSF);
    const size_t o_ahbp = pk.add(bl_ptr.data(), sizeof(int32_t) * (NB + 1));
    const size_t o_ahbq = pk.add(bl_par.data(), sizeof(int32_t) * bl_par.size());
    const size_t o_ahbm = pk.add(bl_mask.data(), sizeof(uint32_t) * bl_mask.size());
    const size_t o_ahsf = pk.add(seg_first.data(), sizeof(int32_t) * (NSEG + 1));
    const size_t o_ahpm = pk.add(par_mult.data(), P);
    pk.add(nullptr, 256);

    cudaError_t e = cudaMalloc(&plan->table_block, pk.host.size());
    if (e != cudaSuccess) {
        hf_set_error("cudaMalloc(%zu) for plan tables: %s", pk.host.size(), cudaGetErrorString(e));
        delete plan;
        return HF_ERR_NOMEM;
    }
    e = cudaMemcpy(plan->table_block, pk.host.data(), pk.host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        hf_set_error("cudaMemcpy plan tables: %s", cudaGetErrorString(e));
        cudaFree(plan->table_block);
        delete plan;
        return HF_ERR_CUDA;
    }
    char* base = (char*)plan->table_block;
#define DP(T, off) reinterpret_cast<const T*>(base + (off))
    dp.nominal = DP(double, o_nom); dp.bin_seg = DP(int, o_seg); dp.hs_par = DP(int, o_hpar);
    dp.hs_row_sample = DP(int, o_hrow); dp.sample_hs_row = DP(int, o_srow);
    dp.t1 = DP(double, o_t1); dp.t2 = DP(double, o_t2); dp.t1t = DP(double, o_t1t);
    dp.t2t = DP(double, o_t2t); dp.sf_ptr = DP(int, o_sfptr); dp.sf_par = DP(int, o_sfpar);
    dp.sf_kind = DP(int, o_sfkind); dp.sf_coef = DP(double, o_sfcoef);
    dp.sfp_ptr = DP(int, o_sfpptr); dp.sfp_entry = DP(int, o_sfpent); dp.sfp_q = DP(int, o_sfpq);
    dp.sfp_kind = DP(int, o_sfpkind); dp.sfp_coef = DP(double, o_sfpcoef);
    dp.bf_par = DP(int, o_bf); dp.g_par = DP(int, o_gpar); dp.g_aux = DP(int, o_gaux);
    dp.g_sigma = DP(double, o_gsig); dp.p_par = DP(int, o_ppar); dp.p_aux = DP(int, o_paux);
    dp.p_factor = DP(double, o_pfac); dp.par_cons_kind = DP(int, o_ck); dp.par_cons_idx = DP(int, o_ci);
    dp.qg_first = DP(int, o_qgf); dp.qg_n = DP(int, o_qgn); dp.n_qg = (int)qg_first.size();
    dp.max_sf_q = 0; dp.max_sf_par = 0;
    for (int q = 0; q < S * NSEG; ++q) dp.max_sf_q = std::max(dp.max_sf_q, d->sf_ptr[q + 1] - d->sf_ptr[q]);
    for (int pp = 0; pp < P; ++pp) dp.max_sf_par = std::max(dp.max_sf_par, sfp_ptr[pp + 1] - sfp_ptr[pp]);
    dp.mma_ok = mma_ok ? 1 : 0; dp.mma_ntiles = mma_ntiles; dp.mma_nchunks = mma_nchunks;
    {
        std::vector<int> seen(P, 0);
        dp.mma_hs_unique = 1;
        for (int h = 0; h < H; ++h)
            if (seen[d->hs_par[h]]++) dp.mma_hs_unique = 0;
        std::vector<int> bin_of(P, -1);
        dp.mma_bf_unique = 1;
        for (size_t i = 0; i < (size_t)K * S * NB; ++i) {
            const int q = d->bf_par[i], b = (int)(i % NB);
            if (q < 0) continue;
            if (bin_of[q] >= 0 && bin_of[q] != b) dp.mma_bf_unique = 0;
            bin_of[q] = b;
        }
    }
    dp.mma_nother = n_other; dp.mma_other[0] = other[0]; dp.mma_other[1] = other[1];
    dp.mma_tt = DP(double, o_tt);
    dp.ah_ok = ah_ok ? 1 : 0; dp.ah_nsfp = (int)sfp_list.size(); dp.ah_nbl = (int)bl_par.size();
    dp.ah_sfp_list = DP(int, o_ahl); dp.ah_sf_j = DP(int, o_ahj); dp.ah_bl_ptr = DP(int, o_ahbp);
    dp.ah_bl_par = DP(int, o_ahbq); dp.ah_bl_mask = DP(unsigned, o_ahbm); dp.ah_seg_first = DP(int, o_ahsf);
    dp.ah_par_mult = DP(unsigned char, o_ahpm);
#undef DP
    plan->h_bf_par.assign(d->bf_par, d->bf_par + (size_t)K * S * NB);
    plan->h_hs_par.assign(d->hs_par, d->hs_par + H);
    plan->h_sf_par.assign(d->sf_par, d->sf_par + NSF);
    cudaGetDevice(&plan->device);
    cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, plan->device);
    *out = plan;
    return HF_OK;
}

void hf_plan_destroy(hf_plan* plan) {
    if (!plan) return;
    cudaFree(plan->table_block);
    cudaFree(plan->work_block);
    cudaFree(plan->dconst);
    if (plan->h_counters) cudaFreeHost(plan->h_counters);
    cudaFree(plan->fit_block);
    cudaFree(plan->hess_block);
    cudaFree(plan->solo_dev);
    cudaFree(plan->solo_scratch);
    cudaFree(plan->stage_block);
    for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
    if (plan->s_in) cudaStreamDestroy(plan->s_in);
    if (plan->s_out) cudaStreamDestroy(plan->s_out);
    delete plan;
}

int64_t hf_plan_launch_count(const hf_plan* plan) { return plan ? plan->launches : 0; }

int hf_plan_set_timing(hf_plan* plan, int enable) {
    if (!plan) return HF_ERR_INVALID;
    for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
    plan->ev.clear();
    plan->timing = enable != 0;
    return HF_OK;
}

int hf_plan_main_kernel_ms(hf_plan* plan, double* ms_out, int64_t* launches_out) {
    if (!plan || !ms_out || !launches_out) return HF_ERR_INVALID;
    double total = 0.0;
    for (size_t i = 0; i + 1 < plan->ev.size(); i += 2) {
        HF_CUDA_OK(cudaEventSynchronize(plan->ev[i + 1]));
        float ms = 0.f;
        HF_CUDA_OK(cudaEventElapsedTime(&ms, plan->ev[i], plan->ev[i + 1]));
        total += ms;
    }
    *ms_out = total;
    *launches_out = (int64_t)(plan->ev.size() / 2);
    for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
    plan->ev.clear();
    return HF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ workspace
int hf_ensure_work(hf_plan* plan, int64_t ld, int64_t data_rows) {
    const DevPlan& pl = plan->d;
    if (ld > plan->work_ld) {
        if (plan->work_block) cudaFree(plan->work_block);
        plan->work_block = nullptr;
        plan->work_ld = 0;
        const int64_t nq = (int64_t)pl.S * pl.NSEG;
        // [parsT P][c1,c2,dc1,dc2 4H][Fseg,Gseg 2nq][gradT P][mainv 1]
        const int64_t rows = 2LL * pl.P + 4LL * pl.H + 2 * nq + 1 + 8;
        const size_t bytes = sizeof(double) * (size_t)rows * (size_t)ld + 256;
        cudaError_t e = cudaMalloc(&plan->work_block, bytes);
        if (e != cudaSuccess) {
            hf_set_error("cudaMalloc(%zu) for the evaluation workspace: %s", bytes, cudaGetErrorString(e));
            return HF_ERR_NOMEM;
        }
        plan->work_ld = ld;
    }
    {
        // carve for the ld of THIS call (leading dimension = ld so padded tiles stay in range)
        const int64_t nq = (int64_t)pl.S * pl.NSEG;
        double* p = (double*)plan->work_block;
        Work& w = plan->w;
        w.parsT = p; p += (int64_t)pl.P * ld;
        w.c1T = p; p += (int64_t)pl.H * ld;
        w.c2T = p; p += (int64_t)pl.H * ld;
        w.dc1T = p; p += (int64_t)pl.H * ld;
        w.dc2T = p; p += (int64_t)pl.H * ld;
        w.FsegT = p; p += nq * ld;
        w.GsegT = p; p += nq * ld;               // GsegT | gradT | mainv stay adjacent: hf_eval_launch
        w.gradT = p; p += (int64_t)pl.P * ld;     // clears the three accumulators with one memset
        w.mainv = p; p += ld;
    }
    if (data_rows > plan->dconst_cap) {
        if (plan->dconst) cudaFree(plan->dconst);
        plan->dconst = nullptr;
        plan->dconst_cap = 0;
        cudaError_t e = cudaMalloc(&plan->dconst, sizeof(double) * (size_t)data_rows);
        if (e != cudaSuccess) {
            hf_set_error("cudaMalloc for data constants: %s", cudaGetErrorString(e));
            return HF_ERR_NOMEM;
        }
        plan->dconst_cap = data_rows;
    }
    plan->w.dconst = plan->dconst;
    return HF_OK;
}

static int ensure_stage(hf_plan* plan, size_t bytes) {
    if (bytes <= plan->stage_bytes) return HF_OK;
    if (plan->stage_block) cudaFree(plan->stage_block);
    plan->stage_block = nullptr;
    plan->stage_bytes = 0;
    cudaError_t e = cudaMalloc(&plan->stage_block, bytes);
    if (e != cudaSuccess) {
        hf_set_error("cudaMalloc(%zu) for host-call staging: %s", bytes, cudaGetErrorString(e));
        return HF_ERR_NOMEM;
    }
    plan->stage_bytes = bytes;
    return HF_OK;
}

extern "C" {

int hf_logpdf(hf_plan* plan, const double* pars, const double* data, int64_t data_rows, int64_t B,
              double* out, void* stream) {
    HfRange nvtx_range("K1 hf_logpdf");
    if (!plan || !pars || !data || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_eval_launch(plan, pars, data, data_rows, B, out, nullptr, (cudaStream_t)stream);
}

int hf_logpdf_grad(hf_plan* plan, const double* pars, const double* data, int64_t data_rows,
                   int64_t B, double* out, double* grad, void* stream) {
    HfRange nvtx_range("K2 hf_logpdf_grad");
    if (!plan || !pars || !data || !out || !grad) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_eval_launch(plan, pars, data, data_rows, B, out, grad, (cudaStream_t)stream);
}

int hf_expected_data(hf_plan* plan, const double* pars, int64_t B, double* out, void* stream) {
    HfRange nvtx_range("hf_expected_data");
    if (!plan || !pars || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_expected_launch(plan, pars, B, out, (cudaStream_t)stream);
}

int hf_logpdf_grad_host(hf_plan* plan, const double* pars_host, const double* data_host,
                        int64_t data_rows, int64_t B, double* out_host, double* grad_host,
                        void* stream) {
    HfRange nvtx_range("K2 hf_logpdf_grad_host");
    if (!plan || !pars_host || !data_host || !out_host) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = plan->P, D = plan->D;
    const size_t n_pars = (size_t)B * P, n_data = (size_t)data_rows * D;
    const size_t bytes = sizeof(double) * (n_pars + n_data + (size_t)B + (grad_host ? n_pars : 0)) + 1024;
    int rc = ensure_stage(plan, bytes);
    if (rc) return rc;
    double* d_pars = (double*)plan->stage_block;
    double* d_data = d_pars + n_pars;
    double* d_out = d_data + n_data;
    double* d_grad = grad_host ? d_out + B : nullptr;
    // Pipeline over chunks of points: H2D of chunk i+1, the kernels of chunk i and D2H of chunk
    // i-1 overlap (PCIe is full duplex); the copy streams are ordered against `st` by events.
    // Chunks are whole waves of the dominant kernel (one CTA of HF_MMA_M points per SM and wave:
    // a 16384-point chunk would be 3.46 waves on 148 SMs and pay for 4).  The first upload and the
    // last download cannot overlap anything, so the schedule ramps up from one wave and down to one
    // wave (below; C4, 65 536 points: 1, 3, 6.84, 2, 1 waves = 10.75 ms against 11.02 ms for
    // 1, 2, 3, 3, 3, 0.84, 1 and 9.73 ms for the kernels alone).
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t WAVE = (int64_t)sms * HF_MMA_M;
    const int64_t CH = WAVE * 3;
    if (B <= 2 * CH || data_rows != 1) {
        HF_CUDA_OK(cudaMemcpyAsync(d_pars, pars_host, sizeof(double) * n_pars, cudaMemcpyHostToDevice, st));
        HF_CUDA_OK(cudaMemcpyAsync(d_data, data_host, sizeof(double) * n_data, cudaMemcpyHostToDevice, st));
        rc = hf_eval_launch(plan, d_pars, d_data, data_rows, B, d_out, d_grad, st);
        if (rc) return rc;
        HF_CUDA_OK(cudaMemcpyAsync(out_host, d_out, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        if (grad_host)
            HF_CUDA_OK(cudaMemcpyAsync(grad_host, d_grad, sizeof(double) * n_pars, cudaMemcpyDeviceToHost, st));
        HF_CUDA_OK(cudaStreamSynchronize(st));
        return HF_OK;
    }
    if (!plan->s_in) {
        HF_CUDA_OK(cudaStreamCreateWithFlags(&plan->s_in, cudaStreamNonBlocking));
        HF_CUDA_OK(cudaStreamCreateWithFlags(&plan->s_out, cudaStreamNonBlocking));
    }
    std::vector<int64_t> ch_b0, ch_nb;
    {
        // Waves per chunk: ramp up (the upload of chunk i+1 has to fit under the kernels of chunk i: an
        // upload or download takes ~0.3 of the kernel time of the same points), a steady size, ramp
        // down (the download of chunk i has to fit under the kernels of chunk i+1; the last download is
        // exposed, so the last chunk is one wave).  HF_E2E_SCHED="1,3;8;2,1" overrides head;steady;tail.
        std::vector<int> head = {1, 3}, tail = {2, 1};
        int steady = 8;
        if (const char* e = getenv("HF_E2E_SCHED")) {
            std::vector<int> part[3];
            int k = 0;
            for (const char* q = e; *q && k < 3;) {
                if (*q == ';') { ++k; ++q; continue; }
                if (*q == ',') { ++q; continue; }
                const int v = atoi(q);
                if (v > 0) part[k].push_back(v);
                while (*q && *q != ',' && *q != ';') ++q;
            }
            if (!part[0].empty() && !part[1].empty() && !part[2].empty()) {
                head = part[0]; steady = part[1][0]; tail = part[2];
            }
        }
        int64_t tail_pts = 0;
        for (int v : tail) tail_pts += WAVE * v;
        int64_t b0 = 0, left = B - tail_pts;  // the closing chunks are set aside
        for (size_t k = 0; left > 0; ++k) {
            const int64_t nb = std::min<int64_t>(left, WAVE * (k < head.size() ? head[k] : steady));
            ch_b0.push_back(b0); ch_nb.push_back(nb);
            b0 += nb; left -= nb;
        }
        for (size_t k = 0; k < tail.size() && b0 < B; ++k) {
            const int64_t nb = (k + 1 == tail.size()) ? B - b0 : std::min<int64_t>(B - b0, WAVE * tail[k]);
            ch_b0.push_back(b0); ch_nb.push_back(nb);
            b0 += nb;
        }
    }
    const int64_t nch = (int64_t)ch_b0.size();
    std::vector<cudaEvent_t> ev_in(nch), ev_done(nch);
    cudaEvent_t ev_start;
    HF_CUDA_OK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    HF_CUDA_OK(cudaEventRecord(ev_start, st));  // everything queued on `st` before this call is visible
    HF_CUDA_OK(cudaStreamWaitEvent(plan->s_in, ev_start, 0));
    HF_CUDA_OK(cudaStreamWaitEvent(plan->s_out, ev_start, 0));
    HF_CUDA_OK(cudaMemcpyAsync(d_data, data_host, sizeof(double) * n_data, cudaMemcpyHostToDevice, plan->s_in));
    for (int64_t c = 0; c < nch; ++c) {
        const int64_t b0 = ch_b0[c], nb = ch_nb[c];
        HF_CUDA_OK(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        HF_CUDA_OK(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        HF_CUDA_OK(cudaMemcpyAsync(d_pars + b0 * P, pars_host + b0 * P, sizeof(double) * nb * P,
                                   cudaMemcpyHostToDevice, plan->s_in));
        HF_CUDA_OK(cudaEventRecord(ev_in[c], plan->s_in));
    }
    for (int64_t c = 0; c < nch; ++c) {
        const int64_t b0 = ch_b0[c], nb = ch_nb[c];
        HF_CUDA_OK(cudaStreamWaitEvent(st, ev_in[c], 0));
        rc = hf_eval_launch(plan, d_pars + b0 * P, d_data, 1, nb, d_out + b0, d_grad ? d_grad + b0 * P : nullptr, st);
        if (rc) return rc;
        HF_CUDA_OK(cudaEventRecord(ev_done[c], st));
        HF_CUDA_OK(cudaStreamWaitEvent(plan->s_out, ev_done[c], 0));
        HF_CUDA_OK(cudaMemcpyAsync(out_host + b0, d_out + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, plan->s_out));
        if (grad_host)
            HF_CUDA_OK(cudaMemcpyAsync(grad_host + b0 * P, d_grad + b0 * P, sizeof(double) * nb * P,
                                       cudaMemcpyDeviceToHost, plan->s_out));
    }
    HF_CUDA_OK(cudaStreamSynchronize(plan->s_out));
    HF_CUDA_OK(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < nch; ++c) {
        cudaEventDestroy(ev_in[c]);
        cudaEventDestroy(ev_done[c]);
    }
    cudaEventDestroy(ev_start);
    return HF_OK;
}

void hf_fit_default_opts(hf_fit_opts* o) {
    if (!o) return;
    o->max_iter = 200;
    o->reserved0 = 0;
    o->tol_edm = 1e-11;
    o->tol_step = 1e-8;
    o->verbose = 0;
    o->reserved = 0;
}

int hf_plan_fit_info(const hf_plan* plan, int32_t info[8]) {
    if (!plan || !info) {
        hf_set_error("hf_plan_fit_info: null argument");
        return HF_ERR_INVALID;
    }
    info[0] = plan->solo_last_nt > 0 ? 1 : 0;
    info[1] = plan->solo_last_grid; info[2] = plan->solo_last_nt; info[3] = plan->solo_last_occ;
    info[4] = plan->solo_last_smem; info[5] = plan->sm_count; info[6] = plan->solo_last_lowrank; info[7] = 0;
    return HF_OK;
}

int hf_fit_batch(hf_plan* plan, const double* data, int64_t data_rows, const double* init,
                 int64_t init_rows, const double* lo, const double* hi, const uint8_t* fixed,
                 int64_t N, const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                 int32_t* niter, void* stream) {
    if (!plan || !data || !init || !lo || !hi || !fixed || !x || !fun || !status) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_fit_launch(plan, data, data_rows, init, init_rows, lo, hi, fixed, N, opts, x, fun,
                         status, niter, (cudaStream_t)stream);
}

int hf_fit_batch_mixed(hf_plan* plan, const double* data, int64_t data_rows, const int32_t* data_map,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* fixed, const uint8_t* fixed_alt, const uint8_t* use_alt,
                       int64_t N, const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                       int32_t* niter, void* stream) {
    if (!plan || !data || !init || !lo || !hi || !fixed || !x || !fun || !status) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_fit_launch(plan, data, data_rows, init, init_rows, lo, hi, fixed, N, opts, x, fun,
                         status, niter, (cudaStream_t)stream, fixed_alt, use_alt, data_map);
}

int hf_fit_batch_masks(hf_plan* plan, const double* data, int64_t data_rows, const int32_t* data_map,
                       const double* init, int64_t init_rows, const double* lo, const double* hi,
                       const uint8_t* masks, int32_t n_masks, const uint8_t* mask_id, int64_t N,
                       const hf_fit_opts* opts, double* x, double* fun, int32_t* status,
                       int32_t* niter, void* stream) {
    if (!plan || !data || !init || !lo || !hi || !masks || !x || !fun || !status) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    if (n_masks < 1 || n_masks > 256 || (n_masks > 1 && !mask_id)) {
        hf_set_error("hf_fit_batch_masks: 1..256 masks, and mask_id when there is more than one");
        return HF_ERR_INVALID;
    }
    if (n_masks == 1)
        return hf_fit_launch(plan, data, data_rows, init, init_rows, lo, hi, masks, N, opts, x, fun,
                             status, niter, (cudaStream_t)stream, nullptr, nullptr, data_map);
    return hf_fit_launch(plan, data, data_rows, init, init_rows, lo, hi, masks, N, opts, x, fun,
                         status, niter, (cudaStream_t)stream, masks + plan->P, mask_id, data_map,
                         n_masks - 1);
}

int hf_fit_batch_host(hf_plan* plan, const double* data_host, int64_t data_rows,
                      const double* init_host, int64_t init_rows, const double* lo_host,
                      const double* hi_host, const uint8_t* fixed_host, int64_t N,
                      const hf_fit_opts* opts, double* x_host, double* fun_host,
                      int32_t* status_host, int32_t* niter_host, void* stream) {
    if (!plan || !data_host || !init_host || !lo_host || !hi_host || !fixed_host || !x_host ||
        !fun_host || !status_host) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = plan->P, D = plan->D;
    const size_t nd = (size_t)data_rows * D, ni = (size_t)init_rows * P, nx = (size_t)N * P;
    const size_t bytes = sizeof(double) * (nd + ni + 2 * P + nx + N) + sizeof(int32_t) * 2 * N + P + 4096;
    int rc = ensure_stage(plan, bytes);
    if (rc) return rc;
    double* d_data = (double*)plan->stage_block;
    double* d_init = d_data + nd;
    double* d_lo = d_init + ni;
    double* d_hi = d_lo + P;
    double* d_x = d_hi + P;
    double* d_fun = d_x + nx;
    int32_t* d_status = (int32_t*)(d_fun + N);
    int32_t* d_niter = d_status + N;
    uint8_t* d_fixed = (uint8_t*)(d_niter + N);
    HF_CUDA_OK(cudaMemcpyAsync(d_data, data_host, sizeof(double) * nd, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_init, init_host, sizeof(double) * ni, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_lo, lo_host, sizeof(double) * P, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_hi, hi_host, sizeof(double) * P, cudaMemcpyHostToDevice, st));
    HF_CUDA_OK(cudaMemcpyAsync(d_fixed, fixed_host, P, cudaMemcpyHostToDevice, st));
    rc = hf_fit_launch(plan, d_data, data_rows, d_init, init_rows, d_lo, d_hi, d_fixed, N, opts, d_x,
                       d_fun, d_status, d_niter, st);
    if (rc) return rc;
    HF_CUDA_OK(cudaMemcpyAsync(x_host, d_x, sizeof(double) * nx, cudaMemcpyDeviceToHost, st));
    HF_CUDA_OK(cudaMemcpyAsync(fun_host, d_fun, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    HF_CUDA_OK(cudaMemcpyAsync(status_host, d_status, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    if (niter_host)
        HF_CUDA_OK(cudaMemcpyAsync(niter_host, d_niter, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    HF_CUDA_OK(cudaStreamSynchronize(st));
    return HF_OK;
}

int hf_sample_toys(hf_plan* plan, const double* pars, uint64_t seed, uint64_t offset, int64_t N,
                   double* out, void* stream) {
    HfRange nvtx_range("K4 hf_sample_toys");
    if (!plan || !pars || !out) {
        hf_set_error("null argument");
        return HF_ERR_INVALID;
    }
    return hf_sample_toys_launch(plan, pars, seed, offset, N, out, (cudaStream_t)stream);
}

}  // extern "C"
