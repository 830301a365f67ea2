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
#include "hf_fit_solo_defs.cuh"

namespace solo {
#define SOLO_NT 256
#define SOLO_NS solo256
#include "hf_fit_solo_impl.cuh"
#undef SOLO_NT
#undef SOLO_NS

}  // namespace solo

using namespace solo;

namespace {

SoloLayout solo_layout(const hf_plan* plan, int ng, int nl, int Hpad, int nj, int Hc, int npart, bool small, bool bigch, int NT, bool lowrank) {
    const DevPlan& pl = plan->d;
    SoloLayout L;
    memset(&L, 0, sizeof(L));
    int o = 0;
    auto take = [&](int n) { const int at = o; o += (n + 1) & ~1; return at; };
    const int P = pl.P, H = pl.H, S = pl.S, NSEG = pl.NSEG;
    L.xs = take(P); L.xc = take(P); L.xt = take(P); L.g = take(P); L.d = take(P);
    L.hc = take(6 * std::max(H, 1)); L.Fq = take(S * NSEG);
    // bins per tile: the whole CTA for plans without histosys (no delta / derivative phases: base, FB and
    // gpart are not used), 32 else
    const int TBv = (H == 0) ? NT : TB;
    L.base = take(H == 0 ? 2 : std::max(SMX, NT / 32) * TB); L.rS = take(std::min(S, SMX) * TBv); L.FB = take(H == 0 ? 2 : SMX * TB);
    L.Wt = take(TBv); L.wgt = take(TBv); L.w2t = take(TBv); L.sqW = take(TBv); L.lwj = take(TBv); L.lw2 = take(TBv);
    L.gpart = take(H == 0 ? 32 : 2 * NT);  // (NT: CTA size of the instance the layout is for; low 16 slots: solve scratch)
    const int npk = ng * (ng + 1) / 2;
    // the Y region doubles as DL | D2 | U0 | Msm | G2sm after the sweep and as the staging tile of C
    L.ltile = small ? 0 : 16;
    L.wstride = (ng + 7) / 8 * 8;
    while (L.wstride % 16 != 8) L.wstride += 8;
    L.ngs = small ? (ng | 1) : ng;
    int njs = (nj + 7) / 8 * 8;
    while (njs % 16 != 8) njs += 8;
    const int chan_small = 2 * S * nj + ((S * nj + 1) & ~1) + SMX * SMX + SMX + ((S * H + 1) & ~1) + (small ? 0 : plan->ah_max_lc * S);
    const int chan_big = 16 * njs + ((S * nj + 1) & ~1) + SMX * SMX + SMX + 8 * Hpad + ((plan->ah_max_lc * S + 1) & ~1) + nj;
    const int chan_words = bigch ? chan_big : chan_small;
    L.ywords = std::max(std::max(H > 0 ? TB * Hpad : 0, chan_words), L.ltile * L.wstride);
    L.Y = take(L.ywords);
    L.S = take(npk); L.r = take(ng); L.z = take(ng); L.linv = take(ng);
    L.dinv = take(std::max(nl, 1)); L.gl = take(std::max(nl, 1)); L.hl = take(std::max(nl, 1));
    L.red = take(32);
    const int npj = nj * (nj + 1) / 2;
    L.lowrank = lowrank ? 1 : 0;
    int n_int = 3 * TBv + ng + nl + 8;
    if (lowrank) {
        L.DLall = take(NSEG * S * nj); L.Qc = take(NSEG * S * S); L.vq = take(NSEG * S); L.Tq = take(NSEG * S * nj);
        L.uq = take(NSEG * S);
        L.ppos = take((npj + 3) / 4);
        n_int += 2 * pl.ah_nbl + nj + (2 * npj + 3) / 4 + 1;
    }
    {
        const int nt8max = std::max(std::max((H + 7) / 8, (ng + 7) / 8), (nj + 7) / 8) + 1;
        L.tdec = take((nt8max * (nt8max + 1) / 2 + 3) / 4 + 1);
    }
    L.ints = take((n_int + 1) / 2 + 1);
    L.small = small ? 1 : 0;
    // external block (shared memory when small, per-CTA global scratch else); SegM | SegG | SegN | LOCw contiguous
    int e = small ? o : 0;
    auto take_e = [&](int n) { const int at = e; e += (n + 1) & ~1; return at; };
    L.Sorig = take_e(npk); L.Sred = take_e(npk); L.C = take_e(lowrank ? 2 : std::max(nl * L.ngs, 1));
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
    const int nj = pl.ah_nsfp;
    const int nt8 = (H + 7) / 8, ntile = nt8 * (nt8 + 1) / 2;

    // CTA size and kernel instance: small models run many narrow CTAs per SM (latency is hidden by
    // other fits), models whose reduced matrix fills the shared memory of an SM run one wide CTA
    std::vector<SoloVariant> variants;
    {
        const SoloVariant* v128 = nullptr;
        const int n128 = solo128_variants(&v128);
        variants.assign(v128, v128 + n128);
        variants.push_back({256, 0, solo256::fit_solo_kernel<0>});
        variants.push_back({256, 1, solo256::fit_solo_kernel<1>});
        variants.push_back({256, 4, solo256::fit_solo_kernel<4>});
        variants.push_back({256, 16, solo256::fit_solo_kernel<16>});
    }
    int force_nt = 0;
    if (const char* e = getenv("HF_FIT_SOLO_NT")) force_nt = atoi(e);  // tuning knob (tools/)
    const SoloVariant* var = nullptr;
    SoloLayout L;
    bool bigch = false;
    int npart = 1, maxt = 0;
    const size_t small_limit = 100 * 1024;  // two or more CTAs per SM
    for (int pass = 0; pass < 2 && !var; ++pass) {
        // pass 0: the preferred CTA size for the placement; pass 1: anything that fits
        for (const SoloVariant& v : variants) {
            if (force_nt && v.nt != force_nt) continue;
            const int warps = v.nt / 32;
            if (warps < pl.R || v.nt < SMX * SMX + pl.S) continue;
            const int need = (H == 0) ? 0 : (ntile + warps - 1) / warps;
            if (need > v.maxt || (need == 0) != (v.maxt == 0)) continue;
            if (H > 0 && Hc > v.nt) continue;
            const int np = (H > 0) ? std::max(1, v.nt / Hc) : 1;
            const bool bc = v.maxt >= 8 && nj >= 16 && nj <= 120;
            // (low-rank elimination of the per-bin block: no histosys, index tables within their integer widths)
            const bool lr = H == 0 && nl > 0 && nl == pl.ah_nbl && nj > 0 && nj <= 255 && ng <= 360 &&
                            pl.NSEG * pl.S * nj <= 4096 && !getenv("HF_FIT_SOLO_DENSE");
            SoloLayout Ls = solo_layout(plan, ng, nl, Hpad, nj, Hc, np, true, bc, v.nt, lr);
            const bool small = sizeof(double) * (size_t)Ls.total <= small_limit;
            if (!small) Ls = solo_layout(plan, ng, nl, Hpad, nj, Hc, np, false, bc, v.nt, false);
            if (sizeof(double) * (size_t)Ls.total + 16 > 225 * 1024) continue;
            if (pass == 0 && !force_nt && v.nt != (small ? 128 : 256)) continue;
            var = &v; L = Ls; bigch = bc; npart = np; maxt = v.maxt;
            break;
        }
    }
    if (!var) return HF_OK;  // the lock-step driver takes the fit
    const int NT = var->nt;
    const size_t smem = sizeof(double) * (size_t)L.total + 16;

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
    a.ngs = L.ngs;
    a.bigch = bigch ? 1 : 0;
    a.njs = (nj + 7) / 8 * 8;
    while (a.njs % 16 != 8) a.njs += 8;
    a.max_lc = plan->ah_max_lc;
    {
        const int nt8max = std::max(std::max((H + 7) / 8, (ng + 7) / 8), (nj + 7) / 8) + 1;
        a.ntdec = nt8max * (nt8max + 1) / 2;
    }

    void (*kern)(DevPlan, SoloArgs) = var->kern;
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
    if (opts.verbose >= 1)
        fprintf(stderr, "[hf_fit solo] %lld fits: grid %lld x %d threads, %d CTAs/SM, %zu bytes shared memory (%s placement), "
                        "ng %d nl %d, tensor-core channel blocks %d, histosys tiles per warp %d, low-rank per-bin block %d\n",
                (long long)N, (long long)grid, NT, occ, smem, L.small ? "small" : "big", ng, nl, a.bigch, maxt, L.lowrank);
    solo_zero_kernel<<<1, 1, 0, st>>>(a.next);
    kern<<<(unsigned)grid, NT, smem, st>>>(pl, a);
    plan->launches += 2;
    HF_CUDA_OK(cudaGetLastError());
    plan->solo_last_grid = (int)grid;
    plan->solo_last_occ = occ;
    plan->solo_last_smem = (int)smem;
    plan->solo_last_nt = NT;
    plan->solo_last_lowrank = L.lowrank;
    *handled = 1;
    return HF_OK;
}
