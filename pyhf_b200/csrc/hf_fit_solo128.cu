// 128-thread kernel instances of the per-fit minimiser (small models: many narrow CTAs per SM); see
// hf_fit_solo.cu.  A translation unit of its own so that the two CTA sizes compile in parallel.
#include "hf_fit_solo_defs.cuh"

namespace solo {
#define SOLO_NT 128
#define SOLO_NS solo128
#include "hf_fit_solo_impl.cuh"
#undef SOLO_NT
#undef SOLO_NS

int solo128_variants(const SoloVariant** out) {
    static const SoloVariant v[] = {
        {128, 0, solo128::fit_solo_kernel<0>}, {128, 1, solo128::fit_solo_kernel<1>},
        {128, 2, solo128::fit_solo_kernel<2>}, {128, 4, solo128::fit_solo_kernel<4>},
    };
    *out = v;
    return 4;
}
}  // namespace solo
