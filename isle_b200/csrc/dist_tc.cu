// dist_tc.cu -- tcgen05 split-TF32 distance contraction (placeholder until the kernel lands;
// distance_pass() falls back to the SIMT engine in kmeans.cu while this reports unsupported).
#include "common.cuh"

namespace isle {

bool dist_tc_supported(const Ctx &, uint32_t, uint32_t) { return false; }

void dist_tc_launch(Ctx &, const float *, const float *, uint32_t, uint32_t, const float *, const float *, uint32_t,
                    int, uint32_t *, float *)
{
    throw Error(ISLE_ERR_ARG, "dist_tc: not built");
}

}  // namespace isle
