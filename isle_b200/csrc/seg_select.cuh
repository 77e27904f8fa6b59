// seg_select.cuh -- exact k-th largest float per segment by a four-round radix select whose per-round histograms add
// across ranks: the document-sharded form of the segmented selects of catchwords.cu and topic_model.cu (a segment's
// values live on several ranks; no values travel, only nseg x 256 counters per round).
//
//   ordered(x)   monotone map float -> uint32 (larger float, larger integer)
//   round q      hist[s][byte q of ordered(x)] += 1 for the values of segment s whose higher bytes equal prefix[s]
//   (allreduce)  hist summed over ranks
//   pick         walk the 256 bins from the top until the remaining rank falls inside one: prefix[s] = prefix[s] << 8 | bin
// After four rounds prefix[s] = ordered(the (kth[s] + 1)-th largest value of segment s).  kth[s] = 0xFFFFFFFF marks a
// segment without a selection (its prefix stays 0).
#pragma once

#include "common.cuh"

namespace isle {
namespace segsel {

__device__ __forceinline__ uint32_t ordered(float x)
{
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unordered(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// does value bits `o` still match the bytes selected in earlier rounds?
__device__ __forceinline__ bool matches(uint32_t o, uint32_t prefix, int round)
{
    return round == 0 || (o >> (32 - 8 * round)) == prefix;
}

static __global__ void __launch_bounds__(256)
hist_pairs_kernel(const uint32_t *__restrict__ seg, const float *__restrict__ val, int64_t n, const uint32_t *__restrict__ prefix,
                  int round, uint32_t *__restrict__ hist)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t s = seg[i], o = ordered(val[i]);
        if (matches(o, prefix[s], round)) atomicAdd(hist + (size_t)s * 256 + ((o >> (24 - 8 * round)) & 255u), 1u);
    }
}

// one thread per segment; kth is the remaining 0-based rank from the top, updated in place
static __global__ void __launch_bounds__(128)
pick_kernel(const uint32_t *__restrict__ hist, uint32_t nseg, uint32_t *__restrict__ kth, uint32_t *__restrict__ prefix)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint32_t rem = kth[s];
    if (rem == 0xFFFFFFFFu) return;
    const uint32_t *h = hist + (size_t)s * 256;
    int bin = 255;
    for (; bin > 0; --bin) {
        const uint32_t n = h[bin];
        if (rem < n) break;
        rem -= n;
    }
    prefix[s] = (prefix[s] << 8) | (uint32_t)bin;
    kth[s] = rem;
}

}  // namespace segsel
}  // namespace isle
