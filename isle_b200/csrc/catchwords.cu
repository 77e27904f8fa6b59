// catchwords.cu -- SURVEY 8(f) row 2, first half: per-cluster catchword thresholds and the catchword test.
//
// Replaces SparseMatrix::rth_highest_element (reference src/sparseMatrix.cpp:491-524, called once per topic under
// pfor from src/trainer.cpp:587-589) and SparseMatrix::find_catchwords (:573-594, src/trainer.cpp:635).
//
//   thresholds[t][w] = the r-th highest normalised value of word w among the documents of cluster t when w occurs in
//                      MORE than r of them; otherwise 0 -- except when r >= |cluster t| and w occurs in every document
//                      of the cluster: then the smallest value.
//   w is a catchword of t  iff  thresholds[t][w] > rho * thresholds[o][w] for every other topic o.
//
// The reference collects a std::vector per word and sorts each; it is the same segmented select as the per-word
// threshold of kernel family (1), with the cluster as an extra key and un-rounded values:
//   1. count[t][w]   (u32 atomics over the clusters' documents of A, doc-major)
//   2. candidates = segments that can yield a non-zero threshold (count > r, or the all-documents case); exclusive
//      scans give every candidate a slot and a value range
//   3. a second pass over the same documents scatters the candidates' values into their ranges
//   4. cub::DeviceSegmentedSort (descending) and one pick per candidate
// Selection of floats is exact, so the result is bit-identical to the reference's.  find_catchwords is O(V k) on the
// device (largest and second largest threshold per word) instead of the reference's O(V k^2) serial loop.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "seg_select.cuh"

namespace isle {

namespace {

constexpr uint32_t kNoCluster = 0xFFFFFFFFu;

// warp per listed document: count[cl V + w] += 1
__global__ void __launch_bounds__(256)
cw_count_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
                const uint32_t *__restrict__ doc_ids, const uint32_t *__restrict__ cl_ids, uint32_t ndocs, uint32_t V,
                uint32_t *__restrict__ count, uint32_t *__restrict__ cl_size)
{
    (void)a_val;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; i < ndocs; i += nw) {
        const uint32_t d = doc_ids ? doc_ids[i] : i;
        const uint32_t cl = cl_ids ? cl_ids[i] : 0u;
        if (cl == kNoCluster) continue;
        const int64_t b = a_off[d], e = a_off[d + 1];
        uint32_t *row = count + (size_t)cl * V;
        for (int64_t p = b + lane; p < e; p += 32) atomicAdd(row + a_row[p], 1u);
        if (lane == 0) atomicAdd(cl_size + cl, 1u);
    }
}

// flag[key] = 1 and take[key] = count when the segment can yield a non-zero threshold
__global__ void __launch_bounds__(256)
cw_candidates_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ cl_size, size_t nkeys, uint32_t V, uint32_t r,
                     uint32_t *__restrict__ flag, uint32_t *__restrict__ take)
{
    const size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    const uint32_t n = count[key], sz = cl_size[key / V];
    const bool cand = n > r || (r >= sz && n == sz && n > 0);        // src/sparseMatrix.cpp:508-520
    flag[key] = cand ? 1u : 0u;
    take[key] = cand ? n : 0u;
}

// segment table of the candidates: begin/end of their value ranges, their key
__global__ void __launch_bounds__(256)
cw_segments_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ slot, const int64_t *__restrict__ voff,
                   const uint32_t *__restrict__ count, size_t nkeys, int64_t *__restrict__ seg_begin, int64_t *__restrict__ seg_end,
                   uint32_t *__restrict__ seg_key)
{
    const size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys || !flag[key]) return;
    const uint32_t s = slot[key];
    seg_begin[s] = voff[key];
    seg_end[s] = voff[key] + count[key];
    seg_key[s] = (uint32_t)key;
}

// second pass over the same documents: values of candidate segments go to their ranges (order inside a range is free)
__global__ void __launch_bounds__(256)
cw_scatter_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
                  const uint32_t *__restrict__ doc_ids, const uint32_t *__restrict__ cl_ids, uint32_t ndocs, uint32_t V,
                  const uint32_t *__restrict__ flag, const int64_t *__restrict__ voff, uint32_t *__restrict__ fill,
                  float *__restrict__ seg_vals)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; i < ndocs; i += nw) {
        const uint32_t d = doc_ids ? doc_ids[i] : i;
        const uint32_t cl = cl_ids ? cl_ids[i] : 0u;
        if (cl == kNoCluster) continue;
        const int64_t b = a_off[d], e = a_off[d + 1];
        for (int64_t p = b + lane; p < e; p += 32) {
            const size_t key = (size_t)cl * V + a_row[p];
            if (flag[key]) seg_vals[voff[key] + atomicAdd(fill + key, 1u)] = a_val[p];
        }
    }
}

// thresholds[key] = sorted (descending) value r-1 of the segment, or its last (smallest) one in the all-documents case
__global__ void __launch_bounds__(256)
cw_pick_kernel(const float *__restrict__ sorted, const int64_t *__restrict__ seg_begin, const int64_t *__restrict__ seg_end,
               const uint32_t *__restrict__ seg_key, uint32_t nseg, uint32_t r, float *__restrict__ thresholds)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int64_t b = seg_begin[s], n = seg_end[s] - b;
    thresholds[seg_key[s]] = sorted[b + (n > (int64_t)r ? (int64_t)r - 1 : n - 1)];
}

// ---- document-sharded form: the candidates' values stay where their documents are; four radix rounds of
// (local histogram over the rank's documents, allreduce, pick) find every segment's threshold exactly
__global__ void __launch_bounds__(256)
cw_hist_docs_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
                    const uint32_t *__restrict__ doc_ids, const uint32_t *__restrict__ cl_ids, uint32_t ndocs, uint32_t V,
                    const uint32_t *__restrict__ flag, const uint32_t *__restrict__ slot, const uint32_t *__restrict__ prefix, int round,
                    uint32_t *__restrict__ hist)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; i < ndocs; i += nw) {
        const uint32_t d = doc_ids ? doc_ids[i] : i;
        const uint32_t cl = cl_ids ? cl_ids[i] : 0u;
        if (cl == kNoCluster) continue;
        for (int64_t p = a_off[d] + lane, e = a_off[d + 1]; p < e; p += 32) {
            const size_t key = (size_t)cl * V + a_row[p];
            if (!flag[key]) continue;
            const uint32_t s = slot[key], o = segsel::ordered(a_val[p]);
            if (segsel::matches(o, prefix[s], round)) atomicAdd(hist + (size_t)s * 256 + ((o >> (24 - 8 * round)) & 255u), 1u);
        }
    }
}

// kth[slot] = 0-based rank from the top of the value wanted from the segment (:508-520); seg_key[slot] = key
__global__ void __launch_bounds__(256)
cw_kth_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ slot, const uint32_t *__restrict__ count, size_t nkeys,
              uint32_t r, uint32_t *__restrict__ kth, uint32_t *__restrict__ seg_key)
{
    const size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys || !flag[key]) return;
    const uint32_t s = slot[key], n = count[key];
    kth[s] = n > r ? r - 1 : n - 1;
    seg_key[s] = (uint32_t)key;
}

__global__ void __launch_bounds__(256)
cw_result_kernel(const uint32_t *__restrict__ prefix, const uint32_t *__restrict__ seg_key, uint32_t nseg, float *__restrict__ thresholds)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nseg) thresholds[seg_key[s]] = segsel::unordered(prefix[s]);
}

struct U32ToI64 {
    __host__ __device__ int64_t operator()(uint32_t x) const { return (int64_t)x; }
};

// topic_of_word[w] = the topic whose threshold beats rho x every other topic's, or -1 (src/sparseMatrix.cpp:573-594).
// thr[t] > rho thr[o] for all o != t  <=>  t is the arg max and thr[t] > rho x (second largest), evaluated in double as the
// reference's float > double * float expression does.
__global__ void __launch_bounds__(256)
find_catchwords_kernel(const float *__restrict__ thr, uint32_t V, uint32_t k, double rho, int32_t *__restrict__ topic_of_word)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V) return;
    float m1 = -1.f, m2 = -1.f;
    int32_t t1 = -1;
    for (uint32_t t = 0; t < k; ++t) {
        const float v = thr[(size_t)t * V + w];
        if (v > m1) { m2 = m1; m1 = v; t1 = (int32_t)t; }
        else if (v > m2) m2 = v;
    }
    const bool is_catch = k > 1 && (double)m1 > rho * (double)m2;      // k == 1: the reference's loop never sets the flag
    topic_of_word[w] = is_catch ? t1 : -1;
}

// thresholds (device, k x V, zeroed here) for the documents listed in doc_ids (NULL: all D documents) with clusters cl_ids
void thresholds_device(Ctx &c, uint32_t k, uint32_t r, const uint32_t *doc_ids, const uint32_t *cl_ids, uint32_t ndocs, float *thr)
{
    const uint32_t V = (uint32_t)c.V;
    const size_t nkeys = (size_t)k * V;
    ISLE_CUDA_CHECK(cudaMemsetAsync(thr, 0, nkeys * sizeof(float), c.stream));
    c.counters["cw_candidates"] = 0.0;
    if (!ndocs && c.world == 1) return;      // a sharded rank without documents still takes part in the collectives
    DevBuf<uint32_t> count(nkeys), cl_size(k), flag(nkeys), take(nkeys), slot(nkeys);
    DevBuf<int64_t> voff(nkeys + 1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(count.p, 0, count.bytes(), c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(cl_size.p, 0, cl_size.bytes(), c.stream));
    const unsigned wgrid = grid_for((size_t)ndocs * 32, 256, c.num_sms * 8);
    {
        StatScope s(c, "cw_count", (double)c.nnzA * 4.0);
        if (ndocs) cw_count_kernel<<<wgrid, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, doc_ids, cl_ids, ndocs, V, count.p, cl_size.p);
        count_launch(c);
    }
    if (c.world > 1) {      // clusters span ranks: counts and sizes are global from here on
        allreduce_sum_u32(c, count.p, count.n);
        allreduce_sum_u32(c, cl_size.p, cl_size.n);
    }
    cw_candidates_kernel<<<(unsigned)((nkeys + 255) / 256), 256, 0, c.stream>>>(count.p, cl_size.p, nkeys, V, r, flag.p, take.p);
    count_launch(c);
    {
        size_t t1 = 0, t2 = 0;
        auto it = cub::TransformInputIterator<int64_t, U32ToI64, uint32_t *>(take.p, U32ToI64());
        cub::DeviceScan::ExclusiveSum(nullptr, t1, flag.p, slot.p, (int)nkeys, c.stream);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, it, voff.p, (int)nkeys, c.stream);
        DevBuf<uint8_t> tmp(std::max(t1, t2));
        ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, t1, flag.p, slot.p, (int)nkeys, c.stream));
        ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, t2, it, voff.p, (int)nkeys, c.stream));
        count_launch(c, 2);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));   // tmp is released on scope exit
    }
    uint32_t last_slot = 0, last_flag = 0, last_take = 0;
    int64_t last_off = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_slot, slot.p + nkeys - 1, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_flag, flag.p + nkeys - 1, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_take, take.p + nkeys - 1, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_off, voff.p + nkeys - 1, 8, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    const uint32_t nseg = last_slot + last_flag;
    const int64_t nvals = last_off + last_take;
    c.counters["cw_candidates"] = nseg;
    if (!nseg) return;
    if (c.world > 1) {
        // the values of a segment live on several ranks: exact distributed radix select (seg_select.cuh), no values travel
        DevBuf<uint32_t> kth(nseg), prefix(nseg), skey(nseg), hist((size_t)nseg * 256);
        ISLE_CUDA_CHECK(cudaMemsetAsync(prefix.p, 0, prefix.bytes(), c.stream));
        cw_kth_kernel<<<(unsigned)((nkeys + 255) / 256), 256, 0, c.stream>>>(flag.p, slot.p, count.p, nkeys, r, kth.p, skey.p);
        count_launch(c);
        for (int round = 0; round < 4; ++round) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(hist.p, 0, hist.bytes(), c.stream));
            if (ndocs) {
                cw_hist_docs_kernel<<<wgrid, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, doc_ids, cl_ids, ndocs, V, flag.p, slot.p,
                                                                 prefix.p, round, hist.p);
                count_launch(c);
            }
            allreduce_sum_u32(c, hist.p, hist.n);
            segsel::pick_kernel<<<(nseg + 127) / 128, 128, 0, c.stream>>>(hist.p, nseg, kth.p, prefix.p);
            count_launch(c);
        }
        cw_result_kernel<<<(nseg + 255) / 256, 256, 0, c.stream>>>(prefix.p, skey.p, nseg, thr);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return;
    }
    ISLE_REQUIRE(nvals < (int64_t)1 << 31, ISLE_ERR_RANGE, "catchword thresholds: more than 2^31 candidate values");
    DevBuf<int64_t> seg_begin(nseg), seg_end(nseg);
    DevBuf<uint32_t> seg_key(nseg);
    DevBuf<float> vals((size_t)nvals), sorted((size_t)nvals);
    cw_segments_kernel<<<(unsigned)((nkeys + 255) / 256), 256, 0, c.stream>>>(flag.p, slot.p, voff.p, count.p, nkeys, seg_begin.p,
                                                                           seg_end.p, seg_key.p);
    ISLE_CUDA_CHECK(cudaMemsetAsync(take.p, 0, take.bytes(), c.stream));    // reused as the per-segment fill counters
    {
        StatScope s(c, "cw_scatter", (double)c.nnzA * 8.0);
        cw_scatter_kernel<<<wgrid, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, doc_ids, cl_ids, ndocs, V, flag.p, voff.p, take.p,
                                                       vals.p);
    }
    count_launch(c, 2);
    {
        StatScope s(c, "cw_sort", (double)nvals * 8.0);
        size_t tb = 0;
        cub::DeviceSegmentedSort::SortKeysDescending(nullptr, tb, vals.p, sorted.p, (int)nvals, (int)nseg, seg_begin.p, seg_end.p, c.stream);
        DevBuf<uint8_t> tmp(tb);
        ISLE_CUDA_CHECK(cub::DeviceSegmentedSort::SortKeysDescending(tmp.p, tb, vals.p, sorted.p, (int)nvals, (int)nseg, seg_begin.p,
                                                                     seg_end.p, c.stream));
        count_launch(c);
        cw_pick_kernel<<<(nseg + 255) / 256, 256, 0, c.stream>>>(sorted.p, seg_begin.p, seg_end.p, seg_key.p, nseg, r, thr);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
}

}  // namespace

// All clusters at once: cluster_of_doc[d] in [0, k) or 0xFFFFFFFF for the D documents of A (original ids).
void catchword_thresholds(Ctx &c, uint64_t k64, uint64_t r, const uint32_t *cluster_of_doc_host, float *thresholds_out)
{
    ISLE_REQUIRE(c.a_off.p != nullptr && c.V > 0, ISLE_ERR_ARG, "catchword_thresholds: upload_A first");
    ISLE_REQUIRE(k64 >= 1 && k64 * c.V < ((uint64_t)1 << 31) && r >= 1 && cluster_of_doc_host, ISLE_ERR_ARG,
                 "catchword_thresholds: bad arguments (k V must stay below 2^31, r >= 1)");
    const uint32_t k = (uint32_t)k64, D = (uint32_t)c.D;
    DevBuf<uint32_t> cl(std::max<uint32_t>(D, 1));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(cl.p, cluster_of_doc_host, (size_t)D * 4, cudaMemcpyHostToDevice, c.stream));
    c.catch_thr.alloc((size_t)k * c.V);
    c.catch_k = k64;
    StatScope s(c, "catch_thresholds", (double)c.nnzA * 16.0);
    thresholds_device(c, k, (uint32_t)std::min<uint64_t>(r, 0xFFFFFFFEu), nullptr, cl.p, D, c.catch_thr.p);
    if (thresholds_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(thresholds_out, c.catch_thr.p, c.catch_thr.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// One cluster, the reference's calling convention (rth_highest_element): thresholds_out has V entries.
void rth_highest_element(Ctx &c, uint64_t r, const uint64_t *docs_host, uint64_t ndocs, float *thresholds_out)
{
    ISLE_REQUIRE(c.a_off.p != nullptr && c.V > 0, ISLE_ERR_ARG, "rth_highest_element: upload_A first");
    ISLE_REQUIRE(r >= 1 && thresholds_out && (docs_host || !ndocs), ISLE_ERR_ARG, "rth_highest_element: bad arguments (r >= 1)");
    std::vector<uint32_t> ids((size_t)ndocs);
    for (uint64_t i = 0; i < ndocs; ++i) {
        ISLE_REQUIRE(docs_host[i] < c.D, ISLE_ERR_RANGE, "rth_highest_element: document id out of range");
        ids[i] = (uint32_t)docs_host[i];
    }
    DevBuf<uint32_t> dids(std::max<uint64_t>(ndocs, 1));
    DevBuf<float> thr((size_t)c.V);
    if (ndocs) ISLE_CUDA_CHECK(cudaMemcpyAsync(dids.p, ids.data(), (size_t)ndocs * 4, cudaMemcpyHostToDevice, c.stream));
    thresholds_device(c, 1, (uint32_t)std::min<uint64_t>(r, 0xFFFFFFFEu), dids.p, nullptr, (uint32_t)ndocs, thr.p);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(thresholds_out, thr.p, thr.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// thresholds_host: k x V topic-major, or NULL = the matrix the last catchword_thresholds call left on the device.
void find_catchwords(Ctx &c, uint64_t k64, const float *thresholds_host, double rho, int32_t *topic_of_word_out)
{
    ISLE_REQUIRE(k64 >= 1 && topic_of_word_out && c.V > 0, ISLE_ERR_ARG, "find_catchwords: bad arguments");
    const uint32_t k = (uint32_t)k64, V = (uint32_t)c.V;
    DevBuf<float> own;
    const float *thr = nullptr;
    if (thresholds_host) {
        own.alloc((size_t)k * V);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(own.p, thresholds_host, own.bytes(), cudaMemcpyHostToDevice, c.stream));
        thr = own.p;
    } else {
        ISLE_REQUIRE(c.catch_thr.p && c.catch_k == k64, ISLE_ERR_ARG, "find_catchwords: no device thresholds of that width");
        thr = c.catch_thr.p;
    }
    DevBuf<int32_t> tw(V);
    {
        StatScope s(c, "find_catchwords", (double)k * V * 4.0);
        find_catchwords_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(thr, V, k, rho, tw.p);
        count_launch(c);
    }
    ISLE_CUDA_CHECK(cudaMemcpyAsync(topic_of_word_out, tw.p, (size_t)V * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

}  // namespace isle
