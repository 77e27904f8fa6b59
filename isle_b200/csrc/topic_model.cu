// topic_model.cu -- SURVEY 8(f) row 2, second half: SparseMatrix::construct_topic_model
// (reference src/sparseMatrix.cpp:597-838, called from src/trainer.cpp:645-651).
//
//   doc_topic_sum[d][t] = fp32 sum, in position order, of document d's normalised values on the catchwords of topic t
//                         (:652-668); the non-zero sums are listed by (document, topic)                        (:669-678)
//   model_threshold[t]  = the rank_threshold-th largest sum of topic t, 0 when the topic has fewer entries or no
//                         catchwords                                                                            (:716-751)
//   Model[:, t]        += every document whose sum for t exceeds the threshold, plus every document of cluster t
//                         (the reference's `doc_in_catchless_topic` holds the cluster of EVERY clustered document)
//                                                                                                               (:787-817)
//   Model[:, t]        *= 1 / ||Model[:, t]||_1                                                                 (:822-826)
//
// Device form.  The sums must be bit-identical to the reference's (they are compared with `>` against a selected
// one), so one thread walks one document in position order into a k-wide row of a blocked scratch matrix, exactly
// the reference's DocTopicSumArray (2^18 documents x k floats per block); the non-zero entries are compacted in
// (document, topic) order with an exclusive scan; the thresholds are a segmented select again: one radix sort of
// 64-bit keys (topic << 32 | ~bits(sum)) and one pick per topic; the accumulation is a warp per selected
// (document, topic) pair doing float atomics into the V x k model (order of additions differs from the reference's
// document order: the model agrees to fp32 rounding, ~1e-7 relative, everything before it exactly).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "seg_select.cuh"

namespace isle {

namespace {

constexpr uint32_t kNoCluster = 0xFFFFFFFFu;
constexpr uint32_t kDocBlock = 1u << 18;      // DOC_BLOCK_SIZE (include/hyperparams.h)

// one thread per document of the block: arr[(d - d0) k + t] += val for every entry on a catchword of t, in position order
__global__ void __launch_bounds__(256)
tm_sums_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
               const int32_t *__restrict__ topic_of_word, uint32_t d0, uint32_t d1, uint32_t k, float *__restrict__ arr)
{
    const uint32_t d = d0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= d1) return;
    float *row = arr + (size_t)(d - d0) * k;
    const int64_t b = a_off[d], e = a_off[d + 1];
    for (int64_t p = b; p < e; ++p) {
        const int32_t t = __ldg(topic_of_word + a_row[p]);
        if (t >= 0) row[t] = __fadd_rn(row[t], a_val[p]);
    }
}

// non-zero sums per document (warp per document)
__global__ void __launch_bounds__(256)
tm_count_kernel(const float *__restrict__ arr, uint32_t ndocs, uint32_t k, uint32_t *__restrict__ cnt)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= ndocs) return;
    const float *row = arr + (size_t)i * k;
    uint32_t n = 0;
    for (uint32_t t = lane; t < k; t += 32) n += row[t] != 0.0f ? 1u : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cnt[i] = n;
}

// entries of the block in (document, topic) order (warp per document, ordered ballot compaction)
__global__ void __launch_bounds__(256)
tm_emit_kernel(const float *__restrict__ arr, uint32_t d0, uint32_t ndocs, uint32_t k, const int64_t *__restrict__ off,
               int64_t base, uint32_t *__restrict__ e_doc, uint32_t *__restrict__ e_topic, float *__restrict__ e_val)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= ndocs) return;
    const float *row = arr + (size_t)i * k;
    int64_t out = base + off[i];
    for (uint32_t t0 = 0; t0 < k; t0 += 32) {
        const uint32_t t = t0 + lane;
        const float v = t < k ? row[t] : 0.0f;
        const uint32_t m = __ballot_sync(0xffffffffu, v != 0.0f);
        if (v != 0.0f) {
            const int64_t o = out + __popc(m & ((1u << lane) - 1u));
            e_doc[o] = d0 + i; e_topic[o] = t; e_val[o] = v;
        }
        out += __popc(m);
    }
}

struct U32ToI64 {
    __host__ __device__ int64_t operator()(uint32_t x) const { return (int64_t)x; }
};

// key = topic << 32 | ~bits(sum): ascending keys = topic ascending, sum descending (sums are positive)
__global__ void __launch_bounds__(256)
tm_keys_kernel(const uint32_t *__restrict__ e_topic, const float *__restrict__ e_val, int64_t n, unsigned long long *__restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t bits = __float_as_uint(e_val[i]);
    const uint32_t ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);      // total order of floats as unsigned
    keys[i] = ((unsigned long long)e_topic[i] << 32) | (unsigned long long)(~ord);
}

// thr[t] = rank-th largest sum of topic t when it has catchwords and at least `rank` entries, else 0
__global__ void tm_threshold_kernel(const unsigned long long *__restrict__ sorted, int64_t n, uint32_t k, uint64_t rank,
                                    const uint32_t *__restrict__ has_catch, float *__restrict__ thr)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    auto lower = [&](unsigned long long key) {
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const int64_t b = lower((unsigned long long)t << 32), e = lower((unsigned long long)(t + 1) << 32);
    float v = 0.0f;
    if (has_catch[t] && rank >= 1 && (uint64_t)(e - b) >= rank) {
        const uint32_t ord = ~(uint32_t)(sorted[b + (int64_t)rank - 1] & 0xFFFFFFFFull);
        const uint32_t bits = (ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord;
        v = __uint_as_float(bits);
    }
    thr[t] = v;
}

// document-sharded thresholds: entries per topic (local), then the selection set-up from the global counts
__global__ void __launch_bounds__(256)
tm_topic_count_kernel(const uint32_t *__restrict__ e_topic, int64_t n, uint32_t *__restrict__ cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(cnt + e_topic[i], 1u);
}

__global__ void tm_kth_kernel(const uint32_t *__restrict__ gcount, const uint32_t *__restrict__ has_catch, uint32_t k, uint64_t rank,
                              uint32_t *__restrict__ kth)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    kth[t] = (has_catch[t] && rank >= 1 && (uint64_t)gcount[t] >= rank) ? (uint32_t)(rank - 1) : 0xFFFFFFFFu;
}

__global__ void tm_thr_from_prefix_kernel(const uint32_t *__restrict__ prefix, const uint32_t *__restrict__ gcount,
                                          const uint32_t *__restrict__ has_catch, uint32_t k, uint64_t rank, float *__restrict__ thr)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    thr[t] = (has_catch[t] && rank >= 1 && (uint64_t)gcount[t] >= rank) ? segsel::unordered(prefix[t]) : 0.0f;
}

__global__ void tm_has_catch_kernel(const int32_t *__restrict__ topic_of_word, uint32_t V, uint32_t k, uint32_t *__restrict__ has_catch)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V) return;
    const int32_t t = topic_of_word[w];
    if (t >= 0 && (uint32_t)t < k) has_catch[t] = 1u;
}

// warp per entry: Model[:, topic] += column(doc) when the sum exceeds the topic's threshold (:799-806)
__global__ void __launch_bounds__(256)
tm_add_entries_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
                      const uint32_t *__restrict__ e_doc, const uint32_t *__restrict__ e_topic, const float *__restrict__ e_val,
                      int64_t n, const float *__restrict__ thr, uint32_t V, float *__restrict__ model)
{
    const uint32_t lane = threadIdx.x & 31;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; i < n; i += nw) {
        const uint32_t t = e_topic[i];
        if (!(e_val[i] > thr[t])) continue;
        const uint32_t d = e_doc[i];
        float *col = model + (size_t)t * V;
        for (int64_t p = a_off[d] + lane, e = a_off[d + 1]; p < e; p += 32) atomicAdd(col + a_row[p], a_val[p]);
    }
}

// warp per document: Model[:, cluster(doc)] += column(doc) (:807-809)
__global__ void __launch_bounds__(256)
tm_add_clusters_kernel(const float *__restrict__ a_val, const uint32_t *__restrict__ a_row, const int64_t *__restrict__ a_off,
                       const uint32_t *__restrict__ cluster_of_doc, uint32_t D, uint32_t V, uint32_t k, float *__restrict__ model)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < D; d += nw) {
        const uint32_t t = cluster_of_doc[d];
        if (t == kNoCluster || t >= k) continue;
        float *col = model + (size_t)t * V;
        for (int64_t p = a_off[d] + lane, e = a_off[d + 1]; p < e; p += 32) atomicAdd(col + a_row[p], a_val[p]);
    }
}

// column t *= (float)(1.0 / (double)asum(column t))   (FPasum + FPscal, :822-826); one CTA per topic
__global__ void __launch_bounds__(256)
tm_normalize_kernel(float *__restrict__ model, uint32_t V)
{
    __shared__ float red[8];
    __shared__ float alpha;
    float *col = model + (size_t)blockIdx.x * V;
    float s = 0.f;
    for (uint32_t w = threadIdx.x; w < V; w += 256) s += fabsf(col[w]);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        alpha = (float)(1.0 / (double)t);
    }
    __syncthreads();
    const float a = alpha;
    for (uint32_t w = threadIdx.x; w < V; w += 256) col[w] *= a;
}

}  // namespace

void construct_topic_model(Ctx &c, uint64_t k64, const int32_t *topic_of_word_host, const uint32_t *cluster_of_doc_host,
                           uint64_t rank_threshold, float *model_out, uint64_t *num_entries_out)
{
    ISLE_REQUIRE(c.a_off.p != nullptr && c.V > 0, ISLE_ERR_ARG, "construct_topic_model: upload_A first");
    ISLE_REQUIRE(k64 >= 1 && k64 <= 65536 && topic_of_word_host && cluster_of_doc_host, ISLE_ERR_ARG, "construct_topic_model: bad arguments");
    // document-sharded: the catchword sums are local to the documents a rank holds; the per-topic thresholds come from an
    // exact distributed radix select over all ranks' sums (seg_select.cuh); the local model contributions are allreduced
    const uint32_t k = (uint32_t)k64, V = (uint32_t)c.V, D = (uint32_t)c.D;
    StatScope total(c, "topic_model", (double)c.nnzA * 24.0 + (double)V * k * 8.0);
    DevBuf<int32_t> tw(V);
    DevBuf<uint32_t> cl(std::max<uint32_t>(D, 1)), has_catch(k), cnt(kDocBlock);
    DevBuf<int64_t> off(kDocBlock);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(tw.p, topic_of_word_host, (size_t)V * 4, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(cl.p, cluster_of_doc_host, (size_t)D * 4, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(has_catch.p, 0, has_catch.bytes(), c.stream));
    tm_has_catch_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(tw.p, V, k, has_catch.p);
    count_launch(c);

    // ---- pass 1: per block of documents, the sums and how many are non-zero; pass 2 re-computes and emits
    // (re-computing the block costs one more read of A, keeping every block's k-wide scratch costs D x k floats)
    const uint32_t nblocks = (D + kDocBlock - 1) / kDocBlock;
    const uint32_t blk_docs = std::min<uint32_t>(D, kDocBlock);
    DevBuf<float> arr((size_t)std::max<uint32_t>(blk_docs, 1) * k);
    size_t scan_bytes = 0;
    auto it = cub::TransformInputIterator<int64_t, U32ToI64, uint32_t *>(cnt.p, U32ToI64());
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, off.p, (int)kDocBlock, c.stream);
    DevBuf<uint8_t> scan_tmp(scan_bytes);
    std::vector<int64_t> block_base(nblocks + 1, 0);
    auto block_sums = [&](uint32_t blk, uint32_t &d0, uint32_t &nd) {
        d0 = blk * kDocBlock;
        nd = std::min<uint32_t>(kDocBlock, D - d0);
        ISLE_CUDA_CHECK(cudaMemsetAsync(arr.p, 0, (size_t)nd * k * 4, c.stream));
        tm_sums_kernel<<<(nd + 255) / 256, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, tw.p, d0, d0 + nd, k, arr.p);
        tm_count_kernel<<<(unsigned)(((size_t)nd * 32 + 255) / 256), 256, 0, c.stream>>>(arr.p, nd, k, cnt.p);
        ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, it, off.p, (int)nd, c.stream));
        count_launch(c, 3);
    };
    for (uint32_t blk = 0; blk < nblocks; ++blk) {
        uint32_t d0, nd;
        block_sums(blk, d0, nd);
        int64_t last_off = 0;
        uint32_t last_cnt = 0;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_off, off.p + nd - 1, 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&last_cnt, cnt.p + nd - 1, 4, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        block_base[blk + 1] = block_base[blk] + last_off + last_cnt;
    }
    const int64_t n = block_base[nblocks];
    c.tm_doc.alloc((size_t)std::max<int64_t>(n, 1));
    c.tm_topic.alloc((size_t)std::max<int64_t>(n, 1));
    c.tm_val.alloc((size_t)std::max<int64_t>(n, 1));
    c.tm_entries = n;
    for (uint32_t blk = 0; blk < nblocks; ++blk) {
        uint32_t d0, nd;
        if (nblocks > 1 || blk > 0) block_sums(blk, d0, nd);       // a single block is still in `arr`
        else { d0 = 0; nd = D; }
        tm_emit_kernel<<<(unsigned)(((size_t)nd * 32 + 255) / 256), 256, 0, c.stream>>>(arr.p, d0, nd, k, off.p, block_base[blk], c.tm_doc.p,
                                                                                       c.tm_topic.p, c.tm_val.p);
        count_launch(c);
    }

    // ---- model thresholds: rank-th largest sum per topic
    DevBuf<float> thr(k);
    if (c.world > 1) {
        DevBuf<uint32_t> gcount(k), kth(k), prefix(k), hist((size_t)k * 256);
        ISLE_CUDA_CHECK(cudaMemsetAsync(gcount.p, 0, gcount.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(prefix.p, 0, prefix.bytes(), c.stream));
        const unsigned egrid = grid_for((size_t)std::max<int64_t>(n, 1), 256, c.num_sms * 8);
        if (n > 0) tm_topic_count_kernel<<<egrid, 256, 0, c.stream>>>(c.tm_topic.p, n, gcount.p);
        allreduce_sum_u32(c, gcount.p, gcount.n);
        tm_kth_kernel<<<(k + 127) / 128, 128, 0, c.stream>>>(gcount.p, has_catch.p, k, rank_threshold, kth.p);
        count_launch(c, 2);
        for (int round = 0; round < 4; ++round) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(hist.p, 0, hist.bytes(), c.stream));
            if (n > 0) segsel::hist_pairs_kernel<<<egrid, 256, 0, c.stream>>>(c.tm_topic.p, c.tm_val.p, n, prefix.p, round, hist.p);
            allreduce_sum_u32(c, hist.p, hist.n);
            segsel::pick_kernel<<<(k + 127) / 128, 128, 0, c.stream>>>(hist.p, k, kth.p, prefix.p);
            count_launch(c, 2);
        }
        tm_thr_from_prefix_kernel<<<(k + 127) / 128, 128, 0, c.stream>>>(prefix.p, gcount.p, has_catch.p, k, rank_threshold, thr.p);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    } else if (n > 0) {
        DevBuf<unsigned long long> keys((size_t)n), sorted((size_t)n);
        tm_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(c.tm_topic.p, c.tm_val.p, n, keys.p);
        int topic_bits = 1;
        while ((1u << topic_bits) < k) ++topic_bits;
        size_t tb = 0;
        ISLE_REQUIRE(n < ((int64_t)1 << 31), ISLE_ERR_RANGE, "construct_topic_model: more than 2^31 (document, topic) sums");
        cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, sorted.p, (int)n, 0, 32 + topic_bits, c.stream);
        DevBuf<uint8_t> tmp(tb);
        ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, sorted.p, (int)n, 0, 32 + topic_bits, c.stream));
        tm_threshold_kernel<<<(k + 127) / 128, 128, 0, c.stream>>>(sorted.p, n, k, rank_threshold, has_catch.p, thr.p);
        count_launch(c, 3);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    } else {
        ISLE_CUDA_CHECK(cudaMemsetAsync(thr.p, 0, thr.bytes(), c.stream));
    }

    // ---- accumulate and normalise
    c.tm_model.alloc((size_t)V * k);
    ISLE_CUDA_CHECK(cudaMemsetAsync(c.tm_model.p, 0, c.tm_model.bytes(), c.stream));
    if (n > 0) {
        tm_add_entries_kernel<<<grid_for((size_t)n * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(
            c.a_val.p, c.a_row.p, c.a_off.p, c.tm_doc.p, c.tm_topic.p, c.tm_val.p, n, thr.p, V, c.tm_model.p);
        count_launch(c);
    }
    if (D) {
        tm_add_clusters_kernel<<<grid_for((size_t)D * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, cl.p, D, V,
                                                                                               k, c.tm_model.p);
        count_launch(c);
    }
    if (c.world > 1) allreduce_sum_f32(c, c.tm_model.p, c.tm_model.n);
    tm_normalize_kernel<<<k, 256, 0, c.stream>>>(c.tm_model.p, V);
    count_launch(c);
    if (model_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(model_out, c.tm_model.p, c.tm_model.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (num_entries_out) *num_entries_out = (uint64_t)n;
}

// The (document, topic, sum) list of the last construct_topic_model, in (document, topic) order.
void download_doc_topic_sums(Ctx &c, uint32_t *docs, uint32_t *topics, float *sums)
{
    const size_t n = (size_t)c.tm_entries;
    if (!n) return;
    ISLE_REQUIRE(c.tm_doc.p && docs && topics && sums, ISLE_ERR_ARG, "doc_topic_sums: construct_topic_model first");
    ISLE_CUDA_CHECK(cudaMemcpyAsync(docs, c.tm_doc.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(topics, c.tm_topic.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(sums, c.tm_val.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

}  // namespace isle
