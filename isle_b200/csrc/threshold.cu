// threshold.cu -- kernel family (1): per-word threshold selection and construction of the
// thresholded matrix B.  HBM-bound integer/compare work; every result is bit-exact with the
// reference because, after rounding, all quantities are small integers:
//   * reference list_word_freqs_by_sorting + compute_thresholds
//     (src/sparseMatrix.cpp:289-333, 357-485) sort all nnz by (word, value desc) and pick
//     zeta_w by a rank rule; here the sort is replaced by a bin-major per-word histogram of
//     the rounded values (atomics commute on integers) and a one-thread-per-word walk;
//   * reference threshold_and_copy (src/sparseMatrix.cpp:1285-1361) is a serial stream
//     compaction; here it is count -> scan -> ordered warp-ballot scatter.
// B is stored as a 0/1 pattern: every nonzero of row w has the value sqrtf(zeta_w)
// (src/sparseMatrix.cpp:1349), so only sqrt_zeta[V] is kept.
#include <cub/cub.cuh>

#include <cmath>

#include <cstring>

#include "common.cuh"
#include "seg_select.cuh"

namespace isle {

// ------------------------------------------------------------------------------- upload
__global__ void narrow_u64_kernel(const unsigned long long *__restrict__ in,
                                  uint32_t *__restrict__ out, int64_t n, unsigned long long limit,
                                  int *__restrict__ err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long v = in[i];
        if (v >= limit) atomicOr(err, 1);
        out[i] = (uint32_t)v;
    }
}

__global__ void check_rows_kernel(const uint32_t *__restrict__ rows, int64_t n, uint32_t limit,
                                  int *__restrict__ err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n; i += stride) bad |= (rows[i] >= limit);
    if (bad) atomicOr(err, 1);
}

void upload_A(Ctx &c, uint64_t V, uint64_t D, int64_t nnz, const float *vals, const void *rows,
              bool rows64, const int64_t *offsets, float avg, uint64_t nz_docs)
{
    download_B_end(c);
    ISLE_REQUIRE(V > 0 && V < (1ull << 32) && D < (1ull << 32) && nnz >= 0, ISLE_ERR_ARG,
                 "upload_A: V, D must fit 32 bits");
    ISLE_REQUIRE(avg >= 1.0f && avg < 1.0e6f, ISLE_ERR_RANGE, "upload_A: avg_doc_sz out of range");
    c.V = V; c.D = D; c.nnzA = nnz; c.avg_doc_sz = avg;
    c.have_zeta = c.have_B = c.have_csr = c.have_U = c.have_P = false;
    c.a_val.alloc((size_t)nnz);
    c.a_row.alloc((size_t)nnz);
    c.a_off.alloc((size_t)D + 1);
    DevBuf<int> err(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(c.a_val.p, vals, (size_t)nnz * 4, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(c.a_off.p, offsets, ((size_t)D + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    if (rows64) {
        DevBuf<unsigned long long> tmp((size_t)nnz);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(tmp.p, rows, (size_t)nnz * 8, cudaMemcpyHostToDevice, c.stream));
        if (nnz) {
            narrow_u64_kernel<<<grid_for((size_t)nnz, 256), 256, 0, c.stream>>>(tmp.p, c.a_row.p, nnz, V, err.p);
            count_launch(c);
        }
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    } else {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(c.a_row.p, rows, (size_t)nnz * 4, cudaMemcpyHostToDevice, c.stream));
        if (nnz) {
            check_rows_kernel<<<grid_for((size_t)nnz, 256), 256, 0, c.stream>>>(c.a_row.p, nnz, (uint32_t)V, err.p);
            count_launch(c);
        }
    }
    int herr = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    ISLE_REQUIRE(herr == 0, ISLE_ERR_RANGE, "upload_A: row index >= vocab size");
    // global number of non-empty docs (reference _nz_docs, src/sparseMatrix.cpp:97)
    unsigned long long nz = nz_docs;
    if (c.world > 1) {
        DevBuf<unsigned long long> d(1);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(d.p, &nz, 8, cudaMemcpyHostToDevice, c.stream));
        allreduce_sum_u64(c, d.p, 1);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&nz, d.p, 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    c.nz_docs = nz;
}

// --------------------------------------------------------------------------- thresholds
// hist is bin-major: hist[v * V + w] = #docs in which word w has rounded value v (v >= 1).
// roundf() is round-half-away-from-zero == std::round (src/sparseMatrix.cpp:381).
__global__ void __launch_bounds__(256)
hist_kernel(const float *__restrict__ val, const uint32_t *__restrict__ row, int64_t nnz,
            uint32_t V, int bins, uint32_t *__restrict__ hist, int *__restrict__ err)
{
    const int64_t nvec = nnz >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float4 *val4 = reinterpret_cast<const float4 *>(val);
    const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
    for (; i < nvec; i += stride) {
        const float4 v = __ldg(val4 + i);
        const uint4 r = __ldg(row4 + i);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = (int)roundf(vv[j]);
            if (q >= bins || q < 0) { atomicOr(err, 1); continue; }
            if (q >= 1) atomicAdd(hist + (size_t)q * V + rr[j], 1u);
        }
    }
    // tail
    i = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) {
        const int q = (int)roundf(val[i]);
        if (q >= bins || q < 0) atomicOr(err, 1);
        else if (q >= 1) atomicAdd(hist + (size_t)q * V + row[i], 1u);
    }
}

// One thread per word; reads are coalesced across words because hist is bin-major.
// Rank rule of compute_thresholds (src/sparseMatrix.cpp:389-481), see SURVEY Appendix B.1.
__global__ void __launch_bounds__(256)
zeta_kernel(const uint32_t *__restrict__ hist, uint32_t V, int bins, uint32_t count_gr,
            uint32_t count_eq, float *__restrict__ zeta, float *__restrict__ sqrt_zeta,
            unsigned long long *__restrict__ kept_total)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long kept = 0;
    if (w < V) {
        uint32_t cum = 0;   // #values >= current bin
        int z = 0;          // 0 until the count_gr-th largest value has been reached
        bool decided = false;
        float zt = 1.0f;
        for (int v = bins - 1; v >= 1; --v) {
            const uint32_t h = hist[(size_t)v * V + w];
            if (h == 0) continue;
            cum += h;
            if (z == 0 && cum < count_gr) continue;   // rank not reached yet
            // candidate: the count_gr-th largest value, or the next smaller distinct value
            // after a rejected candidate (src/sparseMatrix.cpp:445,470)
            z = v;
            if (h < count_eq || z == 1) {             // accept z, or fall to 1 = keep all (:453-468)
                zt = (h < count_eq) ? (float)z : 1.0f;
                decided = true;
                break;
            }
        }
        // not decided: fewer than count_gr values (:395-412), word absent (:476-480), or the
        // walk ran past the smallest value present (:459) -> zeta = 1, keep everything.
        (void)decided;
        kept = cum;
        zeta[w] = zt;
        sqrt_zeta[w] = sqrtf(zt);
    }
    // block reduction of kept counts
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const unsigned long long s = BR(tmp).Sum(kept);
    if (threadIdx.x == 0 && s) atomicAdd(kept_total, s);
}

void compute_thresholds(Ctx &c, uint64_t k, float *zetas_out, int64_t *new_nnz_out)
{
    download_B_end(c);
    ISLE_REQUIRE(c.a_off.p != nullptr, ISLE_ERR_ARG, "thresholds: upload_A first");
    ISLE_REQUIRE(k >= 1, ISLE_ERR_ARG, "thresholds: k must be >= 1");
    // src/sparseMatrix.cpp:370-373, evaluated on the host in double exactly as written there
    // (w0_c = 1.0, eps1_c = 1.0/60.0, include/hyperparams.h:8-9).
    const double w0 = 1.0, eps1 = 1.0 / 60.0;
    unsigned long long count_gr = (unsigned long long)(w0 * (float)c.nz_docs / (2.0 * (float)k));
    unsigned long long count_eq = (unsigned long long)std::ceil(3.0 * eps1 * w0 * (float)c.nz_docs / (float)k);
    if (count_gr == 0) count_gr = 1;
    if (count_eq == 0) count_eq = 1;

    const int bins = (int)c.avg_doc_sz + 2;   // values <= avg_doc_sz (src/sparseMatrix.cpp:380)
    const uint32_t V = (uint32_t)c.V;
    DevBuf<uint32_t> hist((size_t)bins * V);
    DevBuf<int> err(1);
    DevBuf<unsigned long long> kept(1);
    c.zeta.alloc(V);
    c.sqrt_zeta.alloc(V);
    ISLE_CUDA_CHECK(cudaMemsetAsync(hist.p, 0, hist.bytes(), c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(kept.p, 0, 8, c.stream));
    {
        // algorithmic bytes: read val+idx once, histogram read-modify-write once
        StatScope s(c, "thr_hist", (double)c.nnzA * 8.0 + (double)bins * V * 4.0);
        hist_kernel<<<grid_for((size_t)(c.nnzA / 4 + 1), 256, c.num_sms * 8), 256, 0, c.stream>>>(
            c.a_val.p, c.a_row.p, c.nnzA, V, bins, hist.p, err.p);
        count_launch(c);
    }
    if (c.world > 1) allreduce_sum_u32(c, hist.p, hist.n);
    {
        StatScope s(c, "thr_zeta", (double)bins * V * 4.0 + (double)V * 8.0);
        zeta_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(hist.p, V, bins, (uint32_t)count_gr,
                                                             (uint32_t)count_eq, c.zeta.p,
                                                             c.sqrt_zeta.p, kept.p);
        count_launch(c);
    }
    int herr = 0;
    unsigned long long hk = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&hk, kept.p, 8, cudaMemcpyDeviceToHost, c.stream));
    if (zetas_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(zetas_out, c.zeta.p, (size_t)V * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    ISLE_REQUIRE(herr == 0, ISLE_ERR_RANGE,
                 "thresholds: a normalised value exceeds avg_doc_sz (reference assert, sparseMatrix.cpp:380)");
    c.new_nnz = (int64_t)hk;
    c.have_zeta = true;
    c.have_B = c.have_csr = false;
    c.lifted.release();
    c.lifted_cols = 0;
    if (new_nnz_out) *new_nnz_out = (int64_t)hk;
}

// ------------------------------------------------------------------------------ build B
// One warp per document.  keep(entry) = roundf(val) >= zeta[row]  (src/sparseMatrix.cpp:1344-1348)
__global__ void __launch_bounds__(256)
count_kept_kernel(const float *__restrict__ val, const uint32_t *__restrict__ row,
                  const int64_t *__restrict__ off, uint32_t D, const float *__restrict__ zeta,
                  const uint8_t *__restrict__ select, uint32_t *__restrict__ cnt,
                  float *__restrict__ weight)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < D; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        uint32_t n = 0;
        float wsum = 0.f;
        if (select == nullptr || select[d]) {
            for (int64_t p = b + lane; p < e; p += 32) {
                const float z = __ldg(zeta + row[p]);
                if (roundf(val[p]) >= z) { ++n; wsum += z; }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o);
            wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        }
        if (lane == 0) {
            cnt[d] = n;
            if (weight) weight[d] = wsum;   // zeta are small integers: the fp32 sum is exact
        }
    }
}

struct NonZeroFlag {
    __host__ __device__ uint32_t operator()(uint32_t x) const { return x ? 1u : 0u; }
};
struct ToI64 {
    __host__ __device__ int64_t operator()(uint32_t x) const { return (int64_t)x; }
};

__global__ void __launch_bounds__(256)
compact_kernel(const float *__restrict__ val, const uint32_t *__restrict__ row,
               const int64_t *__restrict__ off, uint32_t D, const float *__restrict__ zeta,
               const uint32_t *__restrict__ cnt, const int64_t *__restrict__ pos_scan,
               const uint32_t *__restrict__ id_scan, uint32_t *__restrict__ b_row,
               int64_t *__restrict__ b_off, uint32_t *__restrict__ b_orig)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < D; d += nw) {
        if (cnt[d] == 0) continue;               // dropped doc (src/sparseMatrix.cpp:1355)
        const uint32_t nid = id_scan[d];
        int64_t out = pos_scan[d];
        if (lane == 0) { b_off[nid] = out; b_orig[nid] = d; }
        const int64_t b = off[d], e = off[d + 1];
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const int64_t p = p0 + lane;
            bool keep = false;
            uint32_t w = 0;
            if (p < e) {
                w = row[p];
                keep = roundf(val[p]) >= __ldg(zeta + w);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) b_row[out + __popc(m & ((1u << lane) - 1u))] = w;
            out += __popc(m);
        }
    }
}

void build_B(Ctx &c, const uint8_t *select, int64_t *nnzB_out, uint64_t *DB_out)
{
    download_B_end(c);
    ISLE_REQUIRE(c.have_zeta, ISLE_ERR_ARG, "build_B: compute thresholds first");
    const uint32_t D = (uint32_t)c.D;
    DevBuf<uint8_t> dsel;
    if (select) {
        dsel.alloc(D);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(dsel.p, select, D, cudaMemcpyHostToDevice, c.stream));
    }
    DevBuf<uint32_t> cnt((size_t)D + 1), id_scan((size_t)D + 1);
    DevBuf<int64_t> pos_scan((size_t)D + 1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(cnt.p + D, 0, 4, c.stream));
    const unsigned wgrid = grid_for((size_t)D * 32, 256, c.num_sms * 16);
    {
        StatScope s(c, "b_count", (double)c.nnzA * 8.0 + (double)D * 12.0);
        if (D) {
            count_kept_kernel<<<wgrid, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, D, c.zeta.p,
                                                           select ? dsel.p : nullptr, cnt.p, nullptr);
            count_launch(c);
        }
    }
    // exclusive scans over D+1 entries: positions (i64) and new doc ids (u32)
    {
        StatScope s(c, "b_scan", (double)D * 24.0);
        size_t t1 = 0, t2 = 0;
        auto it_pos = cub::TransformInputIterator<int64_t, ToI64, uint32_t *>(cnt.p, ToI64());
        auto it_id = cub::TransformInputIterator<uint32_t, NonZeroFlag, uint32_t *>(cnt.p, NonZeroFlag());
        cub::DeviceScan::ExclusiveSum(nullptr, t1, it_pos, pos_scan.p, (int)(D + 1), c.stream);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, it_id, id_scan.p, (int)(D + 1), c.stream);
        DevBuf<uint8_t> tmp(std::max(t1, t2));
        ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, t1, it_pos, pos_scan.p, (int)(D + 1), c.stream));
        ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, t2, it_id, id_scan.p, (int)(D + 1), c.stream));
        count_launch(c, 2);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    int64_t nnzB = 0;
    uint32_t DB = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&nnzB, pos_scan.p + D, 8, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&DB, id_scan.p + D, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.nnzB = nnzB;
    c.DB = DB;
    c.b_row.alloc((size_t)nnzB);
    c.b_off.alloc((size_t)DB + 1);
    c.b_orig.alloc((size_t)DB);
    {
        StatScope s(c, "b_compact", (double)c.nnzA * 8.0 + (double)nnzB * 4.0 + (double)D * 24.0);
        if (D) {
            compact_kernel<<<wgrid, 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, D, c.zeta.p, cnt.p,
                                                        pos_scan.p, id_scan.p, c.b_row.p, c.b_off.p,
                                                        c.b_orig.p);
            count_launch(c);
        }
        ISLE_CUDA_CHECK(cudaMemcpyAsync(c.b_off.p + DB, &c.nnzB, 8, cudaMemcpyHostToDevice, c.stream));
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    {   // global numbering of B's columns: rank r owns [db_offset, db_offset + D_B)  
        c.db_all.assign((size_t)c.world, 0);
        DevBuf<unsigned long long> mine(1), all((size_t)c.world);
        unsigned long long v = DB;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(mine.p, &v, 8, cudaMemcpyHostToDevice, c.stream));
        allgather_u64(c, mine.p, all.p);
        std::vector<unsigned long long> h((size_t)c.world);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(h.data(), all.p, h.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        c.db_offset = c.db_total = 0;
        for (int r = 0; r < c.world; ++r) {
            c.db_all[r] = h[r];
            if (r < c.rank) c.db_offset += h[r];
            c.db_total += h[r];
        }
    }
    c.have_B = true;
    c.have_csr = c.have_U = c.have_P = false;
    if (nnzB_out) *nnzB_out = nnzB;
    if (DB_out) *DB_out = DB;
}

void sampling_weights(Ctx &c, float *out)
{
    ISLE_REQUIRE(c.have_zeta, ISLE_ERR_ARG, "sampling_weights: compute thresholds first");
    const uint32_t D = (uint32_t)c.D;
    DevBuf<uint32_t> cnt(D);
    DevBuf<float> w(D);
    if (D) {
        count_kept_kernel<<<grid_for((size_t)D * 32, 256, c.num_sms * 16), 256, 0, c.stream>>>(
            c.a_val.p, c.a_row.p, c.a_off.p, D, c.zeta.p, nullptr, cnt.p, w.p);
        count_launch(c);
    }
    ISLE_CUDA_CHECK(cudaMemcpyAsync(out, w.p, (size_t)D * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// ----------------------------------------------------------------- importance sampling of documents
// sampled_threshold_and_copy (src/sparseMatrix.cpp:1383-1415), SURVEY 8(f) row 4: weight_d = sum of zeta over the kept
// entries of document d; key_d = u_d^(1 / weight_d) (0 when the weight is 0); keep the documents whose key is at least the
// (floor(rate D) + 1)-th largest (A-Res weighted reservoir sampling).  The reference draws u from libc rand() inside a
// parallel loop (racy); here u_d is a counter-based uniform of (seed, d), so the selection is reproducible
// and independent of the launch geometry.
__device__ __forceinline__ uint64_t sample_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
sample_keys_kernel(const float *__restrict__ w, uint32_t D, uint64_t seed, uint64_t d_offset, float *__restrict__ key)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    // d_offset + d = the document's number in the whole corpus (d_offset = 0 on a single GPU)
    const float u = (float)(sample_mix64(seed ^ sample_mix64(d_offset + (uint64_t)d)) >> 40) * (1.0f / 16777216.0f);     // [0, 1), 24 bits
    const float wd = w[d];
    key[d] = wd == 0.0f ? 0.0f : powf(u, 1.0f / wd);
}

__global__ void __launch_bounds__(256)
sample_select_kernel(const float *__restrict__ key, uint32_t D, const float *__restrict__ sorted_desc, uint32_t nth,
                     uint8_t *__restrict__ select, uint32_t *__restrict__ n_selected)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    const bool keep = d < D && (nth >= D || key[d] >= sorted_desc[nth]);      // :1410-1415 (pivot = dice[nth])
    if (d < D) select[d] = keep ? 1 : 0;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_selected, (uint32_t)__popc(m));
}

__global__ void sample_pivot_kernel(const uint32_t *__restrict__ prefix, float *__restrict__ pivot)
{
    *pivot = segsel::unordered(prefix[0]);
}

// Document-sharded contexts: every rank holds a contiguous range of the documents.  The uniforms are keyed by the GLOBAL
// document number (the ranks' document counts are all-gathered), so the keys are those of a single-GPU run on the whole
// corpus; the pivot -- a corpus-wide order statistic -- comes from the exact distributed radix select of seg_select.cuh
// (four rounds of one 256-bin histogram, all-reduced; no keys travel), so the selection is bit-identical to the
// single-GPU one.  select_out / *n_selected_out describe the rank's own documents.
void sample_docs(Ctx &c, float sample_rate, uint64_t seed, uint8_t *select_out, uint64_t *n_selected_out)
{
    ISLE_REQUIRE(c.have_zeta, ISLE_ERR_ARG, "sample_docs: compute thresholds first");
    ISLE_REQUIRE(sample_rate >= 0.0f && select_out, ISLE_ERR_ARG, "sample_docs: bad arguments");
    const uint32_t D = (uint32_t)c.D;
    if (n_selected_out) *n_selected_out = 0;
    uint64_t d_offset = 0, D_total = D;
    if (c.world > 1) {
        DevBuf<unsigned long long> mine(1), all((size_t)c.world);
        const unsigned long long dl = D;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(mine.p, &dl, 8, cudaMemcpyHostToDevice, c.stream));
        allgather_u64(c, mine.p, all.p);
        std::vector<unsigned long long> h((size_t)c.world);
        read_small(c, h.data(), all.p, h.size() * 8);
        D_total = 0;
        for (int r = 0; r < c.world; ++r) { if (r < c.rank) d_offset += h[r]; D_total += h[r]; }
    }
    if (!D && c.world == 1) return;
    const uint32_t Dn = std::max<uint32_t>(D, 1);
    DevBuf<uint32_t> cnt(Dn), nsel(1);
    DevBuf<float> w(Dn), key(Dn), sorted(Dn);
    DevBuf<uint8_t> sel(Dn);
    if (D) {
        count_kept_kernel<<<grid_for((size_t)D * 32, 256, c.num_sms * 16), 256, 0, c.stream>>>(c.a_val.p, c.a_row.p, c.a_off.p, D, c.zeta.p,
                                                                                          nullptr, cnt.p, w.p);
        sample_keys_kernel<<<(D + 255) / 256, 256, 0, c.stream>>>(w.p, D, seed, d_offset, key.p);
        count_launch(c, 2);
    }
    // nth = (size_t)(sample_rate * (float)D)   (:1406), D = the whole corpus
    const float nth_f = sample_rate * (float)D_total;
    const uint64_t nth64 = nth_f >= (float)D_total ? D_total : (uint64_t)nth_f;
    ISLE_CUDA_CHECK(cudaMemsetAsync(nsel.p, 0, 4, c.stream));
    if (c.world == 1) {
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeysDescending(nullptr, tb, key.p, sorted.p, (int)D, 0, 32, c.stream);
        DevBuf<uint8_t> tmp(tb);
        ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortKeysDescending(tmp.p, tb, key.p, sorted.p, (int)D, 0, 32, c.stream));
        const uint32_t nth = (uint32_t)nth64;
        sample_select_kernel<<<(D + 255) / 256, 256, 0, c.stream>>>(key.p, D, sorted.p, nth, sel.p, nsel.p);
        count_launch(c, 2);
    } else {
        ISLE_REQUIRE(D_total < 0xFFFFFFFFull, ISLE_ERR_RANGE, "sample_docs: more than 2^32 - 1 documents");
        DevBuf<float> pivot(1);
        if (nth64 < D_total) {
            DevBuf<uint32_t> seg(Dn), kth(1), prefix(1), hist(256);
            const uint32_t k0 = (uint32_t)nth64;
            ISLE_CUDA_CHECK(cudaMemsetAsync(seg.p, 0, seg.bytes(), c.stream));
            ISLE_CUDA_CHECK(cudaMemsetAsync(prefix.p, 0, 4, c.stream));
            ISLE_CUDA_CHECK(cudaMemcpyAsync(kth.p, &k0, 4, cudaMemcpyHostToDevice, c.stream));
            for (int round = 0; round < 4; ++round) {
                ISLE_CUDA_CHECK(cudaMemsetAsync(hist.p, 0, hist.bytes(), c.stream));
                if (D) segsel::hist_pairs_kernel<<<grid_for((size_t)D, 256), 256, 0, c.stream>>>(seg.p, key.p, (int64_t)D, prefix.p, round, hist.p);
                allreduce_sum_u32(c, hist.p, 256);
                segsel::pick_kernel<<<1, 128, 0, c.stream>>>(hist.p, 1, kth.p, prefix.p);
                count_launch(c, 2);
            }
            sample_pivot_kernel<<<1, 1, 0, c.stream>>>(prefix.p, pivot.p);
            count_launch(c);
        }
        // with nth >= D_total every document is kept (first operand of the test in the kernel: nth >= D)
        if (D) sample_select_kernel<<<(D + 255) / 256, 256, 0, c.stream>>>(key.p, D, pivot.p, nth64 < D_total ? 0u : D, sel.p, nsel.p);
        count_launch(c);
    }
    uint32_t h = 0;
    if (D) ISLE_CUDA_CHECK(cudaMemcpyAsync(select_out, sel.p, D, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&h, nsel.p, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (n_selected_out) *n_selected_out = h;
}

// ---------------------------------------------------------------------------- download B
__global__ void expand_B_kernel(const uint32_t *__restrict__ b_row, const float *__restrict__ sqrt_zeta,
                                int64_t n, float *__restrict__ vals, unsigned long long *__restrict__ rows)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t w = b_row[i];
        if (vals) vals[i] = __ldg(sqrt_zeta + w);
        if (rows) rows[i] = w;
    }
}

__global__ void widen_u32_kernel(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out,
                                 int64_t n, unsigned long long add)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (unsigned long long)in[i] + add;
}

void download_B(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "download_B: build_B first");
    const int64_t n = c.nnzB;
    if ((vals || rows) && n) {
        DevBuf<float> dv(vals ? (size_t)n : 0);
        DevBuf<unsigned long long> dr(rows ? (size_t)n : 0);
        expand_B_kernel<<<grid_for((size_t)n, 256), 256, 0, c.stream>>>(c.b_row.p, c.sqrt_zeta.p, n, dv.p, dr.p);
        count_launch(c);
        if (vals) ISLE_CUDA_CHECK(cudaMemcpyAsync(vals, dv.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
        if (rows) ISLE_CUDA_CHECK(cudaMemcpyAsync(rows, dr.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    if (offsets)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(offsets, c.b_off.p, ((size_t)c.DB + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    if (orig && c.DB) {
        DevBuf<unsigned long long> d((size_t)c.DB);
        widen_u32_kernel<<<grid_for((size_t)c.DB, 256), 256, 0, c.stream>>>(c.b_orig.p, d.p, (int64_t)c.DB, 0ull);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(orig, d.p, (size_t)c.DB * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// The same copy in the background.  threshold_and_copy (src/sparseMatrix.cpp:1285-1361) leaves B in host arrays that
// nothing on the spectral core reads again (the device keeps its own B), so the 12 bytes per nonzero need not hold the
// eigensolver up: the expanded arrays are staged on the device, a host thread moves them over a separate copy stream
// (DMA from the copy engine; staged through the driver's bounce buffers when the caller's arrays are pageable, which is why
// it is a thread and not just an async call), and download_B_end() joins it.  The caller's arrays must stay valid and
// unread until then.  One download at a time; anything that rebuilds B ends a pending one first.
// Small device -> host read that does NOT go through the copy engine: a kernel stores the bytes into host-mapped memory and
// the host waits for the stream.  The copy engine serves queued copies in order, so while a background download of B
// (below) is in flight an ordinary 4-byte cudaMemcpyAsync on the main stream would wait for all 0.8 GB of it; these reads
// are the eigensolver's control flow (operator-layout sizes, residuals, cuSOLVER's info) and must not.
__global__ void publish_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, size_t bytes)
{
    for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}
void read_small(Ctx &c, void *dst, const void *dev_src, size_t bytes)
{
    constexpr size_t kCap = 1 << 20;
    if (bytes == 0) return;
    if (!c.small_host && !c.small_failed) {
        if (cudaHostAlloc((void **)&c.small_host, kCap, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&c.small_dev, c.small_host, 0) != cudaSuccess) {
            cudaGetLastError();
            if (c.small_host) cudaFreeHost(c.small_host);
            c.small_host = nullptr;
            c.small_failed = true;
        }
    }
    if (!c.small_host || bytes > kCap) {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return;
    }
    publish_kernel<<<1, bytes >= 1024 ? 1024 : 64, 0, c.stream>>>((const unsigned char *)dev_src, c.small_dev, bytes);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    std::memcpy(dst, c.small_host, bytes);
}

void download_B_begin(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "download_B_begin: build_B first");
    download_B_end(c);
    const int64_t n = c.nnzB;
    if (!c.copy_stream) {
        ISLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        ISLE_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
    }
    c.dl_vals.alloc((vals && n) ? (size_t)n : 0);
    c.dl_rows.alloc((rows && n) ? (size_t)n : 0);
    c.dl_orig.alloc((orig && c.DB) ? (size_t)c.DB : 0);
    if ((vals || rows) && n) {
        expand_B_kernel<<<grid_for((size_t)n, 256), 256, 0, c.stream>>>(c.b_row.p, c.sqrt_zeta.p, n, c.dl_vals.p, c.dl_rows.p);
        count_launch(c);
    }
    if (orig && c.DB) {
        widen_u32_kernel<<<grid_for((size_t)c.DB, 256), 256, 0, c.stream>>>(c.b_orig.p, c.dl_orig.p, (int64_t)c.DB, 0ull);
        count_launch(c);
    }
    ISLE_CUDA_CHECK(cudaEventRecord(c.ev_copy, c.stream));
    c.dl_error.clear();
    c.dl_active = true;
    Ctx *cp = &c;
    const size_t DB = (size_t)c.DB;
    c.dl_thread = std::thread([cp, vals, rows, offsets, orig, n, DB] {
        Ctx &c = *cp;
        auto chk = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && c.dl_error.empty()) c.dl_error = std::string(what) + ": " + cudaGetErrorString(e);
        };
        chk(cudaSetDevice(c.device), "cudaSetDevice");
        chk(cudaStreamWaitEvent(c.copy_stream, c.ev_copy, 0), "cudaStreamWaitEvent");
        // 32 MB pieces: other copies (the final U of the eigensolver) get their turn between two pieces; the eigensolver's
        // small control-flow reads do not use the copy engine at all (read_small)
        auto copy = [&](void *dst, const void *src, size_t bytes, const char *what) {
            constexpr size_t kPiece = (size_t)32 << 20;
            for (size_t off = 0; off < bytes; off += kPiece)
                chk(cudaMemcpyAsync((char *)dst + off, (const char *)src + off, std::min(kPiece, bytes - off), cudaMemcpyDeviceToHost,
                                    c.copy_stream), what);
        };
        if (vals && n) copy(vals, c.dl_vals.p, (size_t)n * 4, "copy vals");
        if (rows && n) copy(rows, c.dl_rows.p, (size_t)n * 8, "copy rows");
        if (offsets) copy(offsets, c.b_off.p, (DB + 1) * 8, "copy offsets");
        if (orig && DB) copy(orig, c.dl_orig.p, DB * 8, "copy original_cols");
        chk(cudaStreamSynchronize(c.copy_stream), "cudaStreamSynchronize(copy)");
    });
}

void download_B_end(Ctx &c)
{
    if (!c.dl_active) return;
    if (c.dl_thread.joinable()) c.dl_thread.join();
    c.dl_active = false;
    c.dl_vals.release();
    c.dl_rows.release();
    c.dl_orig.release();
    ISLE_REQUIRE(c.dl_error.empty(), ISLE_ERR_CUDA, "download_B (background): " + c.dl_error);
}

// frobenius = sum over nonzeros of sqrt_zeta[row]^2 = sum zeta[row]; the reference sums the
// fp32 squares with cblas_sdot (src/sparseMatrix.cpp:1099) and only logs the value.
__global__ void __launch_bounds__(256)
frob_kernel(const uint32_t *__restrict__ b_row, const float *__restrict__ sqrt_zeta, int64_t n,
            double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double s = 0.0;
    for (; i < n; i += stride) {
        const float v = __ldg(sqrt_zeta + b_row[i]);
        s += (double)(v * v);
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double t = BR(tmp).Sum(s);
    if (threadIdx.x == 0) atomicAdd(out, t);
}

float frobenius(Ctx &c)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "frobenius: build_B first");
    DevBuf<double> d(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(d.p, 0, 8, c.stream));
    if (c.nnzB) {
        frob_kernel<<<grid_for((size_t)c.nnzB, 256, c.num_sms * 8), 256, 0, c.stream>>>(c.b_row.p, c.sqrt_zeta.p, c.nnzB, d.p);
        count_launch(c);
    }
    if (c.world > 1) allreduce_sum_f64(c, d.p, 1);
    double h = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&h, d.p, 8, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    return (float)h;
}

}  // namespace isle
