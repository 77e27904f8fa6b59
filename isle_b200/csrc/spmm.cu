// spmm.cu -- kernel family (2): the B * B^T * X operator that drives block Krylov-Schur.
//
// Replaces MKL_SpSpTrProd (reference include/matUtils.h:14-419): the ctor's CSC->CSR copy by
// sorting (:109-135) is a stable device radix sort of (word, doc) pairs; multiply()
// (:336-365, two mkl_scsrmm passes) is
//     pass 1   Y[d,:] = sum_{w in doc d}  Xs[w,:]          Xs = diag(sqrt_zeta) X
//     pass 2   T[w,:] = sum_{d containing w} Y[d,:]        Z  = diag(sqrt_zeta) T
// B is a 0/1 pattern scaled by sqrt_zeta per row, so no value stream is read.
//
// Layout.  Words are relabelled by decreasing row length ("rank space").  The H most frequent
// words (density >= ~1 % of the documents; more than half of all nonzeros on Zipfian corpora) form
// the dense HEAD, kept as bitmaps and multiplied on the tensor cores (spmm_head.cu).  The TAIL
// keeps index lists (4 bytes per nonzero, doc-major for pass 1, rank-major for pass 2) and is
// processed by the gather kernel below: dense operands are row-major with a fixed 16-float
// (64 B = two 32 B sectors) row stride so a 4-lane group fetches one row with one float4 per lane.
// Work is cut into items (row, first nnz, length <= kChunk) sorted by decreasing length, so the
// eight 4-lane groups of a warp run equally long loops despite Zipfian row lengths;
// rows cut into several items finish with float4 atomics.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"
#include "p2p.cuh"

namespace isle {

static constexpr int kStride = 16;          // floats per padded dense row
static constexpr uint32_t kChunk = 256;     // max nonzeros per work item (bounds the serial length of a 4-lane group)
static constexpr uint32_t kSplitFlag = 0x80000000u;

// ------------------------------------------------------------------------------ CSR copy
__global__ void __launch_bounds__(256)
fill_doc_ids_kernel(const int64_t *__restrict__ off, uint32_t DB, uint32_t *__restrict__ doc_of)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        for (int64_t p = b + lane; p < e; p += 32) doc_of[p] = d;
    }
}

// csr_off[w] = first position of word w in the word-sorted key array (lower bound).
__global__ void row_offsets_kernel(const uint32_t *__restrict__ keys, int64_t n, uint32_t V,
                                   int64_t *__restrict__ off)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > V) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < w) lo = mid + 1; else hi = mid;
    }
    off[w] = lo;
}

// ----------------------------------------------------------------------------- work items
__global__ void count_items_kernel(const int64_t *__restrict__ off, uint32_t nrows, uint32_t chunk,
                                   uint32_t *__restrict__ nitems)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t len = off[r + 1] - off[r];
    nitems[r] = (uint32_t)((len + chunk - 1) / chunk);   // empty rows produce no item
}

__global__ void fill_items_kernel(const int64_t *__restrict__ off, uint32_t nrows, uint32_t chunk,
                                  const uint32_t *__restrict__ item_scan, WorkItem *__restrict__ items,
                                  uint32_t *__restrict__ sort_key)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t b = off[r], e = off[r + 1];
    const uint32_t n = (uint32_t)((e - b + chunk - 1) / chunk);
    uint32_t o = item_scan[r];
    for (uint32_t i = 0; i < n; ++i, ++o) {
        const int64_t ib = b + (int64_t)i * chunk;
        const uint32_t len = (uint32_t)min((int64_t)chunk, e - ib);
        WorkItem it;
        it.out_row = r;
        it.len = len | (n > 1 ? kSplitFlag : 0u);
        it.begin = ib;
        items[o] = it;
        sort_key[o] = chunk - len;      // ascending key == descending length
    }
}

__global__ void iota_kernel(uint32_t *__restrict__ p, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__global__ void gather_items_kernel(const WorkItem *__restrict__ in, const uint32_t *__restrict__ perm,
                                    size_t n, WorkItem *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

static void build_items(Ctx &c, const int64_t *off, uint32_t nrows, DevBuf<WorkItem> &items, size_t &n_items)
{
    DevBuf<uint32_t> cnt((size_t)nrows + 1), scan((size_t)nrows + 1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(cnt.p + nrows, 0, 4, c.stream));
    count_items_kernel<<<(nrows + 255) / 256, 256, 0, c.stream>>>(off, nrows, kChunk, cnt.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, scan.p, (int)(nrows + 1), c.stream);
    DevBuf<uint8_t> tmp(tb);
    ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, scan.p, (int)(nrows + 1), c.stream));
    uint32_t total = 0;
    read_small(c, &total, scan.p + nrows, 4);
    n_items = total;
    items.alloc(total);
    count_launch(c, 2);
    if (!total) return;
    DevBuf<WorkItem> raw(total);
    DevBuf<uint32_t> key(total), key2(total), idx(total), idx2(total);
    fill_items_kernel<<<(nrows + 255) / 256, 256, 0, c.stream>>>(off, nrows, kChunk, scan.p, raw.p, key.p);
    // identity permutation, then a stable sort by key: equal lengths keep row order (locality)
    iota_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(idx.p, total);
    {
        size_t ts = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, ts, key.p, key2.p, idx.p, idx2.p, (int)total, 0, 12, c.stream);
        DevBuf<uint8_t> t2(ts);
        ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(t2.p, ts, key.p, key2.p, idx.p, idx2.p, (int)total, 0, 12, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    gather_items_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(raw.p, idx2.p, total, items.p);
    count_launch(c, 4);
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// ------------------------------------------------------------------- rank space, head / tail
__global__ void row_len_kernel(const int64_t *__restrict__ off, uint32_t V, uint32_t *__restrict__ len,
                               uint32_t *__restrict__ id)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V) return;
    len[w] = (uint32_t)(off[w + 1] - off[w]);
    id[w] = w;
}

__global__ void scatter_rank_kernel(const uint32_t *__restrict__ word_of_rank, uint32_t V, uint32_t *__restrict__ rank_of)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < V) rank_of[word_of_rank[r]] = r;
}

// number of leading entries of the descending array `len` that are >= thr
__global__ void count_ge_kernel(const uint32_t *__restrict__ len, uint32_t V, uint32_t thr, uint32_t *__restrict__ out)
{
    uint32_t lo = 0, hi = V;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (len[mid] >= thr) lo = mid + 1; else hi = mid;
    }
    *out = lo;
}

// lengths of the rank-major tail rows: head ranks become empty rows
__global__ void tail_len_kernel(const uint32_t *__restrict__ len_sorted, uint32_t V, uint32_t H, uint32_t *__restrict__ out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= V) out[r] = (r < H || r == V) ? 0u : len_sorted[r];
}

struct U32ToI64 {
    __host__ __device__ int64_t operator()(uint32_t x) const { return (int64_t)x; }
};

static void scan_to_offsets(Ctx &c, uint32_t *cnt, size_t n_plus_1, int64_t *off)
{
    auto it = cub::TransformInputIterator<int64_t, U32ToI64, uint32_t *>(cnt, U32ToI64());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, it, off, (int)n_plus_1, c.stream);
    DevBuf<uint8_t> tmp(tb);
    ISLE_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, off, (int)n_plus_1, c.stream));
    count_launch(c);
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));   // tmp is released on return
}

// tail nonzeros per document (warp per document)
__global__ void __launch_bounds__(256)
doc_tail_count_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ row, const uint32_t *__restrict__ rank_of,
                      uint32_t H, uint32_t DB, uint32_t *__restrict__ cnt)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        uint32_t n = 0;
        for (int64_t p = b + lane; p < e; p += 32) n += (__ldg(rank_of + row[p]) >= H) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) cnt[d] = n;
    }
}

// Writes the doc-major tail list (ranks, document order preserved) and sets the head bits of the
// document in both bitmaps.  bit position of element k inside its 32-bit word: (k>>1) + 16 (k&1),
// the order expand_word() (spmm_head.cu) unpacks.
__device__ __forceinline__ uint32_t head_bit(uint32_t k) { return 1u << (((k & 31u) >> 1) + ((k & 1u) << 4)); }

// int8 head engine (spmm_head_i8.cu): a row's 256-k super-chunk is 8 words; element k sits in word (k >> 2) & 7 at bit
// 8 (k & 3) + (k >> 5), so that bit j of byte b of word w is k = 32 j + 4 w + b and (word & (0x01010101 << j)) is four
// cells of MMA j.  Word index of (row m, element k) in a bitmap of NSC super-chunks per 128-row tile:
__device__ __forceinline__ size_t head8_word(uint32_t m, uint32_t k, uint32_t NSC)
{
    return ((((size_t)(m >> 7) * NSC + (k >> 8)) * kHeadTile + (m & 127u)) << 3) + ((k >> 2) & 7u);
}
__device__ __forceinline__ uint32_t head8_bit(uint32_t k) { return 1u << (((k & 3u) << 3) + ((k & 255u) >> 5)); }

template <bool I8>
__global__ void __launch_bounds__(256)
doc_split_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ row, const uint32_t *__restrict__ rank_of,
                 uint32_t H, uint32_t DB, uint32_t NC1, uint32_t NC2, const int64_t *__restrict__ t1_off,
                 uint32_t *__restrict__ t1_idx, uint32_t *__restrict__ bits1, uint32_t *__restrict__ bits2)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        int64_t out = t1_off[d];
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const int64_t p = p0 + lane;
            bool tail = false;
            uint32_t r = 0;
            if (p < e) {
                r = __ldg(rank_of + row[p]);
                tail = r >= H;
                if (!tail) {
                    // bits1: rows = documents, k = rank;  bits2: rows = ranks, k = document
                    if (I8) {
                        atomicOr(bits1 + head8_word(d, r, NC1), head8_bit(r));
                        atomicOr(bits2 + head8_word(r, d, NC2), head8_bit(d));
                    } else {
                        const size_t i1 = ((((size_t)(d >> 7) * NC1 + (r >> 7)) * kHeadTile + (d & 127u)) << 2) + ((r >> 5) & 3u);
                        const size_t i2 = ((((size_t)(r >> 7) * NC2 + (d >> 7)) * kHeadTile + (r & 127u)) << 2) + ((d >> 5) & 3u);
                        atomicOr(bits1 + i1, head_bit(r));
                        atomicOr(bits2 + i2, head_bit(d));
                    }
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, tail);
            if (tail) t1_idx[out + __popc(m & ((1u << lane) - 1u))] = r;
            out += __popc(m);
        }
    }
}

// rank-major tail list: row r >= H is the word-major row of word_of_rank[r] (warp per row)
__global__ void __launch_bounds__(256)
rank_rows_copy_kernel(const int64_t *__restrict__ csr_off, const uint32_t *__restrict__ csr_col,
                      const uint32_t *__restrict__ word_of_rank, uint32_t H, uint32_t V,
                      const int64_t *__restrict__ t2_off, uint32_t *__restrict__ t2_idx)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t r = H + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; r < V; r += nw) {
        const uint32_t w = word_of_rank[r];
        const int64_t b = csr_off[w], n = csr_off[w + 1] - b;
        const int64_t o = t2_off[r];
        for (int64_t i = lane; i < n; i += 32) t2_idx[o + i] = csr_col[b + i];
    }
}

void build_csr(Ctx &c, bool head_i8)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "build_csr: build_B first");
    head_i8 = head_i8 && c.opt("spmm_head_i8", 1) != 0;
    if (c.have_csr && (c.H == 0 || c.head_i8 == head_i8)) return;
    c.have_csr = false;
    c.head_i8 = head_i8;
    const uint32_t kTileK = head_i8 ? 256u : (uint32_t)kHeadChunk;      // K granularity of the head engine
    StatScope s(c, "csr_build", (double)c.nnzB * 24.0);
    const int64_t n = c.nnzB;
    const uint32_t V = (uint32_t)c.V, DB = (uint32_t)c.DB;
    const unsigned wgrid = grid_for((size_t)std::max<uint32_t>(DB, 1) * 32, 256, c.num_sms * 16);
    c.csr_col.alloc((size_t)n);
    c.csr_off.alloc((size_t)V + 1);
    {
        DevBuf<uint32_t> doc_of((size_t)n), keys_out((size_t)n);
        if (DB) {
            fill_doc_ids_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_off.p, DB, doc_of.p);
            count_launch(c);
        }
        int bits = 1;
        while ((1ull << bits) < c.V) ++bits;
        size_t ts = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, ts, c.b_row.p, keys_out.p, doc_of.p, c.csr_col.p, n, 0, bits, c.stream);
        DevBuf<uint8_t> tmp(ts);
        if (n)
            ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, ts, c.b_row.p, keys_out.p, doc_of.p, c.csr_col.p, n, 0, bits, c.stream));
        row_offsets_kernel<<<(V + 1 + 255) / 256, 256, 0, c.stream>>>(keys_out.p, n, V, c.csr_off.p);
        count_launch(c, 2);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }

    // ---- rank space: words by decreasing row length
    c.rank_of.alloc(V);
    c.word_of_rank.alloc(V);
    DevBuf<uint32_t> len(V), len_sorted(V), ids(V);
    row_len_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(c.csr_off.p, V, len.p, ids.p);
    {
        size_t ts = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, ts, len.p, len_sorted.p, ids.p, c.word_of_rank.p, (int)V, 0, 32, c.stream);
        DevBuf<uint8_t> tmp(ts);
        ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(tmp.p, ts, len.p, len_sorted.p, ids.p, c.word_of_rank.p, (int)V, 0, 32, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    scatter_rank_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(c.word_of_rank.p, V, c.rank_of.p);
    count_launch(c, 3);

    // ---- head size: words present in at least `density` of the documents, whole 128-row tiles
    uint32_t H = 0;
    if (c.opt("spmm_head", 1) && DB >= 1) {
        const double dens = c.opt("spmm_head_density_ppm", 12000) * 1e-6;
        const uint32_t thr = (uint32_t)std::max(1.0, std::ceil(dens * (double)DB));
        DevBuf<uint32_t> dcnt(1);
        count_ge_kernel<<<1, 1, 0, c.stream>>>(len_sorted.p, V, thr, dcnt.p);
        count_launch(c);
        uint32_t h0 = 0;
        read_small(c, &h0, dcnt.p, 4);
        const uint32_t hmax = (uint32_t)std::max(0, c.opt("spmm_head_max", 4096));
        H = std::min(h0, hmax) / kTileK * kTileK;
    }
    c.H = H;
    c.DBpad = (DB + kTileK - 1) / kTileK * kTileK;
    const uint32_t NC1 = H / kTileK, NC2 = c.DBpad / kTileK;

    // ---- doc-major tail list + head bitmaps
    c.t1_off.alloc((size_t)DB + 1);
    {
        DevBuf<uint32_t> cnt((size_t)DB + 1);
        ISLE_CUDA_CHECK(cudaMemsetAsync(cnt.p + DB, 0, 4, c.stream));
        if (DB) doc_tail_count_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_off.p, c.b_row.p, c.rank_of.p, H, DB, cnt.p);
        count_launch(c);
        scan_to_offsets(c, cnt.p, (size_t)DB + 1, c.t1_off.p);
    }
    int64_t nnz_tail = 0;
    read_small(c, &nnz_tail, c.t1_off.p + DB, 8);
    c.nnz_tail = nnz_tail;
    c.t1_idx.alloc((size_t)nnz_tail);
    const size_t nbits = (size_t)c.DBpad * H / kHeadChunk;       // uint4 entries per bitmap
    c.bits1.alloc(nbits);
    c.bits2.alloc(nbits);
    if (nbits) {
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.bits1.p, 0, c.bits1.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.bits2.p, 0, c.bits2.bytes(), c.stream));
    }
    if (DB) {
        if (head_i8)
            doc_split_kernel<true><<<wgrid, 256, 0, c.stream>>>(c.b_off.p, c.b_row.p, c.rank_of.p, H, DB, NC1, NC2, c.t1_off.p, c.t1_idx.p,
                                                                reinterpret_cast<uint32_t *>(c.bits1.p), reinterpret_cast<uint32_t *>(c.bits2.p));
        else
            doc_split_kernel<false><<<wgrid, 256, 0, c.stream>>>(c.b_off.p, c.b_row.p, c.rank_of.p, H, DB, NC1, NC2, c.t1_off.p, c.t1_idx.p,
                                                                 reinterpret_cast<uint32_t *>(c.bits1.p), reinterpret_cast<uint32_t *>(c.bits2.p));
        count_launch(c);
    }

    // ---- rank-major tail list
    c.t2_off.alloc((size_t)V + 1);
    {
        DevBuf<uint32_t> cnt((size_t)V + 1);
        tail_len_kernel<<<(V + 1 + 255) / 256, 256, 0, c.stream>>>(len_sorted.p, V, H, cnt.p);
        count_launch(c);
        scan_to_offsets(c, cnt.p, (size_t)V + 1, c.t2_off.p);
    }
    c.t2_idx.alloc((size_t)nnz_tail);
    if (V > H) {
        rank_rows_copy_kernel<<<grid_for((size_t)(V - H) * 32, 256, c.num_sms * 16), 256, 0, c.stream>>>(
            c.csr_off.p, c.csr_col.p, c.word_of_rank.p, H, V, c.t2_off.p, c.t2_idx.p);
        count_launch(c);
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.csr_col.release();    // only the rank-major copy is used from here on

    build_items(c, c.t1_off.p, DB, c.items_bt, c.n_items_bt);
    build_items(c, c.t2_off.p, V, c.items_b, c.n_items_b);
    c.xs.alloc((size_t)V * kStride);
    c.ybuf.alloc((size_t)std::max<uint32_t>(c.DBpad, 1) * kStride);
    c.zbuf.alloc((size_t)V * kStride);
    c.xbfp.alloc((size_t)V * 2);
    c.colmax.alloc(kStride);
    c.ybfp.alloc((size_t)std::max<uint32_t>(c.DBpad, 1) * 2);
    c.ycolmax.alloc(kStride);
    if (H && head_i8) {
        c.xdig.alloc((size_t)kHead8Rows * H);           // 4 s8 digits x 10 columns in three groups of 16 rows
        c.ydig.alloc((size_t)kHead8Rows * c.DBpad);
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.xdig.p, 0, c.xdig.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.ydig.p, 0, c.ydig.bytes(), c.stream));
        c.xsplit.release();
        c.ysplit.release();
    } else if (H) {
        const size_t nrows = 48;    // 3 bf16 pieces x up to 16 columns
        c.xsplit.alloc(nrows * H);
        c.ysplit.alloc(nrows * c.DBpad);
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.xsplit.p, 0, c.xsplit.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.ysplit.p, 0, c.ysplit.bytes(), c.stream));
        c.xdig.release();
        c.ydig.release();
    }
    c.have_csr = true;
}

// --------------------------------------------------------------------------- the gather pass
// One 4-lane group per work item; lane `sub` owns floats [4 sub, 4 sub + 4) of the row.
// Per step each lane loads one index (coalesced 16 B per group), indices are exchanged by
// shuffle and eight independent float4 gathers are issued before they are summed, in a fixed
// order so that results are run-to-run deterministic for unsplit rows.
// Index stream loads: read once, so optionally kept out of L1 (ld.global.nc.L1::no_allocate) to leave
// the cache to the gathered dense rows.  ADD: the output row already holds the head part.
template <bool NOALLOC>
__device__ __forceinline__ uint32_t ld_idx(const uint32_t *p)
{
    if (NOALLOC) {
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
    }
    return __ldg(p);
}

template <int NCH, bool NOALLOC, bool ADD>
__global__ void __launch_bounds__(256)
spmm_gather_kernel(const WorkItem *__restrict__ items, size_t n_items, const uint32_t *__restrict__ idx,
                   const float4 *__restrict__ in, float4 *__restrict__ out)
{
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int sub = threadIdx.x & 3;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);
    if (g >= n_items) return;     // whole groups exit together
    const WorkItem it = items[g];
    const uint32_t len = it.len & ~kSplitFlag;
    const uint32_t *p = idx + it.begin;
    const bool active = sub < NCH;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t j = 0;
    for (; j + 8 <= len; j += 8) {
        const uint32_t i0 = ld_idx<NOALLOC>(p + j + sub), i1 = ld_idx<NOALLOC>(p + j + 4 + sub);
        float4 v[8];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t r0 = __shfl_sync(gmask, i0, s, 4);
            const uint32_t r1 = __shfl_sync(gmask, i1, s, 4);
            if (active) {
                v[s] = __ldg(in + (size_t)r0 * 4 + sub);
                v[4 + s] = __ldg(in + (size_t)r1 * 4 + sub);
            }
        }
        if (active) {
#pragma unroll
            for (int s = 0; s < 8; ++s) { acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w; }
        }
    }
    for (; j < len; j += 4) {
        const uint32_t i0 = (j + sub < len) ? ld_idx<NOALLOC>(p + j + sub) : 0u;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t r0 = __shfl_sync(gmask, i0, s, 4);
            if (active && j + s < len) {
                const float4 v = __ldg(in + (size_t)r0 * 4 + sub);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    if (!active) return;
    float4 *o = out + (size_t)it.out_row * 4 + sub;
    if (it.len & kSplitFlag) {
        atomicAdd(o, acc);   // sm_90+: 16-byte vector atomic (RED.128)
    } else if (ADD) {
        const float4 h = *o;
        *o = make_float4(h.x + acc.x, h.y + acc.y, h.z + acc.z, h.w + acc.w);
    } else {
        *o = acc;
    }
}

// ------------------------------------------------------------- one-sector operand rows (BFP)
// For block sizes <= 10 the gathered operand rows are stored as 32 bytes = ONE sector instead of
// a padded 64-byte fp32 row, which halves the L2->SM traffic that bounds the tail passes.  A row
// is two 16-byte units of five columns; a unit holds five 24-bit two's-complement mantissas
// (bytes 0..14) and one shared power-of-two scale (byte 15 = the float exponent field of
// scale * 2^-8, 0 for an all-zero unit).  The largest element of a unit keeps a full fp32
// significand (error <= 2^-24 of the unit maximum per element).
__device__ __forceinline__ uint4 bfp_encode5(const float *v)
{
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) m = fmaxf(m, fabsf(v[j]));
    const int em = (int)((__float_as_uint(m) >> 23) & 0xFFu);     // exponent field of the maximum
    if (em < 31 || em == 255) return make_uint4(0u, 0u, 0u, 0u);  // zero / denormal-range (or non-finite) unit
    const int es = em - 22;                                        // scale s = 2^(es-127): |v| / s < 2^23
    const float inv_s = __uint_as_float((uint32_t)(254 - es) << 23);
    int q[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) q[j] = max(-8388607, min(8388607, __float2int_rn(v[j] * inv_s)));
    uint4 w;
    w.x = ((uint32_t)q[0] & 0xFFFFFFu) | ((uint32_t)q[1] << 24);
    w.y = (((uint32_t)q[1] >> 8) & 0xFFFFu) | ((uint32_t)q[2] << 16);
    w.z = (((uint32_t)q[2] >> 16) & 0xFFu) | ((uint32_t)q[3] << 8);
    w.w = ((uint32_t)q[4] & 0xFFFFFFu) | ((uint32_t)(es - 8) << 24);
    return w;
}

// acc[j] += value j of the unit
__device__ __forceinline__ void bfp_accum5(const uint4 w, float *acc)
{
    const float sc = __uint_as_float((w.w >> 1) & 0x7F800000u);
    acc[0] = fmaf((float)(int)(w.x << 8), sc, acc[0]);
    acc[1] = fmaf((float)(int)(__funnelshift_r(w.x, w.y, 16) & 0xFFFFFF00u), sc, acc[1]);
    acc[2] = fmaf((float)(int)(__funnelshift_r(w.y, w.z, 8) & 0xFFFFFF00u), sc, acc[2]);
    acc[3] = fmaf((float)(int)(w.z & 0xFFFFFF00u), sc, acc[3]);
    acc[4] = fmaf((float)(int)(w.w << 8), sc, acc[4]);
}

// One 4-lane group per work item = two lane pairs; a pair gathers the rows of the four indices its
// two lanes loaded, lane (sub & 1) taking unit (sub & 1) of each row with one 16-byte load, so a
// warp instruction fetches 16 whole sectors.  The two pairs' sums are combined by shuffle; lanes 0 and
// 1 of the group then own columns 0..4 and 5..9 of the output row (fp32, 16-float stride).
// one 16-byte unit of a gathered operand row.  MODE 0: ld.global.nc (allocates an L1 line); 1: L1::no_allocate;
// 2: L1::evict_first
template <int MODE>
__device__ __forceinline__ uint4 ld_row(const uint4 *p)
{
    if (MODE == 0) return __ldg(p);
    uint4 v;
    if (MODE == 1)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::evict_first.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <bool ADD, int PIPE>
__global__ void __launch_bounds__(256)
spmm_gather_bfp_kernel(const WorkItem *__restrict__ items, size_t n_items, const uint32_t *__restrict__ idx,
                       const uint4 *__restrict__ in, float *__restrict__ out, int b)
{
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int sub = threadIdx.x & 3, u = sub & 1;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);   // groups of a warp run loops of different lengths
    if (g >= n_items) return;     // whole groups exit together
    const WorkItem it = items[g];
    const uint32_t len = it.len & ~kSplitFlag;
    const uint32_t *p = idx + it.begin;
    const uint4 *base = in + u;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t j = 0;
    if (PIPE == 1) {
        // software pipeline: the index loads of step j + 1 and the row gathers of step j are in flight while the rows
        // of step j - 1 are accumulated, so a warp keeps two dependent memory round trips overlapped instead of
        // serialising them (the kernel is then far less sensitive to how many warps share the SM with it)
        const uint32_t nfull = len >> 3;
        uint32_t i0 = 0, i1 = 0;
        if (nfull) { i0 = ld_idx<true>(p + sub); i1 = ld_idx<true>(p + 4 + sub); }
        uint4 wa, wb, wc, wd;
        bool have = false;
        for (uint32_t s = 0; s < nfull; ++s) {
            const uint32_t k0 = __shfl_xor_sync(gmask, i0, 1), k1 = __shfl_xor_sync(gmask, i1, 1);
            const uint32_t ra = u ? k0 : i0, rb = u ? i0 : k0, rc = u ? k1 : i1, rd = u ? i1 : k1;
            if (s + 1 < nfull) { i0 = ld_idx<true>(p + 8 * (s + 1) + sub); i1 = ld_idx<true>(p + 8 * (s + 1) + 4 + sub); }
            const uint4 na = __ldg(base + (size_t)ra * 2), nb = __ldg(base + (size_t)rb * 2);
            const uint4 nc = __ldg(base + (size_t)rc * 2), nd = __ldg(base + (size_t)rd * 2);
            if (have) { bfp_accum5(wa, acc); bfp_accum5(wb, acc); bfp_accum5(wc, acc); bfp_accum5(wd, acc); }
            wa = na; wb = nb; wc = nc; wd = nd;
            have = true;
        }
        if (have) { bfp_accum5(wa, acc); bfp_accum5(wb, acc); bfp_accum5(wc, acc); bfp_accum5(wd, acc); }
        j = nfull << 3;
    } else {
        for (; j + 8 <= len; j += 8) {
            const uint32_t i0 = ld_idx<true>(p + j + sub), i1 = ld_idx<true>(p + j + 4 + sub);
            const uint32_t k0 = __shfl_xor_sync(gmask, i0, 1), k1 = __shfl_xor_sync(gmask, i1, 1);
            // the pair's four rows in stream order: (even lane's, odd lane's) x (first, second load)
            const uint32_t ra = u ? k0 : i0, rb = u ? i0 : k0, rc = u ? k1 : i1, rd = u ? i1 : k1;
            constexpr int M = PIPE == 2 ? 1 : (PIPE == 3 ? 2 : 0);
            const uint4 wa = ld_row<M>(base + (size_t)ra * 2), wb = ld_row<M>(base + (size_t)rb * 2);
            const uint4 wc = ld_row<M>(base + (size_t)rc * 2), wd = ld_row<M>(base + (size_t)rd * 2);
            bfp_accum5(wa, acc); bfp_accum5(wb, acc); bfp_accum5(wc, acc); bfp_accum5(wd, acc);
        }
    }
    if (j < len) {
        const uint32_t n = len - j;     // 1..7 left: positions sub and 4 + sub
        const uint32_t i0 = (uint32_t)sub < n ? ld_idx<true>(p + j + sub) : 0u;
        const uint32_t i1 = (uint32_t)sub + 4 < n ? ld_idx<true>(p + j + 4 + sub) : 0u;
        const uint32_t k0 = __shfl_xor_sync(gmask, i0, 1), k1 = __shfl_xor_sync(gmask, i1, 1);
        const uint32_t pe = (uint32_t)(sub & 2), po = pe + 1;       // stream positions of the pair's even / odd lane
        const uint32_t ra = u ? k0 : i0, rb = u ? i0 : k0, rc = u ? k1 : i1, rd = u ? i1 : k1;
        if (pe < n) bfp_accum5(__ldg(base + (size_t)ra * 2), acc);
        if (po < n) bfp_accum5(__ldg(base + (size_t)rb * 2), acc);
        if (pe + 4 < n) bfp_accum5(__ldg(base + (size_t)rc * 2), acc);
        if (po + 4 < n) bfp_accum5(__ldg(base + (size_t)rd * 2), acc);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] += __shfl_xor_sync(gmask, acc[k], 2);
    if (sub >= 2) return;
    float *o = out + (size_t)it.out_row * kStride + 5 * u;
    const int ncol = min(5, b - 5 * u);
    if (ADD || (it.len & kSplitFlag)) {     // ADD: the head engine may be adding into the same row right now
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (k < ncol) atomicAdd(o + k, acc[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (k < ncol) o[k] = acc[k];
    }
}

__device__ __forceinline__ void split3(float v, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l)
{
    h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);          // exact
    m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);         // exact
    l = __float2bfloat16_rn(r2);
}

// colmax_bits[j] = bit pattern of max_w |scale[w] X[w + j ld]| (non-negative floats order like unsigned integers): the
// column maxima of the operand rows as the passes see them (Xs = diag(sqrt_zeta) X)
__global__ void __launch_bounds__(1024)
colmax_kernel(const float *__restrict__ X, size_t ld, uint32_t n, int b, const float *__restrict__ scale,
              uint32_t *__restrict__ colmax_bits)
{
    float m[kStride];
#pragma unroll
    for (int j = 0; j < kStride; ++j) m[j] = 0.f;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
        const float s = scale ? scale[w] : 1.0f;
#pragma unroll
        for (int j = 0; j < kStride; ++j)
            if (j < b) m[j] = fmaxf(m[j], fabsf(s * X[w + (size_t)j * ld]));
    }
    __shared__ uint32_t smax[kStride];
    if (threadIdx.x < kStride) smax[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kStride; ++j) {
        if (j < b) {
            float v = m[j];
#pragma unroll
            for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(smax + j, __float_as_uint(v));
        }
    }
    __syncthreads();
    if (threadIdx.x < b && smax[threadIdx.x]) atomicMax(colmax_bits + threadIdx.x, smax[threadIdx.x]);     // one atomic per column per CTA
}

// Exact power-of-two column equilibration for the block-FP rows: columns that share a unit get the same
// magnitude, so the shared exponent costs no precision whatever the scales of the caller's columns are.
// down = 2^-e, up = 2^e with 2^e <= colmax < 2^(e+1); 1 when the column is zero / not finite.
__device__ __forceinline__ float col_pow2(uint32_t colmax_bits, bool up)
{
    const uint32_t E = (colmax_bits >> 23) & 0xFFu;
    if (E == 0 || E >= 254) return 1.0f;
    return __uint_as_float((up ? E : 254u - E) << 23);
}

// Rank space: Xs[r, j] = sqrt_zeta[w] * X[w + j ld], w = word_of_rank[r]  (j < b), zero padded to 16
// floats; head ranks are also written as three bf16 pieces, K-major: xsplit[(piece*BS + j) * H + r].
__global__ void __launch_bounds__(256)
pack_scaled_kernel(const float *__restrict__ X, size_t ld, uint32_t n, int b, const float *__restrict__ scale,
                   const uint32_t *__restrict__ word_of_rank, float4 *__restrict__ out, uint4 *__restrict__ out_bfp,
                   const uint32_t *__restrict__ colmax_bits, uint32_t H, int BS, __nv_bfloat16 *__restrict__ xsplit,
                   int8_t *__restrict__ xdig)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t w = word_of_rank[r];
    const float s = scale ? scale[w] : 1.0f;
    float v[kStride];
#pragma unroll
    for (int j = 0; j < kStride; ++j)
        v[j] = (j < b) ? (s * X[w + (size_t)j * ld]) * (colmax_bits ? col_pow2(colmax_bits[j], false) : 1.0f) : 0.0f;
    if (out_bfp) {      // b <= 10: one-sector rows
        out_bfp[(size_t)r * 2] = bfp_encode5(v);
        out_bfp[(size_t)r * 2 + 1] = bfp_encode5(v + 5);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            out[(size_t)r * 4 + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (r < H && xdig) {
        // int8 head engine: the equilibrated values (column max in [1, 2)) as 30-bit integers, four s8 digits, K-major
#pragma unroll
        for (int j = 0; j < kStride; ++j) {
            if (j < b) {
                int8_t dg[4];
                digits4(__float2int_rn(v[j] * 536870912.0f), dg);      // |v| < 2: |q| <= 2^30
#pragma unroll
                for (int g = 0; g < 4; ++g) xdig[(size_t)head8_row(j, g) * H + r] = dg[g];
            }
        }
    } else if (r < H) {
#pragma unroll
        for (int j = 0; j < kStride; ++j) {
            if (j < b) {
                __nv_bfloat16 h, m, l;
                split3(v[j], h, m, l);
                xsplit[(size_t)j * H + r] = h;
                xsplit[(size_t)(BS + j) * H + r] = m;
                xsplit[(size_t)(2 * BS + j) * H + r] = l;
            }
        }
    }
}

// ycolmax_bits[j] = bit pattern of max_d |Y[d, j]| over the padded row-major rows of Y
__global__ void __launch_bounds__(1024)
rowmajor_colmax_kernel(const float4 *__restrict__ Y, uint32_t DB, int b, uint32_t *__restrict__ colmax_bits)
{
    float m[kStride];
#pragma unroll
    for (int j = 0; j < kStride; ++j) m[j] = 0.f;
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < DB; d += gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (4 * q < b) {
                const float4 t = Y[(size_t)d * 4 + q];
                m[4 * q] = fmaxf(m[4 * q], fabsf(t.x)); m[4 * q + 1] = fmaxf(m[4 * q + 1], fabsf(t.y));
                m[4 * q + 2] = fmaxf(m[4 * q + 2], fabsf(t.z)); m[4 * q + 3] = fmaxf(m[4 * q + 3], fabsf(t.w));
            }
        }
    }
    __shared__ uint32_t smax[kStride];
    if (threadIdx.x < kStride) smax[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kStride; ++j) {
        if (j < b) {
            float v = m[j];
#pragma unroll
            for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(smax + j, __float_as_uint(v));
        }
    }
    __syncthreads();
    if (threadIdx.x < b && smax[threadIdx.x]) atomicMax(colmax_bits + threadIdx.x, smax[threadIdx.x]);     // one atomic per column per CTA
}

// ysplit[(piece*BS + j) * DBpad + d] = pieces of Y[d, j] (head engine, when ysplit != NULL);
// ybfp[d] = one-sector copy of Y[d, :] (tail gather of pass 2, when ybfp != NULL)
__global__ void __launch_bounds__(256)
ysplit_kernel(const float4 *__restrict__ Y, uint32_t DB, uint32_t DBpad, int b, int BS, __nv_bfloat16 *__restrict__ ysplit,
              uint4 *__restrict__ ybfp, int8_t *__restrict__ ydig, const uint32_t *__restrict__ ycolmax_bits)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= DB) return;
    float v[kStride];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = Y[(size_t)d * 4 + q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    if (ybfp) {
        ybfp[(size_t)d * 2] = bfp_encode5(v);
        ybfp[(size_t)d * 2 + 1] = bfp_encode5(v + 5);
    }
    if (ydig) {      // int8 head engine: quantised with the column maxima of Y
#pragma unroll
        for (int j = 0; j < kStride; ++j) {
            if (j < b) {
                int8_t dg[4];
                digits4(__float2int_rn(v[j] * quant_up(quant_exp(ycolmax_bits, j))), dg);
#pragma unroll
                for (int g = 0; g < 4; ++g) ydig[(size_t)head8_row(j, g) * DBpad + d] = dg[g];
            }
        }
        return;
    }
    if (!ysplit) return;
#pragma unroll
    for (int j = 0; j < kStride; ++j) {
        if (j < b) {
            __nv_bfloat16 h, m, l;
            split3(v[j], h, m, l);
            ysplit[(size_t)j * DBpad + d] = h;
            ysplit[(size_t)(BS + j) * DBpad + d] = m;
            ysplit[(size_t)(2 * BS + j) * DBpad + d] = l;
        }
    }
}

// Z[w + j ld] = sqrt_zeta[w] * T[rank_of[w], j]
__global__ void __launch_bounds__(256)
unpack_scaled_kernel(const float *__restrict__ T, uint32_t n, int b, const float *__restrict__ scale,
                     const uint32_t *__restrict__ rank_of, const uint32_t *__restrict__ colmax_bits, float *__restrict__ Z, size_t ld)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const float s = scale ? scale[w] : 1.0f;
    const float4 *t4 = reinterpret_cast<const float4 *>(T) + (size_t)rank_of[w] * 4;
    float v[kStride];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = t4[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < kStride; ++j)
        if (j < b) Z[w + (size_t)j * ld] = (s * v[j]) * (colmax_bits ? col_pow2(colmax_bits[j], true) : 1.0f);
}

// The same fused with the all-reduce of the document-sharded operator (Z = sum over ranks of B_g (B_g^T X)): ONE
// kernel un-permutes and scales this rank's T straight into the peer-visible stage area, and then reduces and redistributes
// it over NVLink with plain peer loads / stores (p2p.cuh: rank r sums slice r of every rank's stage in rank order and
// stores the sums into every rank's result area).  No NCCL call, no separate un-pack pass, no staging copy.
struct UnpackProducer {
    const float *T;
    uint32_t n;
    int b;
    const float *scale;
    const uint32_t *rank_of, *colmax_bits;
    __device__ __forceinline__ void stage(char *stage_area, size_t nelem, size_t tid, size_t nth) const
    {
        float *Zs = reinterpret_cast<float *>(stage_area);
        for (size_t w = tid; w < n; w += nth) {
            const float s = scale ? scale[w] : 1.0f;
            const float4 *t4 = reinterpret_cast<const float4 *>(T) + (size_t)rank_of[w] * 4;
            float v[kStride];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = t4[q];
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < kStride; ++j)
                if (j < b) Zs[w + (size_t)j * n] = (s * v[j]) * (colmax_bits ? col_pow2(colmax_bits[j], true) : 1.0f);
        }
        // pad lanes of the last 16-byte vector
        const size_t padded = (nelem + 3) / 4 * 4;
        for (size_t i = nelem + tid; i < padded; i += nth) Zs[i] = 0.f;
    }
};
__global__ void __launch_bounds__(512)
unpack_allreduce_kernel(P2pArgs a, UnpackProducer prod, float *__restrict__ Z, size_t nelem)
{
    p2p_allreduce2_body<float, 0>(a, prod, Z, nelem);
}

static void launch_gather(Ctx &c, int nch, bool add, const WorkItem *items, size_t n_items, const uint32_t *idx,
                          const float *in, float *out)
{
    if (!n_items) return;
    const unsigned grid = (unsigned)((n_items * 4 + 255) / 256);
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    const bool na = c.opt("spmm_idx_noalloc", 1) != 0;
#define ISLE_SPMM_LAUNCH(N)                                                                                           \
    do {                                                                                                              \
        if (add) spmm_gather_kernel<N, true, true><<<grid, 256, 0, c.stream>>>(items, n_items, idx, in4, out4);        \
        else if (na) spmm_gather_kernel<N, true, false><<<grid, 256, 0, c.stream>>>(items, n_items, idx, in4, out4);   \
        else spmm_gather_kernel<N, false, false><<<grid, 256, 0, c.stream>>>(items, n_items, idx, in4, out4);          \
    } while (0)
    switch (nch) {
    case 1: ISLE_SPMM_LAUNCH(1); break;
    case 2: ISLE_SPMM_LAUNCH(2); break;
    case 3: ISLE_SPMM_LAUNCH(3); break;
    default: ISLE_SPMM_LAUNCH(4); break;
    }
#undef ISLE_SPMM_LAUNCH
    count_launch(c);
}

static void launch_gather_bfp(Ctx &c, int b, bool add, const WorkItem *items, size_t n_items, const uint32_t *idx,
                              const uint4 *in, float *out, cudaStream_t stream = nullptr)
{
    if (!n_items) return;
    if (!stream) stream = c.stream;
    const unsigned grid = (unsigned)((n_items * 4 + 255) / 256);
    const int pipe = c.opt("spmm_tail_pipe", 0);      // 0 plain, 1 software pipeline, 2 rows with L1::no_allocate, 3 rows with L1::evict_first
    const int carve = c.opt("spmm_tail_carveout", -1);      // experiment: shared-memory carve-out the gather kernel asks for (percent)
#define ISLE_BFP_LAUNCH(A, P)                                                                                             \
    do {                                                                                                                  \
        if (carve >= 0) cudaFuncSetAttribute(spmm_gather_bfp_kernel<A, P>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
        spmm_gather_bfp_kernel<A, P><<<grid, 256, 0, stream>>>(items, n_items, idx, in, out, b);                          \
    } while (0)
    if (add) { if (pipe == 1) ISLE_BFP_LAUNCH(true, 1); else if (pipe == 2) ISLE_BFP_LAUNCH(true, 2); else if (pipe == 3) ISLE_BFP_LAUNCH(true, 3); else ISLE_BFP_LAUNCH(true, 0); }
    else { if (pipe == 1) ISLE_BFP_LAUNCH(false, 1); else if (pipe == 2) ISLE_BFP_LAUNCH(false, 2); else if (pipe == 3) ISLE_BFP_LAUNCH(false, 3); else ISLE_BFP_LAUNCH(false, 0); }
#undef ISLE_BFP_LAUNCH
    count_launch(c);
}

// Experiment hook (option spmm_dummy_head): a kernel that only OCCUPIES what the head engine occupies on every SM -- one CTA,
// 1024 thread slots (the tail keeps 4 of its 8 CTAs, as beside the real head), `smem` bytes of shared memory -- and sleeps
// for `ns`.  Tells how fast the tail gather runs beside an idle co-tenant, i.e. whether sharing the SM or sharing the
// L2 -> SM fabric is what slows it down beside the real head.
__global__ void __launch_bounds__(1024)
occupy_kernel(unsigned long long ns)
{
    extern __shared__ uint8_t dummy_smem[];
    if (threadIdx.x == 0) dummy_smem[0] = 1;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < ns);
}

void spsptr_multiply_dev(Ctx &c, int b, const float *X, float *Z)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "spsptr_multiply: build_B first");
    ISLE_REQUIRE(b >= 1 && b <= kStride, ISLE_ERR_ARG, "spsptr_multiply: block size must be in [1,16]");
    // block sizes <= 10 gather one-sector (32-byte) operand rows; wider blocks use padded fp32 rows
    const bool bfp = b <= 10 && c.opt("spmm_bfp", 1) != 0;
    // the int8 head engine needs the column equilibration of the block-FP path and N = 3 x 10 columns
    build_csr(c, bfp);
    const uint32_t V = (uint32_t)c.V, DB = (uint32_t)c.DB, H = c.H;
    const bool i8 = H && c.head_i8;
    const int nch = (b + 3) / 4;
    const int BS = head_block_stride(b);
    c.counters["ks_ops"] += 1.0;

    if (bfp) {
        ISLE_CUDA_CHECK(cudaMemsetAsync(c.colmax.p, 0, c.colmax.bytes(), c.stream));
        colmax_kernel<<<std::min<unsigned>((V + 1023) / 1024, (unsigned)c.num_sms), 1024, 0, c.stream>>>(X, (size_t)V, V, b, c.sqrt_zeta.p, c.colmax.p);
        count_launch(c);
    }
    pack_scaled_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(X, (size_t)V, V, b, c.sqrt_zeta.p, c.word_of_rank.p,
                                                               reinterpret_cast<float4 *>(c.xs.p), bfp ? c.xbfp.p : nullptr,
                                                               bfp ? c.colmax.p : nullptr, H, BS, c.xsplit.p, i8 ? c.xdig.p : nullptr);
    count_launch(c);
    // T is always zeroed: word rows are split routinely and the head adds partial sums atomically.
    ISLE_CUDA_CHECK(cudaMemsetAsync(c.zbuf.p, 0, c.zbuf.bytes(), c.stream));
    // head engine of pass 1 (rows = documents, K = head ranks) / pass 2 (rows = head ranks, K = documents)
    auto head1 = [&](bool zero_out, bool force_atomic, cudaStream_t st) {
        if (const int dummy_us = c.opt("spmm_dummy_head", 0)) {
            const int smem = c.opt("spmm_dummy_smem_kb", 57) * 1024;
            cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            occupy_kernel<<<c.num_sms, 1024, smem, st>>>((unsigned long long)dummy_us * 1000ull);
            return;
        }
        if (i8) spmm_head_i8_launch(c, b, c.bits1.p, c.DBpad / kHeadTile, H / 256, 1, c.xdig.p, nullptr, c.ybuf.p, DB, zero_out, force_atomic, st);
        else spmm_head_launch(c, b, c.bits1.p, c.DBpad / kHeadTile, H / kHeadChunk, 1, c.xsplit.p, c.ybuf.p, DB, zero_out, force_atomic, st);
    };
    auto head2 = [&](cudaStream_t st) {
        const uint32_t mt = H / kHeadTile;
        // K (documents) split over jobs: whole waves of jobs over the SMs
        const uint32_t nsplit = std::max<uint32_t>(1, ((uint32_t)c.num_sms * (uint32_t)std::max(1, c.opt("spmm_head2_waves", 1))) / mt);
        if (i8) spmm_head_i8_launch(c, b, c.bits2.p, mt, c.DBpad / 256, nsplit, c.ydig.p, c.ycolmax.p, c.zbuf.p, H, false, false, st);
        else spmm_head_launch(c, b, c.bits2.p, mt, c.DBpad / kHeadChunk, nsplit, c.ysplit.p, c.zbuf.p, H, false, false, st);
    };
    {
        // bytes_pass = nnz*4 + (rows+1)*8 + dense_in*4 + dense_out*4
        StatScope s(c, "spmm_bt", (double)c.nnzB * 4.0 + ((double)DB + 1) * 8.0 + ((double)V + DB) * b * 4.0,
                    2.0 * c.nnzB * b);
        if (H && bfp && c.opt("spmm_fork", 1) != 0) {
            // head and tail add into a zeroed Y side by side (TMEM / tensor pipe vs L1TEX-bound gathers).  The head goes
            // first, on the main stream directly behind the memset, so that its one persistent CTA per SM is placed
            // before the tail's thousands of small CTAs fill every SM (a head launched second would wait for them all).
            ISLE_CUDA_CHECK(cudaMemsetAsync(c.ybuf.p, 0, c.ybuf.bytes(), c.stream));
            ISLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, c.stream));
            ISLE_CUDA_CHECK(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
            head1(false, true, c.stream);
            launch_gather_bfp(c, b, true, c.items_bt.p, c.n_items_bt, c.t1_idx.p, c.xbfp.p, c.ybuf.p, c.stream2);
            ISLE_CUDA_CHECK(cudaEventRecord(c.ev_join, c.stream2));
            ISLE_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
        } else if (H) {
            {
                StatScope sh(c, "spmm_head1");
                // every row of Y is written (store mode): Y = head part
                head1(true, false, c.stream);
            }
            StatScope st(c, "spmm_tail1");
            if (bfp) launch_gather_bfp(c, b, true, c.items_bt.p, c.n_items_bt, c.t1_idx.p, c.xbfp.p, c.ybuf.p);
            else launch_gather(c, nch, true, c.items_bt.p, c.n_items_bt, c.t1_idx.p, c.xs.p, c.ybuf.p);
        } else {
            // a document without items (or split into several), and the columns >= b, need zeroed rows
            if (bfp || c.n_items_bt != (size_t)DB) ISLE_CUDA_CHECK(cudaMemsetAsync(c.ybuf.p, 0, c.ybuf.bytes(), c.stream));
            if (bfp) launch_gather_bfp(c, b, false, c.items_bt.p, c.n_items_bt, c.t1_idx.p, c.xbfp.p, c.ybuf.p);
            else launch_gather(c, nch, false, c.items_bt.p, c.n_items_bt, c.t1_idx.p, c.xs.p, c.ybuf.p);
        }
    }
    {
        StatScope s(c, "spmm_b", (double)c.nnzB * 4.0 + ((double)V + 1) * 8.0 + ((double)V + DB) * b * 4.0,
                    2.0 * c.nnzB * b);
        if ((H || bfp) && DB) {
            if (i8) {      // quantisation scale of Y: its column maxima
                ISLE_CUDA_CHECK(cudaMemsetAsync(c.ycolmax.p, 0, c.ycolmax.bytes(), c.stream));
                rowmajor_colmax_kernel<<<std::min<unsigned>((DB + 1023) / 1024, (unsigned)c.num_sms), 1024, 0, c.stream>>>(
                    reinterpret_cast<const float4 *>(c.ybuf.p), DB, b, c.ycolmax.p);
                count_launch(c);
            }
            ysplit_kernel<<<(DB + 255) / 256, 256, 0, c.stream>>>(reinterpret_cast<const float4 *>(c.ybuf.p), DB, c.DBpad, b, BS,
                                                                  (H && !i8) ? c.ysplit.p : nullptr, bfp ? c.ybfp.p : nullptr,
                                                                  i8 ? c.ydig.p : nullptr, c.ycolmax.p);
            count_launch(c);
        }
        if (H) {
            const bool fork = c.opt("spmm_fork", 1) != 0;
            if (fork) {
                // the head and the tail of pass 2 write disjoint rows of T: run them side by side
                ISLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, c.stream));
                ISLE_CUDA_CHECK(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
                if (bfp) {      // head first on the main stream (see pass 1), tail beside it
                    head2(c.stream);
                    launch_gather_bfp(c, b, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybfp.p, c.zbuf.p, c.stream2);
                    ISLE_CUDA_CHECK(cudaEventRecord(c.ev_join, c.stream2));
                } else {
                    head2(c.stream2);
                    ISLE_CUDA_CHECK(cudaEventRecord(c.ev_join, c.stream2));
                    launch_gather(c, nch, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybuf.p, c.zbuf.p);
                }
                ISLE_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
            } else {
                {
                    StatScope sh(c, "spmm_head2");
                    head2(c.stream);
                }
                StatScope st(c, "spmm_tail2");
                if (bfp) launch_gather_bfp(c, b, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybfp.p, c.zbuf.p);
                else launch_gather(c, nch, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybuf.p, c.zbuf.p);
            }
        } else {
            if (bfp) launch_gather_bfp(c, b, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybfp.p, c.zbuf.p);
            else launch_gather(c, nch, false, c.items_b.p, c.n_items_b, c.t2_idx.p, c.ybuf.p, c.zbuf.p);
        }
    }
    if (c.world > 1 && c.opt("spmm_fused_allreduce", 1) != 0) {
        P2pArgs a;
        unsigned grid = 0;
        if (p2p_two_shot_begin(c, (size_t)V * b * 4, &a, &grid)) {
            StatScope s(c, "unpack_allreduce");
            UnpackProducer prod{c.zbuf.p, V, b, c.sqrt_zeta.p, c.rank_of.p, bfp ? c.colmax.p : nullptr};
            unpack_allreduce_kernel<<<grid, 512, 0, c.stream>>>(a, prod, Z, (size_t)V * b);
            ISLE_CUDA_CHECK(cudaGetLastError());
            count_launch(c);
            return;
        }
    }
    unpack_scaled_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(c.zbuf.p, V, b, c.sqrt_zeta.p, c.rank_of.p,
                                                                 bfp ? c.colmax.p : nullptr, Z, (size_t)V);
    count_launch(c);
    if (c.world > 1) {
        // doc-sharded: Z = sum over ranks of B_g (B_g^T X)   
        if (b * (size_t)V) allreduce_sum_f32(c, Z, (size_t)V * b);
    }
}

}  // namespace isle
