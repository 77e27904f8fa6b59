// spmm.cu -- kernel family (2): the B * B^T * X operator that drives block Krylov-Schur.
//
// Replaces MKL_SpSpTrProd (reference include/matUtils.h:14-419): the ctor's CSC->CSR copy by
// sorting (:109-135) is a stable device radix sort of (word, doc) pairs; multiply()
// (:336-365, two mkl_scsrmm passes) is two launches of one gather kernel:
//     pass 1   Y[d,:] = sum_{w in doc d}  Xs[w,:]          Xs = diag(sqrt_zeta) X
//     pass 2   T[w,:] = sum_{d containing w} Y[d,:]        Z  = diag(sqrt_zeta) T
// B is a 0/1 pattern scaled by sqrt_zeta per row, so no value stream is read:
// 4 bytes per nonzero.  Dense operands are row-major with a fixed 16-float (64 B = two
// 32 B sectors) row stride so a 4-lane group fetches one row with one float4 per lane.
// Work is cut into items (row, first nnz, length <= ISLE_SPMM_CHUNK) sorted by decreasing
// length, so the eight 4-lane groups of a warp run equally long loops despite Zipfian row
// lengths; rows cut into several items finish with float4 atomics.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace isle {

static constexpr int kStride = 16;          // floats per padded dense row
static constexpr uint32_t kChunk = 2048;    // max nonzeros per work item
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
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&total, scan.p + nrows, 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
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

void build_csr(Ctx &c)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "build_csr: build_B first");
    if (c.have_csr) return;
    StatScope s(c, "csr_build", (double)c.nnzB * 24.0);
    const int64_t n = c.nnzB;
    const uint32_t V = (uint32_t)c.V, DB = (uint32_t)c.DB;
    c.csr_col.alloc((size_t)n);
    c.csr_off.alloc((size_t)V + 1);
    {
        DevBuf<uint32_t> doc_of((size_t)n), keys_out((size_t)n);
        if (DB) {
            fill_doc_ids_kernel<<<grid_for((size_t)DB * 32, 256, c.num_sms * 16), 256, 0, c.stream>>>(c.b_off.p, DB, doc_of.p);
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
    build_items(c, c.b_off.p, DB, c.items_bt, c.n_items_bt);
    build_items(c, c.csr_off.p, V, c.items_b, c.n_items_b);
    c.xs.alloc((size_t)V * kStride);
    c.ybuf.alloc((size_t)std::max<uint64_t>(DB, 1) * kStride);
    c.zbuf.alloc((size_t)V * kStride);
    c.have_csr = true;
}

// --------------------------------------------------------------------------- the gather pass
// One 4-lane group per work item; lane `sub` owns floats [4 sub, 4 sub + 4) of the row.
// Per step each lane loads one index (coalesced 16 B per group), indices are exchanged by
// shuffle and four independent float4 gathers are issued before they are summed, in a fixed
// order so that results are run-to-run deterministic for unsplit rows.
// Index stream loads: read once, so optionally kept out of L1 (ld.global.nc.L1::no_allocate) to leave
// the cache to the gathered dense rows.
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

template <int NCH, bool NOALLOC>
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
    if (it.len & kSplitFlag) atomicAdd(o, acc);   // sm_90+: 16-byte vector atomic (RED.128)
    else *o = acc;
}

// Xs[w, j] = sqrt_zeta[w] * X[w + j ld]  (j < b), zero padded to 16 floats.
__global__ void __launch_bounds__(256)
pack_scaled_kernel(const float *__restrict__ X, size_t ld, uint32_t n, int b,
                   const float *__restrict__ scale, float4 *__restrict__ out)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const float s = scale ? scale[w] : 1.0f;
    float v[kStride];
#pragma unroll
    for (int j = 0; j < kStride; ++j) v[j] = (j < b) ? s * X[w + (size_t)j * ld] : 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        out[(size_t)w * 4 + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// Z[w + j ld] = sqrt_zeta[w] * T[w, j]
__global__ void __launch_bounds__(256)
unpack_scaled_kernel(const float *__restrict__ T, uint32_t n, int b, const float *__restrict__ scale,
                     float *__restrict__ Z, size_t ld)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const float s = scale ? scale[w] : 1.0f;
    const float4 *t4 = reinterpret_cast<const float4 *>(T) + (size_t)w * 4;
    float v[kStride];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = t4[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < kStride; ++j)
        if (j < b) Z[w + (size_t)j * ld] = s * v[j];
}

static void launch_gather(Ctx &c, int nch, const WorkItem *items, size_t n_items, const uint32_t *idx,
                          const float *in, float *out)
{
    if (!n_items) return;
    const unsigned grid = (unsigned)((n_items * 4 + 255) / 256);
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    const bool na = c.opt("spmm_idx_noalloc", 1) != 0;
#define ISLE_SPMM_LAUNCH(N)                                                                                    \
    do {                                                                                                       \
        if (na) spmm_gather_kernel<N, true><<<grid, 256, 0, c.stream>>>(items, n_items, idx, in4, out4);       \
        else spmm_gather_kernel<N, false><<<grid, 256, 0, c.stream>>>(items, n_items, idx, in4, out4);         \
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

void spsptr_multiply_dev(Ctx &c, int b, const float *X, float *Z)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "spsptr_multiply: build_B first");
    ISLE_REQUIRE(b >= 1 && b <= kStride, ISLE_ERR_ARG, "spsptr_multiply: block size must be in [1,16]");
    build_csr(c);
    const uint32_t V = (uint32_t)c.V, DB = (uint32_t)c.DB;
    const int nch = (b + 3) / 4;
    c.counters["ks_ops"] += 1.0;

    pack_scaled_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(X, (size_t)V, V, b, c.sqrt_zeta.p,
                                                               reinterpret_cast<float4 *>(c.xs.p));
    count_launch(c);
    // Every doc of B has >= 1 nonzero and is never split unless longer than kChunk: only then
    // does Y need zeroing; word rows are split routinely (Zipf head), so T is always zeroed.
    if (c.n_items_bt != (size_t)DB) ISLE_CUDA_CHECK(cudaMemsetAsync(c.ybuf.p, 0, c.ybuf.bytes(), c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(c.zbuf.p, 0, c.zbuf.bytes(), c.stream));
    {
        // bytes_pass = nnz*4 + (rows+1)*8 + dense_in*4 + dense_out*4
        StatScope s(c, "spmm_bt", (double)c.nnzB * 4.0 + ((double)DB + 1) * 8.0 + ((double)V + DB) * b * 4.0,
                    2.0 * c.nnzB * b);
        launch_gather(c, nch, c.items_bt.p, c.n_items_bt, c.b_row.p, c.xs.p, c.ybuf.p);
    }
    {
        StatScope s(c, "spmm_b", (double)c.nnzB * 4.0 + ((double)V + 1) * 8.0 + ((double)V + DB) * b * 4.0,
                    2.0 * c.nnzB * b);
        launch_gather(c, nch, c.items_b.p, c.n_items_b, c.csr_col.p, c.ybuf.p, c.zbuf.p);
    }
    unpack_scaled_kernel<<<(V + 255) / 256, 256, 0, c.stream>>>(c.zbuf.p, V, b, c.sqrt_zeta.p, Z, (size_t)V);
    count_launch(c);
    if (c.world > 1) {
        // doc-sharded: Z = sum over ranks of B_g (B_g^T X)   
        if (b * (size_t)V) allreduce_sum_f32(c, Z, (size_t)V * b);
    }
}

}  // namespace isle
