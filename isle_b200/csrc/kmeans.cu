// kmeans.cu -- kernel family (3): k-means++ seeding and Lloyd's iterations on the rank-k
// projection of the documents.
//
// The reference never materialises the projection (USE_EXPLICIT_PROJECTED_MATRIX false,
// include/hyperparams.h:44): per 2^18-doc block it forms -2 U C^T (vocab x k) and multiplies by
// the sparse block (src/sparseMatrix.cpp:1794-1849).  With 180 GB of HBM the idiomatic form is
// the reference's own disabled "explicit" one (src/denseMatrix.cpp:504-530): materialise
// P = B^T U once (docs x k), after which
//     dist[d,c] = ((-2 P_d . C_c) + ||C_c||^2) + ||P_d||^2          (order of :1820-1846)
// is a dense docs x centers contraction with a fused epilogue:
//     assignment  argmin_c |dist[d,c]|, first index on ties (cblas_isamin, :1868-1870)
//     k-means++   min_dist[d] = min(min_dist[d], max(dist, 0))      (:2112-2126)
// followed by one pass over P that accumulates center sums and counts (:1959-1992).
// dist_kernel selects the contraction engine: 0 = SIMT fp32 FMA tile kernel (this file),
// 1 = tcgen05 split-TF32 kernel (dist_tc.cu).
#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <random>

#include "common.cuh"

namespace isle {

void dist_tc_launch(Ctx &c, const float *P, const float *d2, uint32_t DB, uint32_t kp, const float *C,
                    const float *c2, uint32_t ncent, int mode, uint32_t *assign, float *min_dist);
bool dist_tc_supported(const Ctx &c, uint32_t kp, uint32_t ncent);

// ------------------------------------------------------------------------------ projection
// Us[w, j] = sqrt_zeta[w] * U[w + j V], row-major with row stride kp (zero padded).
__global__ void __launch_bounds__(256)
transpose_scale_kernel(const float *__restrict__ U, uint32_t V, uint32_t k, uint32_t kp,
                       const float *__restrict__ scale, float *__restrict__ Us)
{
    __shared__ float tile[32][33];
    const uint32_t w0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (uint32_t r = ty; r < 32; r += 8) {
        const uint32_t j = j0 + r, w = w0 + tx;
        tile[r][tx] = (j < k && w < V) ? U[(size_t)j * V + w] : 0.0f;
    }
    __syncthreads();
    for (uint32_t r = ty; r < 32; r += 8) {
        const uint32_t w = w0 + r, j = j0 + tx;
        if (w < V && j < kp) Us[(size_t)w * kp + j] = tile[tx][r] * scale[w];
    }
}

// One warp per document: P[d,:] = sum_{w in d} Us[w,:]; l2[d] = ||P[d,:]||^2
// (multiply_with / UT_times_docs + compute_projected_docs_l2sq, src/sparseMatrix.cpp:1749-1791,1888-1918)
__global__ void __launch_bounds__(256)
project_kernel(const uint32_t *__restrict__ b_row, const int64_t *__restrict__ b_off, uint32_t DB,
               const float4 *__restrict__ Us, uint32_t kp4, float4 *__restrict__ P, float *__restrict__ l2)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = b_off[d], e = b_off[d + 1];
        float nrm = 0.f;
        for (uint32_t c0 = 0; c0 < kp4; c0 += 32) {
            const uint32_t col = c0 + lane;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < kp4) {
                int64_t p = b;
                for (; p + 4 <= e; p += 4) {
                    const uint32_t w0 = b_row[p], w1 = b_row[p + 1], w2 = b_row[p + 2], w3 = b_row[p + 3];
                    const float4 v0 = __ldg(Us + (size_t)w0 * kp4 + col);
                    const float4 v1 = __ldg(Us + (size_t)w1 * kp4 + col);
                    const float4 v2 = __ldg(Us + (size_t)w2 * kp4 + col);
                    const float4 v3 = __ldg(Us + (size_t)w3 * kp4 + col);
                    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                    acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                    acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
                }
                for (; p < e; ++p) {
                    const float4 v = __ldg(Us + (size_t)b_row[p] * kp4 + col);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                P[(size_t)d * kp4 + col] = acc;
                nrm += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0) l2[d] = nrm;
    }
}

void project(Ctx &c)
{
    ISLE_REQUIRE(c.have_B && c.have_U, ISLE_ERR_ARG, "project: needs B and U");
    if (c.have_P) return;
    const uint32_t V = (uint32_t)c.V, k = (uint32_t)c.k, DB = (uint32_t)c.DB;
    c.kp = ((uint64_t)k + 31) / 32 * 32;
    const uint32_t kp = (uint32_t)c.kp;
    DevBuf<float> Us((size_t)V * kp);
    c.P.alloc((size_t)std::max<uint32_t>(DB, 1) * kp);
    c.p_l2.alloc(std::max<uint32_t>(DB, 1));
    dim3 tg((V + 31) / 32, kp / 32);
    transpose_scale_kernel<<<tg, 256, 0, c.stream>>>(c.U.p, V, k, kp, c.sqrt_zeta.p, Us.p);
    count_launch(c);
    {
        // SURVEY 8(d) row (2'): flops = 2 nnz k ; compulsory bytes = nnz*8 + D_B k 4 + V k 4
        StatScope s(c, "project", (double)c.nnzB * 8.0 + ((double)DB + V) * k * 4.0, 2.0 * c.nnzB * k);
        if (DB) {
            project_kernel<<<grid_for((size_t)DB * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(
                c.b_row.p, c.b_off.p, DB, reinterpret_cast<const float4 *>(Us.p), kp / 4,
                reinterpret_cast<float4 *>(c.P.p), c.p_l2.p);
            count_launch(c);
        }
    }
    if (c.opt("dist_kernel", 1) == 1 && DB) {
        c.P_lo.alloc((size_t)DB * kp);
        StatScope s(c, "split_p", (double)DB * kp * 8.0);
        split_lo_trunc(c, c.P.p, (size_t)DB * kp, c.P_lo.p);
    } else {
        c.P_lo.release();
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.have_P = true;
}

// ------------------------------------------------------------------- distance contraction
// SIMT engine: 64 docs x 64 centers per CTA step, 16-deep K slabs in shared memory, 4x4
// register tile per thread, fp32 FMA.  MODE 0: abs-argmin (first index).  MODE 1: clamped min.
template <int MODE>
__global__ void __launch_bounds__(256)
dist_simt_kernel(const float *__restrict__ P, const float *__restrict__ d2, uint32_t DB, uint32_t kp,
                 const float *__restrict__ C, const float *__restrict__ c2, uint32_t ncent,
                 uint32_t *__restrict__ assign, float *__restrict__ min_dist)
{
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ float sA[TK][TM + 4];
    __shared__ float sB[TK][TN + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const uint32_t row0 = blockIdx.x * TM;
    float best[4];
    uint32_t besti[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = FLT_MAX; besti[i] = 0; }
    float rd2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = row0 + ty * 4 + i;
        rd2[i] = r < DB ? d2[r] : 0.f;
    }
    for (uint32_t n0 = 0; n0 < ncent; n0 += TN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (uint32_t k0 = 0; k0 < kp; k0 += TK) {
            // 64 rows x 16 k: 1024 floats, 4 per thread as one float4 along k
            {
                const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
                const uint32_t gr = row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < DB) v = *reinterpret_cast<const float4 *>(P + (size_t)gr * kp + k0 + kq);
                sA[kq][r] = v.x; sA[kq + 1][r] = v.y; sA[kq + 2][r] = v.z; sA[kq + 3][r] = v.w;
                const uint32_t gc = n0 + r;
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gc < ncent) u = *reinterpret_cast<const float4 *>(C + (size_t)gc * kp + k0 + kq);
                sB[kq][r] = u.x; sB[kq + 1][r] = u.y; sB[kq + 2][r] = u.z; sB[kq + 3][r] = u.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                const float4 a = *reinterpret_cast<const float4 *>(&sA[kk][ty * 4]);
                const float4 bq = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
        // epilogue for this center tile
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float bl = FLT_MAX;
            uint32_t bi = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t cidx = n0 + tx * 4 + j;
                if (cidx < ncent) {
                    float v = __fadd_rn(__fadd_rn(-2.0f * acc[i][j], __ldg(c2 + cidx)), rd2[i]);
                    v = MODE == 0 ? fabsf(v) : fmaxf(v, 0.0f);
                    if (v < bl) { bl = v; bi = cidx; }
                }
            }
            // reduce over the 16 tx lanes of this row (smaller index wins ties)
#pragma unroll
            for (int o = 8; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bl, o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov < bl || (ov == bl && oi < bi)) { bl = ov; bi = oi; }
            }
            if (bl < best[i]) { best[i] = bl; besti[i] = bi; }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = row0 + ty * 4 + i;
            if (r < DB) {
                if (MODE == 0) assign[r] = besti[i];
                else min_dist[r] = fminf(min_dist[r], best[i]);
            }
        }
    }
}

// k-means++ refresh for a handful of new centers (1 + sqrt(s-5) per round, :2181-2199): one
// HBM-bound pass over P.  The new centers sit in shared memory; a warp owns one document at a time,
// lane l holds columns 4l + 128 i of the row, so the row is read once with coalesced float4 loads and
// each center costs four FMAs per lane per 128 columns plus a shuffle reduction.
// dist = ((-2 P_d . C_c) + ||C_c||^2) + ||P_d||^2, clamped at 0 (:2112-2126).
static constexpr int kSkinnyMax = 16;

// Four documents per warp pass: every center value read from shared memory feeds four rows, which keeps the
// shared-memory traffic (one 16-byte read per center per float4 of P otherwise: 16x the HBM bytes at 16
// centers) under the HBM rate.
static constexpr int kSkinnyDocs = 4;

// NC = 4, 8 or 16: the register tile and the transpose-reduce are sized to the batch (k-means++ adds 1 + sqrt(s - 5) centers
// per round: at k = 100 eleven of the nineteen rounds bring at most 4, seven at most 8)
template <int NC>
__global__ void __launch_bounds__(256)
pp_skinny_kernel(const float4 *__restrict__ P, const float *__restrict__ d2, uint32_t DB, uint32_t kp4,
                 const float4 *__restrict__ C, const float *__restrict__ c2, uint32_t ncent, float *__restrict__ min_dist)
{
    constexpr int LOG = NC == 16 ? 4 : NC == 8 ? 3 : 2;
    extern __shared__ float4 sC[];     // [ncent][kp4]
    for (uint32_t i = threadIdx.x; i < ncent * kp4; i += blockDim.x) sC[i] = C[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kSkinnyDocs;
    const uint32_t stride = ((gridDim.x * blockDim.x) >> 5) * kSkinnyDocs;
    // transpose-reduce: NC partial sums per lane x 32 lanes -> every lane ends with the full sum of ONE center, chosen by
    // its top LOG lane bits (halving exchanges over xor 16, 8, ..., then plain sums over the remaining offsets)
    uint32_t myc = 0;
#pragma unroll
    for (int s_ = 0; s_ < LOG; ++s_)
        if (lane & (16u >> s_)) myc += (uint32_t)(NC >> (s_ + 1));
    const float myc2 = myc < ncent ? c2[myc] : 0.f;
    for (; d0 < DB; d0 += stride) {
        float acc[kSkinnyDocs][NC];
#pragma unroll
        for (int r = 0; r < kSkinnyDocs; ++r)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
        const float4 *rowp[kSkinnyDocs];
#pragma unroll
        for (int r = 0; r < kSkinnyDocs; ++r) rowp[r] = P + (size_t)min(d0 + r, DB - 1) * kp4;   // clamp: tail rows re-read a valid one
        for (uint32_t col = lane; col < kp4; col += 32) {
            float4 p[kSkinnyDocs];
#pragma unroll
            for (int r = 0; r < kSkinnyDocs; ++r) p[r] = __ldg(rowp[r] + col);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if ((uint32_t)c < ncent) {
                    const float4 q = sC[(size_t)c * kp4 + col];
#pragma unroll
                    for (int r = 0; r < kSkinnyDocs; ++r)
                        acc[r][c] = fmaf(p[r].x, q.x, fmaf(p[r].y, q.y, fmaf(p[r].z, q.z, fmaf(p[r].w, q.w, acc[r][c]))));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kSkinnyDocs; ++r) {
            float v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = acc[r][c];
#pragma unroll
            for (int s_ = 0; s_ < LOG; ++s_) {
                const int w = NC >> (s_ + 1);
                const uint32_t off = 16u >> s_;
                const bool hi = lane & off;
#pragma unroll
                for (int jj = 0; jj < w; ++jj) {
                    const float keep = hi ? v[w + jj] : v[jj], send = hi ? v[jj] : v[w + jj];
                    v[jj] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            float sum = v[0];
#pragma unroll
            for (uint32_t off = 16u >> LOG; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            const uint32_t d = d0 + r;
            const float rd2 = d < DB ? d2[d] : 0.f;
            float best = myc < ncent ? fmaxf(__fadd_rn(__fadd_rn(-2.0f * sum, myc2), rd2), 0.0f) : FLT_MAX;
#pragma unroll
            for (uint32_t off = 16; off >= (32u >> LOG); off >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, off));
            if (lane == 0 && d < DB) min_dist[d] = fminf(min_dist[d], best);
        }
    }
}

__global__ void row_l2_kernel(const float *__restrict__ C, uint32_t rows, uint32_t kp, float *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    float s = 0.f;
    for (uint32_t j = lane; j < kp; j += 32) { const float v = C[(size_t)r * kp + j]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
}

// mode 0: assign[d] = argmin |dist| ; mode 1: min_dist[d] = min(min_dist[d], max(dist,0))
static void distance_pass(Ctx &c, const float *C, const float *c2, uint32_t ncent, int mode, uint32_t *assign,
                          float *min_dist)
{
    const uint32_t DB = (uint32_t)c.DB, kp = (uint32_t)c.kp;
    if (!DB) return;
    // engine: 1 = tcgen05 split-TF32 (dist_tc.cu), 0 = SIMT fp32.  Skinny k-means++ updates (a handful of
    // new centers) are a memory-bound pass over P and stay on the SIMT engine, which reads P once.
    const bool tc = c.opt("dist_kernel", 1) == 1 && dist_tc_supported(c, kp, ncent) &&
                    (mode == 0 || (int)ncent >= c.opt("dist_tc_min_centers", kSkinnyMax + 1));
    const size_t skinny_smem = (size_t)ncent * kp * sizeof(float);
    const bool skinny = mode == 1 && !tc && ncent <= (uint32_t)kSkinnyMax && skinny_smem <= 160 * 1024 &&
                        c.opt("pp_skinny", 1) != 0;
    // logical flops = 2 D_B ncent k
    StatScope s(c, mode == 0 ? (tc ? "dist_tc" : "dist_simt") : (tc ? "pp_dist_tc" : skinny ? "pp_dist_skinny" : "pp_dist_simt"),
                (double)DB * kp * 4.0 + (double)DB * 8.0, 2.0 * (double)DB * ncent * (double)c.k);
    if (skinny) {
        const unsigned per_sm = skinny_smem > 100 * 1024 ? 1u : 2u;
        const unsigned grid = std::min<unsigned>((DB + 8 * kSkinnyDocs - 1) / (8 * kSkinnyDocs), (unsigned)c.num_sms * per_sm * 4);
#define ISLE_SKINNY(NC)                                                                                                                  \
    do {                                                                                                                                 \
        ISLE_CUDA_CHECK(cudaFuncSetAttribute(pp_skinny_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skinny_smem));     \
        pp_skinny_kernel<NC><<<grid, 256, skinny_smem, c.stream>>>(reinterpret_cast<const float4 *>(c.P.p), c.p_l2.p, DB, kp / 4,       \
                                                                   reinterpret_cast<const float4 *>(C), c2, ncent, min_dist);          \
    } while (0)
        if (ncent <= 4) ISLE_SKINNY(4);
        else if (ncent <= 8) ISLE_SKINNY(8);
        else ISLE_SKINNY(16);
#undef ISLE_SKINNY
        ISLE_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        return;
    }
    if (tc) {
        dist_tc_launch(c, C, c2, ncent, mode, assign, min_dist);
        return;
    }
    const unsigned grid = (DB + 63) / 64;
    if (mode == 0)
        dist_simt_kernel<0><<<grid, 256, 0, c.stream>>>(c.P.p, c.p_l2.p, DB, kp, C, c2, ncent, assign, min_dist);
    else
        dist_simt_kernel<1><<<grid, 256, 0, c.stream>>>(c.P.p, c.p_l2.p, DB, kp, C, c2, ncent, assign, min_dist);
    count_launch(c);
}

__global__ void iota_u32_kernel(uint32_t *__restrict__ p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// ------------------------------------------------------------------------------ Lloyd step
// One warp per doc: sums[assign[d], :] += P[d, :], counts[assign[d]] += 1   (:1959-1984)
__global__ void __launch_bounds__(256)
accumulate_kernel(const float4 *__restrict__ P, const uint32_t *__restrict__ assign, uint32_t DB, uint32_t kp4,
                  float4 *__restrict__ sums, uint32_t *__restrict__ counts)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const uint32_t cidx = assign[d];
        for (uint32_t col = lane; col < kp4; col += 32)
            atomicAdd(sums + (size_t)cidx * kp4 + col, P[(size_t)d * kp4 + col]);
        if (lane == 0) atomicAdd(counts + cidx, 1u);
    }
}

// Center sums over documents sorted by cluster: a warp walks a run of kSegDocs consecutive entries
// of the sorted (cluster, doc) list, lane l accumulating float4 columns l, l+32, ... of the rows in
// registers, and flushes with one vector atomic per lane only when the cluster changes or the run ends
// (one flush per run instead of one per document).
static constexpr uint32_t kSegDocs = 64;

__global__ void __launch_bounds__(256)
accumulate_sorted_kernel(const float4 *__restrict__ P, const uint32_t *__restrict__ sorted_cluster,
                         const uint32_t *__restrict__ sorted_doc, uint32_t DB, uint32_t kp4, float4 *__restrict__ sums,
                         uint32_t *__restrict__ counts)
{
    constexpr int kMaxCols = 4;    // float4 columns per lane handled in one sweep (kp <= 512 per sweep)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nruns = (DB + kSegDocs - 1) / kSegDocs;
    uint32_t run = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; run < nruns; run += nw) {
        const uint32_t i0 = run * kSegDocs, i1 = min(DB, i0 + kSegDocs);
        for (uint32_t cbase = 0; cbase < kp4; cbase += 32 * kMaxCols) {
            float4 acc[kMaxCols];
#pragma unroll
            for (int q = 0; q < kMaxCols; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t curc = sorted_cluster[i0], n = 0;
            for (uint32_t i = i0; i <= i1; ++i) {
                const uint32_t cl = i < i1 ? sorted_cluster[i] : 0xFFFFFFFFu;
                if (cl != curc) {
#pragma unroll
                    for (int q = 0; q < kMaxCols; ++q) {
                        const uint32_t col = cbase + q * 32 + lane;
                        if (col < kp4) atomicAdd(sums + (size_t)curc * kp4 + col, acc[q]);
                        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (lane == 0 && cbase == 0) atomicAdd(counts + curc, n);
                    curc = cl;
                    n = 0;
                }
                if (i < i1) {
                    const uint32_t d = sorted_doc[i];
#pragma unroll
                    for (int q = 0; q < kMaxCols; ++q) {
                        const uint32_t col = cbase + q * 32 + lane;
                        if (col < kp4) {
                            const float4 v = __ldg(P + (size_t)d * kp4 + col);
                            acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                        }
                    }
                    ++n;
                }
            }
        }
    }
}

// centers = sums * (1/count) for non-empty clusters, zero otherwise (:1988-1992)
__global__ void finalize_centers_kernel(const float *__restrict__ sums, const uint32_t *__restrict__ counts,
                                        uint32_t k, uint32_t kp, float *__restrict__ C)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)k * kp) return;
    const uint32_t cnt = counts[i / kp];
    const float div = (float)cnt;
    C[i] = cnt ? sums[i] * (1.0f / div) : 0.0f;
}

__global__ void count_diff_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t n,
                                  uint32_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t diff = 0;
    for (; i < n; i += stride) diff += (a[i] != b[i]);
    typedef cub::BlockReduce<uint32_t, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const uint32_t t = BR(tmp).Sum(diff);
    if (threadIdx.x == 0 && t) atomicAdd(out, t);
}

__global__ void __launch_bounds__(256)
objective_kernel(const float *__restrict__ P, const float *__restrict__ C, const uint32_t *__restrict__ assign,
                 uint32_t DB, uint32_t kp, double *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    double s = 0.0;
    for (; d < DB; d += nw) {
        const float *p = P + (size_t)d * kp, *cc = C + (size_t)assign[d] * kp;
        for (uint32_t j = lane; j < kp; j += 32) { const double t = (double)p[j] - (double)cc[j]; s += t * t; }
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double t = BR(tmp).Sum(s);
    if (threadIdx.x == 0) atomicAdd(out, t);
}

struct CentersDev {
    DevBuf<float> C, c2;
    uint32_t k, kp;
    CentersDev(Ctx &c, uint32_t k_, uint32_t kp_, const float *host) : C((size_t)k_ * kp_), c2(k_), k(k_), kp(kp_)
    {
        ISLE_CUDA_CHECK(cudaMemsetAsync(C.p, 0, C.bytes(), c.stream));
        if (host)
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(C.p, (size_t)kp * 4, host, (size_t)k * 4, (size_t)k * 4, k,
                                              cudaMemcpyHostToDevice, c.stream));
    }
    void norms(Ctx &c)
    {
        row_l2_kernel<<<(k * 32 + 255) / 256, 256, 0, c.stream>>>(C.p, k, kp, c2.p);
        count_launch(c);
    }
};

static void require_k(Ctx &c, uint64_t k)
{
    ISLE_REQUIRE(c.have_U && c.have_B, ISLE_ERR_ARG, "k-means: needs B and U (run block_ks or set_U)");
    // the reference's projected k-means only supports #centers == k
    ISLE_REQUIRE(k == c.k, ISLE_ERR_ARG, "k-means: number of centers must equal the eigensolver's k");
    project(c);
}

void assign_projected(Ctx &c, uint64_t k, const float *centers, uint32_t *assign_out)
{
    require_k(c, k);
    CentersDev cd(c, (uint32_t)k, (uint32_t)c.kp, centers);
    cd.norms(c);
    DevBuf<uint32_t> a(std::max<uint64_t>(c.DB, 1));
    distance_pass(c, cd.C.p, cd.c2.p, (uint32_t)k, 0, a.p, nullptr);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(assign_out, a.p, (size_t)c.DB * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// update_min_distsq_to_projected_centers (src/sparseMatrix.cpp:2075-2130) over all documents for a batch of
// `ncent` centers (each of length k, contiguous): min_dist[d] = min(min_dist[d], max(dist(d, c), 0)).
// The engine is the one k-means++ would pick for a batch of this size (options dist_kernel, dist_tc_min_centers,
// pp_skinny), so every k-means++ refresh kernel can be checked on its own.
void update_min_dist(Ctx &c, uint64_t ncent, const float *centers, float *min_dist_inout)
{
    ISLE_REQUIRE(c.have_U && c.have_B, ISLE_ERR_ARG, "update_min_dist: needs B and U (run block_ks or set_U)");
    ISLE_REQUIRE(ncent >= 1 && centers && min_dist_inout, ISLE_ERR_ARG, "update_min_dist: bad arguments");
    project(c);
    const uint32_t k = (uint32_t)c.k, kp = (uint32_t)c.kp, DB = (uint32_t)c.DB;
    DevBuf<float> cc((size_t)ncent * kp), c2(ncent), md(std::max<uint32_t>(DB, 1));
    ISLE_CUDA_CHECK(cudaMemsetAsync(cc.p, 0, cc.bytes(), c.stream));
    ISLE_CUDA_CHECK(cudaMemcpy2DAsync(cc.p, (size_t)kp * 4, centers, (size_t)k * 4, (size_t)k * 4, ncent, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(md.p, min_dist_inout, (size_t)DB * 4, cudaMemcpyHostToDevice, c.stream));
    row_l2_kernel<<<(unsigned)((ncent * 32 + 255) / 256), 256, 0, c.stream>>>(cc.p, (uint32_t)ncent, kp, c2.p);
    count_launch(c);
    distance_pass(c, cc.p, c2.p, (uint32_t)ncent, 1, nullptr, md.p);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(min_dist_inout, md.p, (size_t)DB * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

void lloyd_projected(Ctx &c, uint64_t k64, float *centers_inout, int max_reps, uint32_t *assign_out,
                     double *objective_out, int *iters_out)
{
    require_k(c, k64);
    const uint32_t k = (uint32_t)k64, kp = (uint32_t)c.kp, DB = (uint32_t)c.DB;
    CentersDev cd(c, k, kp, centers_inout);
    DevBuf<uint32_t> a0(std::max<uint32_t>(DB, 1)), a1(std::max<uint32_t>(DB, 1)), counts(k), ndiff(1);
    DevBuf<float> sums((size_t)k * kp);
    // documents sorted by cluster for the center sums (one radix pass over log2(k) bits per iteration)
    const bool sorted_accum = c.opt("lloyd_sorted_accum", 1) != 0 && DB > 0;
    DevBuf<uint32_t> doc_ids, skey, sdoc;
    DevBuf<uint8_t> sort_tmp;
    size_t sort_bytes = 0;
    int key_bits = 1;
    while ((1u << key_bits) < k) ++key_bits;
    if (sorted_accum) {
        doc_ids.alloc(DB); skey.alloc(DB); sdoc.alloc(DB);
        iota_u32_kernel<<<grid_for(DB, 256), 256, 0, c.stream>>>(doc_ids.p, DB);
        count_launch(c);
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, a0.p, skey.p, doc_ids.p, sdoc.p, (int)DB, 0, key_bits, c.stream);
        sort_tmp.alloc(sort_bytes);
    }
    uint32_t *cur = a0.p, *prev = a1.p;
    int iters = 0;
    for (int it = 0; it < max_reps; ++it) {
        StatScope s(c, "lloyd_iter");
        cd.norms(c);
        distance_pass(c, cd.C.p, cd.c2.p, k, 0, cur, nullptr);
        ISLE_CUDA_CHECK(cudaMemsetAsync(sums.p, 0, sums.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(counts.p, 0, counts.bytes(), c.stream));
        {
            StatScope s2(c, "lloyd_accum", (double)DB * kp * 4.0 + (double)DB * 4.0);
            if (sorted_accum) {
                ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp.p, sort_bytes, cur, skey.p, doc_ids.p, sdoc.p, (int)DB, 0,
                                                                key_bits, c.stream));
                const uint32_t nruns = (DB + kSegDocs - 1) / kSegDocs;
                accumulate_sorted_kernel<<<grid_for((size_t)nruns * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(
                    reinterpret_cast<const float4 *>(c.P.p), skey.p, sdoc.p, DB, kp / 4, reinterpret_cast<float4 *>(sums.p), counts.p);
                count_launch(c, 2);
            } else if (DB) {
                accumulate_kernel<<<grid_for((size_t)DB * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(
                    reinterpret_cast<const float4 *>(c.P.p), cur, DB, kp / 4, reinterpret_cast<float4 *>(sums.p), counts.p);
                count_launch(c);
            }
        }
        if (c.world > 1) {
            allreduce_sum_f32(c, sums.p, sums.n);
            allreduce_sum_u32(c, counts.p, counts.n);
        }
        finalize_centers_kernel<<<(unsigned)(((size_t)k * kp + 255) / 256), 256, 0, c.stream>>>(sums.p, counts.p, k, kp, cd.C.p);
        count_launch(c);
        ++iters;
        // converged when the partition equals the previous one (:2044-2064)
        uint32_t nd = 1;
        if (it > 0) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(ndiff.p, 0, 4, c.stream));
            if (DB) {
                count_diff_kernel<<<grid_for(DB, 256, c.num_sms * 4), 256, 0, c.stream>>>(cur, prev, DB, ndiff.p);
                count_launch(c);
            }
            if (c.world > 1) allreduce_sum_u32(c, ndiff.p, 1);
            ISLE_CUDA_CHECK(cudaMemcpyAsync(&nd, ndiff.p, 4, cudaMemcpyDeviceToHost, c.stream));
            ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        }
        std::swap(cur, prev);   // prev now holds this iteration's partition
        if (nd == 0) break;
    }
    const uint32_t *final_assign = prev;
    if (objective_out) {
        DevBuf<double> obj(1);
        ISLE_CUDA_CHECK(cudaMemsetAsync(obj.p, 0, 8, c.stream));
        if (DB) {
            objective_kernel<<<grid_for((size_t)DB * 32, 256, c.num_sms * 8), 256, 0, c.stream>>>(c.P.p, cd.C.p, final_assign, DB, kp, obj.p);
            count_launch(c);
        }
        if (c.world > 1) allreduce_sum_f64(c, obj.p, 1);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(objective_out, obj.p, 8, cudaMemcpyDeviceToHost, c.stream));
    }
    if (assign_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(assign_out, final_assign, (size_t)DB * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpy2DAsync(centers_inout, (size_t)k * 4, cd.C.p, (size_t)kp * 4, (size_t)k * 4, k,
                                      cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (iters_out) *iters_out = iters;
}

// ------------------------------------------------------------------------------- k-means++
struct F2D {
    __host__ __device__ double operator()(float x) const { return (double)x; }
};

// D^2 sampling draw (src/sparseMatrix.cpp:2181-2199) on the device: thread i turns the uniform u[i] into the threshold
// total * u[i] (total = sum of the ranks' partial totals, added in rank order), finds the rank whose share of the prefix sums
// holds it (never an empty shard) and, on that rank, the first index with cumul[i] > t == upper_bound(dist_cumul, t) of the
// reference (:2187); the result is the GLOBAL column id + 1 (0 = not this rank's draw).  One host read per round (the ids).
struct PpDraws { double u[64]; };
__global__ void pick_draws_kernel(const double *__restrict__ cumul, uint32_t n, const double *__restrict__ tot_all, int world, int rank,
                                  PpDraws dr, int nt, unsigned long long offset, unsigned long long *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double prefix[17];
    prefix[0] = 0.0;
    for (int r = 0; r < world; ++r) prefix[r + 1] = prefix[r] + tot_all[r];
    const double t = prefix[world] * dr.u[i];
    int r = 0;
    while (r + 1 < world && t >= prefix[r + 1]) ++r;
    while (r > 0 && tot_all[r] <= 0.0) --r;
    if (r != rank || n == 0) { out[i] = 0ull; return; }
    const double tv = fmax(0.0, t - prefix[r]);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cumul[mid] > tv) hi = mid; else lo = mid + 1;
    }
    out[i] = offset + (lo < n ? lo : n - 1) + 1ull;
}

__global__ void fill_kernel(float *__restrict__ p, size_t n, float v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// Document-sharded D^2 sampling: min_dist and its prefix sums are local to the rank
// that owns the documents; per round the ranks exchange their partial totals (allgather of one
// double), every rank draws the same thresholds from the same generator, the owner of each
// threshold locates the document in its local prefix sums, and the chosen ids and then the
// chosen rows of P travel as zero-padded allreduces.  With world == 1 the collectives vanish
// and this is kmeanspp_on_projected_space (src/sparseMatrix.cpp:2133-2209) on one device.
void kmeanspp(Ctx &c, uint64_t k64, uint64_t seed, uint64_t *seeds_out, float *centers_out, float *residual_out)
{
    require_k(c, k64);
    const uint32_t k = (uint32_t)k64, kp = (uint32_t)c.kp, DB = (uint32_t)c.DB;
    const uint64_t DBtot = c.db_total;
    ISLE_REQUIRE(DBtot >= k, ISLE_ERR_ARG, "kmeanspp: fewer documents than centers");   // :2137
    constexpr int kMaxDraw = 64;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    DevBuf<float> min_dist(std::max<uint32_t>(DB, 1)), cc((size_t)k * kp), c2(k), xch((size_t)kMaxDraw * kp);
    DevBuf<double> cumul(std::max<uint32_t>(DB, 1)), tot_mine(1), tot_all((size_t)c.world);
    DevBuf<uint32_t> cand(kMaxDraw);
    DevBuf<unsigned long long> gids(kMaxDraw);
    if (DB) {
        fill_kernel<<<grid_for(DB, 256), 256, 0, c.stream>>>(min_dist.p, DB, FLT_MAX);   // :2149
        count_launch(c);
    }
    size_t tb = 0;
    auto it = cub::TransformInputIterator<double, F2D, float *>(min_dist.p, F2D());
    cub::DeviceScan::InclusiveSum(nullptr, tb, it, cumul.p, (int)std::max<uint32_t>(DB, 1), c.stream);
    DevBuf<uint8_t> tmp(tb);

    std::vector<uint64_t> centers;   // global column ids of B
    auto owner_of = [&](uint64_t g) {
        uint64_t off = 0;
        for (int r = 0; r < c.world; ++r) { if (g < off + c.db_all[r]) return r; off += c.db_all[r]; }
        return c.world - 1;
    };
    // rows P[g] of the given global ids -> cc[centers.size() ...), on every rank
    auto add_centers = [&](const std::vector<uint64_t> &g) {
        if (g.empty()) return;
        const size_t nf = g.size() * (size_t)kp;
        if (c.world > 1) ISLE_CUDA_CHECK(cudaMemsetAsync(xch.p, 0, nf * 4, c.stream));
        for (size_t i = 0; i < g.size(); ++i)
            if (owner_of(g[i]) == c.rank)
                ISLE_CUDA_CHECK(cudaMemcpyAsync(xch.p + i * kp, c.P.p + (size_t)(g[i] - c.db_offset) * kp, (size_t)kp * 4,
                                                cudaMemcpyDeviceToDevice, c.stream));
        if (c.world > 1) allreduce_sum_f32(c, xch.p, nf);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(cc.p + centers.size() * (size_t)kp, xch.p, nf * 4, cudaMemcpyDeviceToDevice, c.stream));
        centers.insert(centers.end(), g.begin(), g.end());
    };
    add_centers({(uint64_t)(rng() % DBtot)});   // :2150 picks an arbitrary first doc
    int new_added = 1;
    double total = 0.0;
    std::vector<double> totals((size_t)c.world);
    while (centers.size() < k) {
        StatScope s(c, "pp_round");
        const uint32_t first = (uint32_t)centers.size() - new_added;
        row_l2_kernel<<<(new_added * 32 + 255) / 256, 256, 0, c.stream>>>(cc.p + (size_t)first * kp, new_added, kp, c2.p + first);
        count_launch(c);
        distance_pass(c, cc.p + (size_t)first * kp, c2.p + first, (uint32_t)new_added, 1, nullptr, min_dist.p);
        if (DB) {
            ISLE_CUDA_CHECK(cub::DeviceScan::InclusiveSum(tmp.p, tb, it, cumul.p, (int)DB, c.stream));
            count_launch(c);
            ISLE_CUDA_CHECK(cudaMemcpyAsync(tot_mine.p, cumul.p + (DB - 1), 8, cudaMemcpyDeviceToDevice, c.stream));
        } else {
            ISLE_CUDA_CHECK(cudaMemsetAsync(tot_mine.p, 0, 8, c.stream));
        }
        allgather_f64(c, tot_mine.p, tot_all.p);
        // draw 1 + sqrt(max(s-5,0)) candidates (:2181-2199); duplicates are skipped.  The uniforms come from the host
        // generator (every rank draws the same), the thresholds total * u and their owners are worked out on the device, so
        // the only host read of the round is the picked ids (with the totals riding along)
        const int s_now = (int)centers.size();
        PpDraws dr;
        int ndraw = 0;
        for (int d = 0; d < 1 + std::sqrt((double)(s_now - 5 > 0 ? s_now - 5 : 0)) && ndraw < kMaxDraw; ++d) dr.u[ndraw++] = uni(rng);
        pick_draws_kernel<<<1, kMaxDraw, 0, c.stream>>>(cumul.p, DB, tot_all.p, c.world, c.rank, dr, ndraw, c.db_offset, gids.p);
        count_launch(c);
        if (c.world > 1) allreduce_sum_u64(c, gids.p, (size_t)ndraw);
        std::vector<unsigned long long> hg(ndraw);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(hg.data(), gids.p, (size_t)ndraw * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaMemcpyAsync(totals.data(), tot_all.p, totals.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        total = 0.0;
        for (int r = 0; r < c.world; ++r) total += totals[r];
        std::vector<uint64_t> fresh;
        for (int d = 0; d < ndraw && centers.size() + fresh.size() < k; ++d) {
            if (hg[d] == 0) continue;                       // no owner (total == 0)
            const uint64_t g = hg[d] - 1;
            if (std::find(centers.begin(), centers.end(), g) == centers.end() &&
                std::find(fresh.begin(), fresh.end(), g) == fresh.end())
                fresh.push_back(g);
        }
        if (fresh.empty() && total <= 0.0) {
            // degenerate corpus (all remaining docs coincide with a center): fall back to unused docs
            for (uint64_t g = 0; g < DBtot && centers.size() + fresh.size() < k && fresh.size() < (size_t)kMaxDraw; ++g)
                if (std::find(centers.begin(), centers.end(), g) == centers.end()) fresh.push_back(g);
        }
        add_centers(fresh);
        new_added = (int)fresh.size();
    }
    if (seeds_out) for (uint32_t i = 0; i < k; ++i) seeds_out[i] = centers[i];
    if (centers_out)   // kmeans_init_on_projected_space copies U^T doc for each seed (:2232-2234)
        ISLE_CUDA_CHECK(cudaMemcpy2DAsync(centers_out, (size_t)k * 4, cc.p, (size_t)kp * 4, (size_t)k * 4, k,
                                          cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (residual_out) *residual_out = (float)total;
}

// --------------------------------------------------------------------------------- lift
void lift_centers(Ctx &c, uint64_t ncols, const float *in, uint64_t ld_in, float *out)
{
    ISLE_REQUIRE(c.have_U, ISLE_ERR_ARG, "lift_centers: needs U");
    ISLE_REQUIRE(ld_in >= c.k, ISLE_ERR_ARG, "lift_centers: ld_in < k");   // :1445
    const int V = (int)c.V, k = (int)c.k, nc = (int)ncols;
    DevBuf<float> din((size_t)ld_in * ncols);
    // the lifted centers stay on the device: they are the input of run_lloyds on the full-dimensional B (lloyd_full.cu)
    c.lifted.alloc((size_t)V * ncols);
    c.lifted_cols = ncols;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(din.p, in, din.bytes(), cudaMemcpyHostToDevice, c.stream));
    if ((size_t)ld_in * ncols > din.n) throw Error(ISLE_ERR_ARG, "lift_centers: bad ld_in");
    gemm_3xtf32(c, V, nc, k, c.U.p, V, din.p, (int)ld_in, c.lifted.p, V);                     // tensor cores above ~20 GFLOP
    if (out) ISLE_CUDA_CHECK(cudaMemcpyAsync(out, c.lifted.p, c.lifted.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

}  // namespace isle
