// lloyd_full.cu -- SURVEY 8(f) row 1: Lloyd's k-means on the full-dimensional thresholded matrix B.
//
// Replaces FPSparseMatrix::run_lloyds and its helpers (reference src/sparseMatrix.cpp:1494-1746:
// distsq_docs_to_centers, closest_centers, compute_centers_l2sq, lloyds_iter, compute_docs_l2sq,
// run_lloyds), the call that follows the spectral core in ISLETrainer::train()
// (src/trainer.cpp:559-571) and produces `closest_docs`.
//
// Centers are dense, k x V with center c contiguous (`centers + c * vocab_size`).  Per iteration:
//   c2[c]   = ||center_c||^2                                                  (:1571-1581)
//   Cs      = centers transposed to V x kp row-major and scaled by sqrt_zeta per row, so that the
//             sparse x dense product of :1533-1537 becomes a pattern gather: B's nonzeros of row w
//             all equal sqrt_zeta[w]
//   assign  warp per document: s[c] = sum_{w in d} Cs[w, c]; dist = ((-2 s) + c2[c]) + d2[d] in the
//           order of :1533-1546; argmin |dist|, first index on ties (cblas_isamin, :1567)
//   update  center_c[w] = sqrt_zeta[w] * #{d in cluster c : w in d} / |cluster c|: member counts are
//           integer atomics (order independent -> run-to-run and rank-count deterministic), an empty
//           cluster's center stays zero (:1626, :1655-1661)
// Stop when the partition repeats (:1718-1737), at most max_reps iterations.
// Document-sharded: counts and cluster sizes are summed over ranks (u32 allreduce); everything
// else is local to the documents a rank holds.
#include <algorithm>

#include "common.cuh"

namespace isle {

namespace {

// Cs[w, j] = scale[w] * C[j V + w], row stride kp, zero padded
__global__ void __launch_bounds__(256)
transpose_scale_centers_kernel(const float *__restrict__ C, uint32_t V, uint32_t k, uint32_t kp,
                               const float *__restrict__ scale, float *__restrict__ Cs)
{
    __shared__ float tile[32][33];
    const uint32_t w0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint32_t r = ty; r < 32; r += 8) {
        const uint32_t j = j0 + r, w = w0 + tx;
        tile[r][tx] = (j < k && w < V) ? C[(size_t)j * V + w] : 0.0f;
    }
    __syncthreads();
    for (uint32_t r = ty; r < 32; r += 8) {
        const uint32_t w = w0 + r, j = j0 + tx;
        if (w < V && j < kp) Cs[(size_t)w * kp + j] = tile[tx][r] * scale[w];
    }
}

// c2[c] = ||C_c||^2, one CTA per center, fixed-order tree (deterministic)
__global__ void __launch_bounds__(256)
center_l2_kernel(const float *__restrict__ C, uint32_t V, float *__restrict__ c2)
{
    __shared__ float red[8];
    const float *row = C + (size_t)blockIdx.x * V;
    float s = 0.f;
    for (uint32_t w = threadIdx.x; w < V; w += 256) s = fmaf(row[w], row[w], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        c2[blockIdx.x] = t;
    }
}

// d2[d] = sum_{w in d} sqrt_zeta[w]^2   (compute_docs_l2sq, :1670-1677)
__global__ void __launch_bounds__(256)
doc_l2_full_kernel(const uint32_t *__restrict__ b_row, const int64_t *__restrict__ b_off, uint32_t DB,
                   const float *__restrict__ sz, float *__restrict__ d2)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = b_off[d], e = b_off[d + 1];
        float s = 0.f;
        for (int64_t p = b + lane; p < e; p += 32) {
            const float v = __ldg(sz + b_row[p]);
            s = fmaf(v, v, s);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) d2[d] = s;
    }
}

// One warp per document; lane l owns columns 4 (c0 + l) .. +3 of each 128-column chunk.
__global__ void __launch_bounds__(256)
assign_full_kernel(const uint32_t *__restrict__ b_row, const int64_t *__restrict__ b_off, uint32_t DB,
                   const float4 *__restrict__ Cs, uint32_t kp4, uint32_t k, const float *__restrict__ c2,
                   const float *__restrict__ d2, uint32_t *__restrict__ assign)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = b_off[d], e = b_off[d + 1];
        const float dd = d2[d];
        float best = INFINITY;
        uint32_t best_c = 0xFFFFFFFFu;
        for (uint32_t c0 = 0; c0 < kp4; c0 += 32) {
            const uint32_t col = c0 + lane;
            if (col < kp4) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int64_t p = b;
                for (; p + 4 <= e; p += 4) {
                    const uint32_t w0 = b_row[p], w1 = b_row[p + 1], w2 = b_row[p + 2], w3 = b_row[p + 3];
                    const float4 v0 = __ldg(Cs + (size_t)w0 * kp4 + col);
                    const float4 v1 = __ldg(Cs + (size_t)w1 * kp4 + col);
                    const float4 v2 = __ldg(Cs + (size_t)w2 * kp4 + col);
                    const float4 v3 = __ldg(Cs + (size_t)w3 * kp4 + col);
                    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                    acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                    acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
                }
                for (; p < e; ++p) {
                    const float4 v = __ldg(Cs + (size_t)b_row[p] * kp4 + col);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const float s[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t c = 4 * col + q;
                    if (c < k) {
                        // ((alpha * s) + c2) + d2 with alpha = -2: the reference's evaluation order; no FMA contraction
                        const float v = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, s[q]), c2[c]), dd));
                        if (v < best) { best = v; best_c = c; }      // ascending c per lane: strict < keeps the first
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const uint32_t oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
        }
        if (lane == 0) assign[d] = best_c == 0xFFFFFFFFu ? 0u : best_c;
    }
}

// cnt[c V + w] += 1 for every nonzero (w, d) with assign[d] = c; sizes[c] += 1 per document
__global__ void __launch_bounds__(256)
count_members_kernel(const uint32_t *__restrict__ b_row, const int64_t *__restrict__ b_off, uint32_t DB,
                     const uint32_t *__restrict__ assign, uint32_t V, uint32_t *__restrict__ cnt,
                     uint32_t *__restrict__ sizes)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < DB; d += nw) {
        const int64_t b = b_off[d], e = b_off[d + 1];
        const uint32_t c = assign[d];
        uint32_t *row = cnt + (size_t)c * V;
        for (int64_t p = b + lane; p < e; p += 32) atomicAdd(row + b_row[p], 1u);
        if (lane == 0) atomicAdd(sizes + c, 1u);
    }
}

// center_c[w] = (sqrt_zeta[w] * cnt) / size  (sum of cnt equal values, then the division of :1655-1661)
__global__ void __launch_bounds__(256)
finalize_full_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ sizes, const float *__restrict__ sz,
                     uint32_t V, uint32_t k, float *__restrict__ C)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)k * V) return;
    const uint32_t c = (uint32_t)(i / V), w = (uint32_t)(i % V);
    const uint32_t n = sizes[c], m = cnt[i];
    C[i] = (n && m) ? __fdiv_rn(__fmul_rn(sz[w], (float)m), (float)n) : 0.0f;
}

__global__ void __launch_bounds__(256)
count_diff_full_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t n, uint32_t *__restrict__ out)
{
    uint32_t local = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) local += a[i] != b[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// obj += sum_d ( d2[d] - 2 <B_d, c_a> + c2[a] ), fp64
__global__ void __launch_bounds__(256)
objective_full_kernel(const uint32_t *__restrict__ b_row, const int64_t *__restrict__ b_off, uint32_t DB,
                      const float *__restrict__ sz, const float *__restrict__ C, uint32_t V,
                      const uint32_t *__restrict__ assign, const float *__restrict__ c2, double *__restrict__ obj)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (; d < DB; d += nw) {
        const int64_t b = b_off[d], e = b_off[d + 1];
        const uint32_t a = assign[d];
        const float *row = C + (size_t)a * V;
        double s = 0.0;
        for (int64_t p = b + lane; p < e; p += 32) {
            const uint32_t w = b_row[p];
            const double v = (double)__ldg(sz + w);
            s += v * v - 2.0 * v * (double)__ldg(row + w);
        }
        if (lane == 0) s += (double)c2[a];
        acc += s;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(obj, acc);
}

}  // namespace

void lloyd_full(Ctx &c, uint64_t k64, float *centers_inout, int max_reps, uint32_t *assign_out,
                double *objective_out, int *iters_out)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "lloyd_full: build_B first");
    ISLE_REQUIRE(k64 >= 1 && k64 <= 65536, ISLE_ERR_ARG, "lloyd_full: bad number of centers");
    const uint32_t k = (uint32_t)k64, V = (uint32_t)c.V, DB = (uint32_t)c.DB;
    const uint32_t kp = (k + 31) / 32 * 32;
    ISLE_REQUIRE(centers_inout || (c.lifted.p && c.lifted_cols == k64), ISLE_ERR_ARG,
                 "lloyd_full: no host centers given and no lifted centers of that width on the device");
    DevBuf<float> Cown;
    float *C = nullptr;
    if (centers_inout) {
        Cown.alloc((size_t)k * V);
        C = Cown.p;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(C, centers_inout, (size_t)k * V * 4, cudaMemcpyHostToDevice, c.stream));
    } else {
        C = c.lifted.p;      // updated in place on the device
    }
    DevBuf<float> Cs((size_t)V * kp), c2(k), d2(std::max<uint32_t>(DB, 1));
    DevBuf<uint32_t> cnt((size_t)k * V), sizes(k), a0(std::max<uint32_t>(DB, 1)), a1(std::max<uint32_t>(DB, 1)), ndiff(1);
    const unsigned wgrid = grid_for((size_t)std::max<uint32_t>(DB, 1) * 32, 256, c.num_sms * 8);
    if (DB) {
        doc_l2_full_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_row.p, c.b_off.p, DB, c.sqrt_zeta.p, d2.p);
        count_launch(c);
    }
    uint32_t *cur = a0.p, *prev = a1.p;
    int iters = 0;
    for (int it = 0; it < max_reps; ++it) {
        StatScope s(c, "lloyd_full_iter");
        center_l2_kernel<<<k, 256, 0, c.stream>>>(C, V, c2.p);
        transpose_scale_centers_kernel<<<dim3((V + 31) / 32, kp / 32), 256, 0, c.stream>>>(C, V, k, kp, c.sqrt_zeta.p, Cs.p);
        count_launch(c, 2);
        if (DB) {
            // the k-wide gather: flops = 2 nnz k; compulsory bytes = nnz 4 + V k 4 + D_B 12
            StatScope s2(c, "lloyd_full_assign", (double)c.nnzB * 4.0 + (double)V * k * 4.0 + (double)DB * 12.0, 2.0 * c.nnzB * k);
            assign_full_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_row.p, c.b_off.p, DB, reinterpret_cast<const float4 *>(Cs.p),
                                                            kp / 4, k, c2.p, d2.p, cur);
            count_launch(c);
        }
        ISLE_CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(sizes.p, 0, sizes.bytes(), c.stream));
        {
            StatScope s3(c, "lloyd_full_update", (double)c.nnzB * 4.0 + 2.0 * (double)V * k * 4.0 + (double)DB * 12.0);
            if (DB) {
                count_members_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_row.p, c.b_off.p, DB, cur, V, cnt.p, sizes.p);
                count_launch(c);
            }
            if (c.world > 1) {
                allreduce_sum_u32(c, cnt.p, cnt.n);
                allreduce_sum_u32(c, sizes.p, sizes.n);
            }
            finalize_full_kernel<<<(unsigned)(((size_t)k * V + 255) / 256), 256, 0, c.stream>>>(cnt.p, sizes.p, c.sqrt_zeta.p, V, k, C);
            count_launch(c);
        }
        ++iters;
        uint32_t nd = 1;
        if (it > 0) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(ndiff.p, 0, 4, c.stream));
            if (DB) {
                count_diff_full_kernel<<<grid_for(DB, 256, c.num_sms * 4), 256, 0, c.stream>>>(cur, prev, DB, ndiff.p);
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
        center_l2_kernel<<<k, 256, 0, c.stream>>>(C, V, c2.p);
        count_launch(c);
        if (DB) {
            objective_full_kernel<<<wgrid, 256, 0, c.stream>>>(c.b_row.p, c.b_off.p, DB, c.sqrt_zeta.p, C, V, final_assign, c2.p, obj.p);
            count_launch(c);
        }
        if (c.world > 1) allreduce_sum_f64(c, obj.p, 1);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(objective_out, obj.p, 8, cudaMemcpyDeviceToHost, c.stream));
    }
    if (assign_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(assign_out, final_assign, (size_t)DB * 4, cudaMemcpyDeviceToHost, c.stream));
    if (centers_inout)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(centers_inout, C, (size_t)k * V * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (iters_out) *iters_out = iters;
}

}  // namespace isle
