// dist_tc.cu -- kernel family (3), tensor-core engine: the docs x centers contraction of the
// projected k-means on tcgen05 (5th-gen tensor cores, TMEM accumulators) in split-TF32.
//
// Reference statement of the operation: distsq_alldocs_to_centers (src/denseMatrix.cpp:504-530,
// the "explicit projection" form of src/sparseMatrix.cpp:1794-1849):
//     dist[d,c] = ((-2 P_d . C_c) + ||C_c||^2) + ||P_d||^2
// followed by cblas_isamin per document (src/sparseMatrix.cpp:1868-1870: argmin |x|, first index
// on ties) or, for k-means++, min(min_dist, max(dist, 0)) (:2112-2126).
//
// 3xTF32: P = P_hi + P_lo and C = C_hi + C_lo; S = P_hi C_hi^T + P_hi C_lo^T + P_lo C_hi^T accumulated in fp32 in
// TMEM (the lo*lo term, ~2^-22 relative, is dropped).  C is split once per pass into hi = tf32(x) (round to nearest) and
// lo = tf32(x - hi).  P has NO hi copy: the tensor core reads a tf32 operand from the upper 19 bits of its 32-bit container
// and ignores the rest, so the TMA tiles of P itself ARE P_hi = trunc_tf32(P); only P_lo = tf32(P - trunc_tf32(P)) (the
// dropped 13 bits, rounded to tf32's 11: error 2^-22 of |P|, the same as the round-to-nearest split) is stored, once per
// projection.  P is therefore held twice, not three times (2 x 8.4 GB per c3 shard instead of 3 x).
//
// Kernel shape (one persistent CTA per SM, 6 warps):
//   warp 0 / lane 0   TMA producer: cp.async.bulk.tensor tiles of P_hi, P_lo (128 docs x 32 k) and
//                     C_hi, C_lo (BN centers x 32 k), 128-byte swizzle, into a ring of stages
//   warp 1 / lane 0   MMA issuer: 12 tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per stage into
//                     one of two TMEM accumulators; tcgen05.commit frees the stage / publishes
//                     the accumulator
//   warps 2-5         epilogue: tcgen05.ld one accumulator row per thread, fused
//                     -2 s + ||c||^2 + ||d||^2, |.|-argmin (or clamped min) carried across center
//                     tiles in registers; the distance matrix is never written
// The D_B x k distance matrix would be 65.6 GB at the PubMed shape; only assign[] / min_dist[]
// leave the SM.
#include <cuda.h>

#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace isle {

namespace {

using namespace tcptx;

constexpr int BM = 128;            // documents per tile (UMMA M, cta_group::1)
constexpr int BK = 32;             // tf32 elements per 128-byte swizzled row
constexpr int kMaxBN = 256;        // centers per tile (UMMA N)
constexpr int kTmemCols = 512;     // two accumulators of up to 256 fp32 columns
constexpr int kThreads = 192;
constexpr uint32_t kSmemBudget = 220 * 1024;

// ---------------------------------------------------------------------------------- PTX

// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; spins < (1u << 24); ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}


struct Params {
    const float *d2;        // ||P_d||^2
    const float *c2;        // ||C_c||^2
    uint32_t *assign;       // mode 0
    float *min_dist;        // mode 1 (read-modify-write)
    uint32_t DB, ncent, kp, BN, stages, mode;
};

__global__ void __launch_bounds__(kThreads, 1)
dist_tc_kernel(const __grid_constant__ CUtensorMap map_p_hi, const __grid_constant__ CUtensorMap map_p_lo,
               const __grid_constant__ CUtensorMap map_c_hi, const __grid_constant__ CUtensorMap map_c_lo, const Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = BM * 128, b_bytes = p.BN * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t *ctrl = smem + (size_t)p.stages * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);            // [stages]
    uint64_t *empty = full + p.stages;                               // [stages]
    uint64_t *tfull = empty + p.stages;                              // [2]
    uint64_t *tempty = tfull + 2;                                    // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    float *c2s = reinterpret_cast<float *>(tmem_slot + 4);           // [ncent]

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t num_m_tiles = (p.DB + BM - 1) / BM;
    const uint32_t num_n_tiles = (p.ncent + p.BN - 1) / p.BN;
    const uint32_t num_k_blocks = p.kp / BK;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_p_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_p_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (uint32_t i = threadIdx.x; i < p.ncent; i += kThreads) c2s[i] = p.c2[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            uint32_t stage = 0, phase = 0;
            for (uint32_t mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
                for (uint32_t nt = 0; nt < num_n_tiles; ++nt) {
                    for (uint32_t kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t *st = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full[stage], stage_bytes);
                        tma_load_2d(st, &map_p_hi, &full[stage], (int)(kb * BK), (int)(mt * BM));
                        tma_load_2d(st + a_bytes, &map_p_lo, &full[stage], (int)(kb * BK), (int)(mt * BM));
                        tma_load_2d(st + 2 * a_bytes, &map_c_hi, &full[stage], (int)(kb * BK), (int)(nt * p.BN));
                        tma_load_2d(st + 2 * a_bytes + b_bytes, &map_c_lo, &full[stage], (int)(kb * BK), (int)(nt * p.BN));
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
                for (uint32_t nt = 0; nt < num_n_tiles; ++nt) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kMaxBN;
                    for (uint32_t kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + a_bytes);
                        const uint64_t b_hi = umma_desc(sa + 2 * a_bytes), b_lo = umma_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
                        for (uint32_t ks = 0; ks < BK / 8; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // 8 tf32 = 32 bytes along K
                            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                        umma_commit(&empty[stage]);          // stage reusable once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);                // accumulator complete
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: thread <-> one document row (TMEM lane), warps 2..5 own lane quarters 2,3,0,1
        const uint32_t quarter = warp & 3;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
            const uint32_t row = mt * BM + quarter * 32 + lane;
            const float rd2 = row < p.DB ? p.d2[row] : 0.f;
            float best = FLT_MAX;
            uint32_t besti = 0;
            for (uint32_t nt = 0; nt < num_n_tiles; ++nt) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kMaxBN;
                const uint32_t n0 = nt * p.BN;
                for (uint32_t cb = 0; cb < p.BN && n0 + cb < p.ncent; cb += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + cb, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t cidx = n0 + cb + i;
                        if (cidx < p.ncent) {
                            float v = __fadd_rn(__fadd_rn(-2.0f * __uint_as_float(r[i]), c2s[cidx]), rd2);
                            v = p.mode == 0 ? fabsf(v) : fmaxf(v, 0.0f);
                            if (v < best) { best = v; besti = cidx; }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < p.DB) {
                if (p.mode == 0) p.assign[row] = besti;
                else p.min_dist[row] = fminf(p.min_dist[row], best);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// hi = tf32(x) (round to nearest, ties away), lo = tf32(x - hi); x - hi is exact in fp32.
__global__ void split_tf32_kernel(const float *__restrict__ x, size_t n, float *__restrict__ hi, float *__restrict__ lo)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = x[i];
        uint32_t h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float r = v - __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}


// rows x kp fp32 row-major, box = 32 k x box_rows, 128-byte swizzle, zero fill out of bounds
CUtensorMap make_map(const float *base, uint64_t rows, uint64_t kp, uint32_t box_rows)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {kp, rows};
    const cuuint64_t gstride[1] = {kp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ISLE_REQUIRE(r == CUDA_SUCCESS, ISLE_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// lo = tf32(x - trunc_tf32(x)): the part of x the tensor core does not see when it reads x itself as a tf32 operand
__global__ void split_lo_trunc_kernel(const float *__restrict__ x, size_t n, float *__restrict__ lo)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = x[i];
        const float r = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);     // exact
        uint32_t l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
        lo[i] = __uint_as_float(l);
    }
}

}  // namespace

void split_lo_trunc(Ctx &c, const float *x, size_t n, float *lo)
{
    if (!n) return;
    split_lo_trunc_kernel<<<grid_for(n, 256, c.num_sms * 8), 256, 0, c.stream>>>(x, n, lo);
    count_launch(c);
}

void split_tf32(Ctx &c, const float *x, size_t n, float *hi, float *lo)
{
    if (!n) return;
    split_tf32_kernel<<<grid_for(n, 256, c.num_sms * 8), 256, 0, c.stream>>>(x, n, hi, lo);
    count_launch(c);
}

// C(m x n, ldc) = A(m x k, lda) B(k x n, ldb), column-major, with fp32 accuracy on the tensor cores: both operands are
// split into hi = tf32(x), lo = tf32(x - hi) and three TF32 tensor-core GEMMs (cuBLAS: these are plain library GEMMs)
// accumulate lo*hi, hi*lo, hi*hi in fp32, smallest terms first (products of tf32 numbers are exact in fp32; the dropped
// lo*lo term is 2^-22 relative).  Used for the two dense contractions outside the hot loops that were fp32 FMA GEMMs:
// the truncation product V S of block Krylov-Schur (restarted_block_ks.h:167) and the lift U C
// (src/sparseMatrix.cpp:1446-1449); small products stay on the fp32 FMA path (dense_tc_min_flops).
void gemm_3xtf32(Ctx &c, int m, int n, int k, const float *A, int lda, const float *B, int ldb, float *C, int ldc)
{
    const float one = 1.f, zero = 0.f;
    const double flops = 2.0 * m * (double)n * k;
    if (c.opt("dense_tc", 1) == 0 || flops < 1e6 * (double)c.opt("dense_tc_min_mflops", 20000)) {
        ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, A, lda, B, ldb, &zero, C, ldc));
        count_launch(c);
        return;
    }
    StatScope s(c, "gemm_3xtf32", 0.0, flops);
    const size_t na = (size_t)lda * k, nb = (size_t)ldb * n;
    DevBuf<float> ahi(na), alo(na), bhi(nb), blo(nb);
    split_tf32(c, A, na, ahi.p, alo.p);
    split_tf32(c, B, nb, bhi.p, blo.p);
    const float *aa[3] = {alo.p, ahi.p, ahi.p}, *bb[3] = {bhi.p, blo.p, bhi.p};
    for (int t = 0; t < 3; ++t) {
        ISLE_CUBLAS_CHECK(cublasGemmEx(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, aa[t], CUDA_R_32F, lda, bb[t], CUDA_R_32F, ldb,
                                       t == 0 ? &zero : &one, C, CUDA_R_32F, ldc, CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_GEMM_DEFAULT));
        count_launch(c);
    }
}

// The tensor-core engine needs whole 32-wide K blocks and at least one 16-wide center tile.
bool dist_tc_supported(const Ctx &c, uint32_t kp, uint32_t ncent)
{
    return c.P.p != nullptr && c.P_lo.p != nullptr && kp % BK == 0 && kp >= BK && ncent >= 1 && ncent <= 16384;
}

void dist_tc_launch(Ctx &c, const float *C, const float *c2, uint32_t ncent, int mode, uint32_t *assign, float *min_dist)
{
    const uint32_t DB = (uint32_t)c.DB, kp = (uint32_t)c.kp;
    uint32_t BN = std::min<uint32_t>(kMaxBN, (ncent + 15) / 16 * 16);
    const int bn_opt = c.opt("dist_tc_bn", 0);
    if (bn_opt >= 16 && bn_opt <= kMaxBN && bn_opt % 16 == 0) BN = std::min<uint32_t>(BN, (uint32_t)bn_opt);
    // centers split into hi/lo for this pass
    DevBuf<float> chi((size_t)ncent * kp), clo((size_t)ncent * kp);
    split_tf32(c, C, (size_t)ncent * kp, chi.p, clo.p);

    const uint32_t stage_bytes = 2 * BM * 128 + 2 * BN * 128;
    const uint32_t ctrl_bytes = 1024 + 16 * 8 * 2 + 64 + ncent * 4;
    uint32_t stages = (kSmemBudget - 1024 - ctrl_bytes) / stage_bytes;
    stages = std::max<uint32_t>(2, std::min<uint32_t>(stages, 8));
    const uint32_t smem_bytes = stages * stage_bytes + 1024 + ctrl_bytes;
    ISLE_REQUIRE(smem_bytes <= 227 * 1024, ISLE_ERR_ARG, "dist_tc: shared memory budget exceeded");
    ISLE_CUDA_CHECK(cudaFuncSetAttribute(dist_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

    const CUtensorMap mp_hi = make_map(c.P.p, DB, kp, BM), mp_lo = make_map(c.P_lo.p, DB, kp, BM);   // P itself is the hi operand
    const CUtensorMap mc_hi = make_map(chi.p, ncent, kp, BN), mc_lo = make_map(clo.p, ncent, kp, BN);
    Params p;
    p.d2 = c.p_l2.p; p.c2 = c2; p.assign = assign; p.min_dist = min_dist;
    p.DB = DB; p.ncent = ncent; p.kp = kp; p.BN = BN; p.stages = stages; p.mode = (uint32_t)mode;
    const uint32_t num_m_tiles = (DB + BM - 1) / BM;
    const unsigned grid = std::min<uint32_t>(num_m_tiles, (uint32_t)c.num_sms);
    dist_tc_kernel<<<grid, kThreads, smem_bytes, c.stream>>>(mp_hi, mp_lo, mc_hi, mc_lo, p);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
}

}  // namespace isle
