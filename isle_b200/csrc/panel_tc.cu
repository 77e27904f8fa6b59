// panel_tc.cu -- the two tall-skinny panel products of the block Gram-Schmidt step of block Krylov-Schur
// (reference block-ks/restarted_block_ks.h:83-90: Hk = W^T F, F -= W Hk, three passes per block step)
// on the 5th-generation tensor cores, in split TF32 with fp32-level accuracy.
//
// With b = 10 columns the products carry 6 flop per byte of the basis W: as fp32 FMA kernels (blockks.cu) they
// are bound by the FMA / shared-memory rate and stream W at 2.2 (W^T F) and 3.4 TB/s (F -= W C) at k = 2000,
// where W (141k x 4010 fp32 = 2.26 GB) comes from HBM on every one of the six passes of a block step and the
// orthogonalisation is 55 % of the whole spectral core (profiles/r1g).  On tcgen05 the arithmetic is free and
// the kernels become streams over W:
//
//   D[m, 0:32]  += A_hi[m, k] * [B_hi | B_lo][k, 0:32]        (tcgen05.mma.kind::tf32, M = 128, N = 32, K = 8)
//   D2[m, 0:16] += A_lo[m, k] *  B_hi[k, 0:16]                (N = 16)
//   result[m, c] = (D[m, c] + D[m, 16 + c]) + D2[m, c]         hi*hi + hi*lo + lo*hi; lo*lo (2^-22) dropped
//
// A = a 128 x 32 tile of W: TMA brings the fp32 tile to shared memory, 128 worker threads (thread = tile row =
// TMEM lane) read their 32 values, split them into hi = tf32(x) and lo = x - hi in registers and write both
// to TMEM with tcgen05.st; the MMAs take A from TMEM.  B = the small operand (F or C), pre-split into
// 32 rows [hi(16) | lo(16)] x K by the kernel that produced it, staged by TMA (128-byte swizzle).
//   MODE 0  C = W^T F : m = column of W, k = row of W.   W tile = [128 columns][32 rows] (K-major, swizzled)
//   MODE 1  F -= W C  : m = row of W,    k = column of W. W tile = [32 columns][128 rows] (read transposed)
// A job is (128-row tile of the result, K segment of 16 chunks = 512 k): it accumulates 64 MMAs per
// accumulator (bounded truncation error of the tensor core's fp32 adds), is drained with tcgen05.ld and
// written as one partial result; a fixed-order reduce kernel sums the partials (deterministic: replicated ranks
// stay bitwise in lock step).  Persistent, one CTA per SM, 18 warps: 4 worker groups with private A stages
// and accumulators, one MMA-issuing thread, one TMA thread, an 8-stage mbarrier ring (same protocol as
// spmm_head.cu).  All waits are bounded: a protocol bug records (role, job, chunk) in host-mapped memory and
// traps instead of hanging the GPU.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace isle {

namespace {

using namespace tcptx;

constexpr uint32_t kGroups = 4;
constexpr uint32_t kWarpMma = 4 * kGroups, kWarpTma = 4 * kGroups + 1;
constexpr int kThreads = 32 * (4 * kGroups + 2);
constexpr uint32_t kTmemCols = 512;      // four accumulator slots of 64 columns (D: 32, D2: 16) at +0, four A stages of 64 (hi 32 | lo 32) at +256
constexpr uint32_t kTmemA = 256;
constexpr uint32_t kKC = 32;             // k per chunk = one 128-byte row of fp32
constexpr uint32_t kATile = 128 * 128;   // bytes: 128 x 32 fp32
constexpr uint32_t kBTile = 32 * 128;    // bytes: 32 rows (hi 16 | lo 16) x 32 k fp32
constexpr uint32_t kStageBytes = kATile + kBTile;
constexpr uint32_t kSpinLimit = 1u << 17;


__device__ __noinline__ void panel_timeout(uint32_t *diag, uint32_t code, uint32_t a, uint32_t b)
{
    if (diag) {
        const uint32_t cls = code & 0xFFu;
        const uint32_t slot = cls == 0x10 ? 0 : cls == 0x20 ? 1 : cls == 0x21 ? 2 : (cls & 0xF0u) == 0x30 ? 3 + (cls & 7u) : 11;
        uint32_t *r = diag + slot * 5;
        if (atomicCAS_system(r, 0u, 1u + (0x1000u | code)) == 0u) {     // 0x1000: panel_tc_kernel
            r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = a; r[4] = b;
        }
        __threadfence_system();
        for (int i = 0; i < 4000; ++i) __nanosleep(1000);
    }
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *diag, uint32_t code, uint32_t da, uint32_t db)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spins = 0; spins < kSpinLimit; ++spins) {
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
    panel_timeout(diag, code, da, db);
}
__device__ __forceinline__ void mbar_wait2(uint64_t *bar_a, uint32_t parity_a, uint64_t *bar_b, uint32_t parity_b, uint32_t *diag,
                                           uint32_t code, uint32_t da, uint32_t db)
{
    const uint32_t addr_a = smem_u32(bar_a), addr_b = smem_u32(bar_b);
    uint32_t done = 0, pa = 0;
#pragma unroll 1
    for (uint32_t spins = 0; spins < kSpinLimit; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p, q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
            "selp.u32 %1, 1, 0, p;\n\t"
            "and.pred p, p, q;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done), "=r"(pa)
            : "r"(addr_a), "r"(parity_a), "r"(addr_b), "r"(parity_b)
            : "memory");
        if (done) return;
    }
    panel_timeout(diag, code + (pa ? 0x100u : 0u), da, db);
}
// D[tmem] (+)= A[tmem] * B[smem]; tf32 operands (fp32 bit patterns, low 13 mantissa bits ignored), fp32 accumulator
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}



struct PanelParams {
    float *out;            // [nseg][m_total][16] partial results
    uint32_t m_total;      // valid rows of the result
    uint32_t num_mtiles;   // 128-row tiles of the result
    uint32_t NC;           // 32-k chunks along K
    uint32_t seg;          // chunks per job
    uint32_t nseg;         // K segments = ceil(NC / seg)
    uint32_t stages;
    uint32_t *diag;
    const uint32_t *skip;  // device flag: non-zero = this launch is an elided Gram-Schmidt pass, return at once
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
panel_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const PanelParams p)
{
    if (p.skip && *reinterpret_cast<const volatile uint32_t *>(p.skip)) return;    // uniform: before any barrier or TMEM allocation
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ctrl = smem + (size_t)p.stages * kStageBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);   // [stages]  TMA landed (W tile + B tile)
    uint64_t *empty = full + p.stages;                      // [stages]  MMAs that read the B tile retired + the group's 4 warps hold the W tile in registers
    uint64_t *a_full = empty + p.stages;                    // [4]       group stored its A stage
    uint64_t *a_empty = a_full + kGroups;                   // [4]       MMAs reading the A stage retired
    uint64_t *acc_full = a_empty + kGroups;                 // [4]       job accumulated
    uint64_t *acc_empty = acc_full + kGroups;               // [4]       its group drained the accumulator
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kGroups);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t njobs = p.num_mtiles * p.nseg;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 5); }
        for (uint32_t a = 0; a < kGroups; ++a) {
            mbar_init(&a_full[a], 4); mbar_init(&a_empty[a], 1);
            mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kWarpTma) {
        if (elect_one_sync()) {
            // ===== TMA producer
            uint32_t stage = 0, phase = 0;
            for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
                const uint32_t mtile = job % p.num_mtiles, sg = job / p.num_mtiles;
                const uint32_t c0 = sg * p.seg, c1 = min(p.NC, c0 + p.seg);
                for (uint32_t ch = c0; ch < c1; ++ch) {
                    mbar_wait(&empty[stage], phase ^ 1, p.diag, 0x10, job, ch);
                    uint8_t *st = smem + (size_t)stage * kStageBytes;
                    mbar_expect_tx(&full[stage], kStageBytes);
                    if (MODE == 0) tma_load_2d(st, &map_a, &full[stage], (int)(ch * kKC), (int)(mtile * 128));
                    else tma_load_2d(st, &map_a, &full[stage], (int)(mtile * 128), (int)(ch * kKC));
                    tma_load_2d(st + kATile, &map_b, &full[stage], (int)(ch * kKC), 0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer.  idesc: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major, N>>3 at 17, M>>4 at 24
        const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        if (elect_one_sync()) {
            const uint64_t desc0 = umma_desc(smem_u32(smem) + kATile);
            uint32_t stage = 0, phase = 0, g = 0, gphase = 0, buf = 0, bphase = 0;
            for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
                const uint32_t sg = job / p.num_mtiles;
                const uint32_t c0 = sg * p.seg, c1 = min(p.NC, c0 + p.seg);
                mbar_wait(&acc_empty[buf], bphase ^ 1, p.diag, 0x20, job, c0);
                const uint32_t d_tmem = tmem_base + buf * 64;
                for (uint32_t ch = c0; ch < c1; ++ch) {
                    mbar_wait2(&full[stage], phase, &a_full[g], gphase, p.diag, 0x21, job, ch);
                    tc_fence_after();
                    const uint64_t bd = desc0 + (uint64_t)(stage * (kStageBytes >> 4));
                    const uint32_t a_tmem = tmem_base + kTmemA + g * 64;
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i) {
                        // K step i: 8 k = 8 TMEM columns of A, 32 bytes along K of B
                        const uint32_t acc = (ch > c0 || i > 0) ? 1u : 0u;
                        umma_tf32_ts(d_tmem, a_tmem + i * 8, bd + (uint64_t)(i * 2), idesc32, acc);             // hi x [hi | lo]
                        umma_tf32_ts(d_tmem + 32, a_tmem + 32 + i * 8, bd + (uint64_t)(i * 2), idesc16, acc);   // lo x hi
                    }
                    umma_commit(&empty[stage]);
                    umma_commit(&a_empty[g]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    if (++g == kGroups) { g = 0; gphase ^= 1; }
                }
                umma_commit(&acc_full[buf]);
                if (++buf == kGroups) { buf = 0; bphase ^= 1; }
            }
        }
    } else {
        // ===== workers: group = warp / 4, lane quarter = warp % 4; thread <-> row of the tile <-> TMEM lane
        const uint32_t grp = warp >> 2, quarter = warp & 3;
        const uint32_t trow = quarter * 32 + lane;
        const uint32_t lane_base = (quarter * 32u) << 16;
        uint32_t stage = 0, phase = 0, g = 0, gphase = 0, buf = 0, bphase = 0;
        bool pend = false;
        uint32_t pend_buf = 0, pend_phase = 0, pend_row = 0, pend_sg = 0;

        auto drain = [&]() {
            mbar_wait(&acc_full[pend_buf], pend_phase, p.diag, 0x30 + grp, pend_row, pend_buf * 2 + pend_phase);
            tc_fence_after();
            uint32_t acc[48];
#pragma unroll
            for (uint32_t cb = 0; cb < 48; cb += 16) tmem_ld16(tmem_base + lane_base + pend_buf * 64 + cb, acc + cb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[pend_buf]);
            if (pend_row < p.m_total) {
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    v[c] = (__uint_as_float(acc[c]) + __uint_as_float(acc[16 + c])) + __uint_as_float(acc[32 + c]);
                float4 *dst = reinterpret_cast<float4 *>(p.out) + ((size_t)pend_sg * p.m_total + pend_row) * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            pend = false;
        };

        for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
            const uint32_t mtile = job % p.num_mtiles, sg = job / p.num_mtiles;
            const uint32_t c0 = sg * p.seg, c1 = min(p.NC, c0 + p.seg);
            for (uint32_t ch = c0; ch < c1; ++ch) {
                if (g == grp) {
                    mbar_wait2(&full[stage], phase, &a_empty[grp], gphase ^ 1, p.diag, 0x34 + grp, job, ch);
                    const uint8_t *tile = smem + (size_t)stage * kStageBytes;
                    float x[32];
                    if (MODE == 0) {
                        // row trow of the [128][32] tile; TMA's 128-byte swizzle: 16-byte unit q sits at q ^ (row & 7)
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q) {
                            const float4 v = *reinterpret_cast<const float4 *>(tile + trow * 128u + ((q ^ (trow & 7u)) << 4));
                            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                        }
                    } else {
                        // column trow of the [32][128] tile (lanes read consecutive words)
                        const float *tf = reinterpret_cast<const float *>(tile);
#pragma unroll
                        for (uint32_t k = 0; k < 32; ++k) x[k] = tf[k * 128u + trow];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);       // the tile is in registers
                    tc_fence_after();
                    const uint32_t a_addr = tmem_base + lane_base + kTmemA + grp * 64;
                    uint32_t r[16];
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
#pragma unroll
                        for (uint32_t k = 0; k < 16; ++k) r[k] = tf32_rna(x[16 * h + k]);
                        tmem_st16(a_addr + 16 * h, r);
#pragma unroll
                        for (uint32_t k = 0; k < 16; ++k) r[k] = __float_as_uint(x[16 * h + k] - __uint_as_float(r[k]));
                        tmem_st16(a_addr + 32 + 16 * h, r);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[grp]);
                    if (pend) drain();
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                if (++g == kGroups) { g = 0; gphase ^= 1; }
            }
            if (buf == grp) {      // job n uses accumulator n mod 4 and is drained by group n mod 4
                if (pend) drain();
                pend = true;
                pend_buf = buf;
                pend_phase = bphase;
                pend_row = mtile * 128 + trow;
                pend_sg = sg;
            }
            if (++buf == kGroups) { buf = 0; bphase ^= 1; }
        }
        if (pend) drain();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}


// fp32 [outer][inner] row-major (inner contiguous, row pitch `pitch` floats), box = box_inner x box_outer, zero fill out of bounds
CUtensorMap make_map(const float *base, uint64_t inner, uint64_t outer, uint64_t pitch, uint32_t box_inner, uint32_t box_outer,
                     bool swizzle128)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {inner, outer};
    const cuuint64_t gstride[1] = {pitch * sizeof(float)};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ISLE_REQUIRE(r == CUDA_SUCCESS, ISLE_ERR_CUDA, "cuTensorMapEncodeTiled (panel operand) failed (" + std::to_string((int)r) + ")");
    return m;
}

// split[c][i] = tf32(X[i + c ld]), split[16 + c][i] = X - that, for c < b, i < len (rows c >= b stay zero from the allocation);
// i in [len, len_pad) is zeroed so that nothing non-finite can meet the zero-filled A tiles
__global__ void __launch_bounds__(256)
split_operand_kernel(const float *__restrict__ X, size_t ld, uint32_t len, uint32_t len_pad, int b, size_t pitch, float *__restrict__ split)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len_pad) return;
    for (int c = 0; c < b; ++c) {
        float hi = 0.f, lo = 0.f;
        if (i < len) {
            const float x = X[i + (size_t)c * ld];
            hi = __uint_as_float(tf32_rna(x));
            lo = x - hi;
        }
        split[(size_t)c * pitch + i] = hi;
        split[(size_t)(16 + c) * pitch + i] = lo;
    }
}

// C[j + c ldc] = sum over segments (fixed order: lane-strided partial sums, then a shuffle tree); Hk (+)= the same;
// csplit rows c / 16 + c = hi / lo of the value (operand of the F -= W C product that follows).  One warp per (j, c).
__global__ void __launch_bounds__(256)
wtf_reduce_tc_kernel(const float *__restrict__ partial, int nseg, int rows, int b, float *__restrict__ C, int ldc,
                     float *__restrict__ Hk, int ldh, int hk_assign, float *__restrict__ csplit, size_t cpitch,
                     uint32_t *__restrict__ cmax_bits, const uint32_t *__restrict__ skip)
{
    if (skip && *skip) return;
    const int lane = threadIdx.x & 31;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= rows * 16) return;
    const int j = t >> 4, c = t & 15;
    if (c >= b) return;
    float v = 0.f;
    for (int s = lane; s < nseg; s += 32) v += partial[((size_t)s * rows + j) * 16 + c];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
        C[j + (size_t)c * ldc] = v;
        if (Hk) Hk[j + (size_t)c * ldh] = hk_assign ? v : Hk[j + (size_t)c * ldh] + v;
        const float hi = __uint_as_float(tf32_rna(v));
        csplit[(size_t)c * cpitch + j] = hi;
        csplit[(size_t)(16 + c) * cpitch + j] = v - hi;
        // largest coefficient of this pass (bit patterns of non-negative floats order like integers: the
        // maximum is independent of the order of the atomics, so replicated ranks decide identically)
        if (cmax_bits) atomicMax(cmax_bits, __float_as_uint(fabsf(v)) & 0x7FFFFFFFu);
    }
}

// Row-sharded basis (SURVEY 8e option B): every rank holds rows [r0, r1) of W and F, so W^T F is a sum over ranks.
// pack: packed[j * 16 + c] = this rank's sum over its segments (same fixed order as above); after the all-reduce,
// finalize writes C, Hk, csplit and the pass maximum from the global coefficients (identical on every rank).
__global__ void __launch_bounds__(256)
wtf_pack_tc_kernel(const float *__restrict__ partial, int nseg, int rows, int b, float *__restrict__ packed,
                   const uint32_t *__restrict__ skip)
{
    if (skip && *skip) return;
    const int lane = threadIdx.x & 31;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= rows * 16) return;
    const int j = t >> 4, c = t & 15;
    float v = 0.f;
    if (c < b)
        for (int s = lane; s < nseg; s += 32) v += partial[((size_t)s * rows + j) * 16 + c];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) packed[t] = v;
}

__global__ void __launch_bounds__(256)
wtf_finalize_tc_kernel(const float *__restrict__ packed, int rows, int b, float *__restrict__ C, int ldc, float *__restrict__ Hk,
                       int ldh, int hk_assign, float *__restrict__ csplit, size_t cpitch, uint32_t *__restrict__ cmax_bits,
                       const uint32_t *__restrict__ skip)
{
    if (skip && *skip) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * 16) return;
    const int j = t >> 4, c = t & 15;
    if (c >= b) return;
    const float v = packed[t];
    C[j + (size_t)c * ldc] = v;
    if (Hk) Hk[j + (size_t)c * ldh] = hk_assign ? v : Hk[j + (size_t)c * ldh] + v;
    const float hi = __uint_as_float(tf32_rna(v));
    csplit[(size_t)c * cpitch + j] = hi;
    csplit[(size_t)(16 + c) * cpitch + j] = v - hi;
    if (cmax_bits) atomicMax(cmax_bits, __float_as_uint(fabsf(v)) & 0x7FFFFFFFu);
}

// skip = 1 when the pass just finished only moved coefficients below `ratio` of the first pass's (they are then
// below fp32 resolution of H and a further pass cannot improve the orthogonality); counts the elisions
__global__ void gs_decide_kernel(const uint32_t *__restrict__ cmax_bits, float ratio, uint32_t *__restrict__ skip,
                                 uint32_t *__restrict__ elided)
{
    const float c1 = __uint_as_float(cmax_bits[0]), c2 = __uint_as_float(cmax_bits[1]);
    const uint32_t s = (c2 <= ratio * c1 && isfinite(c1) && isfinite(c2)) ? 1u : 0u;
    *skip = s;
    *elided += s;
}

// F[i + c n] -= sum over segments of delta[s][i][c] (fixed order); fsplit = hi / lo of the new F (operand of the next W^T F)
__global__ void __launch_bounds__(256)
fsub_reduce_tc_kernel(const float4 *__restrict__ delta, int nseg, uint32_t n, int b, float *__restrict__ F, float *__restrict__ fsplit,
                      size_t fpitch, const uint32_t *__restrict__ skip)
{
    if (skip && *skip) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    for (int s = 0; s < nseg; ++s) {
        const float4 *d = delta + ((size_t)s * n + i) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 t = __ldg(d + q);
            v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        if (c < b) {
            const float x = F[i + (size_t)c * n] - v[c];
            F[i + (size_t)c * n] = x;
            const float hi = __uint_as_float(tf32_rna(x));
            fsplit[(size_t)c * fpitch + i] = hi;
            fsplit[(size_t)(16 + c) * fpitch + i] = x - hi;
        }
    }
}

template <int MODE>
void launch_panel(Ctx &c, const CUtensorMap &ma, const CUtensorMap &mb, PanelParams p)
{
    p.stages = (uint32_t)std::max(2, std::min(10, c.opt("ks_panel_tc_stages", 8)));
    p.diag = c.head_diag_dev;
    const uint32_t smem_bytes = p.stages * kStageBytes + 1024 + 512;
    // per launch, not once per process: the attribute is per device and a process may hold contexts on several
    ISLE_CUDA_CHECK(cudaFuncSetAttribute(panel_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const uint32_t njobs = p.num_mtiles * p.nseg;
    const unsigned grid = std::min<uint32_t>(njobs, (uint32_t)c.num_sms);
    panel_tc_kernel<MODE><<<grid, kThreads, smem_bytes, c.stream>>>(ma, mb, p);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host API
// Scratch of the tensor-core panel engine for a basis of n rows and at most ncv columns.
void PanelTc::init(Ctx &c, int64_t n_, int ncv_)
{
    n = n_;
    ncv = ncv_;
    seg = (uint32_t)std::max(1, std::min(64, c.opt("ks_panel_tc_seg", 16)));
    fpitch = (size_t)((n + 3) / 4 * 4);
    cpitch = (size_t)((ncv + 31) / 32 * 32);
    fsplit.alloc(32 * fpitch);
    csplit.alloc(32 * cpitch);
    ISLE_CUDA_CHECK(cudaMemsetAsync(fsplit.p, 0, fsplit.bytes(), c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(csplit.p, 0, csplit.bytes(), c.stream));
    const size_t nseg_w = ((size_t)(n + kKC - 1) / kKC + seg - 1) / seg;          // W^T F: K = n
    const size_t nseg_f = ((size_t)(ncv + kKC - 1) / kKC + seg - 1) / seg;        // F -= W C: K <= ncv
    part.alloc(std::max(nseg_w * (size_t)ncv, nseg_f * (size_t)n) * 16);
    flags.alloc(4);        // [0], [1]: largest |coefficient| of the first / the latest pass; [2]: skip flag; [3]: elided passes
    ISLE_CUDA_CHECK(cudaMemsetAsync(flags.p, 0, flags.bytes(), c.stream));
    if (sharded) {
        packed.alloc((size_t)ncv * 16);
        ISLE_CUDA_CHECK(cudaMemsetAsync(packed.p, 0, packed.bytes(), c.stream));
    }
}

// Start of a block step: no pass elided yet.
void PanelTc::begin_step(Ctx &c)
{
    ISLE_CUDA_CHECK(cudaMemsetAsync(flags.p, 0, 3 * sizeof(uint32_t), c.stream));
}

// After the second pass: elide what follows when its coefficients were below `ratio` of the first pass's.
void PanelTc::decide_elision(Ctx &c, float ratio)
{
    gs_decide_kernel<<<1, 1, 0, c.stream>>>(flags.p, ratio, flags.p + 2, flags.p + 3);
    count_launch(c);
}

uint32_t PanelTc::elided(Ctx &c)
{
    uint32_t v = 0;
    read_small(c, &v, flags.p + 3, 4);
    return v;
}

bool PanelTc::usable(int64_t n_) { return n_ % 4 == 0; }   // TMA needs 16-byte global strides

// fsplit <- hi / lo of F (n x b, ld n): operand of the first W^T F of a block step
void PanelTc::split_F(Ctx &c, const float *F, int b)
{
    split_operand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(F, (size_t)n, (uint32_t)n, (uint32_t)n, b, fpitch, fsplit.p);
    count_launch(c);
}

// C(rows x b, ld ldc) = W^T F with W = V[:, 0:rows) (ld n) and F as last split; Hk (ld ldh) assigned or incremented;
// leaves csplit = hi / lo of C for the F -= W C that follows.
void PanelTc::wtf(Ctx &c, const float *W, int rows, int b, float *C, int ldc, float *Hk, int ldh, bool assign, int pass)
{
    PanelParams p{};
    p.skip = pass >= 2 ? flags.p + 2 : nullptr;
    uint32_t *cmax = pass == 0 ? flags.p : (pass == 1 ? flags.p + 1 : nullptr);
    p.out = part.p;
    p.m_total = (uint32_t)rows;
    p.num_mtiles = (uint32_t)(rows + 127) / 128;
    p.NC = (uint32_t)((n + kKC - 1) / kKC);
    p.seg = seg;
    p.nseg = (p.NC + seg - 1) / seg;
    // W as [rows (outer: its columns)][n (inner)]: tile = 128 columns x 32 rows, K-major, 128-byte swizzle
    const CUtensorMap ma = make_map(W, (uint64_t)n, (uint64_t)rows, (uint64_t)n, kKC, 128, true);
    const CUtensorMap mb = make_map(fsplit.p, (uint64_t)n, 32, fpitch, kKC, 32, true);
    {
        StatScope s(c, "ks_wtf", (double)n * rows * 4.0, 2.0 * (double)n * rows * b);
        launch_panel<0>(c, ma, mb, p);
    }
    {
        StatScope s(c, "ks_wtfred", (double)p.nseg * rows * 64.0);
        // stale csplit columns of an earlier, wider call cannot matter: the W tile is zero beyond `rows`
        // and csplit only ever holds finite values
        if (sharded) {
            // the basis is row-sharded: the coefficients are a sum over ranks.  An elided pass still takes part in
            // the collective (every rank took the same decision and left `packed` alone: the sum is never read).
            wtf_pack_tc_kernel<<<(unsigned)(((size_t)rows * 16 * 32 + 255) / 256), 256, 0, c.stream>>>(part.p, (int)p.nseg, rows, b,
                                                                                                       packed.p, p.skip);
            allreduce_sum_f32(c, packed.p, (size_t)rows * 16);
            wtf_finalize_tc_kernel<<<(unsigned)(((size_t)rows * 16 + 255) / 256), 256, 0, c.stream>>>(
                packed.p, rows, b, C, ldc, Hk, ldh, assign ? 1 : 0, csplit.p, cpitch, cmax, p.skip);
            count_launch(c, 2);
        } else {
            wtf_reduce_tc_kernel<<<(unsigned)(((size_t)rows * 16 * 32 + 255) / 256), 256, 0, c.stream>>>(
                part.p, (int)p.nseg, rows, b, C, ldc, Hk, ldh, assign ? 1 : 0, csplit.p, cpitch, cmax, p.skip);
            count_launch(c);
        }
    }
}

// F -= W C with C as left in csplit by wtf(); leaves fsplit = hi / lo of the new F.
void PanelTc::fsub(Ctx &c, const float *W, int rows, int b, float *F, int pass)
{
    PanelParams p{};
    p.skip = pass >= 2 ? flags.p + 2 : nullptr;
    p.out = part.p;
    p.m_total = (uint32_t)n;
    p.num_mtiles = (uint32_t)((n + 127) / 128);
    p.NC = (uint32_t)(rows + kKC - 1) / kKC;
    p.seg = seg;
    p.nseg = (p.NC + seg - 1) / seg;
    // W as [rows (outer: its columns)][n (inner)]: tile = 32 columns x 128 rows, read transposed by the workers
    const CUtensorMap ma = make_map(W, (uint64_t)n, (uint64_t)rows, (uint64_t)n, 128, kKC, false);
    const CUtensorMap mb = make_map(csplit.p, (uint64_t)cpitch, 32, cpitch, kKC, 32, true);
    {
        StatScope s(c, "ks_fsub", (double)n * rows * 4.0, 2.0 * (double)n * rows * b);
        launch_panel<1>(c, ma, mb, p);
        fsub_reduce_tc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(reinterpret_cast<const float4 *>(part.p), (int)p.nseg,
                                                                                (uint32_t)n, b, F, fsplit.p, fpitch, p.skip);
        count_launch(c);
    }
}

}  // namespace isle
