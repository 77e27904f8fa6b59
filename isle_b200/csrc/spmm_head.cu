// spmm_head.cu -- kernel family (2), dense-head engine of the B * B^T * X operator.
//
// B is a 0/1 pattern scaled per row and word frequencies are Zipfian: the few
// thousand most frequent words hold more than half of B's nonzeros at densities of 1-90 %.
// For those rows an index list + one 64-byte dense-row gather per nonzero (spmm.cu, bound by
// the L2->SM fill path at ~1.5 clk per nonzero per SM) is the wrong format.  Here the head block
// is kept as a bitmap (1 bit per cell instead of 32 per nonzero) in both orientations and both
// passes of MKL_SpSpTrProd::multiply (reference include/matUtils.h:336-365) become
//     out[m, :] = sum_k bit(m, k) * In[k, :]                m: doc (pass 1) / head word (pass 2)
// on the 5th-gen tensor cores with exact arithmetic: the bits are expanded in registers to bf16
// (value 2.0 = the single bit 0x4000, so one shift + one and makes two cells), written to
// TMEM with tcgen05.st and used as the A operand (tcgen05.mma, A from TMEM); the dense
// operand is split into three bf16 pieces (hi + mid + lo = 24 significand bits, products with 2.0
// are exact, fp32 accumulation in TMEM), staged by TMA (128-byte swizzle) as the B operand.
// The three partial columns are summed and halved in the epilogue.
//
// Kernel shape (persistent, up to two CTAs per SM, 6 warps each):
//   warps 0-3  workers: thread t owns row t of the 128-row tile = TMEM lane t; per 128-k chunk:
//              read 128 bits from the staged bit tile, expand to 64 registers, tcgen05.st into one
//              of two A stages; at the end of a job tcgen05.ld the accumulator and write the rows
//   warp 4     lane 0 issues 8 x tcgen05.mma.kind::f16 (M=128, N=16/32/48, K=16) per chunk
//   warp 5     lane 0 issues TMA: two [N x 64] bf16 tiles + the 2 KB bit tile per stage
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace isle {

namespace {

using namespace tcptx;

constexpr uint32_t kGroups = 4;          // worker groups = A stages in TMEM
constexpr uint32_t kWarpMma = 4 * kGroups, kWarpTma = 4 * kGroups + 1;
constexpr int kThreads = 32 * (4 * kGroups + 2);
constexpr uint32_t kTmemCols = 512;      // four accumulators (<= 48 cols, stride 64) at +0, four A stages of 64 cols at +256
constexpr uint32_t kTmemA = 256;
constexpr uint32_t kBitBytes = kHeadTile * 16;   // 128 rows x 128 bits

// ---------------------------------------------------------------------------------- PTX
// Bounded waits: a protocol bug records where it stopped (role / barrier / job / chunk, first writer wins)
// in host-mapped memory and traps, so the launch fails with a diagnosable error instead of hanging.
__device__ uint32_t *g_head_diag = nullptr;
constexpr uint32_t kSpinLimit = 1u << 17;

__device__ __noinline__ void head_timeout(uint32_t code, uint32_t a, uint32_t b)
{
    // one record per waiter class (code low byte: 0x10 TMA, 0x20/0x21 MMA, 0x30/0x31 + group workers)
    uint32_t *d = g_head_diag;
    if (d) {
        const uint32_t cls = code & 0xFFu;
        const uint32_t slot = cls == 0x10 ? 0 : cls == 0x20 ? 1 : cls == 0x21 ? 2 : (cls & 0xF0u) == 0x30 ? 3 + (cls & 7u) : 11;
        uint32_t *r = d + slot * 5;
        if (atomicCAS_system(r, 0u, 1u + code) == 0u) {
            r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = a; r[4] = b;
        }
        __threadfence_system();
        // give the other stuck roles of this CTA time to leave their records before the context dies
        for (int i = 0; i < 4000; ++i) __nanosleep(1000);
    }
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t code = 0, uint32_t da = 0, uint32_t db = 0)
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
    head_timeout(code, da, db);
}
// Waits for two barriers; the two try_waits are issued back to back so their latencies overlap.
__device__ __forceinline__ void mbar_wait2(uint64_t *bar_a, uint32_t parity_a, uint64_t *bar_b, uint32_t parity_b,
                                           uint32_t code = 0, uint32_t da = 0, uint32_t db = 0)
{
    const uint32_t addr_a = smem_u32(bar_a), addr_b = smem_u32(bar_b);
    uint32_t done = 0;
    uint32_t pa = 0;
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
    head_timeout(code + (pa ? 0x100u : 0u), da, db);     // +0x100: the first barrier had completed
}
// D[tmem] (+)= A[tmem] * B[smem]; bf16 operands, fp32 accumulator
// `issue` != 0 on the one lane that issues (the instruction is predicated, not branched around, so
// the surrounding loop stays warp-uniform and its operands can live in uniform registers)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                             uint32_t issue)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar, uint32_t issue)
{
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(issue)
        : "memory");
}


// 32 bits -> 32 bf16 cells in 16 registers: register j holds k = 2j (low half, bit j) and
// k = 2j+1 (high half, bit j+16); a set bit becomes 0x4000 = 2.0.
__device__ __forceinline__ void expand_word(uint32_t m, uint32_t *r)
{
#pragma unroll
    for (int j = 0; j < 15; ++j) r[j] = (m << (14 - j)) & 0x40004000u;
    r[15] = (m >> 1) & 0x40004000u;
}

struct HeadParams {
    const uint4 *bits;     // [(mtile * NC + chunk) * 128 + row] : 128 k of one row
    float *out;            // [rows][16] fp32
    uint32_t m_valid;      // rows of `out` that exist
    uint32_t num_mtiles;   // 128-row tiles
    uint32_t NC;           // 128-k chunks along K
    uint32_t nsplit;       // K ranges per tile
    uint32_t stages;
    uint32_t seg;          // chunks per accumulation segment (the TMEM accumulator is drained after each)
    uint32_t atomic;       // 1: segments are added into `out` (pre-zeroed); 0: one segment per row, stored
};

// The tensor core adds into its fp32 accumulator with truncation, so the rounding error of a long
// accumulation chain grows linearly.  A job's K range is therefore cut into segments of p.seg chunks
// (8 MMAs each); segment n accumulates in TMEM accumulator n mod 4 and is drained (tcgen05.ld, three-piece
// sum, atomic add or store) by worker group n mod 4 while the next segments are being multiplied.  Each
// accumulator's full/empty barrier pair therefore has exactly one waiter on each side, observing every phase
// in order -- a parity wait cannot tell "phase n done" from "phase n-2 done", so a shared pair would let a
// group that ran two short segments ahead of the tensor core see its wait satisfied early.
//
// Workers: 16 warps = 4 groups x 4 lane quarters.  Group g expands the chunks whose running index is
// g mod 4 into its own A stage (TMEM columns 128 + 64 g), so four chunks are in flight and the
// mbarrier / tcgen05.st latencies of one group hide behind the other three.
//
// BS = column stride of one bf16 piece inside the N dimension: n = piece * BS + c.
template <int BS>
__global__ void __launch_bounds__(kThreads, 1)
spmm_head_kernel(const __grid_constant__ CUtensorMap map_b, const HeadParams p)
{
    constexpr uint32_t N = (3 * BS + 15) / 16 * 16;
    constexpr uint32_t sub_bytes = N * 128;                   // one [N x 64 k] bf16 swizzled tile
    constexpr uint32_t stage_bytes = 2 * sub_bytes + kBitBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ctrl = smem + (size_t)p.stages * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);   // [stages]  TMA landed (B tiles + bits)
    uint64_t *empty = full + p.stages;                      // [stages]  MMA done + the group's 4 warps have read the bits
    uint64_t *a_full = empty + p.stages;                    // [4]       group stored its A stage
    uint64_t *a_empty = a_full + kGroups;                   // [4]       MMAs reading the A stage retired
    uint64_t *acc_full = a_empty + kGroups;                 // [4]       segment accumulated
    uint64_t *acc_empty = acc_full + kGroups;               // [4]       its group drained the accumulator
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kGroups);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t njobs = p.num_mtiles * p.nsplit;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 5); }
        for (uint32_t a = 0; a < kGroups; ++a) {
            mbar_init(&a_full[a], 4); mbar_init(&a_empty[a], 1);
            mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
                const uint32_t mtile = job / p.nsplit, part = job % p.nsplit;
                const uint32_t c0 = (uint32_t)((uint64_t)p.NC * part / p.nsplit);
                const uint32_t c1 = (uint32_t)((uint64_t)p.NC * (part + 1) / p.nsplit);
                for (uint32_t ch = c0; ch < c1; ++ch) {
                    mbar_wait(&empty[stage], phase ^ 1, 0x10, job, ch);
                    uint8_t *st = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full[stage], stage_bytes);
                    tma_load_2d(st, &map_b, &full[stage], (int)(ch * kHeadChunk), 0);
                    tma_load_2d(st + sub_bytes, &map_b, &full[stage], (int)(ch * kHeadChunk + 64), 0);
                    bulk_load_1d(st + 2 * sub_bytes, p.bits + ((size_t)mtile * p.NC + ch) * kHeadTile, kBitBytes, &full[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer (one thread).  Each MMA is only 16 tensor-pipe cycles (N = 32), so the issue
        // sequence is kept to a few instructions per MMA: all descriptor arithmetic is an add of a constant.
        // idesc: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((uint32_t)(kHeadTile >> 4) << 24);
        if (elect_one_sync()) {     // ptxas then knows a single thread is active: plain R2UR, no waterfall loops
            const uint64_t desc0 = umma_desc(smem_u32(smem));
            uint32_t stage = 0, phase = 0, g = 0, gphase = 0, buf = 0, bphase = 0;
            for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
                const uint32_t part = job % p.nsplit;
                const uint32_t c0 = (uint32_t)((uint64_t)p.NC * part / p.nsplit);
                const uint32_t c1 = (uint32_t)((uint64_t)p.NC * (part + 1) / p.nsplit);
                for (uint32_t s0 = c0; s0 < c1; s0 += p.seg) {
                    const uint32_t s1 = min(c1, s0 + p.seg);
                    mbar_wait(&acc_empty[buf], bphase ^ 1, 0x20, job, s0);
                    const uint32_t d_tmem = tmem_base + buf * 64;
                    for (uint32_t ch = s0; ch < s1; ++ch) {
                        mbar_wait2(&full[stage], phase, &a_full[g], gphase, 0x21, job, ch);
                        tc_fence_after();
                        const uint64_t bd = desc0 + (uint64_t)(stage * (stage_bytes >> 4));
                        const uint32_t a_tmem = tmem_base + kTmemA + g * 64;
                        const uint32_t first = ch > s0 ? 1u : 0u;
#pragma unroll
                        for (uint32_t i = 0; i < 8; ++i) {
                            // MMA i: k = 16 i .. 16 i + 15 of the chunk = 8 TMEM columns of A, 32 bytes along K of B
                            constexpr uint32_t kSub = sub_bytes >> 4;
                            umma_bf16_ts(d_tmem, a_tmem + i * 8, bd + (uint64_t)((i >> 2) * kSub + (i & 3) * 2), idesc,
                                         i > 0 ? 1u : first, 1u);
                        }
                        umma_commit(&empty[stage], 1u);
                        umma_commit(&a_empty[g], 1u);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        if (++g == kGroups) { g = 0; gphase ^= 1; }
                    }
                    umma_commit(&acc_full[buf], 1u);
                    if (++buf == kGroups) { buf = 0; bphase ^= 1; }
                }
            }
        }
    } else {
        // ===== workers: group = warp / 4, lane quarter = warp % 4; thread <-> row of the tile <-> TMEM lane
        const uint32_t grp = warp >> 2, quarter = warp & 3;
        const uint32_t trow = quarter * 32 + lane;
        const uint32_t lane_base = (quarter * 32u) << 16;
        uint32_t stage = 0, phase = 0, g = 0, gphase = 0, buf = 0, bphase = 0;
        // the segment this group has to drain: accumulator, barrier phase, output row, store/add
        bool pend = false;
        uint32_t pend_buf = 0, pend_phase = 0, pend_row = 0;

        auto drain = [&]() {
            mbar_wait(&acc_full[pend_buf], pend_phase, 0x30 + grp, pend_row, pend_buf * 2 + pend_phase);
            tc_fence_after();
            uint32_t acc[N];
#pragma unroll
            for (uint32_t cb = 0; cb < N; cb += 16) tmem_ld16(tmem_base + lane_base + pend_buf * 64 + cb, acc + cb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[pend_buf]);
            if (pend_row < p.m_valid) {
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    v[c] = c < BS ? 0.5f * ((__uint_as_float(acc[2 * BS + c]) + __uint_as_float(acc[BS + c])) + __uint_as_float(acc[c]))
                                  : 0.0f;
                float4 *dst = reinterpret_cast<float4 *>(p.out) + (size_t)pend_row * 4;
                if (p.atomic) {
#pragma unroll
                    for (int q = 0; q < (BS + 3) / 4; ++q)
                        atomicAdd(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            pend = false;
        };

        for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
            const uint32_t mtile = job / p.nsplit, part = job % p.nsplit;
            const uint32_t c0 = (uint32_t)((uint64_t)p.NC * part / p.nsplit);
            const uint32_t c1 = (uint32_t)((uint64_t)p.NC * (part + 1) / p.nsplit);
            for (uint32_t s0 = c0; s0 < c1; s0 += p.seg) {
                const uint32_t s1 = min(c1, s0 + p.seg);
                for (uint32_t ch = s0; ch < s1; ++ch) {
                    if (g == grp) {
                        mbar_wait2(&full[stage], phase, &a_empty[grp], gphase ^ 1, 0x34 + grp, job, ch);
                        const uint4 m = reinterpret_cast<const uint4 *>(smem + (size_t)stage * stage_bytes + 2 * sub_bytes)[trow];
                        tc_fence_after();
                        const uint32_t a_addr = tmem_base + lane_base + kTmemA + grp * 64;
                        uint32_t r[16];
                        expand_word(m.x, r); tmem_st16(a_addr, r);
                        expand_word(m.y, r); tmem_st16(a_addr + 16, r);
                        expand_word(m.z, r); tmem_st16(a_addr + 32, r);
                        expand_word(m.w, r); tmem_st16(a_addr + 48, r);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[stage]);       // bits are in registers
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&a_full[grp]);
                        // a finished segment assigned to this group is drained while later chunks multiply
                        if (pend) drain();
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    if (++g == kGroups) { g = 0; gphase ^= 1; }
                }
                if (buf == grp) {      // segment n uses accumulator n mod 4 and is drained by group n mod 4
                    if (pend) drain();
                    pend = true;
                    pend_buf = buf;
                    pend_phase = bphase;
                    pend_row = mtile * kHeadTile + trow;
                }
                if (++buf == kGroups) { buf = 0; bphase ^= 1; }
            }
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


// [nrows x kpad] bf16 row-major (K contiguous), box = 64 k x nrows, 128-byte swizzle
CUtensorMap make_split_map(const __nv_bfloat16 *base, uint64_t nrows, uint64_t kpad)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {kpad, nrows};
    const cuuint64_t gstride[1] = {kpad * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {64, (cuuint32_t)nrows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ISLE_REQUIRE(r == CUDA_SUCCESS, ISLE_ERR_CUDA, "cuTensorMapEncodeTiled (head operand) failed (" + std::to_string((int)r) + ")");
    return m;
}

template <int BS>
void launch_head_t(Ctx &c, const __nv_bfloat16 *split, uint64_t kpad, const HeadParams &p0, cudaStream_t stream)
{
    constexpr uint32_t N = (3 * BS + 15) / 16 * 16;
    constexpr uint32_t stage_bytes = 2 * N * 128 + kBitBytes;
    HeadParams p = p0;
    p.stages = (uint32_t)std::max(2, std::min(16, c.opt("spmm_head_stages", 12)));
    const uint32_t smem_bytes = p.stages * stage_bytes + 1024 + 512;
    ISLE_CUDA_CHECK(cudaFuncSetAttribute(spmm_head_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (c.head_diag_dev) {
        static thread_local uint32_t *installed = nullptr;     // device-side symbol, set once per host thread (= per device) and pointer
        if (installed != c.head_diag_dev) {
            ISLE_CUDA_CHECK(cudaMemcpyToSymbolAsync(g_head_diag, &c.head_diag_dev, sizeof(uint32_t *), 0, cudaMemcpyHostToDevice, stream));
            installed = c.head_diag_dev;
        }
    }
    const CUtensorMap map = make_split_map(split, N, kpad);
    const uint32_t njobs = p.num_mtiles * p.nsplit;
    const unsigned grid = std::min<uint32_t>(njobs, (uint32_t)c.num_sms);
    spmm_head_kernel<BS><<<grid, kThreads, smem_bytes, stream>>>(map, p);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
}

}  // namespace

int head_block_stride(int b) { return b <= 5 ? 5 : (b <= 10 ? 10 : 16); }
int head_split_rows(int b) { return (3 * head_block_stride(b) + 15) / 16 * 16; }

// out[m, 0:b] (+)= sum_k bit(m, k) * In[k, 0:b], In given as its 3-piece bf16 split [N][kpad].
// zero_out: `out` (num_mtiles * 128 rows) is cleared first when partial sums are added atomically.
void spmm_head_launch(Ctx &c, int b, const uint4 *bits, uint32_t num_mtiles, uint32_t NC, uint32_t nsplit,
                      const __nv_bfloat16 *split, float *out, uint32_t m_valid, bool zero_out, bool force_atomic,
                      cudaStream_t stream)
{
    if (!num_mtiles || !NC) return;
    HeadParams p;
    p.bits = bits; p.out = out; p.m_valid = m_valid; p.num_mtiles = num_mtiles; p.NC = NC;
    p.nsplit = std::max<uint32_t>(1, std::min(nsplit, NC));
    p.stages = 0;
    p.seg = (uint32_t)std::max(1, c.opt("spmm_head_seg", 16));
    // rows receive several partial sums (K split over jobs or segments): added into a zeroed output
    p.atomic = (p.nsplit > 1 || NC > p.seg || force_atomic) ? 1u : 0u;
    if (p.atomic && zero_out) ISLE_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)num_mtiles * kHeadTile * 16 * sizeof(float), stream));
    const uint64_t kpad = (uint64_t)NC * kHeadChunk;
    switch (head_block_stride(b)) {
    case 5: launch_head_t<5>(c, split, kpad, p, stream); break;
    case 10: launch_head_t<10>(c, split, kpad, p, stream); break;
    default: launch_head_t<16>(c, split, kpad, p, stream); break;
    }
}

}  // namespace isle
