// spmm_head_i8.cu -- kernel family (2), dense-head engine of the B * B^T * X operator, int8 form.
//
// Same job as spmm_head.cu (the H most frequent words of B kept as bitmaps; both passes of
// MKL_SpSpTrProd::multiply, reference include/matUtils.h:336-365, become
//     out[m, :] = sum_k bit(m, k) * In[k, :]                m: doc (pass 1) / head word (pass 2)
// on tcgen05), but with exact integer arithmetic and a quarter of the bit-expansion work:
//
//   A operand   u8 cells.  The bitmap is laid out so that NO per-cell shifting is needed: a 256-k
//               "super-chunk" of a row is 8 words; bit j of byte b of word w is k = 32 j + 4 w + b.  The
//               register that feeds TMEM column 8 j + w is simply  (m_w >> 4 [j >= 4]) & (0x01010101 << (j & 3)):
//               ONE logic instruction yields FOUR cells (spmm_head.cu: one shift + one and per TWO bf16
//               cells).  A set cell of bit-plane p = j & 3 carries the value 2^p instead of 1, so the eight
//               MMAs of a super-chunk (one per j, K = 32 cells each) accumulate into the accumulator of their
//               plane and the epilogue adds the four planes as acc_p >> p (exact: acc_p is a multiple of 2^p).
//   B operand   the dense operand quantised per column to 30 bits (q = rn(x 2^(29 - E)), 2^E <= column max <
//               2^(E+1): every entry within 2^-6 of its column maximum is represented EXACTLY, smaller ones to
//               2^-30 of the maximum) and split into four balanced base-256 digits, s8, K-major, staged by TMA
//               (128-byte swizzle); 40 rows: columns 0-3 and 4-7 as 4 x 4 blocks, columns 8-9 as a 2 x 4 block.
//   D           s32 in TMEM (tcgen05.mma.kind::i8): integer accumulation is exact, so there is no bound on
//               the length of an accumulation chain other than overflow (segments of <= 65536 cells).
//
// Per super-chunk (128 rows x 256 k) the three busy units are balanced: 72 ALU instructions per thread
// (144 cycles on the half-rate logic pipe), 32 KB of tcgen05.st (128 cycles at 256 B/clk), 8 MMAs of
// M = 128, N = 32, K = 32 (16 cycles each = 128 cycles).  The CTA is lean on purpose -- two worker groups
// of 4 warps + one MMA thread + one TMA thread, ~100 registers -- so that the L1TEX-bound tail gather
// (spmm.cu) keeps half of each SM's warps and registers beside it.
//
// Protocol (persistent, one CTA per SM, bounded waits that record where they stopped and trap):
//   warp 8 lane 0   TMA: per super-chunk two [32 x 128 B] digit tiles + the 4 KB bit tile
//   warp 9 lane 0   MMA: waits (tile landed, A stage stored), issues 8 MMAs, commits stage / A stage
//   warps 0-7       workers: group g = warp / 4 expands the super-chunks of parity g; super-chunk n goes to A stage
//                   n % 4 (TMEM columns 256 + 64 (n % 4)), so a group fills its second stage while the MMAs of its
//                   first are in flight; segment n accumulates in accumulator set n & 1 and is
//                   drained (tcgen05.ld, plane sum, digit recombination, store / atomic add) by group n & 1
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace isle {

namespace {

using namespace tcptx;

constexpr uint32_t kG = 2;                       // worker groups = accumulator sets
constexpr uint32_t kP = 2;                       // bit planes kept in separate accumulators (cell values 1, 2)
constexpr uint32_t kN = 48;                      // UMMA N (multiple of 16); the digit matrix has kHead8Rows = 40 rows (head8_row), the
                                                 // MMA's rows 40..47 read whatever follows the tile and land in accumulator columns nobody reads
constexpr uint32_t kPS = 64;                     // TMEM column stride of one plane accumulator (48 used; power-of-two aligned)
constexpr uint32_t kAccCols = kG * kP * kPS;     // accumulator sets at kP * kPS * set
constexpr uint32_t kA = (512 - kAccCols) / 64;   // A stages in TMEM behind them: super-chunk n uses stage n % kA, expanded by group n % 2
static_assert(kA % kG == 0, "an A stage must always be filled by the same worker group (parity waits skip no phase)");
constexpr uint32_t kWarpMma = 4 * kG, kWarpTma = 4 * kG + 1;
constexpr int kThreads = 32 * (4 * kG + 2);
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kBoxBytes = kHead8Rows * 128; // one swizzled [40 rows x 128 k] s8 tile (5 swizzle atoms of 8 rows)
constexpr uint32_t kScBytes = 2 * kBoxBytes;     // digits of one 256-k super-chunk
constexpr uint32_t kBitBytes = kHeadTile * 32;   // 128 rows x 256 bits
constexpr uint32_t kSpinLimit = 1u << 17;

__device__ uint32_t *g_head8_diag = nullptr;

__device__ __noinline__ void head8_timeout(uint32_t code, uint32_t a, uint32_t b)
{
    uint32_t *d = g_head8_diag;
    if (d) {
        const uint32_t cls = code & 0xFFu;
        const uint32_t slot = cls == 0x10 ? 0 : cls == 0x20 ? 1 : cls == 0x21 ? 2 : (cls & 0xF0u) == 0x30 ? 3 + (cls & 7u) : 11;
        uint32_t *r = d + slot * 5;
        if (atomicCAS_system(r, 0u, 1u + (0x2000u | code)) == 0u) {     // 0x2000: spmm_head_i8_kernel
            r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = a; r[4] = b;
        }
        __threadfence_system();
        for (int i = 0; i < 4000; ++i) __nanosleep(1000);
    }
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t code, uint32_t da, uint32_t db)
{
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spins = 0; spins < kSpinLimit; ++spins)
        if (mbar_try_wait(addr, parity)) return;
    head8_timeout(code, da, db);
}
__device__ __forceinline__ void mbar_wait2(uint64_t *bar_a, uint32_t parity_a, uint64_t *bar_b, uint32_t parity_b, uint32_t code,
                                           uint32_t da, uint32_t db)
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
    head8_timeout(code + (pa ? 0x100u : 0u), da, db);
}

// D[tmem] (s32) (+)= A[tmem] (u8 cells) * B[smem] (s8 digits)
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Head8Params {
    const uint4 *bits;     // [((mtile * NSC + sc) * 128 + row) * 2 + half] : 256 k of one row
    float *out;            // [rows][16] fp32
    const uint32_t *colmax_bits;   // [16] per-column max |In| (bit patterns) the digits were quantised with; NULL: unit scale
    uint32_t m_valid;      // rows of `out` that exist
    uint32_t num_mtiles;   // 128-row tiles
    uint32_t NSC;          // 256-k super-chunks along K
    uint32_t slab;         // resident mode: super-chunks of the digit operand kept in shared memory per K part; 0: digits stream with the bits
    uint32_t nslabs;       // K parts; job = (part, mtile), part-major, one accumulation segment each (<= 256 super-chunks:
                           // int32 / fp32-exactness bound)
    uint32_t stages;       // bit-tile ring
    uint32_t atomic;       // 1: jobs are added into `out` (pre-zeroed); 0: one job per row, stored
    int b;                 // columns in use (<= 10)
    unsigned long long *trace;   // debugging: clock64 stamps of CTA 0's first super-chunks (NULL: off)
};

constexpr uint32_t kTraceChunks = 96, kTraceSlots = 8;
#define HEAD8_TRACE(role_slot, n)                                                                         \
    do {                                                                                                  \
        if (p.trace && blockIdx.x == 0 && (n) < kTraceChunks) p.trace[(n) * kTraceSlots + (role_slot)] = clock64(); \
    } while (0)

// (min blocks = 2 only caps the registers at ~100 per thread: the tail gather shares the SM)
//
// Shared memory: the digit operand of one K slab (p.slab super-chunks x 8 KB) stays resident and is loaded once per
// slab; only the 4 KB bit tiles stream through the ring.  Every CTA works on a contiguous range of the slab-major job
// list, so it changes slab at most a couple of times.  (Re-loading the digit tile with every bit tile, as the first
// version did, tripled the L2 -> SM traffic of the kernel, which is the resource the tail gather is bound by.)
__global__ void __launch_bounds__(kThreads, 2)
spmm_head_i8_kernel(const __grid_constant__ CUtensorMap map_b, const Head8Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool resident = p.slab != 0;
    const uint32_t stage_bytes = resident ? kBitBytes : kScBytes + kBitBytes;     // streaming: [digits 8 KB | bits 4 KB] per stage
    const uint32_t bits_off = resident ? 0u : kScBytes;
    uint8_t *ring = smem + (size_t)p.slab * kScBytes;
    uint8_t *ctrl = ring + (size_t)p.stages * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);   // [stages]  bit tile landed
    uint64_t *empty = full + p.stages;                      // [stages]  the owning group's 4 warps hold the bits in registers
    uint64_t *a_full = empty + p.stages;                    // [kA]      a group stored the A stage
    uint64_t *a_empty = a_full + kA;                        // [kA]      MMAs reading the A stage retired
    uint64_t *acc_full = a_empty + kA;                      // [kG]      job accumulated
    uint64_t *acc_empty = acc_full + kG;                    // [kG]      its group drained the accumulator set
    uint64_t *slab_full = acc_empty + kG;                   // [1]       digit slab landed
    uint64_t *slab_empty = slab_full + 1;                   // [1]       MMAs of the last job on the slab retired
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(slab_empty + 1);
    float *s_scale = reinterpret_cast<float *>(tmem_slot + 4);   // [16]

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t njobs = p.num_mtiles * p.nslabs;
    // this CTA's contiguous share of the slab-major job list
    const uint32_t j_lo = (uint32_t)((uint64_t)njobs * blockIdx.x / gridDim.x);
    const uint32_t j_hi = (uint32_t)((uint64_t)njobs * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], resident ? 4 : 5); }
        for (uint32_t a = 0; a < kA; ++a) { mbar_init(&a_full[a], 4); mbar_init(&a_empty[a], 1); }
        for (uint32_t a = 0; a < kG; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        mbar_init(slab_full, 1); mbar_init(slab_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (threadIdx.x < 16) s_scale[threadIdx.x] = quant_down(quant_exp(p.colmax_bits, (int)threadIdx.x));   // digit sums -> values
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
            uint32_t stage = 0, phase = 0, tn = 0, cur_slab = 0xFFFFFFFFu, sphase = 0;
            for (uint32_t job = j_lo; job < j_hi; ++job) {
                const uint32_t sl = job / p.num_mtiles, mtile = job % p.num_mtiles;
                const uint32_t c0 = (uint32_t)((uint64_t)p.NSC * sl / p.nslabs), c1 = (uint32_t)((uint64_t)p.NSC * (sl + 1) / p.nslabs);
                if (resident && sl != cur_slab) {
                    if (cur_slab != 0xFFFFFFFFu) { mbar_wait(slab_empty, sphase, 0x11, job, sl); sphase ^= 1; }
                    mbar_expect_tx(slab_full, (c1 - c0) * kScBytes);
                    for (uint32_t sc = c0; sc < c1; ++sc) {
                        uint8_t *dst = smem + (size_t)(sc - c0) * kScBytes;
                        tma_load_2d(dst, &map_b, slab_full, (int)(sc * 256), 0);
                        tma_load_2d(dst + kBoxBytes, &map_b, slab_full, (int)(sc * 256 + 128), 0);
                    }
                    cur_slab = sl;
                }
                for (uint32_t sc = c0; sc < c1; ++sc) {
                    mbar_wait(&empty[stage], phase ^ 1, 0x10, job, sc);
                    HEAD8_TRACE(0, tn); ++tn;
                    uint8_t *st = ring + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full[stage], stage_bytes);
                    if (!resident) {
                        tma_load_2d(st, &map_b, &full[stage], (int)(sc * 256), 0);
                        tma_load_2d(st + kBoxBytes, &map_b, &full[stage], (int)(sc * 256 + 128), 0);
                    }
                    bulk_load_1d(st + bits_off, p.bits + ((size_t)mtile * p.NSC + sc) * (kHeadTile * 2), kBitBytes, &full[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer (one thread).  idesc: D = s32 (2 << 4), A = u8 (0 << 7), B = s8 (1 << 10), K-major,
        // N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((kN >> 3) << 17) | ((uint32_t)(kHeadTile >> 4) << 24);
        if (elect_one_sync()) {
            const uint64_t desc0 = umma_desc(smem_u32(smem));
            const uint64_t ring0 = umma_desc(smem_u32(ring));
            uint32_t a = 0, aphase = 0, buf = 0, bphase = 0, tn = 0, cur_slab = 0xFFFFFFFFu, sphase = 0, stage = 0, phase = 0;
            for (uint32_t job = j_lo; job < j_hi; ++job) {
                const uint32_t sl = job / p.num_mtiles;
                const uint32_t c0 = (uint32_t)((uint64_t)p.NSC * sl / p.nslabs), c1 = (uint32_t)((uint64_t)p.NSC * (sl + 1) / p.nslabs);
                if (resident && sl != cur_slab) { mbar_wait(slab_full, sphase, 0x22, job, sl); sphase ^= 1; cur_slab = sl; }
                mbar_wait(&acc_empty[buf], bphase ^ 1, 0x20, job, c0);
                const uint32_t d_tmem = tmem_base + buf * (kP * kPS);
                for (uint32_t sc = c0; sc < c1; ++sc) {
                    if (resident) mbar_wait(&a_full[a], aphase, 0x21, job, sc);
                    else mbar_wait2(&full[stage], phase, &a_full[a], aphase, 0x21, job, sc);
                    HEAD8_TRACE(1, tn);
                    tc_fence_after();
                    const uint64_t bd = resident ? desc0 + (uint64_t)((sc - c0) * (kScBytes >> 4)) : ring0 + (uint64_t)(stage * (stage_bytes >> 4));
                    const uint32_t a_tmem = tmem_base + kAccCols + a * 64;
                    const uint32_t later = sc > c0 ? 1u : 0u;
#pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) {
                        // MMA j: k = 32 j .. 32 j + 31 of the super-chunk = TMEM columns 8 j .. 8 j + 7 of the A
                        // stage (cells worth 2^(j % kP)), 32 bytes along K of digit tile j >> 2
                        umma_i8_ts(d_tmem + (j & (kP - 1)) * kPS, a_tmem + j * 8,
                                   bd + (uint64_t)((j >> 2) * (kBoxBytes >> 4) + (j & 3) * 2), idesc, j >= kP ? 1u : later);
                    }
                    if (!resident) { umma_commit(&empty[stage]); if (++stage == p.stages) { stage = 0; phase ^= 1; } }
                    umma_commit(&a_empty[a]);
                    HEAD8_TRACE(2, tn); ++tn;
                    if (++a == kA) { a = 0; aphase ^= 1; }
                }
                umma_commit(&acc_full[buf]);
                if (++buf == kG) { buf = 0; bphase ^= 1; }
                // last job on this slab: the producer may overwrite the digits once these MMAs have retired
                if (resident && job + 1 < j_hi && (job + 1) / p.num_mtiles != sl) umma_commit(slab_empty);
            }
        }
    } else {
        // ===== workers: group = warp / 4, lane quarter = warp % 4; thread <-> row of the tile <-> TMEM lane
        const uint32_t grp = warp >> 2, quarter = warp & 3;
        const uint32_t trow = quarter * 32 + lane;
        const uint32_t lane_base = (quarter * 32u) << 16;
        uint32_t stage = 0, phase = 0, a = 0, aphase = 0, buf = 0, bphase = 0, tn = 0;
        const bool tracer = quarter == 0 && lane == 0;
        bool pend = false;
        uint32_t pend_buf = 0, pend_phase = 0, pend_row = 0;

        auto drain = [&]() {
            mbar_wait(&acc_full[pend_buf], pend_phase, 0x30 + grp, pend_row, pend_buf * 2 + pend_phase);
            tc_fence_after();
            // digit g of column c sits at n = head8_row(c, g): columns 0-3 in accumulator columns [0, 16), 4-7 in
            // [16, 32) (n = 16 (c / 4) + 4 g + c % 4), columns 8, 9 in [32, 40) (n = 32 + 2 g + c - 8)
            float v[16];
#pragma unroll
            for (int c = 10; c < 16; ++c) v[c] = 0.0f;
#pragma unroll
            for (uint32_t h = 0; h < 3; ++h) {
                int S[16];
#pragma unroll
                for (uint32_t n = 0; n < 16; ++n) S[n] = 0;
#pragma unroll
                for (uint32_t pl = 0; pl < kP; ++pl) {
                    uint32_t t[16];
                    tmem_ld16(tmem_base + lane_base + pend_buf * (kP * kPS) + pl * kPS + 16 * h, t);
                    tmem_ld_wait();
#pragma unroll
                    for (uint32_t n = 0; n < 16; ++n) S[n] += (int)t[n] >> pl;      // exact: plane sums are multiples of 2^pl
                }
                // q = d0 + 2^8 d1 + 2^16 d2 + 2^24 d3 summed per digit; |S| < 2^24 per job, so the conversions are exact
                const int w = h < 2 ? 4 : 2;      // columns in this group
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = 4 * (int)h + cc;
                    if (cc < w) {
                        const float t = fmaf((float)S[3 * w + cc], 16777216.0f,
                                             fmaf((float)S[2 * w + cc], 65536.0f, fmaf((float)S[w + cc], 256.0f, (float)S[cc])));
                        v[c] = c < p.b ? t * s_scale[c] : 0.0f;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[pend_buf]);
            if (pend_row < p.m_valid) {
                float4 *dst = reinterpret_cast<float4 *>(p.out) + (size_t)pend_row * 4;
                if (p.atomic) {
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        if (4 * q < p.b) atomicAdd(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            pend = false;
        };

        for (uint32_t job = j_lo; job < j_hi; ++job) {
            const uint32_t sl = job / p.num_mtiles, mtile = job % p.num_mtiles;
            const uint32_t c0 = (uint32_t)((uint64_t)p.NSC * sl / p.nslabs), c1 = (uint32_t)((uint64_t)p.NSC * (sl + 1) / p.nslabs);
            for (uint32_t sc = c0; sc < c1; ++sc) {
                if ((tn & (kG - 1)) == grp) {
                    mbar_wait2(&full[stage], phase, &a_empty[a], aphase ^ 1, 0x34 + grp, job, sc);
                    if (tracer) HEAD8_TRACE(3, tn);
                    const uint4 *bt = reinterpret_cast<const uint4 *>(ring + (size_t)stage * stage_bytes + bits_off) + trow * 2;
                    const uint4 lo = bt[0], hi = bt[1];
                    const uint32_t m[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    tc_fence_after();
                    const uint32_t a_addr = tmem_base + lane_base + kAccCols + a * 64;
                    // columns 8 j + w, j = bit, w = word: plane j % kP, the higher bit groups through one shift per word
#pragma unroll
                    for (uint32_t q = 0; q < 4; ++q) {
                        uint32_t r[16];
#pragma unroll
                        for (uint32_t jj = 0; jj < 2; ++jj) {
                            const uint32_t j = 2 * q + jj;
                            const uint32_t mask = 0x01010101u << (j & (kP - 1));
                            const uint32_t sh = j & ~(kP - 1);
#pragma unroll
                            for (uint32_t w = 0; w < 8; ++w) r[jj * 8 + w] = (m[w] >> sh) & mask;
                        }
                        tmem_st16(a_addr + 16 * q, r);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);       // bits are in registers
                    if (tracer) HEAD8_TRACE(4, tn);
                    tmem_st_wait();
                    if (tracer) HEAD8_TRACE(5, tn);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[a]);
                    if (pend) { drain(); if (tracer) HEAD8_TRACE(6, tn); }     // a finished job of this group drains while later super-chunks multiply
                }
                ++tn;
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                if (++a == kA) { a = 0; aphase ^= 1; }
            }
            if (buf == grp) {      // job n uses accumulator set n mod kG and is drained by group n mod kG
                if (pend) drain();
                pend = true;
                pend_buf = buf;
                pend_phase = bphase;
                pend_row = mtile * kHeadTile + trow;
            }
            if (++buf == kG) { buf = 0; bphase ^= 1; }
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

// [32 x kpad] s8 row-major (K contiguous), box = 128 k x 32 rows, 128-byte swizzle
CUtensorMap make_digit_map(const int8_t *base, uint64_t kpad)
{
    CUtensorMap m;
    const cuuint64_t gdim[2] = {kpad, (cuuint64_t)kHead8Rows};
    const cuuint64_t gstride[1] = {kpad};
    const cuuint32_t box[2] = {128, (cuuint32_t)kHead8Rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ISLE_REQUIRE(r == CUDA_SUCCESS, ISLE_ERR_CUDA, "cuTensorMapEncodeTiled (int8 head operand) failed (" + std::to_string((int)r) + ")");
    return m;
}

}  // namespace

// out[m, 0:b] (+)= sum_k bit(m, k) * In[k, 0:b]; In given as its three s8 digits [32][NSC * 256] K-major (digit g of
// column c in row head8_row(c, g); rows of columns >= b may hold anything), quantised with the column maxima
// colmax_bits (quant_exp).
// zero_out: `out` (num_mtiles * 128 rows) is cleared first when partial sums are added atomically.
void spmm_head_i8_launch(Ctx &c, int b, const uint4 *bits, uint32_t num_mtiles, uint32_t NSC, uint32_t nsplit,
                         const int8_t *digits, const uint32_t *colmax_bits, float *out, uint32_t m_valid, bool zero_out,
                         bool force_atomic, cudaStream_t stream)
{
    // resident mode (option spmm_head8_slab > 0): the digit operand of a K part stays in shared memory and only the bit tiles
    // stream (a third of the L2 -> SM traffic, 64 KB more shared memory); default: digits stream with the bits, K cut into
    // `nsplit` parts per tile.  Either way the part-major job list is balanced over the CTAs.
    if (!num_mtiles || !NSC) return;
    ISLE_REQUIRE(b >= 1 && b <= 10, ISLE_ERR_ARG, "spmm_head_i8: block size must be in [1,10] (N = 3 digits x 10 columns)");
    Head8Params p;
    p.bits = bits; p.out = out; p.colmax_bits = colmax_bits; p.m_valid = m_valid; p.num_mtiles = num_mtiles; p.NSC = NSC;
    p.slab = std::min<uint32_t>(NSC, (uint32_t)std::max(0, std::min(16, c.opt("spmm_head8_slab", 0))));
    p.nslabs = p.slab ? (NSC + p.slab - 1) / p.slab : std::max<uint32_t>(std::max<uint32_t>(1, std::min(nsplit, NSC)), (NSC + 255) / 256);
    p.b = b;
    // ring length: a multiple of the group count, so that a ring stage is always consumed by the same worker group (a
    // parity wait cannot tell "phase n done" from "phase n - 2 done"; a group must see every phase of the barriers it
    // waits on).  4 stages x 14 KB stay under the 64 KB shared-memory carve-out, which leaves the L1 the tail gather
    // that shares the SM lives on.
    p.stages = (uint32_t)std::max(2, std::min(16, c.opt("spmm_head8_stages", 4))) & ~(kG - 1);
    // rows receive several partial sums (K cut into slabs): added into a zeroed output
    p.atomic = (p.nslabs > 1 || force_atomic) ? 1u : 0u;
    if (p.atomic && zero_out) ISLE_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)num_mtiles * kHeadTile * 16 * sizeof(float), stream));
    const uint32_t smem_bytes = p.slab * kScBytes + p.stages * (p.slab ? kBitBytes : kScBytes + kBitBytes) + 1024 + 512;
    ISLE_CUDA_CHECK(cudaFuncSetAttribute(spmm_head_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (c.head_diag_dev) {
        static thread_local uint32_t *installed = nullptr;     // device-side symbol, set once per host thread (= per device) and pointer
        if (installed != c.head_diag_dev) {
            ISLE_CUDA_CHECK(cudaMemcpyToSymbolAsync(g_head8_diag, &c.head_diag_dev, sizeof(uint32_t *), 0, cudaMemcpyHostToDevice, stream));
            installed = c.head_diag_dev;
        }
    }
    const CUtensorMap map = make_digit_map(digits, (uint64_t)NSC * 256);
    static_assert(kBoxBytes % 1024 == 0 && (2 * kBoxBytes + kBitBytes) % 1024 == 0, "swizzled tiles must stay 1024-byte aligned");
    const uint32_t njobs = p.num_mtiles * p.nslabs;
    const unsigned grid = std::min<uint32_t>(njobs, (uint32_t)c.num_sms);
    // debugging aid: ISLE_HEAD8_TRACE=<file> dumps clock stamps of CTA 0's first super-chunks (synchronises)
    const char *trace_path = std::getenv("ISLE_HEAD8_TRACE");
    DevBuf<unsigned long long> trace;
    p.trace = nullptr;
    if (trace_path) {
        trace.alloc((size_t)kTraceChunks * kTraceSlots);
        ISLE_CUDA_CHECK(cudaMemsetAsync(trace.p, 0, trace.bytes(), stream));
        p.trace = trace.p;
    }
    spmm_head_i8_kernel<<<grid, kThreads, smem_bytes, stream>>>(map, p);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
    if (trace_path) {
        std::vector<unsigned long long> h((size_t)kTraceChunks * kTraceSlots);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(h.data(), trace.p, trace.bytes(), cudaMemcpyDeviceToHost, stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (FILE *f = std::fopen(trace_path, "a")) {
            std::fprintf(f, "# launch mtiles=%u NSC=%u slab=%u nslabs=%u\n", p.num_mtiles, p.NSC, p.slab, p.nslabs);
            for (uint32_t n = 0; n < kTraceChunks; ++n) {
                for (uint32_t s = 0; s < kTraceSlots; ++s) std::fprintf(f, "%llu ", h[(size_t)n * kTraceSlots + s]);
                std::fprintf(f, "\n");
            }
            std::fclose(f);
        }
    }
}

}  // namespace isle
