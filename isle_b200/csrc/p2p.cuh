// p2p.cuh -- device side of the collectives over peer memory (coll.cu): the workspace layout, the release / acquire flags
// between GPUs, and the two-shot all-reduce body, templated on a PRODUCER so that the kernel that creates the data can be
// the kernel that reduces it (spmm.cu: the un-permutation / scaling of the operator output and its all-reduce over NVLink
// are one launch).  See coll.cu for the protocol.
#pragma once

#include "common.cuh"

namespace isle {

constexpr int kMaxPeers = 16;
constexpr size_t kP2pFlagBytes = 4096;            // fa[16], fb[16] (u64, written by the peers), two local CTA counters
constexpr size_t kP2pRegion = (size_t)8 << 20;    // stage / result area of the two-shot form
constexpr size_t kP2pSlot = (size_t)1 << 20;      // payload capacity of a flag-in-data slot (all-gather pieces up to this size)
constexpr size_t kP2pOneShotMax = (size_t)256 << 10;   // one-shot all-reduce up to this size (its traffic grows with the rank count)
constexpr size_t kP2pOffStage = kP2pFlagBytes;
constexpr size_t kP2pOffResult = kP2pOffStage + kP2pRegion;
constexpr size_t kP2pOffSlots = kP2pOffResult + kP2pRegion;                 // [parity 2][src 16][2 * kP2pSlot]
constexpr size_t kP2pBytes = kP2pOffSlots + 2 * kMaxPeers * (2 * kP2pSlot);   // flag-in-data words: 8 bytes per 4 of payload

struct P2pArgs {
    char *ws[kMaxPeers];
    int me, world;
    unsigned long long epoch;
    unsigned tgt_a, tgt_b;
    uint32_t *diag;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Back-off + time limit of every wait on a peer.  Ranks reach a collective seconds apart when their host sides differ
// (corpus generation, rank 0's extra measurements), so the limit is generous: two minutes on %globaltimer, after which the
// kernel records what it was waiting for and traps (a peer that never arrives fails the launch instead of hanging the GPU).
constexpr unsigned long long kP2pTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;
struct P2pSpin {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ bool expired()
    {
        ++spins;
        if (spins <= 64) return false;
        __nanosleep(spins > 4096 ? 1000 : 50);
        if ((spins & 1023u) != 0) return false;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) { t0 = t; return false; }
        return t - t0 > kP2pTimeoutNs;
    }
};
// bounded wait for `*flag >= epoch`
__device__ __forceinline__ void p2p_wait(const unsigned long long *flag, unsigned long long epoch, uint32_t *diag, int what, int peer)
{
    P2pSpin sp;
    for (;;) {
        if (ld_acquire_sys(flag) >= epoch) return;
        if (sp.expired()) {
            if (diag) { diag[0] = 0xDEAD0000u | (uint32_t)what; diag[1] = (uint32_t)peer; diag[2] = (uint32_t)epoch; __threadfence_system(); }
            __trap();
        }
    }
}
// all CTAs of this launch have finished the preceding phase -> tell every rank (flag index `which`: 0 = fa, 1 = fb)
__device__ __forceinline__ void p2p_signal(const P2pArgs &a, int which, unsigned target)
{
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                      // this CTA's stores (local and peer) are visible system-wide ...
        unsigned *cnt = reinterpret_cast<unsigned *>(a.ws[a.me] + 2 * kMaxPeers * 8) + which;
        s_last = atomicAdd(cnt, 1u) + 1u == target;  // ... before it is counted
    }
    __syncthreads();
    // the last CTA to arrive tells every rank, one thread per rank so that the NVLink round trips overlap (one thread
    // issuing `world` release stores in a row paid ~2.5 us for each)
    if (s_last && (int)threadIdx.x < a.world) {
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long *>(a.ws[threadIdx.x]) + which * kMaxPeers + a.me),
                     "l"(a.epoch)
                     : "memory");
    }
}
__device__ __forceinline__ void p2p_wait_all(const P2pArgs &a, int which)
{
    if ((int)threadIdx.x < a.world)
        p2p_wait(reinterpret_cast<const unsigned long long *>(a.ws[a.me]) + which * kMaxPeers + threadIdx.x, a.epoch, a.diag, which,
                 (int)threadIdx.x);
    __syncthreads();
}

template <class T, int OP> struct Vec16;     // 16 bytes of T with the reduction OP (0 sum, 1 max)
template <class T, int OP> struct Vec16 {
    static constexpr int N = 16 / sizeof(T);
    alignas(16) T v[N];
    __device__ __forceinline__ static Vec16 load_cv(const void *p)
    {
        Vec16 r;
        const uint4 u = __ldcv(reinterpret_cast<const uint4 *>(p));
        *reinterpret_cast<uint4 *>(r.v) = u;
        return r;
    }
    __device__ __forceinline__ static Vec16 load(const void *p)
    {
        Vec16 r;
        *reinterpret_cast<uint4 *>(r.v) = *reinterpret_cast<const uint4 *>(p);
        return r;
    }
    __device__ __forceinline__ void store(void *p) const { *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(v); }
    __device__ __forceinline__ void acc(const Vec16 &o)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = OP == 0 ? (T)(v[i] + o.v[i]) : (v[i] > o.v[i] ? v[i] : o.v[i]);
    }
};

// buf (n elements, 16-byte aligned, nv = ceil(n * sizeof(T) / 16) vectors; the pad lanes of the last vector are zero in
// every stage because the stage copy writes them so) -> element-wise reduction over the ranks, in place
// producer that stages a vector that already sits in device memory
template <class T, int OP>
struct P2pCopyProducer {
    const T *src;
    __device__ __forceinline__ void stage(char *stage_area, size_t n, size_t tid, size_t nth) const
    {
        typedef Vec16<T, OP> V;
        const size_t nv = (n * sizeof(T) + 15) / 16;
        const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#pragma unroll 4
        for (size_t i = tid; i < nv; i += nth) {
            V x;
            if (al && (i + 1) * V::N <= n) x = V::load(src + i * V::N);
            else {
#pragma unroll
                for (int j = 0; j < V::N; ++j) x.v[j] = (i * V::N + j < n) ? src[i * V::N + j] : (T)0;
            }
            x.store(stage_area + i * 16);
        }
    }
};

// the two-shot all-reduce, called by every thread of a co-resident grid of 512-thread CTAs.  prod.stage() must write all
// ceil(n sizeof(T) / 16) 16-byte vectors of the stage area (pad lanes zero); the reduced vector lands in buf.
template <class T, int OP, class Producer>
__device__ __forceinline__ void p2p_allreduce2_body(const P2pArgs &a, const Producer &prod, T *__restrict__ buf, size_t n)
{
    typedef Vec16<T, OP> V;
    const size_t nv = (n * sizeof(T) + 15) / 16;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    char *mine = a.ws[a.me];
    const bool al = (reinterpret_cast<uintptr_t>(buf) & 15) == 0;
    // phase A: stage my vector where the peers can read it
    prod.stage(mine + kP2pOffStage, n, tid, nth);
    p2p_signal(a, 0, a.tgt_a);
    p2p_wait_all(a, 0);
    // phase B: my slice of the vectors, summed in rank order, stored into every rank's result area; the loads from all
    // ranks are issued together (NVLink latency is paid once per vector, not once per rank)
    const size_t per = (nv + a.world - 1) / a.world;
    const size_t lo = per * a.me < nv ? per * a.me : nv, hi = lo + per < nv ? lo + per : nv;
    for (size_t i = lo + tid; i < hi; i += nth) {
        V x[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) x[r] = V::load_cv(a.ws[r] + kP2pOffStage + i * 16);
#pragma unroll
        for (int r = 1; r < kMaxPeers; ++r)
            if (r < a.world) x[0].acc(x[r]);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) x[0].store(a.ws[r] + kP2pOffResult + i * 16);
    }
    p2p_signal(a, 1, a.tgt_b);
    p2p_wait_all(a, 1);
    // phase C: copy out
#pragma unroll 4
    for (size_t i = tid; i < nv; i += nth) {
        const V x = V::load_cv(mine + kP2pOffResult + i * 16);
        if (al && (i + 1) * V::N <= n) x.store(buf + i * V::N);
        else
            for (int j = 0; j < V::N; ++j)
                if (i * V::N + j < n) buf[i * V::N + j] = x.v[j];
    }
}


// host side (coll.cu): true when the peer workspace is mapped and `bytes` fits the two-shot form; fills the launch
// arguments (a new epoch, the grid-level arrival targets) and the co-resident grid size
bool p2p_two_shot_begin(Ctx &c, size_t bytes, P2pArgs *args, unsigned *grid);

}  // namespace isle
