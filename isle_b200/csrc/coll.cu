// coll.cu -- the collectives of the document-sharded path over NVLink 5 / NVSwitch.
//
// Every exchange of the path is a small, latency-bound message issued on the context stream directly behind the kernel
// that produced the buffer (the V x b operator block, 4 MB; k x kp center sums; 16 x 16 Gram matrices; a few counters),
// 40 - 7000 of them per step.  Those run as ONE kernel each over peer memory (p2p_* below): every rank keeps a workspace
// that all other ranks map (cudaIpc between processes, peer access inside one process) and the kernel stages, reduces and
// redistributes with plain loads / stores through NVLink, synchronising the GPUs with release / acquire flags in the same
// workspace -- no proxy thread, no ring, one launch.  Sums are formed in rank order by exactly one rank per element and
// broadcast, so every rank receives bit-identical results (the replicated parts of the solver stay in lock step).
//   <= 256 KB   one-shot: push my vector into slot [me] of every peer, one barrier, everybody adds the slots in rank order
//   <= 8 MB     two-shot: stage locally, barrier, rank r reduces slice r reading every peer's stage and stores the sums
//               into every peer's result area, barrier, copy out
//   larger      NCCL (the bins x V threshold histogram: bandwidth-bound, once per step); so is everything when the
//               workspace cannot be mapped (option p2p = 0, no peer access), and broadcast / the unique-id bootstrap.
// With world == 1 all collectives are no-ops.
#include <unistd.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "p2p.cuh"

namespace isle {

#ifdef ISLE_WITH_NCCL
#define ISLE_NCCL_CHECK(expr)                                                                       \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess)                                                                      \
            throw ::isle::Error(ISLE_ERR_CUDA, std::string(#expr) + ": " + ncclGetErrorString(_r)); \
    } while (0)


// ---------------------------------------------------------------------------------------------------------------
// peer-to-peer collectives
// ---------------------------------------------------------------------------------------------------------------
template <class T, int OP>
__global__ void __launch_bounds__(512)
p2p_allreduce2_kernel(P2pArgs a, T *__restrict__ buf, size_t n)
{
    p2p_allreduce2_body<T, OP>(a, P2pCopyProducer<T, OP>{buf}, buf, n);
}

// One-shot form, flag-in-data (the "LL" idea): every 4-byte word of the payload travels as one 8-byte store {word, epoch},
// which NVLink delivers atomically, so the receiver needs no separate flag and the sender no fence: a thread pushes its
// element into slot [me] of every rank, then polls the same element of every slot of its own rank until each carries
// this collective's epoch, and adds them in rank order.  No CTA or grid synchronisation at all; one element per thread.
// Slots alternate with the parity of the epoch (a rank can be at most one collective ahead of another).
__device__ __forceinline__ void st_ll(void *p, uint32_t v, uint32_t flag)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const void *p)
{
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
template <class T, int OP>
__global__ void __launch_bounds__(256)
p2p_allreduce1_kernel(P2pArgs a, T *__restrict__ buf, size_t n)
{
    constexpr int W = sizeof(T) / 4;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t flag = (uint32_t)a.epoch;
    const size_t slots = kP2pOffSlots + ((size_t)(a.epoch & 1) * kMaxPeers) * (2 * kP2pSlot);
    union { T t; uint32_t w[W]; } mine, got;
    mine.t = buf[e];
#pragma unroll
    for (int j = 0; j < W; ++j)
        for (int r = 0; r < a.world; ++r) st_ll(a.ws[r] + slots + (size_t)a.me * (2 * kP2pSlot) + (e * W + j) * 8, mine.w[j], flag);
    const char *in = a.ws[a.me] + slots + e * W * 8;
    uint32_t w[W][kMaxPeers];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        unsigned pending = (1u << a.world) - 1u;
        P2pSpin sp;
        while (pending) {
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (pending >> r & 1u) {
                    const uint2 x = ld_ll(in + (size_t)r * (2 * kP2pSlot) + j * 8);
                    if (x.y == flag) { w[j][r] = x.x; pending &= ~(1u << r); }
                }
            if (pending && sp.expired()) {
                if (a.diag) { a.diag[0] = 0xDEAD0002u; a.diag[1] = pending; a.diag[2] = flag; __threadfence_system(); }
                __trap();
            }
        }
    }
    T s;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if (r < a.world) {
#pragma unroll
            for (int j = 0; j < W; ++j) got.w[j] = w[j][r];
            s = r == 0 ? got.t : (OP == 0 ? (T)(s + got.t) : (s > got.t ? s : got.t));
        }
    buf[e] = s;
}

// all-gather of small pieces (count * sizeof(T) <= kP2pSlot), flag-in-data like the one-shot all-reduce: no barrier
template <class T>
__global__ void __launch_bounds__(256)
p2p_allgather1_kernel(P2pArgs a, const T *__restrict__ send, T *__restrict__ recv, size_t count)
{
    constexpr int W = sizeof(T) / 4;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const uint32_t flag = (uint32_t)a.epoch;
    const size_t slots = kP2pOffSlots + ((size_t)(a.epoch & 1) * kMaxPeers) * (2 * kP2pSlot);
    union { T t; uint32_t w[W]; } x;
    x.t = send[e];
#pragma unroll
    for (int j = 0; j < W; ++j)
        for (int r = 0; r < a.world; ++r) st_ll(a.ws[r] + slots + (size_t)a.me * (2 * kP2pSlot) + (e * W + j) * 8, x.w[j], flag);
    const char *in = a.ws[a.me] + slots + e * W * 8;
    for (int r = 0; r < a.world; ++r) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            P2pSpin sp;
            for (;;) {
                const uint2 y = ld_ll(in + (size_t)r * (2 * kP2pSlot) + j * 8);
                if (y.y == flag) { x.w[j] = y.x; break; }
                if (sp.expired()) {
                    if (a.diag) { a.diag[0] = 0xDEAD0003u; a.diag[1] = (uint32_t)r; a.diag[2] = flag; __threadfence_system(); }
                    __trap();
                }
            }
        }
        recv[(size_t)r * count + e] = x.t;
    }
}

// recv[r * count .. ) = send of rank r  (count floats per rank, world * count * 4 <= kP2pRegion): every rank stores its
// piece into every rank's result area, one barrier, copy out
template <class T>
__global__ void __launch_bounds__(256)
p2p_allgather_kernel(P2pArgs a, const T *__restrict__ send, T *__restrict__ recv, size_t count)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    p2p_signal(a, 0, a.tgt_a);          // every rank has left the previous collective: its result area may be overwritten
    p2p_wait_all(a, 0);
    for (size_t i = tid; i < count; i += nth) {
        const T x = send[i];
        for (int r = 0; r < a.world; ++r) reinterpret_cast<T *>(a.ws[r] + kP2pOffResult)[(size_t)a.me * count + i] = x;
    }
    p2p_signal(a, 1, a.tgt_b);
    p2p_wait_all(a, 1);
    const T *res = reinterpret_cast<const T *>(a.ws[a.me] + kP2pOffResult);
    for (size_t i = tid; i < count * a.world; i += nth) recv[i] = __ldcv(res + i);
}


struct P2pRecord {                 // what the ranks tell each other about their workspace
    unsigned long long pid, ptr;
    int device, ok;
    cudaIpcMemHandle_t handle;
};

// Maps every rank's workspace.  Collective: called by all ranks at their first small collective.  The outcome is agreed
// on (all ranks or none), so the ranks never disagree about which transport carries a message.
static void p2p_init(Ctx &c)
{
    c.p2p_state = -1;
    int ok = c.world <= kMaxPeers && c.opt("p2p", 1) != 0;
    char *ws = nullptr;
    P2pRecord mine{};
    mine.pid = (unsigned long long)getpid();
    mine.device = c.device;
    if (ok && cudaMalloc((void **)&ws, kP2pBytes) != cudaSuccess) { cudaGetLastError(); ok = 0; ws = nullptr; }
    if (ok) {
        ISLE_CUDA_CHECK(cudaMemsetAsync(ws, 0, kP2pBytes, c.stream));
        mine.ptr = (unsigned long long)ws;
        if (cudaIpcGetMemHandle(&mine.handle, ws) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    mine.ok = ok;
    DevBuf<unsigned char> dsend(sizeof(P2pRecord)), drecv(sizeof(P2pRecord) * (size_t)c.world);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(dsend.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
    ISLE_NCCL_CHECK(ncclAllGather(dsend.p, drecv.p, sizeof(P2pRecord), ncclUint8, c.comm, c.stream));
    std::vector<P2pRecord> all((size_t)c.world);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(all.data(), drecv.p, sizeof(P2pRecord) * (size_t)c.world, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    for (int r = 0; r < c.world && ok; ++r) ok = all[r].ok;
    for (int r = 0; r < c.world && ok; ++r) {
        if (r == c.rank) { c.p2p_ws[r] = ws; continue; }
        if (all[r].pid == mine.pid) {          // same process (isle_cuda_create_multi): plain peer access
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c.device, all[r].device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
            cudaGetLastError();
            c.p2p_ws[r] = (char *)all[r].ptr;
        } else {
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            c.p2p_ws[r] = (char *)q;
            c.p2p_ipc[r] = true;
        }
    }
    // all or none
    DevBuf<int> flag(1);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c.stream));
    ISLE_NCCL_CHECK(ncclAllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, c.comm, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    if (!ok) {
        for (int r = 0; r < c.world; ++r) {
            if (c.p2p_ipc[r]) cudaIpcCloseMemHandle(c.p2p_ws[r]);
            c.p2p_ws[r] = nullptr;
            c.p2p_ipc[r] = false;
        }
        if (ws) cudaFree(ws);
        cudaGetLastError();
        c.counters["p2p_enabled"] = 0.0;
        return;
    }
    if (cudaHostAlloc((void **)&c.p2p_diag_host, 64, cudaHostAllocMapped) == cudaSuccess) {
        std::memset(c.p2p_diag_host, 0, 64);
        if (cudaHostGetDevicePointer((void **)&c.p2p_diag_dev, c.p2p_diag_host, 0) != cudaSuccess) c.p2p_diag_dev = nullptr;
    } else {
        cudaGetLastError();
    }
    c.p2p_state = 1;
    c.counters["p2p_enabled"] = 1.0;
}

static bool p2p_ready(Ctx &c)
{
    if (c.p2p_state == 0) p2p_init(c);
    return c.p2p_state == 1;
}

static P2pArgs p2p_args(Ctx &c, unsigned grid, bool uses_a, bool uses_b)
{
    P2pArgs a{};
    for (int r = 0; r < c.world; ++r) a.ws[r] = c.p2p_ws[r];
    a.me = c.rank;
    a.world = c.world;
    a.epoch = ++c.p2p_epoch;
    if (uses_a) c.p2p_cnt_a += grid;
    if (uses_b) c.p2p_cnt_b += grid;
    a.tgt_a = c.p2p_cnt_a;
    a.tgt_b = c.p2p_cnt_b;
    a.diag = c.p2p_diag_dev;
    return a;
}

bool p2p_two_shot_begin(Ctx &c, size_t bytes, P2pArgs *args, unsigned *grid)
{
    if (c.world <= 1 || bytes == 0 || bytes > kP2pRegion || !p2p_ready(c)) return false;
    // every CTA waits for all the others (grid-level phases): the grid must be co-resident, one CTA per SM at most
    *grid = (unsigned)std::min<size_t>((size_t)c.num_sms, std::max<size_t>(4, bytes / 16 / 1024));
    *args = p2p_args(c, *grid, true, true);
    c.counters["p2p_collectives"] += 1.0;
    return true;
}

template <class T, int OP>
static bool p2p_allreduce(Ctx &c, T *buf, size_t n)
{
    const size_t bytes = n * sizeof(T);
    if (bytes > kP2pRegion || !p2p_ready(c)) return false;
    if (bytes <= kP2pOneShotMax) {
        const P2pArgs a = p2p_args(c, 1, false, false);
        p2p_allreduce1_kernel<T, OP><<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(a, buf, n);
    } else {
        P2pArgs a;
        unsigned grid = 0;
        if (!p2p_two_shot_begin(c, bytes, &a, &grid)) return false;
        p2p_allreduce2_kernel<T, OP><<<grid, 512, 0, c.stream>>>(a, buf, n);
        ISLE_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        return true;
    }
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
    c.counters["p2p_collectives"] += 1.0;
    return true;
}

template <class T>
static bool p2p_allgather(Ctx &c, const T *send, T *recv, size_t count)
{
    const size_t bytes = count * sizeof(T) * (size_t)c.world;
    if (bytes > kP2pRegion || !p2p_ready(c)) return false;
    if (count * sizeof(T) <= kP2pSlot) {
        const P2pArgs a = p2p_args(c, 1, false, false);
        p2p_allgather1_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, c.stream>>>(a, send, recv, count);
        ISLE_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        c.counters["p2p_collectives"] += 1.0;
        return true;
    }
    const unsigned grid = (unsigned)std::min<size_t>(64, std::max<size_t>(1, bytes / 32768));
    const P2pArgs a = p2p_args(c, grid, true, true);
    p2p_allgather_kernel<T><<<grid, 256, 0, c.stream>>>(a, send, recv, count);
    ISLE_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
    c.counters["p2p_collectives"] += 1.0;
    return true;
}

void p2p_destroy(Ctx &c)
{
    if (c.p2p_state != 1) return;
    for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) cudaFree(c.p2p_ws[r]);
        else if (c.p2p_ipc[r]) cudaIpcCloseMemHandle(c.p2p_ws[r]);
        c.p2p_ws[r] = nullptr;
    }
    if (c.p2p_diag_host) cudaFreeHost(c.p2p_diag_host);
    c.p2p_diag_host = nullptr;
    cudaGetLastError();
    c.p2p_state = -1;
}


// ---- self-test of the peer-to-peer collectives against NCCL (harness: isle_cuda_selftest_collectives) -----------------
template <class T>
__global__ void fill_pattern_kernel(T *p, size_t n, unsigned rank, unsigned salt)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned h = (unsigned)i * 2654435761u + rank * 40503u + salt * 97u;
        h ^= h >> 15;
        p[i] = (T)(h & 1023u);          // small integers: sums over <= 16 ranks are exact in every type
    }
}
template <class T>
__global__ void count_diff_kernel(const T *a, const T *b, size_t n, unsigned long long *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long d = 0;
    for (; i < n; i += stride) d += a[i] != b[i];
    if (d) atomicAdd(out, d);
}

template <class T, int OP>
static void selftest_one(Ctx &c, size_t n, ncclDataType_t t, unsigned salt, unsigned long long *dmis)
{
    DevBuf<T> a(n + 4), b(n + 4);
    T *pa = a.p + (salt & 1), *pb = b.p + (salt & 1);        // odd salts: a misaligned buffer
    fill_pattern_kernel<T><<<grid_for(n, 256), 256, 0, c.stream>>>(pa, n, (unsigned)c.rank, salt);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(pb, pa, n * sizeof(T), cudaMemcpyDeviceToDevice, c.stream));
    if (!p2p_allreduce<T, OP>(c, pa, n)) throw Error(ISLE_ERR_CUDA, "selftest: the peer-to-peer path refused a message");
    ISLE_NCCL_CHECK(ncclAllReduce(pb, pb, n, t, OP == 0 ? ncclSum : ncclMax, c.comm, c.stream));
    count_diff_kernel<T><<<grid_for(n, 256), 256, 0, c.stream>>>(pa, pb, n, dmis);
}

// Runs every peer-to-peer collective on many sizes (one-shot / two-shot boundaries, odd lengths, misaligned buffers,
// back-to-back epochs) and compares with NCCL bit for bit.  *mismatches_out = differing elements (0 expected);
// *p2p_active_out = 1 when the peer workspace is mapped (0: NCCL carries everything and nothing was compared).
void selftest_collectives(Ctx &c, unsigned long long *mismatches_out, int *p2p_active_out)
{
    *mismatches_out = 0;
    *p2p_active_out = 0;
    if (c.world <= 1 || !p2p_ready(c)) return;
    *p2p_active_out = 1;
    DevBuf<unsigned long long> dmis(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(dmis.p, 0, 8, c.stream));
    const size_t sizes[] = {1, 3, 250, 4096, 16385, 65536, 65537, 100003, 1 << 20, (1 << 21) - 5};
    unsigned salt = 0;
    for (int rep = 0; rep < 3; ++rep)
        for (size_t n : sizes) {
            selftest_one<float, 0>(c, n, ncclFloat32, ++salt, dmis.p);
            selftest_one<uint32_t, 0>(c, n, ncclUint32, ++salt, dmis.p);
            selftest_one<uint32_t, 1>(c, n, ncclUint32, ++salt, dmis.p);
            if (n * 8 <= kP2pRegion) {
                selftest_one<double, 0>(c, n, ncclFloat64, ++salt, dmis.p);
                selftest_one<unsigned long long, 0>(c, n, ncclUint64, ++salt, dmis.p);
            }
        }
    // all-gathers
    for (size_t count : {(size_t)1, (size_t)77, (size_t)20000, std::min<size_t>(262144, kP2pRegion / 4 / c.world), (size_t)(kP2pRegion / 4 / c.world)}) {
        DevBuf<float> send(count), r1(count * c.world), r2(count * c.world);
        fill_pattern_kernel<float><<<grid_for(count, 256), 256, 0, c.stream>>>(send.p, count, (unsigned)c.rank, ++salt);
        if (!p2p_allgather(c, send.p, r1.p, count)) throw Error(ISLE_ERR_CUDA, "selftest: the peer-to-peer all-gather refused a message");
        ISLE_NCCL_CHECK(ncclAllGather(send.p, r2.p, count, ncclFloat32, c.comm, c.stream));
        count_diff_kernel<float><<<grid_for(count * c.world, 256), 256, 0, c.stream>>>(r1.p, r2.p, count * c.world, dmis.p);
    }
    ISLE_CUDA_CHECK(cudaMemcpyAsync(mismatches_out, dmis.p, 8, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    // latency of the message sizes the path uses (V x b operator block, k x kp center sums, a Gram matrix), both transports:
    // counters selftest_{p2p,nccl}_{4mb,50kb,2kb}_us, 20 back-to-back collectives each between two events
    cudaEvent_t e0, e1;
    ISLE_CUDA_CHECK(cudaEventCreate(&e0));
    ISLE_CUDA_CHECK(cudaEventCreate(&e1));
    const struct { const char *name; size_t n; } cases[] = {{"4mb", (size_t)1 << 20}, {"256kb", 65536}, {"50kb", 12800}, {"2kb", 512}};
    for (const auto &cs : cases) {
        DevBuf<float> x(cs.n);
        ISLE_CUDA_CHECK(cudaMemsetAsync(x.p, 0, cs.n * 4, c.stream));
        for (int which = 0; which < 2; ++which) {
            for (int it = 0; it < 23; ++it) {
                if (it == 3) ISLE_CUDA_CHECK(cudaEventRecord(e0, c.stream));
                if (which == 0) p2p_allreduce<float, 0>(c, x.p, cs.n);
                else ISLE_NCCL_CHECK(ncclAllReduce(x.p, x.p, cs.n, ncclFloat32, ncclSum, c.comm, c.stream));
            }
            ISLE_CUDA_CHECK(cudaEventRecord(e1, c.stream));
            ISLE_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f;
            ISLE_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            c.counters[std::string("selftest_") + (which == 0 ? "p2p_" : "nccl_") + cs.name + "_us"] = ms * 1000.0 / 20.0;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

static void allreduce(Ctx &c, void *buf, size_t n, ncclDataType_t t, ncclRedOp_t op)
{
    if (c.world <= 1 || n == 0) return;
    StatScope s(c, "allreduce");
    bool done = false;
    if (op == ncclSum && t == ncclFloat32) done = p2p_allreduce<float, 0>(c, (float *)buf, n);
    else if (op == ncclSum && t == ncclUint32) done = p2p_allreduce<uint32_t, 0>(c, (uint32_t *)buf, n);
    else if (op == ncclSum && t == ncclUint64) done = p2p_allreduce<unsigned long long, 0>(c, (unsigned long long *)buf, n);
    else if (op == ncclSum && t == ncclFloat64) done = p2p_allreduce<double, 0>(c, (double *)buf, n);
    else if (op == ncclMax && t == ncclUint32) done = p2p_allreduce<uint32_t, 1>(c, (uint32_t *)buf, n);
    if (!done) ISLE_NCCL_CHECK(ncclAllReduce(buf, buf, n, t, op, c.comm, c.stream));
}
void allreduce_sum_f32(Ctx &c, float *b, size_t n) { allreduce(c, b, n, ncclFloat32, ncclSum); }
void allreduce_sum_u32(Ctx &c, uint32_t *b, size_t n) { allreduce(c, b, n, ncclUint32, ncclSum); }
void allreduce_sum_u64(Ctx &c, unsigned long long *b, size_t n) { allreduce(c, b, n, ncclUint64, ncclSum); }
void allreduce_sum_f64(Ctx &c, double *b, size_t n) { allreduce(c, b, n, ncclFloat64, ncclSum); }
void allreduce_max_u32(Ctx &c, uint32_t *b, size_t n) { allreduce(c, b, n, ncclUint32, ncclMax); }
void allgather_u64(Ctx &c, const unsigned long long *send, unsigned long long *recv)
{
    if (c.world <= 1) {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, 8, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    if (p2p_allgather(c, send, recv, 1)) return;
    ISLE_NCCL_CHECK(ncclAllGather(send, recv, 1, ncclUint64, c.comm, c.stream));
}
void allgather_f64(Ctx &c, const double *send, double *recv)
{
    if (c.world <= 1) {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, 8, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    if (p2p_allgather(c, send, recv, 1)) return;
    ISLE_NCCL_CHECK(ncclAllGather(send, recv, 1, ncclFloat64, c.comm, c.stream));
}
void bcast_f32(Ctx &c, float *buf, size_t n, int root)
{
    if (c.world <= 1 || n == 0) return;
    ISLE_NCCL_CHECK(ncclBroadcast(buf, buf, n, ncclFloat32, root, c.comm, c.stream));
}
void allgather_f32(Ctx &c, const float *send, float *recv, size_t count)
{
    if (count == 0) return;
    if (c.world <= 1) {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, count * 4, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    StatScope s(c, "allgather");
    if (p2p_allgather(c, send, recv, count)) return;
    ISLE_NCCL_CHECK(ncclAllGather(send, recv, count, ncclFloat32, c.comm, c.stream));
}
#else
static void need_nccl(Ctx &c)
{
    if (c.world > 1) throw Error(ISLE_ERR_ARG, "library built without NCCL");
}
void allreduce_sum_f32(Ctx &c, float *, size_t) { need_nccl(c); }
void allreduce_sum_u32(Ctx &c, uint32_t *, size_t) { need_nccl(c); }
void allreduce_sum_u64(Ctx &c, unsigned long long *, size_t) { need_nccl(c); }
void allreduce_sum_f64(Ctx &c, double *, size_t) { need_nccl(c); }
void allreduce_max_u32(Ctx &c, uint32_t *, size_t) { need_nccl(c); }
void allgather_u64(Ctx &c, const unsigned long long *send, unsigned long long *recv)
{
    need_nccl(c);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, 8, cudaMemcpyDeviceToDevice, c.stream));
}
void allgather_f64(Ctx &c, const double *send, double *recv)
{
    need_nccl(c);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, 8, cudaMemcpyDeviceToDevice, c.stream));
}
void bcast_f32(Ctx &c, float *, size_t, int) { need_nccl(c); }
void p2p_destroy(Ctx &) {}
bool p2p_two_shot_begin(Ctx &, size_t, P2pArgs *, unsigned *) { return false; }
void selftest_collectives(Ctx &, unsigned long long *m, int *a) { *m = 0; *a = 0; }
void allgather_f32(Ctx &c, const float *send, float *recv, size_t count)
{
    need_nccl(c);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, count * 4, cudaMemcpyDeviceToDevice, c.stream));
}
#endif

}  // namespace isle
