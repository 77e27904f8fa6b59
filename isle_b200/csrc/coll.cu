// coll.cu -- the few collectives of the document-sharded path, NCCL over
// NVLink.  All are small, latency-bound messages issued on the context stream directly
// behind the kernel that produced the buffer; with world == 1 they are no-ops.
#include "common.cuh"

namespace isle {

#ifdef ISLE_WITH_NCCL
#define ISLE_NCCL_CHECK(expr)                                                                       \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess)                                                                      \
            throw ::isle::Error(ISLE_ERR_CUDA, std::string(#expr) + ": " + ncclGetErrorString(_r)); \
    } while (0)

static void allreduce(Ctx &c, void *buf, size_t n, ncclDataType_t t, ncclRedOp_t op)
{
    if (c.world <= 1 || n == 0) return;
    StatScope s(c, "allreduce");
    ISLE_NCCL_CHECK(ncclAllReduce(buf, buf, n, t, op, c.comm, c.stream));
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
    ISLE_NCCL_CHECK(ncclAllGather(send, recv, 1, ncclUint64, c.comm, c.stream));
}
void allgather_f64(Ctx &c, const double *send, double *recv)
{
    if (c.world <= 1) {
        ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, 8, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
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
void allgather_f32(Ctx &c, const float *send, float *recv, size_t count)
{
    need_nccl(c);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(recv, send, count * 4, cudaMemcpyDeviceToDevice, c.stream));
}
#endif

}  // namespace isle
