// capi.cu -- extern "C" surface of libisle_cuda.so (include/isle_cuda.h).  Converts C++
// exceptions into return codes + isle_cuda_last_error(), the convention INTEGRATION.md's
// replacement translation unit turns back into std::runtime_error for ISLETrain's catch-all
// (reference drivers/ISLETrain.cpp:48-50).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>

#include "common.cuh"

using namespace isle;

struct Group;

struct isle_cuda_ctx {
    Ctx c;
    Group *g = nullptr;     // multi-GPU context (isle_cuda_create_multi): `c` is unused, the work runs on g's per-GPU contexts
};

// Multi-GPU context behind the same C ABI (isle_cuda_create(ctx**, n_gpus)): one host thread and one
// document-sharded per-GPU context per device inside the library, NCCL communicators from ncclCommInitAll.  The
// caller (ISLETrainer::train(), one host thread) sees one context: every entry point slices its host arrays by
// documents, runs on all GPUs at once and stitches the outputs back together in the reference's layout.
struct Group {
    int n = 0;
    std::vector<isle_cuda_ctx *> ranks;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::function<void(int, Ctx &)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> codes;
    std::vector<std::string> errors;
    std::string last_error;
    // document partition of A (upload) and of B (build_B)
    uint64_t V = 0, D = 0;
    std::vector<uint64_t> d0;        // [n+1] first original document of every rank
    std::vector<int64_t> e0;         // [n+1] first entry of A of every rank
    std::vector<uint64_t> DB, DBoff; // documents of B per rank, prefix
    std::vector<int64_t> nnzB, nnzBoff;
    std::vector<uint64_t> tm_entries, tm_off;   // (document, topic) sums per rank of the last construct_topic_model

    void worker(int r)
    {
        Ctx &c = ranks[r]->c;
        cudaSetDevice(c.device);
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int, Ctx &)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            int code = ISLE_OK;
            std::string msg;
            try {
                tls_stream() = c.stream;
                tls_cache() = &c.cache;
                f(r, c);
            } catch (const Error &e) {
                code = e.code; msg = e.what();
            } catch (const std::exception &e) {
                code = ISLE_ERR_CUDA; msg = e.what();
            }
            {
                std::lock_guard<std::mutex> lk(m);
                codes[r] = code;
                errors[r] = msg;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs f(rank, ctx) on every GPU's thread at once; first failure wins
    int run(std::function<void(int, Ctx &)> f)
    {
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(f);
            pending = n;
            ++gen;
        }
        cv_job.notify_all();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (int r = 0; r < n; ++r)
            if (codes[r] != ISLE_OK) {
                last_error = "[gpu " + std::to_string(r) + "] " + errors[r];
                return codes[r];
            }
        return ISLE_OK;
    }
    int rank_of_doc(uint64_t d) const { return (int)(std::upper_bound(d0.begin(), d0.end(), d) - d0.begin()) - 1; }
};

static thread_local std::string g_create_error;

// timeout record of the tensor-core head kernel (spmm_head.cu), if it left one
static std::string head_diag_text(const Ctx &c)
{
    const volatile uint32_t *d = c.head_diag_host;
    if (!d) return "";
    std::string out;
    for (int slot = 0; slot < 12; ++slot) {
        const volatile uint32_t *r = d + slot * 5;
        if (r[0] == 0) continue;
        char buf[160];
        std::snprintf(buf, sizeof(buf), " [tcgen05 kernel wait timed out (0x1xxx = panel_tc, else spmm_head): code=0x%x block=%u thread=%u a=%u b=%u]", r[0] - 1, r[1],
                      r[2], r[3], r[4]);
        out += buf;
    }
    return out;
}

template <class F>
static int guarded(isle_cuda_ctx *h, F &&f)
{
    if (!h) return ISLE_ERR_ARG;
    try {
        ISLE_CUDA_CHECK(cudaSetDevice(h->c.device));
        tls_stream() = h->c.stream;
        tls_cache() = &h->c.cache;
        f(h->c);
        return ISLE_OK;
    } catch (const Error &e) {
        h->c.last_error = e.what() + head_diag_text(h->c);
        return e.code;
    } catch (const std::exception &e) {
        h->c.last_error = e.what() + head_diag_text(h->c);
        return ISLE_ERR_CUDA;
    }
}

// the same rank-agnostic body on the context's GPU, or on every GPU of a multi-GPU context
template <class F>
static int guarded_all(isle_cuda_ctx *h, F &&f);

static int create_common(isle_cuda_ctx **out, int device, int rank, int world, const void *nccl_id, void *existing_comm = nullptr)
{
    if (!out) return ISLE_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        g_create_error = "no usable CUDA device (libisle_cuda has no CPU fallback)";
        return ISLE_ERR_NOGPU;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        g_create_error = "device is not sm_100 class (libisle_cuda is built for sm_100a only)";
        return ISLE_ERR_NOGPU;
    }
    auto *h = new isle_cuda_ctx();
    Ctx &c = h->c;
    c.device = device;
    c.rank = rank;
    c.world = world;
    c.num_sms = prop.multiProcessorCount;
    try {
        ISLE_CUDA_CHECK(cudaSetDevice(device));
        ISLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        ISLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
        ISLE_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        ISLE_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
        if (cudaHostAlloc((void **)&c.head_diag_host, 256, cudaHostAllocMapped) == cudaSuccess) {
            std::memset(c.head_diag_host, 0, 256);
            if (cudaHostGetDevicePointer((void **)&c.head_diag_dev, c.head_diag_host, 0) != cudaSuccess) c.head_diag_dev = nullptr;
        } else {
            cudaGetLastError();
            c.head_diag_host = nullptr;
        }
        ISLE_CUBLAS_CHECK(cublasCreate(&c.cublas));
        ISLE_CUBLAS_CHECK(cublasSetStream(c.cublas, c.stream));
        // fp32 FMA GEMMs only: no TF32 / reduced-precision paths in the eigensolver
        ISLE_CUBLAS_CHECK(cublasSetMathMode(
            c.cublas, (cublasMath_t)(CUBLAS_DEFAULT_MATH | CUBLAS_MATH_DISALLOW_REDUCED_PRECISION_REDUCTION)));
        ISLE_CUSOLVER_CHECK(cusolverDnCreate(&c.cusolver));
        ISLE_CUSOLVER_CHECK(cusolverDnSetStream(c.cusolver, c.stream));
        if (world > 1 && existing_comm) {
#ifdef ISLE_WITH_NCCL
            c.comm = (ncclComm_t)existing_comm;
#endif
        } else if (world > 1) {
#ifdef ISLE_WITH_NCCL
            ISLE_REQUIRE(nccl_id != nullptr, ISLE_ERR_ARG, "create_sharded: nccl_id is NULL");
            ncclUniqueId id;
            std::memcpy(&id, nccl_id, sizeof(id));
            ncclResult_t r = ncclCommInitRank(&c.comm, world, id, rank);
            if (r != ncclSuccess) throw Error(ISLE_ERR_CUDA, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
#else
            throw Error(ISLE_ERR_ARG, "library built without NCCL");
#endif
        }
    } catch (const std::exception &e) {
        g_create_error = e.what();
        delete h;
        return ISLE_ERR_CUDA;
    }
    *out = h;
    return ISLE_OK;
}


// ------------------------------------------------------------------------------------------------------------
// Multi-GPU forms of the entry points (contexts from isle_cuda_create_multi).  Documents are cut into n contiguous,
// balanced ranges; rank r works on [d0[r], d0[r+1]).  Everything the reference holds per document (B's columns,
// original_cols, assignments, (doc, topic) sums) is stitched back in document order; everything global (zetas,
// eigenpairs, centers, thresholds, the model) is identical on all ranks and taken from rank 0.
// ------------------------------------------------------------------------------------------------------------
namespace {

int m_upload_A(Group &g, uint64_t V, uint64_t D, int64_t nnz, const float *vals, const void *rows, bool rows64, const int64_t *offsets,
               float avg, uint64_t /*nz_docs: recomputed per slice*/)
{
    if (!vals || !rows || !offsets || nnz != offsets[D]) { g.last_error = "upload_A: bad arguments"; return ISLE_ERR_ARG; }
    g.V = V; g.D = D;
    g.d0.assign(g.n + 1, 0); g.e0.assign(g.n + 1, 0);
    for (int r = 0; r <= g.n; ++r) { g.d0[r] = D * (uint64_t)r / (uint64_t)g.n; g.e0[r] = offsets[g.d0[r]]; }
    return g.run([&](int r, Ctx &c) {
        const uint64_t a = g.d0[r], b = g.d0[r + 1];
        std::vector<int64_t> off(b - a + 1);
        uint64_t nz = 0;
        for (uint64_t d = a; d <= b; ++d) off[d - a] = offsets[d] - offsets[a];
        for (uint64_t d = a; d < b; ++d) nz += offsets[d + 1] > offsets[d];
        const void *rp = rows64 ? (const void *)((const uint64_t *)rows + g.e0[r]) : (const void *)((const uint32_t *)rows + g.e0[r]);
        upload_A(c, V, b - a, g.e0[r + 1] - g.e0[r], vals + g.e0[r], rp, rows64, off.data(), avg, nz);
    });
}

int m_build_B(Group &g, const uint8_t *select, int64_t *nnzB_out, uint64_t *DB_out)
{
    g.DB.assign(g.n, 0); g.nnzB.assign(g.n, 0);
    const int rc = g.run([&](int r, Ctx &c) {
        int64_t nb = 0; uint64_t db = 0;
        build_B(c, select ? select + g.d0[r] : nullptr, &nb, &db);
        g.nnzB[r] = nb; g.DB[r] = db;
    });
    g.DBoff.assign(g.n + 1, 0); g.nnzBoff.assign(g.n + 1, 0);
    for (int r = 0; r < g.n; ++r) { g.DBoff[r + 1] = g.DBoff[r] + g.DB[r]; g.nnzBoff[r + 1] = g.nnzBoff[r] + g.nnzB[r]; }
    if (nnzB_out) *nnzB_out = g.nnzBoff[g.n];
    if (DB_out) *DB_out = g.DBoff[g.n];
    return rc;
}

int m_download_B(Group &g, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig)
{
    return g.run([&](int r, Ctx &c) {
        std::vector<int64_t> off(offsets ? g.DB[r] + 1 : 0);
        std::vector<uint64_t> oc(orig ? g.DB[r] : 0);
        download_B(c, vals ? vals + g.nnzBoff[r] : nullptr, rows ? rows + g.nnzBoff[r] : nullptr, offsets ? off.data() : nullptr,
                   orig ? oc.data() : nullptr);
        if (offsets) {      // rank r's columns follow rank r-1's: shift by the entries before them; the last rank closes the array
            const uint64_t cnt = g.DB[r] + (r == g.n - 1 ? 1 : 0);
            for (uint64_t i = 0; i < cnt; ++i) offsets[g.DBoff[r] + i] = off[i] + g.nnzBoff[r];
        }
        if (orig)
            for (uint64_t i = 0; i < g.DB[r]; ++i) orig[g.DBoff[r] + i] = oc[i] + g.d0[r];
    });
}

// global outputs: every rank computes, rank 0 writes the host arrays
template <class F>
int m_rank0_out(Group &g, F f) { return g.run([&](int r, Ctx &c) { f(c, r == 0); }); }

}  // namespace

template <class F>
static int guarded_all(isle_cuda_ctx *h, F &&f)
{
    if (h && h->g) return h->g->run([&](int, Ctx &c) { f(c); });
    return guarded(h, f);
}

static int multi_unsupported(isle_cuda_ctx *h, const char *what)
{
    h->g->last_error = std::string(what) + ": not available on a multi-GPU context";
    return ISLE_ERR_ARG;
}

extern "C" {

int isle_cuda_create(isle_cuda_ctx **ctx, int device) { return create_common(ctx, device, 0, 1, nullptr); }

int isle_cuda_create_sharded(isle_cuda_ctx **ctx, int device, int rank, int world, const void *nccl_id)
{
    if (world < 1 || rank < 0 || rank >= world) return ISLE_ERR_ARG;
    return create_common(ctx, device, rank, world, nccl_id);
}

int isle_cuda_create_multi(isle_cuda_ctx **ctx, int n_gpus, const int *devices)
{
    if (!ctx || n_gpus < 1) return ISLE_ERR_ARG;
    *ctx = nullptr;
    if (n_gpus == 1) return create_common(ctx, devices ? devices[0] : 0, 0, 1, nullptr);
#ifdef ISLE_WITH_NCCL
    std::vector<int> devs(n_gpus);
    for (int r = 0; r < n_gpus; ++r) devs[r] = devices ? devices[r] : r;
    std::vector<ncclComm_t> comms(n_gpus);
    if (ncclCommInitAll(comms.data(), n_gpus, devs.data()) != ncclSuccess) {
        g_create_error = "ncclCommInitAll failed (fewer than n_gpus usable devices?)";
        return ISLE_ERR_NOGPU;
    }
    auto *h = new isle_cuda_ctx();
    auto *g = new Group();
    h->g = g;
    g->n = n_gpus;
    g->codes.assign(n_gpus, ISLE_OK);
    g->errors.assign(n_gpus, "");
    for (int r = 0; r < n_gpus; ++r) {
        isle_cuda_ctx *rc = nullptr;
        const int code = create_common(&rc, devs[r], r, n_gpus, nullptr, comms[r]);
        if (code != ISLE_OK) {
            for (auto *p : g->ranks) isle_cuda_destroy(p);
            delete g;
            delete h;
            return code;
        }
        g->ranks.push_back(rc);
    }
    for (int r = 0; r < n_gpus; ++r) g->threads.emplace_back([g, r] { g->worker(r); });
    *ctx = h;
    return ISLE_OK;
#else
    g_create_error = "library built without NCCL";
    return ISLE_ERR_ARG;
#endif
}

int isle_cuda_nccl_unique_id(void *id128)
{
#ifdef ISLE_WITH_NCCL
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (!id128 || ncclGetUniqueId(&id) != ncclSuccess) return ISLE_ERR_CUDA;
    std::memcpy(id128, &id, sizeof(id));
    return ISLE_OK;
#else
    (void)id128;
    return ISLE_ERR_ARG;
#endif
}

void isle_cuda_destroy(isle_cuda_ctx *h)
{
    if (!h) return;
    if (h->g) {
        Group *g = h->g;
        {
            std::lock_guard<std::mutex> lk(g->m);
            g->stop = true;
        }
        g->cv_job.notify_all();
        for (auto &t : g->threads) t.join();
        for (auto *p : g->ranks) isle_cuda_destroy(p);
        delete g;
        delete h;
        return;
    }
    Ctx &c = h->c;
    cudaSetDevice(c.device);
    tls_stream() = c.stream;
    tls_cache() = &c.cache;
    if (c.dl_active) { try { download_B_end(c); } catch (...) {} }
    if (c.stream) cudaStreamSynchronize(c.stream);
    if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); cudaEventDestroy(c.ev_copy); }
    for (auto &kv : c.stats)
        for (auto &ev : kv.second.pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    p2p_destroy(c);
#ifdef ISLE_WITH_NCCL
    if (c.comm) ncclCommDestroy(c.comm);
#endif
    if (c.timer0) { cudaEventDestroy(c.timer0); cudaEventDestroy(c.timer1); }
    if (c.stream2) { cudaStreamSynchronize(c.stream2); cudaStreamDestroy(c.stream2); }
    if (c.ev_fork) cudaEventDestroy(c.ev_fork);
    if (c.ev_join) cudaEventDestroy(c.ev_join);
    if (c.head_diag_host) cudaFreeHost(c.head_diag_host);
    if (c.small_host) cudaFreeHost(c.small_host);
    if (c.cusolver) cusolverDnDestroy(c.cusolver);
    if (c.cublas) cublasDestroy(c.cublas);
    // device buffers are released by the DevBuf destructors inside Ctx
    cudaStream_t s = c.stream;
    delete h;
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    tls_stream() = nullptr;
    tls_cache() = nullptr;
}

const char *isle_cuda_last_error(const isle_cuda_ctx *h)
{
    if (!h) return g_create_error.c_str();
    return h->g ? h->g->last_error.c_str() : h->c.last_error.c_str();
}

int isle_cuda_upload_A(isle_cuda_ctx *h, uint64_t V, uint64_t D, int64_t nnz, const float *vals, const uint64_t *rows,
                       const int64_t *offsets, float avg, uint64_t nz_docs)
{
    if (h && h->g) return m_upload_A(*h->g, V, D, nnz, vals, rows, true, offsets, avg, nz_docs);
    return guarded(h, [&](Ctx &c) { upload_A(c, V, D, nnz, vals, rows, true, offsets, avg, nz_docs); });
}

int isle_cuda_upload_A_u32(isle_cuda_ctx *h, uint64_t V, uint64_t D, int64_t nnz, const float *vals,
                           const uint32_t *rows, const int64_t *offsets, float avg, uint64_t nz_docs)
{
    if (h && h->g) return m_upload_A(*h->g, V, D, nnz, vals, rows, false, offsets, avg, nz_docs);
    return guarded(h, [&](Ctx &c) { upload_A(c, V, D, nnz, vals, rows, false, offsets, avg, nz_docs); });
}

int isle_cuda_ingest_text(isle_cuda_ctx *h, const char *text, uint64_t size, uint64_t V, uint64_t D, int64_t max_entries,
                          int64_t *nnz_out, float *avg_doc_sz_out, uint64_t *nz_docs_out, uint64_t *tokens_out)
{
    if (h && h->g) return multi_unsupported(h, "ingest_text (split the text by documents and use one sharded context per GPU)");
    return guarded(h, [&](Ctx &c) { ingest_text(c, text, size, V, D, max_entries, nnz_out, avg_doc_sz_out, nz_docs_out, tokens_out); });
}

int isle_cuda_upload_counts(isle_cuda_ctx *h, uint64_t V, uint64_t D, int64_t nnz, const uint32_t *counts, const uint32_t *rows,
                            const int64_t *offsets, float *avg_doc_sz_out, uint64_t *nz_docs_out)
{
    if (h && h->g) return multi_unsupported(h, "upload_counts");
    return guarded(h, [&](Ctx &c) { upload_counts(c, V, D, nnz, counts, rows, offsets, avg_doc_sz_out, nz_docs_out); });
}

int isle_cuda_download_A(isle_cuda_ctx *h, float *normalized_vals, uint64_t *rows, int64_t *offsets)
{
    if (h && h->g) return multi_unsupported(h, "download_A");
    return guarded(h, [&](Ctx &c) { download_A(c, normalized_vals, rows, offsets); });
}

int isle_cuda_thresholds(isle_cuda_ctx *h, uint64_t k, float *zetas_out, int64_t *new_nnz_out)
{
    if (h && h->g) return m_rank0_out(*h->g, [&](Ctx &c, bool out) { compute_thresholds(c, k, out ? zetas_out : nullptr, out ? new_nnz_out : nullptr); });
    return guarded(h, [&](Ctx &c) { compute_thresholds(c, k, zetas_out, new_nnz_out); });
}

int isle_cuda_build_B(isle_cuda_ctx *h, const uint8_t *select, int64_t *nnzB, uint64_t *DB)
{
    if (h && h->g) return m_build_B(*h->g, select, nnzB, DB);
    return guarded(h, [&](Ctx &c) { build_B(c, select, nnzB, DB); });
}

int isle_cuda_sampling_weights(isle_cuda_ctx *h, float *out)
{
    if (h && h->g) { Group &g = *h->g; return g.run([&](int r, Ctx &c) { sampling_weights(c, out + g.d0[r]); }); }
    return guarded(h, [&](Ctx &c) { sampling_weights(c, out); });
}

int isle_cuda_download_B(isle_cuda_ctx *h, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig)
{
    if (h && h->g) return m_download_B(*h->g, vals, rows, offsets, orig);
    return guarded(h, [&](Ctx &c) { download_B(c, vals, rows, offsets, orig); });
}

int isle_cuda_download_B_begin(isle_cuda_ctx *h, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig)
{
    if (h && h->g) return m_download_B(*h->g, vals, rows, offsets, orig);      // multi-GPU: stitched synchronously
    return guarded(h, [&](Ctx &c) { download_B_begin(c, vals, rows, offsets, orig); });
}

int isle_cuda_download_B_end(isle_cuda_ctx *h)
{
    if (h && h->g) return ISLE_OK;
    return guarded(h, [&](Ctx &c) { download_B_end(c); });
}

int isle_cuda_selftest_collectives(isle_cuda_ctx *h, uint64_t *mismatches_out, int *p2p_active_out)
{
    if (!mismatches_out || !p2p_active_out) return ISLE_ERR_ARG;
    *mismatches_out = 0;
    *p2p_active_out = 0;
    std::mutex mu;
    return guarded_all(h, [&](Ctx &c) {
        unsigned long long m = 0;
        int a = 0;
        selftest_collectives(c, &m, &a);
        std::lock_guard<std::mutex> lk(mu);
        *mismatches_out += m;
        *p2p_active_out |= a;
    });
}

int isle_cuda_frobenius(isle_cuda_ctx *h, float *out)
{
    if (h && h->g)      // a sharded context's frobenius() is already the all-reduced, corpus-wide sum
        return m_rank0_out(*h->g, [&](Ctx &c, bool o) { const float f = frobenius(c); if (o) *out = f; });
    return guarded(h, [&](Ctx &c) { *out = frobenius(c); });
}

int isle_cuda_spsptr_multiply(isle_cuda_ctx *h, int b, const float *X, float *Z)
{
    auto body = [&](Ctx &c, bool out) {
        ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "spsptr_multiply: build_B first");
        ISLE_REQUIRE(b >= 1 && b <= 16, ISLE_ERR_ARG, "spsptr_multiply: block size must be in [1,16]");
        DevBuf<float> dX((size_t)c.V * b), dZ((size_t)c.V * b);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(dX.p, X, dX.bytes(), cudaMemcpyHostToDevice, c.stream));
        spsptr_multiply_dev(c, b, dX.p, dZ.p);
        if (out) ISLE_CUDA_CHECK(cudaMemcpyAsync(Z, dZ.p, dZ.bytes(), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    };
    if (h && h->g) return m_rank0_out(*h->g, body);
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "spsptr_multiply: build_B first");
        ISLE_REQUIRE(b >= 1 && b <= 16, ISLE_ERR_ARG, "spsptr_multiply: block size must be in [1,16]");
        DevBuf<float> dX((size_t)c.V * b), dZ((size_t)c.V * b);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(dX.p, X, dX.bytes(), cudaMemcpyHostToDevice, c.stream));
        spsptr_multiply_dev(c, b, dX.p, dZ.p);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(Z, dZ.p, dZ.bytes(), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int isle_cuda_block_ks(isle_cuda_ctx *h, uint64_t k, int b, int max_restarts, float tol, uint64_t seed,
                       float *evalues_out, float *U_out, int *nconv_out)
{
    if (h && h->g)
        return m_rank0_out(*h->g, [&](Ctx &c, bool out) {
            int nconv = 0;
            block_ks(c, k, b, max_restarts, tol, seed, out ? evalues_out : nullptr, out ? U_out : nullptr, &nconv);
            if (out && nconv_out) *nconv_out = nconv;
        });
    return guarded(h, [&](Ctx &c) { block_ks(c, k, b, max_restarts, tol, seed, evalues_out, U_out, nconv_out); });
}

int isle_cuda_set_U(isle_cuda_ctx *h, uint64_t k, const float *U)
{
    auto body = [&](Ctx &c) {
        ISLE_REQUIRE(c.have_B && k >= 1 && U, ISLE_ERR_ARG, "set_U: build_B first");
        c.k = k;
        c.U.alloc((size_t)c.V * k);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(c.U.p, U, c.U.bytes(), cudaMemcpyHostToDevice, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        c.have_U = true;
        c.have_P = false;
    };
    if (h && h->g) return h->g->run([&](int, Ctx &c) { body(c); });
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(c.have_B && k >= 1 && U, ISLE_ERR_ARG, "set_U: build_B first");
        c.k = k;
        c.U.alloc((size_t)c.V * k);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(c.U.p, U, c.U.bytes(), cudaMemcpyHostToDevice, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        c.have_U = true;
        c.have_P = false;
    });
}

int isle_cuda_project(isle_cuda_ctx *h, float *P_out, float *l2_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        return g.run([&](int r, Ctx &c) {
            project(c);
            if (P_out && c.DB)
                ISLE_CUDA_CHECK(cudaMemcpy2DAsync(P_out + (size_t)g.DBoff[r] * c.k, (size_t)c.k * 4, c.P.p, (size_t)c.kp * 4, (size_t)c.k * 4,
                                                  (size_t)c.DB, cudaMemcpyDeviceToHost, c.stream));
            if (l2_out && c.DB)
                ISLE_CUDA_CHECK(cudaMemcpyAsync(l2_out + g.DBoff[r], c.p_l2.p, (size_t)c.DB * 4, cudaMemcpyDeviceToHost, c.stream));
            ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        });
    }
    return guarded(h, [&](Ctx &c) {
        project(c);
        if (P_out && c.DB)
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(P_out, (size_t)c.k * 4, c.P.p, (size_t)c.kp * 4, (size_t)c.k * 4,
                                              (size_t)c.DB, cudaMemcpyDeviceToHost, c.stream));
        if (l2_out && c.DB)
            ISLE_CUDA_CHECK(cudaMemcpyAsync(l2_out, c.p_l2.p, (size_t)c.DB * 4, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int isle_cuda_kmeanspp(isle_cuda_ctx *h, uint64_t k, uint64_t seed, uint64_t *seeds_out, float *centers_out,
                       float *residual_out)
{
    if (h && h->g)
        return m_rank0_out(*h->g, [&](Ctx &c, bool out) {
            kmeanspp(c, k, seed, out ? seeds_out : nullptr, out ? centers_out : nullptr, out ? residual_out : nullptr);
        });
    return guarded(h, [&](Ctx &c) { kmeanspp(c, k, seed, seeds_out, centers_out, residual_out); });
}

int isle_cuda_lloyd_projected(isle_cuda_ctx *h, uint64_t k, float *centers_inout, int max_reps, uint32_t *assign_out,
                              double *objective_out, int *iters_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        if (!centers_inout || max_reps < 1) { g.last_error = "lloyd_projected: bad arguments"; return ISLE_ERR_ARG; }
        return g.run([&](int r, Ctx &c) {
            // every rank iterates on identical (all-reduced) centers; ranks > 0 work on a private copy of the host array
            std::vector<float> mine;
            float *cen = centers_inout;
            if (r != 0) { mine.assign(centers_inout, centers_inout + (size_t)k * k); cen = mine.data(); }
            double obj = 0.0; int iters = 0;
            lloyd_projected(c, k, cen, max_reps, assign_out ? assign_out + g.DBoff[r] : nullptr, &obj, &iters);
            if (r == 0) { if (objective_out) *objective_out = obj; if (iters_out) *iters_out = iters; }
        });
    }
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(centers_inout != nullptr && max_reps >= 1, ISLE_ERR_ARG, "lloyd_projected: bad arguments");
        lloyd_projected(c, k, centers_inout, max_reps, assign_out, objective_out, iters_out);
    });
}

int isle_cuda_lloyd_full(isle_cuda_ctx *h, uint64_t k, float *centers_inout, int max_reps, uint32_t *assign_out,
                         double *objective_out, int *iters_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        if (max_reps < 1) { g.last_error = "lloyd_full: bad arguments"; return ISLE_ERR_ARG; }
        return g.run([&](int r, Ctx &c) {
            std::vector<float> mine;
            float *cen = centers_inout;
            if (r != 0 && centers_inout) { mine.assign(centers_inout, centers_inout + (size_t)k * c.V); cen = mine.data(); }
            double obj = 0.0; int iters = 0;
            lloyd_full(c, k, cen, max_reps, assign_out ? assign_out + g.DBoff[r] : nullptr, &obj, &iters);
            if (r == 0) { if (objective_out) *objective_out = obj; if (iters_out) *iters_out = iters; }
        });
    }
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(max_reps >= 1, ISLE_ERR_ARG, "lloyd_full: bad arguments");
        lloyd_full(c, k, centers_inout, max_reps, assign_out, objective_out, iters_out);
    });
}

int isle_cuda_sample_docs(isle_cuda_ctx *h, float sample_rate, uint64_t seed, uint8_t *select_out, uint64_t *n_selected_out)
{
    if (h && h->g) {      // every GPU selects among its own documents against the corpus-wide pivot; the masks are stitched in document order
        Group &g = *h->g;
        if (!select_out) { g.last_error = "sample_docs: bad arguments"; return ISLE_ERR_ARG; }
        std::mutex mu;
        uint64_t total = 0;
        const int rc = g.run([&](int r, Ctx &c) {
            uint64_t n = 0;
            sample_docs(c, sample_rate, seed, select_out + g.d0[r], &n);
            std::lock_guard<std::mutex> lk(mu);
            total += n;
        });
        if (n_selected_out) *n_selected_out = total;
        return rc;
    }
    return guarded(h, [&](Ctx &c) { sample_docs(c, sample_rate, seed, select_out, n_selected_out); });
}

int isle_cuda_catchword_thresholds(isle_cuda_ctx *h, uint64_t k, uint64_t r, const uint32_t *cluster_of_doc, float *thresholds_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        return g.run([&](int rk, Ctx &c) { catchword_thresholds(c, k, r, cluster_of_doc + g.d0[rk], rk == 0 ? thresholds_out : nullptr); });
    }
    return guarded(h, [&](Ctx &c) { catchword_thresholds(c, k, r, cluster_of_doc, thresholds_out); });
}

int isle_cuda_rth_highest_element(isle_cuda_ctx *h, uint64_t r, const uint64_t *docs, uint64_t ndocs, float *thresholds_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        return g.run([&](int rk, Ctx &c) {     // the cluster's documents that live on this rank, as local ids
            std::vector<uint64_t> mine;
            for (uint64_t i = 0; i < ndocs; ++i)
                if (docs[i] >= g.d0[rk] && docs[i] < g.d0[rk + 1]) mine.push_back(docs[i] - g.d0[rk]);
            std::vector<float> scratch(rk == 0 ? 0 : c.V);
            rth_highest_element(c, r, mine.data(), mine.size(), rk == 0 ? thresholds_out : scratch.data());
        });
    }
    return guarded(h, [&](Ctx &c) { rth_highest_element(c, r, docs, ndocs, thresholds_out); });
}

int isle_cuda_find_catchwords(isle_cuda_ctx *h, uint64_t k, const float *thresholds, double rho, int32_t *topic_of_word_out)
{
    if (h && h->g)
        return h->g->run([&](int rk, Ctx &c) {
            std::vector<int32_t> scratch(rk == 0 ? 0 : c.V);
            find_catchwords(c, k, thresholds, rho, rk == 0 ? topic_of_word_out : scratch.data());
        });
    return guarded(h, [&](Ctx &c) { find_catchwords(c, k, thresholds, rho, topic_of_word_out); });
}

int isle_cuda_construct_topic_model(isle_cuda_ctx *h, uint64_t k, const int32_t *topic_of_word, const uint32_t *cluster_of_doc,
                                    uint64_t rank_threshold, float *model_out, uint64_t *num_entries_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        g.tm_entries.assign(g.n, 0);
        const int rc = g.run([&](int rk, Ctx &c) {
            uint64_t n = 0;
            construct_topic_model(c, k, topic_of_word, cluster_of_doc + g.d0[rk], rank_threshold, rk == 0 ? model_out : nullptr, &n);
            g.tm_entries[rk] = n;
        });
        g.tm_off.assign(g.n + 1, 0);
        for (int rk = 0; rk < g.n; ++rk) g.tm_off[rk + 1] = g.tm_off[rk] + g.tm_entries[rk];
        if (num_entries_out) *num_entries_out = g.tm_off[g.n];
        return rc;
    }
    return guarded(h, [&](Ctx &c) { construct_topic_model(c, k, topic_of_word, cluster_of_doc, rank_threshold, model_out, num_entries_out); });
}

int isle_cuda_doc_topic_sums(isle_cuda_ctx *h, uint32_t *docs, uint32_t *topics, float *sums)
{
    if (h && h->g) {
        Group &g = *h->g;
        return g.run([&](int rk, Ctx &c) {
            uint32_t *d = docs ? docs + g.tm_off[rk] : nullptr;
            download_doc_topic_sums(c, d, topics ? topics + g.tm_off[rk] : nullptr, sums ? sums + g.tm_off[rk] : nullptr);
            if (d) for (uint64_t i = 0; i < g.tm_entries[rk]; ++i) d[i] += (uint32_t)g.d0[rk];      // local -> original document ids
        });
    }
    return guarded(h, [&](Ctx &c) { download_doc_topic_sums(c, docs, topics, sums); });
}

int isle_cuda_panel_products(isle_cuda_ctx *h, int64_t n, int rows, int b, const float *W, float *F_inout, float *C_out, int engine)
{
    if (h && h->g) h = h->g->ranks[0];      // harness call on caller data: one GPU
    return guarded(h, [&](Ctx &c) { panel_products(c, n, rows, b, W, F_inout, C_out, engine); });
}

int isle_cuda_assign_projected(isle_cuda_ctx *h, uint64_t k, const float *centers, uint32_t *assign_out)
{
    if (h && h->g) {
        Group &g = *h->g;
        if (!centers || !assign_out) { g.last_error = "assign_projected: bad arguments"; return ISLE_ERR_ARG; }
        return g.run([&](int r, Ctx &c) { assign_projected(c, k, centers, assign_out + g.DBoff[r]); });
    }
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(centers && assign_out, ISLE_ERR_ARG, "assign_projected: bad arguments");
        assign_projected(c, k, centers, assign_out);
    });
}

int isle_cuda_update_min_dist(isle_cuda_ctx *h, uint64_t num_centers, const float *projected_centers, float *min_dist_inout)
{
    if (h && h->g) {
        Group &g = *h->g;
        return g.run([&](int r, Ctx &c) { update_min_dist(c, num_centers, projected_centers, min_dist_inout + g.DBoff[r]); });
    }
    return guarded(h, [&](Ctx &c) { update_min_dist(c, num_centers, projected_centers, min_dist_inout); });
}

int isle_cuda_lift_centers(isle_cuda_ctx *h, uint64_t ncols, const float *in, uint64_t ld_in, float *out)
{
    if (h && h->g) return m_rank0_out(*h->g, [&](Ctx &c, bool o) { lift_centers(c, ncols, in, ld_in, o ? out : nullptr); });
    return guarded(h, [&](Ctx &c) { lift_centers(c, ncols, in, ld_in, out); });
}

int isle_cuda_cleanup_eigensolver(isle_cuda_ctx *h)
{
    return guarded_all(h, [&](Ctx &c) {
        c.U.release();
        c.P.release();
        c.P_lo.release();
        c.p_l2.release();
        c.have_U = c.have_P = false;
    });
}

int isle_cuda_set_profiling(isle_cuda_ctx *h, int enabled)
{
    return guarded_all(h, [&](Ctx &c) { c.profiling = enabled != 0; });
}

static void drain(Ctx &c)
{
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    for (auto &kv : c.stats) {
        for (auto &ev : kv.second.pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) kv.second.ms += ms;
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        kv.second.pending.clear();
    }
}

int isle_cuda_get_stat(isle_cuda_ctx *h, const char *name, double *out)
{
    if (h && h->g) {      // totals for the per-document quantities, rank 0's numbers for everything else
        Group &g = *h->g;
        const std::string n(name ? name : "");
        if (n == "nnz_B") { *out = g.nnzBoff.empty() ? 0.0 : (double)g.nnzBoff[g.n]; return ISLE_OK; }
        if (n == "D_B") { *out = g.DBoff.empty() ? 0.0 : (double)g.DBoff[g.n]; return ISLE_OK; }
        if (n == "n_gpus") { *out = g.n; return ISLE_OK; }
        h = g.ranks[0];
    }
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(name && out, ISLE_ERR_ARG, "get_stat: bad arguments");
        drain(c);
        const std::string n(name);
        if (n == "launches") { *out = c.launches; return; }
        if (n == "nnz_B") { *out = (double)c.nnzB; return; }
        if (n == "D_B") { *out = (double)c.DB; return; }
        if (n == "alloc_misses") { *out = c.cache.misses; return; }
        if (n == "alloc_hits") { *out = c.cache.hits; return; }
        if (n == "alloc_cached_bytes") { *out = (double)c.cache.cached_bytes; return; }
        if (n == "spmm_head_words") { *out = (double)c.H; return; }
        if (n == "spmm_tail_nnz") { *out = (double)c.nnz_tail; return; }
        auto ci = c.counters.find(n);
        if (ci != c.counters.end()) { *out = ci->second; return; }
        const size_t us = n.rfind('_');
        if (us != std::string::npos) {
            const std::string base = n.substr(0, us), field = n.substr(us + 1);
            auto it = c.stats.find(base);
            if (field == "ms" || field == "calls" || field == "bytes" || field == "flops") {
                if (it == c.stats.end()) { *out = 0.0; return; }
                *out = field == "ms" ? it->second.ms : field == "calls" ? it->second.calls
                     : field == "bytes" ? it->second.bytes : it->second.flops;
                return;
            }
        }
        throw Error(ISLE_ERR_ARG, "get_stat: unknown counter " + n);
    });
}

int isle_cuda_reset_stats(isle_cuda_ctx *h)
{
    return guarded_all(h, [&](Ctx &c) {
        drain(c);
        c.stats.clear();
        c.counters.clear();
        c.cache.misses = c.cache.hits = 0.0;
        c.launches = 0.0;
    });
}

int isle_cuda_timer_start(isle_cuda_ctx *h)
{
    return guarded_all(h, [&](Ctx &c) {
        if (!c.timer0) {
            ISLE_CUDA_CHECK(cudaEventCreate(&c.timer0));
            ISLE_CUDA_CHECK(cudaEventCreate(&c.timer1));
        }
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        ISLE_CUDA_CHECK(cudaEventRecord(c.timer0, c.stream));
    });
}

int isle_cuda_timer_stop(isle_cuda_ctx *h, double *ms_out)
{
    if (h && h->g) {      // the slowest GPU's time
        Group &g = *h->g;
        std::vector<double> ms(g.n, 0.0);
        const int rc = g.run([&](int r, Ctx &c) {
            ISLE_REQUIRE(c.timer0, ISLE_ERR_ARG, "timer_stop: timer_start first");
            ISLE_CUDA_CHECK(cudaEventRecord(c.timer1, c.stream));
            ISLE_CUDA_CHECK(cudaEventSynchronize(c.timer1));
            float t = 0.f;
            ISLE_CUDA_CHECK(cudaEventElapsedTime(&t, c.timer0, c.timer1));
            ms[r] = t;
        });
        if (ms_out) *ms_out = *std::max_element(ms.begin(), ms.end());
        return rc;
    }
    return guarded(h, [&](Ctx &c) {
        ISLE_REQUIRE(c.timer0 && ms_out, ISLE_ERR_ARG, "timer_stop: timer_start first");
        ISLE_CUDA_CHECK(cudaEventRecord(c.timer1, c.stream));
        ISLE_CUDA_CHECK(cudaEventSynchronize(c.timer1));
        float ms = 0.f;
        ISLE_CUDA_CHECK(cudaEventElapsedTime(&ms, c.timer0, c.timer1));
        *ms_out = ms;
    });
}

int isle_cuda_set_option(isle_cuda_ctx *h, const char *name, int value)
{
    return guarded_all(h, [&](Ctx &c) {
        ISLE_REQUIRE(name, ISLE_ERR_ARG, "set_option: bad arguments");
        c.options[name] = value;
        // operator layout options take effect at the next build of the head/tail split
        if (std::strncmp(name, "spmm_head", 9) == 0) c.have_csr = false;
    });
}

}  // extern "C"
