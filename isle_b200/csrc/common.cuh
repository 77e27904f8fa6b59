// common.cuh -- internal declarations shared by the kernel families of libisle_cuda.
#pragma once

#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/isle_cuda.h"

#ifdef ISLE_WITH_NCCL
#include <nccl.h>
#endif

namespace isle {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define ISLE_CUDA_CHECK(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::isle::Error(ISLE_ERR_CUDA, std::string(#expr) + ": " +                     \
                                                   cudaGetErrorString(_e) + " @" + __FILE__ +  \
                                                   ":" + std::to_string(__LINE__));            \
    } while (0)

#define ISLE_CUBLAS_CHECK(expr)                                                                \
    do {                                                                                       \
        cublasStatus_t _s = (expr);                                                            \
        if (_s != CUBLAS_STATUS_SUCCESS)                                                       \
            throw ::isle::Error(ISLE_ERR_CUDA, std::string(#expr) + ": cublas status " +       \
                                                   std::to_string((int)_s) + " @" + __FILE__ + \
                                                   ":" + std::to_string(__LINE__));            \
    } while (0)

#define ISLE_CUSOLVER_CHECK(expr)                                                              \
    do {                                                                                       \
        cusolverStatus_t _s = (expr);                                                          \
        if (_s != CUSOLVER_STATUS_SUCCESS)                                                     \
            throw ::isle::Error(ISLE_ERR_CUDA, std::string(#expr) + ": cusolver status " +     \
                                                   std::to_string((int)_s) + " @" + __FILE__ + \
                                                   ":" + std::to_string(__LINE__));            \
    } while (0)

#define ISLE_REQUIRE(cond, code, msg)                                  \
    do {                                                               \
        if (!(cond)) throw ::isle::Error((code), std::string(msg));    \
    } while (0)

// The stream the calling context works on: set by every C-ABI entry point (one caller thread per
// context).
inline cudaStream_t &tls_stream()
{
    static thread_local cudaStream_t s = nullptr;
    return s;
}

// Per-context caching device allocator.  A step of the spectral core makes ~140 scratch allocations;
// through cudaMallocAsync they cost 4-30 ms of host time per step on the bench box (measured: the
// driver pool maps and unmaps physical memory behind them), which is as much as the whole operator.
// Freed buffers are therefore kept in a size-ordered free list and handed out again (best fit within
// 1/8 slack), so after the first step no allocation reaches the driver.  Reuse straight after a free
// is safe because every kernel of a context is ordered on its one main stream (the second stream only
// touches buffers that live as long as the operator layout).
struct DevCache {
    std::multimap<size_t, void *> free_;
    size_t cached_bytes = 0;
    double misses = 0.0, hits = 0.0;
    static constexpr size_t kMaxCached = (size_t)48 << 30;

    void *get(size_t bytes, size_t &cap)
    {
        auto it = free_.lower_bound(bytes);
        if (it != free_.end() && it->first <= bytes + bytes / 8 + 4096) {
            void *p = it->second;
            cap = it->first;
            cached_bytes -= cap;
            free_.erase(it);
            hits += 1.0;
            return p;
        }
        misses += 1.0;
        void *p = nullptr;
        cap = (bytes + 511) & ~(size_t)511;
        cudaError_t e = cudaMalloc(&p, cap);
        if (e != cudaSuccess) {       // out of memory: drop the cache and retry once
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, cap);
        }
        if (e != cudaSuccess)
            throw Error(ISLE_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(cap) + "): " + cudaGetErrorString(e));
        return p;
    }
    void put(void *p, size_t cap)
    {
        if (cached_bytes + cap > kMaxCached) trim();
        free_.emplace(cap, p);
        cached_bytes += cap;
    }
    void trim()
    {
        if (free_.empty()) return;
        cudaDeviceSynchronize();
        for (auto &kv : free_) cudaFree(kv.second);
        free_.clear();
        cached_bytes = 0;
    }
    ~DevCache() { trim(); }
};

inline DevCache *&tls_cache()
{
    static thread_local DevCache *c = nullptr;
    return c;
}

// Owning device buffer served by the calling context's cache; movable, not copyable.
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    size_t cap = 0;             // bytes actually owned
    DevCache *owner = nullptr;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap), owner(o.owner) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; owner = o.owner; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_) {
        release();
        n = n_;
        owner = tls_cache();
        if (n) {
            if (!owner) throw Error(ISLE_ERR_ARG, "device allocation outside a context call");
            p = static_cast<T *>(owner->get(n * sizeof(T), cap));
        }
    }
    void release() {
        if (p && owner) owner->put(p, cap);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// CUDA-event timing of one kernel family on the context stream.
struct KernelStat {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double ms = 0.0;
    double calls = 0.0;
    double bytes = 0.0;   // algorithmic bytes, summed over calls
    double flops = 0.0;
};

constexpr int kHeadTile = 128;   // rows per tensor-core tile of the head engine (UMMA M)
constexpr int kHeadChunk = 128;  // k per pipeline stage of the head engine

struct WorkItem {          // one 4-lane group's share of a sparse pass
    uint32_t out_row;      // row of the dense output this item accumulates into
    uint32_t len;          // number of nonzeros
    int64_t begin;         // first nonzero in the index stream
};

struct Ctx {
    DevCache cache;            // declared first: destroyed after every DevBuf member below
    int device = 0;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    cublasHandle_t cublas = nullptr;
    cusolverDnHandle_t cusolver = nullptr;
    int num_sms = 148;
    std::string last_error;
#ifdef ISLE_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif

    // ---- A: normalised doc-major CSC (local slice when sharded)
    uint64_t V = 0, D = 0, nz_docs = 0;   // nz_docs: GLOBAL count of non-empty docs
    int64_t nnzA = 0;
    float avg_doc_sz = 0.f;
    DevBuf<float> a_val;
    DevBuf<uint32_t> a_row;
    DevBuf<int64_t> a_off;

    // ---- thresholds
    bool have_zeta = false;
    DevBuf<float> zeta, sqrt_zeta;
    int64_t new_nnz = 0;                   // kept entries over the WHOLE corpus (all ranks: the histogram is all-reduced
                                           // before the rank rule runs), what the reference's compute_thresholds returns

    // ---- B: pattern-only doc-major CSC + word-major CSR copy (values are sqrt_zeta[row])
    bool have_B = false;
    uint64_t DB = 0;
    int64_t nnzB = 0;
    std::vector<uint64_t> db_all;          // D_B of every rank (doc-sharded numbering of B's columns)
    uint64_t db_offset = 0, db_total = 0;  // first global column id of this rank; sum over ranks
    DevBuf<uint32_t> b_row;                // [nnzB] word of each nonzero, doc-major
    DevBuf<int64_t> b_off;                 // [DB+1]
    DevBuf<uint32_t> b_orig;               // [DB] original (local) doc id
    bool have_csr = false;
    DevBuf<uint32_t> csr_col;              // [nnzB] doc of each nonzero, word-major
    DevBuf<int64_t> csr_off;               // [V+1]
    DevBuf<WorkItem> items_bt, items_b;    // work lists for Y=B^T X and Z=B Y (tail part)
    size_t n_items_bt = 0, n_items_b = 0;
    DevBuf<float> xs, ybuf, zbuf;          // padded row-major operands of the SpMM pair (xs, zbuf in rank space)
    // operator layout (spmm.cu): words relabelled by decreasing row length ("rank space"); the H
    // most frequent words form the dense head (bitmaps + tcgen05, spmm_head.cu), the rest the tail
    uint32_t H = 0;                        // head words, multiple of 128 (0: no head engine)
    uint32_t DBpad = 0;                    // DB rounded up to 128
    int64_t nnz_tail = 0;
    DevBuf<uint32_t> rank_of, word_of_rank;   // [V]
    DevBuf<uint32_t> t1_idx;               // doc-major tail: rank of each nonzero
    DevBuf<int64_t> t1_off;                // [DB+1]
    DevBuf<uint32_t> t2_idx;               // rank-major tail: doc of each nonzero
    DevBuf<int64_t> t2_off;                // [V+1] (head ranks are empty rows)
    DevBuf<uint32_t> colmax;               // per-column max |X| (bit patterns) for the exact power-of-two equilibration
    DevBuf<uint4> xbfp, ybfp;              // one-sector (32-byte) copies of the operand rows, two 16-byte units per row
    DevBuf<uint4> bits1, bits2;            // head bitmaps: (doc tile x rank chunk) and (rank tile x doc chunk)
    DevBuf<__nv_bfloat16> xsplit, ysplit;  // 3-piece bf16 splits of the head rows of Xs / of Y, K-major [N][Kpad]
    bool head_i8 = false;                  // layout built for the int8 head engine (spmm_head_i8.cu): 256-k super-chunks, plane-major bits
    DevBuf<int8_t> xdig, ydig;             // three s8 digits of the quantised head rows of Xs / of Y, K-major [32][Kpad]
    DevBuf<uint32_t> ycolmax;              // per-column max |Y| (bit patterns): quantisation scale of pass 2
    uint32_t *head_diag_host = nullptr, *head_diag_dev = nullptr;   // host-mapped timeout record of the head kernel
    cudaStream_t stream2 = nullptr;        // head pass 2 runs beside the tail pass 2
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // ---- eigen / projection state
    uint64_t k = 0;
    bool have_U = false, have_P = false;
    DevBuf<float> U;                       // V x k column-major
    uint64_t kp = 0;                       // padded projection width (multiple of 32)
    DevBuf<float> P;                       // DB x kp row-major, zero padded
    DevBuf<float> P_lo;                    // tf32(P - trunc_tf32(P)): the low part of the split-TF32 pair whose high part is P itself as the tensor core reads it (dist_tc.cu)
    DevBuf<float> p_l2;                    // DB
    DevBuf<float> catch_thr;               // k x V catchword thresholds of the last catchword_thresholds call (catchwords.cu)
    uint64_t catch_k = 0;
    DevBuf<uint32_t> tm_doc, tm_topic;     // (document, topic, sum) list of the last construct_topic_model (topic_model.cu)
    DevBuf<float> tm_val, tm_model;        // ... its sums; the V x k model (column-major)
    int64_t tm_entries = 0;
    DevBuf<float> lifted;                  // last lift_centers result (V x lifted_cols, column-major): input of lloyd_full
    uint64_t lifted_cols = 0;

    // ---- background download of B (download_B_begin / _end): a copy stream and a host thread move B to the caller's arrays
    // while the main stream goes on with the eigensolver
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
    std::thread dl_thread;
    bool dl_active = false;
    std::string dl_error;
    DevBuf<float> dl_vals;
    DevBuf<unsigned long long> dl_rows, dl_orig;

    // ---- peer-to-peer collectives over NVLink / NVSwitch (coll.cu): one workspace per rank that every other rank maps
    // into its own address space (cudaIpc between processes, peer access inside one process)
    int p2p_state = 0;                     // 0 not tried yet, 1 ready, -1 unavailable (NCCL carries every collective)
    char *p2p_ws[16] = {};                 // [world] the workspace of every rank as THIS device addresses it
    bool p2p_ipc[16] = {};                 // mapped with cudaIpcOpenMemHandle (closed on destroy)
    unsigned long long p2p_epoch = 0;      // one per collective; identical on every rank (same call sequence)
    unsigned p2p_cnt_a = 0, p2p_cnt_b = 0; // cumulative CTA-arrival targets of the two grid-level counters
    uint32_t *p2p_diag_host = nullptr, *p2p_diag_dev = nullptr;   // host-mapped record of a timed-out wait

    unsigned char *small_host = nullptr, *small_dev = nullptr;   // host-mapped landing zone of read_small()
    bool small_failed = false;

    // ---- stats / options
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    bool profiling = false;
    double launches = 0.0;
    std::map<std::string, KernelStat> stats;
    std::map<std::string, double> counters;
    std::map<std::string, int> options;

    int opt(const char *name, int dflt) const {
        auto it = options.find(name);
        return it == options.end() ? dflt : it->second;
    }
};

// RAII: brackets a kernel-family launch with events when profiling is on.
struct StatScope {
    Ctx &c;
    KernelStat *st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    StatScope(Ctx &ctx, const char *name, double bytes = 0.0, double flops = 0.0) : c(ctx) {
        st = &c.stats[name];
        st->calls += 1.0;
        st->bytes += bytes;
        st->flops += flops;
        if (c.profiling) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, c.stream);
        }
    }
    ~StatScope() {
        if (e0) {
            cudaEventRecord(e1, c.stream);
            st->pending.emplace_back(e0, e1);
        }
    }
};

inline void count_launch(Ctx &c, int n = 1) { c.launches += n; }

// Quantisation of a dense operand column for the int8 head engine (spmm_head_i8.cu): q = rn(x * 2^(29 - e)) with
// 2^e <= column max < 2^(e+1), so |q| < 2^30 fits four balanced base-256 digits; the value is q * 2^(e - 29).
// `colmax_bits` = bit pattern of the column's max |x| (NULL / 0 / non-finite: e = 0).  Returns the biased exponent.
__device__ __forceinline__ uint32_t quant_exp(const uint32_t *colmax_bits, int c)
{
    if (!colmax_bits) return 127u;
    const uint32_t E = (colmax_bits[c] >> 23) & 0xFFu;
    if (E == 0u || E >= 254u) return 127u;
    return E < 30u ? 30u : (E > 224u ? 224u : E);
}
constexpr int kHead8Rows = 40;      // rows of the digit matrix of the int8 head engine (10 columns x 4 digits)
// row of digit g (0..3) of column c (0..9) in the [40][K] digit matrix: columns 0-3 and 4-7 as 4 x 4 blocks of 16 rows,
// columns 8-9 as a 2 x 4 block of 8 rows
__device__ __forceinline__ int head8_row(int c, int g) { return c < 8 ? 16 * (c / 4) + 4 * g + c % 4 : 32 + 2 * g + (c - 8); }
__device__ __forceinline__ float quant_up(uint32_t E) { return __uint_as_float((254u + 29u - E) << 23); }      // 2^(29 - e)
__device__ __forceinline__ float quant_down(uint32_t E) { return __uint_as_float((E - 29u) << 23); }            // 2^(e - 29)
// four balanced base-256 digits of q, |q| <= 2^30: q = d0 + 2^8 d1 + 2^16 d2 + 2^24 d3, d0..d2 in [-128, 127], |d3| <= 64
__device__ __forceinline__ void digits4(int q, int8_t *d)
{
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int a = ((q + 128) & 255) - 128;
        d[g] = (int8_t)a;
        q = (q - a) >> 8;
    }
    d[3] = (int8_t)q;
}

inline unsigned grid_for(size_t work, unsigned block, unsigned cap = 148u * 16u) {
    size_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}

// ---- threshold.cu
void upload_A(Ctx &c, uint64_t V, uint64_t D, int64_t nnz, const float *vals, const void *rows,
              bool rows64, const int64_t *offsets, float avg, uint64_t nz_docs);
void compute_thresholds(Ctx &c, uint64_t k, float *zetas_out, int64_t *new_nnz_out);
void build_B(Ctx &c, const uint8_t *select, int64_t *nnzB, uint64_t *DB);
void sampling_weights(Ctx &c, float *out);
void sample_docs(Ctx &c, float sample_rate, uint64_t seed, uint8_t *select_out, uint64_t *n_selected_out);
void download_B(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig);
void download_B_begin(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *orig);
void download_B_end(Ctx &c);
void read_small(Ctx &c, void *dst, const void *dev_src, size_t bytes);   // synchronising small D2H read that bypasses the copy engine
float frobenius(Ctx &c);

// ---- ingest.cu (text -> entries -> CSC -> normalised A on the device)
void ingest_text(Ctx &c, const char *text, uint64_t size, uint64_t V, uint64_t D, int64_t max_entries, int64_t *nnz_out,
                 float *avg_out, uint64_t *nz_docs_out, uint64_t *tokens_out);
void upload_counts(Ctx &c, uint64_t V, uint64_t D, int64_t nnz, const uint32_t *counts, const uint32_t *rows, const int64_t *offsets,
                   float *avg_out, uint64_t *nz_docs_out);
void download_A(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets);

// ---- spmm.cu
void build_csr(Ctx &c, bool head_i8 = true);
// Z(V x b, column-major, ld=V) = B (B^T X); X column-major ld=V; both on the device.
void spsptr_multiply_dev(Ctx &c, int b, const float *X, float *Z);

// ---- spmm_head.cu (dense head of the operator on tcgen05)
int head_block_stride(int b);
int head_split_rows(int b);
void spmm_head_launch(Ctx &c, int b, const uint4 *bits, uint32_t num_mtiles, uint32_t NC, uint32_t nsplit,
                      const __nv_bfloat16 *split, float *out, uint32_t m_valid, bool zero_out, bool force_atomic,
                      cudaStream_t stream);

// ---- spmm_head_i8.cu (the same on tcgen05.mma.kind::i8: plane-masked u8 cells x three s8 digits, s32 accumulation)
void spmm_head_i8_launch(Ctx &c, int b, const uint4 *bits, uint32_t num_mtiles, uint32_t NSC, uint32_t nsplit,
                         const int8_t *digits, const uint32_t *colmax_bits, float *out, uint32_t m_valid, bool zero_out,
                         bool force_atomic, cudaStream_t stream);

// ---- panel_tc.cu (tall-skinny panel products of block Gram-Schmidt on tcgen05, split TF32)
struct PanelTc {
    int64_t n = 0;
    int ncv = 0;
    uint32_t seg = 16;                 // 32-k chunks per accumulation segment (= per partial result)
    size_t fpitch = 0, cpitch = 0;
    DevBuf<float> fsplit;              // [32][fpitch]: hi (rows 0-15) / lo (rows 16-31) of the block F, K-major
    DevBuf<float> csplit;              // [32][cpitch]: hi / lo of the coefficient block C
    DevBuf<float> part;                // partial results [segments][m][16]
    DevBuf<uint32_t> flags;            // largest |coefficient| of pass 0 / 1 (bit patterns), skip flag, elided-pass counter
    bool sharded = false;              // n = this rank's rows of a row-sharded basis: W^T F is all-reduced over the ranks
    DevBuf<float> packed;              // [ncv][16] coefficients on their way through the all-reduce
    void init(Ctx &c, int64_t n, int ncv);
    static bool usable(int64_t n);
    void begin_step(Ctx &c);
    void split_F(Ctx &c, const float *F, int b);
    // pass = index of the Gram-Schmidt pass inside the block step (0, 1 record their largest coefficient;
    // passes >= 2 return at once when decide_elision() set the skip flag)
    void wtf(Ctx &c, const float *W, int rows, int b, float *C, int ldc, float *Hk, int ldh, bool assign, int pass = 0);
    void fsub(Ctx &c, const float *W, int rows, int b, float *F, int pass = 0);
    void decide_elision(Ctx &c, float ratio);
    uint32_t elided(Ctx &c);
};

// ---- blockks.cu
void block_ks(Ctx &c, uint64_t k, int b, int max_restarts, float tol, uint64_t seed,
              float *evalues_out, float *U_out, int *nconv_out);
void panel_products(Ctx &c, int64_t n, int rows, int b, const float *W_host, float *F_host_inout, float *C_host_out, int engine);

// ---- kmeans.cu
void project(Ctx &c);
void kmeanspp(Ctx &c, uint64_t k, uint64_t seed, uint64_t *seeds_out, float *centers_out,
              float *residual_out);
void lloyd_projected(Ctx &c, uint64_t k, float *centers_inout, int max_reps,
                     uint32_t *assign_out, double *objective_out, int *iters_out);
void assign_projected(Ctx &c, uint64_t k, const float *centers, uint32_t *assign_out);
void update_min_dist(Ctx &c, uint64_t ncent, const float *centers, float *min_dist_inout);
void lift_centers(Ctx &c, uint64_t ncols, const float *in, uint64_t ld_in, float *out);

// ---- lloyd_full.cu (run_lloyds on the full-dimensional B)
void lloyd_full(Ctx &c, uint64_t k, float *centers_inout, int max_reps, uint32_t *assign_out,
                double *objective_out, int *iters_out);

// ---- catchwords.cu (rth_highest_element per cluster, find_catchwords)
void catchword_thresholds(Ctx &c, uint64_t k, uint64_t r, const uint32_t *cluster_of_doc_host, float *thresholds_out);
void rth_highest_element(Ctx &c, uint64_t r, const uint64_t *docs_host, uint64_t ndocs, float *thresholds_out);
void find_catchwords(Ctx &c, uint64_t k, const float *thresholds_host, double rho, int32_t *topic_of_word_out);

// ---- topic_model.cu (construct_topic_model)
void construct_topic_model(Ctx &c, uint64_t k, const int32_t *topic_of_word_host, const uint32_t *cluster_of_doc_host,
                           uint64_t rank_threshold, float *model_out, uint64_t *num_entries_out);
void download_doc_topic_sums(Ctx &c, uint32_t *docs, uint32_t *topics, float *sums);

// ---- dist_tc.cu (tcgen05 split-TF32 distance contraction)
void split_tf32(Ctx &c, const float *x, size_t n, float *hi, float *lo);
void split_lo_trunc(Ctx &c, const float *x, size_t n, float *lo);
void gemm_3xtf32(Ctx &c, int m, int n, int k, const float *A, int lda, const float *B, int ldb, float *C, int ldc);   // fp32-accurate GEMM on the tensor cores
bool dist_tc_supported(const Ctx &c, uint32_t kp, uint32_t ncent);
void dist_tc_launch(Ctx &c, const float *C, const float *c2, uint32_t ncent, int mode, uint32_t *assign, float *min_dist);

// ---- collectives (no-ops when world == 1)
void allreduce_sum_f32(Ctx &c, float *buf, size_t n);
void allreduce_sum_u32(Ctx &c, uint32_t *buf, size_t n);
void allreduce_sum_u64(Ctx &c, unsigned long long *buf, size_t n);
void allreduce_sum_f64(Ctx &c, double *buf, size_t n);
void allreduce_max_u32(Ctx &c, uint32_t *buf, size_t n);
void allgather_u64(Ctx &c, const unsigned long long *send, unsigned long long *recv);  // 1 per rank
void allgather_f64(Ctx &c, const double *send, double *recv);                         // 1 per rank
void bcast_f32(Ctx &c, float *buf, size_t n, int root);
void allgather_f32(Ctx &c, const float *send, float *recv, size_t count);   // `count` floats per rank
void p2p_destroy(Ctx &c);
void selftest_collectives(Ctx &c, unsigned long long *mismatches_out, int *p2p_active_out);   // P2P vs NCCL, bit for bit                                                    // unmaps / frees the peer workspace

}  // namespace isle
