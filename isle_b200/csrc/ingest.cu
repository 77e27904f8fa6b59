// ingest.cu -- SURVEY 8(f) row 3: text -> (doc, word, count) entries -> sorted, de-duplicated -> doc-major CSC ->
// normalised values, on the device.  Replaces what ISLETrainer does before the spectral core:
//   DocWordEntriesReader::fill_doc_word_entries (reference include/utils.h:160-228): parse `<doc> <word> <count>` lines,
//       1-based ids, blanks or tabs between the fields, optional '\r', a last line without '\n'
//   finalize_data (src/trainer.cpp:232-246): parallel_sort by (doc, word), std::unique on (doc, word)
//   SparseMatrix::populate_CSC (src/sparseMatrix.cpp:58-106): vals / rows / offsets, #tokens, #non-empty docs,
//       avg_doc_sz = (T)(total_word_count / nz_docs)  -- INTEGER division
//   SparseMatrix::normalize_docs (src/sparseMatrix.cpp:136-167): doc_sum = fp32 sum of the counts, value =
//       avg_doc_sz * (count / doc_sum) in fp32 with that association
// HBM-bound byte/integer work: one pass over the text (bytes), one 64-bit radix sort of the entries (the reference sorts
// 24-byte records with __gnu_parallel::sort), a flagged compaction, a binary-searched offset table, a warp-per-document
// normalisation.  Bit-exact: the token sum of a document is an integer, exact (hence order independent) in fp32 below 2^24
// -- checked on the device, larger documents are rejected (ISLE_ERR_RANGE) -- and the divide / multiply per entry are
// IEEE fp32 (no fast-math), the same two roundings the reference makes.
// Of duplicated (doc, word) lines the FIRST in file order survives (a stable sort; the reference's parallel sort leaves
// the survivor unspecified).  The result is left in the context as the uploaded A (ready for isle_cuda_thresholds) and
// can be copied out in the reference's host layout for the members that stay on the host (writers, metrics).
#include <cub/cub.cuh>

#include "common.cuh"

namespace isle {

namespace {

// number of '\n' in [0, size)
__global__ void __launch_bounds__(256)
count_newlines_kernel(const char *__restrict__ buf, size_t size, unsigned long long *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long n = 0;
    for (; i < size; i += stride) n += buf[i] == '\n';
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const unsigned long long t = BR(tmp).Sum(n);
    if (threadIdx.x == 0 && t) atomicAdd(out, t);
}

struct IsNewline {
    const char *buf;
    __host__ __device__ bool operator()(const size_t &i) const { return buf[i] == '\n'; }
};

// One thread per line: [begin, end) without the terminating '\n'.  Mirrors the reference's state machine for well-formed
// lines: digits accumulate into field 1, 2, 3; blanks / tabs separate fields (several allowed); '\r' is ignored.  A line
// that does not hold exactly three numbers, or any other character, sets the error flag (the reference asserts / prints
// "Bad format").
__global__ void __launch_bounds__(256)
parse_lines_kernel(const char *__restrict__ buf, size_t size, const size_t *__restrict__ nl_pos, size_t nlines, size_t n_newlines,
                   unsigned long long *__restrict__ key, uint32_t *__restrict__ count, uint32_t *__restrict__ order,
                   unsigned long long V, unsigned long long D, int *__restrict__ err)
{
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const size_t b = l == 0 ? 0 : nl_pos[l - 1] + 1;
    const size_t e = l < n_newlines ? nl_pos[l] : size;
    unsigned long long f[3] = {0ull, 0ull, 0ull};
    int state = 0;
    bool in_number = false, bad = false;
    for (size_t i = b; i < e; ++i) {
        const char ch = buf[i];
        if (ch >= '0' && ch <= '9') {
            if (!in_number) { in_number = true; if (++state > 3) { bad = true; break; } }
            f[state - 1] = f[state - 1] * 10ull + (unsigned long long)(ch - '0');
        } else if (ch == ' ' || ch == '\t') {
            in_number = false;
        } else if (ch == '\r') {
            in_number = false;
        } else {
            bad = true;
            break;
        }
    }
    if (state != 3 || f[0] == 0 || f[1] == 0 || f[0] > D || f[1] > V || f[2] > 0xFFFFFFFFull) bad = true;
    if (bad) {
        atomicOr(err, 1);
        key[l] = ~0ull;
        count[l] = 0;
        order[l] = (uint32_t)l;
        return;
    }
    key[l] = ((f[0] - 1ull) << 32) | (f[1] - 1ull);     // 1-based -> 0-based (utils.h:172-173)
    count[l] = (uint32_t)f[2];
    order[l] = (uint32_t)l;
}

// keep[i] = first entry of its (doc, word) run in the sorted array (std::unique, trainer.cpp:243-246)
__global__ void __launch_bounds__(256)
flag_first_kernel(const unsigned long long *__restrict__ key, size_t n, uint8_t *__restrict__ keep)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// rows / raw counts of the de-duplicated entries; off[d] = first entry with doc >= d (lower bound)
__global__ void __launch_bounds__(256)
unpack_entries_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ cnt, size_t n, uint32_t *__restrict__ row,
                      float *__restrict__ val)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    row[i] = (uint32_t)(key[i] & 0xFFFFFFFFull);
    val[i] = (float)cnt[i];                  // vals_CSC[pos] = (T)entries[pos].count (sparseMatrix.cpp:69)
}

__global__ void __launch_bounds__(256)
doc_offsets_kernel(const unsigned long long *__restrict__ key, size_t n, uint32_t D, int64_t *__restrict__ off)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > D) return;
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if ((key[mid] >> 32) < (unsigned long long)d) lo = mid + 1; else hi = mid;
    }
    off[d] = (int64_t)lo;
}

// tokens (sum of counts, u64) and non-empty documents; flags documents whose token sum is not exact in fp32
__global__ void __launch_bounds__(256)
doc_stats_kernel(const uint32_t *__restrict__ cnt, const int64_t *__restrict__ off, uint32_t D, unsigned long long *__restrict__ tokens,
                 unsigned long long *__restrict__ nz_docs, int *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    unsigned long long tok = 0, nz = 0;
    for (; d < D; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        unsigned long long s = 0;
        for (int64_t p = b + lane; p < e; p += 32) s += cnt[p];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            tok += s;
            nz += e > b;
            if (s >= (1ull << 24)) atomicOr(err, 2);
        }
    }
    if (lane == 0) {
        if (tok) atomicAdd(tokens, tok);
        if (nz) atomicAdd(nz_docs, nz);
    }
}

// normalized = avg * (count / doc_sum), fp32, that association (sparseMatrix.cpp:145-157); doc_sum is an exact integer
__global__ void __launch_bounds__(256)
normalize_docs_kernel(const uint32_t *__restrict__ cnt, const int64_t *__restrict__ off, uint32_t D, float avg, float *__restrict__ val)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < D; d += nw) {
        const int64_t b = off[d], e = off[d + 1];
        uint32_t s = 0;
        for (int64_t p = b + lane; p < e; p += 32) s += cnt[p];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float doc_sum = (float)s;          // exact below 2^24 (checked by doc_stats_kernel)
        for (int64_t p = b + lane; p < e; p += 32) val[p] = __fmul_rn(avg, __fdiv_rn((float)cnt[p], doc_sum));
    }
}

// key[p] = doc << 32 | row[p] for the entries of a doc-major CSC (warp per document)
__global__ void __launch_bounds__(256)
csc_keys_kernel(const int64_t *__restrict__ off, uint32_t D, const uint32_t *__restrict__ row, unsigned long long *__restrict__ key)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (; d < D; d += nw)
        for (int64_t p = off[d] + lane, e = off[d + 1]; p < e; p += 32) key[p] = ((unsigned long long)d << 32) | row[p];
}

__global__ void widen_rows_kernel(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace

// Sorted, de-duplicated entries (key = doc << 32 | word, raw counts) -> A in the context.  Shared by the text path and
// the entry-array path.  Returns nnz.
static int64_t finish_ingest(Ctx &c, uint64_t V, uint64_t D, DevBuf<unsigned long long> &key, DevBuf<uint32_t> &cnt, size_t n,
                             float *avg_out, uint64_t *nz_docs_out, uint64_t *tokens_out)
{
    c.V = V; c.D = D; c.nnzA = (int64_t)n;
    c.have_zeta = c.have_B = c.have_csr = c.have_U = c.have_P = false;
    c.a_val.alloc(n);
    c.a_row.alloc(n);
    c.a_off.alloc((size_t)D + 1);
    const uint32_t Du = (uint32_t)D;
    if (n) {
        unpack_entries_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(key.p, cnt.p, n, c.a_row.p, c.a_val.p);
        count_launch(c);
    }
    doc_offsets_kernel<<<(Du + 1 + 255) / 256, 256, 0, c.stream>>>(key.p, n, Du, c.a_off.p);
    DevBuf<unsigned long long> stats(2);
    DevBuf<int> err(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(stats.p, 0, 16, c.stream));
    ISLE_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), c.stream));
    const unsigned wgrid = grid_for((size_t)std::max<uint32_t>(Du, 1) * 32, 256, c.num_sms * 16);
    if (Du) doc_stats_kernel<<<wgrid, 256, 0, c.stream>>>(cnt.p, c.a_off.p, Du, stats.p, stats.p + 1, err.p);
    count_launch(c, 2);
    unsigned long long hs[2] = {0, 0};
    int herr = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(hs, stats.p, 16, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    ISLE_REQUIRE(!(herr & 2), ISLE_ERR_RANGE,
                 "ingest: a document holds 2^24 tokens or more: its fp32 token sum would no longer be exact (order dependent)");
    unsigned long long tokens = hs[0], nz = hs[1];
    if (c.world > 1) {      // avg_doc_sz and nz_docs are statistics of the whole corpus
        ISLE_CUDA_CHECK(cudaMemcpyAsync(stats.p, hs, 16, cudaMemcpyHostToDevice, c.stream));
        allreduce_sum_u64(c, stats.p, 2);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(hs, stats.p, 16, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        tokens = hs[0]; nz = hs[1];
    }
    ISLE_REQUIRE(nz > 0, ISLE_ERR_ARG, "ingest: no non-empty document");
    const float avg = (float)(tokens / nz);      // integer division, then the cast (sparseMatrix.cpp:98)
    ISLE_REQUIRE(avg >= 1.0f && avg < 1.0e6f, ISLE_ERR_RANGE, "ingest: avg_doc_sz out of range");
    if (Du) {
        StatScope s(c, "ingest_normalize", (double)n * 12.0 + (double)D * 8.0);
        normalize_docs_kernel<<<wgrid, 256, 0, c.stream>>>(cnt.p, c.a_off.p, Du, avg, c.a_val.p);
        count_launch(c);
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.avg_doc_sz = avg;
    c.nz_docs = nz;
    if (avg_out) *avg_out = avg;
    if (nz_docs_out) *nz_docs_out = nz;
    if (tokens_out) *tokens_out = tokens;
    return (int64_t)n;
}

// sort by (doc, word) keeping file order among equals, then drop all but the first of each (doc, word)
static size_t sort_and_dedupe(Ctx &c, DevBuf<unsigned long long> &key, DevBuf<uint32_t> &cnt, size_t n, int key_bits_doc)
{
    if (!n) return 0;
    DevBuf<unsigned long long> key2(n);
    DevBuf<uint32_t> cnt2(n);
    {
        StatScope s(c, "ingest_sort", (double)n * 12.0 * 2.0);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, cnt.p, cnt2.p, (int64_t)n, 0, 32 + key_bits_doc, c.stream);
        DevBuf<uint8_t> tmp(tb);
        ISLE_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, cnt.p, cnt2.p, (int64_t)n, 0, 32 + key_bits_doc, c.stream));
        count_launch(c);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    DevBuf<uint8_t> keep(n);
    flag_first_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(key2.p, n, keep.p);
    DevBuf<unsigned long long> nsel(1);
    size_t tb = 0, tb2 = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, key2.p, keep.p, key.p, nsel.p, (int64_t)n, c.stream);
    cub::DeviceSelect::Flagged(nullptr, tb2, cnt2.p, keep.p, cnt.p, nsel.p, (int64_t)n, c.stream);
    DevBuf<uint8_t> tmp(std::max(tb, tb2));
    ISLE_CUDA_CHECK(cub::DeviceSelect::Flagged(tmp.p, tb, key2.p, keep.p, key.p, nsel.p, (int64_t)n, c.stream));
    ISLE_CUDA_CHECK(cub::DeviceSelect::Flagged(tmp.p, tb2, cnt2.p, keep.p, cnt.p, nsel.p, (int64_t)n, c.stream));
    count_launch(c, 3);
    unsigned long long h = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&h, nsel.p, 8, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    return (size_t)h;
}

static int bits_for(uint64_t x) { int b = 1; while ((1ull << b) < x) ++b; return b; }

void ingest_text(Ctx &c, const char *text, uint64_t size, uint64_t V, uint64_t D, int64_t max_entries, int64_t *nnz_out,
                 float *avg_out, uint64_t *nz_docs_out, uint64_t *tokens_out)
{
    ISLE_REQUIRE(text && V > 0 && V < (1ull << 32) && D > 0 && D < (1ull << 32), ISLE_ERR_ARG, "ingest_text: bad arguments");
    StatScope total(c, "ingest", (double)size);
    DevBuf<char> buf(std::max<uint64_t>(size, 1));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(buf.p, text, size, cudaMemcpyHostToDevice, c.stream));
    // lines: one per '\n', plus a last line without one (utils.h:219-225)
    DevBuf<unsigned long long> nnl(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(nnl.p, 0, 8, c.stream));
    if (size) count_newlines_kernel<<<grid_for(size, 256, c.num_sms * 16), 256, 0, c.stream>>>(buf.p, size, nnl.p);
    count_launch(c);
    unsigned long long n_newlines = 0;
    char last = '\n';
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&n_newlines, nnl.p, 8, cudaMemcpyDeviceToHost, c.stream));
    if (size) ISLE_CUDA_CHECK(cudaMemcpyAsync(&last, buf.p + size - 1, 1, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    const size_t nlines = (size_t)n_newlines + ((size && last != '\n') ? 1 : 0);
    // the reference sizes `entries` to max_entries and asserts nRead == max_entries (utils.h:166, 227)
    ISLE_REQUIRE(max_entries <= 0 || (size_t)max_entries == nlines, ISLE_ERR_ARG,
                 "ingest_text: max_entries (" + std::to_string(max_entries) + ") is not the number of lines (" + std::to_string(nlines) + ")");
    DevBuf<size_t> nl_pos(std::max<size_t>(n_newlines, 1));
    if (n_newlines) {
        cub::CountingInputIterator<size_t> idx(0);
        IsNewline pred{buf.p};
        size_t tb = 0;
        cub::DeviceSelect::If(nullptr, tb, idx, nl_pos.p, nnl.p, (int64_t)size, pred, c.stream);
        DevBuf<uint8_t> tmp(tb);
        ISLE_CUDA_CHECK(cub::DeviceSelect::If(tmp.p, tb, idx, nl_pos.p, nnl.p, (int64_t)size, pred, c.stream));
        count_launch(c);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    DevBuf<unsigned long long> key(std::max<size_t>(nlines, 1));
    DevBuf<uint32_t> cnt(std::max<size_t>(nlines, 1)), order(std::max<size_t>(nlines, 1));
    DevBuf<int> err(1);
    ISLE_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), c.stream));
    if (nlines) {
        StatScope s(c, "ingest_parse", (double)size + (double)nlines * 16.0);
        parse_lines_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, c.stream>>>(buf.p, size, nl_pos.p, nlines, (size_t)n_newlines, key.p,
                                                                                    cnt.p, order.p, V, D, err.p);
        count_launch(c);
    }
    int herr = 0;
    ISLE_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    ISLE_REQUIRE(herr == 0, ISLE_ERR_RANGE,
                 "ingest_text: bad line (not `<doc> <word> <count>` with 1 <= doc <= D, 1 <= word <= V; the reference prints Bad format)");
    buf.release();
    const size_t n = sort_and_dedupe(c, key, cnt, nlines, bits_for(D));
    const int64_t nnz = finish_ingest(c, V, D, key, cnt, n, avg_out, nz_docs_out, tokens_out);
    if (nnz_out) *nnz_out = nnz;
}

// Copies the ingested A out in the reference's host layout: normalized_vals_CSC f32[nnz], rows_CSC u64[nnz],
// offsets_CSC i64[D+1]; any pointer may be NULL.
void download_A(Ctx &c, float *vals, uint64_t *rows, int64_t *offsets)
{
    ISLE_REQUIRE(c.a_off.p != nullptr, ISLE_ERR_ARG, "download_A: nothing uploaded or ingested");
    const size_t n = (size_t)c.nnzA;
    if (vals && n) ISLE_CUDA_CHECK(cudaMemcpyAsync(vals, c.a_val.p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    if (offsets) ISLE_CUDA_CHECK(cudaMemcpyAsync(offsets, c.a_off.p, ((size_t)c.D + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    if (rows && n) {
        DevBuf<unsigned long long> wide(n);
        widen_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(c.a_row.p, wide.p, n);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(rows, wide.p, n * 8, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

// populate_CSC + normalize_docs for a CSC of raw counts that is already sorted and de-duplicated (what ISLETrainer holds
// after finalize_data's sort): the a2 row of SURVEY 8 on the device, with the 2^24 exactness check.
void upload_counts(Ctx &c, uint64_t V, uint64_t D, int64_t nnz, const uint32_t *counts, const uint32_t *rows, const int64_t *offsets,
                   float *avg_out, uint64_t *nz_docs_out)
{
    ISLE_REQUIRE(V > 0 && V < (1ull << 32) && D < (1ull << 32) && nnz >= 0 && offsets, ISLE_ERR_ARG, "upload_counts: bad arguments");
    const size_t n = (size_t)nnz;
    DevBuf<uint32_t> cnt(std::max<size_t>(n, 1)), row(std::max<size_t>(n, 1));
    DevBuf<int64_t> off((size_t)D + 1);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(cnt.p, counts, n * 4, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(row.p, rows, n * 4, cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(off.p, offsets, ((size_t)D + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    // rebuild the (doc, word) keys from the CSC so that finish_ingest applies unchanged
    DevBuf<unsigned long long> key(std::max<size_t>(n, 1));
    if (D && n) {
        csc_keys_kernel<<<grid_for((size_t)D * 32, 256, c.num_sms * 16), 256, 0, c.stream>>>(off.p, (uint32_t)D, row.p, key.p);
        count_launch(c);
    }
    finish_ingest(c, V, D, key, cnt, n, avg_out, nz_docs_out, nullptr);
}

}  // namespace isle
