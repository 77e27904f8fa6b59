// sparseMatrix_cuda.cpp -- the reference-side binding of libisle_cuda.
//
// This translation unit is compiled WITH the reference's own headers (include/sparseMatrix.h)
// and linked into ISLETrain next to the reference's unmodified objects.  It supplies, as
// explicit specialisations for FPTYPE = float, exactly the member functions that
// ISLETrainer::train() calls on the spectral core (reference src/trainer.cpp:430-554); each one
// forwards to the C ABI in include/isle_cuda.h.  Explicit specialisations are ordinary (strong)
// symbols, while the reference's `template class ISLE::FPSparseMatrix<float>;`
// (src/sparseMatrix.cpp:2495-2510) emits weak ones, so the linker binds trainer.o to the
// functions below and everything else (loaders, metrics, output
// writers) keeps running the reference's host code on the arrays this file fills.
//
//   reference member (include/sparseMatrix.h)            line   C ABI entry point
//   SparseMatrix::list_word_freqs_by_sorting              :116   (no-op: freqs is only a hand-off)
//   SparseMatrix::compute_thresholds                      :128   isle_cuda_upload_A + isle_cuda_thresholds
//   FPSparseMatrix::threshold_and_copy<float>             :278   isle_cuda_build_B + isle_cuda_download_B
//   FPSparseMatrix::sampled_threshold_and_copy<float>     :297   isle_cuda_sampling_weights + build_B(mask)
//   FPSparseMatrix::frobenius                             :242   isle_cuda_frobenius
//   FPSparseMatrix::initialize_for_eigensolver            :258   (allocates U_colmajor as the reference does)
//   FPSparseMatrix::compute_block_ks                      :266   isle_cuda_block_ks
//   FPSparseMatrix::kmeans_init_on_projected_space        :454   isle_cuda_kmeanspp
//   FPSparseMatrix::run_lloyds_on_projected_space         :434   isle_cuda_lloyd_projected
//   FPSparseMatrix::left_multiply_by_U_Spectra            :309   isle_cuda_lift_centers
//   FPSparseMatrix::cleanup_after_eigensolver             :260   isle_cuda_cleanup_eigensolver
//   FPSparseMatrix::run_lloyds          :370   isle_cuda_lloyd_full
//   SparseMatrix::rth_highest_element   :138   isle_cuda_rth_highest_element
//   SparseMatrix::find_catchwords       :157   isle_cuda_find_catchwords
//   SparseMatrix::construct_topic_model (8f row 2)        :162   isle_cuda_construct_topic_model + isle_cuda_doc_topic_sums
//
// Ownership follows the reference: every host array is new[]-allocated here to the size the
// reference would have used and filled by the library; device memory belongs to the context.
// Errors become std::runtime_error, which main()'s catch-all reports as "ISLE Trainer failed"
// (drivers/ISLETrain.cpp:48-50).  train() is single-threaded, so one process-wide context.
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <tuple>
#include <string>
#include <vector>

#include "sparseMatrix.h"

#include "isle_cuda.h"

namespace {

isle_cuda_ctx *g_ctx = nullptr;

isle_cuda_ctx *ctx()
{
    if (!g_ctx) {
        // ISLE_CUDA_NGPUS=n (or ISLE_CUDA_DEVICES=0,1,...): every GPU of the box behind the one context train() uses --
        // the library shards the documents itself (isle_cuda_create_multi); default: one GPU, ISLE_CUDA_DEVICE
        const char *dev = std::getenv("ISLE_CUDA_DEVICE");
        const char *ngpus = std::getenv("ISLE_CUDA_NGPUS");
        const char *list = std::getenv("ISLE_CUDA_DEVICES");
        int rc;
        if (list && *list) {
            std::vector<int> devs;
            for (const char *p = list; *p;) {
                devs.push_back(std::atoi(p));
                while (*p && *p != ',') ++p;
                if (*p == ',') ++p;
            }
            rc = isle_cuda_create_multi(&g_ctx, (int)devs.size(), devs.data());
        } else if (ngpus && std::atoi(ngpus) > 1) {
            rc = isle_cuda_create_multi(&g_ctx, std::atoi(ngpus), nullptr);
        } else {
            rc = isle_cuda_create(&g_ctx, dev ? std::atoi(dev) : 0);
        }
        if (rc != ISLE_OK)
            throw std::runtime_error(std::string("libisle_cuda: ") + isle_cuda_last_error(nullptr));
    }
    return g_ctx;
}

void check(int rc, const char *what)
{
    if (rc != ISLE_OK)
        throw std::runtime_error(std::string("libisle_cuda: ") + what + ": " + isle_cuda_last_error(g_ctx));
}

}  // namespace

namespace ISLE
{
    // ---- stage A ----------------------------------------------------------------------------
    template<>
    void SparseMatrix<float>::list_word_freqs_by_sorting(std::vector<A_TYPE>*)
    {
        // The device path selects thresholds from per-word histograms of the rounded values;
        // the word-major lists the reference builds here are never needed.
    }

    template<>
    offset_t SparseMatrix<float>::compute_thresholds(
        word_id_t word_begin,
        word_id_t word_end,
        std::vector<A_TYPE> *const,
        std::vector<float>& zetas,
        const doc_id_t num_topics)
    {
        if (word_begin != 0 || word_end != vocab_size())
            throw std::runtime_error("libisle_cuda: thresholds are computed for the whole vocabulary at once");
        check(isle_cuda_upload_A(ctx(), vocab_size(), num_docs(), get_nnzs(), normalized_vals_CSC,
                                 (const uint64_t *)rows_CSC, (const int64_t *)offsets_CSC,
                                 (float)avg_doc_sz, (uint64_t)_nz_docs), "upload_A");
        zetas.resize(vocab_size());
        int64_t new_nnzs = 0;
        check(isle_cuda_thresholds(ctx(), num_topics, zetas.data(), &new_nnzs), "thresholds");
        return (offset_t)new_nnzs;
    }

    // ---- stage G: catchword thresholds and catchwords ----------------------
    // rth_highest_element is called once per topic from an OpenMP loop (src/trainer.cpp:587-589): the device
    // context is single-caller, so the calls are serialised here; each one is a few kernels over the
    // cluster's documents of the A matrix compute_thresholds() uploaded.
    template<>
    void SparseMatrix<float>::rth_highest_element(
        const MKL_UINT r,
        const std::vector<doc_id_t>& doc_partition,
        float *thresholds)
    {
        static std::mutex m;
        std::lock_guard<std::mutex> lock(m);
        static_assert(sizeof(doc_id_t) == sizeof(uint64_t), "the C ABI assumes the reference's ILP64 index types");
        check(isle_cuda_rth_highest_element(ctx(), (uint64_t)r, (const uint64_t *)doc_partition.data(),
                                            (uint64_t)doc_partition.size(), thresholds), "rth_highest_element");
    }

    template<>
    void SparseMatrix<float>::find_catchwords(
        const doc_id_t num_topics,
        const float *const thresholds,
        std::vector<word_id_t> *catchwords)
    {
        std::vector<int32_t> topic_of_word(vocab_size());
        check(isle_cuda_find_catchwords(ctx(), num_topics, thresholds, (double)rho_c, topic_of_word.data()), "find_catchwords");
        for (word_id_t word = 0; word < vocab_size(); ++word)        // per topic in ascending word order, as :575-592
            if (topic_of_word[word] >= 0)
                catchwords[topic_of_word[word]].push_back(word);
    }

    // ---- stage H: topic model ---------------------------------------------------
    // construct_topic_model (src/sparseMatrix.cpp:597-838), called at src/trainer.cpp:645-651.  The device builds
    // the (document, topic) catchword sums, the per-topic thresholds and the model; this function fills the three
    // host lists the trainer's writers read (catchword_topics, doc_topic_sum, top_topic_pairs) from what comes back.
    template<>
    void SparseMatrix<float>::construct_topic_model(
        DenseMatrix<FPTYPE>& Model,
        const doc_id_t num_topics,
        const std::vector<doc_id_t> *const closest_docs,
        const std::vector<word_id_t> *const catchwords,
        bool,
        std::vector<std::tuple<int, int, doc_id_t> >* top_topic_pairs,
        std::vector<std::pair<word_id_t, int> >* catchword_topics,
        std::vector<std::tuple<doc_id_t, doc_id_t, FPTYPE> >*  doc_topic_sum)
    {
        assert(Model.vocab_size() == vocab_size());
        assert(Model.num_docs() == num_topics);
        std::vector<int32_t> topic_of_word(vocab_size(), -1);
        for (doc_id_t topic = 0; topic < num_topics; ++topic)
            for (auto w : catchwords[topic]) topic_of_word[w] = (int32_t)topic;
        std::vector<uint32_t> cluster_of_doc(num_docs(), 0xFFFFFFFFu);
        for (doc_id_t topic = 0; topic < num_topics; ++topic)
            for (auto d : closest_docs[topic]) cluster_of_doc[d] = (uint32_t)topic;
        const uint64_t rank_threshold = (doc_id_t)(eps3_c*w0_c*(FPTYPE)num_docs() / ((FPTYPE)num_topics * 2.0));   // :716
        assert(rank_threshold > 0);
        uint64_t n = 0;
        check(isle_cuda_construct_topic_model(ctx(), num_topics, topic_of_word.data(), cluster_of_doc.data(), rank_threshold,
                                              Model.data(), &n), "construct_topic_model");
        if (catchword_topics != NULL)                                   // sorted by word (:619-621)
            for (word_id_t w = 0; w < vocab_size(); ++w)
                if (topic_of_word[w] >= 0) catchword_topics->push_back(std::make_pair(w, (int)topic_of_word[w]));
        if (doc_topic_sum == NULL && top_topic_pairs == NULL) return;
        std::vector<uint32_t> docs(n), topics(n);
        std::vector<float> sums(n);
        check(isle_cuda_doc_topic_sums(ctx(), docs.data(), topics.data(), sums.data()), "doc_topic_sums");
        if (top_topic_pairs != NULL) {                                  // :683-703: strict >, entries in topic order
            uint64_t i = 0;
            while (i < n) {
                const uint32_t doc = docs[i];
                float max = 0.0f, max2 = 0.0f;
                int max_topic = -1, max2_topic = -1;
                for (; i < n && docs[i] == doc; ++i) {
                    if (sums[i] > max) { max2 = max; max2_topic = max_topic; max = sums[i]; max_topic = (int)topics[i]; }
                    else if (sums[i] > max2) { max2 = sums[i]; max2_topic = (int)topics[i]; }
                }
                if (max_topic >= 0 && max2_topic >= 0)
                    top_topic_pairs->push_back(std::make_tuple(max_topic, max2_topic, (doc_id_t)doc));
            }
        }
        if (doc_topic_sum != NULL) {                                    // final order of the reference: (doc, topic) (:789-793)
            doc_topic_sum->reserve(n);
            for (uint64_t i = 0; i < n; ++i)
                doc_topic_sum->emplace_back((doc_id_t)docs[i], (doc_id_t)topics[i], sums[i]);
        }
        std::cout << "Size of doc_topic_sum array: " << n << std::endl;
    }

    // ---- stage B ----------------------------------------------------------------------------
    namespace {
        // background = true: vals / rows / offsets arrive while train() goes on (isle_cuda_download_B_begin); nothing on the
        // spectral core reads them, compute_block_ks() waits for the copy before it returns
        void build_and_download(FPSparseMatrix<float>& B, const uint8_t *mask, const offset_t nnzs,
                                std::vector<doc_id_t>& original_cols,
                                float *&vals, word_id_t *&rows, offset_t *&offsets,
                                offset_t &nnzs_out, doc_id_t &docs_out, bool background = false)
        {
            int64_t nnzB = 0;
            uint64_t DB = 0;
            check(isle_cuda_build_B(ctx(), mask, &nnzB, &DB), "build_B");
            if (nnzB > nnzs + 1000)     // the reference allocates nnzs + 1000 (src/sparseMatrix.cpp:1295-1296)
                throw std::runtime_error("libisle_cuda: B holds more entries than compute_thresholds announced");
            original_cols.resize(DB);
            static_assert(sizeof(doc_id_t) == sizeof(uint64_t) && sizeof(offset_t) == sizeof(int64_t),
                          "the C ABI assumes the reference's ILP64 index types");
            if (background) {
                // original_cols is read by train() right away (src/trainer.cpp:573-575): synchronously; the bulk in the background
                check(isle_cuda_download_B(ctx(), NULL, NULL, NULL, (uint64_t *)original_cols.data()), "download_B");
                check(isle_cuda_download_B_begin(ctx(), vals, (uint64_t *)rows, (int64_t *)offsets, NULL), "download_B_begin");
            } else {
                check(isle_cuda_download_B(ctx(), vals, (uint64_t *)rows, (int64_t *)offsets,
                                           (uint64_t *)original_cols.data()), "download_B");
            }
            nnzs_out = (offset_t)nnzB;
            docs_out = (doc_id_t)DB;
            (void)B;
        }
    }

    template<>
    template<>
    void FPSparseMatrix<float>::threshold_and_copy<float>(
        const SparseMatrix<float>& from,
        const std::vector<float>& zetas,
        const offset_t nnzs,
        std::vector<doc_id_t>& original_cols)
    {
        assert(vocab_size() == from.vocab_size() && num_docs() == from.num_docs());
        assert(original_cols.size() == 0); assert(zetas.size() == vocab_size());
        allocate(nnzs + 1000);                                  // src/sparseMatrix.cpp:1295-1296
        offset_t n = 0; doc_id_t d = 0;
        build_and_download(*this, nullptr, nnzs, original_cols, vals_CSC, rows_CSC, offsets_CSC, n, d, /*background=*/true);
        _num_docs = d;                                          // :1309
        std::cout << "Columns remaining after thresholding: " << d << "\n";
        _nnzs = n;                                              // :1320
    }

    template<>
    template<>
    void FPSparseMatrix<float>::sampled_threshold_and_copy<float>(
        const SparseMatrix<float>& from,
        const std::vector<float>& zetas,
        const offset_t nnzs,
        std::vector<doc_id_t>& original_cols,
        const float sample_rate)
    {
        assert(vocab_size() == from.vocab_size() && num_docs() == from.num_docs());
        assert(original_cols.size() == 0); assert(zetas.size() == vocab_size());
        allocate(nnzs + 1000);
        // weights on the device (:1383-1397); keys, pivot and selection exactly as :1399-1415,
        // but with the rand() draws taken serially (the reference races on rand() under pfor)
        const doc_id_t D = from.num_docs();
        std::vector<float> weights(D), dice(D);
        check(isle_cuda_sampling_weights(ctx(), weights.data()), "sampling_weights");
        for (doc_id_t doc = 0; doc < D; ++doc) {
            dice[doc] = weights[doc] == 0.0f ? 0.0f : (float)std::pow(rand_fraction(), 1 / weights[doc]);
            weights[doc] = dice[doc];
        }
        const size_t nth = (size_t)(sample_rate * (float)D);
        std::nth_element(dice.begin(), dice.begin() + nth, dice.end(), std::greater<float>());
        const float pivot = dice[nth];
        std::cout << "sampling docs: pivot: " << pivot << std::endl;
        std::vector<uint8_t> select(D);
        for (doc_id_t doc = 0; doc < D; ++doc) select[doc] = weights[doc] >= pivot;
        offset_t n = 0; doc_id_t d = 0;
        build_and_download(*this, select.data(), nnzs, original_cols, vals_CSC, rows_CSC, offsets_CSC, n, d);
        _num_docs = d;
        std::cout << "After sampling docs: cols remaining: " << d << "\n";
        _nnzs = n;
        this->shrink(n);                                        // :1428
    }

    template<>
    float FPSparseMatrix<float>::frobenius() const
    {
        float f = 0.0f;
        check(isle_cuda_frobenius(ctx(), &f), "frobenius");
        return f;
    }

    // ---- stage C ----------------------------------------------------------------------------
    template<>
    void FPSparseMatrix<float>::initialize_for_eigensolver(const doc_id_t num_topics)
    {
        U_rows = vocab_size();
        U_cols = num_topics;
        U_colmajor = new float[(size_t)num_topics * (size_t)vocab_size()];   // :1154
    }

    template<>
    void FPSparseMatrix<float>::compute_block_ks(
        const doc_id_t num_topics,
        std::vector<float>& evalues)
    {
        std::vector<float> ev(num_topics);
        int nconv = 0;
        const int rc = isle_cuda_block_ks(ctx(), num_topics, BLOCK_KS_BLOCK_SIZE, BLOCK_KS_MAX_ITERS,
                                          (float)BLOCK_KS_TOLERANCE, /*seed=*/0, ev.data(), U_colmajor, &nconv);
        // the reference asserts num_converged() == num_topics (:1207)
        check(rc, "block_ks");
        check(isle_cuda_download_B_end(ctx()), "download_B_end");   // B's host arrays (threshold_and_copy) are complete from here on
        for (doc_id_t i = 0; i < num_topics; ++i) evalues.push_back(ev[i]);
        // U_rowmajor only feeds the host projection path this library replaces; stages F-H never
        // read it, so it is left unallocated (the destructor handles NULL).
    }

    template<>
    void FPSparseMatrix<float>::cleanup_after_eigensolver()
    {
        assert(U_colmajor != NULL);
        delete[] U_colmajor;
        U_colmajor = NULL;
        check(isle_cuda_cleanup_eigensolver(ctx()), "cleanup_eigensolver");
    }

    // ---- stages D / E -----------------------------------------------------------------------
    template<>
    float FPSparseMatrix<float>::kmeans_init_on_projected_space(
        const int num_centers,
        const int max_reps,
        std::vector<doc_id_t>& best_seed,
        float *const best_centers_coords)
    {
        float best = FP_MAX;
        std::vector<uint64_t> seeds(num_centers);
        std::vector<float> coords((size_t)num_centers * num_centers);
        for (int rep = 0; rep < max_reps; ++rep) {
            float dist = 0.0f;
            check(isle_cuda_kmeanspp(ctx(), num_centers, (uint64_t)rep, seeds.data(), coords.data(), &dist), "kmeanspp");
            std::cout << "k-means init residual: " << dist << std::endl;
            if (dist < best) {
                best = dist;
                best_seed.assign(seeds.begin(), seeds.end());
                if (best_centers_coords)
                    std::copy(coords.begin(), coords.end(), best_centers_coords);
            }
        }
        return best;
    }

    template<>
    float FPSparseMatrix<float>::run_lloyds_on_projected_space(
        const doc_id_t num_centers,
        float *projected_centers,
        std::vector<doc_id_t> *closest_docs,
        const int max_reps)
    {
        std::vector<uint32_t> assign;
        if (closest_docs != NULL) {
            for (doc_id_t center = 0; center < num_centers; ++center)
                assert(closest_docs[center].size() == 0);
            assign.resize(num_docs());
        }
        check(isle_cuda_lloyd_projected(ctx(), num_centers, projected_centers, max_reps,
                                        closest_docs ? assign.data() : NULL, NULL, NULL), "lloyd_projected");
        if (closest_docs != NULL)
            for (doc_id_t d = 0; d < num_docs(); ++d)
                closest_docs[assign[d]].push_back(d);            // ascending doc ids, as :1966-1973
        return 0.0f;                                             // the reference's residual is always 0
    }

    template<>
    void FPSparseMatrix<float>::left_multiply_by_U_Spectra(
        float *const out,
        const float *in,
        const doc_id_t ld_in,
        const doc_id_t ncols)
    {
        assert(U_rows == (MKL_INT)vocab_size());
        assert(ld_in >= (doc_id_t)U_cols);
        check(isle_cuda_lift_centers(ctx(), ncols, in, ld_in, out), "lift_centers");
    }

    // ---- stage F ------------------------------------------------------------
    // run_lloyds (src/sparseMatrix.cpp:1679-1746), called from src/trainer.cpp:566 with the lifted
    // centers and closest_docs: the host `centers` array is the input and receives the result.
    template<>
    float FPSparseMatrix<float>::run_lloyds(
        const doc_id_t num_centers,
        float *centers,
        std::vector<doc_id_t> *closest_docs,
        const int max_reps)
    {
        std::vector<uint32_t> assign;
        if (closest_docs != NULL) {
            for (doc_id_t center = 0; center < num_centers; ++center)
                assert(closest_docs[center].size() == 0);
            assign.resize(num_docs());
        }
        int iters = 0;
        check(isle_cuda_lloyd_full(ctx(), num_centers, centers, max_reps,
                                   closest_docs ? assign.data() : NULL, NULL, &iters), "lloyd_full");
        std::cout << "Lloyd's on B: " << iters << " iterations on the device\n";
        if (closest_docs != NULL)
            for (doc_id_t d = 0; d < num_docs(); ++d)
                closest_docs[assign[d]].push_back(d);            // ascending doc ids, as :1652-1653
        return 0.0f;                                             // compute_residual is off in the reference (:1588)
    }
}
