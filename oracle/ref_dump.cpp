/*
 * oracle/ref_dump.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A tiny driver (ours) linked against the UNMODIFIED reference objects
 * (src/sparseMatrix.cpp, denseMatrix.cpp, utils.cpp, logger.cpp compiled where
 * they lie under /root/reference, see oracle/Makefile).  It runs the spectral
 * core exactly as ISLETrainer::train() does (reference src/trainer.cpp:430-554)
 * by calling the same public SparseMatrix / FPSparseMatrix members in the same
 * order, and dumps every intermediate the parity tests compare against:
 *
 *   stage 0  populate_CSC + normalize_docs     (trainer.cpp:291-293)
 *   stage A  list_word_freqs_by_sorting + compute_thresholds   (:434-435)
 *   stage B  threshold_and_copy                 (:482-483)   [or an injected doc mask]
 *   stage C  initialize_for_eigensolver + compute_block_ks     (:492-497)
 *   stage D  kmeans_init_on_projected_space     (:529-530)
 *   stage E  run_lloyds_on_projected_space + left_multiply_by_U_Spectra (:546-550)
 *   stage F  run_lloyds on the full-dimensional B   (:559-571)
 *   stage G  rth_highest_element per cluster + find_catchwords   (:573-639)
 *   stage H  construct_topic_model                                (:645-651)
 *
 * Input : <corpus.bin>  = int64 V, D, nnz ; int64 offsets[D+1] ; uint32 rows[nnz] ;
 *                         uint32 counts[nnz]      (doc-major CSC of raw counts)
 * Output: <outdir>/<name>.bin raw little-endian arrays + meta.json with sizes and
 *         per-stage wall-clock seconds (used as the CPU baseline by bench.py).
 *
 * usage: ref_dump <corpus.bin> <outdir> <k> [--upto A|B|C|D|E|F|G|H] [--mask mask.u8]
 *                 [--centers centers_lowd.f32]   (override k-means++ seeds for Lloyd)
 *                 [--lloyd-iters n] [--srand seed]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "sparseMatrix.h"
#include "hyperparams.h"
#include "denseMatrix.h"
#include <tuple>

using namespace ISLE;

static double now_s()
{
    return std::chrono::duration<double>(
               std::chrono::high_resolution_clock::now().time_since_epoch())
        .count();
}

template <class T>
static void dump(const std::string &dir, const char *name, const T *p, size_t n)
{
    std::ofstream f(dir + "/" + name + ".bin", std::ios::binary);
    f.write((const char *)p, (std::streamsize)(n * sizeof(T)));
    if (!f) {
        std::fprintf(stderr, "ref_dump: write failed for %s\n", name);
        std::exit(2);
    }
}

/* Subclass only to reach protected bookkeeping when the harness injects a doc mask.
 * threshold_and_copy_doc_block (src/sparseMatrix.cpp:1328-1361) filters every doc
 * independently, so "filter all docs with the reference's threshold_and_copy, then
 * drop the unselected columns" yields exactly what sampled_threshold_and_copy
 * (src/sparseMatrix.cpp:1417-1430) builds for the same selection -- without libc rand(). */
struct MaskedB : public FPSparseMatrix<FPTYPE> {
    MaskedB(word_id_t v, doc_id_t d) : FPSparseMatrix<FPTYPE>(v, d) {}
    void drop_unselected(std::vector<doc_id_t> &original_cols, const bool *mask)
    {
        offset_t w = 0;
        doc_id_t nd = 0;
        for (doc_id_t d = 0; d < num_docs(); ++d) {
            const offset_t b = offsets_CSC[d], e = offsets_CSC[d + 1];
            if (!mask[original_cols[d]]) continue;
            for (offset_t p = b; p < e; ++p, ++w) {
                vals_CSC[w] = vals_CSC[p];
                rows_CSC[w] = rows_CSC[p];
            }
            original_cols[nd] = original_cols[d];
            offsets_CSC[nd + 1] = w; /* nd <= d, offsets_CSC[d+1] already consumed */
            ++nd;
        }
        original_cols.resize(nd);
        _num_docs = nd;
        _nnzs = w;
    }
};

int main(int argc, char **argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: ref_dump corpus.bin outdir k [--upto X] [--mask f] "
                             "[--centers f] [--lloyd-iters n] [--srand s]\n");
        return 1;
    }
    const std::string corpus = argv[1], out = argv[2];
    const doc_id_t k = (doc_id_t)std::atol(argv[3]);
    char upto = 'H';
    std::string mask_file, centers_file;
    int lloyd_iters = MAX_KMEANS_LOWD_REPS;
    for (int i = 4; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--upto") && i + 1 < argc) upto = argv[++i][0];
        else if (!std::strcmp(argv[i], "--mask") && i + 1 < argc) mask_file = argv[++i];
        else if (!std::strcmp(argv[i], "--centers") && i + 1 < argc) centers_file = argv[++i];
        else if (!std::strcmp(argv[i], "--lloyd-iters") && i + 1 < argc) lloyd_iters = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--srand") && i + 1 < argc) std::srand((unsigned)std::atol(argv[++i]));
    }

    /* ---- load corpus ---- */
    int64_t hdr[3];
    std::ifstream in(corpus, std::ios::binary);
    if (!in.read((char *)hdr, sizeof(hdr))) { std::fprintf(stderr, "bad corpus\n"); return 2; }
    const int64_t V = hdr[0], D = hdr[1], nnz = hdr[2];
    std::vector<int64_t> offs((size_t)D + 1);
    std::vector<uint32_t> rows((size_t)nnz), counts((size_t)nnz);
    in.read((char *)offs.data(), (std::streamsize)(sizeof(int64_t) * (D + 1)));
    in.read((char *)rows.data(), (std::streamsize)(sizeof(uint32_t) * nnz));
    in.read((char *)counts.data(), (std::streamsize)(sizeof(uint32_t) * nnz));
    if (!in) { std::fprintf(stderr, "short corpus\n"); return 2; }

    std::vector<DocWordEntry<count_t>> entries;
    entries.reserve((size_t)nnz);
    for (int64_t d = 0; d < D; ++d)
        for (int64_t p = offs[d]; p < offs[d + 1]; ++p)
            entries.emplace_back((word_id_t)rows[p], (doc_id_t)d, (count_t)counts[p]);
    std::vector<int64_t>().swap(offs);
    std::vector<uint32_t>().swap(rows);
    std::vector<uint32_t>().swap(counts);

    FILE *meta = std::fopen((out + "/meta.json").c_str(), "w");
    std::fprintf(meta, "{\"V\": %lld, \"D\": %lld, \"nnz\": %lld, \"k\": %lld",
                 (long long)V, (long long)D, (long long)nnz, (long long)k);

    /* ---- stage 0: CSC + normalise (trainer.cpp:291-293) ---- */
    double t0 = now_s();
    auto A_sp = new SparseMatrix<A_TYPE>((word_id_t)V, (doc_id_t)D);
    A_sp->populate_CSC(entries);
    A_sp->normalize_docs(true);
    std::vector<DocWordEntry<count_t>>().swap(entries);
    double t_norm = now_s() - t0;
    {
        std::vector<float> nv((size_t)nnz);
        for (int64_t p = 0; p < nnz; ++p) nv[p] = A_sp->normalized_val_CSC(p);
        dump(out, "A_normalized_vals", nv.data(), nv.size());
        float avg = A_sp->get_avg_doc_sz();
        dump(out, "A_avg_doc_sz", &avg, 1);
    }
    std::fprintf(meta, ", \"t_normalize\": %.6f", t_norm);

    /* ---- stage A: thresholds (trainer.cpp:430-472) ---- */
    t0 = now_s();
    std::vector<A_TYPE> thresholds((size_t)V, 0);
    auto freqs = new std::vector<A_TYPE>[V];
    A_sp->list_word_freqs_by_sorting(freqs);
    offset_t new_nnzs = A_sp->compute_thresholds(0, (word_id_t)V, freqs, thresholds, k);
    delete[] freqs;
    double t_thr = now_s() - t0;
    dump(out, "zetas", thresholds.data(), thresholds.size());
    std::fprintf(meta, ", \"new_nnzs\": %lld, \"t_thresholds\": %.6f", (long long)new_nnzs, t_thr);
    if (upto == 'A') { std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage B: thresholded matrix (trainer.cpp:475-485) ---- */
    t0 = now_s();
    std::vector<doc_id_t> original_cols;
    MaskedB *B = new MaskedB((word_id_t)V, (doc_id_t)D);
    B->threshold_and_copy<A_TYPE>(*A_sp, thresholds, new_nnzs, original_cols);
    if (!mask_file.empty()) {
        std::vector<char> m((size_t)D);
        std::ifstream mf(mask_file, std::ios::binary);
        mf.read(m.data(), D);
        bool *mask = new bool[D];
        for (int64_t d = 0; d < D; ++d) mask[d] = m[d] != 0;
        B->drop_unselected(original_cols, mask);
        delete[] mask;
    }
    double t_B = now_s() - t0;
    const int64_t DB = (int64_t)B->num_docs(), nnzB = (int64_t)B->get_nnzs();
    {
        std::vector<float> bv((size_t)nnzB);
        std::vector<uint64_t> br((size_t)nnzB);
        std::vector<int64_t> bo((size_t)DB + 1);
        for (int64_t p = 0; p < nnzB; ++p) { bv[p] = B->val_CSC(p); br[p] = B->row_CSC(p); }
        for (int64_t d = 0; d <= DB; ++d) bo[d] = B->offset_CSC((doc_id_t)d);
        dump(out, "B_vals", bv.data(), bv.size());
        dump(out, "B_rows", br.data(), br.size());
        dump(out, "B_offsets", bo.data(), bo.size());
        std::vector<uint64_t> oc(original_cols.begin(), original_cols.end());
        dump(out, "B_original_cols", oc.data(), oc.size());
    }
    float frob = B->frobenius();
    std::fprintf(meta, ", \"D_B\": %lld, \"nnz_B\": %lld, \"frobenius\": %.9g, \"t_build_B\": %.6f",
                 (long long)DB, (long long)nnzB, (double)frob, t_B);
    if (upto == 'B') { std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage C: block Krylov-Schur (trainer.cpp:490-502) ---- */
    t0 = now_s();
    std::vector<FPTYPE> evalues;
    B->initialize_for_eigensolver(k);
    B->compute_block_ks(k, evalues);
    double t_ks = now_s() - t0;
    dump(out, "evalues", evalues.data(), evalues.size());
    {
        /* U itself = left_multiply_by_U_Spectra(I_k)  (src/sparseMatrix.cpp:1438-1450) */
        std::vector<float> I((size_t)k * k, 0.0f), U((size_t)V * k);
        for (doc_id_t i = 0; i < k; ++i) I[(size_t)i * k + i] = 1.0f;
        B->left_multiply_by_U_Spectra(U.data(), I.data(), k, k);
        dump(out, "U_colmajor", U.data(), U.size());
    }
    std::fprintf(meta, ", \"t_block_ks\": %.6f", t_ks);
    if (upto == 'C') { std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage D: k-means++ on the projection (trainer.cpp:511-533) ---- */
    t0 = now_s();
    std::vector<doc_id_t> seeds;
    std::vector<FPTYPE> centers_lowd((size_t)k * k);
    float init_res = B->kmeans_init_on_projected_space((int)k, KMEANS_INIT_REPS, seeds,
                                                       centers_lowd.data());
    double t_pp = now_s() - t0;
    {
        std::vector<uint64_t> s(seeds.begin(), seeds.end());
        dump(out, "seeds", s.data(), s.size());
        dump(out, "centers_lowd_init", centers_lowd.data(), centers_lowd.size());
    }
    std::fprintf(meta, ", \"kmeanspp_residual\": %.9g, \"t_kmeanspp\": %.6f", (double)init_res, t_pp);
    if (upto == 'D') { std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage E: Lloyd on the projection + lift (trainer.cpp:539-553) ---- */
    if (!centers_file.empty()) {
        std::ifstream cf(centers_file, std::ios::binary);
        cf.read((char *)centers_lowd.data(), (std::streamsize)(sizeof(float) * k * k));
        dump(out, "centers_lowd_init", centers_lowd.data(), centers_lowd.size());
    }
    t0 = now_s();
    auto closest = new std::vector<doc_id_t>[k];
    B->run_lloyds_on_projected_space(k, centers_lowd.data(), closest, lloyd_iters);
    std::vector<float> centers((size_t)V * k);
    B->left_multiply_by_U_Spectra(centers.data(), centers_lowd.data(), k, k);
    double t_ll = now_s() - t0;
    {
        std::vector<uint32_t> assign((size_t)DB, 0xffffffffu);
        for (doc_id_t c = 0; c < k; ++c)
            for (auto d : closest[c]) assign[d] = (uint32_t)c;
        dump(out, "lloyd_assign", assign.data(), assign.size());
        dump(out, "centers_lowd_final", centers_lowd.data(), centers_lowd.size());
        dump(out, "centers", centers.data(), centers.size());
    }
    delete[] closest;
    std::fprintf(meta, ", \"t_lloyd\": %.6f", t_ll);
    B->cleanup_after_eigensolver();                       /* trainer.cpp:554 */
    if (upto == 'E') { std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage F: Lloyd on the full-dimensional B (trainer.cpp:559-571) ---- */
    t0 = now_s();
    auto closest_full = new std::vector<doc_id_t>[k];
    B->run_lloyds(k, centers.data(), closest_full, MAX_KMEANS_REPS);
    double t_lf = now_s() - t0;
    {
        std::vector<uint32_t> assign((size_t)DB, 0xffffffffu);
        for (doc_id_t c = 0; c < k; ++c)
            for (auto d : closest_full[c]) assign[d] = (uint32_t)c;
        dump(out, "full_assign", assign.data(), assign.size());
        dump(out, "full_centers", centers.data(), centers.size());
    }
    std::fprintf(meta, ", \"t_lloyd_full\": %.6f", t_lf);
    if (upto == 'F') { delete[] closest_full; std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage G: catchword thresholds and catchwords (trainer.cpp:573-639) ---- */
    t0 = now_s();
    for (doc_id_t topic = 0; topic != k; ++topic)                       /* :573-575 */
        for (auto d = closest_full[topic].begin(); d < closest_full[topic].end(); ++d)
            *d = original_cols[*d];
    const MKL_UINT r = (MKL_UINT)std::floor(eps2_c * w0_c * (FPTYPE)D / (FPTYPE)(2.0 * k));   /* :583 */
    std::vector<A_TYPE> cthr((size_t)k * (size_t)V);
    for (doc_id_t topic = 0; topic < k; ++topic)                         /* :587-589 (pfor in the reference) */
        A_sp->rth_highest_element(r, closest_full[topic], cthr.data() + (size_t)topic * (size_t)V);
    auto catchwords = new std::vector<word_id_t>[k];
    A_sp->find_catchwords(k, cthr.data(), catchwords);                   /* :635 */
    double t_cw = now_s() - t0;
    {
        std::vector<uint32_t> cl((size_t)D, 0xffffffffu);               /* cluster of every ORIGINAL document */
        for (doc_id_t c = 0; c < k; ++c)
            for (auto d : closest_full[c]) cl[d] = (uint32_t)c;
        dump(out, "catch_cluster_of_doc", cl.data(), cl.size());
        dump(out, "catch_thresholds", cthr.data(), cthr.size());
        std::vector<uint32_t> pairs;
        for (doc_id_t c = 0; c < k; ++c)
            for (auto w : catchwords[c]) { pairs.push_back((uint32_t)c); pairs.push_back((uint32_t)w); }
        dump(out, "catchwords", pairs.data(), pairs.size());
    }
    std::fprintf(meta, ", \"catch_r\": %llu, \"t_catchwords\": %.6f", (unsigned long long)r, t_cw);
    if (upto == 'G') { delete[] catchwords; delete[] closest_full; std::fprintf(meta, "}\n"); std::fclose(meta); return 0; }

    /* ---- stage H: topic model (trainer.cpp:645-651) ---- */
    t0 = now_s();
    DenseMatrix<FPTYPE> Model((word_id_t)V, (doc_id_t)k);
    std::vector<std::tuple<int, int, doc_id_t> > top_topic_pairs;
    std::vector<std::pair<word_id_t, int> > catchword_topics;
    std::vector<std::tuple<doc_id_t, doc_id_t, FPTYPE> > doc_topic_sum;
    A_sp->construct_topic_model(Model, k, closest_full, catchwords, AVG_CLUSTER_FOR_CATCHLESS_TOPIC,
                                &top_topic_pairs, &catchword_topics, &doc_topic_sum);
    double t_tm = now_s() - t0;
    {
        dump(out, "model", Model.data(), (size_t)V * (size_t)k);
        std::vector<uint32_t> dts_doc, dts_topic, ttp;
        std::vector<float> dts_val;
        for (auto &e : doc_topic_sum) {
            dts_doc.push_back((uint32_t)std::get<0>(e)); dts_topic.push_back((uint32_t)std::get<1>(e)); dts_val.push_back(std::get<2>(e));
        }
        for (auto &e : top_topic_pairs) {
            ttp.push_back((uint32_t)std::get<0>(e)); ttp.push_back((uint32_t)std::get<1>(e)); ttp.push_back((uint32_t)std::get<2>(e));
        }
        dump(out, "dts_doc", dts_doc.data(), dts_doc.size());
        dump(out, "dts_topic", dts_topic.data(), dts_topic.size());
        dump(out, "dts_val", dts_val.data(), dts_val.size());
        dump(out, "top_topic_pairs", ttp.data(), ttp.size());
    }
    delete[] catchwords;
    delete[] closest_full;
    std::fprintf(meta, ", \"t_topic_model\": %.6f}\n", t_tm);
    std::fclose(meta);
    delete B;
    delete A_sp;
    return 0;
}
