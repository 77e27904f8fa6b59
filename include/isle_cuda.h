/*
 * isle_cuda.h -- C ABI of libisle_cuda.so: the B200 (sm_100a) spectral core of ISLETrain.
 *
 * The reference (microsoft/ISLE) has no plugin/FFI layer; its spectral core is a set of
 * member functions of ISLE::SparseMatrix<float> / ISLE::FPSparseMatrix<float>
 * (reference include/sparseMatrix.h) called in a fixed order by ISLETrainer::train()
 * (reference src/trainer.cpp:430-554).  Each entry point below replaces one of those
 * members; the replacement translation unit that forwards the C++ members to this ABI is
 * isle_b200/shim/sparseMatrix_cuda.cpp.  Plain pointers and sizes
 * only; every host array is borrowed for the duration of the call; device memory is owned
 * by the context.  All functions return 0 on success and a non-zero code otherwise, with a
 * message available from isle_cuda_last_error().  One caller thread per context
 * (train() is single-threaded).
 *
 * Index widths follow the reference build (-DMKL_ILP64: word_id_t/doc_id_t = uint64_t,
 * offset_t = int64_t, reference include/types.h:24-26); they are narrowed to u32 on upload.
 */
#ifndef ISLE_CUDA_H
#define ISLE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct isle_cuda_ctx isle_cuda_ctx;

enum {
    ISLE_OK = 0,
    ISLE_ERR_CUDA = 1,      /* CUDA / cuBLAS / cuSOLVER / NCCL runtime failure            */
    ISLE_ERR_ARG = 2,       /* invalid argument or call order                               */
    ISLE_ERR_RANGE = 3,     /* input violates a reference invariant (e.g. value > avg size) */
    ISLE_ERR_NOCONV = 4,    /* eigensolver did not converge all k pairs                     */
    ISLE_ERR_NOGPU = 5      /* no usable sm_100 device: there is NO CPU fallback            */
};

/* ---- lifetime -------------------------------------------------------------------- */
/* Single-GPU context on CUDA device `device`. */
int isle_cuda_create(isle_cuda_ctx **ctx, int device);
/* Document-sharded context: rank `rank` of `world` ranks (one process per GPU).  `nccl_id`
 * is the 128-byte ncclUniqueId produced by isle_cuda_nccl_unique_id() on rank 0 and
 * broadcast by the host (torch.distributed / MPI / files).  Each rank later uploads its
 * contiguous slice of documents; V, k and all outputs are global. */
int isle_cuda_create_sharded(isle_cuda_ctx **ctx, int device, int rank, int world,
                             const void *nccl_id);
int isle_cuda_nccl_unique_id(void *id128);
/* Multi-GPU context in ONE process (isle_cuda_create(ctx**, n_gpus)): what a single-threaded caller such as
 * ISLETrainer::train() binds to use every GPU of the box.  n_gpus devices (`devices` lists them, NULL = 0 .. n_gpus-1),
 * one host thread and one document-sharded per-GPU context per device inside the library, communicators from
 * ncclCommInitAll.  Every entry point below accepts such a context with the SAME arguments as a single-GPU one: host
 * arrays describe the whole corpus; the library cuts the documents into n contiguous ranges, runs on all GPUs at once
 * and returns per-document outputs (B, original_cols, assignments, (doc, topic) sums) stitched together in document
 * order, global outputs (zetas, eigenpairs, centers, thresholds, the model) from GPU 0.  Not available on it:
 * isle_cuda_ingest_text, isle_cuda_upload_counts, isle_cuda_download_A. */
int isle_cuda_create_multi(isle_cuda_ctx **ctx, int n_gpus, const int *devices);
void isle_cuda_destroy(isle_cuda_ctx *ctx);
const char *isle_cuda_last_error(const isle_cuda_ctx *ctx);

/* ---- stage A/B: thresholds and the thresholded matrix B -------------------------- */
/* Upload the normalised doc-major CSC of A (reference SparseMatrix<float>:
 * normalized_vals_CSC / rows_CSC / offsets_CSC, include/sparseMatrix.h:23-38) together with
 * avg_doc_sz and the number of non-empty docs (src/sparseMatrix.cpp:97-98).
 * In a sharded context D/nnz/arrays describe the local slice, nz_docs the local count. */
int isle_cuda_upload_A(isle_cuda_ctx *ctx, uint64_t V, uint64_t D, int64_t nnz,
                       const float *normalized_vals, const uint64_t *rows,
                       const int64_t *offsets, float avg_doc_sz, uint64_t nz_docs);
/* Same with 32-bit row ids (harness convenience; no reference counterpart). */
int isle_cuda_upload_A_u32(isle_cuda_ctx *ctx, uint64_t V, uint64_t D, int64_t nnz,
                           const float *normalized_vals, const uint32_t *rows,
                           const int64_t *offsets, float avg_doc_sz, uint64_t nz_docs);
/* ingest on the device.  Replaces DocWordEntriesReader::fill_doc_word_entries (include/utils.h:160-228:
 * `<doc> <word> <count>` lines, 1-based ids, blanks / tabs between the fields, optional '\r', a last line without '\n'),
 * the sort + de-duplication of finalize_data (src/trainer.cpp:232-246; of duplicated (doc, word) lines the first in file
 * order survives), SparseMatrix::populate_CSC (src/sparseMatrix.cpp:58-106; avg_doc_sz = (float)(tokens / non-empty docs),
 * integer division) and SparseMatrix::normalize_docs (src/sparseMatrix.cpp:136-167), bit-exactly.  `text` is the file's
 * bytes (mmap it as the reference does), max_entries its line count (<= 0: not checked; the reference asserts equality,
 * utils.h:227).  The result is the uploaded A of the context (isle_cuda_thresholds may follow directly); *nnz_out = entries
 * after de-duplication, *avg_doc_sz_out / *nz_docs_out / *tokens_out the statistics populate_CSC prints.  A malformed line
 * or an id out of range -> ISLE_ERR_RANGE; so is a document of 2^24 tokens or more (its fp32 token sum, exact and order
 * independent below that, is what makes the normalisation reproducible).  In a sharded context `text` holds the rank's
 * documents (D = the local count, doc ids local) and the statistics are those of the whole corpus. */
int isle_cuda_ingest_text(isle_cuda_ctx *ctx, const char *text, uint64_t size, uint64_t V, uint64_t D, int64_t max_entries,
                          int64_t *nnz_out, float *avg_doc_sz_out, uint64_t *nz_docs_out, uint64_t *tokens_out);
/* populate_CSC's statistics + normalize_docs (src/sparseMatrix.cpp:86-98, 136-167) for a doc-major CSC of RAW counts that
 * is already sorted and de-duplicated: uploads it, normalises on the device (same exactness check), leaves A ready for
 * isle_cuda_thresholds. */
int isle_cuda_upload_counts(isle_cuda_ctx *ctx, uint64_t V, uint64_t D, int64_t nnz, const uint32_t *counts, const uint32_t *rows,
                            const int64_t *offsets, float *avg_doc_sz_out, uint64_t *nz_docs_out);
/* Copies the context's A out in the reference's host layout (normalized_vals_CSC f32[nnz], rows_CSC u64[nnz],
 * offsets_CSC i64[D+1], include/sparseMatrix.h:23-38) for the members that stay on the host.  Any pointer may be NULL. */
int isle_cuda_download_A(isle_cuda_ctx *ctx, float *normalized_vals, uint64_t *rows, int64_t *offsets);
/* Replaces list_word_freqs_by_sorting + compute_thresholds
 * (src/sparseMatrix.cpp:289-333, 357-485).  zetas_out: V floats; *new_nnz_out = number of
 * entries with round(value) >= zeta (the function's return value).  In a sharded context both are
 * statistics of the WHOLE corpus (the per-word histograms are all-reduced first): *new_nnz_out is the
 * global count, identical on every rank; the local nnz of B comes from isle_cuda_build_B. */
int isle_cuda_thresholds(isle_cuda_ctx *ctx, uint64_t k, float *zetas_out,
                         int64_t *new_nnz_out);
/* Replaces threshold_and_copy (src/sparseMatrix.cpp:1285-1361); with a non-NULL
 * select_docs[D] mask it is the tail of sampled_threshold_and_copy (:1417-1430) for that
 * selection.  B stays on the device. */
int isle_cuda_build_B(isle_cuda_ctx *ctx, const uint8_t *select_docs_or_null,
                      int64_t *nnz_B_out, uint64_t *D_B_out);
/* Per-document importance-sampling weights, sum of zeta_w over kept entries
 * (src/sparseMatrix.cpp:1383-1397); weights_out: D floats. */
int isle_cuda_sampling_weights(isle_cuda_ctx *ctx, float *weights_out);
/* SURVEY 8(f) row 4.  The selection of sampled_threshold_and_copy (src/sparseMatrix.cpp:1399-1415) on the device:
 * key_d = u_d^(1 / weight_d) (0 when the weight is 0) with u_d a counter-based uniform of (seed, d) in place of the
 * reference's racy libc rand(); select_out[d] = 1 for the documents whose key is at least the (floor(rate D) + 1)-th
 * largest (all of them when floor(rate D) >= D); *n_selected_out (may be NULL) = how many.  Pass select_out to
 * isle_cuda_build_B.  In a sharded context D, rate D and the pivot are those of the WHOLE corpus (the uniforms are keyed by the
 * global document number and the pivot comes from an exact distributed radix select), so the selection is bit-identical to
 * the single-GPU one; select_out / *n_selected_out describe the rank's own documents.  A multi-GPU context returns the whole
 * mask in document order. */
int isle_cuda_sample_docs(isle_cuda_ctx *ctx, float sample_rate, uint64_t seed, uint8_t *select_out,
                          uint64_t *n_selected_out);
/* Copies B back in the reference's layout for the host stages that follow the spectral
 * core (vals f32[nnz_B], rows u64[nnz_B], offsets i64[D_B+1], original_cols u64[D_B]).
 * Any pointer may be NULL to skip that array. */
int isle_cuda_download_B(isle_cuda_ctx *ctx, float *vals, uint64_t *rows, int64_t *offsets,
                         uint64_t *original_cols);
/* The same copy in the background: _begin stages the arrays on the device and returns; a host thread inside the library
 * moves them to the caller's arrays over a separate copy stream while later calls (isle_cuda_block_ks ...) run; _end waits
 * for it.  The arrays must stay valid and unread until _end returns.  threshold_and_copy's host arrays are read by nothing
 * on the spectral core (train() goes on with the device's B), so the replacement translation unit starts the download in
 * threshold_and_copy and ends it after compute_block_ks.  Calls that rebuild A or B end a pending download first. */
int isle_cuda_download_B_begin(isle_cuda_ctx *ctx, float *vals, uint64_t *rows, int64_t *offsets, uint64_t *original_cols);
int isle_cuda_download_B_end(isle_cuda_ctx *ctx);
/* FPSparseMatrix::frobenius (src/sparseMatrix.cpp:1096-1100): sum of squares of B. */
int isle_cuda_frobenius(isle_cuda_ctx *ctx, float *out);

/* ---- stage C: top-k eigenpairs of B B^T ------------------------------------------ */
/* MKL_SpSpTrProd::multiply (include/matUtils.h:336-365): Z = B (B^T X) for a V x b
 * column-major block (b <= 16).  Exposed so the operator can be checked on its own. */
int isle_cuda_spsptr_multiply(isle_cuda_ctx *ctx, int b, const float *X_colmajor,
                              float *Z_colmajor);
/* Replaces initialize_for_eigensolver + compute_block_ks (src/sparseMatrix.cpp:1150-1158,
 * 1195-1220): restarted block Krylov-Schur, nev=k, ncv=2k+b (block-ks/restarted_block_ks.h).
 * evalues_out: k floats (sigma^2, descending); U_colmajor_out: V*k floats or NULL;
 * *nconv_out = converged pairs.  When max_restarts is exhausted the reference's own rule
 * (restarted_block_ks.h:302-315, see blockks.cu) sets nconv = k and training carries on with the Ritz
 * pairs of the last truncation; so does this call, and the counter "ks_unconverged" (isle_cuda_get_stat)
 * holds the number of pairs whose residual is still above tol.  ISLE_ERR_NOCONV is returned when
 * nconv != k (the reference asserts, src/sparseMatrix.cpp:1207); outputs are still filled. */
int isle_cuda_block_ks(isle_cuda_ctx *ctx, uint64_t k, int b, int max_restarts, float tol,
                       uint64_t seed, float *evalues_out, float *U_colmajor_out,
                       int *nconv_out);
/* Harness hook: install an externally computed U (V x k column-major) so that the k-means
 * stages can be checked on an identical projection. */
int isle_cuda_set_U(isle_cuda_ctx *ctx, uint64_t k, const float *U_colmajor);

/* ---- stage D/E: k-means on the rank-k projection ---------------------------------- */
/* Materialises P = B^T U (D_B x k) on the device; P_out (row-major, D_B*k floats) and
 * l2sq_out (D_B floats, compute_projected_docs_l2sq, src/sparseMatrix.cpp:1888-1918) may
 * be NULL.  Called implicitly by the k-means entry points when needed. */
int isle_cuda_project(isle_cuda_ctx *ctx, float *P_out, float *l2sq_out);
/* Replaces kmeans_init_on_projected_space / kmeanspp_on_projected_space
 * (src/sparseMatrix.cpp:2133-2238).  seeds_out: k doc ids (B numbering);
 * centers_lowd_out: k*k floats, center c at [c*k, (c+1)*k). */
int isle_cuda_kmeanspp(isle_cuda_ctx *ctx, uint64_t k, uint64_t seed, uint64_t *seeds_out,
                       float *centers_lowd_out, float *residual_out);
/* Replaces run_lloyds_on_projected_space (src/sparseMatrix.cpp:1921-2072).
 * assign_out (D_B u32, may be NULL) is the final partition; *objective_out (may be NULL)
 * = sum_d ||P_d - c_assign(d)||^2 in fp64 for the returned centers/partition;
 * *iters_out (may be NULL) = Lloyd iterations executed. */
int isle_cuda_lloyd_projected(isle_cuda_ctx *ctx, uint64_t k, float *centers_lowd_inout,
                              int max_reps, uint32_t *assign_out, double *objective_out,
                              int *iters_out);
/* One assignment pass only (projected_closest_centers, src/sparseMatrix.cpp:1852-1871):
 * argmin_c | ||d||^2 + ||c||^2 - 2 d.c |, first index on ties (cblas_isamin). */
int isle_cuda_assign_projected(isle_cuda_ctx *ctx, uint64_t k, const float *centers_lowd,
                               uint32_t *assign_out);
/* Replaces update_min_distsq_to_projected_centers (src/sparseMatrix.cpp:2075-2130) for all documents of B: the
 * k-means++ refresh for a batch of num_centers new centers (projected_centers: num_centers x k floats, center c at
 * [c*k, (c+1)*k)):  min_dist[d] = min(min_dist[d], max(||P_d||^2 + ||c||^2 - 2 P_d.c, 0)).  min_dist_inout: D_B
 * floats.  isle_cuda_kmeanspp runs this internally on device-resident state; the entry point exists so that every
 * refresh kernel (skinny pass, tcgen05 clamped-min mode, fp32 FMA tiles) can be checked against the oracle. */
int isle_cuda_update_min_dist(isle_cuda_ctx *ctx, uint64_t num_centers, const float *projected_centers,
                              float *min_dist_inout);
/* Replaces left_multiply_by_U_Spectra (src/sparseMatrix.cpp:1438-1450):
 * centers_out (V x ncols, column-major) = U * in (k x ncols column-major, leading dim ld_in).
 * centers_out may be NULL (product computed, result left on the device: device-only timing). */
int isle_cuda_lift_centers(isle_cuda_ctx *ctx, uint64_t ncols, const float *in, uint64_t ld_in,
                           float *centers_out);
/* SURVEY 8(f) row 1.  Replaces FPSparseMatrix::run_lloyds (src/sparseMatrix.cpp:1679-1746, with
 * lloyds_iter :1584-1667, distsq_docs_to_centers :1494-1552, closest_centers :1554-1569,
 * compute_docs_l2sq :1670-1677): Lloyd's k-means on the full-dimensional B, the call that follows
 * the spectral core in train() (src/trainer.cpp:566) and produces closest_docs.
 * centers_inout: k centers of length V, center c contiguous (`centers + c * vocab_size`), updated
 * in place; NULL = start from the centers the last isle_cuda_lift_centers left on the device and
 * keep the result there (device-only timing).  Needs B only (may be called after
 * isle_cuda_cleanup_eigensolver, as train() does).  Stops when the partition repeats or after
 * max_reps iterations.  assign_out (D_B u32, may be NULL) = final partition (argmin |dist|, first
 * index on ties; an empty cluster's center stays zero); *objective_out (may be NULL) =
 * sum_d ||B_d - c_assign(d)||^2 in fp64 for the returned centers/partition; *iters_out (may be
 * NULL) = iterations executed.  max_reps >= 1. */
int isle_cuda_lloyd_full(isle_cuda_ctx *ctx, uint64_t k, float *centers_inout, int max_reps,
                         uint32_t *assign_out, double *objective_out, int *iters_out);
/* SURVEY 8(f) row 2, first half.  Replaces SparseMatrix::rth_highest_element (src/sparseMatrix.cpp:491-524),
 * which train() calls once per topic (src/trainer.cpp:587-589): thresholds_out[w] (V floats) = the r-th highest
 * normalised value of word w among the listed documents of A (original ids) when w occurs in MORE than r of
 * them; otherwise 0, except when r >= ndocs and w occurs in every listed document: then the smallest value.
 * Bit-identical to the reference (a selection of the uploaded floats).  r >= 1.  Needs isle_cuda_upload_A only. */
int isle_cuda_rth_highest_element(isle_cuda_ctx *ctx, uint64_t r, const uint64_t *docs, uint64_t ndocs,
                                  float *thresholds_out);
/* The same for all k clusters in one pass: cluster_of_doc[d] in [0, k) or 0xFFFFFFFF (document in no cluster) for
 * the D documents of A; thresholds_out (k x V, topic-major: thresholds[w + t V], the layout of train()'s
 * catchword_thresholds) may be NULL (result kept on the device for isle_cuda_find_catchwords). */
int isle_cuda_catchword_thresholds(isle_cuda_ctx *ctx, uint64_t k, uint64_t r, const uint32_t *cluster_of_doc,
                                   float *thresholds_out);
/* Replaces SparseMatrix::find_catchwords (src/sparseMatrix.cpp:573-594): topic_of_word_out[w] = the topic t with
 * thresholds[w + t V] > rho * thresholds[w + o V] for every other topic o (compared in double, as the reference's
 * expression evaluates), or -1.  At most one topic can qualify when rho >= 1 (rho_c = 1.1, include/hyperparams.h:11).
 * thresholds: k x V host matrix, or NULL = the device matrix of the last isle_cuda_catchword_thresholds. */
int isle_cuda_find_catchwords(isle_cuda_ctx *ctx, uint64_t k, const float *thresholds, double rho,
                              int32_t *topic_of_word_out);
/* SURVEY 8(f) row 2, second half.  Replaces SparseMatrix::construct_topic_model (src/sparseMatrix.cpp:597-838,
 * called at src/trainer.cpp:645-651).  topic_of_word[w] = the topic w is a catchword of, or -1 (what
 * isle_cuda_find_catchwords returns); cluster_of_doc[d] = cluster of original document d or 0xFFFFFFFF (closest_docs);
 * rank_threshold = (uint)(eps3 w0 (float)D / ((float)k 2.0)) as the reference computes it (:716).
 * model_out (V x k column-major, may be NULL): column t = l1-normalised sum of the documents whose catchword mass for t
 * exceeds the topic's threshold plus the documents of cluster t; a topic nobody contributes to divides by zero exactly
 * as the reference's FPscal(1/asum) does.  *num_entries_out = number of non-zero (document, topic) catchword sums,
 * which isle_cuda_doc_topic_sums then copies out in (document, topic) order (the reference's doc_topic_sum list after
 * its final sort, :789-793; sums are bit-identical to the reference's, the model agrees to fp32 rounding). */
int isle_cuda_construct_topic_model(isle_cuda_ctx *ctx, uint64_t k, const int32_t *topic_of_word,
                                    const uint32_t *cluster_of_doc, uint64_t rank_threshold, float *model_out,
                                    uint64_t *num_entries_out);
int isle_cuda_doc_topic_sums(isle_cuda_ctx *ctx, uint32_t *docs, uint32_t *topics, float *sums);
/* Harness only (no reference counterpart): one block Gram-Schmidt pass of BlockKs::expand
 * (block-ks/restarted_block_ks.h:83-84) on caller data, C = W^T F then F -= W C, with a chosen engine
 * (0 = fp32 FMA, 1 = fp32 FMA with vector loads, 2 = tcgen05 split TF32), so the panel engines of the
 * device eigensolver can be checked in isolation.  W: n x rows column-major (ld n); F: n x b
 * column-major (ld n), updated in place; C_out: rows x b column-major (ld rows). */
int isle_cuda_panel_products(isle_cuda_ctx *ctx, int64_t n, int rows, int b, const float *W,
                             float *F_inout, float *C_out, int engine);
/* Harness only: the library's own collectives over peer memory (NVLink loads / stores + flags, coll.cu) run on many message
 * sizes and types and are compared bit for bit with NCCL's.  *mismatches_out = differing elements over all ranks of a
 * multi-GPU context / on this rank of a sharded one (0 expected); *p2p_active_out = 1 when the peer workspaces are mapped
 * (0: single GPU, or NCCL carries every collective and nothing was compared).  Collective: every rank calls it. */
int isle_cuda_selftest_collectives(isle_cuda_ctx *ctx, uint64_t *mismatches_out, int *p2p_active_out);
/* cleanup_after_eigensolver (src/sparseMatrix.cpp:1264-1275): frees U, P and solver state. */
int isle_cuda_cleanup_eigensolver(isle_cuda_ctx *ctx);

/* ---- measurement ------------------------------------------------------------------ */
/* Turn per-kernel CUDA-event timing on/off (off by default; events are recorded on the
 * context's own stream around each launch of the named kernel families). */
int isle_cuda_set_profiling(isle_cuda_ctx *ctx, int enabled);
/* Named counters, e.g. "launches", "spmm_bt_ms", "spmm_bt_calls", "spmm_bt_bytes",
 * "spmm_b_ms", ..., "dist_ms", "dist_flops", "ks_restarts", "ks_ops".  Unknown name ->
 * ISLE_ERR_ARG.  Reading synchronises the stream. */
int isle_cuda_get_stat(isle_cuda_ctx *ctx, const char *name, double *value_out);
int isle_cuda_reset_stats(isle_cuda_ctx *ctx);
/* Device-side stopwatch: records CUDA events on the context's own stream (the stream every
 * kernel of the library is launched on).  stop synchronises and returns the elapsed ms. */
int isle_cuda_timer_start(isle_cuda_ctx *ctx);
int isle_cuda_timer_stop(isle_cuda_ctx *ctx, double *ms_out);
/* Select kernel variants for A/B measurements: name in {"dist_kernel"} etc. */
int isle_cuda_set_option(isle_cuda_ctx *ctx, const char *name, int value);

#ifdef __cplusplus
}
#endif
#endif /* ISLE_CUDA_H */
