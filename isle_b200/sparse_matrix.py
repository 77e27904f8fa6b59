"""Host-side mirror of the reference's matrix interface for the spectral core.

Same method names, argument meaning and error behaviour as ``ISLE::SparseMatrix<float>`` /
``ISLE::FPSparseMatrix<float>`` (reference include/sparseMatrix.h), each method forwarding to
the C ABI of libisle_cuda.so exactly as the C++ replacement TU does
(isle_b200/shim/sparseMatrix_cuda.cpp).  The parity tests therefore read
like ``ISLETrainer::train()`` (reference src/trainer.cpp:430-554):

    A = SparseMatrix(V, D); A.populate_normalized(vals, rows, offsets, avg_doc_sz, nz_docs)
    freqs = A.list_word_freqs_by_sorting()
    zetas, new_nnzs = A.compute_thresholds(0, V, freqs, k)
    B = FPSparseMatrix(A); original_cols = B.threshold_and_copy(A, zetas, new_nnzs)
    B.initialize_for_eigensolver(k); evalues = B.compute_block_ks(k)
    seeds, centers_lowd, res = B.kmeans_init_on_projected_space(k, 1)
    B.run_lloyds_on_projected_space(k, centers_lowd, None, 10)
    centers = B.left_multiply_by_U_Spectra(centers_lowd, k, k)
    B.cleanup_after_eigensolver()

Host arrays are numpy; index widths are the reference's (u64 rows, i64 offsets).  Everything
numeric happens on the GPU inside the library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _capi
from ._capi import Context, IsleCudaError, ptr

# reference include/hyperparams.h
BLOCK_KS_BLOCK_SIZE = 10
BLOCK_KS_MAX_ITERS = 100
BLOCK_KS_TOLERANCE = 1e-4
KMEANS_INIT_REPS = 1
MAX_KMEANS_LOWD_REPS = 10
MAX_KMEANS_REPS = 10           # include/hyperparams.h:68
RHO_C = 1.1                    # include/hyperparams.h:11
W0_C, EPS3_C = 1.0, 5.0        # include/hyperparams.h:8,12


class SparseMatrix:
    """``ISLE::SparseMatrix<float>``: the normalised doc-major CSC of A on the device."""

    def __init__(self, vocab_size: int, num_docs: int, ctx: Optional[Context] = None, device: int = 0):
        self._vocab_size, self._num_docs = int(vocab_size), int(num_docs)
        self.ctx = ctx if ctx is not None else Context(device)
        self._nnzs = 0
        self.avg_doc_sz = 0.0
        self._nz_docs = 0

    def vocab_size(self) -> int:
        return self._vocab_size

    def num_docs(self) -> int:
        return self._num_docs

    def get_nnzs(self) -> int:
        return self._nnzs

    def populate_normalized(self, normalized_vals, rows, offsets, avg_doc_sz: float, nz_docs: int) -> None:
        """State after populate_CSC + normalize_docs (src/sparseMatrix.cpp:58-167): hands the
        normalised CSC to the device (isle_cuda_upload_A)."""
        vals = np.ascontiguousarray(normalized_vals, dtype=np.float32)
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        assert offs.shape[0] == self._num_docs + 1
        self._nnzs = int(offs[-1])
        self.avg_doc_sz, self._nz_docs = float(avg_doc_sz), int(nz_docs)
        rows = np.ascontiguousarray(rows)
        fn = "isle_cuda_upload_A" if rows.dtype == np.uint64 else "isle_cuda_upload_A_u32"
        if rows.dtype not in (np.uint64, np.uint32):
            rows = rows.astype(np.uint32)
        self.ctx.call(fn, self._vocab_size, self._num_docs, self._nnzs, ptr(vals), ptr(rows), ptr(offs),
                      C.c_float(self.avg_doc_sz), self._nz_docs)

    def ingest_text(self, text: bytes, max_entries: int = 0) -> None:
        """ISLETrainer's ingest on the device: DocWordEntriesReader (include/utils.h:160-228), the sort
        and de-duplication of finalize_data (src/trainer.cpp:232-246), populate_CSC and normalize_docs
        (src/sparseMatrix.cpp:58-167).  ``text`` holds `<doc> <word> <count>` lines; afterwards the object is in the
        state populate_normalized leaves it in."""
        buf = np.frombuffer(text, dtype=np.uint8)
        nnz, avg, nz, tok = C.c_int64(), C.c_float(), C.c_uint64(), C.c_uint64()
        self.ctx.call("isle_cuda_ingest_text", ptr(buf) if len(buf) else None, len(buf), self._vocab_size, self._num_docs,
                      int(max_entries), C.byref(nnz), C.byref(avg), C.byref(nz), C.byref(tok))
        self._nnzs, self.avg_doc_sz, self._nz_docs, self.total_tokens = int(nnz.value), float(avg.value), int(nz.value), int(tok.value)

    def populate_CSC_and_normalize(self, counts, rows, offsets) -> None:
        """populate_CSC's statistics + normalize_docs (src/sparseMatrix.cpp:86-98, 136-167) on the device for a sorted,
        de-duplicated doc-major CSC of raw counts."""
        cnt = np.ascontiguousarray(counts, dtype=np.uint32)
        rws = np.ascontiguousarray(rows, dtype=np.uint32)
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        avg, nz = C.c_float(), C.c_uint64()
        self.ctx.call("isle_cuda_upload_counts", self._vocab_size, self._num_docs, int(offs[-1]), ptr(cnt), ptr(rws), ptr(offs),
                      C.byref(avg), C.byref(nz))
        self._nnzs, self.avg_doc_sz, self._nz_docs = int(offs[-1]), float(avg.value), int(nz.value)

    def download(self):
        """The normalised CSC in the reference's host layout: (normalized_vals f32, rows u64, offsets i64)."""
        vals = np.zeros(self._nnzs, dtype=np.float32)
        rows = np.zeros(self._nnzs, dtype=np.uint64)
        offs = np.zeros(self._num_docs + 1, dtype=np.int64)
        self.ctx.call("isle_cuda_download_A", ptr(vals), ptr(rows), ptr(offs))
        return vals, rows, offs

    def list_word_freqs_by_sorting(self):
        """src/sparseMatrix.cpp:289-333.  ``freqs`` is only a hand-off to compute_thresholds;
        the device path needs no word-major lists, so this is a no-op returning a token."""
        return None

    def compute_thresholds(self, word_begin: int, word_end: int, freqs, num_topics: int) -> Tuple[np.ndarray, int]:
        """src/sparseMatrix.cpp:357-485.  Returns (zetas float32[V], #entries >= threshold)."""
        if word_begin != 0 or word_end != self._vocab_size:
            raise ValueError("device path computes thresholds for the whole vocabulary at once")
        zetas = np.zeros(self._vocab_size, dtype=np.float32)
        nn = C.c_int64()
        self.ctx.call("isle_cuda_thresholds", int(num_topics), ptr(zetas), C.byref(nn))
        return zetas, int(nn.value)


    # -- catchword thresholds and catchwords (src/trainer.cpp:577-639)
    def rth_highest_element(self, r: int, doc_partition) -> np.ndarray:
        """src/sparseMatrix.cpp:491-524 for one cluster: returns thresholds float32[V]."""
        docs = np.ascontiguousarray(doc_partition, dtype=np.uint64)
        thr = np.zeros(self._vocab_size, dtype=np.float32)
        self.ctx.call("isle_cuda_rth_highest_element", int(r), ptr(docs) if len(docs) else None, len(docs), ptr(thr))
        return thr

    def catchword_thresholds(self, num_topics: int, r: int, cluster_of_doc: np.ndarray, download: bool = True):
        """All clusters in one device pass: ``cluster_of_doc`` is uint32[D] (0xFFFFFFFF = in no cluster).
        Returns the (k, V) threshold matrix (topic-major, train()'s catchword_thresholds layout)."""
        cl = np.ascontiguousarray(cluster_of_doc, dtype=np.uint32)
        assert cl.shape[0] == self._num_docs
        thr = np.zeros((int(num_topics), self._vocab_size), dtype=np.float32) if download else None
        self.ctx.call("isle_cuda_catchword_thresholds", int(num_topics), int(r), ptr(cl), ptr(thr))
        return thr

    def find_catchwords(self, num_topics: int, thresholds: Optional[np.ndarray], rho: float = RHO_C) -> List[np.ndarray]:
        """src/sparseMatrix.cpp:573-594.  ``thresholds`` (k, V) or None (the device copy of the last
        catchword_thresholds).  Returns k ascending word-id arrays."""
        thr = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float32)
        tw = np.zeros(self._vocab_size, dtype=np.int32)
        self.ctx.call("isle_cuda_find_catchwords", int(num_topics), ptr(thr), C.c_double(rho), ptr(tw))
        return [np.nonzero(tw == t)[0] for t in range(int(num_topics))]


    def construct_topic_model(self, num_topics: int, cluster_of_doc: np.ndarray, catchwords, want_pairs: bool = True,
                              total_docs: Optional[int] = None):
        """src/sparseMatrix.cpp:597-838.  ``cluster_of_doc`` uint32[D] (closest_docs as a map, 0xFFFFFFFF = none),
        ``catchwords`` a list of k word-id arrays.  Returns (Model float32[V, k], (doc, topic, sum) arrays in
        (doc, topic) order, top_topic_pairs int64[n, 3] in document order or None)."""
        k = int(num_topics)
        tw = np.full(self._vocab_size, -1, dtype=np.int32)
        for t in range(k):
            tw[np.asarray(catchwords[t], dtype=np.int64)] = t
        cl = np.ascontiguousarray(cluster_of_doc, dtype=np.uint32)
        assert cl.shape[0] == self._num_docs
        # (doc_id_t)(eps3_c * w0_c * (FPTYPE)num_docs() / ((FPTYPE)num_topics * 2.0))   (:716)
        # document-sharded contexts pass the corpus-wide document count (the rank is a global quantity)
        rank = int(np.uint64(EPS3_C * W0_C * float(np.float32(total_docs or self._num_docs)) / (float(np.float32(k)) * 2.0)))
        model = np.zeros((k, self._vocab_size), dtype=np.float32)
        n = C.c_uint64()
        self.ctx.call("isle_cuda_construct_topic_model", k, ptr(tw), ptr(cl), rank, ptr(model), C.byref(n))
        n = int(n.value)
        docs, topics, sums = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self.ctx.call("isle_cuda_doc_topic_sums", ptr(docs), ptr(topics), ptr(sums))
        pairs = None
        if want_pairs:                                       # :683-703, strict > in topic order
            out = []
            bounds = np.searchsorted(docs, np.arange(self._num_docs + 1))
            for d in range(self._num_docs):
                mx = mx2 = np.float32(0.0)
                t1 = t2 = -1
                for i in range(bounds[d], bounds[d + 1]):
                    x = sums[i]
                    if x > mx:
                        mx2, t2, mx, t1 = mx, t1, x, int(topics[i])
                    elif x > mx2:
                        mx2, t2 = x, int(topics[i])
                if t1 >= 0 and t2 >= 0:
                    out.append((t1, t2, d))
            pairs = np.array(out, dtype=np.int64).reshape(-1, 3)
        return model.T.copy(), (docs, topics, sums), pairs


class FPSparseMatrix:
    """``ISLE::FPSparseMatrix<float>``: the thresholded matrix B and everything train() does
    with it between trainer.cpp:475 and :554."""

    def __init__(self, A: SparseMatrix):
        self.ctx = A.ctx
        self._vocab_size = A.vocab_size()
        self._num_docs = A.num_docs()
        self._nnzs = 0
        self.U_cols = 0

    def vocab_size(self) -> int:
        return self._vocab_size

    def num_docs(self) -> int:
        return self._num_docs

    def get_nnzs(self) -> int:
        return self._nnzs

    # -- stage B
    def threshold_and_copy(self, A: SparseMatrix, zetas, nnzs: int) -> np.ndarray:
        """src/sparseMatrix.cpp:1285-1361.  Returns original_cols (uint64[D_B])."""
        return self._build(None)

    def sampled_threshold_and_copy(self, A: SparseMatrix, zetas, nnzs: int, sample_rate: float,
                                   rng: Optional[np.random.Generator] = None,
                                   select_docs: Optional[np.ndarray] = None,
                                   device_seed: Optional[int] = None) -> np.ndarray:
        """src/sparseMatrix.cpp:1365-1435.  weight_d = sum zeta over kept entries (device);
        key_d = u^(1/weight_d); keep the floor(rate*D) largest keys (A-Res sampling).  The
        reference draws u from libc rand() inside a parallel loop (racy); here the
        caller supplies the generator, or the selection mask itself for parity runs."""
        if select_docs is None and device_seed is not None:      # keys, pivot and selection on the device
            select_docs = np.zeros(A.num_docs(), dtype=np.uint8)
            nsel = C.c_uint64()
            self.ctx.call("isle_cuda_sample_docs", C.c_float(sample_rate), int(device_seed), ptr(select_docs), C.byref(nsel))
            self.last_sample_count = int(nsel.value)
        if select_docs is None:
            w = np.zeros(A.num_docs(), dtype=np.float32)
            self.ctx.call("isle_cuda_sampling_weights", ptr(w))
            rng = rng or np.random.default_rng(0)
            u = rng.random(A.num_docs())
            dice = np.where(w == 0, 0.0, np.power(u, 1.0 / np.maximum(w, 1e-30))).astype(np.float32)
            nth = int(np.float32(sample_rate) * np.float32(A.num_docs()))
            pivot = np.partition(dice, len(dice) - 1 - nth)[len(dice) - 1 - nth] if nth < len(dice) else 0.0
            select_docs = dice >= pivot
        return self._build(np.ascontiguousarray(select_docs, dtype=np.uint8))

    def _build(self, mask) -> np.ndarray:
        nnzB, DB = C.c_int64(), C.c_uint64()
        self.ctx.call("isle_cuda_build_B", ptr(mask), C.byref(nnzB), C.byref(DB))
        self._nnzs, self._num_docs = int(nnzB.value), int(DB.value)
        oc = np.zeros(self._num_docs, dtype=np.uint64)
        self.ctx.call("isle_cuda_download_B", None, None, None, ptr(oc))
        return oc

    def download(self):
        """CSC arrays in the reference's layout (vals f32, rows u64, offsets i64)."""
        vals = np.zeros(self._nnzs, dtype=np.float32)
        rows = np.zeros(self._nnzs, dtype=np.uint64)
        offs = np.zeros(self._num_docs + 1, dtype=np.int64)
        oc = np.zeros(self._num_docs, dtype=np.uint64)
        self.ctx.call("isle_cuda_download_B", ptr(vals), ptr(rows), ptr(offs), ptr(oc))
        return vals, rows, offs, oc

    def frobenius(self) -> float:
        """src/sparseMatrix.cpp:1096-1100."""
        v = C.c_float()
        self.ctx.call("isle_cuda_frobenius", C.byref(v))
        return float(v.value)

    # -- stage C
    def initialize_for_eigensolver(self, num_topics: int) -> None:
        """src/sparseMatrix.cpp:1150-1158 (U is allocated by the library)."""
        self.U_cols = int(num_topics)

    def multiply(self, X: np.ndarray) -> np.ndarray:
        """MKL_SpSpTrProd::multiply (include/matUtils.h:336-365): B (B^T X), X is V x b."""
        X = np.asarray(X, dtype=np.float32)
        b = X.shape[1]
        Xc = np.ascontiguousarray(X.T)          # column-major V x b == C-order b x V
        Zc = np.zeros_like(Xc)
        self.ctx.call("isle_cuda_spsptr_multiply", int(b), ptr(Xc), ptr(Zc))
        return Zc.T.copy()

    def compute_block_ks(self, num_topics: int, *, block_size: int = BLOCK_KS_BLOCK_SIZE,
                         max_iters: int = BLOCK_KS_MAX_ITERS, tol: float = BLOCK_KS_TOLERANCE,
                         seed: int = 0, want_U: bool = False):
        """src/sparseMatrix.cpp:1195-1220.  Returns evalues (float32[k], sigma^2 descending), or
        (evalues, U[V,k]) with want_U.  Raises when nconv != k (the reference asserts, :1207)."""
        k = int(num_topics)
        ev = np.zeros(k, dtype=np.float32)
        U = np.zeros((k, self._vocab_size), dtype=np.float32) if want_U else None
        nconv = C.c_int()
        self.ctx.call("isle_cuda_block_ks", k, int(block_size), int(max_iters), C.c_float(tol), int(seed),
                      ptr(ev), ptr(U), C.byref(nconv))
        self.U_cols = k
        self.nconv = int(nconv.value)
        return (ev, U.T.copy()) if want_U else ev

    def set_U(self, U: np.ndarray) -> None:
        """Harness hook: install U (V x k) computed elsewhere."""
        U = np.asarray(U, dtype=np.float32)
        self.U_cols = U.shape[1]
        Uc = np.ascontiguousarray(U.T)
        self.ctx.call("isle_cuda_set_U", int(self.U_cols), ptr(Uc))

    def cleanup_after_eigensolver(self) -> None:
        """src/sparseMatrix.cpp:1264-1275."""
        self.ctx.call("isle_cuda_cleanup_eigensolver")

    # -- stages D/E
    def projected_docs(self) -> Tuple[np.ndarray, np.ndarray]:
        """P = B^T U (D_B x k) and ||P_d||^2 (UT_times_docs / compute_projected_docs_l2sq)."""
        P = np.zeros((self._num_docs, self.U_cols), dtype=np.float32)
        l2 = np.zeros(self._num_docs, dtype=np.float32)
        self.ctx.call("isle_cuda_project", ptr(P), ptr(l2))
        return P, l2

    def kmeans_init_on_projected_space(self, num_centers: int, max_reps: int = KMEANS_INIT_REPS, *, seed: int = 0):
        """src/sparseMatrix.cpp:2212-2238.  Returns (best_seed u64[k], best_centers_coords f32[k,k],
        min_total_dist)."""
        k = int(num_centers)
        best = None
        for rep in range(max_reps):
            seeds = np.zeros(k, dtype=np.uint64)
            coords = np.zeros((k, k), dtype=np.float32)
            res = C.c_float()
            self.ctx.call("isle_cuda_kmeanspp", k, int(seed) + rep, ptr(seeds), ptr(coords), C.byref(res))
            if best is None or res.value < best[2]:
                best = (seeds, coords, float(res.value))
        return best

    def run_lloyds_on_projected_space(self, num_centers: int, projected_centers: np.ndarray,
                                      closest_docs: Optional[List[list]] = None,
                                      max_reps: int = MAX_KMEANS_LOWD_REPS):
        """src/sparseMatrix.cpp:2016-2072.  ``projected_centers`` (k x k, center c = row c) is
        updated in place; ``closest_docs`` (list of k lists) receives the partition when given.
        Returns the residual the reference returns (always 0); the assignment,
        objective and iteration count are kept in ``self.last_lloyd``."""
        k = int(num_centers)
        assert projected_centers.dtype == np.float32 and projected_centers.flags["C_CONTIGUOUS"]
        assign = np.zeros(self._num_docs, dtype=np.uint32)
        obj, iters = C.c_double(), C.c_int()
        self.ctx.call("isle_cuda_lloyd_projected", k, ptr(projected_centers), int(max_reps), ptr(assign),
                      C.byref(obj), C.byref(iters))
        self.last_lloyd = dict(assign=assign, objective=float(obj.value), iters=int(iters.value))
        if closest_docs is not None:
            order = np.argsort(assign, kind="stable")
            bounds = np.searchsorted(assign[order], np.arange(k + 1))
            for c in range(k):
                closest_docs[c][:] = order[bounds[c]:bounds[c + 1]].tolist()
        return 0.0

    def projected_closest_centers(self, num_centers: int, projected_centers: np.ndarray) -> np.ndarray:
        """src/sparseMatrix.cpp:1852-1871 for all docs: argmin |dist|, first index on ties."""
        assign = np.zeros(self._num_docs, dtype=np.uint32)
        pc = np.ascontiguousarray(projected_centers, dtype=np.float32)
        self.ctx.call("isle_cuda_assign_projected", int(num_centers), ptr(pc), ptr(assign))
        return assign

    def update_min_distsq_to_projected_centers(self, projected_centers: np.ndarray, min_dist: np.ndarray) -> np.ndarray:
        """src/sparseMatrix.cpp:2075-2130 over all documents: ``projected_centers`` is (num_centers, k), ``min_dist``
        float32[D_B] is updated in place with min(min_dist, max(dist, 0)) and returned."""
        pc = np.ascontiguousarray(projected_centers, dtype=np.float32)
        assert pc.ndim == 2 and pc.shape[1] == self.U_cols
        assert min_dist.dtype == np.float32 and min_dist.flags["C_CONTIGUOUS"] and min_dist.shape[0] == self._num_docs
        self.ctx.call("isle_cuda_update_min_dist", pc.shape[0], ptr(pc), ptr(min_dist))
        return min_dist

    def run_lloyds(self, num_centers: int, centers: np.ndarray, closest_docs: Optional[List[list]] = None,
                   max_reps: int = MAX_KMEANS_REPS) -> float:
        """src/sparseMatrix.cpp:1679-1746: Lloyd's on the full-dimensional B.
        ``centers`` is (k, V) float32 C-contiguous, center c = row c (the reference's
        ``centers + c * vocab_size``), updated in place; pass None to start from the centers the last
        left_multiply_by_U_Spectra left on the device.  ``closest_docs`` (list of k lists) receives the
        partition.  Returns the residual the reference returns (0: compute_residual is off, :1588);
        assignment, objective and iteration count are kept in ``self.last_lloyd_full``."""
        k = int(num_centers)
        if centers is not None:
            assert centers.dtype == np.float32 and centers.flags["C_CONTIGUOUS"]
            assert centers.shape == (k, self._vocab_size)
        assign = np.zeros(self._num_docs, dtype=np.uint32)
        obj, iters = C.c_double(), C.c_int()
        self.ctx.call("isle_cuda_lloyd_full", k, ptr(centers), int(max_reps), ptr(assign), C.byref(obj), C.byref(iters))
        self.last_lloyd_full = dict(assign=assign, objective=float(obj.value), iters=int(iters.value))
        if closest_docs is not None:
            order = np.argsort(assign, kind="stable")
            bounds = np.searchsorted(assign[order], np.arange(k + 1))
            for c in range(k):
                closest_docs[c][:] = order[bounds[c]:bounds[c + 1]].tolist()
        return 0.0

    def left_multiply_by_U_Spectra(self, inp: np.ndarray, ld_in: int, ncols: int) -> np.ndarray:
        """src/sparseMatrix.cpp:1438-1450: out (V x ncols) = U * in, ``inp`` holds ncols columns of
        length ld_in (column c = inp[c, :]).  Returns out with out[:, c] the lifted column."""
        a = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.zeros((int(ncols), self._vocab_size), dtype=np.float32)
        self.ctx.call("isle_cuda_lift_centers", int(ncols), ptr(a), int(ld_in), ptr(out))
        return out.T.copy()
