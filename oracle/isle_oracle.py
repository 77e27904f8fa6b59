"""oracle/isle_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy / scipy.sparse) of the reference's spectral core and of the stages
widened from it (full-dimensional Lloyd, catchword thresholds, catchwords, topic model), the
checker the CUDA path is compared against.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module; the product path
(isle_b200/, libisle_cuda.so) never does.

Parity pinning: the reference has no golden vectors of its own,
so this restatement is pinned against outputs of the UNMODIFIED reference C++
compiled in-container (oracle/_ref/ref_dump, see oracle/Makefile) on seeded
corpora; the resulting fixtures live in tests/golden/ together with the script
that made them (tests/golden/make_golden.py).  The only substitution below the
reference's C++ is BLAS/LAPACK = OpenBLAS and the MKL sparse calls = plain loops
(oracle/shim/), which does not touch the bit-exact stages (thresholds, B).

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
import scipy.sparse as sp

F32 = np.float32


# --------------------------------------------------------------------------- ingest
def parse_entries(text: bytes):
    """include/utils.h:160-228 DocWordEntriesReader::fill_doc_word_entries, restated character by character (small
    inputs only): digits accumulate into the 1st / 2nd / 3rd field of the line, a digit after blanks or tabs starts the
    next field, '\r' is ignored, '\n' emits the entry with 0-based doc and word, a last line without '\n' is emitted
    if its third field was reached.  Returns (doc, word, count) int64 arrays in file order."""
    docs, words, counts = [], [], []
    doc = word = count = 0
    state, was_ws = 1, False
    for ch in text:
        c = chr(ch)
        if c == "\r":
            continue
        if c == "\n":
            docs.append(doc - 1); words.append(word - 1); counts.append(count)
            doc = word = count = 0
            state = 1
            was_ws = False      # (the reference leaves was_whitespace set across lines; lines never start with blanks here)
        elif c in " \t":
            was_ws = True
        elif c.isdigit():
            if was_ws:
                state += 1
                was_ws = False
            assert state <= 3, "Bad line"
            if state == 1:
                doc = doc * 10 + int(c)
            elif state == 2:
                word = word * 10 + int(c)
            else:
                count = count * 10 + int(c)
        else:
            raise ValueError("Bad format")
    if state == 3:
        docs.append(doc - 1); words.append(word - 1); counts.append(count)
    return np.array(docs, np.int64), np.array(words, np.int64), np.array(counts, np.int64)


def entries_to_csc(docs, words, counts, D: int):
    """src/trainer.cpp:237-246 (sort by (doc, word), std::unique keeps the first of each (doc, word) -- here the first in
    file order) + src/sparseMatrix.cpp:58-84 populate_CSC: (offsets i64[D+1], rows u32, counts u32)."""
    order = np.lexsort((words, docs))            # stable: file order among equals
    d, w, c = docs[order], words[order], counts[order]
    first = np.ones(len(d), dtype=bool)
    first[1:] = (d[1:] != d[:-1]) | (w[1:] != w[:-1])
    d, w, c = d[first], w[first], c[first]
    offsets = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(np.bincount(d, minlength=D), out=offsets[1:])
    return offsets, w.astype(np.uint32), c.astype(np.uint32)


# --------------------------------------------------------------------------- stage 0
def normalize_docs(counts, offsets):
    """src/sparseMatrix.cpp:86-98 (avg_doc_sz, nz_docs) and :136-167 (normalize_docs).

    avg_doc_sz = (float)(total_tokens / nz_docs) with INTEGER division; doc_sum is a
    sequential fp32 sum of integer-valued floats (exact below 2^24); value =
    avg_doc_sz * ((float)count / doc_sum) in fp32 with that association.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(offsets)
    nz_docs = int((lens > 0).sum())
    total = int(np.asarray(counts, dtype=np.uint64).sum())
    avg = F32(total // nz_docs)
    sums = np.zeros(len(lens), dtype=np.int64)
    nzmask = lens > 0
    sums[nzmask] = np.add.reduceat(np.asarray(counts, dtype=np.int64), offsets[:-1][nzmask])
    assert sums.max(initial=0) < (1 << 24)
    ds = np.repeat(sums.astype(F32), lens)
    vals = (avg * (np.asarray(counts).astype(F32) / ds)).astype(F32)
    return vals, avg, nz_docs


# --------------------------------------------------------------------------- stage A
def threshold_counts(nz_docs: int, k: int):
    """src/sparseMatrix.cpp:370-373: count_gr / count_eq, evaluated in double from
    (float)nz_docs and (float)k with w0_c=1.0, eps1_c=1.0/60.0 (include/hyperparams.h:8-9)."""
    w0, eps1 = 1.0, 1.0 / 60.0
    nzf, kf = float(F32(nz_docs)), float(F32(k))
    count_gr = int(w0 * nzf / (2.0 * kf))
    count_eq = int(math.ceil(3.0 * eps1 * w0 * nzf / kf))
    return max(count_gr, 1), max(count_eq, 1)


def round_half_away(x):
    """std::round on float (src/sparseMatrix.cpp:381,1344): half away from zero."""
    x = np.asarray(x, dtype=F32).astype(np.float64)   # |x| + 0.5 is exact in double for fp32 x
    return (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(F32)


def compute_thresholds(vals, rows, V: int, nz_docs: int, k: int):
    """src/sparseMatrix.cpp:289-333 (word-major descending lists) + :357-485 (rank rule).

    Per word w: S_w = rounded values with zeros removed (:378-387).
      n == 0                -> zeta = 1, kept 0            (:476-480)
      n <  count_gr         -> zeta = 1, keep all          (:395-412)
      z = S_w[count_gr-1] (descending); loop (:444-471):
         eq = #(S_w == z); if eq < count_eq: zeta = z, kept = #(S_w >= z)
         elif z is the smallest value present or z == 1: zeta = 1, keep all
         else z = next smaller distinct value
    Returns (zetas float32[V], new_nnzs).
    """
    return thresholds_from_histogram(word_histogram(vals, rows, V), nz_docs, k)


def word_histogram(vals, rows, V: int, bins: int = 0):
    """hist[w, v] = #documents in which word w has rounded value v >= 1 (:378-387).  Additive over
    disjoint document sets: the doc-sharded path sums these tables across ranks."""
    r = round_half_away(vals).astype(np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    keep = r >= 1
    r, w = r[keep], rows[keep]
    maxv = max(int(r.max(initial=1)) + 2, bins)
    hist = np.zeros((V, maxv), dtype=np.int64)
    np.add.at(hist, (w, r), 1)
    return hist


def thresholds_from_histogram(hist, nz_docs: int, k: int):
    """The rank rule of compute_thresholds (:389-481) on per-word histograms."""
    V = hist.shape[0]
    count_gr, count_eq = threshold_counts(nz_docs, k)
    zetas = np.ones(V, dtype=F32)
    new_nnzs = 0
    n_w = hist.sum(1)
    for word in np.nonzero(n_w)[0]:
        h = hist[word]
        n = int(n_w[word])
        if n < count_gr:
            new_nnzs += n
            continue
        vals_present = np.nonzero(h)[0][::-1]  # descending distinct values
        cum = np.cumsum(h[vals_present])       # #(>= value)
        i = int(np.searchsorted(cum, count_gr, side="left"))  # count_gr-th largest
        while True:
            z = int(vals_present[i])
            if h[z] < count_eq:
                zetas[word] = F32(z)
                new_nnzs += int(cum[i])
                break
            if i == len(vals_present) - 1 or z == 1:
                zetas[word] = F32(1.0)
                new_nnzs += n
                break
            i += 1
    return zetas, int(new_nnzs)


# --------------------------------------------------------------------------- stage B
def sampling_weights(vals, rows, offsets, zetas):
    """src/sparseMatrix.cpp:1383-1397: weight_d = sum of zeta_w over kept entries (fp32)."""
    r = round_half_away(vals)
    z = np.asarray(zetas, dtype=F32)[np.asarray(rows, dtype=np.int64)]
    contrib = np.where(r >= z, z, F32(0)).astype(np.float64)
    cs = np.concatenate([[0.0], np.cumsum(contrib)])
    offsets = np.asarray(offsets, dtype=np.int64)
    return (cs[offsets[1:]] - cs[offsets[:-1]]).astype(F32)


def threshold_and_copy(vals, rows, offsets, zetas, select_docs=None):
    """src/sparseMatrix.cpp:1285-1361 (threshold_and_copy / _doc_block); with a mask it is
    the tail of sampled_threshold_and_copy (:1417-1430).

    Entry kept iff round(val) >= zeta[row]; stored value sqrtf(zeta[row]) (:1344-1349);
    docs with no kept entry are dropped and ids compacted, original_cols[new] = old (:1355-1359).
    Returns (B_vals f32, B_rows u32, B_offsets i64, original_cols u32).
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    z = np.asarray(zetas, dtype=F32)[rows]
    keep = round_half_away(vals) >= z
    D = len(offsets) - 1
    doc_of = np.repeat(np.arange(D, dtype=np.int64), np.diff(offsets))
    if select_docs is not None:
        keep &= np.asarray(select_docs, dtype=bool)[doc_of]
    kept_per_doc = np.bincount(doc_of[keep], minlength=D)
    nonempty = kept_per_doc > 0
    original_cols = np.nonzero(nonempty)[0].astype(np.uint32)
    b_offsets = np.zeros(len(original_cols) + 1, dtype=np.int64)
    np.cumsum(kept_per_doc[nonempty], out=b_offsets[1:])
    b_rows = rows[keep].astype(np.uint32)
    b_vals = np.sqrt(z[keep]).astype(F32)
    return b_vals, b_rows, b_offsets, original_cols


def to_csc(b_vals, b_rows, b_offsets, V):
    D = len(b_offsets) - 1
    return sp.csc_matrix((np.asarray(b_vals, dtype=F32), np.asarray(b_rows, dtype=np.int64),
                          np.asarray(b_offsets, dtype=np.int64)), shape=(V, D))


# --------------------------------------------------------------------------- stage C
def spsptr_multiply(B_csc, X):
    """include/matUtils.h:336-365 MKL_SpSpTrProd::multiply: Z = B (B^T X), fp32."""
    Y = (B_csc.T @ X.astype(F32)).astype(F32)
    return (B_csc @ Y).astype(F32)


def compute_qr(A):
    """block-ks/ks_utils.h:43-127: rank-revealing MGS with one re-orthogonalisation, in
    fp64; v_norm is rounded to float (:66 ARMA_FPTYPE); columns with norm < 1e-6 dropped.
    Returns (Q f32 [n x r], R f32 [r x b], rank)."""
    a = np.array(A, dtype=np.float64)
    n, b = a.shape
    Q = np.zeros((n, b))
    R = np.zeros((b, b))
    rank = 0
    for i in range(b):
        v = a[:, i].copy()
        v_norm = float(F32(math.sqrt(float(v @ v))))
        if v_norm < 1e-6:
            continue
        q = v / v_norm
        Q[:, rank] = q
        bb = q @ a[:, i:]
        a[:, i:] -= np.outer(q, bb)
        cc = q @ a[:, i:]
        a[:, i:] -= np.outer(q, cc)
        R[rank, i:] = bb + cc
        rank += 1
    return Q[:, :rank].astype(F32), R[:rank].astype(F32), rank


class BlockKS:
    """block-ks/restarted_block_ks.h restated.  `op(X)` multiplies an
    n x b fp32 block.  RNG is numpy (the reference uses libc rand()): parity is
    by invariant subspace, never by RNG replay."""

    def __init__(self, op, n, nev, ncv=None, maxit=100, blk=10, tol=1e-4, seed=0):
        self.op, self.n, self.nev = op, n, nev
        self.blk = blk if blk < nev else 1          # restarted_block_ks.h:198
        self.ncv = ncv if ncv is not None else 2 * nev + blk
        self.maxit, self.tol = maxit, F32(tol)
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.nconv = 0
        self.n_op = 0
        self.n_restarts = 0

    def _randu(self, n, m):
        return self.rng.random((n, m), dtype=F32)

    def init(self):
        """restarted_block_ks.h:204-259."""
        b = self.blk
        while True:
            Q, R, rank = compute_qr(self._randu(self.n, b))
            if rank == b:
                break
        V = Q
        V1 = self.op(V); self.n_op += 1
        H = (V.T @ V1).astype(F32)
        V1 = V1 - V @ H
        C = (V.T @ V1).astype(F32)
        H = H + C
        V1 = V1 - V @ C
        Q, R, rank = compute_qr(V1)
        R = np.vstack([R, np.zeros((b - rank, b), dtype=F32)])
        self.H = np.vstack([H, R]).astype(F32)
        V = np.hstack([V, Q])
        if rank < b:
            V = self._refill(V, V.shape[1], 2 * b)
        self.V = V.astype(F32)

    def _refill(self, V, nvecs, target):
        """rank-deficiency fix, restarted_block_ks.h:106-131 / :238-258."""
        tries = 0
        if V.shape[1] < target:
            V = np.hstack([V, np.zeros((self.n, target - V.shape[1]), dtype=F32)])
        while nvecs < target and tries < 100:
            tries += 1
            F2 = self._randu(self.n, target - nvecs)
            W = V[:, :nvecs]
            F2 = F2 - W @ (W.T @ F2)
            F2 = F2 - W @ (W.T @ F2)
            Q2, _, rk2 = compute_qr(F2)
            if rk2 > 0:
                V[:, nvecs:nvecs + rk2] = Q2
                nvecs += rk2
        return V

    def expand(self):
        """restarted_block_ks.h:63-136."""
        b, ncv = self.blk, self.ncv
        V, H = self.V, self.H
        V = np.hstack([V, np.zeros((self.n, ncv - H.shape[0]), dtype=F32)])
        while H.shape[0] < ncv:
            rows, cols = H.shape
            Vk = V[:, cols:rows]
            F = self.op(Vk); self.n_op += 1
            W = V[:, :rows]
            Hk = (W.T @ F).astype(F32)
            F = (F - W @ Hk).astype(F32)
            for _ in range(2):
                Ck = (W.T @ F).astype(F32)
                F = (F - W @ Ck).astype(F32)
                Hk = Hk + Ck
            H = np.hstack([H, Hk])
            H = np.vstack([H, np.zeros((b, H.shape[1]), dtype=F32)])
            Q, R, rk = compute_qr(F)
            V[:, H.shape[1]:H.shape[1] + rk] = Q
            R = np.vstack([R, np.zeros((b - rk, b), dtype=F32)])
            H[H.shape[0] - b:, H.shape[1] - b:] = R
            if rk < b:
                V = self._refill(V, H.shape[1] + rk, H.shape[0])
        self.V, self.H = V, H

    def truncate(self):
        """restarted_block_ks.h:139-187.  eig_sym -> ssyevd with uplo='U'
        (armadillo_bits/auxlib_meat.hpp:1670-1682): only the upper triangle of H is read."""
        b, nev, nconv = self.blk, self.nev, self.nconv
        V, H = self.V, self.H
        m = H.shape[1]
        subH = H[nconv:m, nconv:m]
        eH, vH = scipy.linalg.eigh(subH.astype(F32), lower=False)
        idx = np.argsort(-eH, kind="stable")
        eH, vH = eH[idx].astype(F32), vH[:, idx].astype(F32)
        new_starts = V[:, -b:]
        preserve = V[:, :nconv]
        mid = (V[:, nconv:V.shape[1] - b] @ vH[:, :nev - nconv]).astype(F32)
        V = np.hstack([preserve, mid, new_starts])
        H = H.copy()
        H[nconv:nev, nconv:nev] = np.diag(eH[:nev - nconv])
        H[nev:nev + b, nconv:m] = H[H.shape[0] - b:, m - b:m] @ vH[-b:, :]
        if nconv > 0:
            H[:nconv, nconv:m] = H[:nconv, nconv:m] @ vH
        self.V, self.H = V.astype(F32), H[:nev + b, :nev].astype(F32)

    def compute(self):
        """restarted_block_ks.h:262-321."""
        self.nconv = 0
        self.expand()
        while self.n_restarts < self.maxit:
            self.truncate()
            res = self.H[-self.blk:, :]
            norms = np.sqrt((res.astype(F32) ** 2).sum(0)).astype(F32)
            evs = np.diag(self.H)[:norms.size]
            norms = norms / evs
            bad = np.nonzero(norms >= self.tol)[0]
            if bad.size == 0:
                self.nconv = norms.size
                break
            self.nconv = int(bad[0])
            self.n_restarts += 1
            self.expand()
        self.nconv = min(self.nconv, self.nev)
        return self.nconv

    def eigenvalues(self):
        return np.diag(self.H)[: self.nev].astype(F32)

    def eigenvectors(self):
        return self.V[:, : self.nev].astype(F32)


def block_ks(B_csc, k, blk=10, maxit=100, tol=1e-4, seed=0):
    """src/sparseMatrix.cpp:1195-1220 compute_block_ks: evalues (sigma^2, descending), U (V x k)."""
    ks = BlockKS(lambda X: spsptr_multiply(B_csc, X), B_csc.shape[0], k, 2 * k + blk, maxit, blk, tol, seed)
    ks.init()
    nconv = ks.compute()
    return ks.eigenvalues(), ks.eigenvectors(), nconv, ks


# --------------------------------------------------------------------------- stages D/E
def project(B_csc, U):
    """src/sparseMatrix.cpp:1749-1791 multiply_with/UT_times_docs: P = B^T U (docs x k), fp32."""
    return np.asarray(B_csc.T @ U.astype(F32), dtype=F32)


def docs_l2sq(P):
    """src/sparseMatrix.cpp:1888-1918 compute_projected_docs_l2sq."""
    return np.einsum("ij,ij->i", P, P, dtype=F32).astype(F32)


def dist_matrix(P, d2, C):
    """src/sparseMatrix.cpp:1794-1849: ((-2 P C^T) + ||c||^2) + ||d||^2 in fp32, that order."""
    c2 = np.einsum("ij,ij->i", C, C, dtype=F32).astype(F32)
    G = (F32(-2.0) * (P @ C.T.astype(F32))).astype(F32)
    return ((G + c2[None, :]).astype(F32) + d2[:, None]).astype(F32)


def closest_centers(P, d2, C):
    """src/sparseMatrix.cpp:1852-1871: cblas_isamin = first index of min |x|."""
    return np.argmin(np.abs(dist_matrix(P, d2, C)), axis=1).astype(np.uint32)


def lloyds_iter(P, d2, C):
    """src/sparseMatrix.cpp:1921-2013: assign, then centers = mean of members; an empty
    cluster's center is left at zero (:1988-1992).  Returns (new C, assignment)."""
    k = C.shape[0]
    a = closest_centers(P, d2, C)
    newC = np.zeros_like(C, dtype=np.float64)
    np.add.at(newC, a, P.astype(np.float64))
    cnt = np.bincount(a, minlength=k)
    nz = cnt > 0
    newC[nz] /= cnt[nz][:, None]
    return newC.astype(F32), a


def run_lloyds(P, C0, max_reps=10):
    """src/sparseMatrix.cpp:2016-2072: stop when the partition equals the previous one."""
    d2 = docs_l2sq(P)
    C = C0.astype(F32).copy()
    prev = None
    a = None
    iters = 0
    for _ in range(max_reps):
        C, a = lloyds_iter(P, d2, C)
        iters += 1
        if prev is not None and np.array_equal(prev, a):
            break
        prev = a
    return C, a, iters


def kmeans_objective(P, C, a):
    """Harness-side objective sum_d ||P_d - c_a(d)||^2 in fp64 (the reference never computes
    one: SURVEY Q13); used identically for both sides of the parity check."""
    diff = P.astype(np.float64) - C.astype(np.float64)[a]
    return float((diff * diff).sum())


def kmeanspp(P, k, rng):
    """src/sparseMatrix.cpp:2133-2209: D^2 sampling with batched draws (1 + sqrt(max(s-5,0))
    new centers per distance refresh); clamp >= 0 (:2116,2122); fp32 serial prefix sum
    (:2170-2172); center = upper_bound(cumul, t) - 1; duplicates skipped (:2190)."""
    D = P.shape[0]
    d2 = docs_l2sq(P)
    centers = [int(rng.integers(0, D))]
    min_dist = np.full(D, np.finfo(F32).max, dtype=F32)
    new_added = 1
    while len(centers) < k:
        Cn = P[centers[len(centers) - new_added:]]
        dm = np.maximum(dist_matrix(P, d2, Cn), F32(0))
        min_dist = np.minimum(min_dist, dm.min(1)).astype(F32)
        cumul = np.concatenate([[F32(0)], np.cumsum(min_dist, dtype=F32)])
        s = len(centers)
        new_added = 0
        c = 0
        while c < 1 + math.sqrt(max(s - 5, 0)) and len(centers) < k:
            t = float(cumul[-1]) * rng.random()
            nc = int(np.searchsorted(cumul, t, side="right") - 1)
            if nc not in centers and nc < D:
                centers.append(nc)
                new_added += 1
            c += 1
    return np.array(centers, dtype=np.int64), P[centers].astype(F32)


def lift_centers(U, C_lowd):
    """src/sparseMatrix.cpp:1438-1450 left_multiply_by_U_Spectra: centers (V x k) = U C^T."""
    return (U.astype(F32) @ C_lowd.T.astype(F32)).astype(F32)


# --------------------------------------------------------------------------- stage F
def docs_l2sq_full(B_csc):
    """src/sparseMatrix.cpp:1670-1677 compute_docs_l2sq: sum of squared entries per column of B."""
    Bc = B_csc.tocsc()
    sq = (Bc.data.astype(F32) * Bc.data.astype(F32)).astype(F32)
    return np.add.reduceat(np.concatenate([sq, [F32(0)]]), Bc.indptr[:-1]).astype(F32) * (np.diff(Bc.indptr) > 0)


def dist_matrix_full(B_csc, d2, C):
    """src/sparseMatrix.cpp:1494-1552 distsq_docs_to_centers: ((-2 B^T C^T) + ||c||^2) + ||d||^2.
    C is (k, V): center c = row c (the reference's `centers + c * vocab_size`)."""
    C = C.astype(F32)
    c2 = np.einsum("ij,ij->i", C, C, dtype=F32).astype(F32)
    G = (F32(-2.0) * np.asarray(B_csc.T.astype(F32) @ C.T, dtype=F32)).astype(F32)
    return ((G + c2[None, :]).astype(F32) + d2[:, None]).astype(F32)


def lloyds_iter_full(B_csc, d2, C):
    """src/sparseMatrix.cpp:1584-1667 lloyds_iter: assign by cblas_isamin (first index of min |x|),
    then center = sum of member columns / cluster size; an empty cluster's center stays zero."""
    import scipy.sparse as sp
    k = C.shape[0]
    a = np.argmin(np.abs(dist_matrix_full(B_csc, d2, C)), axis=1).astype(np.uint32)
    D = B_csc.shape[1]
    M = sp.csr_matrix((np.ones(D, dtype=np.float64), (np.arange(D), a.astype(np.int64))), shape=(D, k))
    newC = np.asarray((B_csc.astype(np.float64) @ M).todense()).T.copy()
    cnt = np.bincount(a, minlength=k)
    nz = cnt > 0
    newC[nz] /= cnt[nz][:, None]
    return newC.astype(F32), a


def run_lloyds_full(B_csc, C0, max_reps=10):
    """src/sparseMatrix.cpp:1679-1746 run_lloyds on the full-dimensional B; stops when the partition
    repeats.  (The reference's check lags by up to two iterations on an already fixed partition --
    prev_closest_docs is only refreshed when the cluster sizes did not change, :1718-1733 -- which
    cannot change the centers or the partition it returns.)"""
    d2 = docs_l2sq_full(B_csc)
    C = C0.astype(F32).copy()
    prev, a, iters = None, None, 0
    for _ in range(max_reps):
        C, a = lloyds_iter_full(B_csc, d2, C)
        iters += 1
        if prev is not None and np.array_equal(prev, a):
            break
        prev = a
    return C, a, iters


def kmeans_objective_full(B_csc, C, a):
    """Harness-side objective sum_d ||B_d - c_a(d)||^2 in fp64."""
    Bc = B_csc.tocsc().astype(np.float64)
    C = C.astype(np.float64)
    c2 = np.einsum("ij,ij->i", C, C)
    d2 = np.asarray(Bc.multiply(Bc).sum(0)).ravel()
    docs = np.repeat(np.arange(Bc.shape[1]), np.diff(Bc.indptr))
    dots = np.bincount(docs, weights=Bc.data * C[a.astype(np.int64)[docs], Bc.indices], minlength=Bc.shape[1])
    return float((d2 - 2.0 * dots + c2[a.astype(np.int64)]).sum())


# --------------------------------------------------------------------------- stage G
def catchword_rank(num_docs: int, k: int) -> int:
    """src/trainer.cpp:583: r = floor(eps2 * w0 * (float)num_docs / (float)(2 k)), eps2 = 1/3 and w0 = 1 as
    doubles (include/hyperparams.h:8-10), the document count and 2k as floats."""
    return int(math.floor((1.0 / 3.0) * 1.0 * float(F32(num_docs)) / float(F32(2.0 * k))))


def rth_highest_element(vals, rows, offsets, V: int, docs, r: int):
    """src/sparseMatrix.cpp:491-524: per word the r-th highest normalised value among the documents of one
    cluster when the word occurs in MORE than r of them; otherwise 0 -- except when r >= cluster size and the
    word occurs in every document of the cluster, then the smallest value."""
    thr = np.zeros(V, dtype=F32)
    docs = np.asarray(docs, dtype=np.int64)
    if len(docs) == 0:
        return thr
    lens = (offsets[docs + 1] - offsets[docs]).astype(np.int64)
    idx = np.repeat(offsets[docs], lens) + (np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens))
    w, v = rows[idx].astype(np.int64), vals[idx].astype(F32)
    order = np.lexsort((-v.astype(np.float64), w))          # by word, value descending
    w, v = w[order], v[order]
    cnt = np.bincount(w, minlength=V)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    big = cnt > r
    thr[big] = v[start[big] + r - 1]
    if r >= len(docs):
        full = (~big) & (cnt == len(docs))
        thr[full] = v[start[full] + cnt[full] - 1]
    return thr


def find_catchwords(thr, rho=1.1):
    """src/sparseMatrix.cpp:573-594: word w is a catchword of topic t iff thr[t, w] > rho * thr[o, w] for every
    other topic o (fp32 compare of thr against the fp64 product, as the reference's expression evaluates).
    thr is (k, V).  Returns a list of k ascending word arrays."""
    k, V = thr.shape
    t64 = thr.astype(np.float64)
    out = []
    for t in range(k):
        others = np.delete(t64, t, axis=0)
        ok = np.all(t64[t][None, :] > rho * others, axis=0) if k > 1 else np.zeros(V, bool)
        out.append(np.nonzero(ok)[0].astype(np.int64))
    return out


def construct_topic_model(vals, rows, offsets, V: int, k: int, cluster_of_doc, catchwords, eps3=5.0, w0=1.0):
    """src/sparseMatrix.cpp:597-838 construct_topic_model.  ``cluster_of_doc`` uint32[D] (0xFFFFFFFF = none) is the
    closest_docs partition of the original documents, ``catchwords`` a list of k word arrays.  Returns
    (Model float32[V, k] column-stochastic, dts = (doc, topic, sum) arrays in (doc, topic) order, top_topic_pairs
    int array [n, 3] in document order).

    * doc_topic_sum[doc][topic] = fp32 sum, in position order, of the document's values on topic's catchwords (:652-668);
      non-zero entries are listed by (doc, topic) (:669-678)
    * top two topics per doc by strict > in topic order (:683-703)
    * model_threshold[topic] = rank_threshold-th largest sum of the topic (0 when fewer entries or no catchwords), with
      rank_threshold = (uint)(eps3 w0 (float)D / ((float)k 2.0)) (:716-751)
    * Model[:, topic] += every document whose sum for the topic exceeds the threshold, plus -- for EVERY document of a
      cluster, the variable is misnamed doc_in_catchless_topic -- the document's own cluster (:787-817); columns are then
      scaled to unit l1 norm (:822-826).  The reference accumulates in fp32 in document order; this restatement does the
      same per column through np.add.at on float32 (same order of additions)."""
    D = len(offsets) - 1
    topic_of_word = np.full(V, -1, dtype=np.int64)
    for t in range(k):
        topic_of_word[np.asarray(catchwords[t], dtype=np.int64)] = t
    lens = np.diff(offsets).astype(np.int64)
    doc_of = np.repeat(np.arange(D, dtype=np.int64), lens)
    tw = topic_of_word[rows.astype(np.int64)]
    sel = tw >= 0
    # fp32 sequential sums per (doc, topic) in position order
    key = doc_of[sel] * k + tw[sel]
    v = vals[sel].astype(F32)
    order = np.argsort(key, kind="stable")
    key_s, v_s = key[order], v[order]
    uniq, start = np.unique(key_s, return_index=True)
    sums = np.zeros(len(uniq), dtype=F32)
    pos = start.copy()
    end = np.concatenate([start[1:], [len(key_s)]])
    live = np.arange(len(uniq))
    while len(live):                                   # add the i-th value of every (doc, topic) run in lock step
        sums[live] = (sums[live] + v_s[pos[live]]).astype(F32)
        pos[live] += 1
        live = live[pos[live] < end[live]]
    nzm = sums != 0
    dts_doc, dts_topic, dts_val = (uniq[nzm] // k).astype(np.int64), (uniq[nzm] % k).astype(np.int64), sums[nzm]
    # top two topics per document (strict >, topic order)
    pairs = []
    bounds = np.searchsorted(dts_doc, np.arange(D + 1))
    for d in range(D):
        mx = mx2 = F32(0.0)
        t1 = t2 = -1
        for i in range(bounds[d], bounds[d + 1]):
            x = dts_val[i]
            if x > mx:
                mx2, t2 = mx, t1
                mx, t1 = x, int(dts_topic[i])
            elif x > mx2:
                mx2, t2 = x, int(dts_topic[i])
        if t1 >= 0 and t2 >= 0:
            pairs.append((t1, t2, d))
    rank = int(np.uint64(eps3 * w0 * float(F32(D)) / (float(F32(k)) * 2.0)))
    thr = np.zeros(k, dtype=F32)
    for t in range(k):
        if len(catchwords[t]) > 0:
            x = np.sort(dts_val[dts_topic == t])[::-1]
            if len(x) >= rank and rank >= 1:
                thr[t] = x[rank - 1]
    Model = np.zeros((V, k), dtype=F32)
    cl = np.asarray(cluster_of_doc).astype(np.int64)
    # additions in document order: first the document's qualifying (doc, topic) entries in topic order, then its own cluster
    add_doc = np.concatenate([dts_doc[dts_val > thr[dts_topic]], np.nonzero(cl != 0xFFFFFFFF)[0]])
    add_topic = np.concatenate([dts_topic[dts_val > thr[dts_topic]], cl[cl != 0xFFFFFFFF]])
    add_second = np.concatenate([np.zeros(int((dts_val > thr[dts_topic]).sum()), np.int64), np.ones(int((cl != 0xFFFFFFFF).sum()), np.int64)])
    o = np.lexsort((add_topic, add_second, add_doc))
    add_doc, add_topic = add_doc[o], add_topic[o]
    for t in range(k):
        docs = add_doc[add_topic == t]
        if len(docs) == 0:
            continue
        ln = lens[docs]
        idx = np.repeat(offsets[docs], ln) + (np.arange(ln.sum()) - np.repeat(np.cumsum(ln) - ln, ln))
        col = np.zeros(V, dtype=F32)
        np.add.at(col, rows[idx].astype(np.int64), vals[idx].astype(F32))      # unbuffered, in index order, fp32
        Model[:, t] = col
    s = np.abs(Model).sum(0, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        Model = (Model * (F32(1.0) / s.astype(F32))[None, :]).astype(F32)
    return Model, (dts_doc, dts_topic, dts_val), np.array(pairs, dtype=np.int64).reshape(-1, 3), thr


# --------------------------------------------------------------------------- comparisons
def principal_angle_sin(U1, U2):
    """sin of the largest principal angle between span(U1) and span(U2) (orthonormal cols)."""
    Q1, _ = np.linalg.qr(U1.astype(np.float64))
    Q2, _ = np.linalg.qr(U2.astype(np.float64))
    R = Q2 - Q1 @ (Q1.T @ Q2)
    return float(np.linalg.norm(R, 2))
