"""The spectral core of ``ISLETrainer::train()`` (reference src/trainer.cpp:430-554), stage by
stage, on top of the device-backed matrix classes.  This is the call a user of the Python
side makes; the C++ side is the unchanged reference trainer.cpp linked against
isle_b200/shim/sparseMatrix_cuda.cpp."""
from __future__ import annotations

import dataclasses
import time
from typing import Optional

import numpy as np

from ._capi import Context
from .sparse_matrix import (BLOCK_KS_BLOCK_SIZE, BLOCK_KS_MAX_ITERS, BLOCK_KS_TOLERANCE, KMEANS_INIT_REPS,
                            MAX_KMEANS_LOWD_REPS, FPSparseMatrix, SparseMatrix)


@dataclasses.dataclass
class SpectralCoreResult:
    zetas: np.ndarray            # float32[V]
    new_nnzs: int
    original_cols: np.ndarray    # uint64[D_B]
    D_B: int
    nnz_B: int
    evalues: np.ndarray          # float32[k], sigma^2 descending
    seeds: np.ndarray            # uint64[k]
    centers_lowd: np.ndarray     # float32[k,k] after Lloyd
    centers: np.ndarray          # float32[V,k] = U centers_lowd
    lloyd_iters: int
    objective: float
    stage_seconds: dict
    U: Optional[np.ndarray] = None


def spectral_core(ctx: Context, V: int, D: int, k: int, normalized_vals, rows, offsets, avg_doc_sz: float,
                  nz_docs: int, *, sample_rate: float = 0.0, select_docs=None, seed: int = 0,
                  block_size: int = BLOCK_KS_BLOCK_SIZE, want_U: bool = False,
                  lift: bool = True) -> SpectralCoreResult:
    """Stages A-E of train() with HOST inputs and HOST outputs (the e2e path of bench.py)."""
    t = {}
    t0 = time.perf_counter()
    A = SparseMatrix(V, D, ctx)
    A.populate_normalized(normalized_vals, rows, offsets, avg_doc_sz, nz_docs)
    t["upload_A"] = time.perf_counter() - t0

    t0 = time.perf_counter()                                   # trainer.cpp:430-472
    freqs = A.list_word_freqs_by_sorting()
    zetas, new_nnzs = A.compute_thresholds(0, V, freqs, k)
    t["thresholds"] = time.perf_counter() - t0

    t0 = time.perf_counter()                                   # trainer.cpp:475-485
    B = FPSparseMatrix(A)
    if sample_rate > 0.0 or select_docs is not None:
        oc = B.sampled_threshold_and_copy(A, zetas, new_nnzs, sample_rate, np.random.default_rng(seed), select_docs)
    else:
        oc = B.threshold_and_copy(A, zetas, new_nnzs)
    t["build_B"] = time.perf_counter() - t0

    t0 = time.perf_counter()                                   # trainer.cpp:490-502
    B.initialize_for_eigensolver(k)
    r = B.compute_block_ks(k, block_size=block_size, max_iters=BLOCK_KS_MAX_ITERS, tol=BLOCK_KS_TOLERANCE,
                           seed=seed, want_U=want_U)
    evalues, U = (r if want_U else (r, None))
    t["block_ks"] = time.perf_counter() - t0

    t0 = time.perf_counter()                                   # trainer.cpp:511-533
    seeds, centers_lowd, _res = B.kmeans_init_on_projected_space(k, KMEANS_INIT_REPS, seed=seed)
    t["kmeanspp"] = time.perf_counter() - t0

    t0 = time.perf_counter()                                   # trainer.cpp:539-553
    B.run_lloyds_on_projected_space(k, centers_lowd, None, MAX_KMEANS_LOWD_REPS)
    centers = B.left_multiply_by_U_Spectra(centers_lowd, k, k) if lift else np.zeros((0, 0), np.float32)
    t["lloyd"] = time.perf_counter() - t0
    ll = B.last_lloyd
    B.cleanup_after_eigensolver()                              # trainer.cpp:554
    return SpectralCoreResult(zetas, new_nnzs, oc, B.num_docs(), B.get_nnzs(), evalues, seeds, centers_lowd,
                              centers, ll["iters"], ll["objective"], t, U)
