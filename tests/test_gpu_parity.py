"""GPU parity suite (-m gpu): every call goes through the C ABI of libisle_cuda.so (via the
host mirror isle_b200.sparse_matrix) and is checked against the oracle and the golden
fixtures.  Bars (BASELINE.json north star): thresholds and B bit-exact; singular values
within 1e-4 relative, principal angle < 1e-3; identical Lloyd assignments from identical
projection and initial centers (ties excepted); objective within 1e-4."""
import hashlib
import os

import numpy as np
import pytest

from oracle import isle_oracle as O

pytestmark = pytest.mark.gpu


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def make_AB(ctx, V, D, k, vals, rows, offsets, avg, nz, mask=None):
    from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix
    A = SparseMatrix(V, D, ctx)
    A.populate_normalized(vals, rows, offsets, avg, nz)
    zetas, nn = A.compute_thresholds(0, V, A.list_word_freqs_by_sorting(), k)
    B = FPSparseMatrix(A)
    if mask is None:
        oc = B.threshold_and_copy(A, zetas, nn)
    else:
        oc = B.sampled_threshold_and_copy(A, zetas, nn, 0.0, select_docs=mask)
    return A, B, zetas, nn, oc


def tiny_AB(ctx, g, mask=None):
    return make_AB(ctx, int(g["V"]), int(g["D"]), int(g["k"]), g["A_normalized_vals"], g["rows"], g["offsets"],
                   float(g["avg_doc_sz"]), int(g["D"]), mask)


# ---------------------------------------------------------------- family (1): bit-exact
def test_thresholds_and_B_bit_exact_tiny(ctx, golden_tiny):
    g = golden_tiny
    A, B, zetas, nn, oc = tiny_AB(ctx, g)
    assert np.array_equal(zetas, g["zetas"])
    assert nn == int(g["new_nnzs"])
    vals, rows, offs, oc2 = B.download()
    assert np.array_equal(vals, g["B_vals"])
    assert np.array_equal(rows, g["B_rows"].astype(np.uint64))
    assert np.array_equal(offs, g["B_offsets"])
    assert np.array_equal(oc, g["B_original_cols"].astype(np.uint64)) and np.array_equal(oc, oc2)
    assert abs(B.frobenius() - float(g["frobenius"])) <= 1e-6 * float(g["frobenius"])


def test_background_download_of_B(ctx, golden_c1, corpus_c1):
    """isle_cuda_download_B_begin / _end (what the shim's threshold_and_copy uses): the same arrays as the synchronous copy,
    with the eigensolver running in between."""
    from isle_b200._capi import ptr
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, zetas, nn, oc = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    bv, br, bo, boc = B.download()
    v2, r2, o2, c2 = np.zeros_like(bv), np.zeros_like(br), np.zeros_like(bo), np.zeros_like(boc)
    ctx.call("isle_cuda_download_B_begin", ptr(v2), ptr(r2), ptr(o2), ptr(c2))
    ev = B.compute_block_ks(c.k, seed=1)
    ctx.call("isle_cuda_download_B_end")
    assert np.array_equal(v2, bv) and np.array_equal(r2, br) and np.array_equal(o2, bo) and np.array_equal(c2, boc)
    assert sha(v2, r2.astype(np.uint32), o2, c2.astype(np.uint32)) == str(g["B_sha"])
    ctx.call("isle_cuda_download_B_end")                             # idempotent
    v3 = np.zeros_like(bv)
    ctx.call("isle_cuda_download_B_begin", ptr(v3), None, None, None)
    A2, B2, *_ = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)     # rebuilding B ends the pending download first
    assert np.array_equal(v3, bv)


def test_u64_upload_path(ctx, golden_tiny):
    g = dict(golden_tiny)
    g["rows"] = g["rows"].astype(np.uint64)           # the reference's own index width
    A, B, zetas, nn, oc = tiny_AB(ctx, g)
    assert np.array_equal(zetas, g["zetas"]) and B.get_nnzs() == int(g["nnz_B"])


def test_masked_B_bit_exact(ctx, golden_tiny, golden_tiny_masked):
    g, m = golden_tiny, golden_tiny_masked
    A, B, zetas, nn, oc = tiny_AB(ctx, g, mask=m["mask"])
    vals, rows, offs, oc2 = B.download()
    assert np.array_equal(vals, m["B_vals"]) and np.array_equal(rows, m["B_rows"].astype(np.uint64))
    assert np.array_equal(offs, m["B_offsets"]) and np.array_equal(oc2, m["B_original_cols"].astype(np.uint64))


def test_c1_thresholds_and_B_bit_exact(ctx, golden_c1, corpus_c1):
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, zetas, nn, oc = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    assert np.array_equal(zetas, g["zetas"]) and nn == int(g["new_nnzs"])
    bv, br, bo, boc = B.download()
    assert sha(bv, br.astype(np.uint32), bo, boc.astype(np.uint32)) == str(g["B_sha"])


def test_sampling_weights(ctx, golden_tiny):
    g = golden_tiny
    A, B, zetas, nn, oc = tiny_AB(ctx, g)
    w = np.zeros(int(g["D"]), np.float32)
    from isle_b200._capi import ptr
    ctx.call("isle_cuda_sampling_weights", ptr(w))
    assert np.array_equal(w, O.sampling_weights(g["A_normalized_vals"], g["rows"], g["offsets"], g["zetas"]))


def test_sample_docs_on_device(ctx, golden_tiny):
    """the A-Res selection of sampled_threshold_and_copy (src/sparseMatrix.cpp:1399-1415) on the device
    against a numpy replay of the same counter-based uniforms; powf may differ from numpy in the last place, so keys
    within a few ulps of the pivot may fall on either side."""
    from isle_b200._capi import ptr
    import ctypes as C
    g = golden_tiny
    A, B, zetas, nn, oc = tiny_AB(ctx, g)
    D = int(g["D"])
    w = np.zeros(D, np.float32)
    ctx.call("isle_cuda_sampling_weights", ptr(w))

    def mix(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return x ^ (x >> np.uint64(31))

    for rate, seed in ((0.4, 7), (0.05, 8), (0.999, 9)):
        sel = np.zeros(D, np.uint8)
        nsel = C.c_uint64()
        ctx.call("isle_cuda_sample_docs", C.c_float(rate), seed, ptr(sel), C.byref(nsel))
        with np.errstate(over="ignore"):
            u = (mix(np.uint64(seed) ^ mix(np.arange(D, dtype=np.uint64))) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        with np.errstate(divide="ignore"):
            key = np.where(w == 0, np.float32(0), np.power(u, (np.float32(1) / w).astype(np.float32), dtype=np.float32)).astype(np.float32)
        nth = int(np.float32(rate) * np.float32(D))
        pivot = np.sort(key)[::-1][nth]
        ref = key >= pivot
        assert int(nsel.value) == int(sel.sum())
        assert abs(int(sel.sum()) - int(ref.sum())) <= 2 and int((sel.astype(bool) != ref).sum()) <= 4
        assert int(sel.sum()) >= nth + 1 - 2
        # zero-weight documents are only kept when the pivot itself is 0
        assert pivot == 0 or not np.any(sel.astype(bool) & (w == 0))
    sel = np.zeros(D, np.uint8)
    ctx.call("isle_cuda_sample_docs", C.c_float(1.0), 1, ptr(sel), None)      # floor(rate D) >= D: everything
    assert sel.all()
    # the masked build accepts the device selection (sampled_threshold_and_copy with device_seed)
    from isle_b200.sparse_matrix import FPSparseMatrix
    B2 = FPSparseMatrix(A)
    oc2 = B2.sampled_threshold_and_copy(A, zetas, nn, 0.4, device_seed=7)
    assert B2.num_docs() <= B2.last_sample_count and len(oc2) == B2.num_docs()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_thresholds_ragged_edge_cases(ctx, seed):
    """Empty docs, absent words, one-word docs, values that round to zero, heavy ties."""
    rng = np.random.default_rng(seed)
    V, D, k = 257, 1203, 10
    lens = rng.integers(0, 40, size=D)
    lens[rng.random(D) < 0.1] = 0                       # empty docs
    lens[:5] = [0, 1, 1, 0, 39]
    rows, counts = [], []
    for L in lens:
        w = np.sort(rng.choice(V - 7, size=L, replace=False))     # last 7 words never occur
        rows.append(w)
        counts.append(rng.choice([1, 1, 1, 2, 3, 7, 30], size=L))
    offsets = np.zeros(D + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    rows = np.concatenate(rows).astype(np.uint32)
    counts = np.concatenate(counts).astype(np.uint32)
    vals, avg, nz = O.normalize_docs(counts, offsets)
    z_ref, nn_ref = O.compute_thresholds(vals, rows, V, nz, k)
    bv, br, bo, oc = O.threshold_and_copy(vals, rows, offsets, z_ref)
    A, B, zetas, nn, goc = make_AB(ctx, V, D, k, vals, rows, offsets, avg, nz)
    assert np.array_equal(zetas, z_ref) and nn == nn_ref
    gv, gr, go, goc2 = B.download()
    assert np.array_equal(gv, bv) and np.array_equal(gr, br.astype(np.uint64))
    assert np.array_equal(go, bo) and np.array_equal(goc, oc.astype(np.uint64))
    assert B.num_docs() == len(oc) < D                   # some docs were dropped


def test_value_above_avg_doc_sz_is_rejected(ctx):
    """The reference asserts value <= avg_doc_sz (src/sparseMatrix.cpp:380); we fail loudly."""
    from isle_b200._capi import IsleCudaError
    from isle_b200.sparse_matrix import SparseMatrix
    A = SparseMatrix(4, 2, ctx)
    A.populate_normalized(np.array([50.0, 1.0], np.float32), np.array([0, 1], np.uint32),
                          np.array([0, 1, 2], np.int64), 3.0, 2)
    with pytest.raises(IsleCudaError):
        A.compute_thresholds(0, 4, None, 2)


# ---------------------------------------------------------------- family (2): the operator
@pytest.mark.parametrize("b", [1, 4, 10, 16])
def test_spsptr_multiply_matches_oracle(ctx, golden_tiny, b):
    g = golden_tiny
    A, B, *_ = tiny_AB(ctx, g)
    Bo = O.to_csc(g["B_vals"], g["B_rows"], g["B_offsets"], int(g["V"]))
    X = np.random.default_rng(b).standard_normal((int(g["V"]), b)).astype(np.float32)
    Z = B.multiply(X)
    Zr = (Bo @ (Bo.T @ X.astype(np.float64)))
    assert np.max(np.abs(Z - Zr)) <= 2e-6 * np.max(np.abs(Zr))


def test_spsptr_multiply_c1_split_rows(ctx, golden_c1, corpus_c1):
    """c1 has word rows longer than one work item (split + atomics path) ."""
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, zetas, nn, oc = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    bv, br, bo, _ = O.threshold_and_copy(vals, c.rows, c.offsets, g["zetas"])
    Bo = O.to_csc(bv, br, bo, c.V)
    assert np.diff(Bo.tocsr().indptr).max() > 2048
    X = np.random.default_rng(0).standard_normal((c.V, 10)).astype(np.float32)
    Z = B.multiply(X)
    Zr = (Bo @ (Bo.T @ X.astype(np.float64)))
    assert np.max(np.abs(Z - Zr)) <= 5e-6 * np.max(np.abs(Zr))
    # linearity (size independent property)
    Z2 = B.multiply(2.0 * X + 1.0)
    Z1 = B.multiply(np.ones_like(X))
    assert np.max(np.abs(Z2 - (2 * Z + Z1))) <= 1e-4 * np.max(np.abs(Z2))


# every engine combination of the operator must give the same product: dense head on tcgen05 vs index
# lists only, one-sector (block floating point) operand rows vs padded fp32 rows, forked vs serial streams
ENGINE_OPTS = [
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=0, spmm_head_i8=0),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=1, spmm_head_i8=0, spmm_head_density_ppm=1000, spmm_head_max=4096, spmm_head_seg=2),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=1, spmm_head_density_ppm=1000, spmm_head_max=8192, spmm_head8_slab=1, spmm_head8_stages=2),
    dict(spmm_head=0, spmm_bfp=0, spmm_fork=0),
    dict(spmm_head=0, spmm_bfp=1, spmm_fork=0),
    dict(spmm_head=1, spmm_bfp=0, spmm_fork=0),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=0),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=1),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=1, spmm_head_i8=0, spmm_head_density_ppm=60000, spmm_head_seg=1),
    dict(spmm_head=1, spmm_bfp=1, spmm_fork=1, spmm_head_density_ppm=1000, spmm_head_max=4096, spmm_head_seg=2),
]
ENGINE_DEFAULTS = dict(spmm_head=1, spmm_bfp=1, spmm_fork=1, spmm_head_density_ppm=12000, spmm_head_max=4096,
                       spmm_head_seg=16, spmm_head_i8=1, spmm_head8_slab=0, spmm_head8_stages=4)


@pytest.mark.parametrize("opts", ENGINE_OPTS, ids=lambda o: "-".join(f"{k[5:]}{v}" for k, v in o.items()))
@pytest.mark.parametrize("b", [3, 10, 16])
def test_spsptr_engines_agree(ctx, golden_c1, corpus_c1, opts, b):
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    bv, br, bo, _ = O.threshold_and_copy(vals, c.rows, c.offsets, g["zetas"])
    Bo = O.to_csc(bv, br, bo, c.V)
    X = np.random.default_rng(7 + b).standard_normal((c.V, b)).astype(np.float32)
    X[:, 0] *= 1e-3                      # columns of very different scale share a block-FP unit
    X[::7, :] = 0.0                      # all-zero operand rows
    Zr = (Bo @ (Bo.T @ X.astype(np.float64)))
    try:
        for k_, v_ in {**ENGINE_DEFAULTS, **opts}.items():
            ctx.set_option(k_, v_)
        A, B, zetas, nn, oc = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
        Z = B.multiply(X)
        H = int(ctx.stat("spmm_head_words"))
        assert (H > 0) == bool(opts["spmm_head"])
        if opts.get("spmm_head_density_ppm") == 1000:
            assert H >= 1024             # K split into many segments, partial sums added atomically
        col_scale = np.max(np.abs(Zr), axis=0)
        assert np.max(np.abs(Z - Zr) / col_scale) <= 3e-6
    finally:
        for k_, v_ in ENGINE_DEFAULTS.items():
            ctx.set_option(k_, v_)


# ---------------------------------------------------------------- stage C: eigenpairs
def check_eigs(ev, U, ev_ref, U_ref, frob):
    k = len(ev)
    s, s_ref = np.sqrt(ev), np.sqrt(ev_ref)
    assert np.all(np.diff(ev) <= 1e-3 * ev[:-1])                       # descending
    assert np.max(np.abs(s - s_ref) / s_ref) < 1e-4                    # singular values
    assert O.principal_angle_sin(U, U_ref) < 1e-3                      # principal subspace
    assert np.linalg.norm(U.T.astype(np.float64) @ U - np.eye(k)) < 1e-4
    assert ev.sum() <= frob * (1 + 1e-6)


def test_block_ks_tiny(ctx, golden_tiny):
    g = golden_tiny
    k, V = int(g["k"]), int(g["V"])
    A, B, *_ = tiny_AB(ctx, g)
    B.initialize_for_eigensolver(k)
    ev, U = B.compute_block_ks(k, seed=1, want_U=True)
    assert B.nconv == k
    check_eigs(ev, U, g["evalues"], g["U_colmajor"].reshape(k, V).T, float(g["frobenius"]))
    # residual check against the operator itself: ||B B^T u - lambda u|| <= tol * lambda
    Z = B.multiply(U[:, :10].copy())
    r = np.linalg.norm(Z - U[:, :10] * ev[:10], axis=0) / ev[:10]
    assert r.max() < 5e-4


def test_block_ks_c1(ctx, golden_c1, corpus_c1):
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, *_ = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    ev, U = B.compute_block_ks(c.k, seed=2, want_U=True)
    check_eigs(ev, U, g["evalues"], g["U_colmajor"].reshape(c.k, c.V).T, float(g["frobenius"]))


@pytest.mark.parametrize("engine", [0, 1, 2], ids=["fma-scalar", "fma-vector", "tcgen05"])
@pytest.mark.parametrize("shape", [(600, 20, 10), (1000, 7, 3), (4096, 130, 16), (20004, 129, 1), (20000, 300, 10)])
def test_panel_product_engines(ctx, engine, shape):
    """One block Gram-Schmidt pass (restarted_block_ks.h:83-84), C = W^T F and F -= W C, on random
    orthonormal W against float64: every engine of the device solver, ragged shapes included."""
    from isle_b200._capi import ptr
    n, rows, b = shape
    rng = np.random.default_rng(n + rows + b)
    Q, _ = np.linalg.qr(rng.standard_normal((n, rows)))
    W = Q.astype(np.float32)
    F = (rng.standard_normal((n, b)) * np.logspace(0, -3, b)[None, :]).astype(np.float32)   # columns of very different scale
    Wc, Fc, Cc = np.ascontiguousarray(W.T), F.T.copy(), np.zeros((b, rows), np.float32)     # copy: F.T aliases F when b = 1
    ctx.call("isle_cuda_panel_products", n, rows, b, ptr(Wc), ptr(Fc), ptr(Cc), engine)
    fn = np.linalg.norm(F.astype(np.float64), axis=0)
    C_ref = W.astype(np.float64).T @ F.astype(np.float64)
    assert np.max(np.abs(Cc.T - C_ref) / fn[None, :]) < 2e-6
    F_ref = F.astype(np.float64) - W.astype(np.float64) @ Cc.T.astype(np.float64)
    assert np.max(np.abs(Fc.T - F_ref) / fn[None, :]) < 2e-6


@pytest.mark.parametrize("opts", [dict(ks_panel_tc=0, ks_panel_v2=0), dict(ks_panel_tc=0, ks_panel_v2=1), dict(ks_panel_tc=1),
                                  dict(ks_panel_tc=1, ks_gs_elide=0),
                                  dict(ks_panel_tc=0, ks_panel_v2=0, ks_custom_orth=0, ks_fast_qr=0),
                                  dict(ks_panel_tc=1, ks_defer_rank=0), dict(ks_panel_tc=1, ks_force_qr_fallback=3),
                                  dict(ks_panel_tc=1, dense_tc_min_mflops=0)],
                         ids=["scalar-panels", "vector-panels", "tcgen05-panels", "tcgen05-panels-3-passes", "cublas-mgs",
                              "rank-read-every-step", "forced-qr-fallback-and-redo", "truncation-gemm-3xtf32"])
def test_block_ks_engine_variants_c1(ctx, golden_c1, corpus_c1, opts):
    """Every panel / QR engine of the device solver meets the same bar against the reference's output."""
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, *_ = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    defaults = dict(ks_panel_tc=1, ks_panel_v2=1, ks_custom_orth=1, ks_fast_qr=1, ks_gs_elide=1, ks_defer_rank=1,
                    ks_force_qr_fallback=0, dense_tc_min_mflops=20000)
    try:
        for k_, v_ in {**defaults, **opts}.items():
            ctx.set_option(k_, v_)
        ev, U = B.compute_block_ks(c.k, seed=3, want_U=True)
        elided = ctx.stat("ks_gs_elided")
        if opts.get("ks_force_qr_fallback"):      # the deferred check found the (simulated) bad pivots and re-did those steps
            assert ctx.stat("ks_qr_fallbacks") >= 3
        # the third Gram-Schmidt pass is elided on the device only by the tensor-core engine, and only when allowed
        assert (elided > 0) == (opts.get("ks_panel_tc") == 1 and opts.get("ks_gs_elide", 1) == 1)
    finally:
        for k_, v_ in defaults.items():
            ctx.set_option(k_, v_)
    check_eigs(ev, U, g["evalues"], g["U_colmajor"].reshape(c.k, c.V).T, float(g["frobenius"]))


def test_block_ks_rejects_bad_k(ctx, golden_tiny):
    from isle_b200._capi import IsleCudaError
    A, B, *_ = tiny_AB(ctx, golden_tiny)
    with pytest.raises(IsleCudaError):
        B.compute_block_ks(15)          # reference would write V out of bounds (k % b != 0)


# ---------------------------------------------------------------- family (3): k-means
def setup_projection(ctx, g):
    k, V = int(g["k"]), int(g["V"])
    A, B, *_ = tiny_AB(ctx, g)
    U_ref = g["U_colmajor"].reshape(k, V).T.copy()
    B.set_U(U_ref)
    Bo = O.to_csc(g["B_vals"], g["B_rows"], g["B_offsets"], V)
    return B, O.project(Bo, U_ref), k


@pytest.mark.parametrize("engine", [0, 1])
def test_projection_and_assignment_match_oracle(ctx, golden_tiny, engine):
    g = golden_tiny
    ctx.set_option("dist_kernel", engine)
    B, P_ref, k = setup_projection(ctx, g)
    P, l2 = B.projected_docs()
    assert np.max(np.abs(P - P_ref)) <= 1e-5 * np.max(np.abs(P_ref))
    assert np.allclose(l2, O.docs_l2sq(P_ref), rtol=1e-5)
    C0 = g["centers_lowd_init"].reshape(k, k)
    a = B.projected_closest_centers(k, C0)
    a_ref = O.closest_centers(P_ref, O.docs_l2sq(P_ref), C0)
    assert int((a != a_ref).sum()) == 0
    ctx.set_option("dist_kernel", 1)


@pytest.mark.parametrize("engine", [0, 1])
def test_lloyd_matches_reference(ctx, golden_tiny, engine):
    g = golden_tiny
    ctx.set_option("dist_kernel", engine)
    B, P_ref, k = setup_projection(ctx, g)
    C = g["centers_lowd_init"].reshape(k, k).copy()
    closest = [[] for _ in range(k)]
    B.run_lloyds_on_projected_space(k, C, closest, 10)
    a = B.last_lloyd["assign"]
    assert np.array_equal(a, g["lloyd_assign"])                       # identical partition
    assert sorted(sum(closest, [])) == list(range(B.num_docs()))
    Cr = g["centers_lowd_final"].reshape(k, k)
    assert np.max(np.abs(C - Cr)) < 1e-4
    obj_ref = O.kmeans_objective(P_ref, Cr, g["lloyd_assign"])
    assert abs(B.last_lloyd["objective"] - obj_ref) / obj_ref < 1e-4
    lifted = B.left_multiply_by_U_Spectra(C, k, k)
    assert np.max(np.abs(lifted - O.lift_centers(g["U_colmajor"].reshape(k, int(g["V"])).T, C))) < 1e-4
    lift_ref = g["U_colmajor"].reshape(k, int(g["V"])).T.astype(np.float64) @ C.astype(np.float64).T
    # the same product on the tensor cores (3xTF32 through three TF32 GEMMs, used for large products only by default):
    # fp32-level agreement with the float64 product
    try:
        ctx.set_option("dense_tc_min_mflops", 0)
        lifted_tc = B.left_multiply_by_U_Spectra(C, k, k)
    finally:
        ctx.set_option("dense_tc_min_mflops", 20000)
    scale = np.max(np.abs(lift_ref))
    assert np.max(np.abs(lifted_tc - lift_ref)) <= 2e-6 * scale and np.max(np.abs(lifted - lift_ref)) <= 2e-6 * scale
    ctx.set_option("dist_kernel", 1)


def test_lloyd_c1_matches_reference(ctx, golden_c1, corpus_c1):
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, *_ = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    B.set_U(g["U_colmajor"].reshape(c.k, c.V).T.copy())
    C = g["centers_lowd_init"].reshape(c.k, c.k).copy()
    B.run_lloyds_on_projected_space(c.k, C, None, 10)
    mism = int((B.last_lloyd["assign"] != g["lloyd_assign"]).sum())
    assert mism <= 2, mism                                           # ties excepted
    assert np.max(np.abs(C - g["centers_lowd_final"].reshape(c.k, c.k))) < 1e-3


def test_empty_cluster_center_stays_zero(ctx, golden_tiny):
    """src/sparseMatrix.cpp:1988-1992."""
    g = golden_tiny
    B, P_ref, k = setup_projection(ctx, g)
    C = g["centers_lowd_init"].reshape(k, k).copy()
    C[3] = 1e6                                           # nobody is closest to this center
    B.run_lloyds_on_projected_space(k, C, None, 1)
    assert np.all(C[3] == 0.0)
    assert not np.any(B.last_lloyd["assign"] == 3)


def test_kmeanspp_invariants(ctx, golden_tiny):
    g = golden_tiny
    B, P_ref, k = setup_projection(ctx, g)
    seeds, coords, res = B.kmeans_init_on_projected_space(k, 1, seed=11)
    assert len(set(seeds.tolist())) == k                 # no replicated center (:2176-2178)
    assert np.max(np.abs(coords - P_ref[seeds.astype(np.int64)])) <= 1e-5 * np.max(np.abs(P_ref))
    assert res > 0
    # D^2 seeding must beat uniform seeding on the k-means++ potential, on average
    d2 = O.docs_l2sq(P_ref)
    pot = lambda S: np.maximum(O.dist_matrix(P_ref, d2, P_ref[S]), 0).min(1).sum()
    rng = np.random.default_rng(0)
    uni = np.mean([pot(rng.choice(len(P_ref), k, replace=False)) for _ in range(8)])
    assert pot(seeds.astype(np.int64)) < uni


# ---------------------------------------------------------------- stage F: Lloyd on the full-dimensional B
def _stage_f_B(ctx, g, c):
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A, B, zetas, nn, oc = make_AB(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz)
    assert np.array_equal(zetas, g["zetas"])
    bv, br, bo, _ = O.threshold_and_copy(vals, c.rows, c.offsets, g["zetas"])
    return B, O.to_csc(bv, br, bo, c.V)


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_lloyd_full_matches_reference(ctx, name, golden_tiny_stageF, golden_c1_stageF, corpus_c1):
    """SURVEY 8(f) row 1 (run_lloyds, src/sparseMatrix.cpp:1679-1746) through the C ABI against the
    reference's own output for the same B and the same starting centers."""
    from isle_b200 import corpus
    g = golden_tiny_stageF if name == "tiny" else golden_c1_stageF
    c = corpus.generate("tiny") if name == "tiny" else corpus_c1
    B, Bo = _stage_f_B(ctx, g, c)
    C = np.ascontiguousarray(g["centers_in"].reshape(c.k, c.V).copy())
    closest = [[] for _ in range(c.k)]
    assert B.run_lloyds(c.k, C, closest, 10) == 0.0
    a = B.last_lloyd_full["assign"]
    C_ref = g["centers_out"].reshape(c.k, c.V)
    assert np.mean(a != g["assign"]) <= 1e-3                      # identical partition, near-ties excepted
    assert np.max(np.abs(C - C_ref)) <= 1e-5 * np.max(np.abs(C_ref))
    o_ref = O.kmeans_objective_full(Bo, C_ref, g["assign"])
    assert abs(B.last_lloyd_full["objective"] - o_ref) <= 1e-5 * o_ref
    assert abs(O.kmeans_objective_full(Bo, C, a) - B.last_lloyd_full["objective"]) <= 1e-5 * o_ref
    # closest_docs: a partition of all documents, ascending ids per cluster (trainer.cpp:567-570)
    assert sum(len(x) for x in closest) == B.num_docs()
    assert all(x == sorted(x) for x in closest)
    assert all(a[d] == t for t in range(c.k) for d in closest[t][:3])
    # oracle restatement from the same start agrees as well, iteration count included
    C_o, a_o, it_o = O.run_lloyds_full(Bo, g["centers_in"].reshape(c.k, c.V), 10)
    assert np.mean(a != a_o) <= 1e-3 and B.last_lloyd_full["iters"] == it_o


def test_lloyd_full_single_iteration_and_empty_cluster(ctx, golden_tiny_stageF):
    """One lloyds_iter (src/sparseMatrix.cpp:1584-1667) against the oracle; a center nobody is
    closest to is left at zero (:1626, :1655-1661); max_reps is honoured."""
    from isle_b200 import corpus
    g, c = golden_tiny_stageF, corpus.generate("tiny")
    B, Bo = _stage_f_B(ctx, g, c)
    C0 = g["centers_in"].reshape(c.k, c.V).copy()
    C0[3] = 1e3                                                   # far from every document
    C = np.ascontiguousarray(C0.copy())
    B.run_lloyds(c.k, C, None, 1)
    assert B.last_lloyd_full["iters"] == 1
    C_o, a_o = O.lloyds_iter_full(Bo, O.docs_l2sq_full(Bo), C0)
    assert np.array_equal(B.last_lloyd_full["assign"], a_o)
    assert not np.any(a_o == 3) and np.all(C[3] == 0.0)
    assert np.max(np.abs(C - C_o)) <= 1e-5 * np.max(np.abs(C_o))


def test_lloyd_full_from_device_resident_lifted_centers(ctx, golden_tiny):
    """train() order (src/trainer.cpp:550-566): lift the projected centers, clean up the eigensolver,
    then run_lloyds on B.  With centers=None the library starts from the lifted centers it kept on the
    device; the result must equal the host round trip."""
    g = golden_tiny
    A, B, zetas, nn, oc = tiny_AB(ctx, g)
    k, V = int(g["k"]), int(g["V"])
    B.set_U(g["U_colmajor"].reshape(k, V).T)
    Cl = np.ascontiguousarray(g["centers_lowd_final"].reshape(k, k))
    lifted = B.left_multiply_by_U_Spectra(Cl, k, k)              # V x k, also kept on the device
    B.cleanup_after_eigensolver()
    B.run_lloyds(k, None, None, 10)
    dev = dict(B.last_lloyd_full)
    C = np.ascontiguousarray(lifted.T.copy())
    B.run_lloyds(k, C, None, 10)
    assert np.array_equal(dev["assign"], B.last_lloyd_full["assign"])
    assert dev["iters"] == B.last_lloyd_full["iters"]
    assert abs(dev["objective"] - B.last_lloyd_full["objective"]) <= 1e-12 * dev["objective"]   # fp64 atomics: order varies


# ---------------------------------------------------------------- stage G: catchword thresholds, catchwords
def _upload_A(ctx, c):
    from isle_b200.sparse_matrix import SparseMatrix
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A = SparseMatrix(c.V, c.D, ctx)
    A.populate_normalized(vals, c.rows, c.offsets, avg, nz)
    return A, vals


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_catchwords_match_reference(ctx, name, golden_tiny_stageG, golden_c1_stageG, corpus_c1):
    """SURVEY 8(f) row 2 through the C ABI, both calling conventions, bit-exact against the reference's own
    thresholds and catchwords for the same partition."""
    from isle_b200 import corpus
    g = golden_tiny_stageG if name == "tiny" else golden_c1_stageG
    c = corpus.generate("tiny") if name == "tiny" else corpus_c1
    A, vals = _upload_A(ctx, c)
    r, cl = int(g["r"]), g["cluster_of_doc"]
    thr_ref = g["thresholds"].reshape(c.k, c.V)
    thr = A.catchword_thresholds(c.k, r, cl)                      # all clusters in one pass
    assert np.array_equal(thr.view(np.uint32), thr_ref.view(np.uint32))
    for t in (0, c.k // 2, c.k - 1):                              # the reference's per-topic convention
        one = A.rth_highest_element(r, np.nonzero(cl == t)[0])
        assert np.array_equal(one.view(np.uint32), thr_ref[t].view(np.uint32))
    for source in (thr, None):                                    # host matrix, then the device-resident copy
        cw = A.find_catchwords(c.k, source)
        pairs = np.array([(t, w) for t in range(c.k) for w in cw[t]], dtype=np.uint32).reshape(-1, 2)
        assert np.array_equal(pairs, g["catchwords"])


def test_rth_highest_element_edges(ctx):
    """Branches of src/sparseMatrix.cpp:496-520 against the oracle: empty cluster, count == r, r >= cluster size,
    documents outside every cluster, a cluster id with no documents."""
    from isle_b200 import corpus
    c = corpus.generate("tiny")
    A, vals = _upload_A(ctx, c)
    rng = np.random.default_rng(9)
    assert np.all(A.rth_highest_element(3, []) == 0)
    for docs, r in ((rng.choice(c.D, 40, replace=False), 5), (np.arange(3), 3), (np.arange(7), 50), (np.array([11]), 1)):
        ref = O.rth_highest_element(vals, c.rows, c.offsets, c.V, np.sort(docs), r)
        assert np.array_equal(A.rth_highest_element(r, np.sort(docs)).view(np.uint32), ref.view(np.uint32))
    cl = rng.integers(0, 6, c.D).astype(np.uint32)
    cl[cl == 5] = 0xFFFFFFFF                                      # in no cluster
    cl[cl == 4] = 0                                               # cluster 4 stays empty
    thr = A.catchword_thresholds(5, 20, cl)
    ref = np.stack([O.rth_highest_element(vals, c.rows, c.offsets, c.V, np.nonzero(cl == t)[0], 20) for t in range(5)])
    assert np.array_equal(thr.view(np.uint32), ref.view(np.uint32)) and np.all(thr[4] == 0)
    cw, cw_ref = A.find_catchwords(5, thr), O.find_catchwords(ref)
    assert all(np.array_equal(a, b) for a, b in zip(cw, cw_ref))
    assert all(len(x) == 0 for x in A.find_catchwords(1, thr[:1]))   # k = 1: the reference's loop never sets the flag


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_topic_model_matches_reference(ctx, name, golden_tiny_stageH, golden_c1_stageH, corpus_c1):
    """SURVEY 8(f) row 2, second half (construct_topic_model, src/sparseMatrix.cpp:597-838) through the C ABI against
    the reference's own output: (doc, topic, sum) list and top topic pairs bit-exact, model within 1e-6."""
    from isle_b200 import corpus
    g = golden_tiny_stageH if name == "tiny" else golden_c1_stageH
    c = corpus.generate("tiny") if name == "tiny" else corpus_c1
    A, vals = _upload_A(ctx, c)
    cw = [g["catchwords"][g["catchwords"][:, 0] == t, 1].astype(np.int64) for t in range(c.k)]
    M, (dd, dt, dv), pairs = A.construct_topic_model(c.k, g["cluster_of_doc"], cw)
    assert np.array_equal(dd, g["dts_doc"]) and np.array_equal(dt, g["dts_topic"])
    assert np.array_equal(dv.view(np.uint32), g["dts_val"].view(np.uint32))
    assert np.array_equal(pairs, g["top_topic_pairs"].astype(np.int64))
    M_ref = g["model"].reshape(c.k, c.V).T
    assert np.max(np.abs(M - M_ref)) <= 1e-6 * np.max(np.abs(M_ref))
    assert np.allclose(M.sum(0), 1.0, atol=1e-5)


def test_topic_model_corner_cases(ctx):
    """Topics without catchwords, documents in no cluster, an empty cluster and a topic nobody contributes to (its
    column divides by zero, as the reference's FPscal(1 / asum) does), against the oracle."""
    from isle_b200 import corpus
    c = corpus.generate("tiny")
    A, vals = _upload_A(ctx, c)
    rng = np.random.default_rng(4)
    k = 6
    cl = rng.integers(0, 5, c.D).astype(np.uint32)               # cluster 5 stays empty
    cl[rng.random(c.D) < 0.1] = 0xFFFFFFFF
    words = rng.permutation(c.V)
    cw = [np.sort(words[:40]), np.sort(words[40:45]), np.zeros(0, np.int64), np.sort(words[45:300]), np.zeros(0, np.int64),
          np.zeros(0, np.int64)]
    M, (dd, dt, dv), pairs = A.construct_topic_model(k, cl, cw)
    M_o, (dd_o, dt_o, dv_o), pairs_o, _ = O.construct_topic_model(vals, c.rows, c.offsets, c.V, k, cl, cw)
    assert np.array_equal(dd, dd_o) and np.array_equal(dt, dt_o) and np.array_equal(dv.view(np.uint32), dv_o.view(np.uint32))
    assert np.array_equal(pairs, pairs_o)
    assert np.all(np.isnan(M[:, 5])) and np.all(np.isnan(M_o[:, 5]))      # 0 * (1 / 0)
    assert np.max(np.abs(M[:, :5] - M_o[:, :5])) <= 1e-6 * np.max(np.abs(M_o[:, :5]))


# ---------------------------------------------------------------- ingest
def _corpus_text(c, shuffle_seed=None):
    docs = np.repeat(np.arange(c.D, dtype=np.int64), np.diff(c.offsets))
    arr = np.stack([docs + 1, c.rows.astype(np.int64) + 1, c.counts.astype(np.int64)], 1)
    if shuffle_seed is not None:
        arr = arr[np.random.default_rng(shuffle_seed).permutation(len(arr))]
    import io
    buf = io.BytesIO()
    np.savetxt(buf, arr, fmt="%d")
    return buf.getvalue()


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_ingest_text_matches_reference(ctx, name, golden_tiny, golden_c1, corpus_c1):
    """Text -> CSC -> normalised A on the device against the reference's own ingest output (ref_dump stage 0), from the
    shuffled text form of the corpus: offsets, rows, normalised values bit-exact; then thresholds straight from it."""
    from isle_b200 import corpus
    from isle_b200.sparse_matrix import SparseMatrix
    g = golden_tiny if name == "tiny" else golden_c1
    c = corpus.generate("tiny") if name == "tiny" else corpus_c1
    A = SparseMatrix(c.V, c.D, ctx)
    A.ingest_text(_corpus_text(c, shuffle_seed=5), max_entries=c.nnz)
    vals, rows, offs = A.download()
    assert A.get_nnzs() == c.nnz and A.avg_doc_sz == float(g["avg_doc_sz"]) and A._nz_docs == c.D
    assert np.array_equal(offs, c.offsets) and np.array_equal(rows, c.rows.astype(np.uint64))
    assert sha(vals) == str(g["A_vals_sha"])
    zetas, nn = A.compute_thresholds(0, c.V, None, c.k)
    assert np.array_equal(zetas, g["zetas"]) and nn == int(g["new_nnzs"])


def test_ingest_text_edge_cases_and_errors(ctx):
    from isle_b200._capi import IsleCudaError
    from isle_b200.sparse_matrix import SparseMatrix
    text = b"2 3 4\r\n1\t\t5   6\n2 3 9\n4 1 1\n5 5 7\n5 2 1"     # CRLF, tabs, duplicate (doc, word), empty doc 3, no final newline
    d, w, n = O.parse_entries(text)
    offsets, rows, counts = O.entries_to_csc(d, w, n, 6)
    vals_o, avg_o, nz_o = O.normalize_docs(counts, offsets)
    A = SparseMatrix(5, 6, ctx)
    A.ingest_text(text, max_entries=6)
    vals, r, offs = A.download()
    assert np.array_equal(offs, offsets) and np.array_equal(r, rows.astype(np.uint64))
    assert np.array_equal(vals.view(np.uint32), vals_o.view(np.uint32)) and A.avg_doc_sz == float(avg_o) and A._nz_docs == nz_o
    for bad in (b"1 2 x\n", b"1 2\n", b"7 1 1\n", b"1 6 1\n", b"0 1 1\n", b"1 1 1 1\n"):
        with pytest.raises(IsleCudaError):
            SparseMatrix(5, 6, ctx).ingest_text(bad)
    with pytest.raises(IsleCudaError):                              # the reference asserts nRead == max_entries
        SparseMatrix(5, 6, ctx).ingest_text(text, max_entries=7)
    with pytest.raises(IsleCudaError):                              # 2^24 tokens in one document: fp32 doc_sum no longer exact
        SparseMatrix(5, 6, ctx).ingest_text(b"1 1 16777216\n1 2 5\n")


def test_populate_csc_and_normalize_on_device(ctx, golden_c1, corpus_c1):
    """a2 of SURVEY 8: populate_CSC's statistics + normalize_docs on the device from the sorted CSC of raw counts."""
    from isle_b200.sparse_matrix import SparseMatrix
    g, c = golden_c1, corpus_c1
    A = SparseMatrix(c.V, c.D, ctx)
    A.populate_CSC_and_normalize(c.counts, c.rows, c.offsets)
    vals, rows, offs = A.download()
    vals_o, avg_o, nz_o = O.normalize_docs(c.counts, c.offsets)
    assert np.array_equal(vals.view(np.uint32), vals_o.view(np.uint32)) and sha(vals) == str(g["A_vals_sha"])
    assert A.avg_doc_sz == float(avg_o) and A._nz_docs == nz_o


# ---------------------------------------------------------------- end to end
def test_spectral_core_end_to_end_c1(ctx, golden_c1, corpus_c1):
    """Stages A-E through the public call; k-means is checked through rotation-invariant
    quantities because our U spans the same subspace as the reference's but is not identical."""
    from isle_b200.trainer import spectral_core
    g, c = golden_c1, corpus_c1
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    r = spectral_core(ctx, c.V, c.D, c.k, vals, c.rows, c.offsets, avg, nz, seed=4, want_U=True)
    assert np.array_equal(r.zetas, g["zetas"]) and r.nnz_B == int(g["nnz_B"]) and r.D_B == int(g["D_B"])
    U_ref = g["U_colmajor"].reshape(c.k, c.V).T
    s, s_ref = np.sqrt(r.evalues), np.sqrt(g["evalues"])
    assert np.max(np.abs(s - s_ref) / s_ref) < 1e-4
    assert O.principal_angle_sin(r.U, U_ref) < 1e-3
    # oracle Lloyd from OUR seeds on the oracle's projection (distances are rotation invariant)
    bv, br, bo, _ = O.threshold_and_copy(vals, c.rows, c.offsets, g["zetas"])
    P_ref = O.project(O.to_csc(bv, br, bo, c.V), U_ref)
    C_ref, a_ref, _ = O.run_lloyds(P_ref, P_ref[r.seeds.astype(np.int64)], 10)
    obj_ref = O.kmeans_objective(P_ref, C_ref, a_ref)
    assert abs(r.objective - obj_ref) / obj_ref < 2e-3
    assert r.centers.shape == (c.V, c.k)
    assert ctx.stat("launches") > 0


@pytest.mark.gpu
def test_document_sharded_run_matches_single_gpu():
    """N = 2 ranks over NCCL vs one GPU (tests/multi_gpu_check.py); needs two visible GPUs."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "multi_gpu_check.py"), "c1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
