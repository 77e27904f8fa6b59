"""Full-size (BASELINE.json configs[1]: 300k docs x 102k vocab, ~70M nnz, k = 100) checks of the CUDA
path through the C ABI.  Thresholds and B are still compared bit for bit with the oracle (the numpy
restatement finishes this size in seconds); the eigensolver and k-means stages, which the oracle
cannot finish at this size, are checked through size-independent properties:

  * operator: symmetry  <x, A y> = <A x, y>,  linearity, positivity  <x, A x> = ||B^T x||^2 >= 0
  * eigensolver: nconv = k, sigma descending and positive, sum sigma^2 <= ||B||_F^2
    (src/trainer.cpp:490 logs the Frobenius norm for exactly this comparison), U^T U = I,
    residual ||A u - sigma^2 u|| <= a few tol * sigma^2 for every Ritz pair
  * projection: ||P_d||^2 returned by the library = row norms of P; P = B^T U checked on sampled docs
  * k-means++: k distinct seeds; Lloyd: objective non-increasing from iteration to iteration,
    every document assigned to its nearest center (sampled fp64 check), idempotent at the fixed point
"""
import numpy as np
import pytest

from oracle import isle_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2_state(ctx):
    import torch

    from isle_b200 import corpus
    from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix
    cfg = corpus.CONFIGS["c2"]
    c = corpus.generate(V=cfg["V"], D=cfg["D"], k=cfg["k"], mu=cfg["mu"], seed=cfg["seed"], backend="torch",
                        device="cuda:0")
    torch.cuda.empty_cache()
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    A = SparseMatrix(c.V, c.D, ctx)
    A.populate_normalized(vals, c.rows, c.offsets, avg, nz)
    zetas, nn = A.compute_thresholds(0, c.V, A.list_word_freqs_by_sorting(), c.k)
    B = FPSparseMatrix(A)
    oc = B.threshold_and_copy(A, zetas, nn)
    return dict(c=c, vals=vals, avg=avg, nz=nz, A=A, B=B, zetas=zetas, nn=nn, oc=oc)


def test_c2_thresholds_and_B_bit_exact(ctx, c2_state):
    s = c2_state
    c = s["c"]
    z_ref, nn_ref = O.compute_thresholds(s["vals"], c.rows, c.V, s["nz"], c.k)
    assert np.array_equal(s["zetas"], z_ref)
    assert s["nn"] == nn_ref
    bv, br, bo, oc = O.threshold_and_copy(s["vals"], c.rows, c.offsets, z_ref)
    vals, rows, offs, oc_dev = s["B"].download()
    assert int(offs[-1]) == nn_ref == len(vals)          # kept == nnz_B (src/sparseMatrix.cpp:1312-1318)
    assert np.array_equal(offs, bo)
    assert np.array_equal(rows, br.astype(np.uint64))
    assert np.array_equal(vals.view(np.uint32), bv.view(np.uint32))
    assert np.array_equal(oc_dev, oc.astype(np.uint64)) and np.array_equal(s["oc"], oc_dev)
    # idempotence: B's values are sqrt(zeta_w) >= 1 per row, thresholding B's pattern again keeps all of it
    assert float(vals.min()) >= 1.0


def test_c2_operator_properties(ctx, c2_state):
    B = c2_state["B"]
    V = c2_state["c"].V
    rng = np.random.default_rng(5)
    X = rng.standard_normal((V, 10)).astype(np.float32)
    Y = rng.standard_normal((V, 10)).astype(np.float32)
    AX, AY = B.multiply(X), B.multiply(Y)
    # symmetry, per column pair, relative to the magnitudes involved
    lhs = np.einsum("ij,ij->j", X.astype(np.float64), AY.astype(np.float64))
    rhs = np.einsum("ij,ij->j", AX.astype(np.float64), Y.astype(np.float64))
    scale = np.linalg.norm(X, axis=0) * np.linalg.norm(AY, axis=0)
    assert np.max(np.abs(lhs - rhs) / scale) < 1e-5
    # linearity
    AZ = B.multiply((2.0 * X - 3.0 * Y).astype(np.float32))
    ref = 2.0 * AX.astype(np.float64) - 3.0 * AY.astype(np.float64)
    assert np.max(np.linalg.norm(AZ - ref, axis=0) / np.linalg.norm(ref, axis=0)) < 1e-5
    # positive semi-definite
    assert np.all(np.einsum("ij,ij->j", X.astype(np.float64), AX.astype(np.float64)) > 0)
    # the engines agree: tensor-core head + block-FP tail vs plain fp32 gathers
    ctx.set_option("spmm_head", 0)
    ctx.set_option("spmm_bfp", 0)
    try:
        AX0 = B.multiply(X)
    finally:
        ctx.set_option("spmm_head", 1)
        ctx.set_option("spmm_bfp", 1)
        assert np.max(np.linalg.norm(AX - AX0, axis=0) / np.linalg.norm(AX0, axis=0)) < 2e-6


def test_c2_eigensolver_and_kmeans_properties(ctx, c2_state):
    s = c2_state
    c, B = s["c"], s["B"]
    k, tol = c.k, 1e-4
    B.initialize_for_eigensolver(k)
    ev, U = B.compute_block_ks(k, seed=11, want_U=True)
    assert B.nconv == k
    assert np.all(ev > 0) and np.all(np.diff(ev) <= 0)
    fro = B.frobenius()                                  # sum of squares (src/sparseMatrix.cpp:1096-1100)
    assert float(ev.astype(np.float64).sum()) <= fro * (1 + 1e-5)
    G = U.astype(np.float64).T @ U.astype(np.float64)
    assert np.max(np.abs(G - np.eye(k))) < 5e-5
    # Ritz residuals through the operator, 10 columns per call
    worst = 0.0
    for j0 in range(0, k, 10):
        Uj = np.ascontiguousarray(U[:, j0:j0 + 10])
        R = B.multiply(Uj).astype(np.float64) - Uj.astype(np.float64) * ev[j0:j0 + 10].astype(np.float64)
        worst = max(worst, float(np.max(np.linalg.norm(R, axis=0) / ev[j0:j0 + 10])))
    assert worst < 5 * tol, worst

    # projection
    P, l2 = B.projected_docs()
    assert np.allclose(l2, np.einsum("ij,ij->i", P.astype(np.float64), P.astype(np.float64)), rtol=2e-5)
    vals, rows, offs, _ = B.download()
    rng = np.random.default_rng(3)
    for d in rng.integers(0, B.num_docs(), 64):
        sl = slice(int(offs[d]), int(offs[d + 1]))
        ref = (vals[sl].astype(np.float64)[:, None] * U[rows[sl].astype(np.int64)].astype(np.float64)).sum(0)
        assert np.max(np.abs(P[d] - ref)) <= 2e-5 * max(1.0, float(np.abs(ref).max()))

    # k-means++ and Lloyd
    seeds, centers, _ = B.kmeans_init_on_projected_space(k, 1, seed=11)
    assert len(np.unique(seeds)) == k and int(seeds.max()) < B.num_docs()
    assert np.allclose(centers, P[seeds.astype(np.int64)], atol=1e-5)

    def objective(C, a):
        diff = P.astype(np.float64) - C.astype(np.float64)[a.astype(np.int64)]
        return float(np.einsum("ij,ij->", diff, diff))

    prev, Ccur = None, centers.copy()
    for it in range(4):
        a = B.projected_closest_centers(k, Ccur)
        obj = objective(Ccur, a)
        if prev is not None:
            assert obj <= prev * (1 + 1e-6), (it, obj, prev)
        prev = obj
        # sampled nearest-center check in fp64, F7 semantics (argmin of |dist|), near-ties excepted
        idx = rng.integers(0, B.num_docs(), 2000)
        d2 = ((P[idx].astype(np.float64)[:, None, :] - Ccur.astype(np.float64)[None]) ** 2).sum(-1)
        best = d2.min(1)
        mine = d2[np.arange(len(idx)), a[idx].astype(np.int64)]
        assert np.all(mine <= best + 1e-4 * np.maximum(best, 1.0))
        B.run_lloyds_on_projected_space(k, Ccur, None, 1)      # one Lloyd iteration: centers <- means
        assert objective(Ccur, a) <= obj * (1 + 1e-6)            # the mean minimises the objective of a partition
    # run to the reference's stopping rule; at the fixed point one more iteration changes nothing
    B.run_lloyds_on_projected_space(k, Ccur, None, 10)
    a1 = B.last_lloyd["assign"].copy()
    if B.last_lloyd["iters"] < 10:
        C2 = Ccur.copy()
        B.run_lloyds_on_projected_space(k, C2, None, 1)
        assert np.array_equal(B.last_lloyd["assign"], a1)
    assert np.bincount(a1, minlength=k).sum() == B.num_docs()   # partition covers all docs (trainer.cpp:570)
    B.cleanup_after_eigensolver()
