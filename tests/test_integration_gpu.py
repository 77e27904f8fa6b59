"""Drop-in check of the reference-side binding: the reference's own CLI and
trainer.cpp, unchanged, linked against isle_b200/shim/sparseMatrix_cuda.cpp + libisle_cuda.so
(oracle/_ref/ISLETrain_cuda) must reproduce what the all-CPU reference build
(oracle/_ref/ISLETrain_ref) logs for the spectral core on the same input files:
entries above threshold and ||B||_F^2 exactly, singular values within 1e-4 relative, and a
well-formed topic model at the end of the untouched host stages F-H."""
import os
import re
import subprocess

import numpy as np
import pytest

from isle_b200 import corpus

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "ISLETrain_ref")
CUDA = os.path.join(ROOT, "oracle", "_ref", "ISLETrain_cuda")


def run_cli(exe, wd, c, name, env=None):
    out = os.path.join(wd, name)
    os.makedirs(out)
    # 12 positional arguments (reference drivers/ISLETrain.cpp:9-32)
    args = [exe, os.path.join(wd, "tdf.txt"), os.path.join(wd, "vocab.txt"), out, str(c.V), str(c.D), str(c.nnz),
            str(c.k), "0", "0", "0", "0", "0"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    logdir = os.path.join(out, os.listdir(out)[0])
    log = open(os.path.join(logdir, "diagnosticLog.txt")).read()
    nnz = int(re.search(r"Number of entries above threshold: (\d+)", log).group(1))
    frob = float(re.search(r"Frob\(B_fl_CSC\): ([0-9.eE+-]+)", log).group(1))
    eig = np.array([float(x) for x in re.findall(r"\(\d+\): ([0-9.eE+-]+)", log.split("Eigvals:")[1].split("\n")[0])])
    model = np.loadtxt(os.path.join(logdir, "M_hat_catch_sparse"))
    return nnz, frob, eig, model


@pytest.mark.gpu
def test_isletrain_cli_with_cuda_spectral_core(tmp_path):
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("oracle/_ref binaries are built by __graft_entry__.build() where /root/reference is mounted")
    c = corpus.generate("tiny")
    c.write_text(str(tmp_path / "tdf.txt"), str(tmp_path / "vocab.txt"))
    nnz_r, frob_r, eig_r, model_r = run_cli(REF, str(tmp_path), c, "ref")
    nnz_c, frob_c, eig_c, model_c = run_cli(CUDA, str(tmp_path), c, "cuda")
    assert nnz_c == nnz_r
    assert frob_c == frob_r
    assert eig_c.shape == eig_r.shape == (c.k,)
    assert np.max(np.abs(eig_c - eig_r) / eig_r) < 1e-4
    # host stages F-H ran on the arrays the library filled: a V x k column-stochastic model
    # (M_hat_catch_sparse lines are `<topic> <word> <prob>`, 1-based, small entries dropped)
    M = np.zeros((c.V, c.k))
    M[model_c[:, 1].astype(int) - 1, model_c[:, 0].astype(int) - 1] = model_c[:, 2]
    sums = M.sum(0)
    assert np.all((np.abs(sums - 1.0) < 1e-2) | (sums == 0.0))
    assert (sums > 0).sum() >= (np.unique(model_r[:, 0]).size * 3) // 4


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_isletrain_cli_on_two_gpus(tmp_path):
    """The unmodified CLI / trainer.cpp on a multi-GPU context (ISLE_CUDA_NGPUS=2 -> isle_cuda_create_multi: the library
    shards the documents itself, one host thread per GPU): same log lines as the single-GPU drop-in."""
    if not os.path.exists(CUDA):
        pytest.skip("oracle/_ref binaries are built by __graft_entry__.build() where /root/reference is mounted")
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    c = corpus.generate("c1")
    c.write_text(str(tmp_path / "tdf.txt"), str(tmp_path / "vocab.txt"))
    nnz_1, frob_1, eig_1, model_1 = run_cli(CUDA, str(tmp_path), c, "one")
    nnz_2, frob_2, eig_2, model_2 = run_cli(CUDA, str(tmp_path), c, "two", env={"ISLE_CUDA_NGPUS": "2"})
    assert nnz_2 == nnz_1 and abs(frob_2 - frob_1) <= 1e-6 * frob_1
    assert np.max(np.abs(eig_2 - eig_1) / eig_1) < 1e-4
    assert model_2.shape[1] == 3 and abs(len(model_2) - len(model_1)) <= 0.05 * len(model_1)


@pytest.mark.gpu
def test_multi_gpu_context_matches_single_gpu_through_the_c_abi():
    """isle_cuda_create_multi behind the same host mirror: thresholds, B (stitched in document order), original_cols
    bit-exact; singular values / subspace / Lloyd from identical inputs within the sharded-run tolerances; stage F
    (integer reduction) identical; catchword thresholds identical."""
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    from isle_b200 import _capi
    from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix
    from oracle import isle_oracle as O
    c = corpus.generate("c1")
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)

    def run(ctx, U_in=None, C_in=None, F_in=None):
        A = SparseMatrix(c.V, c.D, ctx)
        A.populate_normalized(vals, c.rows.astype(np.uint64), c.offsets, avg, nz)
        z, nn = A.compute_thresholds(0, c.V, None, c.k)
        B = FPSparseMatrix(A)
        oc = B.threshold_and_copy(A, z, nn)
        bv, br, bo, oc2 = B.download()
        fro = B.frobenius()
        ev, U = B.compute_block_ks(c.k, seed=3, want_U=True)
        if U_in is not None:
            B.set_U(U_in)
        P, l2 = B.projected_docs()
        seeds, coords, _ = B.kmeans_init_on_projected_space(c.k, 1, seed=3)
        C0 = coords.copy() if C_in is None else C_in.copy()
        Cl = C0.copy()
        B.run_lloyds_on_projected_space(c.k, Cl, None, 10)
        lifted = np.ascontiguousarray(B.left_multiply_by_U_Spectra(Cl, c.k, c.k).T) if F_in is None else F_in.copy()
        full_in = lifted.copy()
        B.cleanup_after_eigensolver()
        B.run_lloyds(c.k, lifted, None, 10)
        cl = np.full(c.D, 0xFFFFFFFF, np.uint32)
        cl[oc.astype(np.int64)] = B.last_lloyd_full["assign"]
        thr = A.catchword_thresholds(c.k, O.catchword_rank(c.D, c.k), cl)
        cw = A.find_catchwords(c.k, None)
        model, dts, _ = A.construct_topic_model(c.k, cl, cw, want_pairs=False)
        return dict(z=z, nn=nn, oc=oc, bv=bv, br=br, bo=bo, oc2=oc2, fro=fro, ev=ev, U=U, P=P, l2=l2, C0=C0, Cl=Cl,
                    assign=B.last_lloyd["assign"], obj=B.last_lloyd["objective"], full_in=full_in, full=lifted,
                    full_assign=B.last_lloyd_full["assign"], thr=thr, cw=cw, model=model, dts=dts, seeds=seeds)

    def sample(ctx):      # SURVEY 8(f) row 4 under sharding: same keys (global document numbers), corpus-wide pivot
        from isle_b200._capi import ptr
        import ctypes as C
        A = SparseMatrix(c.V, c.D, ctx)
        A.populate_normalized(vals, c.rows.astype(np.uint64), c.offsets, avg, nz)
        A.compute_thresholds(0, c.V, None, c.k)
        out = {}
        for rate in (0.1, 0.37, 2.0):
            sel, n = np.zeros(c.D, np.uint8), C.c_uint64()
            ctx.call("isle_cuda_sample_docs", C.c_float(rate), 11, ptr(sel), C.byref(n))
            assert int(sel.sum()) == n.value
            out[rate] = sel
        return out

    two = _capi.Context(n_gpus=2)
    sel_two = sample(two)
    one_ = _capi.Context(0)
    sel_one = sample(one_)
    one_.close()
    for rate in sel_one:
        assert np.array_equal(sel_two[rate], sel_one[rate]), rate
    assert sel_one[2.0].all() and 0 < sel_one[0.1].sum() < 0.2 * c.D
    # the library's own collectives over peer memory (in one process: plain peer access) against NCCL, bit for bit
    import ctypes as C
    mism, active = C.c_uint64(), C.c_int()
    two.call("isle_cuda_selftest_collectives", C.byref(mism), C.byref(active))
    assert active.value == 1 and mism.value == 0
    r = run(two)
    assert two.stat("n_gpus") == 2 and two.stat("D_B") == len(r["oc"])
    two.close()
    one = _capi.Context(0)
    s = run(one, U_in=r["U"], C_in=r["C0"], F_in=r["full_in"])
    one.close()
    assert np.array_equal(r["z"], s["z"]) and r["nn"] == s["nn"]
    for key in ("oc", "oc2", "bv", "br", "bo"):
        assert np.array_equal(r[key], s[key]), key
    assert abs(r["fro"] - s["fro"]) <= 1e-6 * s["fro"]
    sv_r, sv_s = np.sqrt(r["ev"]), np.sqrt(s["ev"])
    assert np.max(np.abs(sv_r - sv_s) / sv_s) < 1e-4
    assert np.max(np.abs(r["P"] - s["P"])) <= 1e-5 * np.max(np.abs(s["P"]))         # same U on both sides, rows in document order
    assert len(set(r["seeds"].tolist())) == c.k
    assert abs(r["obj"] - s["obj"]) <= 1e-4 * s["obj"] and np.mean(r["assign"] != s["assign"]) < 5e-3
    assert np.array_equal(r["full"], s["full"]) and np.array_equal(r["full_assign"], s["full_assign"])
    assert np.array_equal(r["thr"].view(np.uint32), s["thr"].view(np.uint32))
    assert all(np.array_equal(a, b) for a, b in zip(r["cw"], s["cw"]))
    for a, b in zip(r["dts"], s["dts"]):
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b)
    ok = ~np.isnan(s["model"])
    assert np.max(np.abs(r["model"][ok] - s["model"][ok])) <= 1e-6 * np.max(np.abs(s["model"][ok]))
