"""Operator check on a B200: B (B^T X) through the C ABI with the dense-head tensor-core engine on
and off, against a float64 scipy product of the downloaded B; then per-kernel timings.

    python tools/spmm_check.py [--config c2] [--docs 300000] [--b 10] [--reps 20]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from isle_b200 import _capi, corpus  # noqa: E402
from isle_b200._capi import ptr  # noqa: E402
from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix  # noqa: E402
from oracle import isle_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--b", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--density-ppm", type=int, nargs="*", default=[12000])
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--seg", type=int, nargs="*", default=[16])
    ap.add_argument("--only-head-serial", action="store_true")
    ap.add_argument("--i8", type=int, nargs="*", default=[1])
    ap.add_argument("--head-max", type=int, default=4096)
    ap.add_argument("--opt", nargs="*", default=[], help="extra name=value options")
    a = ap.parse_args()

    import torch
    cfg = dict(corpus.CONFIGS[a.config])
    D = a.docs or cfg["D"]
    backend = "torch" if D * 50 > 2_000_000 else "numpy"
    c = corpus.generate(V=cfg["V"], D=D, k=cfg["k"], mu=cfg["mu"], seed=cfg["seed"], backend=backend,
                        device="cuda:0" if backend == "torch" else "cpu")
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    ctx = _capi.Context(0)
    A = SparseMatrix(c.V, c.D, ctx)
    A.populate_normalized(vals, c.rows, c.offsets, avg, nz)
    zetas, nn = A.compute_thresholds(0, c.V, A.list_word_freqs_by_sorting(), c.k)
    B = FPSparseMatrix(A)
    B.threshold_and_copy(A, zetas, nn)
    print(f"V={c.V} D={c.D} nnzA={c.nnz} nnzB={B.get_nnzs()} D_B={B.num_docs()}", flush=True)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((c.V, a.b)).astype(np.float32)
    Zref = None
    if not a.no_ref:
        bv, br, bo, _ = B.download()
        Bm = sp.csc_matrix((bv.astype(np.float64), br.astype(np.int64), bo), shape=(c.V, B.num_docs()))
        Zref = Bm @ (Bm.T @ X.astype(np.float64))
        cnt = np.diff(Bm.tocsr().indptr)
        order = np.argsort(-cnt, kind="stable")

    def run(label, reps):
        Z = B.multiply(X)   # builds the layout on first use
        H = int(ctx.stat("spmm_head_words"))
        tail = int(ctx.stat("spmm_tail_nnz"))
        msg = f"[{label}] H={H} tail_nnz={tail} ({tail / max(B.get_nnzs(), 1):.3f})"
        if Zref is not None:
            err = np.linalg.norm(Z - Zref) / np.linalg.norm(Zref)
            rown = np.linalg.norm(Zref, axis=1) + 1e-30
            rerr = np.linalg.norm(Z - Zref, axis=1) / rown
            hh = order[:H] if H else order[:0]
            tt = order[H:]
            msg += f" rel_err={err:.3e} max_row_err head={rerr[hh].max() if H else 0:.3e} tail={rerr[tt].max():.3e}"
            if not np.isfinite(err) or err > 1e-5:
                bad = np.argsort(-rerr)[:8]
                msg += f"\n   worst rows {bad.tolist()} ranks {[int(np.where(order == w)[0][0]) for w in bad]} errs {rerr[bad]}"
                msg += f"\n   Z[bad0]={Z[bad[0]]}\n   R[bad0]={Zref[bad[0]]}"
        print(msg, flush=True)
        ctx.call("isle_cuda_set_profiling", 1)
        ctx.call("isle_cuda_reset_stats")
        Xc = np.ascontiguousarray(X.T)
        Zc = np.zeros_like(Xc)
        for _ in range(reps):
            ctx.call("isle_cuda_spsptr_multiply", int(a.b), ptr(Xc), ptr(Zc))
        names = ["spmm_bt", "spmm_b", "spmm_head1", "spmm_tail1", "spmm_head2", "spmm_tail2"]
        t = {n: ctx.stat(n + "_ms") / reps for n in names}
        by = (ctx.stat("spmm_bt_bytes") + ctx.stat("spmm_b_bytes")) / reps
        tot = t["spmm_bt"] + t["spmm_b"]
        print(f"[{label}] per product: " + " ".join(f"{n}={v * 1e3:.1f}us" for n, v in t.items()) +
              f" | pair={tot * 1e3:.1f}us  algorithmic {by / 1e6:.0f} MB -> {by / tot / 1e6:.0f} GB/s", flush=True)
        ctx.call("isle_cuda_set_profiling", 0)

    if not a.only_head_serial:
        ctx.set_option("spmm_head", 0)
        run("gather only", a.reps)
    ctx.set_option("spmm_head_max", a.head_max)
    for o in a.opt:
        k_, v_ = o.split("=")
        ctx.set_option(k_, int(v_))
    for ppm, seg, i8 in [(p_, s_, i_) for i_ in a.i8 for p_ in a.density_ppm for s_ in a.seg]:
        ctx.set_option("spmm_head_i8", i8)
        ctx.set_option("spmm_head_seg", seg)
        ctx.set_option("spmm_head", 1)
        ctx.set_option("spmm_head_density_ppm", ppm)
        ctx.set_option("spmm_fork", 0)
        run(f"head {ppm}ppm seg{seg} i8={i8} serial", a.reps)
        if a.only_head_serial:
            continue
        ctx.set_option("spmm_fork", 1)
        run(f"head {ppm}ppm seg{seg} i8={i8} fork", a.reps)
    ctx.close()


if __name__ == "__main__":
    main()
