#!/usr/bin/env python
"""GPU check of the tcgen05 split-TF32 distance engine against the SIMT fp32 engine and an fp64
host evaluation, through the C ABI (isle_cuda_set_U / project / assign_projected).

    python tools/dist_tc_check.py [D V k]...

Prints, per shape: #documents whose assignment differs between engines, how many of those are
genuine near-ties in fp64 (relative gap < 1e-5), and the timing of both engines.
"""
import sys
import time
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from isle_b200 import _capi, corpus  # noqa: E402
from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix  # noqa: E402


def run(D, V, k, ctx):
    rng = np.random.default_rng(D + V + k)
    c = corpus.generate(V=V, D=D, k=max(4, min(k, 50)), mu=4.0, seed=7)
    counts = c.counts.astype(np.float32)
    lens = np.diff(c.offsets)
    sums = np.add.reduceat(counts, c.offsets[:-1])
    avg = np.float32(int(counts.sum()) // int((lens > 0).sum()))
    vals = (avg * (counts / np.repeat(sums, lens))).astype(np.float32)
    A = SparseMatrix(V, D, ctx)
    A.populate_normalized(vals, c.rows, c.offsets, float(avg), int((lens > 0).sum()))
    z, nn = A.compute_thresholds(0, V, None, max(2, min(k, 50)))
    B = FPSparseMatrix(A)
    B.threshold_and_copy(A, z, nn)
    U, _ = np.linalg.qr(rng.standard_normal((V, k)).astype(np.float32))
    B.set_U(U.astype(np.float32))
    ctx.set_option("dist_kernel", 1)
    P, l2 = B.projected_docs()
    DB = P.shape[0]
    C = P[rng.choice(DB, k, replace=False)].copy() + 0.01 * rng.standard_normal((k, k)).astype(np.float32)
    out = {}
    for eng in (0, 1):
        ctx.set_option("dist_kernel", eng)
        a = B.projected_closest_centers(k, C)   # warm-up
        ctx.call("isle_cuda_reset_stats")
        ctx.call("isle_cuda_set_profiling", 1)
        t0 = time.perf_counter()
        a = B.projected_closest_centers(k, C)
        wall = time.perf_counter() - t0
        ms = ctx.stat("dist_tc_ms" if eng else "dist_simt_ms")
        ctx.call("isle_cuda_set_profiling", 0)
        out[eng] = (a, ms, wall)
    a0, a1 = out[0][0], out[1][0]
    diff = np.nonzero(a0 != a1)[0]
    # fp64 distances for the differing docs
    P64, C64 = P.astype(np.float64), C.astype(np.float64)
    bad = 0
    worst = 0.0
    for d in diff[:2000]:
        dist = np.abs((P64[d] ** 2).sum() + (C64 ** 2).sum(1) - 2.0 * C64 @ P64[d])
        g = abs(dist[a0[d]] - dist[a1[d]]) / max(dist.min(), 1e-30)
        scale = abs(dist[a0[d]] - dist[a1[d]]) / max((P64[d] ** 2).sum(), 1e-30)
        worst = max(worst, scale)
        if scale > 1e-5:
            bad += 1
    # both engines against fp64 argmin
    ref = np.abs((P64 ** 2).sum(1)[:, None] + (C64 ** 2).sum(1)[None, :] - 2.0 * P64 @ C64.T).argmin(1) if DB * k < 4e8 else None
    m0 = int((ref != a0).sum()) if ref is not None else -1
    m1 = int((ref != a1).sum()) if ref is not None else -1
    flops = 2.0 * DB * k * k
    print(f"D_B={DB} V={V} k={k}: engines differ on {len(diff)} docs, {bad} beyond near-tie (worst gap/||d||^2 {worst:.2e}); "
          f"vs fp64 argmin: simt {m0}, tc {m1} mismatches; simt {out[0][1]:.3f} ms, tc {out[1][1]:.3f} ms "
          f"({flops / max(out[1][1], 1e-9) / 1e9:.1f} logical TFLOP/s, 3x on the tensor pipe)", flush=True)
    return bad == 0 and (m1 <= max(5, 2 * m0 + 5) if ref is not None else True)


if __name__ == "__main__":
    shapes = [(4000, 1500, 40), (30000, 3000, 160), (60000, 4000, 520)]
    if len(sys.argv) > 3:
        v = list(map(int, sys.argv[1:]))
        shapes = [tuple(v[i:i + 3]) for i in range(0, len(v), 3)]
    ctx = _capi.Context(0)
    ok = True
    for s in shapes:
        ok = run(*s, ctx) and ok
    ctx.close()
    print("dist_tc_check:", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
