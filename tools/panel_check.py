"""Panel-product engines of the device eigensolver in isolation (isle_cuda_panel_products): C = W^T F, F -= W C
against float64, for the fp32 FMA kernels (engine 0/1) and the tcgen05 split-TF32 kernel (engine 2).

    python tools/panel_check.py [--big]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from isle_b200 import _capi  # noqa: E402
from isle_b200._capi import ptr  # noqa: E402


def run(ctx, n, rows, b, engine, rng):
    Q, _ = np.linalg.qr(rng.standard_normal((n, rows)))
    W = np.asfortranarray(Q.astype(np.float32))
    F = np.asfortranarray((rng.standard_normal((n, b)) * np.logspace(0, -3, b)[None, :]).astype(np.float32))
    Wc = np.ascontiguousarray(W.T)          # column-major n x rows == C-order rows x n
    Fc = F.T.copy()
    Cc = np.zeros((b, rows), np.float32)
    ctx.call("isle_cuda_panel_products", n, rows, b, ptr(Wc), ptr(Fc), ptr(Cc), engine)
    Cref = W.astype(np.float64).T @ F.astype(np.float64)
    Fref = F.astype(np.float64) - W.astype(np.float64) @ Cc.T.astype(np.float64)     # with the coefficients the engine used
    fn = np.linalg.norm(F, axis=0)
    e_c = np.max(np.abs(Cc.T - Cref) / fn[None, :])
    e_f = np.max(np.abs(Fc.T - Fref) / fn[None, :])
    # orthogonality left after the pass, relative to ||F|| (what the next pass has to remove)
    orth = np.max(np.abs(W.astype(np.float64).T @ Fc.T.astype(np.float64)) / fn[None, :])
    return e_c, e_f, orth


def main():
    ctx = _capi.Context(0)
    rng = np.random.default_rng(0)
    shapes = [(600, 20, 10), (1000, 7, 3), (5000, 50, 10), (4096, 130, 16), (20000, 300, 10), (20004, 129, 1)]
    if "--big" in sys.argv:
        shapes.append((102000, 210, 10))
    ok = True
    for n, rows, b in shapes:
        for engine in (0, 1, 2):
            t = time.time()
            e_c, e_f, orth = run(ctx, n, rows, b, engine, rng)
            good = e_c < 2e-6 and e_f < 2e-6
            ok &= good
            print(f"n={n} rows={rows} b={b} engine={engine}: |C-Cref|/|F|={e_c:.2e} |F-Fref|/|F|={e_f:.2e} "
                  f"|W^T F'|/|F|={orth:.2e} {'ok' if good else 'FAIL'} ({time.time() - t:.2f}s)", flush=True)
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
