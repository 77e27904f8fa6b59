"""Wall-clock per C-ABI call of one end-to-end step of bench.py (host buffers in, host results out), c2."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from isle_b200 import _capi, corpus, sharding
from isle_b200._capi import ptr
ctx = _capi.Context(0)
cfg = corpus.CONFIGS["c2"]; V, k = cfg["V"], cfg["k"]
c = bench.make_corpus("c2", 0, 0); D, nnz = c.D, c.nnz
avg, nz_local, _ = sharding.global_doc_stats(c.counts, c.offsets)
vals = sharding.normalize_shard(c.counts, c.offsets, avg)
P = bench.pinned
h_vals, h_rows, h_offs = P(vals), P(c.rows.astype(np.uint64)), P(c.offsets.astype(np.int64))
zetas, evalues, seeds = P(np.zeros(V, np.float32)), np.zeros(k, np.float32), np.zeros(k, np.uint64)
cl, centers = np.zeros((k, k), np.float32), P(np.zeros((k, V), np.float32))
bv, br, bo, hU = P(np.zeros(nnz + 1000, np.float32)), P(np.zeros(nnz + 1000, np.uint64)), P(np.zeros(D + 1, np.int64)), P(np.zeros((k, V), np.float32))
def step(log):
    def T(name, *a):
        t = time.perf_counter(); ctx.call(name, *a); log.append((name, (time.perf_counter() - t) * 1e3))
    nn, nnzB, DB, nconv, res, obj, it = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_int(), C.c_float(), C.c_double(), C.c_int()
    T("isle_cuda_upload_A", V, D, nnz, ptr(h_vals), ptr(h_rows), ptr(h_offs), C.c_float(float(avg)), nz_local)
    T("isle_cuda_thresholds", k, ptr(zetas), C.byref(nn))
    T("isle_cuda_build_B", None, C.byref(nnzB), C.byref(DB))
    oc = np.zeros(int(DB.value), np.uint64)
    T("isle_cuda_download_B", None, None, None, ptr(oc))
    T("isle_cuda_download_B_begin", ptr(bv), ptr(br), ptr(bo), None)
    T("isle_cuda_block_ks", k, 10, 100, C.c_float(1e-4), 1, ptr(evalues), ptr(hU), C.byref(nconv))
    T("isle_cuda_download_B_end")
    T("isle_cuda_kmeanspp", k, 1, ptr(seeds), ptr(cl), C.byref(res))
    T("isle_cuda_lloyd_projected", k, ptr(cl), 10, None, C.byref(obj), C.byref(it))
    T("isle_cuda_lift_centers", k, ptr(cl), k, ptr(centers))
    T("isle_cuda_cleanup_eigensolver")
for i in range(4):
    log = []; t = time.perf_counter(); step(log); tot = (time.perf_counter() - t) * 1e3
print(f"total {tot:.1f} ms")
for n, ms in log: print(f"  {n:34s} {ms:7.2f} ms")
