"""Where the end-to-end step loses time against the device-resident one: block_ks with / without the U download and with /
without the background download of B; the plain D2H rate of B."""
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
zetas, evalues = P(np.zeros(V, np.float32)), np.zeros(k, np.float32)
bv, br, bo, hU = P(np.zeros(nnz + 1000, np.float32)), P(np.zeros(nnz + 1000, np.uint64)), P(np.zeros(D + 1, np.int64)), P(np.zeros((k, V), np.float32))
nn, nnzB, DB, nconv = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_int()
ctx.call("isle_cuda_upload_A", V, D, nnz, ptr(h_vals), ptr(h_rows), ptr(h_offs), C.c_float(float(avg)), nz_local)
ctx.call("isle_cuda_thresholds", k, ptr(zetas), C.byref(nn))
ctx.call("isle_cuda_build_B", None, C.byref(nnzB), C.byref(DB))
def ks(U, bg, idle_ms=0.0, rebuild=False):
    if rebuild: ctx.call("isle_cuda_build_B", None, C.byref(nnzB), C.byref(DB))
    if idle_ms: time.sleep(idle_ms * 1e-3)
    t = time.perf_counter()
    if bg: ctx.call("isle_cuda_download_B_begin", ptr(bv), ptr(br), ptr(bo), None)
    ctx.call("isle_cuda_block_ks", k, 10, 100, C.c_float(1e-4), 1, ptr(evalues), ptr(hU) if U else None, C.byref(nconv))
    t1 = time.perf_counter()
    if bg: ctx.call("isle_cuda_download_B_end")
    return (t1 - t) * 1e3, (time.perf_counter() - t1) * 1e3
for _ in range(2): ks(False, False)
for U in (False, True):
    for bg in (False, True):
        r = [ks(U, bg) for _ in range(3)]
        print(f"block_ks U={U} background_B={bg}: " + ", ".join(f"{a:.1f}+{b:.1f}" for a, b in r) + " ms")
for idle in (0.0, 15.0, 100.0):
    r = [ks(True, True, idle, True) for _ in range(3)]
    print(f"build_B, {idle:.0f} ms idle, block_ks U=True background_B=True: " + ", ".join(f"{a:.1f}+{b:.1f}" for a, b in r) + " ms")
for idle in (15.0,):
    r = [ks(False, False, idle, True) for _ in range(3)]
    print(f"build_B, {idle:.0f} ms idle, block_ks U=False background_B=False: " + ", ".join(f"{a:.1f}+{b:.1f}" for a, b in r) + " ms")
def full(up):
    if up: ctx.call("isle_cuda_upload_A", V, D, nnz, ptr(h_vals), ptr(h_rows), ptr(h_offs), C.c_float(float(avg)), nz_local)
    ctx.call("isle_cuda_thresholds", k, ptr(zetas), C.byref(nn))
    return ks(True, True, 0.0, True)
for up in (False, True):
    r = [full(up) for _ in range(3)]
    print(f"upload={up}, thresholds, build_B, block_ks U=True background_B=True: " + ", ".join(f"{a:.1f}+{b:.1f}" for a, b in r) + " ms")
t = time.perf_counter(); ctx.call("isle_cuda_download_B", ptr(bv), ptr(br), ptr(bo), None); dt = time.perf_counter() - t
nb = int(nnzB.value) * 12 + (int(DB.value) + 1) * 8
print(f"download_B synchronous: {dt*1e3:.1f} ms for {nb/1e6:.0f} MB = {nb/dt/1e9:.1f} GB/s")
x = torch.empty(200 << 20, dtype=torch.uint8, device="cuda"); hx = torch.empty(200 << 20, dtype=torch.uint8).pin_memory()
for d in ("d2h", "h2d"):
    torch.cuda.synchronize(); t = time.perf_counter()
    (hx.copy_(x, non_blocking=True) if d == "d2h" else x.copy_(hx, non_blocking=True)); torch.cuda.synchronize()
    print(f"torch {d} 200 MiB pinned: {(200 << 20) / (time.perf_counter() - t) / 1e9:.1f} GB/s")
