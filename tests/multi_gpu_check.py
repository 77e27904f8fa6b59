#!/usr/bin/env python
"""Document-sharded spectral core vs the single-GPU run on the same corpus.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py [config]

Every rank holds a contiguous slice of the documents; the library exchanges per-word histograms,
the V x b operator blocks, k-means++ picks and Lloyd center sums over NCCL.  Every rank then
repeats the computation alone on the whole corpus (world = 1 context on its own GPU) and checks:
thresholds and its slice of B bit-exact, singular values within 1e-4 relative, principal angle
below 1e-3, Lloyd objective within 1e-4 from identical initial centers, and the full-dimensional Lloyd
(stage F) bit-identical from identical centers (its sharded reduction is integer).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from isle_b200 import _capi, corpus  # noqa: E402
from isle_b200.sparse_matrix import FPSparseMatrix, SparseMatrix  # noqa: E402
from oracle import isle_oracle as O  # noqa: E402  (checker only)


def stages(ctx, V, D, k, vals, rows, offsets, avg, nz, centers0=None, U_override=None, seed=3, full_centers0=None,
           total_docs=None):
    A = SparseMatrix(V, D, ctx)
    A.populate_normalized(vals, rows, offsets, avg, nz)
    z, nn = A.compute_thresholds(0, V, None, k)
    B = FPSparseMatrix(A)
    B.threshold_and_copy(A, z, nn)
    bv, br, bo, oc = B.download()
    ev, U = B.compute_block_ks(k, seed=seed, want_U=True)
    row_sharded = ctx.stat("ks_row_sharded")
    if U_override is not None:      # identical projection for the k-means comparison
        B.set_U(U_override)
    seeds, coords, res = B.kmeans_init_on_projected_space(k, 1, seed=seed)
    c0 = coords.copy() if centers0 is None else centers0.copy()
    B.run_lloyds_on_projected_space(k, c0, None, 10)
    # stage F: lift, clean up, Lloyd on the full-dimensional B from given centers
    if full_centers0 is None:
        full = np.ascontiguousarray(B.left_multiply_by_U_Spectra(c0, k, k).T)
    else:
        full = full_centers0.copy()
    full_in = full.copy()
    B.cleanup_after_eigensolver()
    B.run_lloyds(k, full, None, 10)
    # stages G / H from that partition: thresholds and catchwords are global quantities, the
    # (doc, topic) sums are local, the model is allreduced
    cl = np.full(D, 0xFFFFFFFF, np.uint32)
    cl[oc.astype(np.int64)] = B.last_lloyd_full["assign"]
    thr = A.catchword_thresholds(k, O.catchword_rank(total_docs or D, k), cl)
    cw = A.find_catchwords(k, None)
    model, dts, _ = A.construct_topic_model(k, cl, cw, want_pairs=False, total_docs=total_docs or D)
    return dict(catch_thr=thr, catchwords=cw, model=model, dts=dts, row_sharded=row_sharded,
                z=z, nn=nn, bv=bv, br=br, bo=bo, oc=oc, ev=ev, U=U, seeds=seeds, coords=coords, centers=c0,
                obj=B.last_lloyd["objective"], assign=B.last_lloyd["assign"], iters=B.last_lloyd["iters"],
                full_in=full_in, full=full, full_obj=B.last_lloyd_full["objective"], full_assign=B.last_lloyd_full["assign"],
                full_iters=B.last_lloyd_full["iters"])


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(_capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    c = corpus.generate(name)
    vals, avg, nz = O.normalize_docs(c.counts, c.offsets)
    d0, d1 = c.D * rank // world, c.D * (rank + 1) // world
    e0, e1 = int(c.offsets[d0]), int(c.offsets[d1])
    lo = (c.offsets[d0:d1 + 1] - c.offsets[d0]).astype(np.int64)
    nz_local = int((np.diff(lo) > 0).sum())

    sh = _capi.Context(local, rank, world, bytes(idt.cpu().numpy().tobytes()))
    # the check covers the row-sharded Krylov basis unless told otherwise (the library's default switches it on only for
    # bases that stream from HBM)
    sh.set_option("ks_row_shard", 0 if os.environ.get("ISLE_KS_ROW_SHARD") == "0" else 1)
    if os.environ.get("ISLE_P2P") == "0":
        sh.set_option("p2p", 0)
    # the library's own collectives over peer memory (coll.cu) against NCCL, bit for bit, before anything depends on them
    import ctypes as C
    st_mism, active = C.c_uint64(), C.c_int()
    sh.call("isle_cuda_selftest_collectives", C.byref(st_mism), C.byref(active))
    p2p_note = f"p2p collectives {'ON' if active.value else 'off (NCCL)'}, self-test mismatches {st_mism.value}"
    p2p_bad = st_mism.value != 0 or (active.value == 0 and os.environ.get("ISLE_P2P") != "0")
    if active.value and rank == 0:
        print("collective latency us (p2p / nccl): " + ", ".join(
            f"{sz} {sh.stat('selftest_p2p_' + sz + '_us'):.1f} / {sh.stat('selftest_nccl_' + sz + '_us'):.1f}" for sz in ("4mb", "256kb", "50kb", "2kb")),
            flush=True)
    r = stages(sh, c.V, d1 - d0, c.k, vals[e0:e1], c.rows[e0:e1], lo, float(avg), nz_local, total_docs=c.D)
    # the single-GPU run; its k-means stages get the sharded run's U and k-means++ centers so that
    # both Lloyd runs see the same projection and start identically
    one = _capi.Context(local)
    s = stages(one, c.V, c.D, c.k, vals, c.rows, c.offsets, float(avg), nz, centers0=r["coords"], U_override=r["U"],
               full_centers0=r["full_in"])

    ok = True

    def check(cond, msg):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAIL: {msg}", flush=True)

    check(not p2p_bad, p2p_note)
    check(np.array_equal(r["z"], s["z"]), "thresholds differ from the single-GPU run")
    # compute_thresholds returns the kept-entry count of the WHOLE corpus on every rank (include/isle_cuda.h)
    check(r["nn"] == s["nn"], f"kept-entry count {r['nn']} differs from the single-GPU run's {s['nn']}")
    tot = torch.tensor([len(r["bv"])], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    check(int(tot.item()) == s["nn"], "local nnz of B do not add up to the global kept-entry count")
    # my slice of the single-GPU B: original docs [d0, d1)
    sel = np.nonzero((s["oc"] >= d0) & (s["oc"] < d1))[0]
    check(len(sel) == len(r["oc"]) and np.array_equal(s["oc"][sel] - d0, r["oc"]), "original_cols differ")
    if len(sel):
        b0, b1 = int(s["bo"][sel[0]]), int(s["bo"][sel[-1] + 1])
        check(np.array_equal(s["br"][b0:b1], r["br"]) and np.array_equal(s["bv"][b0:b1], r["bv"]) and
              np.array_equal(s["bo"][sel[0]:sel[-1] + 2] - b0, r["bo"]), "B slice differs")
    # the Krylov basis of the sharded run is row-sharded over the vocabulary (SURVEY 8e option B) unless switched off
    want_rs = 0.0 if os.environ.get("ISLE_KS_ROW_SHARD") == "0" else 1.0
    check(r["row_sharded"] == want_rs and s["row_sharded"] == 0.0, f"row sharding flag {r['row_sharded']} / {s['row_sharded']}")
    sv_r, sv_s = np.sqrt(r["ev"]), np.sqrt(s["ev"])
    rel = float(np.max(np.abs(sv_r - sv_s) / sv_s))
    check(rel < 1e-4, f"singular values differ: {rel:.2e}")
    ang = O.principal_angle_sin(r["U"], s["U"])
    if name == "c3m":
        # the k-th eigenvalue of c3m is 0.46 % above the next: a solver stopping at residual 1e-4 pins span(U) only to
        # ~tol / gap = 2e-2 (tests/test_gpu_largek.py); the well-separated leading 300 Ritz vectors must agree to 1e-3
        U1, U2 = r["U"].astype(np.float64), s["U"].astype(np.float64)
        lead = float(np.linalg.norm(U1[:, :300] - U2 @ (U2.T @ U1[:, :300]), 2))
        check(lead < 1e-3 and ang < 1e-2, f"principal angle {ang:.2e}, leading 300 vectors {lead:.2e}")
    else:
        check(ang < 1e-3, f"principal angle {ang:.2e}")
    # Lloyd: the sharded run used its own seeds; re-run it from the same centers as `s`
    # (already the case: s started from r's coords).  Compare objective and local assignments.
    obj_rel = abs(r["obj"] - s["obj"]) / s["obj"]
    check(obj_rel < 1e-4, f"Lloyd objective differs: {obj_rel:.2e} ({r['obj']} vs {s['obj']})")
    if len(sel):
        mism = float((s["assign"][sel] != r["assign"]).mean())
        # The two runs add the same center sums in different orders (per-rank partial sums, then the all-reduce), so centers
        # can differ in the last bit; usually no document sits that close to a tie and the partitions are identical (0.00 %,
        # objective equal to 1e-15), but one flipped near-tie in an early iteration moves two centers and cascades over the
        # ten iterations (seen: 0.1 % - 0.7 % at k = 320 with either transport, objective still within 1e-5).  The bar on the
        # objective above is the north star's; this one bounds the cascade.
        check(mism < (2e-2 if name == "c3m" else 5e-3), f"Lloyd assignments differ on {mism:.2%} of the local docs")
    else:
        mism = 0.0
    # stage F from identical centers: integer member counts are allreduced, so centers and partition agree exactly
    check(np.array_equal(r["full"], s["full"]), f"full-dimensional Lloyd centers differ: {np.abs(r['full'] - s['full']).max():.2e}")
    check(r["full_iters"] == s["full_iters"], "full-dimensional Lloyd iteration counts differ")
    check(abs(r["full_obj"] - s["full_obj"]) <= 1e-9 * s["full_obj"], "full-dimensional Lloyd objective differs")
    if len(sel):
        check(np.array_equal(s["full_assign"][sel], r["full_assign"]), "full-dimensional Lloyd assignments differ")
    # stages G / H: thresholds (distributed radix select) and catchwords bit-identical, local (doc, topic) sums = the
    # slice of the single-GPU list, the allreduced model within fp32 rounding
    check(np.array_equal(r["catch_thr"].view(np.uint32), s["catch_thr"].view(np.uint32)), "catchword thresholds differ")
    check(all(np.array_equal(a, b) for a, b in zip(r["catchwords"], s["catchwords"])), "catchwords differ")
    sd, st_, sv = s["dts"]
    m = (sd >= d0) & (sd < d1)
    rd, rt, rv = r["dts"]
    check(np.array_equal(sd[m] - d0, rd) and np.array_equal(st_[m], rt) and np.array_equal(sv[m].view(np.uint32), rv.view(np.uint32)),
          "document-topic sums differ")
    okm = ~np.isnan(s["model"])
    check(np.array_equal(np.isnan(r["model"]), ~okm) and
          np.max(np.abs(r["model"][okm] - s["model"][okm])) <= 1e-6 * np.max(np.abs(s["model"][okm])), "topic model differs")
    # k-means++ seeds are global column ids of B and must be distinct
    check(len(set(r["seeds"].tolist())) == c.k and int(r["seeds"].max()) < len(s["oc"]), "bad k-means++ seeds")
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(f"multi_gpu_check[{name}, world={world}]: sigma rel {rel:.2e}, angle {ang:.2e}, objective rel {obj_rel:.2e}, "
              f"assign mismatch {mism:.2%}, lloyd iters {r['iters']}/{s['iters']}, {p2p_note}, "
              f"p2p collectives used {sh.stat('p2p_collectives') if active.value else 0:.0f} -> "
              + ("OK" if int(flag.item()) == 0 else "FAILED"), flush=True)
    sh.close()
    one.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
