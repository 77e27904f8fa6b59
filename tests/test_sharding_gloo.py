"""world_size = 2 over gloo on CPU: the host-side logic of the document-sharded path and the
algebra the device collectives rely on -- global ingest statistics from slices,
per-word histograms that add across ranks to the single-process thresholds, B slices and the
global column numbering, Lloyd center sums, the integer member counts of the full-dimensional Lloyd.
No GPU, no CUDA calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isle_b200 import corpus, sharding
from oracle import isle_oracle as O

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, golden_path, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dict(np.load(golden_path))
        c = corpus.generate("tiny")
        d0, d1 = sharding.shard_bounds(c.D, rank, world)
        lo, lr, lc = sharding.slice_corpus(c.offsets, c.rows, c.counts, d0, d1)
        avg, nz_local, nz_global = sharding.global_doc_stats(lc, lo)
        vals = sharding.normalize_shard(lc, lo, avg)
        full_vals, full_avg, full_nz = O.normalize_docs(c.counts, c.offsets)
        e0, e1 = int(c.offsets[d0]), int(c.offsets[d1])
        assert np.float32(full_avg) == avg and full_nz == nz_global
        assert np.array_equal(vals, full_vals[e0:e1])                      # bit-exact normalisation

        # thresholds: local histograms summed across ranks == single-process rule
        bins = int(avg) + 2
        h = torch.from_numpy(O.word_histogram(vals, lr, c.V, bins))
        dist.all_reduce(h)
        z, nn = O.thresholds_from_histogram(h.numpy(), nz_global, c.k)
        assert np.array_equal(z, g["zetas"]) and nn == int(g["new_nnzs"])

        # B: the local build is the slice of the global B; column numbering by all-gathered D_B
        bv, br, bo, oc = O.threshold_and_copy(vals, lr, lo, z)
        db = torch.tensor([len(oc)], dtype=torch.int64)
        all_db = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_db, db)
        off = int(sum(int(x) for x in all_db[:rank]))
        G_oc = g["B_original_cols"].astype(np.int64)
        sel = np.nonzero((G_oc >= d0) & (G_oc < d1))[0]
        assert sel[0] == off and len(sel) == len(oc)
        assert np.array_equal(G_oc[sel] - d0, np.asarray(oc, dtype=np.int64))
        b0, b1 = int(g["B_offsets"][sel[0]]), int(g["B_offsets"][sel[-1] + 1])
        assert np.array_equal(g["B_rows"][b0:b1].astype(np.int64), np.asarray(br, dtype=np.int64))
        assert np.array_equal(g["B_vals"][b0:b1], bv)

        # operator: Z = sum over ranks of B_g (B_g^T X)
        B_local = O.to_csc(bv, br, bo, c.V)
        X = np.random.default_rng(5).standard_normal((c.V, 10)).astype(np.float32)
        Zl = torch.from_numpy(O.spsptr_multiply(B_local, X).astype(np.float64))
        dist.all_reduce(Zl)
        B_full = O.to_csc(g["B_vals"], g["B_rows"], g["B_offsets"], c.V)
        Zf = O.spsptr_multiply(B_full, X).astype(np.float64)
        assert np.max(np.abs(Zl.numpy() - Zf)) <= 1e-4 * np.max(np.abs(Zf))

        # Lloyd: center sums / counts add across ranks
        U = np.ascontiguousarray(g["U_colmajor"].reshape(c.k, c.V).T) if g["U_colmajor"].ndim == 1 else g["U_colmajor"]
        P_local, P_full = O.project(B_local, U), O.project(B_full, U)
        C0 = g["centers_lowd_init"].reshape(c.k, c.k)
        a_local = O.closest_centers(P_local, O.docs_l2sq(P_local), C0)
        a_full = O.closest_centers(P_full, O.docs_l2sq(P_full), C0)
        assert np.array_equal(a_local, a_full[sel])
        k = C0.shape[0]
        sums = np.zeros((k, P_local.shape[1])); np.add.at(sums, a_local, P_local.astype(np.float64))
        cnt = np.bincount(a_local, minlength=k).astype(np.int64)
        ts, tc = torch.from_numpy(sums), torch.from_numpy(cnt)
        dist.all_reduce(ts); dist.all_reduce(tc)
        fs = np.zeros_like(sums); np.add.at(fs, a_full, P_full.astype(np.float64))
        assert np.array_equal(tc.numpy(), np.bincount(a_full, minlength=k))
        assert np.allclose(ts.numpy(), fs, rtol=1e-9, atol=1e-9)
        # stage F (Lloyd on the full-dimensional B): assignments are local, integer member counts per (cluster, word)
        # and cluster sizes add across ranks, so the centers built from the totals equal the single-process ones
        gF = dict(np.load(os.path.join(os.path.dirname(golden_path), "tiny_stageF.npz")))
        CF = gF["centers_in"].reshape(c.k, c.V)
        aF_local = np.argmin(np.abs(O.dist_matrix_full(B_local, O.docs_l2sq_full(B_local), CF)), axis=1)
        aF_full = np.argmin(np.abs(O.dist_matrix_full(B_full, O.docs_l2sq_full(B_full), CF)), axis=1)
        assert np.array_equal(aF_local, aF_full[sel])

        def member_counts(Bm, a):
            import scipy.sparse as sp
            M = sp.csr_matrix((np.ones(Bm.shape[1], np.int64), (np.arange(Bm.shape[1]), a)), shape=(Bm.shape[1], c.k))
            return np.asarray(((Bm != 0).astype(np.int64) @ M).todense()).T.copy()          # [k, V]

        tcnt = torch.from_numpy(member_counts(B_local, aF_local))
        tsz = torch.from_numpy(np.bincount(aF_local, minlength=c.k).astype(np.int64))
        dist.all_reduce(tcnt); dist.all_reduce(tsz)
        assert np.array_equal(tcnt.numpy(), member_counts(B_full, aF_full))
        assert np.array_equal(tsz.numpy(), np.bincount(aF_full, minlength=c.k))
        sz = np.sqrt(z).astype(np.float32)
        cen = np.where(tsz.numpy()[:, None] > 0, (sz[None, :] * tcnt.numpy().astype(np.float32)) / np.maximum(tsz.numpy(), 1)[:, None].astype(np.float32), 0)
        C1, _ = O.lloyds_iter_full(B_full, O.docs_l2sq_full(B_full), CF)
        assert np.max(np.abs(cen - C1)) <= 1e-5 * np.max(np.abs(C1))
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_world2(tmp_path):
    golden = os.path.join(os.path.dirname(__file__), "golden", "tiny.npz")
    keys = set(np.load(golden).keys())
    need = {"zetas", "new_nnzs", "B_original_cols", "B_offsets", "B_rows", "B_vals", "U_colmajor", "centers_lowd_init"}
    if not need <= keys:
        pytest.skip(f"golden fixture lacks {sorted(need - keys)}")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(WORLD, _free_port(), golden, out), nprocs=WORLD, join=True)
    assert sorted(out.keys()) == list(range(WORLD))


def _ordered(x):
    """seg_select.cuh ordered(): monotone map float32 -> uint32 (larger float, larger integer)."""
    b = np.asarray(x, np.float32).view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)


def _unordered(o):
    o = np.uint32(o)
    return np.array([o & 0x7FFFFFFF if o & 0x80000000 else ~o], np.uint32).view(np.float32)[0]


def _radix_select_worker(rank, world, port, out):
    """The exact distributed select of seg_select.cuh (stages G / H and the importance-sampling pivot under sharding): four
    rounds of a 256-bin histogram of the next byte of ordered(x) over the values that still match the prefix, summed over
    the ranks; no values travel.  Must return the (kth + 1)-th largest of the union, ties and signs included."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                                   # same stream on every rank: everybody knows the union
        parts = [np.concatenate([rng.standard_normal(5000).astype(np.float32), np.zeros(40, np.float32),
                                 np.float32([1e-42, -1e-42, 3.5, 3.5, 3.5, -0.0]), rng.random(2000).astype(np.float32)])
                 for _ in range(world)]
        mine, union = parts[rank], np.sort(np.concatenate(parts))[::-1]
        for kth in (0, 1, 17, 4999, len(union) // 2, len(union) - 1):
            prefix, rem = np.uint32(0), kth
            o = _ordered(mine)
            for rnd in range(4):
                match = np.ones(len(o), bool) if rnd == 0 else (o >> np.uint32(32 - 8 * rnd)) == prefix
                h = torch.from_numpy(np.bincount((o[match] >> np.uint32(24 - 8 * rnd)) & 255, minlength=256).astype(np.int64))
                dist.all_reduce(h)
                h = h.numpy()
                b = 255
                while b > 0 and rem >= h[b]:
                    rem -= int(h[b]); b -= 1
                prefix = np.uint32((int(prefix) << 8) | b)
            got = _unordered(prefix)
            assert got == union[kth] or (got == 0 and union[kth] == 0), (kth, got, union[kth])
        # importance sampling (isle_cuda_sample_docs under sharding): keys from GLOBAL document numbers, pivot = the
        # (floor(rate D) + 1)-th largest key of the union -> the selection equals the single-process one
        D = 9001
        d0, d1 = sharding.shard_bounds(D, rank, world)
        keys = np.random.default_rng(3).random(D).astype(np.float32) ** 3          # stand-in for u_d^(1 / w_d), keyed by global d
        nth = int(np.float32(0.1) * np.float32(D))
        pivot = np.sort(keys)[::-1][nth]
        o, prefix, rem = _ordered(keys[d0:d1]), np.uint32(0), nth
        for rnd in range(4):
            match = np.ones(len(o), bool) if rnd == 0 else (o >> np.uint32(32 - 8 * rnd)) == prefix
            h = torch.from_numpy(np.bincount((o[match] >> np.uint32(24 - 8 * rnd)) & 255, minlength=256).astype(np.int64))
            dist.all_reduce(h)
            h = h.numpy()
            b = 255
            while b > 0 and rem >= h[b]:
                rem -= int(h[b]); b -= 1
            prefix = np.uint32((int(prefix) << 8) | b)
        assert _unordered(prefix) == pivot
        sel = torch.tensor([int((keys[d0:d1] >= _unordered(prefix)).sum())])
        dist.all_reduce(sel)
        assert int(sel) == int((keys >= pivot).sum())
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_distributed_radix_select_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_radix_select_worker, args=(WORLD, _free_port(), out), nprocs=WORLD, join=True)
    assert sorted(out.keys()) == list(range(WORLD))


def test_shard_bounds_cover_all_documents():
    for D in (0, 1, 7, 1500, 300000):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(D, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == D
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
