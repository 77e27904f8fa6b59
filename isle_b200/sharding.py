"""Host-side helpers of the document-sharded launch (one process per GPU):
which documents a rank owns and how the ingest statistics that ISLETrainer computes over the
whole corpus (reference src/sparseMatrix.cpp:86-98: avg_doc_sz = total tokens // non-empty docs,
integer division, then cast to float) are obtained when every rank only sees its slice.
Collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(num_docs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced document range [d0, d1) of ``rank``."""
    return num_docs * rank // world, num_docs * (rank + 1) // world


def slice_corpus(offsets, rows, counts, d0: int, d1: int):
    """Local CSC (offsets rebased to 0) of documents [d0, d1)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    e0, e1 = int(offsets[d0]), int(offsets[d1])
    return (offsets[d0:d1 + 1] - offsets[d0]).astype(np.int64), rows[e0:e1], counts[e0:e1]


def global_doc_stats(counts, offsets, group=None) -> tuple[np.float32, int, int]:
    """(avg_doc_sz, local non-empty docs, global non-empty docs).  Token totals and document counts
    are exact integers, so the all-reduce order cannot change the result."""
    import torch
    import torch.distributed as dist

    lens = np.diff(np.asarray(offsets, dtype=np.int64))
    nz_local = int((lens > 0).sum())
    tot = torch.tensor([int(np.asarray(counts).astype(np.uint64).sum()), nz_local], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            tot = tot.cuda()
        dist.all_reduce(tot, group=group)
    tokens, nz_global = int(tot[0].item()), int(tot[1].item())
    return np.float32(tokens // max(nz_global, 1)), nz_local, nz_global


def normalize_shard(counts, offsets, avg_doc_sz) -> np.ndarray:
    """normalize_docs (reference src/sparseMatrix.cpp:136-167) for a slice, with the GLOBAL average:
    value = avg_doc_sz * ((float)count / doc_sum), doc_sum exact in fp32 below 2^24 tokens."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(offsets)
    c = np.asarray(counts).astype(np.float32)
    nonempty = lens > 0
    sums = np.zeros(len(lens), dtype=np.int64)
    if c.size:
        sums[nonempty] = np.add.reduceat(np.asarray(counts).astype(np.int64), offsets[:-1][nonempty])
    assert sums.max(initial=0) < (1 << 24), "document longer than 2^24 tokens: fp32 doc_sum no longer exact"
    return (np.float32(avg_doc_sz) * (c / np.repeat(sums.astype(np.float32), lens))).astype(np.float32)
