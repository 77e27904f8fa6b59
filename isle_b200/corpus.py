"""Synthetic Zipfian dominant-admixture corpora (harness, not product code).

The reference ships no data, so every test/bench input is generated here with the
model SURVEY.md section 8(d) specifies:

* k topics over V words.  Base word distribution Zipf(s=1.07) under a random
  permutation; topic t owns a disjoint catchword set of ``max(5, V // (20 k))``
  words boosted x50 and is otherwise ``base * Gamma(0.3)`` noise, renormalised.
* doc d: dominant topic t_d ~ Uniform(k); mixture 0.8 e_{t_d} + 0.2 Dirichlet(0.1 1_k);
  length L_d ~ LogNormal(mu, 0.5) clipped to [20, 2000]; counts ~ Multinomial(L_d, mixture @ topics).

Output is the doc-major CSC of raw counts that ``ISLETrainer`` builds from its
``<doc> <word> <count>`` text input (reference src/trainer.cpp:232-293,
src/sparseMatrix.cpp:58-106): ``offsets int64[D+1]``, ``rows uint32[nnz]`` (ascending
inside a doc), ``counts uint32[nnz]``.  ``normalize`` restates
SparseMatrix::normalize_docs (src/sparseMatrix.cpp:136-167) in numpy.

Two backends draw the same model: numpy (CPU, small configs / tests) and torch on
CUDA (bench-sized configs; the Dirichlet part is then approximated by 8 atoms per
doc, see ``_doc_topic_tokens_torch``).  RNG streams differ between backends; every
comparison in this repo is made on one corpus inside one process or via files.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Optional

import numpy as np


@dataclasses.dataclass
class Corpus:
    V: int
    D: int
    k: int
    offsets: np.ndarray  # int64[D+1]
    rows: np.ndarray     # uint32[nnz]
    counts: np.ndarray   # uint32[nnz]
    dominant: Optional[np.ndarray] = None  # int32[D] planted dominant topic

    @property
    def nnz(self) -> int:
        return int(self.offsets[-1])

    def write_bin(self, path: str) -> None:
        """corpus.bin for oracle/ref_dump (see oracle/ref_dump.cpp header)."""
        with open(path, "wb") as f:
            np.array([self.V, self.D, self.nnz], dtype=np.int64).tofile(f)
            self.offsets.astype(np.int64).tofile(f)
            self.rows.astype(np.uint32).tofile(f)
            self.counts.astype(np.uint32).tofile(f)

    def write_text(self, tdf_path: str, vocab_path: str) -> None:
        """1-based ``<doc> <word> <count>`` lines + a vocab file, the ISLETrain CLI input
        (reference README.md:44-47, include/utils.h:104-228)."""
        docs = np.repeat(np.arange(self.D, dtype=np.int64), np.diff(self.offsets))
        arr = np.stack([docs + 1, self.rows.astype(np.int64) + 1, self.counts.astype(np.int64)], 1)
        np.savetxt(tdf_path, arr, fmt="%d")
        with open(vocab_path, "w") as f:
            for w in range(self.V):
                f.write(f"w{w}\n")


# Named shapes (BASELINE.json configs).  mu is calibrated so distinct words/doc ~ nnz/D.
CONFIGS = {
    "tiny": dict(V=600, D=1500, k=20, mu=3.9, seed=20239),
    "c1": dict(V=5000, D=10000, k=20, mu=5.19, seed=20241),
    "c2": dict(V=102000, D=300000, k=100, mu=5.93, seed=20242),
    "c3": dict(V=141000, D=8200000, k=2000, mu=4.29, seed=20243),
    "c4": dict(V=100000, D=11000000, k=2000, mu=4.485, seed=20244),
    # one of the eight document shards of c3 (what each GPU of an 8 x B200 box holds): same V, k, mu
    "c3s": dict(V=141000, D=1025000, k=2000, mu=4.29, seed=20243),
    # c3-shaped miniature the reference finishes in seconds: k > 256 (several center tiles in the tcgen05 distance
    # kernel, >= 17 new k-means++ centers per round), ncv = 650 (panel products over more than one K segment)
    "c3m": dict(V=6000, D=40000, k=320, mu=4.29, seed=20250),
}


def _topics(V: int, k: int, rng: np.random.Generator) -> np.ndarray:
    base = 1.0 / np.arange(1, V + 1, dtype=np.float64) ** 1.07
    base = base[rng.permutation(V)]
    ncatch = max(5, V // (20 * k))
    catch = rng.permutation(V)[: ncatch * k].reshape(k, ncatch)
    T = np.empty((k, V), dtype=np.float64)
    for t in range(k):
        row = base * rng.gamma(0.3, size=V)
        row[catch[t]] = base[catch[t]] * 50.0
        T[t] = row / row.sum()
    return T


def generate(name: str | None = None, *, V: int = 0, D: int = 0, k: int = 0, mu: float = 4.5,
             seed: int = 0, backend: str = "numpy", device: str = "cuda", doc_seed: int | None = None) -> Corpus:
    """``seed`` fixes the topic model; ``doc_seed`` (default: continue the same stream) draws a
    different set of documents from it -- document shards of one corpus for the multi-GPU runs."""
    if name is not None:
        cfg = dict(CONFIGS[name])
        V, D, k, mu, seed = cfg["V"], cfg["D"], cfg["k"], cfg["mu"], cfg["seed"]
    if backend == "torch":
        return _generate_torch(V, D, k, mu, seed, device, doc_seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    T = _topics(V, k, rng)
    if doc_seed is not None:
        rng = np.random.Generator(np.random.PCG64([seed, doc_seed]))
    cdf = np.cumsum(T, axis=1)
    cdf /= cdf[:, -1:]
    flat = (cdf + np.arange(k)[:, None]).ravel()  # monotone over (topic, word)

    dominant = rng.integers(0, k, size=D)
    L = np.clip(np.rint(rng.lognormal(mu, 0.5, size=D)), 20, 2000).astype(np.int64)
    offs_parts, rows_parts, cnt_parts = [], [], []
    chunk = max(1, min(D, 4_000_000 // max(1, int(L.mean()))))
    for d0 in range(0, D, chunk):
        d1 = min(D, d0 + chunk)
        n = d1 - d0
        Lc = L[d0:d1]
        tok_doc = np.repeat(np.arange(n), Lc)
        ntok = tok_doc.size
        # token topic: 0.8 dominant, 0.2 from the doc's Dirichlet(0.1)
        mix = rng.dirichlet(np.full(k, 0.1), size=n)
        mcdf = np.cumsum(mix, axis=1)
        mcdf /= mcdf[:, -1:]
        u = rng.random(ntok)
        sec = (mcdf[tok_doc] < u[:, None]).sum(1) if k <= 64 else _rowwise_search(mcdf, tok_doc, u)
        sec = np.minimum(sec, k - 1)
        topic = np.where(rng.random(ntok) < 0.8, dominant[d0:d1][tok_doc], sec)
        w = np.searchsorted(flat, topic + rng.random(ntok), side="right") - topic * V
        w = np.clip(w, 0, V - 1)
        key = tok_doc.astype(np.int64) * V + w
        uk, cnt = np.unique(key, return_counts=True)
        dd = uk // V
        rows_parts.append((uk % V).astype(np.uint32))
        cnt_parts.append(cnt.astype(np.uint32))
        offs_parts.append(np.bincount(dd, minlength=n).astype(np.int64))
    per_doc = np.concatenate(offs_parts)
    offsets = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(per_doc, out=offsets[1:])
    return Corpus(V, D, k, offsets, np.concatenate(rows_parts), np.concatenate(cnt_parts),
                  dominant.astype(np.int32))


def _rowwise_search(mcdf: np.ndarray, tok_doc: np.ndarray, u: np.ndarray) -> np.ndarray:
    n, k = mcdf.shape
    flat = (mcdf + np.arange(n)[:, None]).ravel()
    return np.searchsorted(flat, tok_doc + u, side="right") - tok_doc * k


def _generate_torch(V: int, D: int, k: int, mu: float, seed: int, device: str, doc_seed: int | None = None) -> Corpus:
    """Same model drawn with torch on ``device`` (bench-sized corpora in seconds)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    dev = torch.device(device)

    base = 1.0 / torch.arange(1, V + 1, dtype=torch.float64, device=dev) ** 1.07
    base = base[torch.randperm(V, generator=g, device=dev)]
    ncatch = max(5, V // (20 * k))
    catch = torch.randperm(V, generator=g, device=dev)[: ncatch * k].reshape(k, ncatch)
    # Gamma(0.3) noise: torch's gamma sampler has no generator argument -> seed the global one
    torch.manual_seed(seed)
    flat_parts = []
    tchunk = max(1, (1 << 27) // V)
    for t0 in range(0, k, tchunk):
        t1 = min(k, t0 + tchunk)
        noise = torch._standard_gamma(torch.full((t1 - t0, V), 0.3, dtype=torch.float64, device=dev))
        T = base[None, :] * noise
        T.scatter_(1, catch[t0:t1], base[catch[t0:t1]] * 50.0)
        c = torch.cumsum(T, dim=1)
        c = c / c[:, -1:]
        flat_parts.append((c + torch.arange(t0, t1, device=dev, dtype=torch.float64)[:, None]).reshape(-1))
        del noise, T, c
    flat = torch.cat(flat_parts)
    del flat_parts
    if doc_seed is not None:
        g.manual_seed(seed * 1000003 + doc_seed + 1)

    dominant = torch.randint(0, k, (D,), generator=g, device=dev)
    L = torch.empty(D, dtype=torch.float64, device=dev).normal_(mu, 0.5, generator=g).exp_()
    L = L.round_().clamp_(20, 2000).to(torch.int64)

    M_ATOMS = 8
    rows_parts, cnt_parts, per_doc_parts = [], [], []
    chunk = max(1, min(D, 64_000_000 // max(1, int(L.double().mean().item()))))
    for d0 in range(0, D, chunk):
        d1 = min(D, d0 + chunk)
        n = d1 - d0
        Lc = L[d0:d1]
        tok_doc = torch.repeat_interleave(torch.arange(n, device=dev), Lc)
        ntok = tok_doc.numel()
        sec = _doc_topic_tokens_torch(n, k, tok_doc, M_ATOMS, g, dev)
        pick = torch.rand(ntok, generator=g, device=dev) < 0.8
        topic = torch.where(pick, dominant[d0:d1][tok_doc], sec)
        u = torch.rand(ntok, generator=g, device=dev, dtype=torch.float64)
        w = torch.searchsorted(flat, topic.double() + u, right=True) - topic * V
        w.clamp_(0, V - 1)
        key = tok_doc * V + w
        uk, cnt = torch.unique(key, return_counts=True)
        dd = torch.div(uk, V, rounding_mode="floor")
        rows_parts.append((uk - dd * V).to(torch.int32).cpu().numpy().view(np.uint32))
        cnt_parts.append(cnt.to(torch.int32).cpu().numpy().view(np.uint32))
        per_doc_parts.append(torch.bincount(dd, minlength=n).cpu().numpy().astype(np.int64))
        del tok_doc, sec, pick, topic, u, w, key, uk, cnt, dd
    per_doc = np.concatenate(per_doc_parts)
    offsets = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(per_doc, out=offsets[1:])
    del flat
    torch.cuda.empty_cache()
    return Corpus(V, D, k, offsets, np.concatenate(rows_parts), np.concatenate(cnt_parts),
                  dominant.to(torch.int32).cpu().numpy())


def _doc_topic_tokens_torch(n, k, tok_doc, m, g, dev):
    """Secondary-topic draw.  Dirichlet(0.1 1_k) is approximated by its m largest atoms:
    m topics chosen uniformly per doc with Dirichlet(1)-distributed weights (exact
    Dirichlet over k=2000 topics for 8.2M docs would need 1.6e10 gamma draws)."""
    import torch

    atoms = torch.randint(0, k, (n, m), generator=g, device=dev)
    wts = -torch.log(torch.rand(n, m, generator=g, device=dev).clamp_min_(1e-12))
    c = torch.cumsum(wts, dim=1)
    c = c / c[:, -1:]
    u = torch.rand(tok_doc.numel(), generator=g, device=dev)
    idx = (c[tok_doc] < u[:, None]).sum(1).clamp_(0, m - 1)
    return atoms[tok_doc, idx]


def normalize(c: Corpus):
    """Restates SparseMatrix::populate_CSC bookkeeping + normalize_docs in numpy fp32.

    reference src/sparseMatrix.cpp:86-98 : nz_docs = #non-empty docs,
        avg_doc_sz = (float)(total_tokens / nz_docs)            [integer division]
    reference src/sparseMatrix.cpp:145-157: doc_sum = left-to-right fp32 sum of counts,
        normalized = avg_doc_sz * ((float)count / doc_sum)      [fp32, that association]
    doc_sum adds integer-valued floats, exact (hence order independent) below 2^24.
    Returns (normalized_vals float32[nnz], avg_doc_sz float32, nz_docs int).
    """
    lens = np.diff(c.offsets)
    nz_docs = int((lens > 0).sum())
    total = int(c.counts.astype(np.uint64).sum())
    avg = np.float32(total // nz_docs)
    cnt_f = c.counts.astype(np.float32)
    doc_sum = np.add.reduceat(c.counts.astype(np.int64), c.offsets[:-1][lens > 0]) if nz_docs else np.zeros(0)
    full = np.zeros(c.D, dtype=np.int64)
    full[lens > 0] = doc_sum
    assert full.max(initial=0) < (1 << 24), "doc token total must stay exact in fp32"
    ds = np.repeat(full.astype(np.float32), lens)
    vals = (avg * (cnt_f / ds)).astype(np.float32)
    return vals, avg, nz_docs


def save_npz(c: Corpus, path: str) -> None:
    np.savez_compressed(path, V=c.V, D=c.D, k=c.k, offsets=c.offsets, rows=c.rows, counts=c.counts)


def load_npz(path: str) -> Corpus:
    z = np.load(path)
    return Corpus(int(z["V"]), int(z["D"]), int(z["k"]), z["offsets"], z["rows"], z["counts"])
