#!/usr/bin/env python
"""bench.py -- ISLETrain spectral core (threshold -> block-KS SVD -> k-means on the projection)
docs/sec on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl ours|reference]

A "step" is one pass of stages A-E of ISLETrainer::train() (reference src/trainer.cpp:430-554)
over one synthetic corpus.  At N=1 the corpus is BASELINE.json configs[1] (NYTimes-shaped:
300k docs x 102k vocab, ~70M nnz, k=100); at N>1 every rank holds one such shard of documents
(weak scaling, vocabulary and k shared).

  value     docs/sec with the normalised CSC of A already resident in HBM, timed with CUDA
            events on the library's own stream, max over ranks
  e2e       the same through the reference-facing C ABI with HOST buffers: upload of A
            (pinned host memory, u64 row ids as the reference holds them) and download of
            zetas, original_cols, eigenvalues and the lifted centers inside the timed region
  roofline  the B*B^T*X SpMM passes: algorithmic bytes / CUDA-event time, against
            the measured HBM copy bandwidth in MEASURED_PEAKS.json
  tensor    the one dense contraction of the path (docs x centers distances on tcgen05, split TF32):
            logical and tensor-pipe TFLOP/s of the Lloyd assignment passes against the TF32 peak
  cpu_baseline  the UNMODIFIED reference C++ (oracle/_ref/ref_dump: reference sources over
            OpenBLAS + the MKL shim, not Intel MKL) on the host cores, on a bounded document
            slice of the same corpus
  next_rows the stages widened from the spectral core (Lloyd on the full-dimensional B,
            catchword thresholds + catchwords, topic model), timed beside the metric, not part of it
  panel_gbs, stage_ms_per_step   where the step's time goes

--impl reference times only that CPU arm and prints the same line shape.  --config c3s runs one of the
eight document shards of the PubMed-shaped c3 (k = 2000), the per-GPU workload of the 8 x B200 target.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ISLETrain spectral-core docs/sec (threshold -> block-KS SVD -> k-means on projection)"
UNIT = "docs/s"
SHAPES = {"c1": "small synthetic", "c2": "NYTimes-shaped synthetic", "c3": "PubMed-shaped synthetic",
          "c3s": "PubMed-shaped synthetic, one of 8 document shards", "c4": "Wikipedia-shaped synthetic"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-docs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--opt", nargs="*", default=[], help="library options name=value (isle_cuda_set_option), for A/B runs")
    ap.add_argument("--e2e-skip-B-U", dest="e2e_skip_B_U", action="store_true",
                    help="e2e leg without the B / U downloads the reference-side shim makes (round-1 behaviour)")
    return ap.parse_args()


# ------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler process is
    started before the warm-up (its start-up initialises NVML on every GPU of the box, which stalls
    CUDA calls of running processes for ~100 ms when other GPUs are idle); mark()/stop() keep only the
    samples whose timestamps fall inside the timed region."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        if os.environ.get("ISLE_BENCH_NO_SAMPLER"):
            return
        try:
            fd, self.path = tempfile.mkstemp(prefix="isle_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            # nvidia-smi's start-up (NVML init on every GPU of the box) stalls CUDA calls of running
            # processes for up to ~0.7 s: wait for its first sample so that never lands in a timed step
            t_end = time.time() + 15.0
            while time.time() < t_end and os.path.getsize(self.path) == 0 and self.proc.poll() is None:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, mx, reasons, all_sm, all_mx = [], [], set(), [], []
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    ts = None
                all_sm.append(float(f[2])); all_mx.append(float(f[3]))
                if ts is None or self.t0 is None or not (self.t0 - 0.05 <= ts <= self.t1 + 0.05):
                    continue
                sm.append(float(f[2])); mx.append(float(f[3]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                       "samples": len(sm)}
            elif all_sm:   # timed region shorter than the sampling period: nearest samples of the run
                out = {"sm_mhz": float(np.median(all_sm[-5:])), "sm_max_mhz": float(max(all_mx)), "reasons": [],
                       "samples": 0, "note": "no sample fell inside the timed region; last samples of the run"}
            os.unlink(self.path)
        except Exception:
            pass
        return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_tf32_peak():
    """Fallback dense TF32 tensor-pipe peak in TFLOP/s when it cannot be measured in the run: half the measured bf16
    burst figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if d.get("bf16_tflops"):
            return 0.5 * float(d["bf16_tflops"]), "measured bf16 burst / 2 (MEASURED_PEAKS.json)"
    return 0.5 * 2250.0, "nominal bf16 / 2"


def measure_tf32_peak():
    """Dense TF32 peak measured the way MEASURED_PEAKS.json measured bf16: torch.matmul fp32 with allow_tf32 on
    8192^3, best of 10, CUDA events (library GEMM, outside every timed region)."""
    try:
        import torch
        n = 8192
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device="cuda", dtype=torch.float32)
        b = torch.randn(n, n, device="cuda", dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "measured in this run: torch.matmul fp32 allow_tf32 8192^3, best of 10 (burst)"
    except Exception:
        return load_tf32_peak()


def corpus_sha16(c):
    import hashlib
    h = hashlib.sha256()
    for a in (c.offsets, c.rows, c.counts):
        h.update(np.ascontiguousarray(a).view(np.uint8).data)
    return h.hexdigest()[:16]


def workload_config(name, D, V, nnz, k, sha16=None):
    """The `config` object both arms print (identical for the same corpus): what is computed, nothing about how."""
    return {"workload": f"{name} {SHAPES.get(name, 'synthetic')} per GPU: {D} docs x {V} vocab, {nnz} nnz, k={k}",
            "block_size": 10, "tol": 1e-4, "l2": "inputs (A: %.0f MB) larger than the 126 MB L2" % (nnz * 8 / 1e6),
            "corpus_sha16": sha16}


def load_traffic(config):
    """DRAM bytes per SpMM pass from the committed ncu --set full capture (profiles/spmm_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("workload") == config:
            return 0.5 * (float(d["pass1"]) + float(d["pass2"]))
    except Exception:
        pass
    return None


def run_ref_dump(c, k, nthreads, workdir):
    """Times the unmodified reference (oracle/_ref/ref_dump) on corpus `c`; returns stage seconds."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/ref_dump missing (built by __graft_entry__.build() in the container)")
    path = os.path.join(workdir, "corpus.bin")
    c.write_bin(path)
    env = dict(os.environ, OMP_THREAD_LIMIT=str(nthreads), OMP_NUM_THREADS=str(nthreads),
               OPENBLAS_NUM_THREADS=str(nthreads))
    t0 = time.perf_counter()
    subprocess.run([exe, path, workdir, str(k), "--upto", "E"], check=True, env=env, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    meta = json.load(open(os.path.join(workdir, "meta.json")))
    core = sum(meta[x] for x in ("t_thresholds", "t_build_B", "t_block_ks", "t_kmeanspp", "t_lloyd"))
    return core, wall, meta


def slice_corpus(c, ndocs):
    from isle_b200.corpus import Corpus
    ndocs = min(ndocs, c.D)
    e = int(c.offsets[ndocs])
    return Corpus(c.V, ndocs, c.k, c.offsets[: ndocs + 1].copy(), c.rows[:e].copy(), c.counts[:e].copy())


def make_corpus(name, rank, seed):
    import torch
    from isle_b200 import corpus
    cfg = dict(corpus.CONFIGS[name])
    backend = "torch" if (torch.cuda.is_available() and cfg["D"] * 50 > 2_000_000) else "numpy"
    # every rank draws its own documents from the SAME topic model (one corpus, sharded by documents)
    return corpus.generate(V=cfg["V"], D=cfg["D"], k=cfg["k"], mu=cfg["mu"], seed=cfg["seed"] + seed,
                           doc_seed=rank if rank else None,
                           backend=backend, device=f"cuda:{torch.cuda.current_device()}" if backend == "torch" else "cpu")


def pinned(a: np.ndarray) -> np.ndarray:
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    try:
        return t.pin_memory().numpy()
    except Exception:
        return t.numpy()


# ------------------------------------------------------------------------------- reference arm
REF_WALL_BUDGET_S = 150.0        # wall-clock budget of the reference arm's repetitions (corpus generation excluded)
REF_MAX_FULL_NNZ = 120_000_000   # corpora above this (the k = 2000 shapes) cannot finish on the host cores in minutes


def reference_line(args, cfg_name, D, V, nnz, k, ncores, times_s, docs_timed, reps_run, warmup_run, same_corpus, note, sha16=None):
    """The JSON line of the reference arm (pure: tested on CPU)."""
    t = float(np.mean(times_s))
    val = docs_timed / t
    sample = (f"{'the identical corpus of the GPU arm' if same_corpus else 'a document slice of the GPU arm corpus'}: "
              f"{docs_timed} docs, V={V}, k={k}, stages A-E (thresholds, B, block-KS, k-means++, Lloyd on the projection), "
              f"{reps_run} timed repetition(s) + {warmup_run} warm-up within a {REF_WALL_BUDGET_S:.0f} s wall budget; "
              f"unmodified reference C++ over OpenBLAS + MKL shim (not Intel MKL); stage seconds from ref_dump's own "
              f"stopwatch around the same calls the reference's timerLog.txt phases bracket{note}")
    return {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(cfg_name, D, V, nnz, k, sha16),
        "same_corpus_as_gpu_arm": bool(same_corpus), "reps_run": reps_run, "warmup_run": warmup_run,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def bench_reference(args):
    """Times the reference's own CPU path on the corpus the GPU arm uses: the same generator call (make_corpus with
    rank 0's document seed), hence the identical (D, V, nnz, seed) object on the same box.  Repetitions are
    whole runs of the spectral core; as many as fit the wall budget (at least one)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from isle_b200 import corpus
    ncores = os.cpu_count() or 1
    cfg = corpus.CONFIGS[args.config]
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        pass
    c = make_corpus(args.config, 0, args.seed)
    D_full, nnz_full = c.D, c.nnz
    sha16 = corpus_sha16(c)
    same = True
    note = ""
    if args.cpu_sample_docs or nnz_full > REF_MAX_FULL_NNZ or cfg["k"] > 500:
        nd = args.cpu_sample_docs or max(4 * cfg["k"], min(c.D, 30000))
        c = slice_corpus(c, nd)
        same = c.D == D_full
        note = "" if same else f"; the full corpus ({D_full} docs, k={cfg['k']}) does not finish on the host cores in minutes"
    if args.gpus > 1:
        note += f"; N={args.gpus}: one rank's shard timed, docs/s of the host is what is reported"
    times, warm = [], 0
    t_begin = time.perf_counter()
    with tempfile.TemporaryDirectory(prefix="isle_ref_") as wd:
        core, wall, meta = run_ref_dump(c, cfg["k"], ncores, wd)
        first_wall = wall
        # the first run counts as warm-up only when another one fits the budget
        if args.warmup > 0 and (time.perf_counter() - t_begin) + first_wall <= REF_WALL_BUDGET_S:
            warm = 1
        else:
            times.append(core)
        while len(times) < max(args.steps, 1) and ((time.perf_counter() - t_begin) + first_wall <= REF_WALL_BUDGET_S or not times):
            core, wall, meta = run_ref_dump(c, cfg["k"], ncores, wd)
            times.append(core)
    line = reference_line(args, args.config, D_full, cfg["V"], nnz_full, cfg["k"], ncores, times, c.D, len(times), warm, same, note, sha16)
    emit(json.dumps(line))


# ------------------------------------------------------------------------------- our arm
STAT_NAMES = ("launches", "spmm_bt_ms", "spmm_bt_bytes", "spmm_bt_calls", "spmm_b_ms", "spmm_b_bytes", "spmm_b_calls",
              "ks_op_ms", "ks_orth_ms", "ks_qr_ms", "ks_truncate_ms", "ks_restarts", "ks_gs_elided", "ks_ops", "project_ms",
              "lloyd_iter_ms", "pp_round_ms", "thr_hist_ms", "thr_zeta_ms", "b_count_ms", "b_compact_ms", "csr_build_ms",
              "dist_tc_ms", "dist_simt_ms", "dist_tc_flops", "dist_simt_flops", "lloyd_accum_ms", "ks_wtf_ms", "ks_wtf_bytes",
              "ks_wtfred_ms", "ks_fsub_ms", "ks_fsub_bytes", "pp_dist_tc_ms", "pp_dist_skinny_ms", "pp_dist_simt_ms",
              "split_p_ms", "spmm_head1_ms", "spmm_tail1_ms", "spmm_head2_ms", "spmm_tail2_ms", "spmm_head_words",
              "spmm_tail_nnz", "alloc_misses", "alloc_hits", "allreduce_ms", "allgather_ms", "lift_ms", "ks_evd_ms", "ks_row_sharded",
              "allreduce_calls", "allgather_calls", "unpack_allreduce_ms", "unpack_allreduce_calls", "gemm_3xtf32_ms")


def ours_line(*, args, world, cfg_name, D, V, nnz, k, sha16, total_docs, dev_ms, e2e_s, h2d, d2h, st, state, clocks,
              hbm_peak, tf32_peak, traffic, step_wall, next_rows):
    """The JSON line of our arm from the measured numbers (pure: tested on CPU with synthetic stats)."""
    peak, peak_src = hbm_peak
    spmm_ms = st["spmm_bt_ms"] + st["spmm_b_ms"]
    spmm_bytes = st["spmm_bt_bytes"] + st["spmm_b_bytes"]
    ach = spmm_bytes / (spmm_ms * 1e-3) / 1e9 if spmm_ms > 0 else 0.0
    ncalls = st["spmm_bt_calls"] + st["spmm_b_calls"]
    tms, tfl = st["dist_tc_ms"], st["dist_tc_flops"]
    pipe = 3.0 * tfl / (tms * 1e-3) / 1e12 if tms else None
    live = 1.0 - (st["ks_gs_elided"] * args.steps) / max(3.0 * st["ks_ops"], 1.0)
    return {
        "metric": METRIC, "value": total_docs * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg_name, D, V, nnz, k, sha16),
        # what the run did with that workload (kept out of `config`, which both arms print identically)
        "run": {"D_B": state["DB"], "nnz_B": state["nnzB"], "ks_restarts": st["ks_restarts"],
                "ks_block_steps": st["ks_ops"], "ks_gs_third_passes_elided": st["ks_gs_elided"],
                "spmm_head_words": int(st["spmm_head_words"]), "spmm_tail_nnz": int(st["spmm_tail_nnz"]),
                "lloyd_iters": state["iters"], "nconv": state["nconv"], "ks_row_sharded": bool(st.get("ks_row_sharded", 0.0)),
                # collectives carried by the library's own kernels over peer memory (coll.cu) instead of NCCL, per step
                "p2p_collectives_per_step": state.get("p2p_collectives", 0.0) / max(args.steps, 1),
                "options": list(getattr(args, "opt", []) or [])},
        # the e2e leg moves what the reference-side shim moves: A up (u64 row ids), zetas, all of B (vals, u64 rows,
        # offsets, original_cols), eigenvalues, U, seeds, projected centers and the lifted centers down
        "e2e": {"value": (total_docs * args.steps / e2e_s) if e2e_s and e2e_s != float("inf") else None, "unit": UNIT,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(st["launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "B^T X and B Y passes of the operator (spmm_head_i8_kernel: dense head on tcgen05 "
                                               "kind::i8, beside it spmm_gather_bfp_kernel: sparse tail; one pass = one launch of each)",
                     "binds": "L2 -> SM fabric (one 32-byte operand row per nonzero: ~1.2 GB per pass at 7-8 TB/s), not HBM: "
                              "profiles/r2_spmm_ncu_full.md",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                     "traffic": traffic, "traffic_source": "stored ncu --set full capture (profiles/spmm_traffic.json), not measured in this run" if traffic else None,
                     "peak_source": peak_src, "launches": int(ncalls),
                     "avg_launch_ms": spmm_ms / ncalls if ncalls else None,
                     "bytes_per_launch": spmm_bytes / ncalls if ncalls else None},
        # the one dense contraction of the path (docs x centers distances, Lloyd assignment passes): logical flops
        # 2 D_B k k per pass; the tensor pipe does 3x that (split TF32: hi*hi + hi*lo + lo*hi)
        "tensor": {"kernel": "dist_tc_kernel (tcgen05 kind::tf32, split TF32)", "bound": "tensor",
                   "logical_tflops": tfl / (tms * 1e-3) / 1e12 if tms else None, "pipe_tflops": pipe,
                   "peak": tf32_peak[0], "unit": "TFLOP/s", "peak_source": tf32_peak[1],
                   "frac": pipe / tf32_peak[0] if pipe else None},
        "stage_ms_per_step": {n[:-3]: st[n] / args.steps for n in st if n.endswith("_ms")},
        # block Gram-Schmidt panel products: algorithmic bytes (n x rows x 4 per product) of the passes that ran
        # (elided third passes return at once and move nothing) / CUDA-event time, reduce kernels included
        "panel_gbs": {"wtf": live * st["ks_wtf_bytes"] / st["ks_wtf_ms"] / 1e6 if st["ks_wtf_ms"] else None,
                      "fsub": live * st["ks_fsub_bytes"] / st["ks_fsub_ms"] / 1e6 if st["ks_fsub_ms"] else None,
                      "passes_run_fraction": live},
        "step_wall_ms": [round(x, 2) for x in step_wall],
        "next_rows": next_rows,
        "alloc": {"driver_allocs_in_timed_region": st["alloc_misses"], "cache_hits_per_step": st["alloc_hits"] / args.steps},
    }


def bench_ours(args):
    import torch
    import torch.distributed as dist

    from isle_b200 import _capi, corpus, sharding
    from isle_b200._capi import ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(_capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    ctx = _capi.Context(local, rank, world, nccl_id)
    for o in args.opt:
        name, val = o.split("=")
        ctx.set_option(name, int(val))

    cfg = corpus.CONFIGS[args.config]
    V, k = cfg["V"], cfg["k"]
    c = make_corpus(args.config, rank, args.seed)
    D, nnz = c.D, c.nnz
    sha16 = corpus_sha16(c) if rank == 0 else None
    tf32_peak = measure_tf32_peak() if rank == 0 else (None, None)
    # input normalisation = ISLETrainer's own ingest (populate_CSC + normalize_docs), outside the path;
    # avg_doc_sz is a statistic of the whole corpus, so shards exchange token / document totals
    avg, nz_local, _nz_global = sharding.global_doc_stats(c.counts, c.offsets)
    vals = sharding.normalize_shard(c.counts, c.offsets, avg)
    h_vals = pinned(vals)
    h_rows64 = pinned(c.rows.astype(np.uint64))
    h_offs = pinned(c.offsets.astype(np.int64))
    torch.cuda.empty_cache()

    zetas = pinned(np.zeros(V, np.float32))
    evalues = np.zeros(k, np.float32)
    seeds = np.zeros(k, np.uint64)
    centers_lowd = np.zeros((k, k), np.float32)
    centers = pinned(np.zeros((k, V), np.float32))
    # host arrays the reference-side shim fills every run (threshold_and_copy: allocate(nnzs + 1000), U_colmajor)
    full_dl = not args.e2e_skip_B_U
    h_bvals = pinned(np.zeros(nnz + 1000, np.float32)) if full_dl else None
    h_brows = pinned(np.zeros(nnz + 1000, np.uint64)) if full_dl else None
    h_boffs = pinned(np.zeros(D + 1, np.int64)) if full_dl else None
    h_U = pinned(np.zeros((k, V), np.float32)) if full_dl else None
    state = {}

    def upload():
        ctx.call("isle_cuda_upload_A", V, D, nnz, ptr(h_vals), ptr(h_rows64), ptr(h_offs), C.c_float(float(avg)), nz_local)

    def core(host_outputs: bool, step_seed: int):
        nn, nnzB, DB, nconv = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_int()
        res, obj, iters = C.c_float(), C.c_double(), C.c_int()
        ctx.call("isle_cuda_thresholds", k, ptr(zetas) if host_outputs else None, C.byref(nn))
        ctx.call("isle_cuda_build_B", None, C.byref(nnzB), C.byref(DB))
        if host_outputs:
            oc = np.zeros(int(DB.value), np.uint64)
            ctx.call("isle_cuda_download_B", None, None, None, ptr(oc))
            if full_dl:     # as the shim does: B's bulk travels in the background while the eigensolver runs
                ctx.call("isle_cuda_download_B_begin", ptr(h_bvals), ptr(h_brows), ptr(h_boffs), None)
        ctx.call("isle_cuda_block_ks", k, 10, 100, C.c_float(1e-4), step_seed, ptr(evalues),
                 ptr(h_U) if (host_outputs and full_dl) else None, C.byref(nconv))
        if host_outputs and full_dl:
            ctx.call("isle_cuda_download_B_end")
        ctx.call("isle_cuda_kmeanspp", k, step_seed, ptr(seeds), ptr(centers_lowd), C.byref(res))
        ctx.call("isle_cuda_lloyd_projected", k, ptr(centers_lowd), 10, None, C.byref(obj), C.byref(iters))
        ctx.call("isle_cuda_lift_centers", k, ptr(centers_lowd), k, ptr(centers) if host_outputs else None)
        ctx.call("isle_cuda_cleanup_eigensolver")
        state.update(nnzB=int(nnzB.value), DB=int(DB.value), nconv=int(nconv.value), iters=int(iters.value),
                     objective=float(obj.value))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    barrier()       # ranks leave corpus generation (and rank 0 its TF32 peak measurement) seconds apart: start the library together
    # ---- warm-up: full end-to-end steps
    for i in range(args.warmup):
        upload()
        core(True, args.seed + i)

    # ---- (1) device-resident: A stays in HBM, CUDA events on the library stream
    upload()
    ctx.call("isle_cuda_reset_stats")
    ctx.call("isle_cuda_set_profiling", 1)
    barrier()
    sampler.mark_begin()
    ctx.call("isle_cuda_timer_start")
    step_wall = []
    for i in range(args.steps):
        tw = time.perf_counter()
        core(False, args.seed + 100 + i)
        step_wall.append((time.perf_counter() - tw) * 1e3)
    ms = C.c_double()
    ctx.call("isle_cuda_timer_stop", C.byref(ms))
    barrier()
    sampler.mark_end()
    dev_ms = ms.value
    st = {n: ctx.stat(n) for n in STAT_NAMES}
    try:
        state["p2p_collectives"] = ctx.stat("p2p_collectives")
    except Exception:       # single GPU, or the peer workspace could not be mapped: the counter does not exist
        state["p2p_collectives"] = 0.0
    # ---- SURVEY 8(f) row 1, reported beside the metric (not part of it): Lloyd on the full-dimensional B from
    # the lifted centers the last step left on the device (trainer.cpp:566)
    ctx.call("isle_cuda_reset_stats")
    fobj, fit, fms = C.c_double(), C.c_int(), C.c_double()
    ctx.call("isle_cuda_timer_start")
    f_assign = np.zeros(max(state["DB"], 1), np.uint32)
    ctx.call("isle_cuda_lloyd_full", k, None, 10, ptr(f_assign), C.byref(fobj), C.byref(fit))
    ctx.call("isle_cuda_timer_stop", C.byref(fms))
    stf = {n: ctx.stat(n) for n in ("lloyd_full_iter_ms", "lloyd_full_assign_ms", "lloyd_full_update_ms",
                                    "lloyd_full_assign_flops")}
    # ---- SURVEY 8(f) row 2, first half: catchword thresholds of all clusters + catchwords (trainer.cpp:577-639), from the
    # partition stage F just produced; single-GPU contexts only for now
    cw = None
    if world == 1:
        oc = np.zeros(max(state["DB"], 1), np.uint64)
        ctx.call("isle_cuda_download_B", None, None, None, ptr(oc))
        cl = np.full(D, 0xFFFFFFFF, np.uint32)
        cl[oc[:state["DB"]].astype(np.int64)] = f_assign[:state["DB"]]
        r_catch = int(np.floor((1.0 / 3.0) * 1.0 * float(np.float32(D)) / float(np.float32(2.0 * k))))   # trainer.cpp:583
        tw = np.zeros(V, np.int32)
        cms = C.c_double()
        ctx.call("isle_cuda_timer_start")
        ctx.call("isle_cuda_catchword_thresholds", k, r_catch, ptr(cl), None)
        ctx.call("isle_cuda_find_catchwords", k, None, C.c_double(1.1), ptr(tw))
        ctx.call("isle_cuda_timer_stop", C.byref(cms))
        cw = {"what": "SURVEY 8(f) row 2 first half: rth_highest_element for all clusters + find_catchwords, device-resident A, "
                      "not part of the metric", "ms": cms.value, "r": r_catch, "catchwords": int((tw >= 0).sum()),
              "candidate_segments": int(ctx.stat("cw_candidates")),
              "count_ms": ctx.stat("cw_count_ms"), "scatter_ms": ctx.stat("cw_scatter_ms"), "sort_ms": ctx.stat("cw_sort_ms"),
              "algorithmic_GBps": (2.0 * nnz * 8.0 + 2.0 * k * V * 4.0) / (cms.value * 1e-3) / 1e9 if cms.value > 0 else None}
        # ---- SURVEY 8(f) row 2, second half: construct_topic_model from those catchwords and that partition
        rank_tm = int(np.uint64(5.0 * 1.0 * float(np.float32(D)) / (float(np.float32(k)) * 2.0)))      # sparseMatrix.cpp:716
        n_tm, tms = C.c_uint64(), C.c_double()
        ctx.call("isle_cuda_timer_start")
        ctx.call("isle_cuda_construct_topic_model", k, ptr(tw), ptr(cl), rank_tm, None, C.byref(n_tm))
        ctx.call("isle_cuda_timer_stop", C.byref(tms))
        cw["topic_model"] = {"what": "construct_topic_model (sparseMatrix.cpp:597-838), model left on the device", "ms": tms.value,
                             "doc_topic_sums": int(n_tm.value), "rank_threshold": rank_tm}
    ctx.call("isle_cuda_set_profiling", 0)

    # ---- (2) end to end: host buffers in, host results out, every step
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0 if os.environ.get("ISLE_BENCH_SKIP_E2E") else args.steps      # profiler runs (ncu launch lists) skip this leg
    for i in range(e2e_steps):
        upload()
        core(True, args.seed + 200 + i)
    barrier()
    e2e_s = max(time.perf_counter() - t0, 1e-9) if e2e_steps else float("inf")
    clocks = sampler.stop()

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s, float(D)], dtype=torch.float64, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, total_docs = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        total_docs = float(D)

    if rank == 0:
        h2d = nnz * (4 + 8) + (D + 1) * 8
        d2h = V * 4 + state["DB"] * 8 + k * 4 + k * 8 + 2 * k * k * 4 + V * k * 4
        if full_dl:
            d2h += state["nnzB"] * (4 + 8) + (state["DB"] + 1) * 8 + V * k * 4
        next_rows = {"lloyd_full": {"what": "SURVEY 8(f) row 1: run_lloyds on the full-dimensional B (trainer.cpp:566), "
                                            "device-resident, not part of the metric", "ms": fms.value, "iters": fit.value,
                                    "objective": fobj.value, "assign_ms": stf["lloyd_full_assign_ms"],
                                    "update_ms": stf["lloyd_full_update_ms"],
                                    "assign_gather_tflops": (stf["lloyd_full_assign_flops"] / (stf["lloyd_full_assign_ms"] * 1e-3) / 1e12
                                                             if stf["lloyd_full_assign_ms"] > 0 else None)},
                     "catchwords": cw}
        line = ours_line(args=args, world=world, cfg_name=args.config, D=D, V=V, nnz=nnz, k=k, sha16=sha16,
                         total_docs=total_docs, dev_ms=dev_ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, st=st, state=state,
                         clocks=clocks, hbm_peak=load_peaks(), tf32_peak=tf32_peak, traffic=load_traffic(args.config),
                         step_wall=step_wall, next_rows=next_rows)
        if not args.no_cpu_baseline and world == 1:
            try:
                ncores = os.cpu_count() or 1
                full = not args.cpu_sample_docs and nnz <= REF_MAX_FULL_NNZ and k <= 500
                cs = c if full else slice_corpus(c, args.cpu_sample_docs or max(4 * k, 30000))
                with tempfile.TemporaryDirectory(prefix="isle_ref_") as wd:
                    core_s, wall, meta = run_ref_dump(cs, k, ncores, wd)
                line["cpu_baseline"] = {
                    "value": cs.D / core_s, "unit": UNIT, "cores": ncores, "kind": "reference",
                    "sample": (f"{'the identical corpus' if cs.D == D else 'the first ' + str(cs.D) + ' docs of the same corpus'} "
                               f"({cs.D} docs, {cs.nnz} nnz, V={V}, k={k}), one run of stages A-E in {core_s:.1f} s; unmodified "
                               f"reference C++ over OpenBLAS + MKL shim (not Intel MKL), ref_dump's own stopwatch"),
                    "stage_s": {x: meta[x] for x in ("t_thresholds", "t_build_B", "t_block_ks", "t_kmeanspp", "t_lloyd")}}
            except Exception as e:  # the baseline is reported, never allowed to sink the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {e}"}
        emit(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def keep_stdout_for_the_line():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout at
    init), so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(text: str) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


if __name__ == "__main__":
    a = parse()
    keep_stdout_for_the_line()
    if a.impl == "reference":
        bench_reference(a)
    else:
        bench_ours(a)
