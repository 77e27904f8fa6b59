#!/usr/bin/env python
"""Times the real drop-in: the reference's unmodified CLI (drivers/ISLETrain.cpp + trainer.cpp) linked against the
all-CPU reference objects (oracle/_ref/ISLETrain_ref) and against the replacement translation unit + libisle_cuda.so
(oracle/_ref/ISLETrain_cuda), on the same input files, with the per-phase seconds taken from the reference's own
timerLog.txt (include/timer.h:72-86: the number tagged "(sys)" is the wall-clock delta, "(user)" the CPU time).

    python tools/dropin_timing.py [--docs 60000] [--config c2] [--ngpus 1 2 ...] > gpurun_out/dropin.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isle_b200 import corpus  # noqa: E402

CORE = ("Computing thresholds", "Creating thresholded and scaled matrix", "eigen solver init", "Spectra eigen solve",
        "K-means seeds initialization", "Converging LLoyds k-means on B_k")


def run(exe, wd, c, name, env=None):
    out = os.path.join(wd, name)
    os.makedirs(out)
    args = [exe, os.path.join(wd, "tdf.txt"), os.path.join(wd, "vocab.txt"), out, str(c.V), str(c.D), str(c.nnz), str(c.k),
            "0", "0", "0", "0", "0"]
    t0 = time.perf_counter()
    r = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        return {"error": (r.stdout[-500:] + r.stderr[-500:])}
    logdir = os.path.join(out, os.listdir(out)[0])
    phases = {}
    for line in open(os.path.join(logdir, "timerLog.txt")):
        m = re.match(r".*Time for (.+?)\.*(\d[0-9.eE+-]*)s\(user\)\s+([0-9.eE+-]+)s\(sys\)", line)
        if m:
            phases[m.group(1).rstrip(".")] = phases.get(m.group(1).rstrip("."), 0.0) + float(m.group(3))
    core = sum(v for k, v in phases.items() if any(k.startswith(p) for p in CORE))
    return {"wall_s": wall, "spectral_core_s": core, "phases_wall_s": phases}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--docs", type=int, default=60000)
    ap.add_argument("--ngpus", type=int, nargs="*", default=[1])
    a = ap.parse_args()
    cfg = corpus.CONFIGS[a.config]
    c = corpus.generate(V=cfg["V"], D=min(a.docs, cfg["D"]), k=cfg["k"], mu=cfg["mu"], seed=cfg["seed"])
    res = {"workload": f"{a.config} shape, first {c.D} docs x {c.V} vocab, {c.nnz} nnz, k={c.k}", "host_cores": os.cpu_count()}
    with tempfile.TemporaryDirectory(prefix="isle_dropin_") as wd:
        c.write_text(os.path.join(wd, "tdf.txt"), os.path.join(wd, "vocab.txt"))
        res["ISLETrain_ref"] = run(os.path.join(ROOT, "oracle/_ref/ISLETrain_ref"), wd, c, "ref")
        for n in a.ngpus:
            env = {"ISLE_CUDA_NGPUS": str(n)} if n > 1 else {}
            run(os.path.join(ROOT, "oracle/_ref/ISLETrain_cuda"), wd, c, f"warm{n}", env)        # CUDA context / module load warm-up
            res[f"ISLETrain_cuda_{n}gpu"] = run(os.path.join(ROOT, "oracle/_ref/ISLETrain_cuda"), wd, c, f"cuda{n}", env)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
