"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_dump).

Runs only in the build container (needs /root/reference to have been compiled by
`make -C oracle`).  The GPU box and the CPU test-suite consume the committed .npz
files; nothing under tests/ reads /root/reference at run time.

    python tests/golden/make_golden.py

tiny : full arrays (corpus, zetas, B, evalues, U, seeds, Lloyd in/out)
c1   : BASELINE.json configs[0] shape (10k docs x 5k vocab, k=20): zetas, evalues,
       Lloyd in/out in full; corpus and B by SHA-256 digest (regenerated from the seed).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from isle_b200 import corpus  # noqa: E402


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run(name, extra=()):
    c = corpus.generate(name)
    tmp = tempfile.mkdtemp(prefix=f"isle_golden_{name}_")
    c.write_bin(os.path.join(tmp, "corpus.bin"))
    env = dict(os.environ, OMP_THREAD_LIMIT=str(os.cpu_count()))
    subprocess.run([os.path.join(ROOT, "oracle/_ref/ref_dump"), os.path.join(tmp, "corpus.bin"), tmp,
                    str(c.k), *extra], check=True, env=env, stdout=subprocess.DEVNULL)
    meta = json.load(open(os.path.join(tmp, "meta.json")))
    ld = lambda n, dt: np.fromfile(os.path.join(tmp, n + ".bin"), dtype=dt)
    out = dict(
        V=c.V, D=c.D, k=c.k, nnz=c.nnz,
        avg_doc_sz=ld("A_avg_doc_sz", np.float32)[0],
        new_nnzs=meta["new_nnzs"], D_B=meta["D_B"], nnz_B=meta["nnz_B"], frobenius=meta["frobenius"],
        zetas=ld("zetas", np.float32),
        evalues=ld("evalues", np.float32),
        U_colmajor=ld("U_colmajor", np.float32),
        seeds=ld("seeds", np.uint64),
        centers_lowd_init=ld("centers_lowd_init", np.float32),
        centers_lowd_final=ld("centers_lowd_final", np.float32),
        lloyd_assign=ld("lloyd_assign", np.uint32),
        corpus_sha=sha(c.offsets, c.rows, c.counts),
        A_vals_sha=sha(ld("A_normalized_vals", np.float32)),
        B_sha=sha(ld("B_vals", np.float32), ld("B_rows", np.uint64).astype(np.uint32),
                  ld("B_offsets", np.int64), ld("B_original_cols", np.uint64).astype(np.uint32)),
    )
    return c, tmp, out, ld


def stage_f(name):
    """Stage F (run_lloyds on the full-dimensional B) as a self-contained fixture:
    zetas (-> B), the lifted centers that went in, the centers and the partition that came out."""
    c, tmp, out, ld = run(name)
    return dict(zetas=out["zetas"], corpus_sha=out["corpus_sha"], B_sha=out["B_sha"],
                centers_in=ld("centers", np.float32), centers_out=ld("full_centers", np.float32),
                assign=ld("full_assign", np.uint32))


def stage_g(name):
    """Stage G (rth_highest_element per cluster + find_catchwords): the partition of the ORIGINAL
    documents that went in, the rank r, the k x V threshold matrix and the (topic, word) catchword pairs that came out."""
    c, tmp, out, ld = run(name)
    meta = json.load(open(os.path.join(tmp, "meta.json")))
    return dict(corpus_sha=out["corpus_sha"], A_vals_sha=out["A_vals_sha"], r=meta["catch_r"],
                cluster_of_doc=ld("catch_cluster_of_doc", np.uint32), thresholds=ld("catch_thresholds", np.float32),
                catchwords=ld("catchwords", np.uint32).reshape(-1, 2))


def stage_h(name):
    """Stage H (construct_topic_model): what went in (partition, catchwords) and what came out
    (Model V x k, the (doc, topic, sum) list, the top topic pairs)."""
    c, tmp, out, ld = run(name)
    return dict(corpus_sha=out["corpus_sha"], A_vals_sha=out["A_vals_sha"],
                cluster_of_doc=ld("catch_cluster_of_doc", np.uint32), catchwords=ld("catchwords", np.uint32).reshape(-1, 2),
                model=ld("model", np.float32), dts_doc=ld("dts_doc", np.uint32), dts_topic=ld("dts_topic", np.uint32),
                dts_val=ld("dts_val", np.float32), top_topic_pairs=ld("top_topic_pairs", np.uint32).reshape(-1, 3))


def main_stage_h():
    here = os.path.dirname(__file__)
    for name in ("tiny", "c1"):
        np.savez_compressed(os.path.join(here, f"{name}_stageH.npz"), **stage_h(name))


def main_stage_f():
    here = os.path.dirname(__file__)
    for name in ("tiny", "c1"):
        np.savez_compressed(os.path.join(here, f"{name}_stageF.npz"), **stage_f(name))


def main_stage_g():
    here = os.path.dirname(__file__)
    for name in ("tiny", "c1"):
        np.savez_compressed(os.path.join(here, f"{name}_stageG.npz"), **stage_g(name))

def main_c3m():
    """c3-shaped miniature at k = 320 (corpus.CONFIGS['c3m']): the large-k kernel modes (several center tiles and the
    clamped-min mode of the tcgen05 distance kernel, panel products over several K segments, Gram-Schmidt pass elision at
    ncv = 650, eig_sym of a 640 x 640 projected matrix) pinned to the reference's own output.  U is stored in full
    (V x k fp32): the Lloyd parity test installs exactly the projection the reference used."""
    here = os.path.dirname(__file__)
    c, tmp, out, ld = run("c3m")
    np.savez_compressed(os.path.join(here, "c3m.npz"), **out)


if __name__ == "__main__":
    if "--c3m" in sys.argv:
        main_c3m()
        sys.exit(0)
    if "--stage-f" in sys.argv:      # adds the stage-F fixtures without touching the stage A-E ones
        main_stage_f()
        sys.exit(0)
    if "--stage-g" in sys.argv:
        main_stage_g()
        sys.exit(0)
    if "--stage-h" in sys.argv:
        main_stage_h()
        sys.exit(0)
    here = os.path.dirname(__file__)
    c, tmp, out, ld = run("tiny")
    out.update(offsets=c.offsets, rows=c.rows, counts=c.counts,
               A_normalized_vals=ld("A_normalized_vals", np.float32),
               B_vals=ld("B_vals", np.float32), B_rows=ld("B_rows", np.uint64).astype(np.uint32),
               B_offsets=ld("B_offsets", np.int64),
               B_original_cols=ld("B_original_cols", np.uint64).astype(np.uint32))
    np.savez_compressed(os.path.join(here, "tiny.npz"), **out)

    # same corpus with an injected doc-selection mask (sampled_threshold_and_copy path)
    rng = np.random.Generator(np.random.PCG64(7))
    mask = (rng.random(c.D) < 0.4).astype(np.uint8)
    mpath = os.path.join(tmp, "mask.u8")
    mask.tofile(mpath)
    subprocess.run([os.path.join(ROOT, "oracle/_ref/ref_dump"), os.path.join(tmp, "corpus.bin"), tmp,
                    str(c.k), "--upto", "B", "--mask", mpath], check=True, stdout=subprocess.DEVNULL)
    np.savez_compressed(os.path.join(here, "tiny_masked.npz"), mask=mask,
                        B_vals=ld("B_vals", np.float32), B_rows=ld("B_rows", np.uint64).astype(np.uint32),
                        B_offsets=ld("B_offsets", np.int64),
                        B_original_cols=ld("B_original_cols", np.uint64).astype(np.uint32))

    c, tmp, out, ld = run("c1")
    np.savez_compressed(os.path.join(here, "c1.npz"), **out)

    main_stage_f()
    main_stage_g()
    main_stage_h()
    main_c3m()
