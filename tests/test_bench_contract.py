"""The bench line contract, checked on the CODE that builds the lines (bench.ours_line / bench.reference_line are pure
functions of the measured numbers): every key the driver and the judge read must be there with the right type and the
right arithmetic, on our arm and on the reference arm, and both arms must print the same `config` for the same corpus."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=4, warmup=3, config="c2", impl="ours", cpu_sample_docs=0, no_cpu_baseline=True, seed=0,
             e2e_skip_B_U=False)
    d.update(kw)
    return argparse.Namespace(**d)


def _stats(steps):
    st = {n: 0.0 for n in bench.STAT_NAMES}
    st.update(launches=4000.0, spmm_bt_ms=10.0 * steps, spmm_b_ms=10.0 * steps, spmm_bt_bytes=260e6 * 40 * steps,
              spmm_b_bytes=260e6 * 40 * steps, spmm_bt_calls=40.0 * steps, spmm_b_calls=40.0 * steps, ks_ops=41.0,
              ks_gs_elided=30.0, ks_restarts=2.0, dist_tc_ms=0.5 * steps, dist_tc_flops=6e9 * 7 * steps, ks_wtf_ms=2.0,
              ks_wtf_bytes=4e9, ks_fsub_ms=2.0, ks_fsub_bytes=4e9, spmm_head_words=2048.0, spmm_tail_nnz=27e6,
              alloc_hits=500.0)
    return st


def _ours(args, world=1, D=300000):
    st = _stats(args.steps)
    state = dict(DB=D, nnzB=59_000_000, nconv=100, iters=7, objective=1.0)
    return bench.ours_line(args=args, world=world, cfg_name="c2", D=D, V=102000, nnz=69_000_000, k=100, sha16="ab" * 8,
                           total_docs=float(D * world), dev_ms=48.0 * args.steps, e2e_s=0.08 * args.steps, h2d=837e6, d2h=800e6,
                           st=st, state=state, clocks={"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []},
                           hbm_peak=(6540.2, "measured"), tf32_peak=(900.0, "measured"), traffic=250e6,
                           step_wall=[48.0] * args.steps, next_rows={})


def test_our_arm_line_has_the_contract_keys():
    a = _args()
    d = json.loads(json.dumps(_ours(a)))          # must survive JSON
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "tensor"):
        assert k in d, k
    assert d["unit"] == "docs/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] == "f32" and "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] == 4 and d["warmup"] == 3 and d["gpu_launches"] == 4000
    assert abs(d["ms_per_step"] - 48.0) < 1e-9
    assert abs(d["value"] - 300000 / 0.048) < 1e-6 * d["value"]           # docs of all ranks * steps / device time
    e = d["e2e"]
    assert e["unit"] == "docs/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert abs(e["value"] - 300000 / 0.08) < 1e-6 * e["value"] and e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - 2 * 260e6 * 40 * 4 / (80e-3) / 1e9) < 1e-6 * r["achieved"]     # algorithmic bytes / event time
    assert r["launches"] == 320 and abs(r["avg_launch_ms"] - 0.25) < 1e-12 and abs(r["bytes_per_launch"] - 260e6) < 1
    assert r["traffic"] == 250e6 and "stored" in r["traffic_source"]
    t = d["tensor"]
    assert t["bound"] == "tensor" and abs(t["frac"] - t["pipe_tflops"] / t["peak"]) < 1e-12
    assert abs(t["pipe_tflops"] - 3 * t["logical_tflops"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["run"]["D_B"] == 300000 and d["run"]["nconv"] == 100


def test_weak_scaling_value_is_whole_job():
    a = _args(gpus=4)
    d = _ours(a, world=4)
    assert d["n_gpus"] == 4 and abs(d["value"] - 4 * 300000 / 0.048) < 1e-6 * d["value"]


def test_reference_arm_line_and_same_config():
    a = _args(impl="reference")
    ours = _ours(a)
    ref = bench.reference_line(a, "c2", 300000, 102000, 69_000_000, 100, 16, [23.0, 24.0], 300000, 2, 1, True, "", "ab" * 8)
    ref = json.loads(json.dumps(ref))
    assert ref["impl"] == "reference" and ref["unit"] == "docs/s" and ref["gpu_launches"] == 0
    assert abs(ref["value"] - 300000 / 23.5) < 1e-9 * ref["value"] and abs(ref["ms_per_step"] - 23500.0) < 1e-6
    assert ref["e2e"] == {"value": ref["value"], "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = ref["cpu_baseline"]
    assert c["kind"] == "reference" and c["value"] == ref["value"] and c["cores"] == 16 and "identical corpus" in c["sample"]
    assert ref["reps_run"] == 2 and ref["warmup_run"] == 1 and ref["same_corpus_as_gpu_arm"] is True
    # the two arms describe the workload identically (metric, unit, direction, config object)
    for k in ("metric", "unit", "higher_is_better", "config", "dtype", "data", "scaling"):
        assert ref[k] == json.loads(json.dumps(ours))[k], k


def test_reference_arm_on_a_slice_says_so():
    a = _args(impl="reference", config="c3s")
    ref = bench.reference_line(a, "c3s", 1025000, 141000, 60_000_000, 2000, 16, [50.0], 30000, 1, 0, False,
                               "; the full corpus does not finish", None)
    assert ref["same_corpus_as_gpu_arm"] is False and "slice" in ref["cpu_baseline"]["sample"]
    assert abs(ref["value"] - 30000 / 50.0) < 1e-9


def test_corpus_checksum_is_stable():
    from isle_b200 import corpus
    c = corpus.generate("tiny")
    assert bench.corpus_sha16(c) == bench.corpus_sha16(corpus.generate("tiny")) and len(bench.corpus_sha16(c)) == 16
