"""bench.py is the driver's contract: both arms must run (here: the CPU dry run on gloo with a tiny model), print ONE JSON
line with the agreed keys, and spell the benchmark configuration identically."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--device", "cpu", "--model", "bloom-tiny", "--seq-len", "32", "--batch-per-gpu", "2", "--steps", "2", "--warmup", "3"]


def _run(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + COMMON + extra, capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(lines[0])


def test_layout_and_config_helpers():
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    def args(**kw):
        base = dict(gpus=8, tp=0, pp=1, model="bloom-560m", batch_per_gpu=8, seq_len=1024, microbatches=8, experts=0)
        base.update(kw)
        return argparse.Namespace(**base)

    assert bench.layout_of(args(gpus=1)) == (1, 1, 1)
    assert bench.layout_of(args(gpus=2)) == (2, 1, 1)
    assert bench.layout_of(args(gpus=8)) == (2, 1, 4)                      # BASELINE.json's headline: TP2 x DP4
    assert bench.layout_of(args(gpus=8, tp=8)) == (8, 1, 1)                # config #3 / #4
    assert bench.layout_of(args(gpus=8, tp=2, pp=2)) == (2, 2, 2)          # config #5
    cfg = bench.config_of(args(gpus=8), 2, 1, 4)
    assert cfg["parallelism"] == "tp2dp4+zero1" and cfg["global_batch"] == 64 and cfg["model"] == "bloom-560m"
    assert bench.config_of(args(gpus=8, tp=8, experts=8), 8, 1, 1)["parallelism"] == "tp8dp1+moe8e"


@pytest.mark.parametrize("extra", [["--gpus", "1"], ["--gpus", "2"], ["--gpus", "2", "--hf"]])
def test_both_arms_print_one_json_line_with_the_same_config(extra):
    ours = _run(extra)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks"):
        assert key in ours, key
    assert ours["e2e"]["h2d_bytes_per_step"] > 0 and ours["e2e"]["d2h_bytes_per_step"] == 4
    if "--hf" in extra:
        return
    ref = _run(extra + ["--impl", "reference"])
    assert ref["impl"] == "reference"
    assert ref["config"] == ours["config"] and ref["metric"] == ours["metric"]
    if extra[-1] != "1":
        assert ours["numerics_ok"] is True      # the self-check ran (on CPU both engines are the library path)


def test_moe_config_both_arms_and_the_reference_arm_in_bf16():
    """BASELINE.json config #4 shape (Switch-MoE, experts sharded over the tensor group).  The reference arm runs with the
    model cast to bf16 as on the GPU: its router stays fp32 (it casts its own input), everything else must cope."""
    extra = ["--gpus", "2", "--tp", "2", "--experts", "2"]
    ours = _run(extra + ["--no-self-check"])
    ref = _run(extra + ["--impl", "reference"], env={"PIPEGOOSE_B200_BENCH_CPU_BF16": "1"})
    assert ref["impl"] == "reference" and ref["config"] == ours["config"]
    assert ours["config"]["parallelism"] == "tp2dp1+moe2e"
    for line in (ours, ref):
        assert line["final_loss"] == line["final_loss"] and 0 < line["final_loss"] < 20   # finite, of the order of ln(vocab)
