"""Compatibility probe: run the REFERENCE's own test-suite against ``pipegoose_b200``.

    python tools/run_reference_tests.py [--reference /root/reference] [--out profiles/reference_testsuite_compat.txt]

The reference's ``tests/`` tree is copied to a temporary directory (nothing of it enters this repository), every
``pipegoose`` import is rewritten to ``pipegoose_b200``, the Hugging Face hub is replaced by offline stand-ins (a tiny
random Bloom / GPT-2 and a whitespace tokenizer with Bloom's left padding — there is no network here), and pytest runs
the tree file by file (one pytest process each, hard time limit) on CPU / gloo.  The report lists what passes unmodified and the
reason for everything that does not.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HUB_STAND_IN = '''
"""Offline stand-ins for the Hugging Face hub (written by tools/run_reference_tests.py)."""
import hashlib

import torch
import transformers
from transformers import BatchEncoding, BloomConfig, BloomForCausalLM, GPT2Config, GPT2Model

VOCAB = 1024


def _tiny_bloom(*a, **k):
    torch.manual_seed(1234)
    return BloomForCausalLM(BloomConfig(vocab_size=VOCAB, hidden_size=64, n_layer=2, n_head=8))


def _tiny_gpt2(*a, **k):
    torch.manual_seed(1234)
    return GPT2Model(GPT2Config(vocab_size=VOCAB, n_embd=64, n_layer=2, n_head=4))


class _Tok:
    pad_token, eos_token, pad_token_id, eos_token_id, padding_side = "<pad>", "</s>", 3, 2, "left"

    def _ids(self, text):
        return [4 + int(hashlib.md5(w.encode()).hexdigest(), 16) % (VOCAB - 4) for w in text.split()] or [self.eos_token_id]

    def __call__(self, text, return_tensors=None, padding=False, truncation=False, max_length=None, **_):
        rows = [self._ids(t) for t in ([text] if isinstance(text, str) else list(text))]
        if max_length:
            rows = [r[:max_length] for r in rows]
        n = max(len(r) for r in rows)
        ids = [[self.pad_token_id] * (n - len(r)) + r for r in rows]
        mask = [[0] * (n - len(r)) + [1] * len(r) for r in rows]
        return BatchEncoding({"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(mask)})

    def decode(self, ids, **_):
        return " ".join(str(int(i)) for i in ids)

    def batch_decode(self, ids, **_):
        return [self.decode(r) for r in ids]


transformers.AutoModelForCausalLM.from_pretrained = staticmethod(_tiny_bloom)
transformers.BloomForCausalLM.from_pretrained = classmethod(lambda cls, *a, **k: _tiny_bloom())
transformers.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: _tiny_gpt2() if "gpt2" in name else _tiny_bloom().transformer)
transformers.AutoTokenizer.from_pretrained = staticmethod(lambda *a, **k: _Tok())
'''

# why a reference test is not expected to pass against this library: pattern on the test id -> reason
KNOWN = [
    (r"test_initialize_expert_parallel_group", "deliberate (MIGRATION Q3): EXPERT_DATA groups hold replicas of the SAME experts here; the "
                                                "reference's are the tensor-parallel rank sets"),
    (r"with_expert_capacity", "deliberate: an expert takes `capacity` tokens here; the reference keeps positions < capacity of a 1-based "
                              "count, i.e. capacity - 1 (off by one), and its test asserts the strict bound"),
    (r"test_hybrid\.py", "the reference's own test compares EVERY parameter with a dim-0 shard, LayerNorms (replicated in both libraries) "
                        "included: it cannot pass against the reference either (it is excluded from its CI)"),
    (r"pipeline_parallel/(job/|sync/|test_comm|test_pipeline_context|test_pipeline_parallel|test_worker)",
     "RPC-era job runtime mechanics (packages arrive through RPC callbacks, a clock thread drives the schedule generator); the runtime "
     "here moves packages with p2p and static tables — SURVEY §9 asks for no parity of these internals"),
]


DATASET_STAND_IN = '''

class _DS(list):
    def map(self, fn):
        return _DS([dict(r, **fn(r)) for r in self])


def load_dataset(name, split=None):
    import random

    random.seed(0)
    words = "the movie was a long slow beautiful mess of great acting and bad writing".split()
    return _DS([{"text": " ".join(random.choice(words) for _ in range(random.randint(3, 12))), "label": 0} for _ in range(64)])
'''


def run_example(reference: str):
    """The reference's user scripts against this package on CPU ranks: imports rewritten, "cuda" -> "cpu", one epoch, the
    hub / datasets / wandb stand-ins.

    * examples/hybrid_parallelism.py (TP2 x DP2, 🤗 Bloom, tokenizer with padding, stock SGD): through the class-swap path
      (what an fp32 🤗 model gets) and through the fused sequence-parallel path;
    * tests/convergence/run_ep.py (Switch-MoE next to the dense model, ExpertLoss): experts on one rank and sharded over two;
    * tests/convergence/run_hybrid_parallel.py (TP2 x DP2 + ZeRO-1 DistributedOptimizer next to a DDP replica of the model)."""
    tmp = tempfile.mkdtemp(prefix="pgb200_refscripts_")
    with open(os.path.join(tmp, "pgb200_hub_stand_in.py"), "w") as f:
        f.write(HUB_STAND_IN.replace("n_layer=2, n_head=8", "n_layer=4, n_head=8") + DATASET_STAND_IN
                + "\n_Tok.add_special_tokens = lambda self, d: 0\n")
    with open(os.path.join(tmp, "wandb.py"), "w") as f:
        f.write("def init(*a, **k): pass\ndef log(*a, **k): pass\ndef finish(*a, **k): pass\n")

    def prepare(path):
        src = re.sub(r"\bpipegoose\b", "pipegoose_b200", open(path).read())
        src = src.replace("from datasets import load_dataset", "from pgb200_hub_stand_in import load_dataset")
        for a, b in (('model.to("cuda")', 'model.to("cpu")'), ("ref_model.cuda()", "ref_model.cpu()"), ("tensor.cuda()", "tensor.cpu()"),
                     ("torch.cuda.manual_seed_all(seed)", "pass"), ("torch.cuda.empty_cache()", "pass"),
                     ("range(100)", "range(1)"), ("NUM_EPOCHS = 100", "NUM_EPOCHS = 1"), ("NUM_EPOCHS = 4", "NUM_EPOCHS = 1"),
                     (", device_ids=[device]", "")):
            src = src.replace(a, b)
        return "import pgb200_hub_stand_in  # noqa: F401\n" + src

    runs = []
    example = os.path.join(reference, "examples", "hybrid_parallelism.py")
    if os.path.exists(example):
        src = prepare(example)
        runs.append(("examples/hybrid_parallelism.py, class-swap path", src, 4, r"rank=0, loss=([0-9.]+)"))
        runs.append(("examples/hybrid_parallelism.py, fused sequence-parallel path",
                     src.replace("TensorParallel(model, parallel_context)", "TensorParallel(model, parallel_context, sequence_parallel=True)"),
                     4, r"rank=0, loss=([0-9.]+)"))
    run_ep = os.path.join(reference, "tests", "convergence", "run_ep.py")
    if os.path.exists(run_ep):
        src = prepare(run_ep)
        runs.append(("tests/convergence/run_ep.py, 4 experts on one rank", src, 1, r"rank=0, train_loss=([0-9.]+)"))
        runs.append(("tests/convergence/run_ep.py, 4 experts sharded over 2 ranks",
                     src.replace("TENSOR_PARALLEL_SIZE = 1", "TENSOR_PARALLEL_SIZE = 2"), 2, r"rank=0, train_loss=([0-9.]+)"))
    hybrid = os.path.join(reference, "tests", "convergence", "run_hybrid_parallel.py")
    if os.path.exists(hybrid):
        runs.append(("tests/convergence/run_hybrid_parallel.py (TP2 x DP2 + DistributedOptimizer, next to a DDP replica)",
                     prepare(hybrid), 4, r"rank=0, train_loss=([0-9.]+)"))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + tmp + os.pathsep + os.environ.get("PYTHONPATH", ""), HF_HUB_OFFLINE="1")
    lines = []
    for i, (name, text, nproc, pattern) in enumerate(runs):
        script = os.path.join(tmp, f"script_{i}.py")
        with open(script, "w") as f:
            f.write(text)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
               "--master-port", str(29580 + i), script]
        try:
            out = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
            losses = [float(x) for x in re.findall(pattern, out.stdout)]
            if out.returncode == 0 and len(losses) >= 2:
                lines.append(f"    {name}: ran, {len(losses)} training steps on rank 0, loss {losses[0]:.4f} -> {losses[-1]:.4f}")
            else:
                lines.append(f"    {name}: FAILED (exit {out.returncode})")
        except subprocess.TimeoutExpired:
            lines.append(f"    {name}: FAILED (timeout)")
    shutil.rmtree(tmp, ignore_errors=True)
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=None)
    ap.add_argument("--timeout", type=int, default=90, help="per test, seconds")
    ap.add_argument("--file-timeout", type=int, default=420, help="per test file, seconds (hard kill)")
    ap.add_argument("--only", default=None, help="sub-path under tests/ (default: everything but convergence/)")
    ap.add_argument("--example-only", action="store_true", help="only run the reference's example script")
    args = ap.parse_args()
    if args.example_only:
        print("\n".join(["== the reference's user scripts against this package (CPU ranks, stand-in data)"] + run_example(args.reference)))
        return 0
    src = os.path.join(args.reference, "tests")
    if not os.path.isdir(src):
        print(f"no reference tests at {src}")
        return 0
    tmp = tempfile.mkdtemp(prefix="pgb200_reftests_")
    dst = os.path.join(tmp, "reftests")
    shutil.copytree(src, dst)
    for dirpath, _dirs, files in os.walk(dst):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(dirpath, f)
                text = open(p).read()
                new = re.sub(r"\bpipegoose\b", "pipegoose_b200", text)
                if new != text:
                    open(p, "w").write(new)
    # the stand-ins must also be active inside the ranks a test spawns (they import the test module, not conftest.py):
    # one module next to the tests, imported by conftest.py and by every test file
    with open(os.path.join(tmp, "pgb200_hub_stand_in.py"), "w") as f:
        f.write(HUB_STAND_IN)
    with open(os.path.join(dst, "conftest.py"), "w") as f:
        f.write("import pgb200_hub_stand_in  # noqa: F401\n")
    for dirpath, _dirs, names in os.walk(dst):
        for f in names:
            if f.startswith("test_") and f.endswith(".py"):
                p = os.path.join(dirpath, f)
                text = open(p).read()
                open(p, "w").write("import pgb200_hub_stand_in  # noqa: F401\n" + text)
    # one pytest process PER TEST FILE with a hard limit: a rank blocked in a collective or a recv cannot stall the probe
    # (pytest-timeout cannot interrupt a main thread that is blocked in C)
    files = []
    for dirpath, _dirs, names in os.walk(os.path.join(dst, args.only) if args.only and os.path.isdir(os.path.join(dst, args.only)) else dst):
        for f in sorted(names):
            if f.startswith("test_") and f.endswith(".py") and "/convergence" not in dirpath:
                files.append(os.path.relpath(os.path.join(dirpath, f), tmp))
    if args.only and args.only.endswith(".py"):
        files = [os.path.join("reftests", args.only)]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + tmp + os.pathsep + os.environ.get("PYTHONPATH", ""), HF_HUB_OFFLINE="1")
    # RPC agents are opt-in here (enable_rpc=True / PIPEGOOSE_B200_ENABLE_RPC=1; nothing in the library needs them); the
    # reference starts them whenever there is more than one rank, and its RPC test relies on that
    env_rpc = dict(env, PIPEGOOSE_B200_ENABLE_RPC="1")
    results = {}
    import signal

    for t in sorted(files):
        cmd = [sys.executable, "-m", "pytest", t, "-q", "-p", "no:cacheprovider", "--timeout", str(args.timeout),
               "-rA", "--no-header", "-W", "ignore"]
        proc = subprocess.Popen(cmd, cwd=tmp, env=env_rpc if "test_rpc" in t else env, stdout=subprocess.PIPE,
                                stderr=subprocess.STDOUT, text=True, start_new_session=True)
        try:
            out, _ = proc.communicate(timeout=args.file_timeout)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGKILL)      # the pytest process and the ranks it started (its own session)
            out, _ = proc.communicate()
            results[t.replace("reftests/", "") + "::<file did not finish in %d s>" % args.file_timeout] = "FAILED"
        for line in (out or "").splitlines():
            m = re.match(r"^(PASSED|FAILED|ERROR|SKIPPED)\s+(\S+)", line)
            if m:
                results[m.group(2).replace("reftests/", "")] = m.group(1)
        print(f"[{t}] done", file=sys.stderr, flush=True)
    # a failure nobody expects gets ONE more chance alone with a doubled limit (16-rank tests exceed a per-test limit on a
    # box that is still reaping the ranks of a killed file); what passes then is reported as "passed on retry"
    retried = set()

    def expected(test_id):
        return any(re.search(pat, test_id) for pat, _ in KNOWN)

    for tid in [t for t, r in results.items() if r in ("FAILED", "ERROR") and not expected(t) and "::<file" not in t]:
        cmd = [sys.executable, "-m", "pytest", "reftests/" + tid, "-q", "-p", "no:cacheprovider", "--timeout", str(2 * args.timeout),
               "-rA", "--no-header", "-W", "ignore"]
        proc = subprocess.Popen(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
        try:
            out, _ = proc.communicate(timeout=args.file_timeout)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGKILL)
            out, _ = proc.communicate()
        if re.search(r"^PASSED\s+\S+", out or "", re.M):
            results[tid] = "PASSED"
            retried.add(tid)
    shutil.rmtree(tmp, ignore_errors=True)

    def reason(test_id):
        for pat, why in KNOWN:
            if re.search(pat, test_id):
                return why
        return None

    passed = sorted(t for t, r in results.items() if r == "PASSED")
    skipped = sorted(t for t, r in results.items() if r == "SKIPPED")
    failed = sorted(t for t, r in results.items() if r in ("FAILED", "ERROR"))
    explained = [(t, reason(t)) for t in failed if reason(t)]
    unexplained = [t for t in failed if not reason(t)]
    lines = ["The reference's own tests (tests/, convergence scripts excluded) run UNMODIFIED against pipegoose_b200",
             "(imports rewritten, hub replaced by offline stand-ins; CPU / gloo; tools/run_reference_tests.py)", "",
             f"passed {len(passed)}   failed {len(failed)} (explained {len(explained)}, unexplained {len(unexplained)})   "
             f"skipped by their own markers {len(skipped)}", ""]
    by_reason = {}
    for t, why in explained:
        by_reason.setdefault(why, []).append(t)
    lines.append("== failing, with the reason")
    for why, ts in by_reason.items():
        lines.append(f"* {why}")
        lines += [f"    {t}" for t in ts]
    if unexplained:
        lines.append("== failing, UNEXPLAINED")
        lines += [f"    {t}" for t in unexplained]
    lines.append("")
    lines.append("== passing")
    lines += [f"    {t}" + ("   (on retry, alone)" if t in retried else "") for t in passed]
    if not args.only:
        lines += ["", "== the reference's user scripts against this package (CPU ranks, one epoch, stand-in data)"]
        lines += run_example(args.reference)
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    return 1 if unexplained else 0


if __name__ == "__main__":
    sys.exit(main())
