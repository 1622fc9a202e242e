"""One profiled training step of the flagship model (bloom-560m, 8x1024 tokens, 1 GPU).

Run under ``ncu --profile-from-start off`` (the step is bracketed by cudaProfilerStart/Stop):

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/step_launches.csv python tools/step_profile.py
    python tools/step_profile.py --aggregate gpurun_out/step_launches.csv profiles/step_breakdown.json

Without ncu it prints device-event timings of the fwd / bwd / optimizer phases.
"""
import csv
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def aggregate(csv_path, out_path):
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns, grid, block in rows:
        short = name.split("(")[0]
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    table = sorted(([k, v[0], v[1] / 1e3, v[1] / total] for k, v in agg.items()), key=lambda t: -t[2])
    out = {"total_us": total / 1e3, "launches": len(rows),
           "kernels": [{"name": k, "count": c, "us": round(us, 1), "frac": round(fr, 4)} for k, c, us, fr in table]}
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
    for k in out["kernels"][:25]:
        print(f"{k['us']:10.1f} us  {k['frac']*100:5.1f}%  x{k['count']:<4d} {k['name'][:100]}")
    print(f"total {out['total_us']/1e3:.2f} ms over {out['launches']} launches")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--aggregate":
        aggregate(sys.argv[2], sys.argv[3])
        return
    import torch

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.optim.fused_adam import FusedAdam

    model_name = os.environ.get("PG_MODEL", "bloom_560m")
    B, S = int(os.environ.get("PG_B", 8)), int(os.environ.get("PG_S", 1024))
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = getattr(BloomConfig, model_name)()
    model = BloomForCausalLM(cfg).to(torch.bfloat16).to(dev)
    optim = FusedAdam(model.parameters(), lr=1e-4)
    ids = torch.randint(0, cfg.vocab_size, (B, S), device=dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        if timed:
            ev[0].record()
        loss = model(ids, labels=ids).loss
        if timed:
            ev[1].record()
        optim.zero_grad()
        loss.backward()
        if timed:
            ev[2].record()
        optim.step()
        if timed:
            ev[3].record()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    import time

    # host-side cost of enqueueing one step (GPU queue empty, no sync inside): if this is close to the device
    # time of a step, the step is launch-bound
    cpu = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        cpu.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(timed=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"fwd_ms": ev[0].elapsed_time(ev[1]), "bwd_ms": ev[1].elapsed_time(ev[2]),
                      "optim_ms": ev[2].elapsed_time(ev[3]), "cpu_enqueue_ms": cpu}))
    if os.environ.get("PG_CPROFILE"):
        import cProfile
        import pstats

        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
