"""Summarise `ncu --set full` reports into profiles/*.txt (run here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/ncu_gemm_pair.ncu-rep profiles/ncu_gemm_pair.txt ["note"]

Writes the roofline-relevant raw metrics plus the top stall reasons from the source page (needs -lineinfo)."""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(io.StringIO(ncu([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}", f"# {note}" if note else "#"]
    for r in rows[2:]:
        kv = dict(zip(hdr, zip(units, r)))
        lines.append(f"kernel: {kv.get('Kernel Name', ('', '?'))[1]}")
        for m in RAW:
            if m in kv:
                lines.append(f"  {m:75s} {kv[m][1]:>16s} {kv[m][0]}")
        stalls = sorted(((float(v[1]), k) for k, v in kv.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")
                         and v[1] not in ("", "n/a")), reverse=True)
        if stalls:
            lines.append("  top warp stall reasons (warps stalled per issue slot):")
            for val, k in stalls[:6]:
                lines.append(f"    {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {val:8.2f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
