"""Summarise an .ncu-rep (read here, without a GPU) into a small text file for profiles/:
    python tools/summarize_ncu.py gpurun_out/prof_fit.ncu-rep profiles/r01_fit_kernel.txt
Headline raw metrics (duration, DRAM bytes, pipe utilisation, registers), the SASS opcode mix and the top stall sites
from the source page (needs -lineinfo / --import-source on at capture time)."""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
       "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    head, units = rows[0], rows[1]
    for rec in rows[2:]:
        name = rec[head.index("Kernel Name")]
        lines.append(f"== kernel: {name}")
        for m in RAW:
            if m in head:
                lines.append(f"  {m:70s} {rec[head.index(m)]:>18s} {units[head.index(m)]}")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h, data = src[hi], [r for r in src[hi + 1:] if len(r) >= len(src[hi])]
    ci = {n: i for i, n in enumerate(h)}
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
        op = m.group(2) if m else "?"
        base = op if op.startswith("MUFU") else op.split(".")[0]
        ops[base] += int(r[ci["Instructions Executed"]])
        samp[base] += int(r[ci["# Samples"]])
        for c in stall_cols:
            stalls[c] += int(r[ci[c]])
    ti, ts = sum(ops.values()), max(1, sum(samp.values()))
    lines.append(f"\n== SASS opcode mix (warp instructions executed: {ti}; stall samples: {ts})")
    for k, v in ops.most_common(18):
        lines.append(f"  {k:14s} {100 * v / ti:5.1f} % of instructions   {100 * samp[k] / ts:5.1f} % of samples")
    for tag in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS"):
        lines.append(f"  {tag:14s} {ops.get(tag, 0)} executed")
    lines.append("\n== warp stall reasons (samples)")
    for k, v in stalls.most_common(8):
        lines.append(f"  {k:28s} {v:8d}  {100 * v / ts:5.1f} %")
    lines.append("\n== top stall sites")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:14]:
        why = {c: int(r[ci[c]]) for c in stall_cols if int(r[ci[c]]) > 0.25 * max(1, int(r[ci["# Samples"]]))}
        lines.append(f"  {int(r[ci['# Samples']]):6d}  {r[ci['Source']].strip()[:70]:70s} {why}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
