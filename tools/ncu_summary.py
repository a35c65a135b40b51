#!/usr/bin/env python
"""Condense an ncu report (`ncu --set full`) into the markdown kept under profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/rN_<kernel>_ncu.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        row = dict(zip(hdr, zip(units, r)))
        print(f"## `{row['Kernel Name'][1][:120]}`\n")
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in row:
                print(f"| {name} (`{k}`) | {row[k][1]} {row[k][0]} |")
        st = sorted(((float(v[1]), k[len(STALLS):-len('_per_issue_active.ratio')]) for k, v in row.items()
                     if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio")), reverse=True)
        print("\nWarp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{n} {v:.2f}" for v, n in st[:7]) + "\n")
    src = run(["ncu", "-i", rep, "--page", "source", "--csv"])
    blocks = src.split('"Kernel Name"')
    for b in blocks[1:2]:
        lines = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
        h = lines[1]
        ix = {k: i for i, k in enumerate(h)}
        data = [l for l in lines[2:] if len(l) == len(h)]
        tot = sum(int(l[ix["# Samples"]] or 0) for l in data) or 1
        print(f"### Hottest SASS instructions by stall samples (first captured launch; {tot} samples)\n")
        print("| samples | share | SASS | dominant stalls |\n|---|---|---|---|")
        for l in sorted(data, key=lambda l: -int(l[ix["# Samples"]] or 0))[:top]:
            s = int(l[ix["# Samples"]] or 0)
            stl = sorted(((int(l[ix[k]] or 0), k[6:]) for k in h if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
            print(f"| {s} | {100 * s / tot:.1f}% | `{l[ix['Source']][:80]}` | " + ", ".join(f"{n} {v}" for v, n in stl if v) + " |")


if __name__ == "__main__":
    main()
