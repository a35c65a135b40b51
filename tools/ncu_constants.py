#!/usr/bin/env python
"""profiles/ncu_constants.json from an `ncu --set full` report: the profiler-only counters bench.py quotes beside its live
timings (DRAM bytes per launch, pipe utilisations), each tagged with the capture file and the sha256 of the kernel's source
file AT CAPTURE TIME -- bench.py drops an entry as soon as that source changes, so a stale counter can never be printed.

  python tools/ncu_constants.py gpurun_out/prof_kernels.ncu-rep profiles/r2_kernels_ncu.md
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# bench stage name -> (substring of the kernel name, source file)
KERNELS = {
    "encode_sample": ("encode_sample_tc_kernel", "cppf_b200/csrc/encode_tc.cu"),
    "vote": ("vote_private_kernel", "cppf_b200/csrc/vote_private.cu"),
    "backvote": ("backvote_bins_kernel", "cppf_b200/csrc/vote_private.cu"),
    "stats": ("survivor_stats_kernel", "cppf_b200/csrc/vote_private.cu"),
    "point_encoder": ("point_encode_tc_kernel", "cppf_b200/csrc/point_encoder.cu"),
}
METRICS = {
    "gpu__time_duration.sum": "duration_ms_under_ncu",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct_of_peak",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
}


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def main():
    rep, capture = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"capture": capture, "report": os.path.basename(rep), "kernels": {}}
    for r in rows[2:]:
        row = dict(zip(hdr, zip(units, r)))
        name = row["Kernel Name"][1]
        for stage, (needle, src) in KERNELS.items():
            if needle not in name or stage in out["kernels"]:
                continue
            ent = {"kernel": name[:100], "capture": capture, "source": src,
                   "source_sha256": hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()}
            if "dram__bytes_read.sum" in row and "dram__bytes_write.sum" in row:
                ent["dram_bytes_per_launch"] = to_bytes(row["dram__bytes_read.sum"][1], row["dram__bytes_read.sum"][0]) + \
                    to_bytes(row["dram__bytes_write.sum"][1], row["dram__bytes_write.sum"][0])
            for m, key in METRICS.items():
                if m in row:
                    try:
                        ent[key] = float(row[m][1].replace(",", ""))
                    except ValueError:
                        pass
            out["kernels"][stage] = ent
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, list(out["kernels"]))


if __name__ == "__main__":
    main()
