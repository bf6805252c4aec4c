#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` raw CSV of tools/profile_step.py <slots>:
DRAM read+write bytes per slot for each captured stage kernel."""
import csv
import json
import sys

raw, slots, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
names = {"stft1920": "stft_db", "stft_db_kernel": "stft_db", "sync_scores": "costas_sync", "select_candidates": "select",
         "llr_kernel": "llr", "decode_kernel": "decode"}


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * f[unit]


res = {}
n_decode = 0
for r in data:
    k = next((v for n, v in names.items() if n in r[ix["Kernel Name"]]), None)
    if not k:
        continue
    if k == "decode":                # BP + CRC (captures made before llr_kernel existed: the first decode_kernel launch
        k = "ldpc_crc" if ("llr" in res or n_decode > 0) else "llr"      # of a step extracted the LLRs)
        n_decode += 1
        if k in res:
            continue
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    res[k] = {"kernel": r[ix["Kernel Name"]][:60], "slots_in_launch": slots, "dram_bytes_read": rd, "dram_bytes_write": wr,
              "dram_bytes_per_slot": (rd + wr) / slots,
              "fp64_pipe_pct": float(r[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
              "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
              "lsu_pipe_pct": float(r[ix["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]]),
              "dram_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
              "duration_us_under_ncu": float(r[ix["gpu__time_duration.sum"]]) * (1e3 if units[ix["gpu__time_duration.sum"]] == "ms" else 1)}
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1))
