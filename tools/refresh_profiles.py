#!/usr/bin/env python
"""Turn the gpurun_out/<tag>_* files written by tools/gpu_round_artifacts.sh into the tracked files under profiles/.
    python tools/refresh_profiles.py r43 [round-prefix, default r1]
Needs `ncu` on PATH (reads the .ncu-rep written on the GPU box)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json_line(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in " + path)


# ---- bench lines
bench = last_json_line(os.path.join(G, tag + "_bench_default.log"))
json.dump(bench, open(os.path.join(P, rnd + "_bench_default.json"), "w"), indent=1)
ref_log = os.path.join(G, tag + "_bench_reference.log")
if os.path.exists(ref_log):
    json.dump(last_json_line(ref_log), open(os.path.join(P, rnd + "_bench_reference_arm.json"), "w"), indent=1)
ldpc_log = os.path.join(G, tag + "_ldpc_c5.log")
if os.path.exists(ldpc_log):
    json.dump(last_json_line(ldpc_log), open(os.path.join(P, rnd + "_ldpc_c5_1e6.json"), "w"), indent=1)

# ---- launch list
src = os.path.join(G, tag + "_launches.csv")
shutil.copy(src, os.path.join(P, rnd + "_ncu_launches.csv"))
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ft8b::", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
plain = last_json_line(os.path.join(G, tag + "_bench_plain.log"))
with open(os.path.join(P, rnd + "_ncu_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stft|sync_scores|select_candidates|llr_kernel|decode_kernel -c 400\n")
    f.write("    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu      (%d slots; first 400 launches)\n" % plain["config"]["slots_per_gpu"])
    f.write("per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `stages[*].share`\n\n")
    f.write("%-40s %8s %12s %8s\n" % ("kernel", "launches", "total ms", "share"))
    for k, a in agg.items():
        f.write("%-40s %8d %12.3f %7.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))
    f.write("\nbench.py (CUDA events, same command without ncu) shares: " +
            ", ".join("%s %.1f%%" % (k, 100 * v["share"]) for k, v in plain["stages"].items()) + "\n")

# ---- full capture
rep = os.path.join(G, tag + "_prof.ncu-rep")
raw = os.path.join(G, tag + "_raw.csv")
subprocess.run("ncu -i %s --page raw --csv > %s 2>/dev/null" % (rep, raw), shell=True, check=True)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), raw, "256", os.path.join(P, "ncu_traffic.json")],
               check=True, stdout=subprocess.DEVNULL)
srcp = os.path.join(G, tag + "_src_all.csv")
subprocess.run("ncu -i %s --page source --csv > %s 2>/dev/null" % (rep, srcp), shell=True, check=True)
out = open(os.path.join(P, rnd + "_ncu_full_256slots.txt"), "w")
out.write("ncu --set full --clock-control none --import-source on -k regex:stft1920|sync_scores|llr_kernel|decode_kernel|select_cand -s 5 -c 5 python tools/profile_step.py 256 2\n")
out.write("(one decode step over 256 slots: stft1920 -> sync_scores -> select -> llr_kernel -> decode_kernel[BP])\n\n")
out.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_metrics.py"), raw], capture_output=True, text=True).stdout)
names = ["stft1920", "sync_scores", "select_candidates", "llr_kernel", "decode_kernel (BP launch)"]
for sec, name in enumerate(names):
    out.write("\n== SASS mix / stall reasons: %s\n" % name)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), srcp, str(2 * sec)], capture_output=True, text=True)
    out.write("\n".join(r.stdout.splitlines()[:16]) + "\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phase_stalls.py"), srcp, str(2 * sec)], capture_output=True, text=True)
    out.write(r.stdout.splitlines()[0] + "\n" if r.stdout else "")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_smem_conflicts.py"), srcp, str(2 * sec)], capture_output=True, text=True)
    out.write("\n".join(r.stdout.splitlines()[:6]) + "\n")
out.close()
print("profiles refreshed from", tag, "->", rnd)
print(json.dumps({k: bench[k] for k in ("value", "ms_per_step", "e2e", "roofline")}, indent=1)[:1500])
