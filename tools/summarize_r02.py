"""Copies the judged artefacts of tools/gpu_profile_r02.sh / gpu_scale_r02.sh from gpurun_out/ into profiles/ (round 2):
bench lines, the ncu launch list with a per-kernel share table, the kernels' metric / hot-line texts, and
profiles/traffic.json (DRAM bytes per uncompressed byte, read by bench.py's roofline.traffic)."""
import csv
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
import sys
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
for f in glob.glob(os.path.join(G, TAG + "_*")):
    b = os.path.basename(f)
    if b.endswith((".json", ".txt", "_launches.csv")) and os.path.getsize(f) > 0:
        shutil.copy(f, os.path.join(P, b))
lp = os.path.join(G, f"{TAG}_launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows:
        if len(r) != len(hdr) or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1
    with open(os.path.join(P, f"{TAG}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list summary ({TAG}): `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |\n")
# traffic: uncompressed bytes covered by the captured launch (tools/gpu_profile_r02.sh)
LAUNCH_GIB = {"lz4_copy_kernel": 16, "lz4_region_kernel": 16, "lz4_parse_kernel": 16, "lz4_parse_wide_kernel": 0.5, "zstd_frames_lane_kernel": 4}
tp = os.path.join(P, "traffic.json")
traffic = json.load(open(tp)) if os.path.exists(tp) else {}
for k, gib in LAUNCH_GIB.items():
    mp = os.path.join(G, f"{TAG}_{k}_metrics.txt")
    if not os.path.exists(mp):
        continue
    txt = open(mp).read()
    def num(name):
        m = re.search(name + r"\s+([\d.,]+)\s+(\w+)", txt)
        return float(m.group(1).replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m.group(2)]
    try:
        traffic[k] = {"dram_bytes_per_uncompressed_byte": (num(r"dram__bytes_read\.sum") + num(r"dram__bytes_write\.sum")) / (gib * 2 ** 30),
                      "source": f"profiles/{TAG}_{k}_metrics.txt"}
    except Exception as e:      # noqa: BLE001
        print("no traffic for", k, e)
json.dump(traffic, open(tp, "w"), indent=1)
print("profiles/ updated")
